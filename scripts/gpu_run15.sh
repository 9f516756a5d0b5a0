# packed-FP32 phase B: parity suite + bench of the default build and the 2-rows-per-trip phase C variant + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r15_pytest.log; cat gpurun_out/r15_pytest.log
for v in default c2; do
  if [ $v = default ]; then unset PD_B200_LIB; else export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r15_grid139_$v.json 2> gpurun_out/r15_grid139_$v.err; tail -2 gpurun_out/r15_grid139_$v.err
done
unset PD_B200_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local -s 12 -c 1 -f -o gpurun_out/r15_k_local_grid139 python scripts/profile_step.py grid139 2 10 > gpurun_out/r15_prof.log 2>&1
python - <<PY
import json
for v in ["default","c2"]:
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r15_grid139_{v}.json") if l.startswith("{")][-1]; r=d["roofline"]
        print(v, "ms/step %.3f value %.0f e2e %.0f local %.1f us vertex %.1f us frac %.3f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3, r["frac"]), d["clocks"]["sm_mhz"])
    except Exception as e: print(v,"failed",e)
PY
