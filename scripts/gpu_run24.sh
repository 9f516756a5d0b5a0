# A/B on one box: heaviest vertex groups on the highest warp ids (default) vs group w on warp w
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r24_pytest.log; cat gpurun_out/r24_pytest.log
for rep in 1 2; do for v in default rev0; do
  if [ $v = default ]; then unset PD_B200_LIB; else export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r24_grid139_${v}_$rep.json 2> gpurun_out/r24_grid139_${v}_$rep.err; tail -2 gpurun_out/r24_grid139_${v}_$rep.err
done; done
unset PD_B200_LIB
python - <<PY
import json
for rep in [1,2]:
  for v in ["default","rev0"]:
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r24_grid139_{v}_{rep}.json") if l.startswith("{")][-1]; r=d["roofline"]
        print(v, rep, "ms/step %.3f value %.0f e2e %.0f local %.1f us vertex %.1f us frac %.3f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3, r["frac"]), d["clocks"]["sm_mhz"])
    except Exception as e: print(v,"failed",e)
PY
