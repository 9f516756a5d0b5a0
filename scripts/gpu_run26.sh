# A/B: vertex kernel at 6 vs 8 resident blocks per SM (40 vs 32 registers)
mkdir -p gpurun_out
for rep in 1 2; do for v in default vb8; do
  if [ $v = default ]; then unset PD_B200_LIB; else export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r26_grid139_${v}_$rep.json 2> gpurun_out/r26_grid139_${v}_$rep.err; tail -2 gpurun_out/r26_grid139_${v}_$rep.err
done; done
unset PD_B200_LIB
python - <<PY
import json
for rep in [1,2]:
  for v in ["default","vb8"]:
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r26_grid139_{v}_{rep}.json") if l.startswith("{")][-1]; r=d["roofline"]
        print(v, rep, "ms/step %.3f value %.0f local %.1f us vertex %.1f us"%(d["ms_per_step"], d["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3), d["clocks"]["sm_mhz"])
    except Exception as e: print(v,"failed",e)
PY
