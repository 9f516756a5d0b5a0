# Round 2, call 26 (N=1, grid139, same box): phase-C groups with two lanes per vertex for the heaviest vertices of a tile (default) vs one
# lane per vertex everywhere (variants/libpd_lpv1.so = the previous commit); whole GPU suite through the new layout first
mkdir -p gpurun_out
T=${T:-r2c26}; W=${W:-grid139}
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
for rep in 1 2 3; do for v in default lpv1; do
  if [ $v = default ]; then unset PD_B200_LIB; else export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so; fi
  timeout 300 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-faithful > gpurun_out/${T}_${v}_$rep.json 2> gpurun_out/${T}_${v}_$rep.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${T}_${v}_$rep.json") if l.startswith("{")][-1]; r=d["roofline"]
    print("$v rep $rep $W ms/step %.3f local %.1f us vertex (alone) %.1f us frac %.3f"%(d["ms_per_step"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3, r["frac"]), d["clocks"]["sm_mhz"])
except Exception as e: print("$v rep $rep failed", e)
PY
done; done
unset PD_B200_LIB
for w in grid55 armadillo; do for v in default lpv1; do
  if [ $v = default ]; then unset PD_B200_LIB; else export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so; fi
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-faithful 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$v $w ms/step %.4f'%d['ms_per_step'])"
done; done
unset PD_B200_LIB
