# Round 2, call 23 (N=1, grid139, same box): L2 prefetch of the staging lists (vstage) two tiles ahead (default) vs none (variants/libpd_nopf.so);
# per-phase clock profile of warp 0 for both
mkdir -p gpurun_out
T=${T:-r2c23}; W=${W:-grid139}
for v in default nopf; do
  if [ $v = default ]; then unset PD_B200_LIB; else export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so; fi
  echo "== $v"; timeout 300 python scripts/phase_profile.py $W 2>&1 | tail -8
done
for rep in 1 2 3; do for v in default nopf; do
  if [ $v = default ]; then unset PD_B200_LIB; else export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so; fi
  timeout 300 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-faithful > gpurun_out/${T}_${v}_$rep.json 2> gpurun_out/${T}_${v}_$rep.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${T}_${v}_$rep.json") if l.startswith("{")][-1]; r=d["roofline"]
    print("$v rep $rep $W ms/step %.3f local %.1f us vertex (alone) %.1f us"%(d["ms_per_step"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3), d["clocks"]["sm_mhz"])
except Exception as e: print("$v rep $rep failed", e)
PY
done; done
unset PD_B200_LIB
