# Round 2, call 17 (N=1): vertex kernel experiments on grid139, same box: c and c + md folded into the w lanes (default) vs read from cc
# (variants/libpd_nofold.so); blocks walking the vertices from the end (PD_VERTEX_REVERSE=1)
mkdir -p gpurun_out
T=${T:-r2c17}; W=${W:-grid139}
for rep in 1 2 3; do for v in default reverse nofold; do
  unset PD_B200_LIB PD_VERTEX_REVERSE
  [ $v = nofold ] && export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_nofold.so
  [ $v = reverse ] && export PD_VERTEX_REVERSE=1
  timeout 300 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-faithful > gpurun_out/${T}_${v}_$rep.json 2> gpurun_out/${T}_${v}_$rep.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${T}_${v}_$rep.json") if l.startswith("{")][-1]; r=d["roofline"]
    print("$v rep $rep $W ms/step %.3f local %.1f us vertex (alone) %.1f us"%(d["ms_per_step"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3), d["clocks"]["sm_mhz"])
except Exception as e: print("$v rep $rep failed", e)
PY
done; done
unset PD_B200_LIB PD_VERTEX_REVERSE
