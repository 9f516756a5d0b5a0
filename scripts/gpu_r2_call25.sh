# Round 2, call 25 (N=1, grid139, same box): position gather issued by the upper half of the CTA (default) vs by every thread for its own
# slot (variants/libpd_allgather.so = the previous commit); phase profile of warp 0; parity tests through the new kernel
mkdir -p gpurun_out
T=${T:-r2c25}; W=${W:-grid139}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -2
echo "== default"; timeout 300 python scripts/phase_profile.py $W 2>&1 | tail -8
for rep in 1 2 3; do for v in default allgather; do
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
for w in grid55 armadillo; do for v in default allgather; do
  if [ $v = default ]; then unset PD_B200_LIB; else export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so; fi
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-faithful 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$v $w ms/step %.4f'%d['ms_per_step'])"
done; done
unset PD_B200_LIB
