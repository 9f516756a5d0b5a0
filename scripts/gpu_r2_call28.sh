# Round 2, call 28 (N=1, batch64, same box): per-body kernel with the slow rotation path on a compacted list (default) vs taken in place
# (variants/libpd_inplace.so = the previous commit); the GPU tests that run the per-body kernel first
mkdir -p gpurun_out
T=${T:-r2c28}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_collision.py tests/test_gpu_zz_drag.py -m gpu -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for rep in 1 2 3; do for v in default inplace; do
  if [ $v = default ]; then unset PD_B200_LIB; else export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so; fi
  timeout 300 python bench.py --workload batch64 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$v rep $rep batch64 ms/step %.4f value %.0f e2e %.0f finite %s'%(d['ms_per_step'], d['value'], d['e2e']['value'], d['run']['finite']))"
done; done
unset PD_B200_LIB
