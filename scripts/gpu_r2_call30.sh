# Round 2, call 30 (N=1): block size of the cooperative solve kernels (256 = default, 512, 1024 threads; at most 3 blocks per SM) on config 3,
# the solver tests through the default first
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -2
for v in default st512 st1024; do
  if [ $v = default ]; then unset PD_B200_LIB; else export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so; fi
  for rep in 1 2; do
  timeout 300 python bench.py --workload grid55-pcg --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$v rep $rep grid55-pcg ms/step %.3f value %.0f frac %.3f parity %s'%(d['ms_per_step'], d['value'], d['roofline']['frac'], (d.get('parity') or {}).get('rel_err')))"
  done
done
unset PD_B200_LIB
timeout 300 python bench.py --workload grid55-pcg-tol --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('default grid55-pcg-tol ms/step %.3f value %.0f frac %.3f parity %s'%(d['ms_per_step'], d['value'], d['roofline']['frac'], (d.get('parity') or {}).get('rel_err')))"
