# Round 2, call 29 (N=1): grid size of the cooperative PCG kernel on config 3 (grid55-pcg, 50 fixed inner iterations)
mkdir -p gpurun_out
for k in 0 1 2 3; do
  if [ $k = 0 ]; then unset PD_SOLVE_BLOCKS_PER_SM; else export PD_SOLVE_BLOCKS_PER_SM=$k; fi
  for rep in 1 2; do
  timeout 300 python bench.py --workload grid55-pcg --steps 5 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('blocks/SM $k rep $rep grid55-pcg ms/step %.3f value %.0f frac %.3f'%(d['ms_per_step'], d['value'], d['roofline']['frac']))"
  done
done
unset PD_SOLVE_BLOCKS_PER_SM
