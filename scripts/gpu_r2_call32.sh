# Round 2, call 32 (N=1): programmatic dependent launch off (0) / late trigger (2) on the small and medium workloads
mkdir -p gpurun_out
for w in armadillo grid55 grid70 grid139; do for rep in 1 2; do for v in 0 2; do
  PD_PDL=$v timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-faithful 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('PD_PDL=$v rep $rep $w ms/step %.4f e2e %.4f'%(d['ms_per_step'], d['e2e']['ms_per_step']))"
done; done; done
