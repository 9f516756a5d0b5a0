# Round 2, call 7 (N=1): nested-dissection Cholesky + warp-per-row solves, PCG with two grid syncs per CG iteration; config 3 bench lines
mkdir -p gpurun_out
T=r2c7
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${T}_pytest.log
timeout 400 python scripts/solver_bench.py > gpurun_out/${T}_solver_bench.jsonl 2> gpurun_out/${T}_solver_bench.err; tail -3 gpurun_out/${T}_solver_bench.err; cut -c1-700 gpurun_out/${T}_solver_bench.jsonl
for w in grid55-pcg grid55-pcg-tol; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_$w.json 2> gpurun_out/${T}_$w.err; tail -2 gpurun_out/${T}_$w.err; cut -c1-1500 gpurun_out/${T}_$w.json
done
timeout 300 python bench.py --impl reference --workload grid55-pcg-tol --steps 3 --warmup 3 > gpurun_out/${T}_ref_grid55-pcg-tol.json 2> gpurun_out/${T}_ref_grid55-pcg-tol.err; tail -2 gpurun_out/${T}_ref_grid55-pcg-tol.err; cut -c1-900 gpurun_out/${T}_ref_grid55-pcg-tol.json
