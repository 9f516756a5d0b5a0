# Round 2, call 9 (N=1): mesh-mesh collision pass (first GPU run), adapter with handleCollision, Cholesky with pipelined gathers
mkdir -p gpurun_out
T=r2c9
timeout 900 python -m pytest tests/test_gpu_collision.py -m gpu -q -s -x 2>&1 | grep -vE "^$" | tail -30 | tee gpurun_out/${T}_collision.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/${T}_pytest.log
timeout 400 python scripts/solver_bench.py > gpurun_out/${T}_solver_bench.jsonl 2> gpurun_out/${T}_solver_bench.err; tail -3 gpurun_out/${T}_solver_bench.err; cut -c1-420 gpurun_out/${T}_solver_bench.jsonl
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_collision.py -m gpu -q -x -k "pass_vs_oracle" 2>&1 | tail -5 | tee gpurun_out/${T}_memcheck.log
