# Round 2, call 8 (N=1): Cholesky solves with 128-bit tagged entries, PCG with one sync per reduction
mkdir -p gpurun_out
T=r2c8
timeout 600 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${T}_pytest.log
timeout 400 python scripts/solver_bench.py > gpurun_out/${T}_solver_bench.jsonl 2> gpurun_out/${T}_solver_bench.err; tail -3 gpurun_out/${T}_solver_bench.err; cut -c1-520 gpurun_out/${T}_solver_bench.jsonl
