# First GPU call of round 2 (one gpurun call, N=1): everything the last session of round 1 wrote without GPU time, then the
# usual validation.  gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh'
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2a_pytest.log          # no -x: see every new test
timeout 120 python -m pytest tests/test_gpu_zz_drag.py tests/test_pd_run.py -m gpu -q -s 2>&1 | grep -E "rel err|max_abs_diff|drag|passed|failed" | tee gpurun_out/r2a_drag.log
timeout 200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_grid139.json 2> gpurun_out/r2a_bench_grid139.err; tail -2 gpurun_out/r2a_bench_grid139.err
timeout 200 python scripts/solver_bench.py > gpurun_out/r2a_solver_bench.jsonl 2> gpurun_out/r2a_solver_bench.err; tail -3 gpurun_out/r2a_solver_bench.err; cat gpurun_out/r2a_solver_bench.jsonl
# memcheck of the paths written without GPU time (drag kernels, live mu edit, per-body kernel)
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/dbg_small.py 6 2>&1 | tail -4 | tee gpurun_out/r2a_memcheck.log
PD_BODY_KERNEL=1 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/dbg_small.py 6 2>&1 | tail -4 | tee gpurun_out/r2a_memcheck_body.log
