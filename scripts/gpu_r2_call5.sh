# Round 2, call 5 (N=1): full GPU suite after the body-kernel default / perf timers / test fixes; smoke; batch64 line
mkdir -p gpurun_out
T=r2c5
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/${T}_pytest.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/${T}_smoke.log
timeout 200 python bench.py --workload batch64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_batch64.json 2> gpurun_out/${T}_batch64.err; tail -2 gpurun_out/${T}_batch64.err; tail -c 1500 gpurun_out/${T}_batch64.json
