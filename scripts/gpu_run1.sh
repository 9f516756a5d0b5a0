set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r1_pytest.log; cat gpurun_out/r1_pytest.log
timeout 600 python tests/golden/make_reference_golden.py gpurun_out/reference_b200.npz 2>&1 | tail -10
timeout 600 python bench.py --workload grid55 --steps 5 --warmup 3 > gpurun_out/r1_bench_grid55.json 2> gpurun_out/r1_bench_grid55.err; cat gpurun_out/r1_bench_grid55.json; tail -5 gpurun_out/r1_bench_grid55.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r1_bench_grid139.json 2> gpurun_out/r1_bench_grid139.err; cat gpurun_out/r1_bench_grid139.json; tail -5 gpurun_out/r1_bench_grid139.err
timeout 900 python bench.py --impl reference --workload grid55 --steps 3 --warmup 3 > gpurun_out/r1_ref_grid55.json 2> gpurun_out/r1_ref_grid55.err; cat gpurun_out/r1_ref_grid55.json; tail -5 gpurun_out/r1_ref_grid55.err
