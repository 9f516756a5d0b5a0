set -x
mkdir -p gpurun_out
timeout 1200 python scripts/noise_floor.py > gpurun_out/noise_floor3.txt 2>&1; tail -60 gpurun_out/noise_floor3.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r3_pytest_all.log; cat gpurun_out/r3_pytest_all.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench_grid139.json 2> gpurun_out/r3_bench_grid139.err; cat gpurun_out/r3_bench_grid139.json; tail -5 gpurun_out/r3_bench_grid139.err
timeout 600 python bench.py --workload grid55 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench_grid55.json 2> gpurun_out/r3_bench_grid55.err; cat gpurun_out/r3_bench_grid55.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_local -s 12 -c 1 -f -o gpurun_out/r3_k_local_grid139 python scripts/profile_step.py grid139 2 10 > gpurun_out/prof3.log 2>&1
tail -3 gpurun_out/prof3.log
