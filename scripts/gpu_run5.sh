set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r5_pytest.log; cat gpurun_out/r5_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r5_bench_grid139.json 2> gpurun_out/r5_bench_grid139.err; cat gpurun_out/r5_bench_grid139.json; tail -5 gpurun_out/r5_bench_grid139.err
timeout 600 python bench.py --workload grid55 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r5_bench_grid55.json 2> gpurun_out/r5_bench_grid55.err; cat gpurun_out/r5_bench_grid55.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r5_ref_grid139.json 2> gpurun_out/r5_ref_grid139.err; cat gpurun_out/r5_ref_grid139.json; tail -5 gpurun_out/r5_ref_grid139.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r5_launches_grid139.csv python scripts/profile_step.py grid139 2 10 > gpurun_out/prof5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_local -s 12 -c 1 -f -o gpurun_out/r5_k_local_grid139 python scripts/profile_step.py grid139 2 10 >> gpurun_out/prof5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vertex_jacobi -s 12 -c 1 -f -o gpurun_out/r5_k_vertex_grid139 python scripts/profile_step.py grid139 2 10 >> gpurun_out/prof5.log 2>&1
tail -3 gpurun_out/prof5.log
ls -la gpurun_out
