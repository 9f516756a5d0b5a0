set -x
mkdir -p gpurun_out
TAG=${TAG:-r6}
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_pytest.log
[ -x oracle/_ref/adapter_test ] && (timeout 120 oracle/_ref/adapter_test 6 3 2>&1 | tee gpurun_out/${TAG}_adapter.log)
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_grid139.json 2> gpurun_out/${TAG}_bench_grid139.err; cat gpurun_out/${TAG}_bench_grid139.json; tail -5 gpurun_out/${TAG}_bench_grid139.err
timeout 600 python bench.py --workload grid55 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_grid55.json 2> gpurun_out/${TAG}_bench_grid55.err; cat gpurun_out/${TAG}_bench_grid55.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_local -s 12 -c 1 -f -o gpurun_out/${TAG}_k_local_grid139 python scripts/profile_step.py grid139 2 10 > gpurun_out/prof_${TAG}.log 2>&1
tail -3 gpurun_out/prof_${TAG}.log
