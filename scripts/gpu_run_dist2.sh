mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
nvidia-smi topo -m | head -8
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -8 | tee gpurun_out/dist2_pytest.log
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/dist_bench_n$N.json 2> gpurun_out/dist_bench_n$N.err; tail -3 gpurun_out/dist_bench_n$N.err; cat gpurun_out/dist_bench_n$N.json | cut -c1-1500
