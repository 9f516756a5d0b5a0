# Round 2, call 13 (N=8): grid139 strong scaling with the trimmed boundary tiles and the deferred publish; in-situ kernel times
mkdir -p gpurun_out
T=r2c13; N=${N:-8}; W=${W:-grid139}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_n${N}_${W}.json 2> gpurun_out/${T}_n${N}_${W}.err
tail -c 2500 gpurun_out/${T}_n${N}_${W}.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 scripts/dist_perf.py $W 3 2>&1 | grep -v "^\*\|OMP_NUM\|NCCL version\|\[bench\]" | tee gpurun_out/${T}_dist_perf_n${N}.txt
