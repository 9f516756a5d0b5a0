# Round 2, call 4 (N=1): ncu of the per-body kernel on batch64 (why 14 us per iteration?)
mkdir -p gpurun_out
T=r2c4
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_body_step -s 3 -c 1 -o gpurun_out/${T}_k_body_batch64 python bench.py --workload batch64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log
