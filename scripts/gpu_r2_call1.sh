# Round 2, first GPU call (N=1): the whole GPU suite, the prepared same-box A/Bs of round 1's unmeasured variants
# (PD_H_PLANES, PD_PHASEC_PRED, PD_BODY_KERNEL), the faithful mode's first driver-style timing, config 3's solvers.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r2_call1.sh'
mkdir -p gpurun_out
T=r2c1
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${T}_pytest.log; tail -5 gpurun_out/${T}_pytest.log
V="default planes pred" bash scripts/gpu_ab.sh 2>&1 | tee gpurun_out/${T}_ab.log
timeout 300 python bench.py --rot-mode 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_faithful_grid139.json 2> gpurun_out/${T}_faithful_grid139.err; tail -2 gpurun_out/${T}_faithful_grid139.err
for w in grid55 batch64; do
  timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_$w.json 2> gpurun_out/${T}_$w.err; tail -2 gpurun_out/${T}_$w.err
done
bash scripts/gpu_body_ab.sh 2>&1 | tee gpurun_out/${T}_body.log
timeout 300 python scripts/solver_bench.py > gpurun_out/${T}_solver_bench.jsonl 2> gpurun_out/${T}_solver_bench.err; tail -3 gpurun_out/${T}_solver_bench.err; cat gpurun_out/${T}_solver_bench.jsonl
for v in planes pred; do
  PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_local -s 210 -c 1 -o gpurun_out/${T}_${v}_k_local_grid139 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_${v}_ncu.log 2>&1; tail -2 gpurun_out/${T}_${v}_ncu.log
done
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/${T}_reference_grid139.json 2> gpurun_out/${T}_reference_grid139.err; tail -c 600 gpurun_out/${T}_reference_grid139.json
