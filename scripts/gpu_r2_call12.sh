# Round 2, call 12 (N=2): halo push published after the first tile's barrier (default) vs in the prologue (variants/libpd_early.so)
mkdir -p gpurun_out
T=r2c12; N=${N:-2}; W=${W:-grid70}
timeout 500 python -m pytest tests/test_gpu_dist.py tests/test_gpu_linear.py -m gpu -q -s 2>&1 | grep -E "passed|failed|PCG-Jacobi|CG \+ IC|vs the reference" | cut -c1-250 | tee gpurun_out/${T}_pytest.log
for rep in 1 2; do for v in default early; do
  if [ $v = default ]; then unset PD_B200_LIB; else export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$rep bench.py --gpus $N --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_${v}_n${N}_${W}_$rep.json 2> gpurun_out/${T}_${v}_n${N}_${W}_$rep.err
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$rep scripts/dist_perf.py $W 3 2>&1 | grep -v "^\*\|OMP_NUM\|NCCL version\|\[bench\]" | tee gpurun_out/${T}_${v}_dist_perf_n${N}_$rep.txt
done; done
unset PD_B200_LIB
python - <<PY
import json
for rep in (1, 2):
    for v in ("default", "early"):
        try:
            d=[json.loads(l) for l in open(f"gpurun_out/${T}_{v}_n${N}_${W}_{rep}.json") if l.startswith("{")][-1]; r=d["roofline"]; m=d["run"]["multi_gpu"]
            print(v, "rep", rep, "$W N=$N ms/step %.3f value %.0f halo_ok %s bit_identical %s"%(d["ms_per_step"], d["value"], d["run"]["halo_ok"], (d.get("parity") or {}).get("bit_identical_to_n1")))
        except Exception as e: print(v, rep, "failed", e)
PY
