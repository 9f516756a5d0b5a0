# PD_DIST_TRIM experiment (trimmed + re-packed boundary tiles, csrc/layout.hpp) on N GPUs of one box:
#   gpurun --gpus 2 --timeout 900 -- 'N=2 W=grid70 bash scripts/gpu_trim_ab.sh'      then N=8 W=grid139
# 1. the bit-identity tests of tests/test_gpu_dist.py with the trimmed layouts (one-process lock-step needs 1 GPU, the
#    two-process test 2); 2. bench at N with and without, same box; 3. in-situ kernel times with trimming.
mkdir -p gpurun_out
N=${N:-2}; W=${W:-grid70}
PD_DIST_TRIM=1 timeout 500 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/trim_pytest.log
for rep in 1 2; do for t in 0 1; do
  PD_DIST_TRIM=$t timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$t bench.py --gpus $N --workload $W --steps 5 --warmup 3 > gpurun_out/trim${t}_n${N}_${W}_$rep.json 2> gpurun_out/trim${t}_n${N}_${W}_$rep.err
done; done
PD_DIST_TRIM=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 scripts/dist_perf.py $W 3 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/trim1_dist_perf_n$N.txt
python - <<PY
import json
for rep in (1, 2):
    for t in (0, 1):
        try:
            d=[json.loads(l) for l in open(f"gpurun_out/trim{t}_n${N}_${W}_{rep}.json") if l.startswith("{")][-1]; r=d["roofline"]; m=d["run"]["multi_gpu"]
            print("trim", t, "rep", rep, "$W N=$N ms/step %.3f value %.0f local %.1f us vertex %.1f us halo_ok %s redundant %.3f"%(d["ms_per_step"], d["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3, d["run"]["halo_ok"], m["redundant_tet_fraction"]), max(m["tets_evaluated_per_rank"]))
        except Exception as e: print("trim", t, rep, "failed", e)
PY
