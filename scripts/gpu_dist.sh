# Multi-GPU run on one box (gpurun --gpus N -- 'N=8 bash scripts/gpu_dist.sh'): distributed GPU tests (need 2 GPUs), the
# in-situ kernel times of every rank (scripts/dist_perf.py) and the torchrun bench line at N
mkdir -p gpurun_out
N=${N:-2}; W=${W:-grid139}
timeout 500 python -m pytest tests/test_gpu_dist.py -q 2>&1 | tail -4 | tee gpurun_out/dist_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 scripts/dist_perf.py $W 3 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/dist_perf_n$N.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --workload $W --steps 5 --warmup 3 > gpurun_out/dist_bench_n${N}_$W.json 2> gpurun_out/dist_bench_n${N}_$W.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/dist_bench_n${N}_$W.json") if l.startswith("{")][-1]; r=d["roofline"]
print("$W N=$N ms/step %.3f value %.0f e2e %.0f local %.1f us vertex %.1f us halo_ok %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3, d["run"]["halo_ok"]), d["clocks"]["sm_mhz"])
print(d["run"]["multi_gpu"])
PY
