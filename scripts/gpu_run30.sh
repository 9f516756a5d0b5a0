# N=8 with the halo push at the start of the local kernel (push-list order, spread over the CTAs)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29546 scripts/dist_perf.py grid139 3 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/r30_dist_perf_n8.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r30_n8_grid139.json 2> gpurun_out/r30_n8_grid139.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r30_n8_grid139.json") if l.startswith("{")][-1]; r=d["roofline"]
print("grid139 N=8 ms/step %.3f value %.0f e2e %.0f local %.1f us vertex %.1f us halo_ok %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3, d["config"]["halo_ok"]), d["clocks"]["sm_mhz"])
PY
