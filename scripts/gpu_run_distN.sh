mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/dist_bench_n$N.json 2> gpurun_out/dist_bench_n$N.err; tail -3 gpurun_out/dist_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/dist_bench_n$N.json") if l.startswith("{")][-1]; r=d["roofline"]
print("N=$N ms/step %.2f value %.0f e2e %.0f local %.1f us vertex %.1f us halo_ok %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3, d["config"]["halo_ok"]))
print(d["config"]["multi_gpu"])
PY
