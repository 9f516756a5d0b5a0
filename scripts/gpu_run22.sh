# N=2 validation of the v7 local kernel: distributed GPU tests + torchrun bench on grid139 and grid70
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r22_pytest.log
for w in grid139 grid70; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload $w --steps 3 --warmup 3 > gpurun_out/r22_n2_$w.json 2> gpurun_out/r22_n2_$w.err; tail -3 gpurun_out/r22_n2_$w.err | cut -c1-300
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r22_n2_$w.json") if l.startswith("{")][-1]; r=d["roofline"]
print("$w N=2 ms/step %.2f value %.0f e2e %.0f local %.1f us vertex %.1f us halo_ok %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3, d["config"]["halo_ok"]))
print(d["config"]["multi_gpu"])
PY
done
timeout 300 python bench.py --workload grid70 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r22_n1_grid70.json 2>/dev/null
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r22_n1_grid70.json") if l.startswith("{")][-1]; r=d["roofline"]
print("grid70 N=1 ms/step %.2f value %.0f local %.1f us vertex %.1f us"%(d["ms_per_step"], d["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3))
PY
