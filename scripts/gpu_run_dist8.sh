mkdir -p gpurun_out
N=8 bash scripts/gpu_run_distN.sh
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --workload batch64 --steps 5 --warmup 3 > gpurun_out/batch64_n8.json 2> gpurun_out/batch64_n8.err; tail -2 gpurun_out/batch64_n8.err | cut -c1-200
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/batch64_n8.json") if l.startswith("{")][-1]
print("batch64 N=8 ms/step %.3f value %.0f e2e %.0f tets %d"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["num_tets"]))
PY
