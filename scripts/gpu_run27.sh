# scaling record on one 8-GPU box: grid139 at N=8 and N=4, batch64 at N=8
mkdir -p gpurun_out
run() { # N workload tag
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $1 --workload $2 --steps 5 --warmup 3 > gpurun_out/r27_n$1_$2.json 2> gpurun_out/r27_n$1_$2.err; tail -2 gpurun_out/r27_n$1_$2.err | cut -c1-200
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r27_n$1_$2.json") if l.startswith("{")][-1]; r=d["roofline"]
    print("$2 N=$1 ms/step %.3f value %.0f e2e %.0f local %.1f us vertex %.1f us halo_ok %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3, d["config"]["halo_ok"]), d["clocks"]["sm_mhz"])
    print(d["config"]["multi_gpu"])
except Exception as e: print("$2 N=$1 failed", e)
PY
}
run 8 grid139
run 4 grid139
run 8 batch64
