# Round 2, call 27 (N=8): config 5 as specified (512 contexts, 64 per GPU, per-body kernel, no communication) and grid139 with the final code
mkdir -p gpurun_out
T=${T:-r2c27}; N=8
for w in batch64 grid139; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus $N --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_n${N}_$w.json 2> gpurun_out/${T}_n${N}_$w.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${T}_n${N}_$w.json") if l.startswith("{")][-1]
    print("$w N=$N ms/step %.3f value %.0f e2e %.3f ms (%.0f) scaling %s halo_ok %s parity %s"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["scaling"], d["run"]["halo_ok"], d.get("parity")), d["clocks"]["sm_mhz"], d["gpu_launches"])
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/${T}_n${N}_$w.err").read()[-1500:])
PY
done
timeout 300 python bench.py --workload batch64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_n1_batch64.json 2>/dev/null
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${T}_n1_batch64.json") if l.startswith("{")][-1]
print("batch64 N=1 ms/step %.3f value %.0f e2e %.0f"%(d["ms_per_step"], d["value"], d["e2e"]["value"]))
PY
