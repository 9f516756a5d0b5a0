# Round 2, call 14 (N=2): where the multi-GPU iteration loses time -- timing probes (PD_DIST_DEBUG: 1 no wait, 2 no push, 3 neither)
mkdir -p gpurun_out
T=${T:-r2c14}; N=${N:-2}; W=${W:-grid70}
for rep in 1 2; do for v in 0 1 2 3; do
  PD_DIST_DEBUG=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$v bench.py --gpus $N --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-parity --rot-mode ${RM:-2} > gpurun_out/${T}_dbg${v}_n${N}_${W}_$rep.json 2> gpurun_out/${T}_dbg${v}_n${N}_${W}_$rep.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${T}_dbg${v}_n${N}_${W}_$rep.json") if l.startswith("{")][-1]
    print("PD_DIST_DEBUG=$v rep $rep $W N=$N ms/step %.3f e2e %.3f local alone %.1f us vertex alone %.1f us"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["launch_ms"]*1e3, d["roofline"]["fused_iteration"]["vertex_kernel_ms"]*1e3), d["clocks"]["sm_mhz"])
except Exception as e: print("PD_DIST_DEBUG=$v rep $rep failed", e)
PY
done; done
# the same per-rank problem size on ONE GPU (no partition at all): grid55 = 0.998 M tets
for rep in 1 2; do
  timeout 300 python bench.py --workload grid55 --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-faithful --rot-mode ${RM:-2} > gpurun_out/${T}_grid55_n1_$rep.json 2>/dev/null
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${T}_grid55_n1_$rep.json") if l.startswith("{")][-1]
print("grid55 N=1 rep $rep ms/step %.3f local alone %.1f us vertex alone %.1f us"%(d["ms_per_step"], d["roofline"]["launch_ms"]*1e3, d["roofline"]["fused_iteration"]["vertex_kernel_ms"]*1e3))
PY
done
