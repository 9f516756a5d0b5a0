# Round 2, call 33 (N=8): grid139 with the final code (late-trigger PDL default), PD_PDL=0 beside it
mkdir -p gpurun_out
T=${T:-r2c33}; N=8; W=grid139
for v in 2 0; do
PD_PDL=$v timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2970$v bench.py --gpus $N --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_pdl${v}_n${N}.json 2> gpurun_out/${T}_pdl${v}_n${N}.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${T}_pdl${v}_n${N}.json") if l.startswith("{")][-1]
    print("PD_PDL=$v $W N=$N ms/step %.3f value %.0f e2e %.3f halo_ok %s bit_identical %s"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["run"]["halo_ok"], (d.get("parity") or {}).get("bit_identical_to_n1")), d["clocks"]["sm_mhz"])
except Exception as e:
    print("PD_PDL=$v failed", e); print(open("gpurun_out/${T}_pdl${v}_n${N}.err").read()[-1500:])
PY
done
