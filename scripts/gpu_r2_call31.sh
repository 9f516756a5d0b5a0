# Round 2, call 31 (N=2, grid70 = the per-rank problem size of grid139 on 8 GPUs ... halved): programmatic dependent launch off / early / late
mkdir -p gpurun_out
T=${T:-r2c31}; N=2; W=grid70
for rep in 1 2; do for v in 0 2 1; do
  PD_PDL=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2969$v bench.py --gpus $N --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_pdl${v}_$rep.json 2> gpurun_out/${T}_pdl${v}_$rep.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${T}_pdl${v}_$rep.json") if l.startswith("{")][-1]
    print("PD_PDL=$v rep $rep $W N=$N ms/step %.3f e2e %.3f halo_ok %s bit_identical %s"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["run"]["halo_ok"], (d.get("parity") or {}).get("bit_identical_to_n1")))
except Exception as e:
    print("PD_PDL=$v rep $rep failed", e); print(open("gpurun_out/${T}_pdl${v}_$rep.err").read()[-800:])
PY
done; done
