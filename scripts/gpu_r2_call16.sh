# Round 2, call 16 (N=2): halo push in the first blocks of the vertex kernel (default) vs at the start of the local kernel (PD_PUSH_IN_VERTEX=0)
mkdir -p gpurun_out
T=${T:-r2c16}; N=${N:-2}; W=${W:-grid70}
timeout 500 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
for rep in 1 2; do for v in 1 0; do
  PD_PUSH_IN_VERTEX=$v timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2959$v bench.py --gpus $N --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_piv${v}_n${N}_${W}_$rep.json 2> gpurun_out/${T}_piv${v}_n${N}_${W}_$rep.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${T}_piv${v}_n${N}_${W}_$rep.json") if l.startswith("{")][-1]
    print("PD_PUSH_IN_VERTEX=$v rep $rep $W N=$N ms/step %.3f e2e %.3f halo_ok %s bit_identical %s launches %d"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["run"]["halo_ok"], (d.get("parity") or {}).get("bit_identical_to_n1"), d["gpu_launches"]), d["clocks"]["sm_mhz"])
except Exception as e:
    print("PD_PUSH_IN_VERTEX=$v rep $rep failed", e); print(open("gpurun_out/${T}_piv${v}_n${N}_${W}_$rep.err").read()[-1500:])
PY
done; done
