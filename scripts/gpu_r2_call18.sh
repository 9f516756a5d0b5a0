# Round 2, call 18 (N=2): halo push with the flag IN the data (default) vs fence + ticket + flags (variants/libpd_flags.so = the previous commit)
mkdir -p gpurun_out
T=${T:-r2c18}; N=${N:-2}; W=${W:-grid70}
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_solvers.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.log
for rep in 1 2; do for v in default flags; do
  if [ $v = default ]; then unset PD_B200_LIB; else export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$rep bench.py --gpus $N --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_${v}_n${N}_${W}_$rep.json 2> gpurun_out/${T}_${v}_n${N}_${W}_$rep.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${T}_${v}_n${N}_${W}_$rep.json") if l.startswith("{")][-1]
    print("$v rep $rep $W N=$N ms/step %.3f e2e %.3f halo_ok %s bit_identical %s"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["run"]["halo_ok"], (d.get("parity") or {}).get("bit_identical_to_n1")), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$v rep $rep failed", e); print(open("gpurun_out/${T}_${v}_n${N}_${W}_$rep.err").read()[-1500:])
PY
done; done
unset PD_B200_LIB
