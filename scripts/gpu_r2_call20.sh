mkdir -p gpurun_out
T=${T:-r2c20}; N=${N:-2}; W=${W:-grid70}
timeout 400 python -m pytest tests/test_gpu_dist.py tests/test_gpu_solvers.py -m gpu -q 2>&1 | tail -12 | cut -c1-300 > gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_pytest.log
PD_BENCH_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_n${N}_${W}.json 2> gpurun_out/${T}_n${N}_${W}.err
grep "trace" gpurun_out/${T}_n${N}_${W}.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${T}_n${N}_${W}.json") if l.startswith("{")][-1]
print("$W N=$N ms/step %.3f e2e %.3f halo_ok %s bit_identical %s"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["run"]["halo_ok"], (d.get("parity") or {}).get("bit_identical_to_n1")))
PY
