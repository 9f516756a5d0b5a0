# Round 2, call 21 (N=8): grid139 strong scaling with the tagged halo push; dist tests on 8 GPUs' box; in-situ kernel times
mkdir -p gpurun_out
T=${T:-r2c21}; N=${N:-8}; W=${W:-grid139}
timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
for rep in 1 2; do
PD_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$rep bench.py --gpus $N --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_n${N}_${W}_$rep.json 2> gpurun_out/${T}_n${N}_${W}_$rep.err
grep "trace\] rank 0" gpurun_out/${T}_n${N}_${W}_$rep.err | tr '\n' ';'; echo
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${T}_n${N}_${W}_$rep.json") if l.startswith("{")][-1]
print("$W N=$N rep $rep ms/step %.3f value %.0f e2e %.3f halo_ok %s bit_identical %s"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["run"]["halo_ok"], (d.get("parity") or {}).get("bit_identical_to_n1")), d["clocks"]["sm_mhz"])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 4 --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_n4_${W}.json 2> gpurun_out/${T}_n4_${W}.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${T}_n4_${W}.json") if l.startswith("{")][-1]
print("$W N=4 ms/step %.3f value %.0f e2e %.3f halo_ok %s bit_identical %s"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["run"]["halo_ok"], (d.get("parity") or {}).get("bit_identical_to_n1")), d["clocks"]["sm_mhz"])
PY
timeout 300 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-faithful > gpurun_out/${T}_n1_${W}.json 2>/dev/null
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${T}_n1_${W}.json") if l.startswith("{")][-1]
print("$W N=1 ms/step %.3f value %.0f e2e %.3f"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), d["clocks"]["sm_mhz"])
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29632 scripts/dist_perf.py $W 3 2>&1 | grep -v "^\*\|OMP_NUM\|NCCL version\|\[bench\]" | tee gpurun_out/${T}_dist_perf_n${N}.txt
