# vertex kernel with batched slot loads: parity suite + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r25_pytest.log; cat gpurun_out/r25_pytest.log
for rep in 1 2; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r25_grid139_$rep.json 2> gpurun_out/r25_grid139_$rep.err; tail -2 gpurun_out/r25_grid139_$rep.err
done
python - <<PY
import json
for rep in [1,2]:
    d=[json.loads(l) for l in open(f"gpurun_out/r25_grid139_{rep}.json") if l.startswith("{")][-1]; r=d["roofline"]
    print(rep, "ms/step %.3f value %.0f e2e %.0f local %.1f us vertex %.1f us frac %.3f traffic %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3, r["frac"], r["traffic"]), d["clocks"]["sm_mhz"])
PY
