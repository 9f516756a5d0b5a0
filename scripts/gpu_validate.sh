# final N=1 validation of the round: parity suite + default bench line (with cpu_baseline)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/val_pytest.log
timeout 200 python bench.py --steps 3 --warmup 3 > gpurun_out/val_bench_grid139.json 2> gpurun_out/val_bench_grid139.err; tail -2 gpurun_out/val_bench_grid139.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/val_bench_grid139.json") if l.startswith("{")][-1]; r=d["roofline"]
print("ms/step %.3f value %.0f e2e %.0f local %.1f us vertex %.1f us frac %.3f fused %.3f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3, r["frac"], r["fused_iteration"]["frac"]), d["clocks"]["sm_mhz"], d["cpu_baseline"]["value"])
PY
