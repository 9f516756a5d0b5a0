mkdir -p gpurun_out
TAG=${TAG:-r9}
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_grid139.json 2> gpurun_out/${TAG}_bench_grid139.err; tail -5 gpurun_out/${TAG}_bench_grid139.err
timeout 600 python bench.py --workload grid55 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_grid55.json 2> gpurun_out/${TAG}_bench_grid55.err
python - <<PY
import json
for w in ("grid139","grid55"):
    d=json.load(open("gpurun_out/${TAG}_bench_%s.json"%w)); r=d["roofline"]
    print(w, "ms/step %.2f value %.0f e2e %.0f local %.1f us (frac %.3f) vertex %.1f us"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["frac"], r["fused_iteration"]["vertex_kernel_ms"]*1e3))
PY
timeout 300 python scripts/phase_profile.py grid139 0 0 2>&1 | tail -9
