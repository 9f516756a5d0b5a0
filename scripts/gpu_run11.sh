mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --workload batch64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r11_batch64_n1.json 2> gpurun_out/r11_batch64_n1.err; tail -3 gpurun_out/r11_batch64_n1.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r11_batch64_n1.json") if l.startswith("{")][-1]; r=d["roofline"]
print("batch64 N=1 ms/step %.3f value %.0f e2e %.0f local %.1f us vertex %.1f us"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3), d["config"]["num_tets"])
PY
