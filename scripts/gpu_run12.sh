mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r12_pytest.log; cat gpurun_out/r12_pytest.log
for pdl in 0 1; do
  PD_PDL=$pdl timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r12_grid139_pdl$pdl.json 2> gpurun_out/r12_grid139_pdl$pdl.err; tail -2 gpurun_out/r12_grid139_pdl$pdl.err
  PD_PDL=$pdl timeout 300 python bench.py --workload batch64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r12_batch64_pdl$pdl.json 2> gpurun_out/r12_batch64_pdl$pdl.err; tail -2 gpurun_out/r12_batch64_pdl$pdl.err
done
python - <<PY
import json
for w in ["grid139","batch64"]:
    for pdl in [0,1]:
        try:
            d=[json.loads(l) for l in open(f"gpurun_out/r12_{w}_pdl{pdl}.json") if l.startswith("{")][-1]; r=d["roofline"]
            print(w, "pdl",pdl, "ms/step %.3f value %.0f e2e %.0f local %.1f us vertex %.1f us"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3), d["clocks"])
        except Exception as e: print(w,pdl,"failed",e)
PY
