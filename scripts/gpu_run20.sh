# local kernel v7d: staging-slot colouring (fewer bank conflicts on the position loads) + deeper table prefetch
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r20_pytest.log; cat gpurun_out/r20_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r20_grid139.json 2> gpurun_out/r20_grid139.err; tail -2 gpurun_out/r20_grid139.err
PD_NO_STAGE_COLOR=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r20_grid139_nocolor.json 2> gpurun_out/r20_grid139_nocolor.err; tail -2 gpurun_out/r20_grid139_nocolor.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local -s 12 -c 1 -f -o gpurun_out/r20_k_local_grid139 python scripts/profile_step.py grid139 2 10 > gpurun_out/r20_prof.log 2>&1
python - <<PY
import json
for v in ["grid139","grid139_nocolor"]:
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r20_{v}.json") if l.startswith("{")][-1]; r=d["roofline"]
        print(v, "ms/step %.3f value %.0f e2e %.0f local %.1f us vertex %.1f us frac %.3f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3, r["frac"]), d["clocks"]["sm_mhz"])
    except Exception as e: print(v,"failed",e)
PY
