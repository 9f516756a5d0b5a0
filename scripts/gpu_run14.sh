# HEAD validation: GPU parity suite, default bench (with cpu_baseline), small-size bench, ncu launch list of the
# bench command, full ncu captures of the two per-iteration kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r14_pytest.log; cat gpurun_out/r14_pytest.log
timeout 600 python bench.py > gpurun_out/r14_bench_grid139.json 2> gpurun_out/r14_bench_grid139.err; tail -2 gpurun_out/r14_bench_grid139.err
timeout 300 python bench.py --workload grid70 --no-cpu-baseline > gpurun_out/r14_bench_grid70.json 2> gpurun_out/r14_bench_grid70.err; tail -2 gpurun_out/r14_bench_grid70.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/r14_launches_grid139.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r14_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local -s 12 -c 2 -f -o gpurun_out/r14_k_local_grid139 python scripts/profile_step.py grid139 2 10 > gpurun_out/r14_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vertex_jacobi -s 12 -c 2 -f -o gpurun_out/r14_k_vertex_grid139 python scripts/profile_step.py grid139 2 10 >> gpurun_out/r14_prof.log 2>&1
python - <<PY
import json
for w in ["grid139","grid70"]:
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/r14_bench_{w}.json") if l.startswith("{")][-1]; r=d["roofline"]
        print(w, "ms/step %.3f value %.0f e2e %.0f local %.1f us vertex %.1f us frac %.3f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3, r["frac"]), d["clocks"], d["cpu_baseline"])
    except Exception as e: print(w,"failed",e)
PY
ls -la gpurun_out | tail -12
