# round-1 artefacts of the v7 local kernel: parity suite, default bench (with cpu_baseline) + reference arm, ncu launch
# list of the bench command, full ncu captures of the two per-iteration kernels, batch64 and grid55 lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/art_pytest_full.log 2>&1; tail -3 gpurun_out/art_pytest_full.log
timeout 600 python bench.py > gpurun_out/art_bench_grid139.json 2> gpurun_out/art_bench_grid139.err; tail -2 gpurun_out/art_bench_grid139.err
timeout 600 python bench.py --impl reference > gpurun_out/art_reference_grid139.json 2> gpurun_out/art_reference_grid139.err; tail -2 gpurun_out/art_reference_grid139.err
for w in batch64 grid55; do timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/art_bench_$w.json 2> gpurun_out/art_bench_$w.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/art_launches_grid139.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/art_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local -s 12 -c 2 -f -o gpurun_out/art_k_local_grid139 python scripts/profile_step.py grid139 2 10 > gpurun_out/art_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vertex_jacobi -s 12 -c 2 -f -o gpurun_out/art_k_vertex_grid139 python scripts/profile_step.py grid139 2 10 >> gpurun_out/art_prof.log 2>&1
timeout 300 python scripts/phase_profile.py grid139 > gpurun_out/art_phase.txt 2>&1
python - <<PY
import json
for w in ["bench_grid139","reference_grid139","bench_batch64","bench_grid55"]:
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/art_{w}.json") if l.startswith("{")][-1]; r=d.get("roofline") or {}
        print(w, "ms/step %.3f value %.0f e2e %.0f"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), "local %.1f us frac %.3f"%(r.get("launch_ms",0)*1e3, r.get("frac",0)), d.get("cpu_baseline"))
    except Exception as e: print(w,"failed",e)
PY
