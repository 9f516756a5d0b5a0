# round-2 artefacts of the final code (N=1): whole GPU suite, default bench (with cpu_baseline) + reference arm, the other
# workloads, ncu launch list of the bench command, full ncu captures of the two per-iteration kernels
mkdir -p gpurun_out
T=${T:-r2fin}
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${T}_pytest_full.log 2>&1; tail -3 gpurun_out/${T}_pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench_grid139.json 2> gpurun_out/${T}_bench_grid139.err; tail -2 gpurun_out/${T}_bench_grid139.err
timeout 600 python bench.py --impl reference > gpurun_out/${T}_reference_grid139.json 2> gpurun_out/${T}_reference_grid139.err; tail -2 gpurun_out/${T}_reference_grid139.err
for w in armadillo grid55 batch64 grid55-pcg grid55-pcg-tol; do timeout 400 python bench.py --workload $w --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; tail -1 gpurun_out/${T}_bench_$w.err; done
timeout 400 python bench.py --workload armadillo --impl reference --no-cpu-baseline > gpurun_out/${T}_reference_armadillo.json 2> gpurun_out/${T}_reference_armadillo.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/${T}_launches_grid139.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-faithful > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_local -s 12 -c 2 -f -o gpurun_out/${T}_k_local_grid139 python scripts/profile_step.py grid139 2 10 > gpurun_out/${T}_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vertex_jacobi -s 12 -c 2 -f -o gpurun_out/${T}_k_vertex_grid139 python scripts/profile_step.py grid139 2 10 >> gpurun_out/${T}_prof.log 2>&1
python - <<PY
import json
for w in ["bench_grid139","reference_grid139","bench_armadillo","reference_armadillo","bench_grid55","bench_batch64","bench_grid55-pcg","bench_grid55-pcg-tol"]:
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/${T}_{w}.json") if l.startswith("{")][-1]; r=d.get("roofline") or {}
        print(w, "ms/step %.3f value %.0f e2e %.0f"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), "kernel %.1f us frac %.3f"%(r.get("launch_ms",0)*1e3, r.get("frac",0)), "parity", (d.get("parity") or {}).get("rel_err"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(w,"failed",e)
PY
