# Round 2, call 10 (N=1): collision tests again, Cholesky critical path, late-trigger PDL A/B on grid139
mkdir -p gpurun_out
T=r2c10
timeout 900 python -m pytest tests/test_gpu_collision.py tests/test_gpu_solvers.py -m gpu -q -s 2>&1 | grep -E "passed|failed|ccd queries|toi bit|oracle vs|collision step|house|Cholesky|PCG" | cut -c1-260 | tee gpurun_out/${T}_pytest.log
for b in 1 2 4 8; do
  PD_CHOL_BLOCKS_PER_SM=$b timeout 300 python - <<PY 2>&1 | tail -1 | tee -a gpurun_out/${T}_chol.log
import sys; sys.path.insert(0, "scripts"); sys.argv = ["x"]
import json, solver_bench as S
print("blocks/SM $b", json.dumps({k: v for k, v in S.case(24, 1, 5, 0, 0.0).items() if k in ("ms_per_step", "nnz_L", "pd_iterations", "finite")}))
PY
done
for rep in 1 2; do for v in 0 2; do
  PD_PDL=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-faithful > gpurun_out/${T}_pdl${v}_$rep.json 2> gpurun_out/${T}_pdl${v}_$rep.err
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${T}_pdl${v}_$rep.json") if l.startswith("{")][-1]
print("PD_PDL=$v rep $rep ms/step %.3f value %.0f"%(d["ms_per_step"], d["value"]), d["clocks"]["sm_mhz"])
PY
done; done
# split phase C (default library) vs -DPD_SPLIT_C=0 (variants/libpd_nosplit.so), same box
V="default nosplit" bash scripts/gpu_ab.sh 2>&1 | tee gpurun_out/${T}_ab_split.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_local -s 210 -c 1 -o gpurun_out/${T}_k_local_grid139_split python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-faithful > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log
