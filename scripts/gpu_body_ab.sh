# PD_BODY_KERNEL experiment (csrc/pd_body_kernel.cuh: one CTA per small body, one launch per step) on the batch workload
# (BASELINE config 5), same box:   gpurun --timeout 900 -- 'bash scripts/gpu_body_ab.sh'
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zz_drag.py -m gpu -q -s -k "per_body" 2>&1 | tail -6 | tee gpurun_out/body_pytest.log
for rep in 1 2; do for b in 0 1; do  # PD_BODY_KERNEL=0: tile path; 1: the default (auto)
  PD_BODY_KERNEL=$b timeout 300 python bench.py --workload batch64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/body${b}_batch64_$rep.json 2> gpurun_out/body${b}_batch64_$rep.err; tail -2 gpurun_out/body${b}_batch64_$rep.err
done; done
python - <<PY
import json
for rep in (1, 2):
    for b in (0, 1):
        try:
            d=[json.loads(l) for l in open(f"gpurun_out/body{b}_batch64_{rep}.json") if l.startswith("{")][-1]
            print("PD_BODY_KERNEL", b, "rep", rep, "batch64 ms/step %.3f value %.0f e2e %.0f launches %d finite %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["run"]["finite"]), d["clocks"]["sm_mhz"])
        except Exception as e: print(b, rep, "failed", e)
PY
