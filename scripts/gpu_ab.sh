# Same-box A/B of kernel variants (box-to-box noise is 2-3 %: never compare numbers of different gpurun calls).
# Build the variants here first, e.g.
#   PD_OUT=soft-body-simulation-cuda_b200/variants/libpd_g16.so PD_DEFS="-DPD_TILE_GROUP=16" python soft-body-simulation-cuda_b200/build.py
# then  gpurun -- 'V="default g16" bash scripts/gpu_ab.sh'   (PD_B200_LIB selects the library; experiments only)
mkdir -p gpurun_out
W=${W:-grid139}
for rep in 1 2; do for v in ${V:-default}; do
  if [ $v = default ]; then unset PD_B200_LIB; else export PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so; fi
  timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-faithful > gpurun_out/ab_${W}_${v}_$rep.json 2> gpurun_out/ab_${W}_${v}_$rep.err; tail -2 gpurun_out/ab_${W}_${v}_$rep.err
done; done
unset PD_B200_LIB
python - <<PY
import json
for rep in [1, 2]:
    for v in "${V:-default}".split():
        try:
            d=[json.loads(l) for l in open(f"gpurun_out/ab_${W}_{v}_{rep}.json") if l.startswith("{")][-1]; r=d["roofline"]
            print(v, rep, "ms/step %.3f value %.0f local %.1f us vertex %.1f us"%(d["ms_per_step"], d["value"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3), d["clocks"]["sm_mhz"])
        except Exception as e: print(v, "failed", e)
PY
