mkdir -p gpurun_out
for cfg in "0 0" "2 0" "0 3" "0 2" "2 2" "1 0"; do set -- $cfg
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --rot-mode $1 --ctas-per-sm $2 > gpurun_out/exp1_$1_$2.json 2>gpurun_out/exp1_$1_$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/exp1_$1_$2.json")); r=d["roofline"]
print("rot $1 ctas $2: ms/step %.2f local %.1f us vertex %.1f us"%(d["ms_per_step"], r["launch_ms"]*1e3, r["fused_iteration"]["vertex_kernel_ms"]*1e3))
PY
done
