# Round 2, call 2 (N=1): the GPU suite with the new grid-family / C1 parity tests, the bench line with its parity and
# faithful records, the reference arm.   gpurun --timeout 900 -- 'bash scripts/gpu_r2_call2.sh'
mkdir -p gpurun_out
T=r2c2
timeout 600 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|error|Error|grid[0-9]+ |C1 step|C1 cube|assert" | tail -60 > gpurun_out/${T}_pytest.log; tail -40 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_grid139.json 2> gpurun_out/${T}_bench_grid139.err; tail -3 gpurun_out/${T}_bench_grid139.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${T}_bench_grid139.json") if l.startswith("{")][-1]
print("ms/step", d["ms_per_step"], "value", d["value"], "parity", d["parity"], "faithful", d["faithful"])
PY
timeout 400 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_reference_grid139.json 2> gpurun_out/${T}_reference_grid139.err; tail -c 400 gpurun_out/${T}_reference_grid139.json
