# Round 2, call 3 (N=1): per-body kernel as the default for small-body scenes (A/B on batch64), PCG parity diagnostic.
mkdir -p gpurun_out
T=r2c3
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${T}_pytest.log
timeout 300 python scripts/diag/pcg_parity.py 12 2>&1 | grep -v "^$" | tee gpurun_out/${T}_pcg_parity.txt
bash scripts/gpu_body_ab.sh 2>&1 | tee gpurun_out/${T}_body.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/dbg_small.py 6 2>&1 | tail -4 | tee gpurun_out/${T}_memcheck.log
