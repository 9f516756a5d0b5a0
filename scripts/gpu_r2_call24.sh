mkdir -p gpurun_out
timeout 300 python scripts/phase_profile.py grid139 2>&1 | tail -8
timeout 300 python scripts/phase_profile.py grid139 0 2 2>&1 | tail -8
