mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_full.log 2>&1
grep -n "steps 1..3\|pinned sphere\|free fall 1M\|C1 step\|C1 cube rot\|engine vs reference\|centroid\|passed\|failed\|FAILED\|^E  " gpurun_out/pytest_gpu_full.log | head -80
