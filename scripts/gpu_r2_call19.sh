mkdir -p gpurun_out
T=${T:-r2c19}
timeout 300 python -m pytest tests/test_gpu_dist.py tests/test_gpu_solvers.py -m gpu -q -x 2>&1 | tail -60 | cut -c1-400 > gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_pytest.log
