# Round 2, call 11 (N=1): whole GPU suite (first run of test_gpu_linear), default bench line, launch list
mkdir -p gpurun_out
T=r2c11
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | cut -c1-300 > gpurun_out/${T}_pytest.log; tail -5 gpurun_out/${T}_pytest.log
grep -E "PCG-Jacobi|CG \+ IC|vs the reference's" gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err; cut -c1-1500 gpurun_out/${T}_bench.json
