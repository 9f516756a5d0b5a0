mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r13_pytest.log; cat gpurun_out/r13_pytest.log
PD_PDL=1 timeout 300 python -m pytest tests -m gpu -q -k "batch_of_contexts or bit_for_bit" 2>&1 | tail -8 > gpurun_out/r13_pytest_pdl.log; cat gpurun_out/r13_pytest_pdl.log
