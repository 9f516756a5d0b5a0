# A/B of the PD_H_PLANES experiment (H scratch as x | y | z planes) against the default library on ONE box.
# Build the variant HERE first (variants/ is git-ignored but travels with the gpurun snapshot):
#   PD_OUT=$PWD/soft-body-simulation-cuda_b200/variants/libpd_planes.so PD_DEFS="-DPD_H_PLANES=1" python soft-body-simulation-cuda_b200/build.py
#   PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_planes.so python scripts/check_h_planes_layout.py     # host-side emulation
#   gpurun --timeout 900 -- 'bash scripts/gpu_planes_ab.sh'
mkdir -p gpurun_out
VAR=$PWD/soft-body-simulation-cuda_b200/variants/libpd_planes.so
# 1. correctness of the variant: the oracle / reference / golden parity tests through the variant library (the layout
#    bit-exactness tests of tests/test_host_logic.py are for the default layout and are not run here)
PD_B200_LIB=$VAR timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/planes_pytest.log
# 2. speed, same box, alternating
V="default planes" bash scripts/gpu_ab.sh
# 3. where the wavefronts went: ncu on the variant's local kernel
PD_B200_LIB=$VAR timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_local -s 210 -c 2 -o gpurun_out/planes_k_local_grid139 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/planes_ncu.log 2>&1; tail -2 gpurun_out/planes_ncu.log
