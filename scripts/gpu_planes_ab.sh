# Same-box A/B of the prepared k_local experiments against the default library:
#   planes = -DPD_H_PLANES=1    the H scratch as x | y | z planes with a 32-colouring (3 wavefronts per 32 entries instead of 4)
#   pred   = -DPD_PHASEC_PRED=1 pads of the incidence rows are not loaded (quarter-warps of pads cost no wavefront)
# Build the variants HERE first (variants/ is git-ignored but travels with the gpurun snapshot):
#   for v in "planes -DPD_H_PLANES=1" "pred -DPD_PHASEC_PRED=1"; do set -- $v; PD_OUT=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$1.so PD_DEFS="$2" python soft-body-simulation-cuda_b200/build.py; done
#   PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_planes.so python scripts/check_h_planes_layout.py     # host-side check of the layout
#   python -m pytest tests/test_kernel_emulation.py -q                                                                 # the variants' kernels on the host
#   gpurun --timeout 1200 -- 'bash scripts/gpu_planes_ab.sh'
mkdir -p gpurun_out
for v in ${VARIANTS:-planes pred}; do
  VAR=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so
  # 1. correctness of the variant: the oracle / reference / golden parity tests through the variant library (the layout
  #    bit-exactness tests of tests/test_host_logic.py are for the default layout and are not run here)
  PD_B200_LIB=$VAR timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/${v}_pytest.log
done
# 2. speed, same box, alternating
V="default ${VARIANTS:-planes pred}" bash scripts/gpu_ab.sh
# 3. where the wavefronts went: ncu on each variant's local kernel
for v in ${VARIANTS:-planes pred}; do
  PD_B200_LIB=$PWD/soft-body-simulation-cuda_b200/variants/libpd_$v.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_local -s 210 -c 2 -o gpurun_out/${v}_k_local_grid139 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${v}_ncu.log 2>&1; tail -2 gpurun_out/${v}_ncu.log
done
