set -x
mkdir -p gpurun_out
# launch list of 2 steps x 10 iterations (plain launches), grid139
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_grid139.csv python scripts/profile_step.py grid139 2 10 > gpurun_out/prof1.log 2>&1
# full capture of the local kernel and the vertex kernel (second step)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_local -s 12 -c 2 -f -o gpurun_out/r1_k_local_grid139 python scripts/profile_step.py grid139 2 10 >> gpurun_out/prof1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vertex_jacobi -s 12 -c 2 -f -o gpurun_out/r1_k_vertex_grid139 python scripts/profile_step.py grid139 2 10 >> gpurun_out/prof1.log 2>&1
tail -5 gpurun_out/prof1.log
ls -la gpurun_out
