#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_step.py 2 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:potrf_diag -s 80 -c 1 \
    -o gpurun_out/prof_potrf_diag python scripts/profile_step.py 2 > gpurun_out/ncu_a.log 2>&1
echo "ncu_a exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 158 -c 8 \
    -o gpurun_out/prof_gemm_nt python scripts/profile_step.py 2 > gpurun_out/ncu_b.log 2>&1
echo "ncu_b exit $?"
ncu --set full --clock-control none --import-source on -k regex:"postcov|gemm_nn|solve_alpha" -s 17 -c 6 \
    -o gpurun_out/prof_misc python scripts/profile_step.py 2 > gpurun_out/ncu_c.log 2>&1
echo "ncu_c exit $?"
