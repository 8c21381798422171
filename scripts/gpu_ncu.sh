#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_step.py 2 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_step.py 2 > gpurun_out/ncu1.log 2>&1
echo "ncu1 exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 160 -c 10 \
    -o gpurun_out/prof_gemm_nt python scripts/profile_step.py 2 > gpurun_out/ncu_b.log 2>&1
echo "ncu_b exit $?"
ncu --set full --clock-control none --import-source on -k regex:"postcov|gemm_nn" -s 16 -c 4 \
    -o gpurun_out/prof_misc python scripts/profile_step.py 2 > gpurun_out/ncu_c.log 2>&1
echo "ncu_c exit $?"
