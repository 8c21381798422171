#!/bin/bash
cd "$(dirname "$0")/.."
NCU="timeout 600 ncu --set full --clock-control none --import-source on"
OCBO_POTRF_LOOKAHEAD=0 $NCU -k regex:ozaki_update_kernel --launch-skip 15 --launch-count 1 -o gpurun_out/r02_full_ozaki_update -f \
    python scripts/potrf_ws_once.py > gpurun_out/ncu_f1.log 2>&1; echo "ozaki update exit $?"
OCBO_POTRF_LOOKAHEAD=0 $NCU -k regex:ozaki_split_fixed_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/r02_full_ozaki_split -f \
    python scripts/potrf_ws_once.py > gpurun_out/ncu_f2.log 2>&1; echo "ozaki split exit $?"
$NCU --kernel-name-base mangled -k regex:ozaki_update_kernelILi1E --launch-count 1 -o gpurun_out/r02_full_ozaki_sample -f \
    python scripts/profile_step.py 1 > gpurun_out/ncu_f5.log 2>&1; echo "sample exit $?"
ls -la gpurun_out/r02_full_*.ncu-rep
