#!/bin/bash
# Round 2: ncu --set full captures of one launch of each DMMA kernel class of the sweep step
# (TMA-fed operand loads), each only after the same command has exited 0 without ncu.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python scripts/potrf_once.py > gpurun_out/plain_potrf.log 2>&1 || { echo "potrf_once failed"; exit 1; }
python scripts/profile_step.py 1 > gpurun_out/plain_step.log 2>&1 || { echo "profile_step failed"; exit 1; }
# gemm_nt launches of one m = 8192 factorisation (outer 2048, middle 1024, no look-ahead): 31 panel solves /
# in-panel / middle updates precede the first rank-2048 trailing update (grid 2352)
OCBO_POTRF_LOOKAHEAD=0 $NCU -k regex:gemm_nt_kernel --launch-skip 31 --launch-count 1 -o gpurun_out/r02_full_outer_update \
    python scripts/potrf_once.py > gpurun_out/ncu_f1.log 2>&1; echo "outer update exit $?"
$NCU -k regex:"postcov_kernel" --launch-count 1 -o gpurun_out/r02_full_postcov \
    python scripts/profile_step.py 1 > gpurun_out/ncu_f3.log 2>&1; echo "postcov exit $?"
$NCU -k regex:"sample_argmax_kernel" --launch-count 1 -o gpurun_out/r02_full_sample_argmax \
    python scripts/profile_step.py 1 > gpurun_out/ncu_f5.log 2>&1; echo "sample exit $?"
$NCU -k regex:"gram_kernel" --launch-count 2 -o gpurun_out/r02_full_gram \
    python scripts/profile_step.py 1 > gpurun_out/ncu_f4.log 2>&1; echo "gram exit $?"
ls -la gpurun_out/*.ncu-rep
