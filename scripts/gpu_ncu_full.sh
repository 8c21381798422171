#!/bin/bash
# ncu --set full captures of one launch of each kernel class of the sweep step
# (each only after the same command has exited 0 without ncu).
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python scripts/potrf_once.py > gpurun_out/plain_potrf.log 2>&1 || { echo "potrf_once failed"; exit 1; }
python scripts/profile_step.py 1 > gpurun_out/plain_step.log 2>&1 || { echo "profile_step failed"; exit 1; }
# gemm_nt launches of one m = 8192 factorisation, outer block 2048: 1 panel solve (column 0), then
# (update, panel) for columns 1..15 = 31 launches, then the first rank-2048 trailing update
$NCU -k regex:gemm_nt_kernel --launch-skip 31 --launch-count 1 -o gpurun_out/full_outer_update \
    python scripts/potrf_once.py > gpurun_out/ncu_f1.log 2>&1; echo "outer update exit $?"
$NCU -k regex:potrf_diag_blk --launch-count 1 -o gpurun_out/full_diag_blk \
    python scripts/potrf_once.py 1024 > gpurun_out/ncu_f2.log 2>&1; echo "diag exit $?"
$NCU -k regex:"gram_kernel|postcov_kernel" --launch-count 3 -o gpurun_out/full_gram_postcov \
    python scripts/profile_step.py 1 > gpurun_out/ncu_f3.log 2>&1; echo "gram/postcov exit $?"
$NCU -k regex:"post_var_ei|segment_argmax|philox_normal|joint_pick|solve_alpha" --launch-count 5 \
    -o gpurun_out/full_reduce python scripts/profile_step.py 1 > gpurun_out/ncu_f4.log 2>&1; echo "reduce exit $?"
$NCU -k regex:gemm_nn_kernel --launch-skip 15 --launch-count 1 -o gpurun_out/full_sample \
    python scripts/profile_step.py 1 > gpurun_out/ncu_f5.log 2>&1; echo "sample exit $?"
ls -la gpurun_out/*.ncu-rep
python -c "
from tests.tools import bench_small
print(bench_small.measure_revi(1))" > gpurun_out/plain_revi.log 2>&1 && \
$NCU -k regex:kg_kernel --launch-count 1 -o gpurun_out/full_kg python -c "
from tests.tools import bench_small
print(bench_small.measure_revi(1))" > gpurun_out/ncu_kg.log 2>&1; echo "kg exit $?"
