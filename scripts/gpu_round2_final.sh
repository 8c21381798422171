#!/bin/bash
# Round-2 artefacts (int8-tensor-core path default): parity tests, smoke, default bench line, ncu launch list of a
# short bench command, DRAM traffic and --set full captures of the Ozaki kernels.  Each ncu command runs only after
# the same command has exited 0 without ncu; no number printed under ncu is a bench value.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
OCBO_OZAKI=0 timeout 1200 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_parity.py -m gpu -q --timeout=900 > gpurun_out/pytest_gpu_fp64_path.log 2>&1
echo "pytest (OCBO_OZAKI=0: sweep + parity files) exit $?" >> gpurun_out/pytest_gpu_fp64_path.log; tail -2 gpurun_out/pytest_gpu_fp64_path.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -2 gpurun_out/bench.err
BENCH_SHORT="python bench.py --steps 1 --warmup 1 --streams 1 --no-cpu-baseline --no-peaks --sweep-trials 0 --no-ladder-variant"
$BENCH_SHORT > gpurun_out/plain_short.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bench_launches.csv \
    $BENCH_SHORT > gpurun_out/ncu_bench.log 2>&1
echo "ncu bench exit $?"
python scripts/profile_step.py 2 > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"ozaki_update" --csv --log-file gpurun_out/r02_traffic_ozaki.csv \
    python scripts/profile_step.py 2 > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic exit $?"
NCU="timeout 600 ncu --set full --clock-control none --import-source on"
OCBO_POTRF_LOOKAHEAD=0 python scripts/potrf_ws_once.py > gpurun_out/plain_potrf_ws.log 2>&1 || { echo "potrf_ws_once failed"; exit 1; }
# first trailing update of the m = 8192 factorisation (6144 x 2048, lower tiles): before it come the in-panel updates
# of block columns 1-7, the rank-1024 middle update and the in-panel updates of block columns 9-15 (15 launches)
OCBO_POTRF_LOOKAHEAD=0 $NCU -k regex:ozaki_update_kernel --launch-skip 15 --launch-count 1 -o gpurun_out/r02_full_ozaki_update \
    python scripts/potrf_ws_once.py > gpurun_out/ncu_f1.log 2>&1; echo "ozaki update exit $?"
OCBO_POTRF_LOOKAHEAD=0 $NCU -k regex:ozaki_split_fixed_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/r02_full_ozaki_split \
    python scripts/potrf_ws_once.py > gpurun_out/ncu_f2.log 2>&1; echo "ozaki split exit $?"
# the step's first ozaki_update launch is the posterior covariance (MODE 2), preceded by the column-wise split of V
$NCU -k regex:"ozaki_update_kernel|ozaki_split_t_kernel" --launch-count 2 -o gpurun_out/r02_full_ozaki_postcov \
    python scripts/profile_step.py 1 > gpurun_out/ncu_f3.log 2>&1; echo "postcov exit $?"
# the sampling tail: ozaki_update_kernel<1> (the only launch whose name ends in <1>)
$NCU --kernel-name-base mangled -k regex:ozaki_update_kernelILi1E --launch-count 1 -o gpurun_out/r02_full_ozaki_sample \
    python scripts/profile_step.py 1 > gpurun_out/ncu_f5.log 2>&1; echo "sample exit $?"
ls -la gpurun_out/r02_full_*.ncu-rep
