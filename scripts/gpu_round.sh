#!/bin/bash
# Round artefacts: parity tests, smoke, the default bench line, ncu launch list of a
# short bench command, DRAM traffic of the DMMA update kernel.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -2 gpurun_out/bench.err
BENCH_SHORT="python bench.py --steps 1 --warmup 1 --streams 1 --no-cpu-baseline --no-peaks --sweep-trials 0 --no-ladder-variant"
$BENCH_SHORT > gpurun_out/plain_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bench_launches.csv \
    $BENCH_SHORT > gpurun_out/ncu_bench.log 2>&1
echo "ncu bench exit $?"
python scripts/profile_step.py 2 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"gemm_nt_kernel|postcov|gemm_nn|sample_argmax" --csv --log-file gpurun_out/traffic.csv \
    python scripts/profile_step.py 2 > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic exit $?"
