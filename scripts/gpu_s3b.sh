#!/bin/bash
# session 3, call b: in-flight tests, then whole-loop iterations/s with trials in flight
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_inflight.py tests/test_gpu_graphs.py tests/test_gpu_strategies.py tests/test_options_drivers.py -m gpu -q -x --timeout=300 > gpurun_out/pytest_inflight.log 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/pytest_inflight.log
timeout 600 python -c "
import json
from tests.tools import bench_small
print(json.dumps(bench_small.measure_bo_loop()))" > gpurun_out/bo_loop.log 2>&1; tail -3 gpurun_out/bo_loop.log
timeout 600 python tests/tools/bench_inflight.py > gpurun_out/inflight.log 2>&1; tail -3 gpurun_out/inflight.log
