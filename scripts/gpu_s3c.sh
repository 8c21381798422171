#!/bin/bash
# session 3, call c: full GPU suite after the staging rewrite, in-flight sweeps
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tests/tools/bench_inflight.py > gpurun_out/inflight.log 2>&1; tail -3 gpurun_out/inflight.log
timeout 600 python -c "
import json
from tests.tools import bench_small
print(json.dumps(bench_small.measure_bo_loop()))" > gpurun_out/bo_loop.log 2>&1; tail -3 gpurun_out/bo_loop.log
