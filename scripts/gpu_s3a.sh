#!/bin/bash
# session 3, call a: parity tests + smoke + short bench (gpu_check), REVI decisions/s, trials-in-flight sweep
bash scripts/gpu_check.sh
python -c "
import json
from tests.tools import bench_small
print(json.dumps(bench_small.measure_revi(1)))" > gpurun_out/revi.log 2>&1; tail -2 gpurun_out/revi.log
timeout 600 python tests/tools/bench_inflight.py > gpurun_out/inflight.log 2>&1; tail -60 gpurun_out/inflight.log
