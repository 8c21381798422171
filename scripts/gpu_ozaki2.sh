#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OCBO_OZAKI=1 timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q > gpurun_out/r02_oz_pytest_sweep.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_oz_pytest_sweep.log
OCBO_OZAKI=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-ladder-variant --sweep-trials 0 --no-one-thread > gpurun_out/r02_bench_oz2.json 2> gpurun_out/r02_bench_oz2.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_oz2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('e2e', {}).get('value'), d.get('parity_checked'), d.get('clocks'))
print(d.get('parity'))
PY
