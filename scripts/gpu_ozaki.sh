#!/bin/bash
# Ozaki (int8 tensor core) trailing updates: headline parity test + bench with and without
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OCBO_OZAKI=1 timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q > gpurun_out/r02_oz_pytest_sweep.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_oz_pytest_sweep.log
OCBO_OZAKI=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-ladder-variant --sweep-trials 0 --no-one-thread > gpurun_out/r02_bench_oz1.json 2> gpurun_out/r02_bench_oz1.err
echo "bench oz1 rc=$?"
OCBO_OZAKI=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-ladder-variant --sweep-trials 0 --no-one-thread --no-cpu-baseline --no-parity > gpurun_out/r02_bench_oz0.json 2> gpurun_out/r02_bench_oz0.err
echo "bench oz0 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_oz1.json', 'gpurun_out/r02_bench_oz0.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d.get('e2e', {}).get('value'), d.get('parity_checked'), d.get('clocks'))
    except Exception as e:
        print(f, 'ERR', e)
PY
