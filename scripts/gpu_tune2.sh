#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
: > gpurun_out/tune.log
for inner in left right; do
for outer in 512 1024 2048 4096 8192; do
  for sb in "8 1" "8 2"; do
    set -- $sb
    echo "== inner=$inner outer=$outer streams=$1 batch=$2" >> gpurun_out/tune.log
    OCBO_POTRF_INNER=$inner OCBO_POTRF_OUTER=$outer timeout 300 python bench.py --steps 5 --warmup 3 --streams $1 --batch $2 \
      --no-cpu-baseline --no-peaks >> gpurun_out/tune.log 2>&1
  done
done
done
python - <<'PY'
import json
cur=None
for line in open('gpurun_out/tune.log'):
    if line.startswith('=='): cur=line.strip()
    elif line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print(cur, 'samples/s %.0f  ms/trial %.2f  potrf_ms %s' % (d['value'], d['ms_per_step']/d['config']['trials_per_step_per_gpu'], {k:round(v,2) for k,v in r['potrf_ms'].items()}))
    elif 'rror' in line: print(line.strip()[:200])
PY
