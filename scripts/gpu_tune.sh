#!/bin/bash
# Parity tests, then a sweep of the Cholesky outer block and the stream count.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
: > gpurun_out/tune.log
for outer in ${OUTERS:-512 1024 2048}; do
  for streams in ${STREAMS:-1 4 8}; do
    echo "== outer=$outer streams=$streams" >> gpurun_out/tune.log
    OCBO_POTRF_OUTER=$outer timeout 300 python bench.py --steps 4 --warmup 3 --streams $streams \
      --no-cpu-baseline ${PEAKS:---no-peaks} >> gpurun_out/tune.log 2>&1
  done
done
python - <<'PY'
import json
cur=None
for line in open('gpurun_out/tune.log'):
    if line.startswith('=='): cur=line.strip()
    elif line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print(cur, 'samples/s %.0f  ms/trial %.2f  potrf_ms %s upd TF %.1f' % (d['value'], d['ms_per_step']/d['config']['trials_per_step_per_gpu'], {k:round(v,2) for k,v in r['potrf_ms'].items()}, r['achieved'] or 0), r.get('fp64_probes_tflops'))
    elif 'rror' in line: print(line.strip()[:200])
PY
