#!/bin/bash
# Trials in flight on the int8-tensor-core path: streams x batch (the round-2a sweep of these two was made on the
# fp64 path, whose step is 2.3x longer per launch).  Prints samples/s, ms per step and the host's issue time per step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  tag=s$1_b$2
  timeout 300 python bench.py --steps 8 --warmup 3 --streams $1 --batch $2 --no-parity --no-ladder-variant --sweep-trials 0 \
      --no-one-thread --no-cpu-baseline --no-peaks > gpurun_out/r02_sb_$tag.json 2> gpurun_out/r02_sb_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/r02_sb_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'host issue', round(d['run']['host_issue_ms_per_step'], 2),
          'launches/step', round(d['run']['launches_per_step']), d.get('clocks', {}).get('sm_mhz'), d.get('clocks', {}).get('reasons'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
for sb in "$@"; do run ${sb%x*} ${sb#*x}; done
