#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env OCBO_OZAKI=1 OCBO_POTRF_OUTER=2048 OCBO_POTRF_MID=512 timeout 300 python bench.py --steps 8 --warmup 3 --no-ladder-variant --sweep-trials 0 --no-one-thread --no-cpu-baseline --no-peaks "$@" > gpurun_out/r02_oztune_$tag.json 2> gpurun_out/r02_oztune_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/r02_oztune_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'], 2), d.get('clocks', {}).get('sm_mhz'), d.get('clocks', {}).get('reasons'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
run s1b1 --streams 1 --batch 1
run s1b2 --streams 1 --batch 2
run s2b2 --streams 2 --batch 2
run s4b2 --streams 4 --batch 2
run s4b4 --streams 4 --batch 4
run s8b1 --streams 8 --batch 1
run s12b2 --streams 12 --batch 2
run s16b1 --streams 16 --batch 1
