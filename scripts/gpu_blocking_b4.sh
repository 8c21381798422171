#!/bin/bash
# Panel widths / CTAs of the persistent kernel once more at 4 streams x 4 trials per launch (environment switches only).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-parity --no-ladder-variant --sweep-trials 0 --no-one-thread --no-cpu-baseline --no-peaks > gpurun_out/r02_b4_$tag.json 2> gpurun_out/r02_b4_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/r02_b4_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'], 2), d.get('clocks', {}).get('sm_mhz'), d.get('clocks', {}).get('reasons'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
run default X=1
run o2048_m512 OCBO_POTRF_OUTER=2048 OCBO_POTRF_MID=512
run o1024_m512 OCBO_POTRF_OUTER=1024 OCBO_POTRF_MID=512
run o4096_m1024 OCBO_POTRF_OUTER=4096 OCBO_POTRF_MID=1024
run sms144 OCBO_OZAKI_SMS=144
run inpanel256 OCBO_OZAKI_INPANEL=256
run default2 X=1
