#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" OCBO_OZAKI=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-ladder-variant --sweep-trials 0 --no-one-thread --no-cpu-baseline --no-peaks > gpurun_out/r02_oztune_$tag.json 2> gpurun_out/r02_oztune_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/r02_oztune_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'], 2), d.get('clocks', {}).get('sm_mhz'), d.get('clocks', {}).get('reasons'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
run sms148 OCBO_POTRF_OUTER=2048 OCBO_POTRF_MID=512
run sms144 OCBO_POTRF_OUTER=2048 OCBO_POTRF_MID=512 OCBO_OZAKI_SMS=144
run sms140 OCBO_POTRF_OUTER=2048 OCBO_POTRF_MID=512 OCBO_OZAKI_SMS=140
run sms132 OCBO_POTRF_OUTER=2048 OCBO_POTRF_MID=512 OCBO_OZAKI_SMS=132
run sms120 OCBO_POTRF_OUTER=2048 OCBO_POTRF_MID=512 OCBO_OZAKI_SMS=120
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ozaki_update -c 1 -o gpurun_out/r02_ozaki_v3 -f python tests/tools/ozaki_once.py > gpurun_out/ncu_oz3.log 2>&1; echo ncu rc=$?
