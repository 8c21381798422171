#!/bin/bash
cd "$(dirname "$0")/.."
for rep in 1 2; do
for cfg in "1024 512" "1024 1024" "2048 1024" "2048 2048" "4096 1024"; do
  set -- $cfg
  OCBO_POTRF_OUTER=$1 OCBO_POTRF_MID=$2 timeout 300 python bench.py --steps 8 --warmup 3 --no-ladder-variant --sweep-trials 0 --no-one-thread --no-cpu-baseline --no-peaks 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('outer $1 mid $2', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
done
