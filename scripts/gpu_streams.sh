#!/bin/bash
cd "$(dirname "$0")/.."
for cfg in "8 2" "4 4" "6 3" "8 3" "12 2" "16 1" "8 2"; do
  set -- $cfg
  timeout 300 python bench.py --steps 8 --warmup 3 --streams $1 --batch $2 --no-ladder-variant --sweep-trials 0 --no-one-thread --no-cpu-baseline --no-peaks 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams $1 batch $2', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
