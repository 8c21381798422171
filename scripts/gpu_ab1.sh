#!/bin/bash
# like gpu_ab.sh with extra bench arguments: VAR A B REPS -- args...
cd "$(dirname "$0")/.."
VAR=$1; A=$2; B=$3; REPS=$4; shift 4
for i in $(seq 1 $REPS); do
  for v in $A $B; do
    env $VAR=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-ladder-variant --sweep-trials 0 --no-one-thread --no-cpu-baseline --no-peaks "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
  done
done
