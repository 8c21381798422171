#!/bin/bash
# sweep step: more trials in flight than the default 8 streams x batch 2?
mkdir -p gpurun_out; : > gpurun_out/streams.log
for sb in "8 2" "12 1" "12 2" "16 1" "6 3"; do
  set -- $sb
  echo "== streams=$1 batch=$2" >> gpurun_out/streams.log
  timeout 300 python bench.py --steps 5 --warmup 3 --streams $1 --batch $2 --no-cpu-baseline --no-peaks 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step']/d['config']['trials_per_step_per_gpu'])" >> gpurun_out/streams.log
done
cat gpurun_out/streams.log
