#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/mid.log
for cfg in "2048 0" "2048 512" "2048 1024" "4096 512" "4096 1024" "8192 1024" "8192 512" "2048 256"; do
  set -- $cfg
  echo "== outer=$1 mid=$2" >> gpurun_out/mid.log
  OCBO_POTRF_OUTER=$1 OCBO_POTRF_MID=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-peaks 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step']/16, d['e2e']['picks_checksum'], d['roofline']['potrf_ms'])" >> gpurun_out/mid.log
done
cat gpurun_out/mid.log
