#!/bin/bash
cd "$(dirname "$0")/.."
run() {
  tag=$1; shift
  env "$@" timeout 300 python tests/tools/stage_times.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', 'whole', round(d['whole_step'],3), 'potrf', round(d['potrf_sigma'],3), 'postcov', round(d['post_cov'],3), 'V', round(d['gram_and_solve_for_V'],3), 'sample', round(d['sampling_tail'],3))"
}
run default X=1
run mid256_k256 OCBO_POTRF_MID=256 OCBO_OZAKI_MINK=256
run o2048_m512 OCBO_POTRF_OUTER=2048 OCBO_POTRF_MID=512
run o1024_m512_right OCBO_POTRF_INNER=right
run o512_m512 OCBO_POTRF_OUTER=512 OCBO_POTRF_MID=512
run default2 X=1
