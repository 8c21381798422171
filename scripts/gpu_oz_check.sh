#!/bin/bash
cd "$(dirname "$0")/.."
timeout 200 python tests/tools/ozaki_fused_probe.py 2>&1 >/dev/null | tail -3
timeout 600 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_sweep.py -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-ladder-variant --sweep-trials 0 --no-one-thread "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['parity'] and (d['parity']['ok'], d['parity']['max_rel_err_samples']), d['roofline']['potrf_ms'], d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['by_launch_class']['posterior_covariance_k1024'], d['clocks'])"
