#!/bin/bash
# Round-2 artefacts with the int8-tensor-core path as default: parity tests, smoke, default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_checked'], d['clocks'])
r = d['roofline']
print({k: r[k] for k in ('achieved', 'peak', 'unit', 'frac', 'cublaslt_int8_tops_this_run', 'fp64_equivalent_tflops') if k in r})
print(r.get('potrf_ms')); print(r.get('by_launch_class'))
print(d.get('sweep_fixed_trials')); print(d.get('ladder_variant'))
PY
