#!/bin/bash
# One point of the scaling series: bash scripts/gpu_scale.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench$N.log 2> gpurun_out/bench$N.err
echo "bench$N exit $?"; tail -c 400 gpurun_out/bench$N.log | head -c 300; echo
