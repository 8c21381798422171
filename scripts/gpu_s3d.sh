#!/bin/bash
# session 3, call d (2 GPUs): option-file driver sharded over ranks (NCCL gather of trial results), bench at N=2
mkdir -p gpurun_out /tmp/w2
cat > /tmp/w2/opt.txt <<EOT
--run_id w2
--trials 6
--max_capital 8
--tune_every 1
--tuning_methods ml
--methods mts
--functions branin,parabaloid,parabaloid
--write_dir /tmp/w2/out
EOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    -m ocbo_b200.ocbo --options /tmp/w2/opt.txt --trials_in_flight 2 > gpurun_out/driver_w2.log 2>&1
echo "driver w2 exit $?"; ls /tmp/w2/out/w2 | tr '\n' ' '; echo
sed -i 's#/tmp/w2/out#/tmp/w2/out1#' /tmp/w2/opt.txt
timeout 300 python -m ocbo_b200.ocbo --options /tmp/w2/opt.txt --trials_in_flight 3 > gpurun_out/driver_w1.log 2>&1
echo "driver w1 exit $?"
python - <<EOT
import pickle, numpy as np
same = True
for t in range(6):
    a = pickle.load(open('/tmp/w2/out/w2/trial_%d.pkl' % t, 'rb'))['mts']
    b = pickle.load(open('/tmp/w2/out1/w2/trial_%d.pkl' % t, 'rb'))['mts']
    same &= a.choice_timeline == b.choice_timeline and a.total_best == b.total_best
print('2 ranks x 2 in flight == 1 rank x 3 in flight, all 6 trials:', same)
EOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench2.log 2> gpurun_out/bench2.err
echo "bench2 exit $?"; tail -c 1500 gpurun_out/bench2.log
