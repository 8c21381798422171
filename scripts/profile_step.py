"""One sweep step (1 stream, batch 1) for ncu: 1 warm-up run + 1 profiled run."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocbo_b200 import sweep
step = sweep.SweepStep(1024, 64, 128, 256, 6, batch=1)
step.set_inputs_device([0], [sweep.make_trial_inputs(0)])
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(runs):
    step.run()
torch.cuda.synchronize()
t, a, g = step.result_host() if False else (None, None, None)
print('ok')
