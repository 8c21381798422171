"""Whole-loop BO iterations/s with P driver PROCESSES sharing one GPU (each with R trials in
flight): in flight a loop is bound by the per-trial Python of the strategy code, which only more
processes can run in parallel; their device passes time-slice the GPU (or share it under MPS).
Usage: python tests/tools/bench_procs.py [set2d|rand6d|jointbran] [P ...]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHILD = r'''
import os, sys, time
sys.path.insert(0, %(root)r)
import torch
from ocbo_b200 import %(drv)s as drv
argv = %(argv)r
drv.run(argv=argv); torch.cuda.synchronize()
open(%(ready)r + str(os.getpid()), 'w').close()
while not os.path.exists(%(go)r):
    time.sleep(0.002)
t0 = time.perf_counter(); drv.run(argv=argv); torch.cuda.synchronize()
print('WALL %%.6f' %% (time.perf_counter() - t0))
'''


def measure(name, P, R=8, capital=40):
    from tests.tools.profile_driver import OPTS
    kind, text = OPTS[name]
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, 'o.txt')
        with open(path, 'w') as fh:
            fh.write(text % capital)
        argv = ['--options', path, '--trials', str(R), '--trials_in_flight', str(R)]
        ready, go = os.path.join(tmp, 'ready_'), os.path.join(tmp, 'go')
        code = CHILD % dict(root=ROOT, drv='ocbo' if kind == 'd' else 'cts_ocbo', argv=argv, ready=ready, go=go)
        env = dict(os.environ, OMP_NUM_THREADS='1', OPENBLAS_NUM_THREADS='1', MKL_NUM_THREADS='1')
        for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'LOCAL_WORLD_SIZE', 'MASTER_ADDR', 'MASTER_PORT'):
            env.pop(k, None)  # independent processes: no launcher, no process group
        procs = [subprocess.Popen([sys.executable, '-c', code], stdout=subprocess.PIPE, text=True, env=env)
                 for _ in range(P)]
        t_end = time.time() + 300
        while sum(f.startswith('ready_') for f in os.listdir(tmp)) < P and time.time() < t_end:
            time.sleep(0.01)
        open(go, 'w').close()
        walls = []
        for p in procs:
            out = p.communicate(timeout=300)[0]
            walls += [float(l.split()[1]) for l in out.splitlines() if l.startswith('WALL')]
        if len(walls) != P:
            return {'error': 'only %d of %d processes finished' % (len(walls), P)}
        return {'processes': P, 'trials_in_flight_each': R, 'iterations_per_s': P * R * capital / max(walls),
                'slowest_wall_s': max(walls), 'fastest_wall_s': min(walls)}


if __name__ == '__main__':
    names = [a for a in sys.argv[1:] if not a.isdigit()] or ['set2d', 'rand6d', 'jointbran']
    Ps = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 2, 4, 8]
    for n in names:
        for P in Ps:
            print(n, json.dumps(measure(n, P)), flush=True)
