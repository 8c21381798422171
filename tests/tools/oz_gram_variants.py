"""Diagnostic: where does the Gram epilogue of the covariance mode of the int8 kernel (ozaki_update_kernel<2>) spend
its time?  Builds variants of the library on the GPU box (-DOCBO_OZ_NOEXP: no exp chain, -DOCBO_OZ_NODOT: no
shared-memory reads of the row points) and times ocbo_post_cov_ws (m = 8192, n = 1024, batch 2) with each, next to
the shipped library with the Gram fused and unfused.   python tests/tools/oz_gram_variants.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CHILD = r'''
import ctypes, sys, os
sys.path.insert(0, %r)
import torch
from ocbo_b200 import sweep
from ocbo_b200._lib import call
from ocbo_b200.engine import _p
grp = sweep.SweepStep(1024, 64, 128, 256, 6, batch=2)
grp.set_inputs_device([0, 1], [sweep.make_trial_inputs(t) for t in (0, 1)])
grp.run(); torch.cuda.synchronize()
st = ctypes.c_void_p(grp.stream.cuda_stream)
def postcov():
    call('ocbo_post_cov_ws', grp.kind, 6, _p(grp.Xs), grp.m, grp.Xs.stride(0), None, _p(grp.V), grp.n, grp.m, grp.V.stride(0),
         _p(grp.theta), grp.theta.stride(0), _p(grp.Sigma), grp.m, grp.Sigma.stride(0), 1, 2, _p(grp.ws), grp.ws.numel(), st)
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(grp.stream); postcov(); e1.record(grp.stream); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print('%%.3f ms per ocbo_post_cov_ws (batch 2)' %% best)
''' % ROOT


def main():
    from ocbo_b200 import build
    runs = [('shipped, Gram fused', build.LIB_PATH, {}), ('shipped, Gram unfused', build.LIB_PATH, {'OCBO_OZAKI_FUSE_GRAM': '0'})]
    for tag, defs in (('no exp chain', ['OCBO_OZ_NOEXP=1']), ('no row-point reads', ['OCBO_OZ_NODOT=1']),
                      ('neither', ['OCBO_OZ_NOEXP=1', 'OCBO_OZ_NODOT=1'])):
        d = os.path.join(ROOT, 'gpurun_out', 'lib_' + tag.replace(' ', '_'))
        runs.append(('variant: ' + tag, build.build(defines=defs, out_dir=d), {}))
    for name, lib, env in runs:
        out = subprocess.run([sys.executable, '-c', CHILD], env=dict(os.environ, OCBO_B200_LIB=lib, **env),
                             capture_output=True, text=True)
        print('%-28s %s' % (name, (out.stdout.strip() or out.stderr.strip()[-300:])))
        sys.stdout.flush()


if __name__ == '__main__':
    main()
