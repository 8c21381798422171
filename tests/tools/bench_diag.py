"""Time the 128 x 128 diagonal-block kernel (factor + inverse) alone: one problem (the
latency that sits on the blocked Cholesky's serial chain) and one problem per SM.
Usage: python tests/tools/bench_diag.py            (runs every OCBO_DIAG_KERNEL variant)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from ocbo_b200 import engine
    rs = np.random.RandomState(0)
    for B, nvalid in ((1, 128), (148, 128), (1, 48), (5, 50)):
        G = rs.normal(size=(B, nvalid, nvalid + 8))
        Ah = np.tile(np.eye(128), (B, 1, 1))
        Ah[:, :nvalid, :nvalid] = G @ G.transpose(0, 2, 1) / nvalid + 0.1 * np.eye(nvalid)
        A0 = engine._h2d(Ah)
        nv = engine._h2d(np.full((B,), nvalid, dtype=np.int32), engine.I32)
        invd = torch.zeros((B, 1, 128, 128), dtype=torch.float64, device=A0.device)
        info = torch.zeros((B,), dtype=torch.int32, device=A0.device)
        reps = 50
        As = [A0.clone() for _ in range(reps + 5)]
        for i in range(5):
            engine.potrf(As[i], invd, info, nv)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            engine.potrf(As[5 + i], invd, info, nv)
        e1.record()
        torch.cuda.synchronize()
        assert int(info.abs().sum().item()) == 0
        print('%-5s batch %3d valid %3d: %.2f us per launch' % (
            os.environ.get('OCBO_DIAG_KERNEL', 'blk'), B, nvalid, e0.elapsed_time(e1) * 1e3 / reps))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        child()
    else:
        for variant in ('reg', 'blk'):
            subprocess.run([sys.executable, os.path.abspath(__file__), 'child'],
                           env=dict(os.environ, OCBO_DIAG_KERNEL=variant), check=False)
