"""Correctness of the TMA main-loop variants (OCBO_TMA_FREE = 0 CTA barrier per k-tile, 1 free-running
warps) and of the cp.async path (OCBO_TMA=0): the library's DGEMM probe
against torch.matmul, and one m = 4096 Cholesky's backward error.  One process per mode."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def child():
    import torch
    from ocbo_b200 import engine
    from ocbo_b200._lib import call
    torch.manual_seed(0)
    dev = torch.device('cuda')
    for n in (1024, 4096):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        c = torch.zeros(n, n, dtype=torch.float64, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for rep in range(3):
            call('ocbo_probe_dgemm_nt', ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                 ctypes.c_void_p(c.data_ptr()), n, st)
        torch.cuda.synchronize()
        ref = a @ b.t()
        print('mode %s dgemm n=%d max rel err %.3e' % (os.environ.get('OCBO_TMA_FREE', 'default'), n,
                                                       float((c - ref).abs().max() / ref.abs().max())))
    m = 4096
    g = torch.randn(m, m + 8, dtype=torch.float64, device=dev)
    S = g @ g.t() / m + 0.1 * torch.eye(m, dtype=torch.float64, device=dev)
    L = S.clone().unsqueeze(0)
    invd = torch.empty((1, m // 128, 128, 128), dtype=torch.float64, device=dev)
    info = torch.zeros((1,), dtype=torch.int32, device=dev)
    engine.potrf(L, invd, info)
    torch.cuda.synchronize()
    Lt = torch.tril(L[0])
    print('mode %s potrf m=%d info %d backward err %.3e' % (os.environ.get('OCBO_TMA_FREE', 'default'), m,
          int(info.item()), float((Lt @ Lt.t() - S).abs().max() / S.abs().max())))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        child()
    else:
        for mode in sys.argv[1:] or ['0', '1']:
            subprocess.run([sys.executable, os.path.abspath(__file__), 'child'],
                           env=dict(os.environ, OCBO_TMA_FREE=mode, OCBO_POTRF_LOOKAHEAD='0'), check=False, timeout=300)
