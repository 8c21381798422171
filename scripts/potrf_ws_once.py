"""One chol(Sigma)-sized factorisation through ocbo_potrf_ws (m = 8192, batch 1, no look-ahead) for ncu:
python scripts/potrf_ws_once.py [runs]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocbo_b200._lib import call, LIB

m = 8192
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device('cuda')
torch.manual_seed(0)
G = torch.randn(m, 1024, dtype=torch.float64, device=dev)
A0 = G @ G.t() / 1024.0 + torch.eye(m, dtype=torch.float64, device=dev) * 1e-2
invd = torch.empty(m // 128, 128, 128, dtype=torch.float64, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
nb = int(LIB.ocbo_potrf_ws_bytes(m, 1))
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(runs):
    A = A0.clone()
    call('ocbo_potrf_ws', p(A), m, m, 0, p(invd), p(info), 1, p(ws), nb, st)
torch.cuda.synchronize()
assert int(info.item()) == 0
print('ok')
