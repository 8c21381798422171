"""One split + fused update of the rank-2048 trailing update shape (for ncu):  python tests/tools/ozaki_once.py [rows kw]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from ocbo_b200._lib import call, LIB

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 6144
kw = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device('cuda')
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: ctypes.c_void_p(t.data_ptr())
torch.manual_seed(0)
P = torch.randn(rows, kw, dtype=torch.float64, device=dev) * torch.logspace(0, -6, kw, dtype=torch.float64, device=dev)
C = torch.zeros(rows, rows, dtype=torch.float64, device=dev)
nb = int(LIB.ocbo_ozaki_ws_bytes(rows, kw, 1))
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
for _ in range(reps):
    call('ocbo_probe_ozaki_update', p(P), kw, rows, kw, p(C), rows, p(ws), nb, 1, st)
torch.cuda.synchronize()
print('ok')
