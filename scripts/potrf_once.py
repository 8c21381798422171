"""One Cholesky of a seeded SPD m x m matrix (default 8192) through ocbo_potrf, for ncu:
with OCBO_POTRF_OUTER=2048 the 16th gemm_nt_kernel<128,64> launch is the first rank-2048
trailing update (15 in-panel updates precede it)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocbo_b200 import engine
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator(device='cuda').manual_seed(0)
G = torch.randn((m, m // 4), dtype=torch.float64, device='cuda', generator=g)
A = (G @ G.t() / (m // 4) + torch.eye(m, dtype=torch.float64, device='cuda')).unsqueeze(0).contiguous()
invd = torch.empty((1, m // 128, 128, 128), dtype=torch.float64, device='cuda')
info = torch.zeros((1,), dtype=torch.int32, device='cuda')
engine.potrf(A, invd, info)
torch.cuda.synchronize()
print('info', int(info.item()))
