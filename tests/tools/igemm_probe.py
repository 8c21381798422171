"""ocbo_probe_igemm_nt (hand-written tcgen05.mma kind::i8 + TMA + TMEM, csrc/igemm_probe.cu) against
torch._int_mm (cuBLASLt): exact int32 results, then the rate on the Ozaki-concatenated shape of a
rank-2048 update (6144 x 6144 x 16384).
    python tests/tools/igemm_probe.py > gpurun_out/r02_igemm_probe.json"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from ocbo_b200._lib import call
    dev = torch.device('cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    out = {'checks': [], 'rates': []}
    g = torch.Generator(device='cuda').manual_seed(0)
    for (m, n, k) in ((128, 128, 128), (256, 384, 512), (1024, 1024, 2048), (3072, 3072, 4096)):
        A = torch.randint(-127, 128, (m, k), dtype=torch.int8, device=dev, generator=g)
        B = torch.randint(-127, 128, (n, k), dtype=torch.int8, device=dev, generator=g)
        C = torch.full((m, n), -7, dtype=torch.int32, device=dev)
        call('ocbo_probe_igemm_nt', p(A), k, p(B), k, p(C), n, m, n, k, st)
        torch.cuda.synchronize()
        ref = torch._int_mm(A, B.t())
        bad = int((C != ref).sum().item())
        out['checks'].append({'m': m, 'n': n, 'k': k, 'mismatches': bad})
        sys.stderr.write('igemm %dx%dx%d mismatches %d\n' % (m, n, k, bad))
        if bad:
            print(json.dumps(out))
            return
    for (m, n, k) in ((6144, 6144, 16384), (8192, 8192, 8192)):
        A = torch.randint(-127, 128, (m, k), dtype=torch.int8, device=dev, generator=g)
        B = torch.randint(-127, 128, (n, k), dtype=torch.int8, device=dev, generator=g)
        C = torch.empty((m, n), dtype=torch.int32, device=dev)

        def timed(fn):
            fn()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            return best
        ms_own = timed(lambda: call('ocbo_probe_igemm_nt', p(A), k, p(B), k, p(C), n, m, n, k, st))
        ms_lib = timed(lambda: torch._int_mm(A, B.t()))
        ops = 2.0 * m * n * k
        out['rates'].append({'m': m, 'n': n, 'k': k, 'own_tops': ops / ms_own / 1e9, 'cublaslt_tops': ops / ms_lib / 1e9,
                             'own_ms': ms_own, 'cublaslt_ms': ms_lib})
        sys.stderr.write('igemm %dx%dx%d own %.0f TOP/s  cuBLASLt %.0f TOP/s\n' % (m, n, k, ops / ms_own / 1e9, ops / ms_lib / 1e9))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
