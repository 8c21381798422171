"""ocbo_probe_ozaki_update (csrc/ozaki.cu: split into int8 slices + fused tcgen05.mma kind::i8 update with the fp64
recombination in the TMEM epilogue) against cuBLAS DGEMM: error on random and Cholesky-panel-like operands, then the
time of the rank-2048 trailing update of an m = 8192 factorisation (6144 x 2048, lower tiles).
    python tests/tools/ozaki_fused_probe.py > gpurun_out/r02_ozaki_fused.json"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from ocbo_b200._lib import call, LIB
    dev = torch.device('cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    out = {'checks': [], 'rates': []}
    torch.manual_seed(0)

    def run(P, C, tri):
        rows, kw = P.shape
        nb = int(LIB.ocbo_ozaki_ws_bytes(rows, kw, 1))
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        call('ocbo_probe_ozaki_update', p(P), P.stride(0), rows, kw, p(C), C.stride(0), p(ws), nb, tri, st)
        return ws

    for (rows, kw, kind) in ((128, 64, 'randn'), (128, 128, 'randn'), (256, 256, 'randn'), (1024, 2048, 'randn'),
                             (1024, 2048, 'panel'), (2048, 1024, 'panel')):
        P = torch.randn(rows, kw, dtype=torch.float64, device=dev)
        if kind == 'panel':
            P = P * torch.logspace(0, -6, kw, dtype=torch.float64, device=dev)
            P = P * torch.exp(torch.randn(rows, 1, dtype=torch.float64, device=dev))
        C0 = torch.randn(rows, rows, dtype=torch.float64, device=dev)
        for tri in (0, 1):
            C = C0.clone()
            run(P, C, tri)
            torch.cuda.synchronize()
            ref = C0 - P @ P.t()
            bound = P.abs() @ P.abs().t()
            if tri:
                # tiles on and below the diagonal (128 x 64 tiles: tile (ti, tj) with tj <= 2 ti + 1)
                i = torch.arange(rows, device=dev)
                mask = (i[None, :] // 64) <= 2 * (i[:, None] // 128) + 1
                untouched = bool((C[~mask] == C0[~mask]).all().item())
            else:
                mask = torch.ones_like(C, dtype=torch.bool)
                untouched = True
            err = float(((C - ref).abs()[mask] / bound[mask]).max().item())
            out['checks'].append({'rows': rows, 'kw': kw, 'kind': kind, 'tri': tri, 'max_err_over_absP_absPt': err,
                                  'tiles_above_diagonal_untouched': untouched})
            sys.stderr.write('ozaki %dx%d %s tri=%d err %.3g untouched %s\n' % (rows, kw, kind, tri, err, untouched))
    rows, kw = 6144, 2048
    P = torch.randn(rows, kw, dtype=torch.float64, device=dev) * torch.logspace(0, -6, kw, dtype=torch.float64, device=dev)
    C = torch.zeros(rows, rows, dtype=torch.float64, device=dev)
    nb = int(LIB.ocbo_ozaki_ws_bytes(rows, kw, 1))
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    ms_oz = timed(lambda: call('ocbo_probe_ozaki_update', p(P), kw, rows, kw, p(C), rows, p(ws), nb, 1, st))
    ms_lib = timed(lambda: torch.mm(P, P.t()))
    flops_tri = 2.0 * kw * (rows * (rows + 128) / 2.0)  # lower 128 x 64 tiles incl. the diagonal blocks
    out['rates'].append({'rows': rows, 'kw': kw, 'ozaki_split_plus_update_ms': ms_oz,
                         'ozaki_fp64_equivalent_tflops': flops_tri / ms_oz / 1e9,
                         'cublas_dgemm_full_square_ms': ms_lib,
                         'cublas_dgemm_tflops': 2.0 * rows * rows * kw / ms_lib / 1e9,
                         'int8_tops': float(LIB.ocbo_ozaki_slice_pairs()) * flops_tri / ms_oz / 1e9})
    sys.stderr.write('ozaki 6144x2048 lower: %.3f ms (%.1f TF/s fp64-equivalent); cuBLAS full square %.3f ms\n'
                     % (ms_oz, flops_tri / ms_oz / 1e9, ms_lib))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
