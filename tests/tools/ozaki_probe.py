"""Go / no-go probe: the Cholesky's rank-2048 update C = P P^T in fp64 through the INT8 tensor cores
(Ozaki-style error-free splitting), next to the DMMA kernel that does it today.  

  P (m x k fp64) = 2^e_i * sum_t 2^(-w t) p_t,   p_t int8 slices (row-wise exponent e_i, w bits each)
  P P^T          = 2^(e_i + e_j) * sum_g 2^(-w g) [ sum_{t+u=g} p_t p_u^T ],   g = 2 .. s+1
                   (slice pairs beyond g = s+1 dropped: below 2^-(w (s+1)) relative to the row maxima)
The pairs of one g are ONE integer GEMM with the slices concatenated along k (exact int32 accumulation:
8 * 2048 * 2^12 < 2^31), so s int32 products are formed instead of s (s+1) / 2.

The integer GEMMs here are cuBLASLt's (torch._int_mm) -- an UPPER bound on what a hand-written
tcgen05.mma kind::i8 kernel of ours would reach, and enough to decide: adopt only if the whole update
(split + GEMMs + fp64 accumulation) is >= 1.5x the DMMA kernel at <= 1e-13 relative error.

    python tests/tools/ozaki_probe.py > gpurun_out/r02_ozaki_probe.json
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def timed(torch, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def split(torch, P, s, w):
    """Row-wise scaled int8 slices of P: returns (exponents e [m], slices [s, m, k] int8)."""
    amax = P.abs().amax(dim=1).clamp_min(1e-300)
    e = torch.ceil(torch.log2(amax)) + 1.0            # |P / 2^e| < 1/2: every slice fits in w+1 signed bits
    r = P * torch.exp2(-e)[:, None]
    out = torch.empty((s,) + P.shape, dtype=torch.int8, device=P.device)
    for t in range(s):
        r = r * float(2 ** w)
        v = torch.trunc(r)
        out[t] = v.to(torch.int8)
        r = r - v
    return e, out


def main():
    import torch
    from ocbo_b200._lib import call
    dev = torch.device('cuda')
    torch.manual_seed(0)
    m, k = 6144, 2048
    # a Cholesky-panel-like operand: columns of decaying magnitude, rows of varying scale
    P = torch.randn(m, k, dtype=torch.float64, device=dev) * torch.logspace(0, -6, k, dtype=torch.float64, device=dev)
    P = P * torch.exp(torch.randn(m, 1, dtype=torch.float64, device=dev))
    ref = P @ P.t()                                                         # cuBLAS DGEMM
    C = torch.zeros(m, m, dtype=torch.float64, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    # the kernel in use today, on the same product (full square through the probe entry: 2 m^2 k flops)
    def dmma():
        # ocbo_probe_dgemm_nt is square-only: time the m x m x k product as (m/k) k-chunks of an n=m probe
        # is not possible; use torch's DGEMM-equivalent shape for the reference rate instead
        return P @ P.t()
    ms_dgemm, _ = timed(torch, dmma)
    out = {'shape': 'C = P P^T, P %d x %d fp64 (the first rank-2048 trailing update of an m = 8192 factorisation)' % (m, k),
           'fp64_full_square_ms': ms_dgemm, 'fp64_tflops': 2.0 * m * m * k / (ms_dgemm * 1e-3) / 1e12,
           'note_fp64': 'cuBLAS DGEMM on the full square; the library\'s own DMMA kernel does the lower triangle only '
                        '(half the flops) at 0.96 of this rate'}
    res = {}
    for s, w in ((7, 7), (8, 7), (9, 6), (8, 6)):
        def run_split():
            return split(torch, P, s, w)
        ms_split, (e, sl) = timed(torch, run_split, reps=2)

        def gemms():
            acc = []
            for g in range(2, s + 2):                    # g = t + u, 1-based slice indices
                ts = [t for t in range(1, s + 1) if 1 <= g - t <= s]
                A = torch.cat([sl[t - 1] for t in ts], dim=1)                  # m x (cnt k)
                B = torch.cat([sl[g - t - 1] for t in ts], dim=1)              # m x (cnt k)
                acc.append(torch._int_mm(A, B.t()))
            return acc

        def gemms_only(prep):
            return [torch._int_mm(A, Bt) for A, Bt in prep]
        prep = []
        for g in range(2, s + 2):
            ts = [t for t in range(1, s + 1) if 1 <= g - t <= s]
            prep.append((torch.cat([sl[t - 1] for t in ts], dim=1).contiguous(),
                         torch.cat([sl[g - t - 1] for t in ts], dim=1).contiguous().t()))
        ms_gemm, G = timed(torch, lambda: gemms_only(prep), reps=2)
        int_ops = sum(2.0 * m * m * A.shape[1] for A, _ in prep)

        def accumulate():
            acc = torch.zeros(m, m, dtype=torch.float64, device=dev)
            for i, g in enumerate(range(2, s + 2)):
                acc.add_(G[i].to(torch.float64), alpha=float(2.0 ** (-w * g)))
            return acc * torch.exp2(e)[:, None] * torch.exp2(e)[None, :]
        ms_acc, Coz = timed(torch, accumulate, reps=2)
        err = float((Coz - ref).abs().max() / ref.abs().max())
        # element-wise relative to |P||P|^T (the bound a backward-stable fp64 product satisfies)
        bound = P.abs() @ P.abs().t()
        err_cw = float(((Coz - ref).abs() / bound).max())
        total = ms_split + ms_gemm + ms_acc
        res['s%d_w%d' % (s, w)] = {
            'slices': s, 'bits_per_slice': w, 'int8_gemms': len(prep), 'int8_tops': int_ops / (ms_gemm * 1e-3) / 1e12,
            'ms_split': ms_split, 'ms_int8_gemms': ms_gemm, 'ms_fp64_accumulate': ms_acc, 'ms_total': total,
            'fp64_equivalent_tflops_gemms_only': 2.0 * m * m * k / (ms_gemm * 1e-3) / 1e12,
            'fp64_equivalent_tflops_total': 2.0 * m * m * k / (total * 1e-3) / 1e12,
            'speedup_vs_fp64_gemms_only': ms_dgemm / ms_gemm, 'speedup_vs_fp64_total': ms_dgemm / total,
            'max_rel_err_vs_dgemm': err, 'max_err_over_absP_absPt': err_cw}
        del prep, G, Coz, sl
        torch.cuda.empty_cache()
    out['ozaki'] = res
    best = max(res.values(), key=lambda r: r['speedup_vs_fp64_total'] if r['max_rel_err_vs_dgemm'] <= 1e-13 else 0.0)
    out['decision'] = {
        'rule': 'adopt only if >= 1.5x the fp64 tensor path at <= 1e-13 relative error (whole update, not GEMMs only)',
        'best_accurate_variant_speedup_total': best['speedup_vs_fp64_total'] if best['max_rel_err_vs_dgemm'] <= 1e-13 else None,
        'go': bool(best['max_rel_err_vs_dgemm'] <= 1e-13 and best['speedup_vs_fp64_total'] >= 1.5)}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
