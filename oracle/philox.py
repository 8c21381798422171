"""Philox4x32-10 + Box-Muller normals in numpy -- the definition the device RNG
(``ocbo_b200/csrc/philox.cuh``) must reproduce.  TEST INFRASTRUCTURE ONLY.

The reference draws ``np.random.normal(size=(m, S))`` on the host
(draw_gaussian_samples); the sweep configuration replaces
that with a counter-based generator keyed by (seed, trial) so that results do
not depend on the world size.  Layout contract:

  element e = i * S + s of the (m x S) row-major normal matrix Z
  pair    q = e >> 1                       (two normals per Philox call)
  counter   = (q & 0xffffffff, q >> 32, trial, 0x0CB0)
  key       = (seed & 0xffffffff, seed >> 32)
  u1 = (((r0 << 32 | r1) >> 11) + 0.5) * 2^-53,   u2 likewise from (r2, r3)
  z_even = sqrt(-2 ln u1) * cos(2 pi u2),  z_odd = sqrt(-2 ln u1) * sin(2 pi u2)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
CTR3 = 0x0CB0


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all inputs broadcastable uint32-valued arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, \
                         (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def normals(seed, trial, m, S):
    """(m x S) row-major standard normals per the layout contract above."""
    total = m * S
    npairs = (total + 1) // 2
    q = np.arange(npairs, dtype=np.uint64)
    r0, r1, r2, r3 = philox4x32_10(q & MASK, q >> np.uint64(32),
                                   np.uint64(trial & 0xFFFFFFFF), np.uint64(CTR3),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    a = ((r0 << np.uint64(32)) | r1) >> np.uint64(11)
    b = ((r2 << np.uint64(32)) | r3) >> np.uint64(11)
    u1 = (a.astype(np.float64) + 0.5) * 2.0 ** -53
    u2 = (b.astype(np.float64) + 0.5) * 2.0 ** -53
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.empty((npairs, 2), dtype=np.float64)
    z[:, 0] = rad * np.cos(2.0 * np.pi * u2)
    z[:, 1] = rad * np.sin(2.0 * np.pi * u2)
    return z.ravel()[:total].reshape(m, S)
