"""numpy restatement of the knowledge-gradient pieces behind REVI.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PINNED: unlike the Dragonfly
arithmetic, ``knowledge_gradient`` is in-tree (/root/reference/src/util/misc_util.py:180-212,
plain numpy + scipy.stats.norm), so this restatement is checked bit for bit against the
reference's own function (tests/golden/make_kg_golden.py -> tests/golden/kg_vectors.json,
tests/test_oracle.py).
"""
import numpy as np
from scipy.special import ndtr

_PDF_C = np.sqrt(2.0 * np.pi)


def norm_cdf(z):
    """scipy.stats.norm.cdf is scipy.special.ndtr (cephes: with x = z / sqrt 2,
    0.5 + 0.5 erf(x) for |x| < sqrt(1/2), else 0.5 erfc(|x|) reflected)."""
    return float(ndtr(z))


def norm_pdf(z):
    """scipy.stats.norm.pdf: exp(-z^2 / 2) / sqrt(2 pi)."""
    return float(np.exp(-z ** 2 / 2.0) / _PDF_C)


def knowledge_gradient(a, b):
    """misc_util.py:180-212: E[max_i (a_i + b_i Z)] - max_i a_i for Z ~ N(0, 1):
    sort the lines by slope, build the upper envelope with a stack, sum the segments."""
    with np.errstate(divide='ignore', invalid='ignore'):
        pairs = sorted(zip(np.asarray(b, dtype=np.float64).tolist(), np.asarray(a, dtype=np.float64).tolist()))
        b = np.asarray([p[0] for p in pairs])
        a = np.asarray([p[1] for p in pairs])
        a = a - np.max(a)
        idxs = [0, 1]
        zs = [float('-inf'), (a[0] - a[1]) / (b[1] - b[0])]
        for i in range(2, len(a)):
            j = idxs[-1]
            z = (a[i] - a[j]) / (b[j] - b[i])
            while z < zs[-1]:
                idxs.pop()
                zs.pop()
                j = idxs[-1]
                z = (a[i] - a[j]) / (b[j] - b[i])
            idxs.append(i)
            zs.append(z)
        zs.append(float('inf'))
        result = 0.0
        cdfs = [norm_cdf(float(z)) for z in zs]
        pdfs = [norm_pdf(float(z)) for z in zs]
        for i, idx in enumerate(idxs):
            result += a[idx] * (cdfs[i + 1] - cdfs[i]) + b[idx] * (pdfs[i] - pdfs[i + 1])
        return float(result)


def revi_improvements(judge_means, interactions, cand_vars, noise):
    """corr_opt.py:103-119 / postmax_cts.py:147-160: for every candidate c the sum over the
    judgement groups g of KG(judge_means[g], interactions[c, g] / sqrt(noise + var_c)),
    accumulated in group order."""
    G, E = judge_means.shape
    out = np.zeros(len(interactions))
    for c in range(len(interactions)):
        inter = np.asarray(interactions[c]).reshape(G, E)
        imp = 0
        for g in range(G):
            imp += knowledge_gradient(judge_means[g], inter[g] / np.sqrt(noise + cand_vars[c]))
        out[c] = imp
    return out
