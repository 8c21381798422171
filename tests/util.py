"""Shared fixtures for the parity tests: seeded synthetic GP problems."""
import numpy as np


def hartmann6(x):
    A = np.asarray([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14],
                    [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14]])
    P = 1e-4 * np.asarray([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                           [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]])
    alpha = np.asarray([1.0, 1.2, 3.0, 3.2])
    x = np.atleast_2d(x)
    e = -np.einsum('ij,nij->ni', A, (x[:, None, :] - P[None]) ** 2)
    return (alpha * np.exp(e)).sum(axis=1)


def problem(seed, n, D, noise=1e-2, bw=0.25, scale=1.0, mu0=0.0):
    rs = np.random.RandomState(seed)
    X = rs.uniform(size=(n, D))
    if D == 6:
        y = hartmann6(X) + rs.normal(scale=1e-2, size=n)
    else:
        y = np.sin(3.0 * X.sum(axis=1)) + rs.normal(scale=1e-2, size=n)
    return X, y, dict(scale=scale, bandwidths=np.full(D, bw), noise_var=noise, mu0=mu0)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
