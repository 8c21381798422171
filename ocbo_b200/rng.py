"""Source of the standard normals behind ``draw_sample``.

The reference draws ``np.random.normal(size=(m, S))`` from the GLOBAL numpy
stream inside Dragonfly's draw_gaussian_samples, so the
default here does exactly that on the host and ships the (m x S) block to the
GPU: a seeded run consumes the global stream like the reference does.  Tests
inject identical normals into the oracle and this path through ``NORMAL_SOURCE``;
the sweep path uses the device Philox generator instead (engine.philox_normals).
"""
import numpy as np

NORMAL_SOURCE = None  # callable(size) -> ndarray, or None for np.random.normal


def normals(size):
    if NORMAL_SOURCE is not None:
        return np.asarray(NORMAL_SOURCE(size), dtype=np.float64)
    return np.random.normal(size=size)
