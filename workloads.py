"""Seeded synthetic workloads for bench.py and the parity tests.

Fixed *density* clouds (h0 = 1e-2 point spacing scale) so conditioning does not depend on n;
neighbourhoods = k nearest neighbours (self excluded) by ``scipy.spatial.cKDTree``, exactly as
the reference's examples build them (examples/expertsolver_example.py:59-66 -- ``hoods`` is
caller-side data, the library receives the gathered ``xk = x[hoods]``, ``fk = f[hoods]``).
"""
from __future__ import annotations

import numpy as np

H0 = 1e-2


def cloud(n: int, dim: int, seed: int = 42, unit_box: bool = False) -> np.ndarray:
    rng = np.random.default_rng(seed)
    scale = 1.0 if unit_box else H0 * n ** (1.0 / dim)
    x = scale * rng.random((n, dim))
    return x[:, 0].copy() if dim == 1 else x


def hoods_knn(x: np.ndarray, k: int) -> np.ndarray:
    from scipy.spatial import cKDTree
    x2 = x.reshape(len(x), -1)
    return cKDTree(x2).query(x2, k + 1, workers=-1)[1][:, 1:].astype(np.int32)


def field(x: np.ndarray) -> np.ndarray:
    """sin(pi x) cos(pi y) exp(z) (examples/wlsqm_example.py:65,288)."""
    x2 = x.reshape(len(x), -1)
    f = np.sin(np.pi * x2[:, 0])
    if x2.shape[1] >= 2:
        f = f * np.cos(np.pi * x2[:, 1])
    if x2.shape[1] >= 3:
        f = f * np.exp(x2[:, 2])
    return f


def field_step(f: np.ndarray, t: int) -> np.ndarray:
    """time step t of the synthetic stream: f cos(0.1 t) + 0.01 N(0,1), rng seeded 1000+t."""
    return f * np.cos(0.1 * t) + 0.01 * np.random.default_rng(1000 + t).standard_normal(len(f))


def dof_orders(dim: int, order: int) -> np.ndarray:
    """total derivative order of every DOF slot (defs.pyx slot order)."""
    counts = {1: [1, 1, 1, 1, 1], 2: [1, 2, 3, 4, 5], 3: [1, 3, 6, 10, 15]}[dim]
    return np.concatenate([np.full(counts[d], d) for d in range(order + 1)])


def parity_report(a: np.ndarray, b: np.ndarray, dim: int, order: int):
    """DOF-scaled error |a-b| / max_i |b[:, j]| grouped by derivative order -> {d: (p50,p99,max)}."""
    no = a.shape[1]
    scale = np.nanmax(np.abs(b), axis=0)
    scale[scale == 0] = 1.0
    e = np.abs(a - b) / scale
    d = dof_orders(dim, order)[:no]
    out = {}
    for dd in range(order + 1):
        ee = e[:, d == dd].ravel()
        ee = ee[np.isfinite(ee)]
        if ee.size:
            out[dd] = (float(np.median(ee)), float(np.quantile(ee, 0.99)), float(ee.max()))
    return out
