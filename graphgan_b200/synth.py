"""Synthetic benchmark inputs (SURVEY.md section 8d / BASELINE.json configs C2-C5).

Edge lists are produced in a fixed "file order" (the order matters: it is the adjacency order
the reference would read, src/utils.py:27-37), self-loops and duplicate undirected edges are
dropped so that the raw and the walk CSR coincide.
"""
import numpy as np


def _dedupe(a, b):
    keep = a != b
    a, b = a[keep], b[keep]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = lo.astype(np.int64) * (int(hi.max()) + 1 if hi.size else 1) + hi
    _, first = np.unique(key, return_index=True)
    first.sort()
    return np.stack([a[first], b[first]], axis=1)


def erdos_renyi(n, avg_deg, seed=0):
    """C2: N nodes, N*avg_deg/2 undirected edges with uniform endpoints."""
    rs = np.random.RandomState(seed)
    m = n * avg_deg // 2
    a = rs.randint(0, n, size=m)
    b = rs.randint(0, n, size=m)
    return _dedupe(a, b)


def power_law(n, avg_deg, gamma=2.5, seed=0):
    """C3/C4: Chung-Lu style, endpoints drawn proportionally to w_i = (i + 10)^(-1/(gamma-1))."""
    rs = np.random.RandomState(seed)
    m = n * avg_deg // 2
    w = (np.arange(n, dtype=np.float64) + 10.0) ** (-1.0 / (gamma - 1.0))
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    a = np.searchsorted(cdf, rs.random_sample(m), side="right").astype(np.int64)
    b = np.searchsorted(cdf, rs.random_sample(m), side="right").astype(np.int64)
    np.minimum(a, n - 1, out=a)
    np.minimum(b, n - 1, out=b)
    return _dedupe(a, b)


def embeddings(n, d, seed=1, sigma=0.5):
    """E ~ N(0, sigma^2) as fp32 (score std ~2.8 at d=128)."""
    rs = np.random.RandomState(seed)
    out = np.empty((n, d), np.float32)
    step = 1 << 16
    for i in range(0, n, step):
        out[i:i + step] = rs.normal(0.0, sigma, size=(min(step, n - i), d)).astype(np.float32)
    return out


def pick_roots(degrees, n_roots, seed=0):
    """a seeded subset of non-isolated roots, in ascending id order (the reference walks roots
    in id order, graph_gan.py:188)."""
    rs = np.random.RandomState(seed + 7)
    cand = np.flatnonzero(degrees > 0)
    if n_roots >= cand.shape[0]:
        return cand.astype(np.int32)
    return np.sort(rs.choice(cand, size=n_roots, replace=False)).astype(np.int32)
