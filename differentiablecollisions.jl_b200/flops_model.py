"""Algorithmic FLOP model of the proximity path (SURVEY.md §8d / BASELINE.md §3), frozen in code.

FMA = 2 flops, div/sqrt = 1.  n = nx (pair variables), m = ns (cone rows), socs = SOC dims > 0,
k = that pair's actual interior-point iteration count (returned by the kernels).
"""
from __future__ import annotations

import numpy as np

# per shape type: (extra variables, SOC dimension); orthant rows depend on the shape instance
TYPE_NE = {0: 0, 1: 1, 2: 1, 3: 0, 4: 0, 5: 0, 6: 2}
TYPE_D = {0: 0, 1: 4, 2: 4, 3: 3, 4: 4, 5: 4, 6: 4}


def f_iter(n, m, socs):
    return m * n * (n + 14) + 52 * m + n**3 / 3 + 4 * n**2 + 3 * n + sum((13 + n) * 2 * d * d + 240 for d in socs if d > 0)


def f_fixed(n, m):
    return m * n * (n + 9) + 600


def f_prox(n, m, socs, k):
    return f_fixed(n, m) + k * f_iter(n, m, socs)


def f_grad(n, m, socs, k, nt=14):
    return f_prox(n, m, socs, k) + nt * (4 * m + 60)


def f_jac(n, m, socs, k, nt=14):
    sd = sum(d * d for d in socs if d > 0)
    return f_prox(n, m, socs, k) + nt * (6 * m * n + 4 * m + 2 * n * n + 2 * sd + 250) + 2 * m * n * n + 2 * n**3 / 3


def batch_flops(kind, types, n_ort, shape1, shape2, iters, att_kind=0):
    """Total algorithmic FLOPs of a batch: types/n_ort per shape id, shape ids and iteration counts per pair."""
    types = np.asarray(types); n_ort = np.asarray(n_ort)
    ne = np.array([TYPE_NE[int(t)] for t in types]); dd = np.array([TYPE_D[int(t)] for t in types])
    s1 = np.asarray(shape1); s2 = np.asarray(shape2); k = np.asarray(iters, dtype=np.float64)
    n = 4 + ne[s1] + ne[s2]
    d1 = dd[s1]; d2 = dd[s2]
    m = n_ort[s1] + n_ort[s2] + d1 + d2
    soc_term = np.where(d1 > 0, (13 + n) * 2 * d1 * d1 + 240, 0) + np.where(d2 > 0, (13 + n) * 2 * d2 * d2 + 240, 0)
    fi = m * n * (n + 14) + 52 * m + n**3 / 3 + 4 * n**2 + 3 * n + soc_term
    fp = m * n * (n + 9) + 600 + k * fi
    nt = 14 if att_kind == 0 else 12
    if kind == 0:
        return float(fp.sum())
    if kind == 1:
        return float((fp + nt * (4 * m + 60)).sum())
    sd = d1 * d1 + d2 * d2
    return float((fp + nt * (6 * m * n + 4 * m + 2 * n * n + 2 * sd + 250) + 2 * m * n * n + 2 * n**3 / 3).sum())
