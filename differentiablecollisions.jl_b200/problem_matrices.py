"""Host-side dense conic form of ONE pair: the (c, G, h) the reference hands to `solve_socp`.

The batched CUDA path never forms G.  This module exists for the two places where a caller sees
the dense problem: `verbose=True` on the single-pair calls (the reference prints one row per interior-point iteration,
`src/solvers/coneqp/static_solver2.jl:117-120,179-183`; here the dense problem is solved by `dcol_solve_socp_trace` on
the GPU and the table is printed from its per-iteration record) and users of `Context.solve_socp`.

What it follows: `src/problem_matrices.jl:30-191` (one block pair per primitive kind: rows `h - G y` in the
non-negative orthant, then rows in a second-order cone; `y = [x; alpha; extra variables of the primitive]`) and
`src/combine_problem_matrices.jl:15-86` (orthant rows of both primitives first, then the two cone blocks; each
primitive's extra variables get their own columns after the common four).
"""
from __future__ import annotations

import math
import numpy as np

from .primitives import (ATT_QUAT, CAPSULE, CONE, CYLINDER, ELLIPSOID, POLYGON, POLYTOPE, SPHERE, dcm_from_q)


def _dcm_from_mrp(p):
    # src/problem_matrices.jl:18-27 (the closed form, not the detour through a quaternion)
    p1, p2, p3 = (float(v) for v in p)
    den = (p1 * p1 + p2 * p2 + p3 * p3 + 1.0) ** 2
    a = 4.0 * (p1 * p1 + p2 * p2 + p3 * p3) - 4.0
    return np.array([[den - 8 * (p2 * p2 + p3 * p3), 8 * p1 * p2 + p3 * a, 8 * p1 * p3 - p2 * a],
                     [8 * p1 * p2 - p3 * a, den - 8 * (p1 * p1 + p3 * p3), 8 * p2 * p3 + p1 * a],
                     [8 * p1 * p3 + p2 * a, 8 * p2 * p3 - p1 * a, den - 8 * (p1 * p1 + p2 * p2)]]) / den


def _ball_block(R, r, extra_cols):
    """SOC rows `[R alpha; x - r - extra_cols e]` (sphere, capsule, cylinder, polygon)."""
    ne = extra_cols.shape[1]
    G = np.zeros((4, 4 + ne))
    G[0, 3] = -R
    G[1:, :3] = -np.eye(3)
    G[1:, 4:] = extra_cols
    return G, np.concatenate([[0.0], -r])


def problem_matrices(prim):
    """(G_ort, h_ort, G_soc, h_soc) of one primitive at its current pose (`prim.r`, `prim.q` | `prim.p`)."""
    Q = dcm_from_q(prim.q) if prim.att_kind == ATT_QUAT else _dcm_from_mrp(prim.p)
    r = np.asarray(prim.r, dtype=np.float64) + Q @ np.asarray(prim.r_offset, dtype=np.float64)
    Q = Q @ np.asarray(prim.Q_offset, dtype=np.float64)
    t = prim.shape_type
    none_o = (np.zeros((0, 4)), np.zeros(0))
    if t == POLYTOPE:                                           # :71-80  A Q'(x - r) <= alpha b
        AQt = prim.A @ Q.T
        return np.hstack([AQt, -prim.b[:, None]]), AQt @ r, np.zeros((0, 4)), np.zeros(0)
    if t == SPHERE:                                             # :92-107
        G, h = _ball_block(prim.R, r, np.zeros((3, 0)))
        return none_o[0], none_o[1], G, h
    if t in (CAPSULE, CYLINDER):                                # :30-39, :148-163  axis point r + e bx, |e| <= alpha L/2
        bx = Q[:, 0]
        G, h = _ball_block(prim.R, r, bx[:, None])
        Go = np.array([[0, 0, 0, -prim.L / 2, 1.0], [0, 0, 0, -prim.L / 2, -1.0]])
        ho = np.zeros(2)
        if t == CYLINDER:                                       # two end planes
            Go = np.vstack([Go, np.concatenate([-bx, [-prim.L / 2, 0.0]]), np.concatenate([bx, [-prim.L / 2, 0.0]])])
            ho = np.array([0.0, 0.0, -bx @ r, bx @ r])
        return Go, ho, G, h
    if t == CONE:                                               # :56-66
        tb = math.tan(prim.beta)
        EQt = np.diag([tb, 1.0, 1.0]) @ Q.T
        G = np.hstack([-EQt, -np.array([[tb * 3 * prim.H / 4], [0.0], [0.0]])])
        bx = Q[:, 0]
        return np.concatenate([bx, [-prim.H / 4]])[None, :], np.array([bx @ r]), G, -EQt @ r
    if t == POLYGON:                                            # :118-129  in-plane point r + Q~ e, A e <= alpha b
        nh = prim.A.shape[0]
        Go = np.hstack([np.zeros((nh, 3)), -prim.b[:, None], prim.A])
        G, h = _ball_block(prim.R, r, Q[:, :2])
        return Go, np.zeros(nh), G, h
    if t == ELLIPSOID:                                          # :168-178  |U Q'(x - r)| <= alpha
        UQt = prim.U @ Q.T
        G = np.zeros((4, 4)); G[0, 3] = -1.0; G[1:, :3] = -UQt
        return none_o[0], none_o[1], G, np.concatenate([[0.0], -UQt @ r])
    raise TypeError("not a primitive")


def combine_problem_matrices(Go1, ho1, Gs1, hs1, Go2, ho2, Gs2, hs2):
    """(c, G, h, n_ort, d1, d2): rows ordered [orthant 1; orthant 2; cone 1; cone 2], columns
    [x, alpha, extras of primitive 1, extras of primitive 2] (`src/combine_problem_matrices.jl:15-86`)."""
    e1 = max(Go1.shape[1], Gs1.shape[1]) - 4
    e2 = max(Go2.shape[1], Gs2.shape[1]) - 4
    nx = 4 + e1 + e2

    def widen(B, first, ne):
        out = np.zeros((B.shape[0], nx))
        if B.shape[0]:
            out[:, :4] = B[:, :4]
            out[:, first:first + ne] = B[:, 4:4 + ne]
        return out

    G = np.vstack([widen(Go1, 4, e1), widen(Go2, 4 + e1, e2), widen(Gs1, 4, e1), widen(Gs2, 4 + e1, e2)])
    h = np.concatenate([ho1, ho2, hs1, hs2])
    c = np.zeros(nx); c[3] = 1.0
    return c, G, h, Go1.shape[0] + Go2.shape[0], Gs1.shape[0], Gs2.shape[0]


def pair_problem(prim1, prim2):
    """The dense conic program of a pair, `min alpha  s.t.  h - G y in K` — what `proximity` solves."""
    return combine_problem_matrices(*problem_matrices(prim1), *problem_matrices(prim2))


TRACE_HEADER = ("iter     objv        μ       |Gx+s-h|      |s∘z|      α\n"
                "---------------------------------------------------------")


def format_trace(trace, iters):
    """The table `solve_socp(...; verbose = true)` prints (`static_solver2.jl:117-120,179-183`), from a trace record."""
    lines = [TRACE_HEADER]
    for k in range(min(int(iters), trace.shape[0])):
        o, mu, pres, cent, al = trace[k]
        lines.append("%3d   %10.3e  %9.2e  %9.2e  %9.2e  % 6.4f" % (k + 1, o, mu, pres, cent, al))
    return "\n".join(lines)
