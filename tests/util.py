"""Shared synthetic-input generators for tests, smoke() and bench.py.

Pose law (copied from the reference's test/polytope_derivs_test.jl:27-30):
r1 = 2*N(0,I3), r2 = N(0,I3), q = normalize(N(0,I4)); MRP p = q[2:4]/(1+q[1]) after flipping
q -> -q if q[1] < 0.  Counter-based RNG: numpy Philox, seed 1234 (test/runtests.jl:13).
Shapes are the ones the reference's own tests/examples use.
"""
from __future__ import annotations

import json
import os

import numpy as np

import dcol_b200 as dc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ELL_P = np.array([[1.1, .1, .1], [.1, 2.1, .1], [.1, .1, .6]])


def load_polytopes():
    d = json.load(open(os.path.join(GOLDEN, "polytopes.json")))
    return (np.array(d["A1"]), np.array(d["b1"]), np.array(d["A2"]), np.array(d["b2"]))


def rng(seed=1234):
    return np.random.Generator(np.random.Philox(seed))


def cls_name(att_kind, base):
    return getattr(dc, base + ("MRP" if att_kind == dc.ATT_MRP else ""))


def standard_shapes(att_kind=dc.ATT_QUAT, variant="cube"):
    """The 7 primitive kinds, in type order [polytope, capsule, cylinder, cone, sphere, ellipsoid, polygon].
    variant 'cube': shapes of SURVEY.md Appendix C; 'scene': shapes of examples/contact_physics (:264-267)."""
    A5, b5 = dc.create_n_sided(5, 0.6)
    g = lambda b: cls_name(att_kind, b)
    if variant == "cube":
        poly = dc.create_rect_prism(1.0, 1.0, 1.0, attitude="quat" if att_kind == dc.ATT_QUAT else "MRP")[0]
        caps = g("Capsule")(0.3, 1.2)
    else:
        _, _, A2, b2 = load_polytopes()
        poly = g("Polytope")(A2, b2)
        caps = g("Capsule")(0.2, 5.0)
    return [poly, caps, g("Cylinder")(0.6, 3.0), g("Cone")(2.0, np.deg2rad(22)), g("Sphere")(0.8),
            g("Ellipsoid")(ELL_P), g("Polygon")(A5, b5, 0.2)]


def random_poses(n, att_kind=dc.ATT_QUAT, seed=1234, r1_scale=2.0, r2_scale=1.0):
    g = rng(seed)
    r1 = r1_scale * g.standard_normal((n, 3))
    r2 = r2_scale * g.standard_normal((n, 3))
    q1 = g.standard_normal((n, 4)); q1 /= np.linalg.norm(q1, axis=1, keepdims=True)
    q2 = g.standard_normal((n, 4)); q2 /= np.linalg.norm(q2, axis=1, keepdims=True)
    if att_kind == dc.ATT_QUAT:
        return r1, q1, r2, q2
    q1 = np.where(q1[:, :1] < 0, -q1, q1); q2 = np.where(q2[:, :1] < 0, -q2, q2)
    return r1, q1[:, 1:] / (1 + q1[:, :1]), r2, q2[:, 1:] / (1 + q2[:, :1])


def mixed_batch(n, att_kind=dc.ATT_QUAT, seed=1234, variant="cube", table=None):
    """n pairs, type pair uniform over the 7x7 grid (BASELINE config 2 / 3).  Returns (table, s1, s2, r1, a1, r2, a2)."""
    table = table or dc.ShapeTable()
    ids = np.array([table.add(p) for p in standard_shapes(att_kind, variant)], dtype=np.int32)
    g = rng(seed + 1)
    s1 = ids[g.integers(0, 7, n)]
    s2 = ids[g.integers(0, 7, n)]
    r1, a1, r2, a2 = random_poses(n, att_kind, seed)
    return table, s1, s2, r1, a1, r2, a2


def cube_batch(n, att_kind=dc.ATT_QUAT, seed=1234, table=None):
    """BASELINE config 1: polytope(cube)-polytope(cube)."""
    table = table or dc.ShapeTable()
    cube = dc.create_rect_prism(1.0, 1.0, 1.0, attitude="quat" if att_kind == dc.ATT_QUAT else "MRP")[0]
    sid = table.add(cube)
    s = np.full(n, sid, dtype=np.int32)
    r1, a1, r2, a2 = random_poses(n, att_kind, seed)
    return table, s, s.copy(), r1, a1, r2, a2


def rel_err(a, b, floor=1e-12):
    """max-norm relative error per row: |a-b|_inf / max(|b|_inf, floor)"""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.ndim == 1:
        return np.abs(a - b) / np.maximum(np.abs(b), floor)
    ax = tuple(range(1, a.ndim))
    return np.abs(a - b).max(axis=ax) / np.maximum(np.abs(b).max(axis=ax), floor)


def oracle_spread(oracle, kind, ta, s1, s2, r1, a1, r2, a2, att_kind, tol, base, trials=3, seed=99):
    """Intrinsic sensitivity of the oracle's own outputs: row error between the oracle on the given inputs (`base`) and on
    inputs whose every pose entry is multiplied by 1 + 2.2e-16*U(-1,1) (a 1-ulp perturbation), max over `trials`.
    The witness point x (and with it J) is ill-conditioned for a small tail of pairs: there the
    spread exceeds the parity tolerance for ANY implementation, including the oracle compiled with other flags."""
    g = rng(seed)
    out = {}
    for _ in range(trials):
        pr = [a * (1.0 + 2.2e-16 * (2 * g.random(np.shape(a)) - 1)) for a in (r1, a1, r2, a2)]
        op = oracle.query(kind, ta, s1, s2, pr[0], pr[1], pr[2], pr[3], att_kind, tol)
        for k in ("x", "grad", "J"):
            if k in base:
                e = rel_err(op[k], base[k])
                out[k] = np.maximum(out.get(k, 0.0), e)
    return out
