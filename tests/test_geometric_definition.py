"""CPU: the oracle's alpha against a generic NLP solver on the GEOMETRIC definition of the problem.

The reference states each primitive as a convex set S(alpha) that scales uniformly about its origin and
defines proximity as  min alpha  s.t.  x in S1(alpha) and x in S2(alpha)  (src/problem_matrices.jl:30-190 give the
sets in conic form).  Here the sets are written down directly as inequalities in the body frame — no
G/h matrices, no cones, no interior-point method — and handed to scipy's SLSQP.  Agreement pins the optimum the
oracle (and through it the CUDA path) converges to, independently of how the conic program is assembled or solved.
"""
import numpy as np
import pytest
from scipy.optimize import minimize

import dcol_b200 as dc
import util

NAMES = ["polytope", "capsule", "cylinder", "cone", "sphere", "ellipsoid", "polygon"]


def n_slack(P):
    name = type(P).__name__.replace("MRP", "")
    return {"Capsule": 1, "Polygon": 2}.get(name, 0)


def norm(v):
    return np.sqrt(v @ v + 1e-24)               # smooth at 0, concave constraint functions keep SLSQP's linearisations outer


def membership(P, Q, r, xw, a, sl):
    """g >= 0  <=>  world point xw lies in primitive P (pose Q, r) scaled by a; sl = the primitive's own slack variables."""
    name = type(P).__name__.replace("MRP", "")
    Qb = Q @ P.Q_offset
    xb = Qb.T @ (xw - r - Q @ P.r_offset)
    if name == "Polytope":                      # A x~ <= alpha b
        return a * P.b - P.A @ xb
    if name == "Sphere":
        return np.array([a * P.R - norm(xb)])
    if name == "Capsule":                       # within alpha R of a point gamma on the axis segment |gamma| <= alpha L/2
        d = xb - np.array([sl[0], 0.0, 0.0])
        return np.array([a * P.R - norm(d), a * P.L / 2 - sl[0], a * P.L / 2 + sl[0]])
    if name == "Cylinder":
        return np.array([a * P.R - norm(xb[1:]), a * P.L / 2 - xb[0], a * P.L / 2 + xb[0]])
    if name == "Cone":                          # apex at -3H/4 on the body x axis, base plane at +H/4, half angle beta
        return np.array([np.tan(P.beta) * (xb[0] + 0.75 * a * P.H) - norm(xb[1:]), 0.25 * a * P.H - xb[0]])
    if name == "Ellipsoid":                     # x~' P x~ <= alpha^2
        return np.array([a - norm(P.U @ xb)])
    if name == "Polygon":                       # planar polygon A y <= alpha b in the body xy plane, padded by alpha R
        d = xb - np.array([sl[0], sl[1], 0.0])
        return np.concatenate([a * P.b - P.A @ sl, [a * P.R - norm(d)]])
    raise AssertionError(name)


def slsqp_alpha(P1, Q1, r1, P2, Q2, r2):
    k1, k2 = n_slack(P1), n_slack(P2)

    def cons(v):
        x, a = v[:3], v[3]
        return np.concatenate([[a], membership(P1, Q1, r1, x, a, v[4:4 + k1]), membership(P2, Q2, r2, x, a, v[4 + k1:4 + k1 + k2])])
    best = None
    for a0, w in ((10.0, 0.5), (30.0, 0.5), (20.0, 0.3), (20.0, 0.7)):   # strictly feasible starts: large alpha, between the origins
        v0 = np.zeros(4 + k1 + k2); v0[:3] = w * r1 + (1 - w) * r2; v0[3] = a0
        res = minimize(lambda v: v[3], v0, jac=lambda v: np.eye(len(v))[3], method="SLSQP",
                       constraints=[{"type": "ineq", "fun": cons}], options={"ftol": 1e-13, "maxiter": 500})
        # SLSQP often stops with "positive directional derivative" at the optimum; any FEASIBLE end point is an
        # upper bound on the optimum, the smallest one over the starts is the estimate
        if cons(res.x).min() > -1e-9 and (best is None or res.x[3] < best):
            best = res.x[3]
    return best


@pytest.mark.parametrize("att_kind", [dc.ATT_QUAT, dc.ATT_MRP])
def test_alpha_matches_slsqp_on_geometric_sets(oracle, att_kind):
    shapes = util.standard_shapes(att_kind, "cube")
    shapes[3].r_offset = np.array([0.1, -0.2, 0.05])                       # offsets take part (src/problem_matrices.jl:46-47)
    shapes[1].Q_offset = dc.dcm_from_q(np.array([0.9, 0.1, -0.3, 0.2]))
    t = dc.ShapeTable()
    ids = [t.add(s) for s in shapes]
    n_pose = 3
    r1, a1, r2, a2 = util.random_poses(n_pose, att_kind, 77)
    dcm = dc.dcm_from_q if att_kind == dc.ATT_QUAT else dc.dcm_from_mrp
    s1, s2, R1, A1, R2, A2 = [], [], [], [], [], []
    for k in range(n_pose):
        for i in range(7):
            for j in range(7):
                s1.append(ids[i]); s2.append(ids[j]); R1.append(r1[k]); A1.append(a1[k]); R2.append(r2[k]); A2.append(a2[k])
    out = oracle.query(0, t.arrays(), np.array(s1, np.int32), np.array(s2, np.int32), np.array(R1), np.array(A1),
                       np.array(R2), np.array(A2), att_kind, 1e-10)
    assert (out["status"] == 0).all()
    worst = 0.0
    for n in range(len(s1)):
        k, i, j = n // 49, (n % 49) // 7, n % 7
        ref = slsqp_alpha(shapes[i], dcm(a1[k]), r1[k], shapes[j], dcm(a2[k]), r2[k])
        assert ref is not None, (NAMES[i], NAMES[j], k)
        err = abs(out["alpha"][n] - ref) / max(1.0, abs(ref))
        worst = max(worst, err)
        assert err < 2e-6, (NAMES[i], NAMES[j], k, out["alpha"][n], ref)   # SLSQP's own accuracy is the limit here
        # the witness point lies in both sets at the returned alpha (slack-free primitives; capsule: nearest axis point)
        for P, Q, r in ((shapes[i], dcm(a1[k]), r1[k]), (shapes[j], dcm(a2[k]), r2[k])):
            if n_slack(P) == 2:
                continue
            a = out["alpha"][n] * (1 + 1e-7)
            xb = (Q @ P.Q_offset).T @ (out["x"][n] - r - Q @ P.r_offset)
            assert membership(P, Q, r, out["x"][n], a, [np.clip(xb[0], -a * getattr(P, "L", 0) / 2, a * getattr(P, "L", 0) / 2)]).min() > -1e-7
