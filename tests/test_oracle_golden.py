"""CPU: pin the oracle (C++ restatement) to the committed fixtures.

* tests/golden/example_socp_2.json / polytopes.json : the reference's own JLD2 input fixtures
  (test/solver_tests.jl:38-64, test/polytope_derivs_test.jl:3-8) decoded by make_fixtures.py.
* tests/golden/appendix_c.json : known answers from an independent NumPy restatement
   — 49 pair types, 5 Jacobians/gradients, the fixture SOCP solution.
"""
import json
import os

import numpy as np
import pytest

import dcol_b200 as dc
import util

G = json.load(open(os.path.join(util.GOLDEN, "appendix_c.json")))
NAMES = ["polytope", "capsule", "cylinder", "cone", "sphere", "ellipsoid", "polygon"]


def fixed_poses(att_kind):
    q1 = np.array(G["pose1"]["q_unnormalised"], dtype=float); q1 /= np.linalg.norm(q1)
    q2 = np.array(G["pose2"]["q_unnormalised"], dtype=float); q2 /= np.linalg.norm(q2)
    r1 = np.array(G["pose1"]["r"]); r2 = np.array(G["pose2"]["r"])
    if att_kind == dc.ATT_QUAT:
        return r1, q1, r2, q2
    return r1, q1[1:] / (1 + q1[0]), r2, q2[1:] / (1 + q2[0])


def run_pair(oracle, kind, n1, n2, att_kind, tol=1e-6):
    shapes = util.standard_shapes(att_kind, "cube")
    t = dc.ShapeTable()
    ids = [t.add(s) for s in shapes]
    r1, a1, r2, a2 = fixed_poses(att_kind)
    s1 = np.array([ids[NAMES.index(n1)]], dtype=np.int32); s2 = np.array([ids[NAMES.index(n2)]], dtype=np.int32)
    return oracle.query(kind, t.arrays(), s1, s2, r1[None], a1[None], r2[None], a2[None], att_kind, tol)


def test_fixture_socp_solution(oracle):
    """test/solver_tests.jl:63-64: |s'z| < 1e-10 at pdip_tol=1e-12; plus the Appendix C iterates."""
    d = json.load(open(os.path.join(util.GOLDEN, "example_socp_2.json")))
    Gm = np.array(d["G"]); h = np.array(d["h"]); c = np.array(d["c"])
    x, s, z, it, st = oracle.solve_socp(c, Gm, h, d["n_ort"], d["n_soc1"], d["n_soc2"], 1e-12)
    assert st == 0 and abs(s @ z) < 1e-10
    assert it == G["fixture_socp"]["tol1e-12"]["iters"]
    np.testing.assert_allclose(x, G["fixture_socp"]["tol1e-12"]["x"], atol=2e-10)
    x, s, z, it, st = oracle.solve_socp(c, Gm, h, d["n_ort"], d["n_soc1"], d["n_soc2"], 1e-6)
    assert st == 0 and it == G["fixture_socp"]["tol1e-6"]["iters"]
    np.testing.assert_allclose(x, G["fixture_socp"]["tol1e-6"]["x"], atol=2e-10)
    # KKT residuals (test/deprecated_tests/lp_solver_tests.jl:17-41 property)
    x, s, z, it, st = oracle.solve_socp(c, Gm, h, d["n_ort"], d["n_soc1"], d["n_soc2"], 1e-10)
    assert np.abs(Gm.T @ z + c).max() < 1e-8 and np.abs(Gm @ x + s - h).max() < 1e-8


@pytest.mark.parametrize("att_kind", [dc.ATT_QUAT, dc.ATT_MRP])
def test_49_pair_table(oracle, att_kind):
    for row in G["pairs"]:
        o = run_pair(oracle, 0, row["p1"], row["p2"], att_kind)
        assert o["status"][0] == 0
        assert o["iters"][0] == row["iters"], (row["p1"], row["p2"], o["iters"][0], row["iters"])
        assert abs(o["alpha"][0] - row["alpha"]) < 5e-10 * max(1, abs(row["alpha"])), (row, o["alpha"][0])
        np.testing.assert_allclose(o["x"][0], row["x"], atol=2e-8)


def test_fixture_polytopes(oracle):
    A1, b1, A2, b2 = util.load_polytopes()
    t = dc.ShapeTable()
    i1 = t.add(dc.Polytope(A1, b1)); i2 = t.add(dc.Polytope(A2, b2))
    r1, a1, r2, a2 = fixed_poses(dc.ATT_QUAT)
    for tol, key in ((1e-6, "tol1e-6"), (1e-8, "tol1e-8")):
        o = oracle.query(0, t.arrays(), [i1], [i2], r1[None], a1[None], r2[None], a2[None], 0, tol)
        g = G["fixture_polytopes"][key]
        assert o["iters"][0] == g["iters"]
        assert abs(o["alpha"][0] - g["alpha"]) < 5e-10
        if "x" in g:
            np.testing.assert_allclose(o["x"][0], g["x"], atol=2e-8)


@pytest.mark.parametrize("key", list(G["jacobians"].keys()))
def test_jacobian_tables(oracle, key):
    name, att = key.split("/")
    n1, n2 = name.replace("cube", "polytope").split("–")
    att_kind = dc.ATT_QUAT if att == "quat" else dc.ATT_MRP
    j = run_pair(oracle, 2, n1, n2, att_kind)
    g = run_pair(oracle, 1, n1, n2, att_kind)
    Jg = np.array(G["jacobians"][key]["J"]); gg = np.array(G["jacobians"][key]["grad"])
    # 9 printed decimals; the table used complex-step + dense solves, the oracle uses forward-mode duals + block solves
    np.testing.assert_allclose(j["J"][0], Jg, atol=5e-8, rtol=1e-7)
    np.testing.assert_allclose(g["grad"][0], gg, atol=5e-8, rtol=1e-7)


def test_sphere_sphere_analytic(oracle):
    """alpha = |r2 - r1| / (R1 + R2) + O(mu), x on the segment (closed form)."""
    o = run_pair(oracle, 0, "sphere", "sphere", dc.ATT_QUAT, tol=1e-10)
    r1, _, r2, _ = fixed_poses(0)
    exact = np.linalg.norm(r2 - r1) / 1.6
    assert abs(o["alpha"][0] - exact) < 1e-8
    np.testing.assert_allclose(o["x"][0], 0.5 * (r1 + r2), atol=1e-7)
