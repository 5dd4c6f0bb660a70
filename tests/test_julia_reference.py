"""Parity against REAL Julia output, when tests/golden/julia_reference.json exists (made by tests/golden/dump_reference.jl
with the DifferentiableCollisions.jl package; Julia is not in the build image, so the file is absent there and these
tests skip).  north_star tolerances: alpha, x 1e-6 relative; gradient / Jacobian 1e-5 relative; alpha <= 1 identical."""
import json
import os

import numpy as np
import pytest

import dcol_b200 as dc
import util

PATH = os.path.join(util.GOLDEN, "julia_reference.json")
NAMES = ["polytope", "capsule", "cylinder", "cone", "sphere", "ellipsoid", "polygon"]


def _load():
    if not os.path.exists(PATH):
        pytest.skip("tests/golden/julia_reference.json not present (run tests/golden/dump_reference.jl with Julia)")
    return json.load(open(PATH))


def _groups(ref):
    """records grouped by (shapes variant, attitude kind) -> arrays"""
    out = {}
    for r in ref["records"]:
        if r.get("status", 0) != 0:
            continue
        out.setdefault((r["shapes"], r["att"]), []).append(r)
    for (variant, att), recs in out.items():
        att_kind = 0 if att == "quat" else 1
        t = dc.ShapeTable(); ids = [t.add(s) for s in util.standard_shapes(att_kind, variant)]
        s1 = np.array([ids[NAMES.index(r["p1"])] for r in recs], np.int32)
        s2 = np.array([ids[NAMES.index(r["p2"])] for r in recs], np.int32)
        arr = lambda k: np.array([r[k] for r in recs], float)
        nt = 14 if att_kind == 0 else 12
        J = arr("J_colmajor").reshape(len(recs), nt, 4).transpose(0, 2, 1)
        yield att_kind, t, s1, s2, arr("r1"), arr("a1"), arr("r2"), arr("a2"), arr("alpha"), arr("x"), arr("grad"), J, ref.get("pdip_tol", 1e-6)


def _check(q, ref):
    for att_kind, t, s1, s2, r1, a1, r2, a2, alpha, x, grad, J, tol in _groups(ref):
        p = q(0, t.arrays(), s1, s2, r1, a1, r2, a2, att_kind, tol)
        g = q(1, t.arrays(), s1, s2, r1, a1, r2, a2, att_kind, tol)
        j = q(2, t.arrays(), s1, s2, r1, a1, r2, a2, att_kind, tol)
        assert (p["status"] == 0).all()
        assert util.rel_err(p["alpha"], alpha).max() < 1e-6
        assert ((p["alpha"] <= 1) == (alpha <= 1)).all()
        assert np.quantile(util.rel_err(p["x"], x), 0.995) < 1e-6
        assert np.quantile(util.rel_err(g["grad"], grad), 0.995) < 1e-5
        assert np.quantile(util.rel_err(j["J"], J), 0.995) < 1e-5


def test_generator_twin_matches_the_dump():
    ref = _load()
    from golden.reference_inputs import SplitMix64, poses
    g = SplitMix64(1234)
    for r in ref["records"][:50]:
        r1, q1, r2, q2 = poses(g)
        np.testing.assert_allclose(r["r1"], r1, rtol=1e-13); np.testing.assert_allclose(r["r2"], r2, rtol=1e-13)
        if r["att"] == "quat":
            np.testing.assert_allclose(r["a1"], q1, rtol=1e-13, atol=1e-15)


def test_oracle_matches_julia(oracle):
    _check(oracle.query, _load())


@pytest.mark.gpu
def test_cuda_matches_julia(request):
    ref = _load()
    ctx = request.getfixturevalue("ctx")

    def q(kind, ta, s1, s2, r1, a1, r2, a2, att, tol):
        ctx.set_shape_arrays(*ta)
        return ctx._batch(kind, s1, s2, r1, a1, r2, a2, att, tol)
    _check(q, ref)
