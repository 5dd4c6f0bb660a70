"""GPU (-m gpu): the CUDA library, called through its C ABI, against the CPU oracle on the same inputs.

Tolerances are BASELINE.json's: alpha and x within 1e-6 relative, Jacobian / gradient within 1e-5
relative (FP64, same pdip_tol), classification alpha <= 1 identical.  The witness point x (and with it
J) is ill-conditioned for a small tail of pairs (SURVEY.md §7 "Parity tail": 1e-15 noise in one
interior-point step moves x by up to 3e-5 relative), so x and J are asserted on the pass fraction plus
a hard cap, while alpha, classification, status and iteration counts are asserted on every pair.
"""
import json
import os

import numpy as np
import pytest

import dcol_b200 as dc
import props
import util

pytestmark = pytest.mark.gpu

TOL_AX = 1e-6     # alpha, x  (north_star)
TOL_J = 1e-5      # Jacobian / gradient (north_star)


def gq(ctx):
    def query(kind, ta, s1, s2, r1, a1, r2, a2, att, tol):
        ctx.set_shape_arrays(*ta)
        return ctx._batch(kind, s1, s2, r1, a1, r2, a2, att, tol)
    return query


def compare(g, o, kind, strict_frac=0.999, spread=None):
    """alpha, classification and status on EVERY pair.  x / grad / J (row max-norm relative error):
    with `spread` (util.oracle_spread: the oracle's own response to a 1-ulp input perturbation) the tail is bounded pair by
    pair and as a population — every pair within max(tol, 20 x its spread) up to a handful of exceptions (the spread is a
    3-sample estimate), no more pairs over tolerance than the perturbed oracle itself produces, and a hard cap of
    10 x the largest spread; without it (small special-case batches) by pass fraction plus a hard cap."""
    assert (g["status"] == o["status"]).all()
    ok = o["status"] == 0
    n = int(ok.sum())
    assert (g["iters"] != o["iters"]).sum() <= max(3, 5e-4 * n)      # same iterate sequence up to round-off at the mu test
    ea = util.rel_err(g["alpha"], o["alpha"])[ok]
    assert ea.max() < TOL_AX, ea.max()
    assert ((g["alpha"] <= 1) == (o["alpha"] <= 1))[ok].all()
    for key, tol, cap in (("x", TOL_AX, 1e-2), ("grad", TOL_J, 1e-2), ("J", TOL_J, 1e-2)):
        if key not in g or (key == "x" and kind == 1):
            continue
        e = util.rel_err(g[key], o[key])[ok]
        frac = (e < tol).mean()
        if spread is not None:
            sp = spread[key][ok]
            viol = int((e > np.maximum(tol, 20.0 * sp)).sum())
            assert viol <= max(2, 1e-4 * n), (key, viol, e.max())
            assert (e >= tol).sum() <= 1.5 * (sp >= tol).sum() + 5, (key, int((e >= tol).sum()), int((sp >= tol).sum()))
            assert e.max() <= 10.0 * max(tol, sp.max()), (key, e.max(), sp.max())
            assert frac >= 0.995, (key, frac)
        else:
            assert frac >= strict_frac and e.max() < cap, (key, e.max(), frac)


@pytest.mark.parametrize("att_kind", [dc.ATT_QUAT, dc.ATT_MRP])
@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("variant", ["cube", "scene"])
def test_mixed_7x7_vs_oracle(ctx, oracle, att_kind, kind, variant):
    """BASELINE config 2/3 at an oracle-sized n: all 49 type pairs, random poses."""
    n = 20000
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att_kind, seed=1234 + kind, variant=variant)
    g = gq(ctx)(kind, t.arrays(), s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
    o = oracle.query(kind, t.arrays(), s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
    assert (o["status"] == 0).all()
    # 'scene' has the long thin Capsule(0.2, 5.0) of the reference's contact example: near-parallel capsule pairs have a
    # non-unique witness point, the ill-conditioned tail of SURVEY.md §7 (see test_oracle_x_sensitivity_tail); the tail is
    # bounded against the oracle's own sensitivity, pair by pair.  The 'cube' set has no such pairs: every pair in tolerance.
    sp = util.oracle_spread(oracle, kind, t.arrays(), s1, s2, r1, a1, r2, a2, att_kind, 1e-6, o)
    compare(g, o, kind, spread=sp)
    if variant == "cube":
        for key, tol in (("x", TOL_AX), ("grad", TOL_J), ("J", TOL_J)):
            if key in g and key in o:
                assert util.rel_err(g[key], o[key]).max() < tol, key


@pytest.mark.parametrize("tol", [1e-4, 1e-8, 1e-10])
def test_cube_cube_config1(ctx, oracle, tol):
    """BASELINE config 1: polytope(cube)-polytope(cube), 10k random poses."""
    t, s1, s2, r1, a1, r2, a2 = util.cube_batch(10000, 0)
    # the reduced KKT matrix G'S^-1ZG has cond ~ 1/mu^2: below pdip_tol ~1e-8 the reference formula for J is
    # round-off dominated in ANY implementation, so J parity is asserted at 1e-4 / 1e-8 and alpha, x at every tol
    for kind in ((0, 2) if tol >= 1e-8 else (0,)):
        g = gq(ctx)(kind, t.arrays(), s1, s2, r1, a1, r2, a2, 0, tol)
        o = oracle.query(kind, t.arrays(), s1, s2, r1, a1, r2, a2, 0, tol)
        compare(g, o, kind, strict_frac=0.99)


@pytest.mark.parametrize("tol", [1e-3, 1e-8, 1e-10])
@pytest.mark.parametrize("att_kind", [dc.ATT_QUAT, dc.ATT_MRP])
def test_mixed_tolerance_sweep(ctx, oracle, tol, att_kind):
    """All 49 type pairs away from the default pdip_tol: the cone-block passes (NT scaling, line searches on the scaled
    point, unit-triangular factor) must follow the reference's iterate sequence down to 1e-10 — status identical,
    iteration counts identical up to round-off at the mu test, alpha to 1e-6 relative (measured ~1e-12)."""
    n = 12000
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att_kind, seed=77, variant="cube")
    g = gq(ctx)(0, t.arrays(), s1, s2, r1, a1, r2, a2, att_kind, tol)
    o = oracle.query(0, t.arrays(), s1, s2, r1, a1, r2, a2, att_kind, tol)
    assert (o["status"] == 0).all() and (g["status"] == 0).all()
    assert (g["iters"] != o["iters"]).sum() <= max(3, 5e-4 * n)
    same = g["iters"] == o["iters"]
    assert util.rel_err(g["alpha"], o["alpha"])[same].max() < TOL_AX
    assert (util.rel_err(g["x"], o["x"])[same] < TOL_AX).mean() > 0.999
    if tol <= 1e-8:
        assert np.median(util.rel_err(g["alpha"], o["alpha"])[same]) < 1e-11
    assert ((g["alpha"] <= 1) == (o["alpha"] <= 1))[same].all()


def test_golden_table(ctx):
    """49 pair types at the fixed poses of tests/golden/appendix_c.json."""
    G = json.load(open(os.path.join(util.GOLDEN, "appendix_c.json")))
    names = ["polytope", "capsule", "cylinder", "cone", "sphere", "ellipsoid", "polygon"]
    for att_kind in (0, 1):
        shapes = util.standard_shapes(att_kind, "cube")
        t = dc.ShapeTable(); ids = [t.add(s) for s in shapes]
        q1 = np.array(G["pose1"]["q_unnormalised"], float); q2 = np.array(G["pose2"]["q_unnormalised"], float)
        if att_kind == 1:
            q1 /= np.linalg.norm(q1); q2 /= np.linalg.norm(q2)
            a1, a2 = q1[1:] / (1 + q1[0]), q2[1:] / (1 + q2[0])
        else:
            a1, a2 = q1, q2                      # un-normalised quaternions are accepted (dcm_from_q normalises)
        rows = G["pairs"]
        n = len(rows)
        s1 = np.array([ids[names.index(r["p1"])] for r in rows], np.int32)
        s2 = np.array([ids[names.index(r["p2"])] for r in rows], np.int32)
        r1 = np.tile(G["pose1"]["r"], (n, 1)); r2 = np.tile(G["pose2"]["r"], (n, 1))
        g = gq(ctx)(0, t.arrays(), s1, s2, r1, np.tile(a1, (n, 1)), r2, np.tile(a2, (n, 1)), att_kind, 1e-6)
        assert (g["status"] == 0).all()
        assert (g["iters"] == np.array([r["iters"] for r in rows])).all()
        np.testing.assert_allclose(g["alpha"], [r["alpha"] for r in rows], rtol=1e-9)
        np.testing.assert_allclose(g["x"], [r["x"] for r in rows], atol=5e-8)
    for key, val in G["jacobians"].items():
        name, att = key.split("/")
        n1, n2 = name.replace("cube", "polytope").split("–")
        att_kind = 0 if att == "quat" else 1
        shapes = util.standard_shapes(att_kind, "cube")
        t = dc.ShapeTable(); ids = [t.add(s) for s in shapes]
        q1 = np.array(G["pose1"]["q_unnormalised"], float); q1 /= np.linalg.norm(q1)
        q2 = np.array(G["pose2"]["q_unnormalised"], float); q2 /= np.linalg.norm(q2)
        a1, a2 = (q1, q2) if att_kind == 0 else (q1[1:] / (1 + q1[0]), q2[1:] / (1 + q2[0]))
        args = (t.arrays(), [ids[names.index(n1)]], [ids[names.index(n2)]], np.array([G["pose1"]["r"]]), a1[None],
                np.array([G["pose2"]["r"]]), a2[None], att_kind, 1e-6)
        j = gq(ctx)(2, *args); gr = gq(ctx)(1, *args)
        np.testing.assert_allclose(j["J"][0], val["J"], atol=5e-8, rtol=1e-7)
        np.testing.assert_allclose(gr["grad"][0], val["grad"], atol=5e-8, rtol=1e-7)


def test_reference_properties_on_gpu(ctx):
    """The reference's own derivative testsets, run against the CUDA path."""
    props.cone_capsule_derivative_properties(gq(ctx), has_slow=False, att_kind=0)
    props.cone_capsule_derivative_properties(gq(ctx), has_slow=False, att_kind=1)
    props.polytope_derivative_properties(gq(ctx))


def test_fast_jacobian_vs_oracle_slow(ctx, oracle):
    """test/proximity_test.jl:48: fast IFT (GPU) vs dense-KKT `proximity_jacobian_slow` (oracle)."""
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(5000, 0, seed=77)
    g = gq(ctx)(2, t.arrays(), s1, s2, r1, a1, r2, a2, 0, 1e-6)
    o = oracle.query(3, t.arrays(), s1, s2, r1, a1, r2, a2, 0, 1e-6)
    fro = np.sqrt(((g["J"] - o["J"]) ** 2).sum(axis=(1, 2)))
    assert np.median(fro) < 1e-5


@pytest.mark.parametrize("att_kind", [0, 1])
def test_offsets_and_unnormalised_quaternions(ctx, oracle, att_kind):
    """r_offset / Q_offset (untested in the reference) and un-normalised q (App. D-9) against the oracle."""
    g = util.rng(5)
    shapes = util.standard_shapes(att_kind, "scene")
    for s in shapes:
        s.r_offset = 0.3 * g.standard_normal(3)
        q = g.standard_normal(4)
        s.Q_offset = dc.dcm_from_q(q)
    t = dc.ShapeTable(); ids = np.array([t.add(s) for s in shapes], np.int32)
    n = 4000
    s1 = ids[g.integers(0, 7, n)]; s2 = ids[g.integers(0, 7, n)]
    r1, a1, r2, a2 = util.random_poses(n, att_kind, 99)
    if att_kind == 0:
        a1 = a1 * (0.5 + g.random((n, 1))); a2 = a2 * (0.5 + 2 * g.random((n, 1)))
    for kind in (0, 1, 2):
        gg = gq(ctx)(kind, t.arrays(), s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
        o = oracle.query(kind, t.arrays(), s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
        compare(gg, o, kind, strict_frac=0.995)
    props.structural_jacobian_properties(gg, a1, a2, att_kind) if att_kind == 1 else None


def test_big_polytopes_and_polygons(ctx, oracle):
    """variable face counts: 14- and 8-face fixture polytopes, 30-face polytope, 3..12-sided polygons."""
    A1, b1, A2, b2 = util.load_polytopes()
    g = util.rng(8)
    dirs = g.standard_normal((30, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    shapes = [dc.Polytope(A1, b1), dc.Polytope(A2, b2), dc.Polytope(dirs, 0.7 + 0.3 * g.random(30))]
    for nh in (3, 8, 12):
        A, b = dc.create_n_sided(nh, 0.5)
        shapes.append(dc.Polygon(A, b, 0.15))
    t = dc.ShapeTable(); ids = np.array([t.add(s) for s in shapes], np.int32)
    n = 6000
    s1 = ids[g.integers(0, len(ids), n)]; s2 = ids[g.integers(0, len(ids), n)]
    r1, a1, r2, a2 = util.random_poses(n, 0, 123)
    for kind in (0, 2):
        gg = gq(ctx)(kind, t.arrays(), s1, s2, r1, a1, r2, a2, 0, 1e-6)
        o = oracle.query(kind, t.arrays(), s1, s2, r1, a1, r2, a2, 0, 1e-6)
        compare(gg, o, kind, strict_frac=0.995)


def test_edge_cases(ctx, oracle):
    """empty batch, batch of one, identical poses (deep penetration), far apart, device-memory mode parity."""
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(64, 0)
    q = gq(ctx)
    e = q(0, t.arrays(), s1[:0], s2[:0], r1[:0], a1[:0], r2[:0], a2[:0], 0, 1e-6)
    assert e["alpha"].shape == (0,)
    one = q(2, t.arrays(), s1[:1], s2[:1], r1[:1], a1[:1], r2[:1], a2[:1], 0, 1e-6)
    o1 = oracle.query(2, t.arrays(), s1[:1], s2[:1], r1[:1], a1[:1], r2[:1], a2[:1], 0, 1e-6)
    np.testing.assert_allclose(one["alpha"], o1["alpha"], rtol=1e-9)
    np.testing.assert_allclose(one["J"], o1["J"], rtol=1e-5, atol=1e-8)
    # coincident centres and 1e3-separated centres
    for r2x in (r1.copy(), r1 + 1e3):
        g = q(0, t.arrays(), s1, s2, r1, a1, r2x, a2, 0, 1e-6)
        o = oracle.query(0, t.arrays(), s1, s2, r1, a1, r2x, a2, 0, 1e-6)
        assert (g["status"] == o["status"]).all()
        ok = o["status"] == 0
        assert (np.abs(g["alpha"] - o["alpha"])[ok] < 1e-6 * np.maximum(1e-3, np.abs(o["alpha"][ok]))).all()
    # bad shape index is an API error, not a crash
    bad = s1.copy(); bad[3] = 1000
    with pytest.raises(RuntimeError):
        q(0, t.arrays(), bad, s2, r1, a1, r2, a2, 0, 1e-6)


def test_pdip_failed_status(ctx, oracle):
    """An unreachable pdip_tol (-1: `mu < -1` is never true, static_solver2.jl:144) -> nonzero status for every pair
    ('pdip failed' after 50 iterations, :188, or a numerical breakdown on the way), reported per pair by the
    batched call, never as an exception or a crash."""
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(512, 0)
    for kind in (0, 1, 2):
        g = gq(ctx)(kind, t.arrays(), s1, s2, r1, a1, r2, a2, 0, -1.0)
        assert (g["status"] != 0).all()
        if kind:
            assert np.isnan(g["grad" if kind == 1 else "J"]).all()
    o = oracle.query(0, t.arrays(), s1, s2, r1, a1, r2, a2, 0, -1.0)
    assert (o["status"] != 0).all()


def test_single_pair_dropin_api():
    """dc.proximity(P1, P2) etc. keep the reference's call surface (README.md:50-59)."""
    cone = dc.Cone(2.0, np.deg2rad(22)); cone.r = np.array([0.1, -0.2, 0.3]); cone.q = np.array([1, .1, .2, .3]) / np.linalg.norm([1, .1, .2, .3])
    caps = dc.Capsule(0.3, 1.2); caps.r = np.array([1.5, 0.3, -0.2]); caps.q = np.array([1, .2, -.3, .4]) / np.linalg.norm([1, .2, -.3, .4])
    a, x = dc.proximity(cone, caps)
    a2, x2, J = dc.proximity_jacobian(cone, caps)
    a3, g = dc.proximity_gradient(cone, caps)
    assert abs(a - 1.306940177502) < 1e-9 and abs(a2 - a) < 1e-12 and abs(a3 - a) < 1e-12
    assert J.shape == (4, 14) and g.shape == (14,) and x.shape == (3,)
    assert np.abs(J[3] - g).max() < 1e-3
    with pytest.raises(TypeError):
        dc.proximity(cone, dc.CapsuleMRP(0.3, 1.2))
    with pytest.raises((RuntimeError, ArithmeticError)):      # error("pdip failed") or a numerical exception in the reference
        dc.proximity(cone, caps, pdip_tol=-1.0)


def test_device_memory_mode(ctx, oracle):
    """mem=DEVICE: torch CUDA tensors in, results left in HBM (the path bench.py's `value` times)."""
    import torch
    n = 50000
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, 1, seed=5)
    ctx.set_shape_arrays(*t.arrays())
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ds1, ds2, dr1, da1, dr2, da2 = T(s1), T(s2), T(r1), T(a1), T(r2), T(a2)
    alpha = torch.empty(n, dtype=torch.float64, device=dev); x = torch.empty((n, 3), dtype=torch.float64, device=dev)
    J = torch.empty((n, 12, 4), dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev); it = torch.empty(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.query_raw(2, n, ds1, ds2, dr1, da1, dr2, da2, 1, 1e-6, dc.MEM_DEVICE, alpha, x, J, st, it)
    torch.cuda.synchronize()
    h = ctx._batch(2, s1, s2, r1, a1, r2, a2, 1, 1e-6)
    assert np.array_equal(alpha.cpu().numpy(), h["alpha"])           # same kernels, same bits
    assert np.array_equal(J.cpu().numpy().transpose(0, 2, 1), h["J"], equal_nan=True)
    assert ctx.launch_count() > 0 and ctx.last_kernel_ms() > 0


def test_full_size_properties(ctx, oracle):
    """BASELINE config 2 at full size (1M mixed pairs): size-independent properties — argument-order symmetry of alpha
    (test/proximity_test.jl:12-15), structural Jacobian identities, permutation invariance (binning must not change
    per-pair results) — and, because the GPU box's host cores make it affordable, the oracle on all 1M pairs."""
    n = 1_000_000
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, 0, seed=2024)
    ctx.set_shape_arrays(*t.arrays())
    a = ctx._batch(0, s1, s2, r1, a1, r2, a2, 0, 1e-6)
    b = ctx._batch(0, s2, s1, r2, a2, r1, a1, 0, 1e-6)
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert np.abs(a["alpha"] - b["alpha"]).max() < 1e-4
    perm = util.rng(1).permutation(n)
    c = ctx._batch(0, s1[perm], s2[perm], r1[perm], a1[perm], r2[perm], a2[perm], 0, 1e-6)
    assert np.array_equal(c["alpha"], a["alpha"][perm]) and np.array_equal(c["x"], a["x"][perm])
    m = 200_000
    j = ctx._batch(2, s1[:m], s2[:m], r1[:m], a1[:m], r2[:m], a2[:m], 0, 1e-6)
    np.testing.assert_allclose(j["alpha"], a["alpha"][:m], rtol=1e-10)   # different kernel variant, same algorithm
    props.structural_jacobian_properties(j, a1[:m], a2[:m], 0)
    o = oracle.query(0, t.arrays(), s1, s2, r1, a1, r2, a2, 0, 1e-6)
    assert (o["status"] == 0).all() and (a["iters"] == o["iters"]).mean() > 0.9999
    assert util.rel_err(a["alpha"], o["alpha"]).max() < TOL_AX
    assert ((a["alpha"] <= 1) == (o["alpha"] <= 1)).all()
    ex = util.rel_err(a["x"], o["x"])
    assert (ex < TOL_AX).mean() > 0.9999 and ex.max() < 1e-3
    oj = oracle.query(2, t.arrays(), s1[:m], s2[:m], r1[:m], a1[:m], r2[:m], a2[:m], 0, 1e-6)
    ej = util.rel_err(j["J"], oj["J"])
    assert (ej < TOL_J).mean() > 0.9999 and np.quantile(ej, 0.99999) < 1e-2


def test_async_submit_matches_sync(ctx):
    """dcol_submit / dcol_wait with three submissions in flight return the same bits as the synchronous calls."""
    import torch
    n = 300000
    sets = []
    for k, (kind, att) in enumerate(((0, 0), (2, 1), (1, 0), (0, 0), (2, 0))):
        t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att, seed=50 + k)
        sets.append((kind, att, t, s1, s2, r1, a1, r2, a2))
    ctx.set_shape_arrays(*sets[0][2].arrays())          # all sets use the same 7 shapes per attitude kind
    outs, tickets = [], []
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    for kind, att, t, s1, s2, r1, a1, r2, a2 in sets:
        na = 4 if att == 0 else 3; nt = 6 + 2 * na
        ins = [pin(a) for a in (s1, s2, r1, a1, r2, a2)]
        o = {"alpha": torch.empty(n, dtype=torch.float64).pin_memory(), "x": torch.empty((n, 3), dtype=torch.float64).pin_memory(),
             "deriv": (torch.empty((n, nt) if kind == 1 else (n, nt, 4), dtype=torch.float64).pin_memory() if kind else None),
             "status": torch.empty(n, dtype=torch.int32).pin_memory(), "iters": torch.empty(n, dtype=torch.int32).pin_memory(), "ins": ins}
        if len(tickets) >= 3:
            ctx.wait(tickets[-3])
        tickets.append(ctx.submit_raw(kind, n, *ins, att, 1e-6, o["alpha"], o["x"] if kind != 1 else None, o["deriv"], o["status"], o["iters"]))
        outs.append(o)
    for tk in tickets[-3:]:
        ctx.wait(tk)
    for (kind, att, t, s1, s2, r1, a1, r2, a2), o in zip(sets, outs):
        ref = ctx._batch(kind, s1, s2, r1, a1, r2, a2, att, 1e-6)
        assert np.array_equal(o["alpha"].numpy(), ref["alpha"]) and np.array_equal(o["status"].numpy(), ref["status"])
        assert np.array_equal(o["iters"].numpy(), ref["iters"])
        if kind == 0:
            assert np.array_equal(o["x"].numpy(), ref["x"])
        if kind == 1:
            assert np.array_equal(o["deriv"].numpy(), ref["grad"], equal_nan=True)
        if kind == 2:
            assert np.array_equal(o["deriv"].numpy().transpose(0, 2, 1), ref["J"], equal_nan=True)


def test_registered_host_buffers(ctx):
    """dcol_host_register: numpy arrays page-locked in place give the same bits as pageable ones (and may be reused);
    bad arguments are API errors, not crashes."""
    import time
    n = 200000
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, 1, seed=77)
    ctx.set_shape_arrays(*t.arrays())
    ref = ctx.proximity_jacobian_batch(s1, s2, r1, a1, r2, a2, att_kind=1)
    ins = [ctx.pin(np.ascontiguousarray(a)) for a in (s1, s2, r1, a1, r2, a2)]
    out = ctx.alloc_outputs(2, n, att_kind=1, pinned=True)
    try:
        t_pin = []
        for rep in range(3):
            for a in out.values():
                a.fill(0)
            t0 = time.perf_counter()
            res = ctx.proximity_jacobian_batch(*ins, att_kind=1, out=out)
            t_pin.append(time.perf_counter() - t0)
            assert res["alpha"] is out["alpha"]
            assert np.array_equal(res["alpha"], ref["alpha"]) and np.array_equal(res["x"], ref["x"])
            assert np.array_equal(res["J"], ref["J"], equal_nan=True) and np.array_equal(res["iters"], ref["iters"])
        with pytest.raises(RuntimeError):
            ctx.pin(out["alpha"])                       # already registered
        with pytest.raises(ValueError):
            ctx.proximity_gradient_batch(*ins, att_kind=1, out=out)   # buffers of another query kind
    finally:
        for a in ins + list(out.values()):
            ctx.unpin(a)
    with pytest.raises(RuntimeError):
        ctx.unpin(out["alpha"])                         # not registered any more


@pytest.mark.parametrize("att_kind", [0, 1])
def test_random_shape_parameters(ctx, oracle, att_kind):
    """Random geometry instead of the reference's fixed test shapes: 6 shapes per kind with random sizes, random bounded
    polytopes (5-20 faces) and polygons (3-9 sides), half of them with random r_offset / Q_offset."""
    g = util.rng(2025 + att_kind)
    C = lambda b: util.cls_name(att_kind, b)
    shapes = []
    for k in range(6):
        nf = int(g.integers(5, 21))
        dirs = np.vstack([np.eye(3), -np.eye(3), g.standard_normal((max(nf - 6, 0), 3))])[:max(nf, 6)]
        dirs = dirs * (0.5 + g.random((dirs.shape[0], 1)))                  # rows are NOT unit-normalised (like the fixtures)
        shapes.append(C("Polytope")(dirs, (0.3 + g.random(dirs.shape[0])) * np.linalg.norm(dirs, axis=1)))
        shapes.append(C("Capsule")(0.1 + g.random(), 0.2 + 3 * g.random()))
        shapes.append(C("Cylinder")(0.1 + g.random(), 0.2 + 3 * g.random()))
        shapes.append(C("Cone")(0.5 + 2.5 * g.random(), np.deg2rad(10 + 40 * g.random())))
        shapes.append(C("Sphere")(0.1 + g.random()))
        M = g.standard_normal((3, 3)); shapes.append(C("Ellipsoid")(M @ M.T + 0.3 * np.eye(3)))
        nh = int(g.integers(3, 10)); A, b = dc.create_n_sided(nh, 0.2 + g.random())
        shapes.append(C("Polygon")(A, b * (0.7 + 0.6 * g.random(nh)), 0.05 + 0.3 * g.random()))
    for s in shapes[::2]:
        s.r_offset = 0.2 * g.standard_normal(3)
        s.Q_offset = dc.dcm_from_q(g.standard_normal(4))
    t = dc.ShapeTable(); ids = np.array([t.add(s) for s in shapes], np.int32)
    n = 30000
    s1 = ids[g.integers(0, len(ids), n)]; s2 = ids[g.integers(0, len(ids), n)]
    r1, a1, r2, a2 = util.random_poses(n, att_kind, 314)
    for kind in (0, 1, 2):
        gg = gq(ctx)(kind, t.arrays(), s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
        o = oracle.query(kind, t.arrays(), s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
        assert (o["status"] == 0).mean() > 0.999
        compare(gg, o, kind, strict_frac=0.99)


def test_straggler_handover_is_bit_identical(ctx):
    """Parking the last stragglers of a tile and finishing them in resume passes must not change a single bit:
    no hand-over (DCOL_SPILL=0), the default (one resume level) and three resume levels give identical outputs."""
    import os
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(300000, 0, seed=99, variant="scene")
    ctx.set_shape_arrays(*t.arrays())
    outs = []
    saved = {k: os.environ.get(k) for k in ("DCOL_SPILL", "DCOL_LEVELS")}
    try:
        for env in ({"DCOL_SPILL": "0"}, {}, {"DCOL_LEVELS": "3"}, {"DCOL_SPILL": "2", "DCOL_LEVELS": "2"}):
            for k in saved:
                os.environ.pop(k, None)
            os.environ.update(env)
            outs.append([ctx._batch(kind, s1, s2, r1, a1, r2, a2, 0, 1e-6) for kind in (0, 2)])
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            for key in a:
                assert np.array_equal(a[key], b[key], equal_nan=True), key


def test_convergence_on_the_last_allowed_step(ctx, oracle):
    """static_solver2.jl:132,143-146,188: mu is tested only at the START of iterations 1..cap and "pdip failed" is raised
    after the last step without another test, so a pair whose mu drops below pdip_tol on exactly the last allowed step
    FAILS.  The cap is 50 in the reference; the testing knob lowers it on both sides so that thousands of pairs land on
    that edge robustly (mu falls by orders of magnitude per step there).  Status and iteration counts bit-identical."""
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(20000, 0, seed=31)
    ta = t.arrays()
    full = oracle.query(0, ta, s1, s2, r1, a1, r2, a2, 0, 1e-6)
    try:
        for cap in (6, 8, 10):
            oracle.set_max_iters(cap); ctx.set_max_iters(cap)
            o = oracle.query(0, ta, s1, s2, r1, a1, r2, a2, 0, 1e-6)
            g = gq(ctx)(0, ta, s1, s2, r1, a1, r2, a2, 0, 1e-6)
            edge = full["iters"] == cap                       # converge on exactly the cap-th step
            assert edge.sum() > 100
            assert (o["status"][edge] == 1).all() and (o["iters"][edge] == cap).all()
            assert (o["status"][full["iters"] < cap] == 0).all()
            assert np.array_equal(g["status"], o["status"])
            assert (g["iters"] == o["iters"]).mean() > 0.999 and np.array_equal(g["iters"][edge], o["iters"][edge])
    finally:
        oracle.set_max_iters(50); ctx.set_max_iters(50)


def test_many_faces(ctx, oracle):
    """Polytopes / polygons beyond 32 rows (the oracle's round-1 limit): 48-, 64- and 80-face polytopes, 40-sided polygon
    (a pair may carry up to 172 orthant rows: one warp per CTA with the whole shared memory as its row store)."""
    g = util.rng(18)
    shapes = []
    for nf in (48, 64, 80):
        d = g.standard_normal((nf, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        shapes.append(dc.Polytope(d, 0.6 + 0.4 * g.random(nf)))
    A, b = dc.create_n_sided(40, 0.7)
    shapes += [dc.Polygon(A, b, 0.1), dc.Capsule(0.3, 1.2), dc.Cone(2.0, np.deg2rad(22))]
    t = dc.ShapeTable(); ids = np.array([t.add(s) for s in shapes], np.int32)
    n = 3000
    s1 = ids[g.integers(0, len(ids), n)]; s2 = ids[g.integers(0, len(ids), n)]
    r1, a1, r2, a2 = util.random_poses(n, 0, 321)
    for kind in (0, 2):
        gg = gq(ctx)(kind, t.arrays(), s1, s2, r1, a1, r2, a2, 0, 1e-6)
        o = oracle.query(kind, t.arrays(), s1, s2, r1, a1, r2, a2, 0, 1e-6)
        assert (o["status"] == 0).all()
        compare(gg, o, kind, strict_frac=0.995)


def test_final_mu_output(ctx):
    """dcol_set_mu_output: mu = s'z/deg at the iteration that ended the solve (< pdip_tol exactly for converged pairs)."""
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(6000, 1, seed=3)
    ctx.set_shape_arrays(*t.arrays())
    g = ctx._batch(1, s1, s2, r1, a1, r2, a2, 1, 1e-6, want_mu=True)
    assert (g["status"] == 0).all() and (g["mu"] < 1e-6).all() and (g["mu"] > 0).all()
    try:
        ctx.set_max_iters(5)
        h = ctx._batch(0, s1, s2, r1, a1, r2, a2, 1, 1e-6, want_mu=True)
    finally:
        ctx.set_max_iters(50)
    failed = h["status"] == 1
    assert failed.any() and (h["mu"][~failed] < 1e-6).all() and np.isfinite(h["mu"]).all()
    small = ctx._batch(0, s1[:100], s2[:100], r1[:100], a1[:100], r2[:100], a2[:100], 1, 1e-6, want_mu=True)   # small-batch path
    big = ctx._batch(0, s1, s2, r1, a1, r2, a2, 1, 1e-6, want_mu=True)
    assert np.array_equal(small["mu"], big["mu"][:100])


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_small_batch_path_matches_the_binned_path(ctx, kind):
    """n <= 16384 runs the single-block preparation without hand-over and un-permute pass; results are bit-identical to the
    same pairs inside a large batch (sorted, handed over, un-permuted)."""
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(30000, 0, seed=11, variant="scene")
    ctx.set_shape_arrays(*t.arrays())
    big = ctx._batch(kind, s1, s2, r1, a1, r2, a2, 0, 1e-6)
    for n in (1, 2, 33, 4096, 16384):
        sm = ctx._batch(kind, s1[:n], s2[:n], r1[:n], a1[:n], r2[:n], a2[:n], 0, 1e-6)
        for k in sm:
            assert np.array_equal(sm[k], big[k][:n], equal_nan=True), (k, n)


def test_generic_socp_entry_on_the_reference_fixture(ctx, oracle):
    """dcol_solve_socp (the solver subsystem on its own, static_solver2.jl:106-191) on the reference's stored problem
    example_socp_2.jld2 with the property its own test asserts (test/solver_tests.jl:63-64: |s'z| < 1e-10 at
    pdip_tol = 1e-12), against the oracle's iterates, and on dense (c, G, h) of mixed primitive pairs against the
    proximity entry point that never forms G."""
    d = json.load(open(os.path.join(util.GOLDEN, "example_socp_2.json")))
    Gm = np.array(d["G"], float); h = np.array(d["h"], float); c = np.array(d["c"], float)
    x, s, z, st, it = ctx.solve_socp(c, Gm, h, d["n_ort"], d["n_soc1"], d["n_soc2"], pdip_tol=1e-12)
    assert st == 0 and abs(s @ z) < 1e-10
    ox, os_, oz, oit, ost = oracle.solve_socp(c, Gm, h, d["n_ort"], d["n_soc1"], d["n_soc2"], pdip_tol=1e-12)
    assert it == oit and ost == 0
    np.testing.assert_allclose(x, ox, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(s, os_, rtol=1e-6, atol=1e-9); np.testing.assert_allclose(z, oz, rtol=1e-6, atol=1e-9)
    # batched: the dense problems of random pairs of one structure (capsule-cone: nx 5, 3 orthant rows, SOC 4 and 3)
    g = util.rng(6)
    caps = dc.Capsule(0.3, 1.2); cone = dc.Cone(2.0, np.deg2rad(22))
    t = dc.ShapeTable(); i1 = t.add(caps); i2 = t.add(cone)
    n = 200
    r1, a1, r2, a2 = util.random_poses(n, 0, 5)
    Gs, hs = [], []
    for k in range(n):
        caps.r, caps.q, cone.r, cone.q = r1[k], a1[k], r2[k], a2[k]
        cc, Gk, hk, n_ort, d1, d2 = oracle.pair_problem(caps, cone)
        Gs.append(Gk); hs.append(hk)
    X, S, Z, stt, its = ctx.solve_socp(cc, np.array(Gs), np.array(hs), n_ort, d1, d2, pdip_tol=1e-6)
    ctx.set_shape_arrays(*t.arrays())
    p = ctx._batch(0, np.full(n, i1, np.int32), np.full(n, i2, np.int32), r1, a1, r2, a2, 0, 1e-6)
    assert (stt == 0).all() and np.array_equal(its, p["iters"])
    np.testing.assert_allclose(X[:, 3], p["alpha"], rtol=1e-9)
    np.testing.assert_allclose(X[:, :3], p["x"], rtol=1e-6, atol=1e-8)
    # dimensions out of range are an API error
    with pytest.raises(RuntimeError):
        ctx.solve_socp(np.zeros(20), np.zeros((30, 20)), np.zeros(30), 30, 0, 0)


def test_verbose_trace(ctx, capsys):
    """verbose = true (static_solver2.jl:117-120, 179-183): dcol_solve_socp_trace records, per iteration, the five numbers
    the reference prints.  The last row describes the returned iterate, so it can be recomputed from (x, s, z); rows past
    the last iteration are untouched; the single-pair front ends print the table and return the fast path's values."""
    d = json.load(open(os.path.join(util.GOLDEN, "example_socp_2.json")))
    Gm = np.array(d["G"], float); h = np.array(d["h"], float); c = np.array(d["c"], float)
    x, s, z, st, it, tr = ctx.solve_socp(c, Gm, h, d["n_ort"], d["n_soc1"], d["n_soc2"], pdip_tol=1e-8, trace=True)
    x0, s0, z0, st0, it0 = ctx.solve_socp(c, Gm, h, d["n_ort"], d["n_soc1"], d["n_soc2"], pdip_tol=1e-8)
    assert st == 0 and it == it0 and np.array_equal(x, x0) and np.array_equal(s, s0) and np.array_equal(z, z0)
    assert np.isfinite(tr[:it]).all() and np.isnan(tr[it:]).all()
    last = tr[it - 1]
    assert last[0] == pytest.approx(c @ x, rel=1e-13)
    assert last[1] == pytest.approx(s @ z / (d["n_ort"] + 1), rel=1e-9)          # the reference divides by n_ort + 1 here (:181)
    assert last[2] == pytest.approx(np.linalg.norm(Gm @ x + s - h), abs=1e-12)
    assert (tr[:it, 4] > 0).all() and (tr[:it, 4] <= 1).all() and (np.diff(tr[:it, 1]) < 0).all()
    # a batch: every problem its own rows
    caps = dc.Capsule(0.3, 1.2); cone = dc.Cone(2.0, np.deg2rad(22))
    r1, a1, r2, a2 = util.random_poses(70, 0, 11)
    Gs, hs = [], []
    for k in range(70):
        caps.r, caps.q, cone.r, cone.q = r1[k], a1[k], r2[k], a2[k]
        cc, Gk, hk, n_ort, d1, d2 = dc.pair_problem(caps, cone)
        Gs.append(Gk); hs.append(hk)
    X, S, Z, stt, its, T = ctx.solve_socp(cc, np.array(Gs), np.array(hs), n_ort, d1, d2, pdip_tol=1e-6, trace=True)
    assert (stt == 0).all() and len(set(its.tolist())) > 1
    for k in range(70):
        assert np.isfinite(T[k, :its[k]]).all() and np.isnan(T[k, its[k]:]).all()
        assert T[k, its[k] - 1, 0] == pytest.approx(X[k, 3], rel=1e-13)
    # the front end: header, one row per iteration, objective of the last row = alpha of the fast path
    capsys.readouterr()
    alpha, xw = dc.proximity(caps, cone, pdip_tol=1e-6, verbose=True, ctx=ctx)
    out = capsys.readouterr().out.strip().split("\n")
    assert out[0].startswith("iter     objv") and set(out[1]) == {"-"} and len(out) == 2 + its[69]
    assert int(out[-1].split()[0]) == its[69] and float(out[-1].split()[1]) == pytest.approx(alpha, rel=2e-3)
    a2_, g = dc.proximity_gradient(caps, cone, verbose=True, ctx=ctx)
    assert len(capsys.readouterr().out.strip().split("\n")) == len(out) and a2_ == pytest.approx(alpha, rel=1e-9)


@pytest.mark.parametrize("att_kind", [0, 1])
def test_single_pair_one_launch_path(ctx, att_kind):
    """One pair on host pointers runs as ONE launch on a record the host packs itself (proximity_single): bit-identical to
    the same pair inside a two-pair batch (preparation kernel + persistent kernel) for all 49 ordered type pairs (both class
    orders: the host applies the exchange the packing kernels apply), every query kind, shapes with offsets; and the same
    error for a bad shape index."""
    g = util.rng(41 + att_kind)
    shapes = util.standard_shapes(att_kind, "scene")
    for p in shapes:
        p.r_offset = 0.2 * g.standard_normal(3); p.Q_offset = dc.dcm_from_q(g.standard_normal(4))
    t = dc.ShapeTable(); ids = [t.add(p) for p in shapes]
    ctx.set_shape_arrays(*t.arrays())
    r1, a1, r2, a2 = util.random_poses(49, att_kind, seed=77)
    k = 0
    for i in ids:
        for j in ids:
            for kind in (0, 1, 2):
                one = ctx._batch(kind, [i], [j], r1[k:k + 1], a1[k:k + 1], r2[k:k + 1], a2[k:k + 1], att_kind, 1e-6, want_mu=(kind == 0))
                two = ctx._batch(kind, [i, i], [j, j], r1[[k, k]], a1[[k, k]], r2[[k, k]], a2[[k, k]], att_kind, 1e-6, want_mu=(kind == 0))
                assert one["status"][0] == 0
                for key in one:
                    assert np.array_equal(one[key][0], two[key][0], equal_nan=True), (i, j, kind, key)
            k += 1
    # a handful of pairs (up to 128 on host pointers): packed on the host as well, one block per class pair present --
    # bit-identical to the same pairs inside a batch that takes the preparation-kernel path (129 pairs and more)
    n = 300
    tm, s1, s2, q1, b1, q2, b2 = util.mixed_batch(n, att_kind, seed=9, variant="scene", table=t)
    ctx.set_shape_arrays(*t.arrays())
    for kind in (0, 1, 2):
        big = ctx._batch(kind, s1, s2, q1, b1, q2, b2, att_kind, 1e-6)
        for m in (2, 3, 31, 32, 33, 78, 128):
            sm = ctx._batch(kind, s1[:m], s2[:m], q1[:m], b1[:m], q2[:m], b2[:m], att_kind, 1e-6)
            for key in sm:
                assert np.array_equal(sm[key], big[key][:m], equal_nan=True), (kind, m, key)
        one_class = ctx._batch(kind, np.full(128, s1[0]), np.full(128, s2[0]), q1[:128], b1[:128], q2[:128], b2[:128], att_kind, 1e-6)   # one group of 128
        ref = ctx._batch(kind, np.full(n, s1[0]), np.full(n, s2[0]), q1, b1, q2, b2, att_kind, 1e-6)
        for key in one_class:
            assert np.array_equal(one_class[key], ref[key][:128], equal_nan=True), (kind, key)
    with pytest.raises(RuntimeError, match="shape index out of range"):
        ctx._batch(0, [0, len(ids) + 40, 1], [0, 1, 2], r1[:3], a1[:3], r2[:3], a2[:3], att_kind, 1e-6)
    with pytest.raises(RuntimeError, match="shape index out of range"):
        ctx._batch(0, [len(ids) + 40], [0], r1[:1], a1[:1], r2[:1], a2[:1], att_kind, 1e-6)
    ok = ctx._batch(0, [0], [1], r1[:1], a1[:1], r2[:1], a2[:1], att_kind, 1e-6)          # the context stays usable
    assert ok["status"][0] == 0
