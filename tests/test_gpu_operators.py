"""GPU (-m gpu): the next-tier operators of SURVEY.md §8f — pose-table (indexed) queries and the all-pairs contact scene —
against the plain batched calls, the oracle, and at BASELINE config 4 / 5 sizes."""
import itertools

import numpy as np
import pytest

import dcol_b200 as dc
import props
import util

pytestmark = pytest.mark.gpu


def scene_bodies(n_bodies, att_kind, seed=3, spread=5.0):
    """Types cycled over the 12-body list of examples/contact_physics/ncp_contact_simulation.jl:264-267, positions
    spread*N(0,I) (:281 uses 5)."""
    A1, b1, A2, b2 = util.load_polytopes()
    A5, b5 = dc.create_n_sided(5, 0.6)
    A8, b8 = dc.create_n_sided(8, 0.8)
    g = lambda b: util.cls_name(att_kind, b)
    protos = [g("Cylinder")(0.6, 3.0), g("Capsule")(0.2, 5.0), g("Sphere")(0.8), g("Cone")(2.0, np.deg2rad(22)),
              g("Polytope")(A2, b2), g("Polygon")(A5, b5, 0.2), g("Cylinder")(1.1, 2.3), g("Capsule")(0.8, 1.0), g("Sphere")(0.5),
              g("Cone")(3.0, np.deg2rad(18)), g("Polytope")(A1, b1), g("Polygon")(A8, b8, 0.15)]
    t = dc.ShapeTable()
    ids = np.array([t.add(p) for p in protos], np.int32)
    shape = ids[np.arange(n_bodies) % len(ids)]
    rg = util.rng(seed)
    r = spread * rg.standard_normal((n_bodies, 3))
    q = rg.standard_normal((n_bodies, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    if att_kind == dc.ATT_MRP:
        q = np.where(q[:, :1] < 0, -q, q)
        att = q[:, 1:] / (1 + q[:, :1])
    else:
        att = q
    return t, shape, r, att


@pytest.mark.parametrize("att_kind", [0, 1])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_scene_allpairs_matches_explicit_pairs(ctx, oracle, att_kind, kind):
    nb = 61
    t, shape, r, att = scene_bodies(nb, att_kind, spread=2.0)
    ctx.set_shape_arrays(*t.arrays())
    s = ctx.scene_allpairs(kind, shape, r, att, att_kind)
    ij = np.array(list(itertools.combinations(range(nb), 2)))          # Combinatorics.combinations order
    assert s["alpha"].shape[0] == len(ij) == nb * (nb - 1) // 2
    e = ctx._batch(kind, shape[ij[:, 0]], shape[ij[:, 1]], r[ij[:, 0]], att[ij[:, 0]], r[ij[:, 1]], att[ij[:, 1]], att_kind, 1e-6)
    for k in e:
        assert np.array_equal(s[k], e[k], equal_nan=True), k                # same kernels, same pairs, same order
    o = oracle.query(kind, t.arrays(), shape[ij[:, 0]], shape[ij[:, 1]], r[ij[:, 0]], att[ij[:, 0]], r[ij[:, 1]], att[ij[:, 1]], att_kind, 1e-6)
    ok = o["status"] == 0
    assert (s["status"] == o["status"]).all()
    assert util.rel_err(s["alpha"], o["alpha"])[ok].max() < 1e-6
    assert ((s["alpha"] <= 1) == (o["alpha"] <= 1))[ok].all()
    if kind == 1:
        assert (util.rel_err(s["grad"], o["grad"])[ok] < 1e-5).mean() > 0.99
    if kind == 2:
        assert (util.rel_err(s["J"], o["J"])[ok] < 1e-5).mean() > 0.99


def test_indexed_query_matches_replicated_poses(ctx, oracle):
    """BASELINE config 5 shape: vehicle = create_rect_prism(2.5, .15, .01) against the three boxes of piano_mover.jl:74-82,
    MRP p = (0, 0, tan(theta/4)) (:41); pose table = (rollout, knot), obstacle table = 3 static poses."""
    B, N, K = 64, 100, 3
    veh = dc.create_rect_prism(2.5, 0.15, 0.01)[0]
    obs = [dc.create_rect_prism(3.0, 3.0, 0.02)[0], dc.create_rect_prism(4.0, 1.0, 0.02)[0], dc.create_rect_prism(1.0, 3.0, 0.02)[0]]
    t = dc.ShapeTable()
    sv = t.add(veh); so = [t.add(o) for o in obs]
    ctx.set_shape_arrays(*t.arrays())
    g = util.rng(17)
    theta = g.uniform(-np.pi, np.pi, B * N)
    r1 = np.stack([g.uniform(-4, 4, B * N), g.uniform(-4, 4, B * N), np.zeros(B * N)], axis=1)
    p1 = np.stack([np.zeros(B * N), np.zeros(B * N), np.tan(theta / 4)], axis=1)
    r2 = np.array([[-1.5, 1.5, 0.0], [2.0, -2.5, 0.0], [3.0, 2.5, 0.0]]); p2 = np.zeros((3, 3))
    pose1 = np.repeat(np.arange(B * N, dtype=np.int32), K)
    pose2 = np.tile(np.arange(K, dtype=np.int32), B * N)
    s1 = np.full(B * N * K, sv, np.int32); s2 = np.array(so, np.int32)[pose2]
    for kind in (0, 1, 2):
        a = ctx.query_indexed(kind, s1, s2, pose1, pose2, r1, p1, r2, p2, dc.ATT_MRP)
        b = ctx._batch(kind, s1, s2, r1[pose1], p1[pose1], r2[pose2], p2[pose2], dc.ATT_MRP, 1e-6)
        for k in a:
            assert np.array_equal(a[k], b[k], equal_nan=True), k
    o = oracle.query(2, t.arrays(), s1, s2, r1[pose1], p1[pose1], r2[pose2], p2[pose2], dc.ATT_MRP, 1e-6)
    ok = (o["status"] == 0)
    assert (a["status"] == o["status"]).all() and ok.mean() > 0.99
    assert util.rel_err(a["alpha"], o["alpha"])[ok].max() < 1e-6
    # thin plates in a plane: many face-parallel (degenerate) configurations, so assert J on the pass fraction
    assert (util.rel_err(a["J"], o["J"])[ok] < 1e-5).mean() > 0.97
    with pytest.raises(IndexError):
        ctx.query_indexed(0, s1, s2, pose1 + B * N, pose2, r1, p1, r2, p2, dc.ATT_MRP)


def test_config4_full_size_scene(ctx):
    """BASELINE config 4: 4096 mixed bodies, all 8 386 560 i<j pairs, proximity_gradient, quaternion.  Size-independent
    properties: translation invariance (d alpha/d r_i + d alpha/d r_j = 0), tangency of the quaternion blocks, and
    agreement with explicitly listed pairs on a random sample."""
    nb = 4096
    t, shape, r, att = scene_bodies(nb, 0, spread=5.0)
    ctx.set_shape_arrays(*t.arrays())
    s = ctx.scene_allpairs(1, shape, r, att, 0)
    n = nb * (nb - 1) // 2
    assert s["alpha"].shape == (n,) and s["grad"].shape == (n, 14)
    ok = s["status"] == 0
    assert ok.mean() > 0.9999
    gsum = s["grad"][:, 0:3] + s["grad"][:, 7:10]
    assert np.abs(gsum[ok]).max() < 1e-6 * max(1.0, np.abs(s["grad"][ok]).max())
    rg = util.rng(5)
    k = rg.integers(0, n, 20000)
    # invert the combinations index on the host
    i = np.floor(((2 * nb - 1) - np.sqrt((2 * nb - 1) ** 2 - 8.0 * k)) / 2).astype(np.int64)
    i = np.where(i * (2 * nb - i - 1) // 2 > k, i - 1, i)
    i = np.where((i + 1) * (2 * nb - i - 2) // 2 <= k, i + 1, i)
    j = k - i * (2 * nb - i - 1) // 2 + i + 1
    assert (i < j).all() and (j < nb).all()
    e = ctx._batch(1, shape[i], shape[j], r[i], att[i], r[j], att[j], 0, 1e-6)
    assert np.array_equal(s["alpha"][k], e["alpha"]) and np.array_equal(s["grad"][k], e["grad"], equal_nan=True)
    tq = np.einsum("nj,nj->n", s["grad"][k, 3:7], att[i])
    assert np.abs(tq[e["status"] == 0]).max() < 1e-8 * max(1.0, np.abs(e["grad"]).max())


def wrench_reference(grad, lam, att, ij, att_kind):
    """The accumulation of examples/contact_physics/ncp_contact_simulation.jl:133-139 written as the example writes it
    (dense Gbar = blockdiag(I3, G(q_i), I3, G(q_j)), G = L(q) H from variational_utils.jl:1-26), per pair, in float64."""
    nb = att.shape[0]
    out = np.zeros((nb, 6))

    def hat(w):
        return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])

    def G(q):
        L = np.zeros((4, 4)); L[0, 0] = q[0]; L[0, 1:] = -q[1:]; L[1:, 0] = q[1:]; L[1:, 1:] = q[0] * np.eye(3) + hat(q[1:])
        H = np.vstack([np.zeros((1, 3)), np.eye(3)])
        return L @ H
    for k, (i, j) in enumerate(ij):
        if att_kind == dc.ATT_QUAT:
            Gbar = np.zeros((14, 12)); Gbar[0:3, 0:3] = np.eye(3); Gbar[3:7, 3:6] = G(att[i]); Gbar[7:10, 6:9] = np.eye(3); Gbar[10:14, 9:12] = G(att[j])
            E = (grad[k][None, :] @ Gbar).T[:, 0] * lam[k]
        else:
            E = grad[k] * lam[k]
        out[i] += E[:6]; out[j] += E[6:]
    return out


@pytest.mark.parametrize("att_kind", [0, 1])
def test_scene_wrench_matches_per_pair_accumulation(ctx, oracle, att_kind):
    """dcol_scene_wrench against (a) the same accumulation done on the host from dcol_scene_allpairs' gradient rows and
    (b) the oracle's gradients pushed through the example's dense formula."""
    nb = 47
    t, shape, r, att = scene_bodies(nb, att_kind, spread=2.0)
    ctx.set_shape_arrays(*t.arrays())
    ij = np.array(list(itertools.combinations(range(nb), 2)))
    lam = util.rng(11).uniform(0.0, 2.0, len(ij))
    w = ctx.scene_wrench(shape, r, att, lam, att_kind)
    s = ctx.scene_allpairs(1, shape, r, att, att_kind)
    assert np.array_equal(w["alpha"], s["alpha"]) and np.array_equal(w["status"], s["status"]) and np.array_equal(w["iters"], s["iters"])
    assert (s["status"] == 0).all()
    ref = wrench_reference(s["grad"], lam, att, ij, att_kind)
    scale = np.abs(ref).max()
    assert np.abs(w["wrench"] - ref).max() < 1e-12 * scale            # same gradients, different summation order
    o = oracle.query(1, t.arrays(), shape[ij[:, 0]], shape[ij[:, 1]], r[ij[:, 0]], att[ij[:, 0]], r[ij[:, 1]], att[ij[:, 1]], att_kind, 1e-6)
    ref_o = wrench_reference(o["grad"], lam, att, ij, att_kind)
    assert np.abs(w["wrench"] - ref_o).max() < 1e-5 * scale
    w2 = ctx.scene_wrench(shape, r, att, lam, att_kind, per_pair=False)          # NULL per-pair outputs
    assert np.array_equal(w2["wrench"], w["wrench"]) and "alpha" not in w2
    # degenerate scenes
    one = ctx.scene_wrench(shape[:1], r[:1], att[:1], np.zeros(0), att_kind)
    assert one["wrench"].shape == (1, 6) and (one["wrench"] == 0).all()
    with pytest.raises(ValueError):
        ctx.scene_wrench(shape, r, att, lam[:-1], att_kind)
    bad = shape.copy(); bad[3] = 12345
    with pytest.raises(RuntimeError, match="shape index out of range"):
        ctx.scene_wrench(bad, r, att, lam, att_kind)


def test_scene_wrench_config4_size(ctx):
    """BASELINE config 4 (4096 bodies, 8 386 560 pairs) through the fused operator: finite, and equal to the accumulation
    of the per-pair rows on a sample of bodies."""
    nb = 4096
    t, shape, r, att = scene_bodies(nb, 0, spread=30.0)
    ctx.set_shape_arrays(*t.arrays())
    n = nb * (nb - 1) // 2
    lam = util.rng(5).uniform(0.0, 1.0, n)
    w = ctx.scene_wrench(shape, r, att, lam, 0, per_pair=True)
    assert (w["status"] == 0).mean() > 0.9999
    ok_body = np.isfinite(w["wrench"]).all(axis=1)
    assert ok_body.mean() > 0.99
    s = ctx.scene_allpairs(1, shape, r, att, 0)
    assert np.array_equal(s["alpha"], w["alpha"])
    k_of = lambda i, j: i * (2 * nb - i - 1) // 2 + (j - i - 1)
    for b in (0, 1, 2047, 4095):
        acc = np.zeros(7)
        for j in range(nb):
            if j == b:
                continue
            k = k_of(b, j) if j > b else k_of(j, b)
            acc += lam[k] * (s["grad"][k, :7] if j > b else s["grad"][k, 7:])
        q = att[b]; g0, gv = acc[3], acc[4:7]
        ref = np.concatenate([acc[:3], -q[1:] * g0 + q[0] * gv - np.cross(q[1:], gv)])
        if np.isfinite(ref).all():
            assert np.abs(w["wrench"][b] - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


def test_se2_constraints_match_the_example_formulas(ctx, oracle):
    """dcol_se2_constraints against ineq_con_x / ineq_con_x_jac of piano_mover.jl:39-53 evaluated on the host from per-pair
    proximity_gradient results (GPU and oracle), and against central finite differences of c in the state."""
    n, K = 3000, 3
    veh = dc.create_rect_prism(2.5, 0.15, 0.01)[0]
    obs = [dc.create_rect_prism(3.0, 3.0, 1.0)[0], dc.create_rect_prism(4.0, 1.0, 1.0)[0], dc.create_rect_prism(1.0, 5.0, 1.1)[0]]   # :70-76
    t = dc.ShapeTable()
    sv = t.add(veh); so = np.array([t.add(o) for o in obs], np.int32)
    ctx.set_shape_arrays(*t.arrays())
    g = util.rng(23)
    xyt = np.stack([g.uniform(-1, 6, n), g.uniform(-1, 6, n), g.uniform(-np.pi, np.pi, n)], axis=1)
    r_obs = np.array([[1.5, 3.5, 0.0], [4.5, 1.5, 0.0], [5.0, 4.0, 0.0]]); p_obs = np.zeros((K, 3))
    out = ctx.se2_constraints(xyt, sv, so, r_obs, p_obs)
    assert out["c"].shape == (n, K) and out["Jc"].shape == (n, K, 3)

    def example(query, X):
        m = X.shape[0]
        r1 = np.stack([X[:, 0], X[:, 1], np.zeros(m)], axis=1); p1 = np.stack([np.zeros(m), np.zeros(m), np.tan(X[:, 2] / 4)], axis=1)
        i1 = np.repeat(np.arange(m), K); i2 = np.tile(np.arange(K), m)
        res = query(np.full(m * K, sv, np.int32), so[i2], r1[i1], p1[i1], r_obs[i2], p_obs[i2])
        dp = 1.0 / (4 * np.cos(X[i1, 2] / 4) ** 2)                                   # dp_dθ, :46
        J = np.stack([-res["grad"][:, 0], -res["grad"][:, 1], -res["grad"][:, 5] * dp], axis=1)   # :50
        return (1 - res["alpha"]).reshape(m, K), J.reshape(m, K, 3), res["status"].reshape(m, K)
    c_g, J_g, st_g = example(lambda *a: ctx._batch(1, *a, dc.ATT_MRP, 1e-6), xyt)
    # the device evaluates tan(theta/4) itself (CUDA libm vs numpy: last-bit differences in p), so agreement is to round-off
    assert np.array_equal(out["status"], st_g)
    assert np.abs(out["c"] - c_g).max() < 1e-9
    okg = st_g == 0
    eg = np.abs(out["Jc"] - J_g).max(axis=2)[okg] / np.maximum(1.0, np.abs(J_g).max(axis=2)[okg])
    assert np.quantile(eg, 0.99) < 1e-8
    c_o, J_o, st_o = example(lambda *a: oracle.query(1, t.arrays(), *a, dc.ATT_MRP, 1e-6), xyt)
    ok = (st_o == 0) & (st_g == 0)
    assert ok.mean() > 0.99 and (st_o == st_g).all()
    assert np.abs(out["c"] - c_o)[ok].max() < 1e-6 * max(1.0, np.abs(c_o[ok]).max())
    err = np.abs(out["Jc"] - J_o).max(axis=2)[ok] / np.maximum(1.0, np.abs(J_o).max(axis=2)[ok])
    assert (err < 1e-5).mean() > 0.97                       # thin plates: degenerate face-parallel configurations in the tail
    # finite differences of the constraint itself (tight tolerance so that the difference quotient is meaningful)
    m = 200; eps = 1e-6
    fd = np.zeros((m, K, 3))
    for a in range(3):
        d = np.zeros(3); d[a] = eps
        cp = ctx.se2_constraints(xyt[:m] + d, sv, so, r_obs, p_obs, pdip_tol=1e-10)["c"]
        cm = ctx.se2_constraints(xyt[:m] - d, sv, so, r_obs, p_obs, pdip_tol=1e-10)["c"]
        fd[:, :, a] = (cp - cm) / (2 * eps)
    Jt = ctx.se2_constraints(xyt[:m], sv, so, r_obs, p_obs, pdip_tol=1e-10)["Jc"]
    e = np.abs(fd - Jt).max(axis=2) / np.maximum(1.0, np.abs(Jt).max(axis=2))
    assert np.median(e) < 1e-5 and (e < 1e-3).mean() > 0.95    # kinks where the active face changes
    empty = ctx.se2_constraints(np.zeros((0, 3)), sv, so, r_obs, p_obs)
    assert empty["c"].shape == (0, K)
    with pytest.raises(RuntimeError, match="shape index out of range"):
        ctx.se2_constraints(xyt[:10], 999, so, r_obs, p_obs)
    again = ctx.se2_constraints(xyt[:10], sv, so, r_obs, p_obs)              # the context stays usable after an API error
    assert np.array_equal(again["c"], out["c"][:10])


def test_se3_constraints_match_the_example_formulas(ctx, oracle):
    """dcol_se3_constraints against ineq_con_x / ineq_con_x_jac of cone_through_wall.jl:56-68 and
    cluttered_hallway_quadrotor.jl:83-104 (state [r; v; p; omega], rows [-J[1:3]' 0 -J[4:6]' 0]) evaluated on the host from
    per-pair proximity_gradient results: the same kernels on the same poses, so bit-identical; against the oracle; and
    against central finite differences of c in the six pose columns."""
    n = 2500
    g = util.rng(29)
    # the cone of cone_through_wall.jl against its four walls (:96-110), and the quadrotor sphere against the hallway set
    for veh, obs, r_obs, p_obs in (
            (dc.ConeMRP(2.0, np.deg2rad(22)),
             [dc.create_rect_prism(10.0, 10.0, 1.0)[0], dc.create_rect_prism(10.0, 10.0, 1.0)[0], dc.create_rect_prism(4.1, 4.1, 1.1)[0], dc.create_rect_prism(4.1, 4.1, 1.1)[0]],
             np.array([[-6, 0, 5.0], [6, 0, 5.0], [0, 0, 2.05], [0, 0, 7.96]]), np.tile(dc.mrp_from_q([np.cos(np.pi / 4), np.sin(np.pi / 4), 0, 0]), (4, 1))),
            (dc.SphereMRP(0.25),
             [dc.CylinderMRP(0.6, 3.0), dc.CapsuleMRP(0.2, 5.0), dc.ConeMRP(2.0, np.deg2rad(22)), dc.PolygonMRP(*dc.create_n_sided(5, 0.6), 0.2), dc.create_rect_prism(20, 5.0, 0.2)[0]],
             np.array([[-5, 0.3, 3.2], [-2.5, -0.4, 2.1], [0, 0.8, 3.9], [2.5, 0.1, 2.7], [0, 0, 0.9]]), 0.3 * g.standard_normal((5, 3)))):
        K = len(obs)
        t = dc.ShapeTable(); sv = t.add(veh); so = np.array([t.add(o) for o in obs], np.int32)
        ctx.set_shape_arrays(*t.arrays())
        X = np.zeros((n, 12))
        X[:, 0:3] = np.array([0.0, 0.0, 4.0]) + g.uniform(-7, 7, (n, 3)) * np.array([1.0, 1.0, 0.5])
        X[:, 3:6] = g.standard_normal((n, 3)); X[:, 9:12] = g.standard_normal((n, 3))       # velocities: ignored by the constraint
        q = g.standard_normal((n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True); q = np.where(q[:, :1] < 0, -q, q)
        X[:, 6:9] = q[:, 1:] / (1 + q[:, :1])
        out = ctx.se3_constraints(X, sv, so, r_obs, p_obs)
        assert out["c"].shape == (n, K) and out["Jc"].shape == (n, K, 6)
        i1 = np.repeat(np.arange(n), K); i2 = np.tile(np.arange(K), n)
        args = (np.full(n * K, sv, np.int32), so[i2], X[i1, 0:3], X[i1, 6:9], r_obs[i2], p_obs[i2])
        e = ctx._batch(1, *args, dc.ATT_MRP, 1e-6)
        assert np.array_equal(out["status"].reshape(-1), e["status"]) and (e["status"] == 0).all()
        assert np.array_equal(out["c"].reshape(-1), 1 - e["alpha"])
        assert np.array_equal(out["Jc"].reshape(-1, 6), -e["grad"][:, :6])
        o = oracle.query(1, t.arrays(), *args, dc.ATT_MRP, 1e-6)
        assert np.abs(out["c"].reshape(-1) - (1 - o["alpha"])).max() < 1e-6 * max(1.0, np.abs(o["alpha"]).max())
        err = np.abs(out["Jc"].reshape(-1, 6) + o["grad"][:, :6]).max(axis=1) / np.maximum(1.0, np.abs(o["grad"]).max(axis=1))
        assert np.quantile(err, 0.99) < 1e-5
        # another state layout: [p; r] in a 7-column row (the columns are arguments, not a convention)
        Y = np.concatenate([X[:, 6:9], X[:, 0:3], np.ones((n, 1))], axis=1)
        alt = ctx.se3_constraints(Y, sv, so, r_obs, p_obs, r_col=3, p_col=0)
        assert np.array_equal(alt["c"], out["c"]) and np.array_equal(alt["Jc"], out["Jc"])
        m = 150; eps = 1e-6
        Jt = ctx.se3_constraints(X[:m], sv, so, r_obs, p_obs, pdip_tol=1e-10)["Jc"]
        fd = np.zeros((m, K, 6))
        for a, col in enumerate((0, 1, 2, 6, 7, 8)):
            d = np.zeros(12); d[col] = eps
            fd[:, :, a] = (ctx.se3_constraints(X[:m] + d, sv, so, r_obs, p_obs, pdip_tol=1e-10)["c"]
                           - ctx.se3_constraints(X[:m] - d, sv, so, r_obs, p_obs, pdip_tol=1e-10)["c"]) / (2 * eps)
        ef = np.abs(fd - Jt).max(axis=2) / np.maximum(1.0, np.abs(Jt).max(axis=2))
        assert np.median(ef) < 1e-5 and (ef < 1e-3).mean() > 0.95
    assert ctx.se3_constraints(np.zeros((0, 12)), sv, so, r_obs, p_obs)["c"].shape == (0, K)
    with pytest.raises(RuntimeError, match="disjoint"):
        ctx.se3_constraints(X[:4], sv, so, r_obs, p_obs, r_col=0, p_col=2)
    with pytest.raises(RuntimeError, match="disjoint"):
        ctx.se3_constraints(X[:4], sv, so, r_obs, p_obs, r_col=10, p_col=0)
    with pytest.raises(RuntimeError, match="shape index out of range"):
        ctx.se3_constraints(X[:4], 999, so, r_obs, p_obs)
    assert np.array_equal(ctx.se3_constraints(X[:4], sv, so, r_obs, p_obs)["c"], out["c"][:4])


@pytest.mark.parametrize("att_kind", [0, 1])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_scene_cull_keeps_near_pairs_bit_identical(ctx, att_kind, kind):
    """dcol_scene_set_cull (SURVEY.md §8f-1 broad phase): pairs that survive the bounding-ball test are bit-identical to
    the call without the cull; culled pairs truly have alpha > alpha_cut, carry a valid lower bound of it, status 3,
    zero derivative rows."""
    nb = 300
    t, shape, r, att = scene_bodies(nb, att_kind, spread=4.0)
    ctx.set_shape_arrays(*t.arrays())
    full = ctx.scene_allpairs(kind, shape, r, att, att_kind)
    for cut in (1.0, 2.5):
        try:
            ctx.scene_set_cull(cut)
            c = ctx.scene_allpairs(kind, shape, r, att, att_kind)
        finally:
            ctx.scene_set_cull(0.0)
        culled = c["status"] == 3
        assert culled.any() and (~culled).any()
        for k in c:
            assert np.array_equal(c[k][~culled], full[k][~culled], equal_nan=True), (k, cut)
        ok = culled & (full["status"] == 0)
        assert (full["alpha"][ok] > cut).all()                       # nothing within the interaction radius is lost
        assert (c["alpha"][culled] > cut).all() and (c["alpha"][ok] <= full["alpha"][ok] * (1 + 1e-6)).all()   # a lower bound (exact for sphere pairs, hence the solver-accuracy slack)
        assert (c["iters"][culled] == 0).all()
        if kind == 1:
            assert (c["grad"][culled] == 0).all()
        if kind == 2:
            assert (c["J"][culled] == 0).all() and np.isnan(c["x"][culled]).all()
        if cut == 1.0:
            assert ((full["alpha"] <= 1) & (full["status"] == 0) <= ~culled).all()      # every colliding pair is solved


def test_scene_cull_with_offsets_and_wrench(ctx):
    """the cull respects r_offset (ball centre r + Qn r_offset) and dcol_scene_wrench skips culled pairs consistently"""
    nb = 200
    A5, b5 = dc.create_n_sided(5, 0.6)
    protos = [dc.Capsule(0.3, 1.2), dc.Sphere(0.8), dc.Cone(2.0, np.deg2rad(22)), dc.create_rect_prism(1.0, 1.0, 1.0, attitude="quat")[0],
              dc.Polygon(A5, b5, 0.2), dc.Ellipsoid(util.ELL_P), dc.Cylinder(0.6, 3.0)]
    g = util.rng(4)
    for p in protos:
        p.r_offset = 1.5 * g.standard_normal(3)
        p.Q_offset = dc.dcm_from_q(g.standard_normal(4))
    t = dc.ShapeTable(); ids = np.array([t.add(p) for p in protos], np.int32)
    shape = ids[np.arange(nb) % len(ids)]
    r = 4.0 * g.standard_normal((nb, 3)); q = g.standard_normal((nb, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    ctx.set_shape_arrays(*t.arrays())
    full = ctx.scene_allpairs(1, shape, r, q, 0)
    lam = g.random(full["alpha"].shape[0])
    lam_near = np.where(full["alpha"] <= 2.0, lam, 0.0)             # the caller's multipliers vanish outside the interaction radius
    w_full = ctx.scene_wrench(shape, r, q, lam_near, 0)
    try:
        ctx.scene_set_cull(2.0)
        c = ctx.scene_allpairs(1, shape, r, q, 0)
        w_cull = ctx.scene_wrench(shape, r, q, lam_near, 0)
    finally:
        ctx.scene_set_cull(0.0)
    culled = c["status"] == 3
    assert culled.mean() > 0.3 and (full["alpha"][culled & (full["status"] == 0)] > 2.0).all()
    assert np.array_equal(c["grad"][~culled], full["grad"][~culled], equal_nan=True)
    np.testing.assert_allclose(w_cull["wrench"], w_full["wrench"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(w_cull["status"] == 3, culled)


def test_scene_cull_config4(ctx):
    """BASELINE config 4 with the contact example's interaction radius (alpha_cut = 1: only colliding pairs matter to the
    complementarity rows, ncp_contact_simulation.jl:143-149): same colliding set as the full solve."""
    nb = 4096
    t, shape, r, att = scene_bodies(nb, 0, spread=5.0)
    ctx.set_shape_arrays(*t.arrays())
    full = ctx.scene_allpairs(0, shape, r, att, 0)
    try:
        ctx.scene_set_cull(1.0)
        c = ctx.scene_allpairs(0, shape, r, att, 0)
    finally:
        ctx.scene_set_cull(0.0)
    culled = c["status"] == 3
    hit = (full["alpha"] <= 1) & (full["status"] == 0)
    assert not (hit & culled).any()
    assert np.array_equal(c["alpha"][~culled], full["alpha"][~culled]) and np.array_equal(c["x"][~culled], full["x"][~culled])
    assert culled.mean() > 0.5


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_warm_start_is_opt_in_and_consistent(ctx, kind):
    """dcol_set_warm_start (not in the reference, excluded from parity): the default path is untouched
    (bit-identical with the warm buffer attached but empty), a warm solve of slightly moved poses needs fewer iterations
    and agrees with the cold solve of the same poses to the solver tolerance."""
    import torch
    n = 40000
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, 1, seed=23, variant="cube")
    ctx.set_shape_arrays(*t.arrays())
    cold = ctx._batch(kind, s1, s2, r1, a1, r2, a2, 1, 1e-6)
    state = torch.zeros(n * ctx.warm_stride(), dtype=torch.float64, device="cuda:0")
    g = util.rng(77)
    r1b = r1 + 0.01 * g.standard_normal(r1.shape); r2b = r2 + 0.01 * g.standard_normal(r2.shape)
    a1b = a1 + 0.005 * g.standard_normal(a1.shape); a2b = a2 + 0.005 * g.standard_normal(a2.shape)
    cold_b = ctx._batch(kind, s1, s2, r1b, a1b, r2b, a2b, 1, 1e-6)
    try:
        ctx.set_warm_start(state, n, 1e-2)
        first = ctx._batch(kind, s1, s2, r1, a1, r2, a2, 1, 1e-6)          # empty records: cold starts, iterates saved
        for k in cold:
            assert np.array_equal(first[k], cold[k], equal_nan=True), k
        warm_b = ctx._batch(kind, s1, s2, r1b, a1b, r2b, a2b, 1, 1e-6)       # starts from the saved iterates
    finally:
        ctx.set_warm_start(None)
    assert (warm_b["status"] == 0).all()
    assert warm_b["iters"].mean() < 0.7 * cold_b["iters"].mean(), (warm_b["iters"].mean(), cold_b["iters"].mean())
    ea = util.rel_err(warm_b["alpha"], cold_b["alpha"])          # both are O(pdip_tol)-accurate solutions of the same problem
    assert np.quantile(ea, 0.999) < 1e-5 and ea.max() < 1e-3, (np.quantile(ea, 0.999), ea.max())
    assert ((warm_b["alpha"] <= 1) == (cold_b["alpha"] <= 1)).mean() > 0.9999
    if kind == 1:
        # the duals z (and with them grad = -z' du/dtheta) are only O(sqrt(pdip_tol))-accurate at mu < 1e-6 in either solve:
        # the reference's own tests compare the gradient with J[4,:] at 1e-3 and with finite differences at 1e-1
        assert np.quantile(util.rel_err(warm_b["grad"], cold_b["grad"]), 0.99) < 1e-2
    again = ctx._batch(kind, s1, s2, r1b, a1b, r2b, a2b, 1, 1e-6)            # warm start off again: the cold result, bit for bit
    for k in cold_b:
        assert np.array_equal(again[k], cold_b[k], equal_nan=True), k
