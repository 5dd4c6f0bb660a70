"""GPU (-m gpu): device groups (`dcol_group_*`, SURVEY.md §8e from ONE process): contiguous index slices on several
contexts, one host thread per member.  Members may share a device, so the single-GPU run covers the slicing, the
threading and the error path; with more devices visible the members spread over them."""
import numpy as np
import pytest

import dcol_b200 as dc
import util

pytestmark = pytest.mark.gpu


def _devices(k):
    import torch
    nd = max(1, torch.cuda.device_count())
    return [i % nd for i in range(k)]


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("kind,att_kind,n,members", [(0, 0, 100003, 3), (1, 1, 40001, 2), (2, 0, 30011, 4), (2, 1, 7, 3), (0, 1, 2, 5)])
def test_group_matches_single_context(ctx, kind, att_kind, n, members):
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att_kind, seed=31 + kind, variant="scene")
    ctx.set_shape_arrays(*t.arrays())
    e = ctx._batch(kind, s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
    g = dc.Group(_devices(members))
    try:
        assert len(g) == members
        g.set_shape_arrays(*t.arrays())
        o = g._batch(kind, s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
        _same(o, e)                                   # same kernels on every member: bit-identical, in the caller's order
        assert g.last_kernel_ms() > 0
        o2 = g._batch(kind, s1, s2, r1, a1, r2, a2, att_kind, 1e-6)     # contexts are reused across calls
        _same(o2, e)
    finally:
        g.close()


def test_group_shape_table_and_errors(ctx):
    g = dc.Group(_devices(2))
    try:
        shapes = util.standard_shapes(0, "cube")
        ids = np.array([g.add_shape(s) for s in shapes], np.int32)
        ids_c = np.array([ctx.add_shape(s) for s in shapes], np.int32)
        n = 5000
        r1, a1, r2, a2 = util.random_poses(n, 0, seed=5)
        rg = util.rng(6)
        k1, k2 = rg.integers(0, 7, n), rg.integers(0, 7, n)
        o = g.proximity_gradient_batch(ids[k1], ids[k2], r1, a1, r2, a2)
        e = ctx.proximity_gradient_batch(ids_c[k1], ids_c[k2], r1, a1, r2, a2)
        _same(o, e)
        # an out-of-range shape index in the second member's slice: the call fails and names the member
        bad = ids[k1].copy(); bad[n - 3] = 1000
        with pytest.raises(RuntimeError, match=r"member 1 .*shape index out of range"):
            g.proximity_batch(bad, ids[k2], r1, a1, r2, a2)
        o3 = g.proximity_batch(ids[k1], ids[k2], r1, a1, r2, a2)         # the group stays usable
        _same(o3, ctx.proximity_batch(ids_c[k1], ids_c[k2], r1, a1, r2, a2))
        assert g.proximity_batch(ids[:0], ids[:0], r1[:0], a1[:0], r2[:0], a2[:0])["alpha"].shape == (0,)
        with pytest.raises(RuntimeError, match="query must be"):
            g._batch(7, ids[k1], ids[k2], r1, a1, r2, a2)
    finally:
        g.close()
    with pytest.raises(RuntimeError, match="device"):
        dc.Group([4096])


@pytest.mark.parametrize("att_kind", [0, 1])
def test_scene_slices_and_group_scene(ctx, att_kind):
    """SURVEY.md 8e for the all-pairs scene: every device gets the body arrays and one contiguous slice of the i<j
    enumeration (dcol_scene_allpairs_range / dcol_scene_wrench_range).  Slices are bit-identical to the rows of the whole
    call, with and without the broad phase; the per-body partial sums of a partition add up to the whole; the group
    operators do exactly that over their members."""
    from test_gpu_operators import scene_bodies
    nb = 211
    t, shape, r, att = scene_bodies(nb, att_kind, spread=3.0)
    n = nb * (nb - 1) // 2
    lam = util.rng(5).uniform(0.0, 2.0, n)
    cuts = [0, 1, 5000, 5001, 17000, n]
    g = dc.Group(_devices(3))
    try:
        ctx.set_shape_arrays(*t.arrays()); g.set_shape_arrays(*t.arrays())
        for cull in (0.0, 1.5):
            ctx.scene_set_cull(cull); g.scene_set_cull(cull)
            for kind in (0, 1, 2):
                whole = ctx.scene_allpairs(kind, shape, r, att, att_kind)
                if cull:
                    assert (whole["status"] == 3).any() and (whole["status"] == 0).any()
                for k0, k1 in zip(cuts[:-1], cuts[1:]):
                    part = ctx.scene_allpairs(kind, shape, r, att, att_kind, pair_range=(k0, k1))
                    _same(part, {k: v[k0:k1] for k, v in whole.items()})
                _same(g.scene_allpairs(kind, shape, r, att, att_kind), whole)
            w = ctx.scene_wrench(shape, r, att, lam, att_kind)
            acc = np.zeros((nb, 6))
            for k0, k1 in zip(cuts[:-1], cuts[1:]):
                p = ctx.scene_wrench(shape, r, att, lam[k0:k1], att_kind, pair_range=(k0, k1))
                for k in ("alpha", "status", "iters"):
                    assert np.array_equal(p[k], w[k][k0:k1]), k
                acc += p["wrench"]
            scale = np.abs(w["wrench"]).max()
            assert np.abs(acc - w["wrench"]).max() < 1e-12 * scale          # same terms, another order of summation
            gw = g.scene_wrench(shape, r, att, lam, att_kind)
            for k in ("alpha", "status", "iters"):
                assert np.array_equal(gw[k], w[k]), k
            assert np.abs(gw["wrench"] - w["wrench"]).max() < 1e-12 * scale
            assert np.array_equal(g.scene_wrench(shape, r, att, lam, att_kind, per_pair=False)["wrench"], gw["wrench"])   # reproducible
        empty = ctx.scene_wrench(shape, r, att, np.zeros(0), att_kind, pair_range=(7, 7))
        assert (empty["wrench"] == 0).all() and empty["alpha"].shape == (0,)
        with pytest.raises(ValueError):
            ctx.scene_allpairs(0, shape, r, att, att_kind, pair_range=(5, n + 1))
        # the C entry checks the range itself
        import ctypes as C
        a = np.empty(4); s = np.empty(4, np.int32)
        rc = ctx._lib.dcol_scene_allpairs_range(ctx._h, 0, nb, n - 1, n + 3, shape.ctypes.data, r.ctypes.data, att.ctypes.data, att_kind, 1e-6, 0,
                                                a.ctypes.data, np.empty((4, 3)).ctypes.data, None, s.ctypes.data, None)
        assert rc == -1 and b"k_begin" in ctx._lib.dcol_last_error(ctx._h)
    finally:
        ctx.scene_set_cull(0.0)
        g.close()
