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
