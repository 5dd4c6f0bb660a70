"""GPU (-m gpu): the staging pipeline of large host-pointer calls on PAGEABLE caller memory (plain numpy / Julia arrays):
copy threads + pinned staging inside the library (dcol_capi.cu, HostCopier).  Same kernels, same pairs: results must be
bit-identical to the driver's pageable path (DCOL_HOST_PIPE=0), to page-locked caller buffers and to device-resident
calls; small calls and page-locked buffers must not take the pipeline."""
import os

import numpy as np
import pytest

import dcol_b200 as dc
import util
from test_gpu_operators import scene_bodies

pytestmark = pytest.mark.gpu


def _ctx(pipe, threads=None):
    """The switch is read once per context, at its first host-pointer query."""
    old = {k: os.environ.get(k) for k in ("DCOL_HOST_PIPE", "DCOL_HOST_THREADS")}
    os.environ["DCOL_HOST_PIPE"] = "1" if pipe else "0"
    if threads is not None:
        os.environ["DCOL_HOST_THREADS"] = str(threads)
    c = dc.Context(0)
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(60000, 0, seed=1)       # > 8 MB: triggers the probe of the environment
    c.set_shape_arrays(*t.arrays())
    c._batch(0, s1, s2, r1, a1, r2, a2, 0, 1e-6)
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    return c


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("kind,att_kind,n", [(0, 0, 450000), (1, 1, 230000), (2, 0, 170000), (2, 1, 70000)])
def test_pageable_pipeline_is_bit_identical(kind, att_kind, n):
    on, off = _ctx(True, threads=5), _ctx(False)
    try:
        t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att_kind, seed=11 + kind, variant="scene")
        for c in (on, off):
            c.set_shape_arrays(*t.arrays())
        a = on._batch(kind, s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
        b = off._batch(kind, s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
        _same(a, b)
        assert (a["status"] == 0).all()
        # page-locked result buffers keep the direct path; the values do not depend on the path
        out = on.alloc_outputs(kind, n, att_kind, pinned=True)
        try:
            c2 = on._batch(kind, s1, s2, r1, a1, r2, a2, att_kind, 1e-6, out=out)
            _same(a, c2)
        finally:
            for arr in out.values():
                on.unpin(arr)
        # optional per-pair final mu travels through the staging area as well
        m1 = on._batch(kind, s1, s2, r1, a1, r2, a2, att_kind, 1e-6, want_mu=True)
        m0 = off._batch(kind, s1, s2, r1, a1, r2, a2, att_kind, 1e-6, want_mu=True)
        assert np.array_equal(m1["mu"], m0["mu"]) and (m1["mu"] < 1e-6).all()
    finally:
        on.close(); off.close()


def test_pageable_pipeline_indexed_and_scene():
    on, off = _ctx(True), _ctx(False)
    try:
        # pose-table query: the tables go up once, the pair lists chunk by chunk
        g = util.rng(5)
        t, s1, s2, _, _, _, _ = util.mixed_batch(400000, 1, seed=3)
        n = s1.shape[0]
        np1, np2 = 5000, 700
        r1, a1, _, _ = util.random_poses(np1, 1, seed=21)
        _, _, r2, a2 = util.random_poses(np2, 1, seed=22)
        pose1 = g.integers(0, np1, n).astype(np.int32); pose2 = g.integers(0, np2, n).astype(np.int32)
        for c in (on, off):
            c.set_shape_arrays(*t.arrays())
        a = on.query_indexed(1, s1, s2, pose1, pose2, r1, a1, r2, a2, dc.ATT_MRP)
        b = off.query_indexed(1, s1, s2, pose1, pose2, r1, a1, r2, a2, dc.ATT_MRP)
        _same(a, b)
        e = off._batch(1, s1, s2, r1[pose1], a1[pose1], r2[pose2], a2[pose2], dc.ATT_MRP, 1e-6)
        _same(a, e)
        # all-pairs scene: pair list generated on the device, results to pageable memory
        t, shape, r, att = scene_bodies(1100, 0, spread=3.0)
        for c in (on, off):
            c.set_shape_arrays(*t.arrays())
        _same(on.scene_allpairs(1, shape, r, att, 0), off.scene_allpairs(1, shape, r, att, 0))
    finally:
        on.close(); off.close()


def test_pageable_pipeline_reports_bad_shape_index():
    on = _ctx(True)
    try:
        t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(300000, 0, seed=4)
        on.set_shape_arrays(*t.arrays())
        s1 = s1.copy(); s1[250001] = 99
        with pytest.raises(Exception, match="shape index out of range"):
            on._batch(0, s1, s2, r1, a1, r2, a2, 0, 1e-6)
        s1[250001] = 0
        assert (on._batch(0, s1, s2, r1, a1, r2, a2, 0, 1e-6)["status"] == 0).all()      # the context stays usable
    finally:
        on.close()


def test_operator_results_through_the_pipeline():
    """Operators that solve into device scratch (culled scene, scene wrench, planar constraint rows) return their results
    to pageable memory in slabs through the same staging area (results_to_host): bit-identical to the driver's copies;
    a pinned-memory budget too small for the call falls back to the driver path."""
    on, off = _ctx(True), _ctx(False)
    try:
        t, shape, r, att = scene_bodies(1300, 0, spread=3.0)             # 844 350 pairs: 100 MB of gradient rows
        for c in (on, off):
            c.set_shape_arrays(*t.arrays())
            c.scene_set_cull(1.5)
        a, b = on.scene_allpairs(1, shape, r, att, 0), off.scene_allpairs(1, shape, r, att, 0)
        _same(a, b)
        assert (a["status"] == 3).sum() > 0 and (a["status"] == 0).sum() > 0
        lam = util.rng(3).uniform(0, 1, a["alpha"].shape[0])
        _same(on.scene_wrench(shape, r, att, lam, 0, per_pair=True), off.scene_wrench(shape, r, att, lam, 0, per_pair=True))
        for c in (on, off):
            c.scene_set_cull(0.0)
        # planar constraint rows: 250 000 states x 3 obstacles (6 MB + 18 MB + 3 MB)
        veh = dc.create_rect_prism(2.5, 0.15, 0.01)[0]
        obs = [dc.create_rect_prism(3.0, 3.0, 0.02)[0], dc.create_rect_prism(4.0, 1.0, 0.02)[0], dc.create_rect_prism(1.0, 3.0, 0.02)[0]]
        tb = dc.ShapeTable(); sv = tb.add(veh); so = np.array([tb.add(x) for x in obs], np.int32)
        g = util.rng(9)
        xyt = np.stack([g.uniform(-4, 4, 250000), g.uniform(-4, 4, 250000), g.uniform(-3, 3, 250000)], axis=1)
        r2 = np.array([[-1.5, 1.5, 0.0], [2.0, -2.5, 0.0], [3.0, 2.5, 0.0]]); p2 = np.zeros((3, 3))
        for c in (on, off):
            c.set_shape_arrays(*tb.arrays())
        _same(on.se2_constraints(xyt, sv, so, r2, p2), off.se2_constraints(xyt, sv, so, r2, p2))
        # no pinned memory to spare (DCOL_PIN_MAX_MB): the calls still work, through the driver's pageable copies
        os.environ["DCOL_PIN_MAX_MB"] = "1"
        try:
            _same(on.se2_constraints(xyt, sv, so, r2, p2), off.se2_constraints(xyt, sv, so, r2, p2))
            tm, s1, s2, r1, a1, rr2, a2 = util.mixed_batch(120000, 0, seed=2)
            for c in (on, off):
                c.set_shape_arrays(*tm.arrays())
            _same(on._batch(0, s1, s2, r1, a1, rr2, a2, 0, 1e-6), off._batch(0, s1, s2, r1, a1, rr2, a2, 0, 1e-6))
        finally:
            os.environ.pop("DCOL_PIN_MAX_MB", None)
    finally:
        on.close(); off.close()
