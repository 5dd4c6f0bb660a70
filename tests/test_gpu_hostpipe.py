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
