"""Latency of the contact example's own scene (13 bodies = 78 pairs, ncp_contact_simulation.jl:264-269) through the scene operators
and through the plain batched call on host arrays (wall clock, Python wrapper)."""
import sys, os, time, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import dcol_b200 as dc
import util
from test_gpu_operators import scene_bodies
ctx = dc.Context(0)
for nb in (13, 16, 32):
    t, shape, r, att = scene_bodies(nb, 0, spread=2.0)
    ctx.set_shape_arrays(*t.arrays())
    ij = np.array(list(itertools.combinations(range(nb), 2)))
    n = len(ij); lam = np.ones(n)
    def wall(f, reps=300):
        for _ in range(20): f()
        t0 = time.perf_counter()
        for _ in range(reps): f()
        return 1e6 * (time.perf_counter() - t0) / reps
    a = wall(lambda: ctx.scene_wrench(shape, r, att, lam, 0))
    b = wall(lambda: ctx.scene_allpairs(1, shape, r, att, 0))
    c = wall(lambda: ctx.scene_allpairs(0, shape, r, att, 0))
    s1, s2 = shape[ij[:, 0]], shape[ij[:, 1]]
    r1, a1, r2, a2 = r[ij[:, 0]], att[ij[:, 0]], r[ij[:, 1]], att[ij[:, 1]]
    d = wall(lambda: ctx._batch(1, s1, s2, r1, a1, r2, a2, 0, 1e-6))
    e = wall(lambda: ctx._batch(0, s1, s2, r1, a1, r2, a2, 0, 1e-6))
    print(f"{nb} bodies, {n} pairs: scene_wrench {a:7.1f} us | scene_allpairs gradient {b:7.1f} us, proximity {c:7.1f} us | plain batched call on the same pairs: gradient {d:7.1f} us, proximity {e:7.1f} us", flush=True)
