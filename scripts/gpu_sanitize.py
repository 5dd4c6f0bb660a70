"""Development aid: small mixed batches of all three queries, both attitude kinds, offsets and big polytopes — for compute-sanitizer."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import dcol_b200 as dc
import util
ctx = dc.Context(0)
for att in (0, 1):
    for variant in ("cube", "scene"):
        t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(3000, att, seed=11, variant=variant)
        ctx.set_shape_arrays(*t.arrays())
        for kind in (0, 1, 2):
            o = ctx._batch(kind, s1, s2, r1, a1, r2, a2, att, 1e-6)
            assert (o["status"] == 0).all()
A1, b1, A2, b2 = util.load_polytopes()
g = util.rng(8)
dirs = g.standard_normal((30, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
shapes = [dc.Polytope(A1, b1), dc.Polytope(dirs, 0.7 + 0.3 * g.random(30)), dc.Sphere(0.5)]
t = dc.ShapeTable(); ids = np.array([t.add(s) for s in shapes], np.int32)
n = 2000
s1 = ids[g.integers(0, 3, n)]; s2 = ids[g.integers(0, 3, n)]
r1, a1, r2, a2 = util.random_poses(n, 0, 5)
ctx.set_shape_arrays(*t.arrays())
for kind in (0, 2):
    o = ctx._batch(kind, s1, s2, r1, a1, r2, a2, 0, 1e-6)
nb = 40
sh = ids[np.arange(nb) % 3]
o = ctx.scene_allpairs(1, sh, 2 * g.standard_normal((nb, 3)), a1[:nb], 0)
print("sanitize workload done", o["alpha"].shape)
