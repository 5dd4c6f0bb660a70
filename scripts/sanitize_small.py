"""A small pass over every kernel family for compute-sanitizer (memcheck / racecheck / initcheck): binned path with
hand-over and resume (20 000 mixed pairs, three query kinds), small-batch path, scene operators with the broad phase,
pair slices, planar constraints, the generic conic solve with its trace."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import dcol_b200 as dc
import util
from test_gpu_operators import scene_bodies

ctx = dc.Context(0)
for att_kind in (0, 1):
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(20000, att_kind, seed=5, variant="scene")
    ctx.set_shape_arrays(*t.arrays())
    for kind in (0, 1, 2):
        o = ctx._batch(kind, s1, s2, r1, a1, r2, a2, att_kind, 1e-6)
        assert (o["status"] == 0).all()
        o = ctx._batch(kind, s1[:300], s2[:300], r1[:300], a1[:300], r2[:300], a2[:300], att_kind, 1e-6)
        assert (o["status"] == 0).all()
t, shape, r, att = scene_bodies(120, 0, spread=3.0)
ctx.set_shape_arrays(*t.arrays())
n = 120 * 119 // 2
lam = np.ones(n)
for cull in (0.0, 1.5):
    ctx.scene_set_cull(cull)
    ctx.scene_allpairs(2, shape, r, att, 0)
    ctx.scene_wrench(shape, r, att, lam, 0)
    ctx.scene_wrench(shape, r, att, lam[100:3000], 0, pair_range=(100, 3000))
ctx.scene_set_cull(0.0)
veh = dc.create_rect_prism(2.5, 0.15, 0.01)[0]; ob = dc.create_rect_prism(3.0, 3.0, 0.02)[0]
tb = dc.ShapeTable(); sv = tb.add(veh); so = np.array([tb.add(ob)], np.int32)
ctx.set_shape_arrays(*tb.arrays())
ctx.se2_constraints(np.array([[0.5, 4.0, 0.3], [4.0, 0.1, -1.0]]), sv, so, np.array([[0.0, 0.0, 0.0]]), np.zeros((1, 3)))
caps = dc.Capsule(0.3, 1.2); cone = dc.Cone(2.0, np.deg2rad(22)); cone.r = np.array([1.0, 0.5, -0.2])
x, s, z, st, it, tr = ctx.solve_socp(*dc.pair_problem(caps, cone), pdip_tol=1e-6, trace=True)
assert st == 0
print("sanitize_small: all calls returned, launches:", ctx.launch_count())
ctx.close()
