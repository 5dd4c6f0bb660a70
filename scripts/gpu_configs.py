"""Throughput of the five BASELINE.json configs through the public Python API (host buffers, one blocking call at a time)
and, where it exists, the kernel time of the same call.  Informational: bench.py reports configs 2 and 3."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import dcol_b200 as dc
import util
from test_gpu_operators import scene_bodies

ctx = dc.Context(0)

def timed(fn, reps=5):
    fn(); best = 1e9; kms = None
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0); kms = ctx.last_kernel_ms()
    return best, kms, out

# 1: cube-cube, 10k random poses, proximity
t, s1, s2, r1, a1, r2, a2 = util.cube_batch(10000, 0)
ctx.set_shape_arrays(*t.arrays())
w, k, o = timed(lambda: ctx._batch(0, s1, s2, r1, a1, r2, a2, 0, 1e-6))
print(f"config 1  cube-cube proximity, 10k pairs:            call {w*1e3:8.3f} ms  kernels {k:7.3f} ms  {10000/w/1e6:8.2f} Mpairs/s  (mean iters {o['iters'].mean():.2f})")
# 2: 7x7 mixed, 1M, proximity
t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(1000000, 0)
ctx.set_shape_arrays(*t.arrays())
w, k, o = timed(lambda: ctx._batch(0, s1, s2, r1, a1, r2, a2, 0, 1e-6), 3)
print(f"config 2  7x7 mixed proximity, 1M pairs:             call {w*1e3:8.3f} ms  kernels {k:7.3f} ms  {1e6/w/1e6:8.2f} Mpairs/s")
# 3: mixed MRP jacobian, 2M per GPU
t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(2000000, 1, seed=4321)
ctx.set_shape_arrays(*t.arrays())
w, k, o = timed(lambda: ctx._batch(2, s1, s2, r1, a1, r2, a2, 1, 1e-6), 2)
print(f"config 3  mixed MRP jacobian, 2M pairs (1 of 8 GPUs): call {w*1e3:8.3f} ms  kernels {k:7.3f} ms  {2e6/w/1e6:8.2f} Mpairs/s")
# 4: 4096-body scene, all pairs, gradient
nb = 4096
t, shape, r, att = scene_bodies(nb, 0, spread=5.0)
ctx.set_shape_arrays(*t.arrays())
w, k, o = timed(lambda: ctx.scene_allpairs(1, shape, r, att, 0), 2)
npairs = nb * (nb - 1) // 2
print(f"config 4  4096-body scene, {npairs} gradient pairs:  call {w*1e3:8.3f} ms  kernels {k:7.3f} ms  {npairs/w/1e6:8.2f} Mpairs/s  (collisions {int((o['alpha'] <= 1).sum())})")
lam = util.rng(5).uniform(0.0, 1.0, npairs)
w2, k2, o2 = timed(lambda: ctx.scene_wrench(shape, r, att, lam, 0, per_pair=False), 2)
print(f"config 4f same scene through dcol_scene_wrench (per-body sums only): call {w2*1e3:8.3f} ms  {npairs/w2/1e6:8.2f} Mpairs/s")
w3, k3, o3 = timed(lambda: ctx.scene_wrench(shape, r, att, lam, 0, per_pair=True), 2)
print(f"config 4f same, alpha/status/iters per pair returned as well:        call {w3*1e3:8.3f} ms  {npairs/w3/1e6:8.2f} Mpairs/s")
# 4c: the same scene with the opt-in broad phase at the example's interaction radius (only colliding pairs matter to the
# complementarity rows) and at twice that
for cut in (1.0, 2.0):
    ctx.scene_set_cull(cut)
    wc, kc, oc = timed(lambda: ctx.scene_allpairs(1, shape, r, att, 0), 2)
    wc2, kc2, oc2 = timed(lambda: ctx.scene_wrench(shape, r, att, lam, 0, per_pair=False), 2)
    ctx.scene_set_cull(0.0)
    print(f"config 4c same scene, dcol_scene_set_cull({cut}): {int((oc['status'] != 3).sum())} of {npairs} pairs solved: allpairs call {wc*1e3:8.3f} ms  kernels {kc:7.3f} ms"
          f" | scene_wrench (per-body sums only) call {wc2*1e3:8.3f} ms")
# 5: piano mover batch: 1024 rollouts x 100 knots x 3 obstacles, jacobian, MRP p = (0,0,tan(theta/4))
B, N, K = 1024, 100, 3
veh = dc.create_rect_prism(2.5, 0.15, 0.01)[0]
obs = [dc.create_rect_prism(3.0, 3.0, 0.02)[0], dc.create_rect_prism(4.0, 1.0, 0.02)[0], dc.create_rect_prism(1.0, 3.0, 0.02)[0]]
t = dc.ShapeTable(); sv = t.add(veh); so = [t.add(x) for x in obs]
ctx.set_shape_arrays(*t.arrays())
g = util.rng(17)
theta = g.uniform(-np.pi, np.pi, B * N)
r1 = np.stack([g.uniform(-4, 4, B * N), g.uniform(-4, 4, B * N), np.zeros(B * N)], axis=1)
p1 = np.stack([np.zeros(B * N), np.zeros(B * N), np.tan(theta / 4)], axis=1)
r2 = np.array([[-1.5, 1.5, 0.0], [2.0, -2.5, 0.0], [3.0, 2.5, 0.0]]); p2 = np.zeros((3, 3))
pose1 = np.repeat(np.arange(B * N, dtype=np.int32), K); pose2 = np.tile(np.arange(K, dtype=np.int32), B * N)
s1 = np.full(B * N * K, sv, np.int32); s2 = np.array(so, np.int32)[pose2]
w, k, o = timed(lambda: ctx.query_indexed(2, s1, s2, pose1, pose2, r1, p1, r2, p2, dc.ATT_MRP), 3)
print(f"config 5  piano-mover batch, {B*N*K} box-box MRP jacobians: call {w*1e3:8.3f} ms  kernels {k:7.3f} ms  {B*N*K/w/1e6:8.2f} Mpairs/s  (failed {int((o['status'] != 0).sum())})")
xyt = np.stack([r1[:, 0], r1[:, 1], theta], axis=1)
w5, k5, o5 = timed(lambda: ctx.se2_constraints(xyt, sv, np.array(so, np.int32), r2, p2), 3)
print(f"config 5f same batch through dcol_se2_constraints (c and dc/d(x,y,theta) per pair): call {w5*1e3:8.3f} ms  kernels {k5:7.3f} ms  {B*N*K/w5/1e6:8.2f} Mpairs/s")
# 5g: the quadrotor of cluttered_hallway_quadrotor.jl (a sphere, state [r; v; p; omega]) against its 11 obstacles (:139-153):
# 1024 rollouts x 100 knots x 11 obstacles through dcol_se3_constraints, and the same pairs as plain gradient rows
A1, b1, A2, b2 = util.load_polytopes()
veh = dc.SphereMRP(0.25)
bot = dc.create_rect_prism(20, 5.0, 0.2)[0]; top = dc.create_rect_prism(20, 5.0, 0.2)[0]
obs = [dc.CylinderMRP(0.6, 3.0), dc.CapsuleMRP(0.2, 5.0), dc.SphereMRP(0.8), dc.ConeMRP(2.0, np.deg2rad(22)), dc.PolytopeMRP(A2, b2),
       dc.PolygonMRP(*dc.create_n_sided(5, 0.6), 0.2), dc.CylinderMRP(1.1, 2.3), dc.CapsuleMRP(0.8, 1.0), dc.SphereMRP(0.5), bot, top]
K = len(obs)
t = dc.ShapeTable(); sv = t.add(veh); so = np.array([t.add(x) for x in obs], np.int32)
ctx.set_shape_arrays(*t.arrays())
g = util.rng(19)
r_obs = np.zeros((K, 3)); r_obs[:9, 0] = np.linspace(-5, 5, 9); r_obs[:9, 1] = g.standard_normal(9); r_obs[:9, 2] = 3 + g.standard_normal(9)
r_obs[9] = [0, 0, 0.9]; r_obs[10] = [0, 0, 6.0]
q = g.standard_normal((K, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True); q = np.where(q[:, :1] < 0, -q, q)
p_obs = q[:, 1:] / (1 + q[:, :1]); p_obs[9:] = 0.0
X = np.zeros((B * N, 12))
X[:, 0] = g.uniform(-8, 8, B * N); X[:, 1] = g.uniform(-2, 2, B * N); X[:, 2] = g.uniform(1.5, 5.5, B * N); X[:, 6:9] = 0.1 * g.standard_normal((B * N, 3))
w6, k6, o6 = timed(lambda: ctx.se3_constraints(X, sv, so, r_obs, p_obs), 3)
print(f"config 5g quadrotor hallway batch, {B*N*K} pairs through dcol_se3_constraints (c and the two 1x3 Jacobian blocks per pair): call {w6*1e3:8.3f} ms  kernels {k6:7.3f} ms  {B*N*K/w6/1e6:8.2f} Mpairs/s  (failed {int((o6['status'] != 0).sum())})")
i1 = np.repeat(np.arange(B * N, dtype=np.int32), K); i2 = np.tile(np.arange(K, dtype=np.int32), B * N)
w7, k7, o7 = timed(lambda: ctx.query_indexed(1, np.full(B * N * K, sv, np.int32), so[i2], i1, i2, np.ascontiguousarray(X[:, 0:3]), np.ascontiguousarray(X[:, 6:9]), r_obs, p_obs, dc.ATT_MRP), 3)
print(f"config 5h same pairs through dcol_query_indexed (gradient rows, 13 numbers per pair): call {w7*1e3:8.3f} ms  kernels {k7:7.3f} ms  identical c: {np.array_equal(1 - o7['alpha'], o6['c'].reshape(-1))}")
