"""Warm start on BASELINE config 5 (piano-mover batch: 1024 rollouts x 100 knots x 3 obstacles, box-box MRP): iteration
counts and call time of dcol_proximity_gradient for a cold solve, a first solve that saves its iterates, and a warm solve of
slightly moved states (what successive AL-iLQR iterations / line-search trials evaluate).  Writes to stdout."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dcol_b200 as dc
import util
ctx = dc.Context(0)
B, N, K = 1024, 100, 3
veh = dc.create_rect_prism(2.5, 0.15, 0.01)[0]
obs = [dc.create_rect_prism(3.0, 3.0, 0.02)[0], dc.create_rect_prism(4.0, 1.0, 0.02)[0], dc.create_rect_prism(1.0, 3.0, 0.02)[0]]
t = dc.ShapeTable(); sv = t.add(veh); so = [t.add(x) for x in obs]
ctx.set_shape_arrays(*t.arrays())
g = util.rng(17)
theta = g.uniform(-np.pi, np.pi, B * N)
xy = np.stack([g.uniform(-4, 4, B * N), g.uniform(-4, 4, B * N)], axis=1)
r2 = np.array([[-1.5, 1.5, 0.0], [2.0, -2.5, 0.0], [3.0, 2.5, 0.0]]); p2 = np.zeros((3, 3))
n = B * N * K
def pairs(xy, theta):
    r1 = np.repeat(np.concatenate([xy, np.zeros((B * N, 1))], axis=1), K, axis=0)
    p1 = np.repeat(np.stack([np.zeros(B * N), np.zeros(B * N), np.tan(theta / 4)], axis=1), K, axis=0)
    return r1, p1, np.tile(r2, (B * N, 1)), np.tile(p2, (B * N, 1))
s1 = np.full(n, sv, np.int32); s2 = np.tile(np.array(so, np.int32), B * N)
dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
alpha = torch.empty(n, dtype=torch.float64, device=dev); grad = torch.empty((n, 12), dtype=torch.float64, device=dev)
st = torch.empty(n, dtype=torch.int32, device=dev); it = torch.empty(n, dtype=torch.int32, device=dev)
ds1, ds2 = T(s1), T(s2)
def run(xy, theta, reps=5):
    d = [T(a) for a in pairs(xy, theta)]
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.query_raw(1, n, ds1, ds2, *d, 1, 1e-6, dc.MEM_DEVICE, alpha, None, grad, st, it)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best, ctx.last_kernel_ms(), it.cpu().numpy().copy(), alpha.cpu().numpy().copy(), st.cpu().numpy().copy()
w, k, its, a0, s0 = run(xy, theta)
print(f"cold                      : call {w*1e3:7.3f} ms  solver {k:6.3f} ms  mean iters {its.mean():.2f}  failed {(s0 != 0).sum()}")
state = torch.zeros(n * ctx.warm_stride(), dtype=torch.float64, device=dev)
for step, (dxy, dth) in enumerate(((0.0, 0.0), (0.005, 0.002), (0.02, 0.01), (0.1, 0.05))):
    state.zero_()
    ctx.set_warm_start(state, n, 1e-2)
    run(xy, theta, reps=1)                                                  # saves the iterates of the nominal states
    xy2 = xy + dxy * g.standard_normal(xy.shape); th2 = theta + dth * g.standard_normal(theta.shape)
    d = [T(a) for a in pairs(xy2, th2)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.query_raw(1, n, ds1, ds2, *d, 1, 1e-6, dc.MEM_DEVICE, alpha, None, grad, st, it)
    torch.cuda.synchronize(); ww = time.perf_counter() - t0
    kw = ctx.last_kernel_ms(); iw = it.cpu().numpy().copy(); aw = alpha.cpu().numpy().copy(); sw = st.cpu().numpy().copy()
    ctx.set_warm_start(None)
    wc, kc, ic, ac, sc = run(xy2, th2)
    print(f"states moved by sigma=({dxy} m, {dth} rad): warm call {ww*1e3:7.3f} ms solver {kw:6.3f} ms mean iters {iw.mean():.2f} failed {(sw != 0).sum()}"
          f" | cold call {wc*1e3:7.3f} ms solver {kc:6.3f} ms mean iters {ic.mean():.2f} | max rel diff of alpha {np.abs(aw - ac).max() / np.abs(ac).max():.2e}")
