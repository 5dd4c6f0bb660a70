"""Development aid: wall-clock latency of small host-buffer calls (single pair, 100, 1k, 10k pairs)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import dcol_b200 as dc
import util

cone = dc.Cone(2.0, np.deg2rad(22)); caps = dc.Capsule(0.3, 1.2)
cone.r = np.array([0.3, 0.2, 1.0]); caps.r = np.array([-1.0, 0.4, 0.1])
for _ in range(20):
    dc.proximity(cone, caps)
t0 = time.perf_counter()
for _ in range(200):
    a, x = dc.proximity(cone, caps)
print(f"dc.proximity(P1,P2) single pair: {(time.perf_counter()-t0)/200*1e6:.1f} us per call  alpha={a:.6f}")
t0 = time.perf_counter()
for _ in range(200):
    a, x, J = dc.proximity_jacobian(cone, caps)
print(f"dc.proximity_jacobian(P1,P2) single pair: {(time.perf_counter()-t0)/200*1e6:.1f} us per call")
ctx = dc.Context(0)
for kind, name in ((0, "proximity"), (2, "jacobian")):
    for n in (1, 100, 1000, 10000, 100000):
        table, s1, s2, r1, a1, r2, a2 = util.cube_batch(n, 0, 5)
        ctx.set_shape_arrays(*table.arrays())
        f = ctx.proximity_batch if kind == 0 else ctx.proximity_jacobian_batch
        for _ in range(5):
            f(s1, s2, r1, a1, r2, a2, att_kind=0, pdip_tol=1e-6)
        t0 = time.perf_counter()
        reps = 50 if n < 100000 else 10
        for _ in range(reps):
            f(s1, s2, r1, a1, r2, a2, att_kind=0, pdip_tol=1e-6)
        dt = (time.perf_counter() - t0) / reps
        print(f"{name:10s} cube-cube n={n:6d}: {dt*1e6:9.1f} us per call  kernels {ctx.last_kernel_ms()*1e3:8.1f} us  {n/dt/1e6:8.3f} Mpairs/s")

# the same 100k-pair Jacobian call with page-locked (registered) numpy buffers
n = 100000
table, s1, s2, r1, a1, r2, a2 = util.cube_batch(n, 0, 5)
ctx.set_shape_arrays(*table.arrays())
ins = [ctx.pin(np.ascontiguousarray(a)) for a in (s1, s2, r1, a1, r2, a2)]
out = ctx.alloc_outputs(2, n, att_kind=0, pinned=True)
for _ in range(3):
    ctx.proximity_jacobian_batch(*ins, att_kind=0, out=out)
t0 = time.perf_counter()
for _ in range(10):
    ctx.proximity_jacobian_batch(*ins, att_kind=0, out=out)
dt = (time.perf_counter() - t0) / 10
print(f"jacobian   cube-cube n={n:6d} registered buffers: {dt*1e6:9.1f} us per call  {n/dt/1e6:8.3f} Mpairs/s")
