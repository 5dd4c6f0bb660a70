"""Development aid: blocking host-pointer calls from PAGEABLE numpy arrays (the default Python / Julia path), with the
library's staging pipeline on (default) and off (DCOL_HOST_PIPE=0, the driver's pageable copies); results must be identical.
usage: gpu_hostpipe.py   (spawns itself once per mode)"""
import sys, os, time, subprocess, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

def child():
    import numpy as np
    import dcol_b200 as dc
    import util
    ctx = dc.Context(0)
    def timed(fn, reps=4):
        fn(); best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
        return best, ctx.last_kernel_ms(), out
    def digest(o):
        h = hashlib.sha1()
        for k in sorted(o):
            h.update(np.ascontiguousarray(o[k]).tobytes())
        return h.hexdigest()[:12]
    for name, kind, att, n, seed in (("config 2 prox 1M quat", 0, 0, 1000000, 1234), ("config 3 jac 2M mrp", 2, 1, 2000000, 4321),
                                     ("prox 300k quat", 0, 0, 300000, 7), ("grad 600k mrp", 1, 1, 600000, 8)):
        t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att, seed=seed)
        ctx.set_shape_arrays(*t.arrays())
        w, k, o = timed(lambda: ctx._batch(kind, s1, s2, r1, a1, r2, a2, att, 1e-6))
        na = 4 if att == 0 else 3; nt = 6 + 2 * na
        out = {"alpha": np.zeros(n), "status": np.zeros(n, np.int32), "iters": np.zeros(n, np.int32)}
        if kind != 1: out["x"] = np.zeros((n, 3))
        if kind == 1: out["deriv"] = np.zeros((n, nt))
        if kind == 2: out["deriv"] = np.zeros((n, nt, 4))
        w2, k2, o2 = timed(lambda: ctx._batch(kind, s1, s2, r1, a1, r2, a2, att, 1e-6, out=out))
        print(f"{name:24s} fresh outputs {w*1e3:8.2f} ms | reused outputs {w2*1e3:8.2f} ms | kernels {k2:6.2f} ms | {n/w2/1e6:7.1f} Mpairs/s | sha {digest(o)} {digest(o2)}", flush=True)
    # config 4: all-pairs scene, device-generated pair list, results to pageable memory
    from test_gpu_operators import scene_bodies
    t, shape, r, att = scene_bodies(4096, 0, spread=5.0)
    ctx.set_shape_arrays(*t.arrays())
    w, k, o = timed(lambda: ctx.scene_allpairs(1, shape, r, att, 0), 2)
    print(f"{'config 4 scene allpairs':24s} fresh outputs {w*1e3:8.2f} ms | kernels {k:6.2f} ms | sha {digest(o)}", flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    child()
else:
    for mode in ("1", "0"):
        print(f"=== DCOL_HOST_PIPE={mode}", flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, DCOL_HOST_PIPE=mode))
    for th in ("2", "4", "16"):
        print(f"=== DCOL_HOST_PIPE=1 DCOL_HOST_THREADS={th}", flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, DCOL_HOST_PIPE="1", DCOL_HOST_THREADS=th))
