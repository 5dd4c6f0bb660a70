"""Development aid: where the pageable-memory pipeline starts to pay.  Blocking host-pointer calls from pageable numpy
arrays (reused output buffers), pipeline on / off, over a range of call sizes."""
import sys, os, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

def child():
    import numpy as np
    import dcol_b200 as dc
    import util
    ctx = dc.Context(0)
    for kind, att in ((0, 0), (2, 1)):
        na = 4 if att == 0 else 3; nt = 6 + 2 * na
        row = []
        for n in ((60000, 100000, 150000, 200000, 300000, 500000) if kind == 0 else (20000, 40000, 70000, 100000, 200000)):
            t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att, seed=5)
            ctx.set_shape_arrays(*t.arrays())
            out = {"alpha": np.zeros(n), "status": np.zeros(n, np.int32), "iters": np.zeros(n, np.int32), "x": np.zeros((n, 3))}
            if kind == 2: out["deriv"] = np.zeros((n, nt, 4))
            f = lambda: ctx._batch(kind, s1, s2, r1, a1, r2, a2, att, 1e-6, out=out)
            f(); f(); best = 1e9
            for _ in range(8):
                t0 = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0)
            mb = n * (8 + 8 * (6 + 2 * na) + 40 + (32 * nt if kind == 2 else 0)) / 1e6
            row.append(f"n={n} ({mb:.0f} MB) {best*1e3:.2f} ms")
        print(("prox quat: " if kind == 0 else "jac mrp:   ") + " | ".join(row), flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    child()
else:
    for mode in ("1", "0"):
        print(f"=== DCOL_HOST_PIPE={mode}", flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, DCOL_HOST_PIPE=mode))
