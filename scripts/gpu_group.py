"""Single-process multi-GPU through dcol_group: blocking calls on pageable numpy arrays, 1..N devices."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dcol_b200 as dc
import util
nd = torch.cuda.device_count()
print("devices visible:", nd, "cpus:", len(os.sched_getaffinity(0)))
for kind, att, n in ((0, 0, 4000000), (2, 1, 2000000)):
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att, seed=3)
    ref = None
    for m in sorted({1, 2, 4, 8} & set(range(1, nd + 1))):
        g = dc.Group(list(range(m)))
        g.set_shape_arrays(*t.arrays())
        g._batch(kind, s1, s2, r1, a1, r2, a2, att, 1e-6)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); o = g._batch(kind, s1, s2, r1, a1, r2, a2, att, 1e-6); best = min(best, time.perf_counter() - t0)
        if ref is None:
            ref = o
        same = all(np.array_equal(o[k], ref[k], equal_nan=True) for k in o)
        print(f"{'proximity' if kind == 0 else 'jacobian '} n={n} group of {m}: {best*1e3:8.2f} ms per call  {n/best/1e6:8.1f} Mpairs/s  solver {g.last_kernel_ms():6.2f} ms  identical to 1 GPU: {same}", flush=True)
        # the same call on page-locked arrays (what a caller that reuses its buffers should do): PCIe-bound per device
        na = 4 if att == 0 else 3; nt = 6 + 2 * na
        pins = [np.ascontiguousarray(a).copy() for a in (s1, s2, r1, a1, r2, a2)]
        out = {"alpha": np.empty(n), "status": np.empty(n, np.int32), "iters": np.empty(n, np.int32), "x": np.empty((n, 3))}
        if kind == 2: out["deriv"] = np.empty((n, nt, 4))
        for a in pins + list(out.values()): g.pin(a)
        g._batch(kind, *pins, att, 1e-6, out=out)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); o = g._batch(kind, *pins, att, 1e-6, out=out); best = min(best, time.perf_counter() - t0)
        same = all(np.array_equal(o[k], ref[k], equal_nan=True) for k in o)
        print(f"{'proximity' if kind == 0 else 'jacobian '} n={n} group of {m}, page-locked arrays: {best*1e3:8.2f} ms per call  {n/best/1e6:8.1f} Mpairs/s  identical: {same}", flush=True)
        for a in pins + list(out.values()): g.unpin(a)
        g.close()
