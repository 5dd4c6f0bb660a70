"""Blocking host-pointer calls on page-locked arrays (bench.py's e2e.one_call_at_a_time): chunk size of the in-call pipeline."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dcol_b200 as dc
import util
ctx = dc.Context(0)
for kind, att, n in ((0, 0, 1_000_000), (2, 1, 2_000_000), (0, 0, 4_000_000)):
    na = 4 if att == 0 else 3; nt = 6 + 2 * na
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att, seed=5)
    ctx.set_shape_arrays(*t.arrays())
    host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (s1, s2, r1, a1, r2, a2)]
    alpha = torch.empty(n, dtype=torch.float64).pin_memory(); x = torch.empty((n, 3), dtype=torch.float64).pin_memory()
    J = torch.empty((n, nt, 4), dtype=torch.float64).pin_memory() if kind == 2 else None
    st = torch.empty(n, dtype=torch.int32).pin_memory(); it = torch.empty(n, dtype=torch.int32).pin_memory()
    row = []
    for chunk in (None, 250000, 175000, 125000, 90000, 60000):
        if chunk is None: os.environ.pop("DCOL_CHUNK", None)
        else: os.environ["DCOL_CHUNK"] = str(chunk)
        f = lambda: ctx.query_raw(kind, n, *host, att, 1e-6, dc.MEM_HOST, alpha, x, J, st, it)
        f(); f()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            for _ in range(4): f()
            best = min(best, (time.perf_counter() - t0) / 4)
        row.append(f"chunk {chunk or 'default'}: {best*1e3:6.3f} ms ({n/best/1e6:6.1f} Mpairs/s)")
    print(("proximity" if kind == 0 else "jacobian ") + f" n={n}: " + " | ".join(row), flush=True)
