import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import dcol_b200 as dc, util
n = 1000000
ctx = dc.Context(0); dev = torch.device("cuda:0")
for kind, att in ((0, 0), (1, 0), (2, 0)):
    na = 4 if att == 0 else 3; nt = 6 + 2 * na
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att, seed=1234)
    ctx.set_shape_arrays(*t.arrays())
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d = [T(a) for a in (s1, s2, r1, a1, r2, a2)]
    alpha = torch.empty(n, dtype=torch.float64, device=dev); x = torch.empty((n, 3), dtype=torch.float64, device=dev)
    der = torch.empty((n, nt) if kind == 1 else (n, nt, 4), dtype=torch.float64, device=dev) if kind else None
    st = torch.empty(n, dtype=torch.int32, device=dev); it = torch.empty(n, dtype=torch.int32, device=dev)
    ms = []
    for rep in range(4):
        ctx.query_raw(kind, n, *d, att, 1e-6, dc.MEM_DEVICE, alpha, x if kind != 1 else None, der, st, it)
        ms.append(ctx.last_kernel_ms())
    print("kind", kind, "1M mixed quat: kernel %.3f ms -> %.1f Mpairs/s" % (min(ms[1:]), n / min(ms[1:]) / 1e3))
