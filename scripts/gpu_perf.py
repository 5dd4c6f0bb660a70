"""Development aid: kernel time per primitive-type pair (device-resident inputs) and for the mixed batch."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dcol_b200 as dc
import util

names = ["polytope", "capsule", "cylinder", "cone", "sphere", "ellipsoid", "polygon"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
combos = [("capsule", "capsule"), ("polytope", "polytope"), ("polygon", "polygon"), ("sphere", "sphere"), ("cone", "ellipsoid"),
          ("cylinder", "polytope"), ("mixed", "mixed")]
ctx = dc.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
res = {}
for kind, att in ((0, 0), (2, 1)):
    na = 4 if att == 0 else 3; nt = 6 + 2 * na
    shapes = util.standard_shapes(att, "cube")
    t = dc.ShapeTable(); ids = np.array([t.add(s) for s in shapes], np.int32)
    ctx.set_shape_arrays(*t.arrays())
    r1, a1, r2, a2 = util.random_poses(n, att, 1234)
    g = util.rng(7)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    dr1, da1, dr2, da2 = T(r1), T(a1), T(r2), T(a2)
    alpha = torch.empty(n, dtype=torch.float64, device=dev); x = torch.empty((n, 3), dtype=torch.float64, device=dev)
    J = torch.empty((n, nt, 4), dtype=torch.float64, device=dev) if kind == 2 else None
    st = torch.empty(n, dtype=torch.int32, device=dev); it = torch.empty(n, dtype=torch.int32, device=dev)
    for c in combos:
        if c[0] == "mixed":
            s1 = ids[g.integers(0, 7, n)]; s2 = ids[g.integers(0, 7, n)]
        else:
            s1 = np.full(n, ids[names.index(c[0])], np.int32); s2 = np.full(n, ids[names.index(c[1])], np.int32)
        ds1, ds2 = T(s1), T(s2)
        ms = []
        for rep in range(4):
            ctx.query_raw(kind, n, ds1, ds2, dr1, da1, dr2, da2, att, 1e-6, dc.MEM_DEVICE, alpha, x, J, st, it)
            ms.append(ctx.last_kernel_ms())
        m = min(ms[1:])
        its = it.cpu().numpy()
        print(f"kind={kind} {c[0]:>9s}-{c[1]:<9s} n={n} kernel {m:8.3f} ms  {n/m/1e3:8.1f} Mpairs/s  mean iters {its.mean():.2f} max {its.max()}  ns/pair-iter {1e6*m/ (its.sum()):.3f}"
              f"  failed {(st.cpu().numpy()!=0).sum()}", flush=True)
