"""Development aid: kernel time vs batch size (intercept = fixed per-call cost)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, time
import dcol_b200 as dc
import util
names = ["polytope", "capsule", "cylinder", "cone", "sphere", "ellipsoid", "polygon"]
ctx = dc.Context(0); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
att = 0
shapes = util.standard_shapes(att, "cube")
t = dc.ShapeTable(); ids = np.array([t.add(s) for s in shapes], np.int32)
ctx.set_shape_arrays(*t.arrays())
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for combo in ("mixed", "capsule", "polytope"):
    for n in (1000, 10000, 50000, 100000, 200000, 400000, 1000000, 2000000, 4000000):
        r1, a1, r2, a2 = util.random_poses(n, att, 1234)
        g = util.rng(7)
        if combo == "mixed":
            s1 = ids[g.integers(0, 7, n)]; s2 = ids[g.integers(0, 7, n)]
        else:
            s1 = np.full(n, ids[names.index(combo)], np.int32); s2 = s1.copy()
        d = [T(a) for a in (s1, s2, r1, a1, r2, a2)]
        alpha = torch.empty(n, dtype=torch.float64, device=dev); x = torch.empty((n, 3), dtype=torch.float64, device=dev)
        st = torch.empty(n, dtype=torch.int32, device=dev); it = torch.empty(n, dtype=torch.int32, device=dev)
        ms = []; wall = []
        for rep in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ctx.query_raw(0, n, *d, att, 1e-6, dc.MEM_DEVICE, alpha, x, None, st, it)
            wall.append(1e3 * (time.perf_counter() - t0)); ms.append(ctx.last_kernel_ms())
        print(f"{combo:9s} n={n:8d} kernels {min(ms[1:]):8.3f} ms  call wall {min(wall[1:]):8.3f} ms  {n/min(wall[1:])/1e3:8.1f} Mpairs/s (wall)", flush=True)
