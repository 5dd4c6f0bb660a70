"""Solver time against batch size (device-resident, min of 4 calls): where the fixed part of a step goes.
Workloads: the mixed 7x7 batch and single type pairs; DCOL_SKIP_RESUME=1 in the environment leaves the stragglers unfinished."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dcol_b200 as dc
import util
dev = torch.device("cuda:0")
ctx = dc.Context(0)
stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
NMAX = 16_000_000
table = dc.ShapeTable()
ids = np.array([table.add(p) for p in util.standard_shapes(0, "cube")], dtype=np.int32)
ctx.set_shape_arrays(*table.arrays())
g = torch.Generator(device=dev); g.manual_seed(7)
r1 = 2.0 * torch.randn((NMAX, 3), dtype=torch.float64, device=dev, generator=g)
r2 = torch.randn((NMAX, 3), dtype=torch.float64, device=dev, generator=g)
q1 = torch.randn((NMAX, 4), dtype=torch.float64, device=dev, generator=g); q1 /= q1.norm(dim=1, keepdim=True)
q2 = torch.randn((NMAX, 4), dtype=torch.float64, device=dev, generator=g); q2 /= q2.norm(dim=1, keepdim=True)
mix1 = torch.from_numpy(ids).to(dev)[torch.randint(0, 7, (NMAX,), device=dev, generator=g)].contiguous()
mix2 = torch.from_numpy(ids).to(dev)[torch.randint(0, 7, (NMAX,), device=dev, generator=g)].contiguous()
alpha = torch.empty(NMAX, dtype=torch.float64, device=dev); x = torch.empty((NMAX, 3), dtype=torch.float64, device=dev)
st = torch.empty(NMAX, dtype=torch.int32, device=dev); it = torch.empty(NMAX, dtype=torch.int32, device=dev)
names = ["polytope", "capsule", "cylinder", "cone", "sphere", "ellipsoid", "polygon"]
loads = [("mixed", mix1, mix2)] + [(f"{names[a]}-{names[b]}", torch.full((NMAX,), int(ids[a]), dtype=torch.int32, device=dev), torch.full((NMAX,), int(ids[b]), dtype=torch.int32, device=dev))
                                   for a, b in ((1, 1), (0, 0), (6, 6), (4, 4))]
for name, s1, s2 in loads:
    row = []
    for n in (125_000, 250_000, 500_000, 1_000_000, 2_000_000, 4_000_000, 16_000_000):
        bk = 1e9; bc = 1e9
        for rep in range(5):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.query_raw(0, n, s1, s2, r1, q1, r2, q2, 0, 1e-6, dc.MEM_DEVICE, alpha, x, None, st, it)
            e1.record(stream); torch.cuda.synchronize()
            if rep: bk = min(bk, ctx.last_kernel_ms()); bc = min(bc, e0.elapsed_time(e1))
        row.append((n, bk, bc))
    # fit solver = a + b n on the two largest sizes
    (n1, k1, _), (n2, k2, _) = row[-2], row[-1]
    b = (k2 - k1) / (n2 - n1); a = k1 - b * n1
    print(f"{name:20s} " + " | ".join(f"n={n/1e6:5.3f}M solver {k:7.3f} call {c:7.3f} ({1e3*k/(n/1e6):6.1f} us/M)" for n, k, c in row))
    print(f"{'':20s} steady state {b*1e9:6.1f} us per 1M pairs, fixed part {a*1e3:6.1f} us; at 1M: {row[3][1]*1e3 - b*1e9:6.1f} us above steady state", flush=True)
