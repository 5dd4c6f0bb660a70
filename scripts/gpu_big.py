"""Maximum-size sanity (one call far above the bench's batch): N device-resident mixed pairs generated on the GPU, one
dcol_proximity / dcol_proximity_jacobian call, every status 0, and random windows of the result bit-identical to small
calls on the same pairs (32-bit index arithmetic, scratch sizing, tile tables at 10^8 pairs)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dcol_b200 as dc
import util

dev = torch.device("cuda", 0)
ctx = dc.Context(0)
for kind, att_kind, n in ((0, 0, int(os.environ.get("DCOL_BIG_PROX", 200_000_000))), (2, 1, int(os.environ.get("DCOL_BIG_JAC", 40_000_000)))):
    na = 4 if att_kind == 0 else 3; nt = 6 + 2 * na
    table = dc.ShapeTable()
    ids = torch.tensor([table.add(p) for p in util.standard_shapes(att_kind, "cube")], dtype=torch.int32, device=dev)
    ctx.set_shape_arrays(*table.arrays())
    g = torch.Generator(device=dev); g.manual_seed(1234 + kind)
    s1 = ids[torch.randint(0, 7, (n,), device=dev, generator=g)].contiguous()
    s2 = ids[torch.randint(0, 7, (n,), device=dev, generator=g)].contiguous()
    r1 = 2.0 * torch.randn((n, 3), dtype=torch.float64, device=dev, generator=g)
    r2 = torch.randn((n, 3), dtype=torch.float64, device=dev, generator=g)
    def att():
        q = torch.randn((n, 4), dtype=torch.float64, device=dev, generator=g)
        q /= q.norm(dim=1, keepdim=True)
        if att_kind == 0:
            return q
        q = torch.where(q[:, :1] < 0, -q, q)
        return (q[:, 1:] / (1 + q[:, :1])).contiguous()
    a1, a2 = att(), att()
    alpha = torch.empty(n, dtype=torch.float64, device=dev); x = torch.empty((n, 3), dtype=torch.float64, device=dev)
    J = torch.empty((n, nt, 4), dtype=torch.float64, device=dev) if kind == 2 else None
    st = torch.empty(n, dtype=torch.int32, device=dev); it = torch.empty(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.query_raw(kind, n, s1, s2, r1, a1, r2, a2, att_kind, 1e-6, dc.MEM_DEVICE, alpha, x, J, st, it)
    dt = time.perf_counter() - t0
    ok = int((st == 0).sum().item())
    print(f"kind {kind}: n = {n}: one call {dt*1e3:.1f} ms wall, solver {ctx.last_kernel_ms():.1f} ms ({n/ctx.last_kernel_ms()/1e3:.1f} Mpairs/s), status 0 on {ok} of {n}, "
          f"mean iterations {it.double().mean().item():.3f}, peak device memory {torch.cuda.max_memory_allocated()/2**30:.1f} GiB (torch side)", flush=True)
    assert ok == n
    same = True
    for w0 in (0, n // 3 + 12345, n - 150_000):
        w1 = w0 + 150_000
        b_alpha = torch.empty(w1 - w0, dtype=torch.float64, device=dev); b_x = torch.empty((w1 - w0, 3), dtype=torch.float64, device=dev)
        b_J = torch.empty((w1 - w0, nt, 4), dtype=torch.float64, device=dev) if kind == 2 else None
        b_st = torch.empty(w1 - w0, dtype=torch.int32, device=dev); b_it = torch.empty(w1 - w0, dtype=torch.int32, device=dev)
        ctx.query_raw(kind, w1 - w0, s1[w0:w1], s2[w0:w1], r1[w0:w1], a1[w0:w1], r2[w0:w1], a2[w0:w1], att_kind, 1e-6, dc.MEM_DEVICE, b_alpha, b_x, b_J, b_st, b_it)
        same = same and torch.equal(b_alpha, alpha[w0:w1]) and torch.equal(b_x, x[w0:w1]) and torch.equal(b_it, it[w0:w1]) and (kind != 2 or torch.equal(b_J, J[w0:w1]))
    print(f"kind {kind}: three 150 000-pair windows bit-identical to separate calls: {same}", flush=True)
    assert same
    del s1, s2, r1, r2, a1, a2, alpha, x, J, st, it
    torch.cuda.empty_cache()
