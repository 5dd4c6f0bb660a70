"""Development aid: device-resident kernel+call time of the 1M mixed proximity batch and the 2M mixed MRP Jacobian batch
(whole-call time with CUDA events on the launching stream, min of 5)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dcol_b200 as dc
import util
dev = torch.device("cuda:0")
ctx = dc.Context(0)
stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
res = []
for kind, att, n, variant in ((0, 0, 1000000, "cube"), (0, 0, 1000000, "scene"), (2, 1, 2000000, "cube")):
    na = 4 if att == 0 else 3; nt = 6 + 2 * na
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att, seed=1234, variant=variant)
    ctx.set_shape_arrays(*t.arrays())
    d = [T(a) for a in (s1, s2, r1, a1, r2, a2)]
    alpha = torch.empty(n, dtype=torch.float64, device=dev); x = torch.empty((n, 3), dtype=torch.float64, device=dev)
    J = torch.empty((n, nt, 4), dtype=torch.float64, device=dev) if kind == 2 else None
    st = torch.empty(n, dtype=torch.int32, device=dev); it = torch.empty(n, dtype=torch.int32, device=dev)
    best = 1e9; bk = 1e9
    for rep in range(6):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.query_raw(kind, n, *d, att, 1e-6, dc.MEM_DEVICE, alpha, x, J, st, it)
        e1.record(stream); torch.cuda.synchronize()
        if rep: best = min(best, e0.elapsed_time(e1)); bk = min(bk, ctx.last_kernel_ms())
    res.append(f"{'prox' if kind == 0 else 'jac'}-{variant} call {best:.3f} ms solver {bk:.3f} ms")
print(" | ".join(res))
