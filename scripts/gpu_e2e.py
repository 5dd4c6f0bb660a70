"""Development aid: end-to-end (host pinned buffers) timing of proximity on 1M mixed pairs, sync and pipelined."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dcol_b200 as dc
import util
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dev = torch.device("cuda:0")
h = torch.empty(120 << 20, dtype=torch.uint8).pin_memory(); g = torch.empty(120 << 20, dtype=torch.uint8, device=dev)
for _ in range(2): g.copy_(h, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): g.copy_(h, non_blocking=True)
torch.cuda.synchronize(); print("H2D 120MB %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
ctx = dc.Context(0)
t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, 0, seed=1234)
ctx.set_shape_arrays(*t.arrays())
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
ins = [pin(a) for a in (s1, s2, r1, a1, r2, a2)]
outs = [[torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty((n, 3), dtype=torch.float64).pin_memory(),
         torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()] for _ in range(3)]
def sync_step(k):
    o = outs[k % 3]
    ctx.query_raw(0, n, *ins, 0, 1e-6, dc.MEM_HOST, o[0], o[1], None, o[2], o[3])
for k in range(3): sync_step(k)
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(10): sync_step(k)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("sync  e2e %.3f ms/step  %.1f Mpairs/s (kernels %.3f ms)" % (dt * 1e3, n / dt / 1e6, ctx.last_kernel_ms()))
if hasattr(ctx, "submit_raw") and hasattr(ctx._lib, "dcol_submit"):
    def sub(k):
        o = outs[k % 3]
        return ctx.submit_raw(0, n, *ins, 0, 1e-6, o[0], o[1], None, o[2], o[3])
    for k in range(3): ctx.wait(sub(k))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tk = []
    tsub = 0.0
    for k in range(10):
        if len(tk) >= 3: ctx.wait(tk[-3])
        ts = time.perf_counter(); tk.append(sub(k)); tsub += time.perf_counter() - ts
    for x in tk[-3:]: ctx.wait(x)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print("async e2e %.3f ms/step  %.1f Mpairs/s  (host time inside submit %.3f ms/step)" % (dt * 1e3, n / dt / 1e6, tsub / 10 * 1e3))
