"""Ad-hoc GPU-vs-oracle comparison (development aid; the judged tests are tests/ -m gpu)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import dcol_b200 as dc
from oracle import oracle as orc
import util

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
kinds = [int(k) for k in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]
ctx = dc.Context(0)
for att in (0, 1):
    table, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att, variant="cube")
    ctx.table = table; ctx._table_version = -1
    for kind in kinds:
        t0 = time.time()
        g = ctx._batch(kind, s1, s2, r1, a1, r2, a2, att, 1e-6)
        t1 = time.time()
        o = orc.query(kind, table.arrays(), s1, s2, r1, a1, r2, a2, att, 1e-6)
        t2 = time.time()
        ea = util.rel_err(g["alpha"], o["alpha"])
        print(f"att={att} kind={kind} n={n} gpu {t1-t0:.3f}s (kernels {ctx.last_kernel_ms():.3f} ms) oracle {t2-t1:.3f}s")
        print("  status gpu", np.bincount(g["status"], minlength=3), "oracle", np.bincount(o["status"], minlength=3))
        print("  iters equal frac", (g["iters"] == o["iters"]).mean(), "mean iters", g["iters"].mean(), "max", g["iters"].max())
        print("  alpha rel err max %.3e  p99.9 %.3e" % (ea.max(), np.quantile(ea, 0.999)))
        print("  class mismatch", int(((g["alpha"] <= 1) != (o["alpha"] <= 1)).sum()))
        if "x" in g:
            ex = util.rel_err(g["x"], o["x"])
            print("  x rel err max %.3e p99.9 %.3e frac>1e-6 %.2e" % (ex.max(), np.quantile(ex, 0.999), (ex > 1e-6).mean()))
        if "grad" in g:
            eg = util.rel_err(g["grad"], o["grad"])
            print("  grad rel err max %.3e p99.9 %.3e frac>1e-5 %.2e" % (eg.max(), np.quantile(eg, 0.999), (eg > 1e-5).mean()))
        if "J" in g:
            ej = util.rel_err(g["J"], o["J"])
            print("  J rel err max %.3e p99.9 %.3e frac>1e-5 %.2e" % (ej.max(), np.quantile(ej, 0.999), (ej > 1e-5).mean()))
        bad = np.argsort(-ea)[:3]
        for b in bad:
            print("   worst alpha", b, s1[b], s2[b], g["alpha"][b], o["alpha"][b], g["iters"][b], o["iters"][b], g["status"][b])
print("fp64 peak TFLOP/s", ctx.measure_fp64_tflops())
