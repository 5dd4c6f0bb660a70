"""Development aid: GPU-vs-oracle diagnostics over tolerances and extreme poses."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import dcol_b200 as dc
from oracle import oracle as orc
import util

ctx = dc.Context(0)
names = ["polytope", "capsule", "cylinder", "cone", "sphere", "ellipsoid", "polygon"]

def report(tag, g, o, s1, s2, types):
    print(f"== {tag}: status gpu {np.bincount(g['status'], minlength=3)} oracle {np.bincount(o['status'], minlength=3)}"
          f" iters eq {(g['iters'] == o['iters']).mean():.4f}")
    ok = (o["status"] == 0) & (g["status"] == 0)
    ea = util.rel_err(g["alpha"], o["alpha"])
    print("   alpha relerr max %.2e p99.9 %.2e" % (ea[ok].max(), np.quantile(ea[ok], .999)))
    if "x" in g:
        ex = util.rel_err(g["x"], o["x"]); print("   x relerr max %.2e frac>1e-6 %.2e" % (ex[ok].max(), (ex[ok] > 1e-6).mean()))
    if "J" in g:
        ej = util.rel_err(g["J"], o["J"]); print("   J relerr max %.2e frac>1e-5 %.2e" % (ej[ok].max(), (ej[ok] > 1e-5).mean()))
    bad = np.where(g["status"] != o["status"])[0][:8]
    for b in bad:
        print("   status mismatch pair", b, names[types[s1[b]]], names[types[s2[b]]], "gpu", g["status"][b], g["iters"][b], "oracle", o["status"][b], o["iters"][b], o["alpha"][b])
    worst = np.argsort(-np.where(ok, ea, 0))[:4]
    for b in worst:
        print("   worst alpha", b, names[types[s1[b]]], names[types[s2[b]]], g["alpha"][b], o["alpha"][b], g["iters"][b], o["iters"][b])

n = 20000
for variant in ("cube", "scene"):
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, 0, seed=5, variant=variant)
    types = t.arrays()[0]
    ctx.set_shape_arrays(*t.arrays())
    for tol in (1e-6, 1e-8, 1e-10, 1e-12):
        g = ctx._batch(2, s1, s2, r1, a1, r2, a2, 0, tol)
        o = orc.query(2, t.arrays(), s1, s2, r1, a1, r2, a2, 0, tol)
        report(f"{variant} tol={tol}", g, o, s1, s2, types)
    for tag, r2x in (("coincident", r1.copy()), ("far1e3", r1 + 1e3), ("far30", r1 + 30.0)):
        g = ctx._batch(0, s1, s2, r1, a1, r2x, a2, 0, 1e-6)
        o = orc.query(0, t.arrays(), s1, s2, r1, a1, r2x, a2, 0, 1e-6)
        report(f"{variant} {tag}", g, o, s1, s2, types)
