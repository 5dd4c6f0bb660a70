"""Parity table (run on the GPU box): CUDA path vs the CPU oracle per class pair x query x attitude kind, with the
oracle's own sensitivity to a 1-ulp perturbation of its inputs beside it, so that the intrinsic (conditioning) tail is
separated from implementation error.  Writes gpurun_out/parity_r02.json (copied to profiles/ by hand).

err      : row max-norm relative error  |gpu - oracle|_inf / max(|oracle|_inf, 1e-12)   (x: per pair, J/grad: per pair)
err_elem : per-element error with an absolute floor  |gpu - oracle| / max(|oracle_ij|, 1e-3 * |oracle|_inf)
spread   : max over 3 perturbed oracle runs (every pose entry times 1 + 2.2e-16 * U(-1,1)) of the same row error
           between perturbed and unperturbed oracle outputs
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import dcol_b200 as dc
import util
from oracle import oracle as orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
names = ["polytope", "capsule", "cylinder", "cone", "sphere", "ellipsoid", "polygon"]
ctx = dc.Context(0)
out = {"n_per_config": N, "pdip_tol": 1e-6, "tolerances": {"alpha": 1e-6, "x": 1e-6, "deriv": 1e-5}, "rows": []}


def elem_err(a, b):
    ax = tuple(range(1, a.ndim))
    scale = np.abs(b).max(axis=ax, keepdims=True)
    return (np.abs(a - b) / np.maximum(np.abs(b), np.maximum(1e-3 * scale, 1e-12))).max(axis=ax)


def stats(e):
    return {"max": float(e.max()), "p999": float(np.quantile(e, 0.999)), "median": float(np.median(e))}


for variant in ("cube", "scene"):
    for att in (0, 1):
        t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(N, att, seed=4321, variant=variant)
        ids = np.unique(s1)
        ctx.set_shape_arrays(*t.arrays())
        for kind, qname in ((0, "proximity"), (1, "gradient"), (2, "jacobian")):
            g = ctx._batch(kind, s1, s2, r1, a1, r2, a2, att, 1e-6)
            o = orc.query(kind, t.arrays(), s1, s2, r1, a1, r2, a2, att, 1e-6)
            key = "x" if kind == 0 else ("grad" if kind == 1 else "J")
            sp_main = np.zeros(N); sp_alpha = np.zeros(N)
            gen = util.rng(99)
            for trial in range(3):
                pr = [a * (1.0 + 2.2e-16 * (2 * gen.random(a.shape) - 1)) for a in (r1, a1, r2, a2)]
                op = orc.query(kind, t.arrays(), s1, s2, pr[0], pr[1], pr[2], pr[3], att, 1e-6)
                sp_main = np.maximum(sp_main, util.rel_err(op[key], o[key]))
                sp_alpha = np.maximum(sp_alpha, util.rel_err(op["alpha"], o["alpha"]))
            e_alpha = util.rel_err(g["alpha"], o["alpha"])
            e_main = util.rel_err(g[key], o[key]); e_elem = elem_err(np.asarray(g[key]), np.asarray(o[key]))
            tol = 1e-6 if kind == 0 else 1e-5
            # class pair of every pair (unordered type names)
            t1 = np.searchsorted(ids, s1); t2 = np.searchsorted(ids, s2)
            lo = np.minimum(t1, t2); hi = np.maximum(t1, t2)
            for a in range(7):
                for b in range(a, 7):
                    m = (lo == a) & (hi == b)
                    if not m.any():
                        continue
                    bad = m & (e_main >= tol)
                    ratio = e_main[bad] / np.maximum(sp_main[bad], 1e-16)
                    out["rows"].append({
                        "shapes": variant, "att": "quat" if att == 0 else "mrp", "query": qname, "pair": f"{names[a]}-{names[b]}",
                        "n": int(m.sum()), "status_mismatch": int((g["status"] != o["status"])[m].sum()),
                        "iters_mismatch": int((g["iters"] != o["iters"])[m].sum()),
                        "class_mismatch": int(((g["alpha"] <= 1) != (o["alpha"] <= 1))[m].sum()),
                        "alpha": stats(e_alpha[m]), key: stats(e_main[m]), key + "_elem": stats(e_elem[m]),
                        "n_over_tol": int(bad.sum()),
                        "oracle_spread_" + key: stats(sp_main[m]), "oracle_spread_alpha": stats(sp_alpha[m]),
                        "n_oracle_spread_over_tol": int((m & (sp_main >= tol)).sum()),
                        "max_err_over_spread_among_over_tol": float(ratio.max()) if bad.any() else 0.0,
                    })
            tot_bad = e_main >= tol
            print(f"{variant:5s} att={att} {qname:9s}: alpha max {e_alpha.max():.2e}  {key} max {e_main.max():.2e} p99.9 {np.quantile(e_main, 0.999):.2e} "
                  f"over tol {tot_bad.sum()} ({100 * tot_bad.mean():.3f}%)  oracle-spread over tol {(sp_main >= tol).sum()}  "
                  f"iters mismatch {(g['iters'] != o['iters']).sum()}  elem max {e_elem.max():.2e}", flush=True)
            if tot_bad.any():
                r = e_main[tot_bad] / np.maximum(sp_main[tot_bad], 1e-16)
                print(f"      err/spread among over-tol pairs: median {np.median(r):.1f} p90 {np.quantile(r, 0.9):.1f} max {r.max():.1f}", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_r02.json"), "w"), indent=0)
print("wrote gpurun_out/parity_r02.json with", len(out["rows"]), "rows")
