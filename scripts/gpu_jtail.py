"""Development aid: Jacobian tail against the oracle (100k mixed pairs per shape set, quaternion and MRP) for the library
selected by DCOL_B200_LIB — used for the before/after of partial pivoting in the Jacobian epilogue."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import dcol_b200 as dc
import util
from oracle import oracle as orc
ctx = dc.Context(0)
n = 100000
for variant in ("cube", "scene"):
    for att in (0, 1):
        t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, att, seed=4321, variant=variant)
        ctx.set_shape_arrays(*t.arrays())
        g = ctx._batch(2, s1, s2, r1, a1, r2, a2, att, 1e-6)
        o = orc.query(2, t.arrays(), s1, s2, r1, a1, r2, a2, att, 1e-6)
        e = util.rel_err(g["J"], o["J"])
        print(f"{os.environ.get('DCOL_B200_LIB', 'default'):40s} {variant:5s} att={att}: J over 1e-5: {(e >= 1e-5).sum():4d} of {n}  max {e.max():.3e}  p99.99 {np.quantile(e, 0.9999):.3e}  p99.9 {np.quantile(e, 0.999):.3e}")
