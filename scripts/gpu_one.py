"""Development aid: one device-resident proximity call (for ncu).  usage: gpu_one.py <n> <mixed|type> [kind]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dcol_b200 as dc
import util
names = ["polytope", "capsule", "cylinder", "cone", "sphere", "ellipsoid", "polygon"]
n = int(sys.argv[1]); combo = sys.argv[2]; kind = int(sys.argv[3]) if len(sys.argv) > 3 else 0
att = 0 if kind == 0 else 1
na = 4 if att == 0 else 3; nt = 6 + 2 * na
ctx = dc.Context(0); dev = torch.device("cuda:0")
shapes = util.standard_shapes(att, "cube")
t = dc.ShapeTable(); ids = np.array([t.add(s) for s in shapes], np.int32)
ctx.set_shape_arrays(*t.arrays())
r1, a1, r2, a2 = util.random_poses(n, att, 1234)
g = util.rng(7)
if combo == "mixed":
    s1 = ids[g.integers(0, 7, n)]; s2 = ids[g.integers(0, 7, n)]
else:
    s1 = np.full(n, ids[names.index(combo)], np.int32); s2 = s1.copy()
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d = [T(a) for a in (s1, s2, r1, a1, r2, a2)]
alpha = torch.empty(n, dtype=torch.float64, device=dev); x = torch.empty((n, 3), dtype=torch.float64, device=dev)
J = torch.empty((n, nt, 4), dtype=torch.float64, device=dev) if kind == 2 else None
st = torch.empty(n, dtype=torch.int32, device=dev); it = torch.empty(n, dtype=torch.int32, device=dev)
for rep in range(3):
    ctx.query_raw(kind, n, *d, att, 1e-6, dc.MEM_DEVICE, alpha, x, J, st, it)
    print("kernel ms", ctx.last_kernel_ms())
