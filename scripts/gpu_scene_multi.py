"""BASELINE config 4 (4096 bodies, 8 386 560 i<j pairs, proximity_gradient) over several GPUs (body arrays on
every device, one contiguous slice of the pair enumeration each).
  python scripts/gpu_scene_multi.py                  one process, dcol_group over 1, 2, 4, 8 visible devices
  torchrun --nproc-per-node N scripts/gpu_scene_multi.py   one rank per GPU: dcol_scene_wrench_range + ONE NCCL all-reduce
                                                     of the per-body partial sums (dc.scene_wrench_sharded)"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import dcol_b200 as dc
import util
from test_gpu_operators import scene_bodies

NB = int(os.environ.get("DCOL_SCENE_BODIES", "4096"))
t, shape, r, att = scene_bodies(NB, 0, spread=5.0)
n = NB * (NB - 1) // 2
lam = util.rng(11).uniform(0.0, 2.0, n)


def best_of(f, reps=3):
    f(); b = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); o = f(); b = min(b, time.perf_counter() - t0)
    return b, o


if "RANK" in os.environ:
    import torch.distributed as dist
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = dc.Context(lr); ctx.set_shape_arrays(*t.arrays())
    for cull in (0.0, 1.0):
        ctx.scene_set_cull(cull)
        def step():
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            w, mine, rng_ = dc.scene_wrench_sharded(ctx, rank, world, shape, r, att, lam, device=torch.device("cuda", lr), per_pair=False)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device="cuda"); dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item()), w, mine
        step(); res = [step() for _ in range(3)]
        dt = min(x[0] for x in res); w = res[-1][1]
        ref = None
        if rank == 0:
            ref = ctx.scene_wrench(shape, r, att, lam, 0, per_pair=False)["wrench"]
            err = np.abs(w - ref).max() / np.abs(ref).max()
            print(f"torchrun x{world}  cull alpha_cut={cull}: scene_wrench_sharded {dt*1e3:8.2f} ms per call (max over ranks; {n/dt/1e6:7.1f} Mpairs/s), "
                  f"solver {ctx.last_kernel_ms():6.2f} ms on rank 0, all-reduced wrench vs one GPU: rel. diff {err:.1e}", flush=True)
    dist.barrier(); dist.destroy_process_group()
    sys.exit(0)

nd = torch.cuda.device_count()
print("devices visible:", nd, "cpus:", len(os.sched_getaffinity(0)), "bodies:", NB, "pairs:", n)
ref = {}
for m in sorted({1, 2, 4, 8} & set(range(1, nd + 1))):
    g = dc.Group(list(range(m))); g.set_shape_arrays(*t.arrays())
    for cull in (0.0, 1.0):
        g.scene_set_cull(cull)
        bw, w = best_of(lambda: g.scene_wrench(shape, r, att, lam, 0, per_pair=False))
        ms_w = g.last_kernel_ms()
        bp, p = best_of(lambda: g.scene_allpairs(1, shape, r, att, 0), reps=2)
        if (cull, "w") not in ref:
            ref[(cull, "w")] = w["wrench"]; ref[(cull, "p")] = p
        dw = np.abs(w["wrench"] - ref[(cull, "w")]).max() / np.abs(ref[(cull, "w")]).max()
        same = all(np.array_equal(p[k], ref[(cull, "p")][k], equal_nan=True) for k in p)
        print(f"group of {m}  cull alpha_cut={cull}: scene_wrench {bw*1e3:8.2f} ms per call ({n/bw/1e6:7.1f} Mpairs/s, solver {ms_w:6.2f} ms, wrench vs 1 GPU rel. diff {dw:.1e})"
              f" | scene_allpairs gradient rows to pageable arrays {bp*1e3:8.2f} ms, identical to 1 GPU: {same}", flush=True)
    g.close()
