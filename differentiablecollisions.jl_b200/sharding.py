"""Index sharding of a pair batch over the GPUs of one box.

Every pair query is independent, so the N>1 path is: rank g takes the g-th contiguous slice of the
pair index range, runs the same C-ABI call on its own GPU, and — only if the caller wants the rows on
every rank — an all-gather of the result rows (NCCL over NVLink on GPUs, gloo in the CPU tests).
There is no collective on the data path itself.
"""
from __future__ import annotations


def shard_bounds(n: int, world: int):
    """Contiguous [begin, end) per rank; sizes differ by at most one; empty shards allowed when n < world."""
    if world < 1 or n < 0:
        raise ValueError("need world >= 1 and n >= 0")
    base, rem = divmod(n, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < rem else 0)
        out.append((b, e))
        b = e
    return out


def shard(n: int, rank: int, world: int):
    return shard_bounds(n, world)[rank]


def gather_rows(local, n_total: int, rank: int, world: int, group=None):
    """All-gather per-pair result rows (first dim = this rank's shard) into the full (n_total, ...) tensor on every
    rank.  Shards are padded to the largest shard so a plain all_gather works on both NCCL and gloo."""
    import torch
    import torch.distributed as dist
    bounds = shard_bounds(n_total, world)
    b, e = bounds[rank]
    if local.shape[0] != e - b:
        raise ValueError(f"rank {rank}: local rows {local.shape[0]} != shard size {e - b}")
    if world == 1:
        return local
    mx = max(e2 - b2 for b2, e2 in bounds)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: e - b] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][: bounds[r][1] - bounds[r][0]] for r in range(world)], dim=0)


def scene_wrench_sharded(ctx, rank: int, world: int, shape, r, att, lam, att_kind=0, pdip_tol=1e-6, group=None, device=None, per_pair=True):
    """The contact-scene accumulation (`dcol_scene_wrench`, examples/contact_physics/ncp_contact_simulation.jl:133-139) over
    `world` ranks: every rank holds the body arrays and the full weight vector `lam`, solves its contiguous slice of the
    i<j enumeration on its own GPU (`dcol_scene_wrench_range`) and contributes per-body partial sums; ONE all-reduce (sum)
    of the n_bodies x 6 partial sums is the exchange step of this operator (NCCL over NVLink when `device` is a CUDA
    device, gloo on CPU tensors).  Returns (wrench on every rank, this rank's per-pair dict (alpha / status / iters of its slice unless per_pair=False), (k0, k1))."""
    import numpy as np
    import torch
    import torch.distributed as dist
    nb = int(np.asarray(shape).reshape(-1).shape[0])
    k0, k1 = shard(nb * (nb - 1) // 2, rank, world)
    lam = np.asarray(lam, dtype=np.float64).reshape(-1)
    out = ctx.scene_wrench(shape, r, att, lam[k0:k1], att_kind, pdip_tol, per_pair=per_pair, pair_range=(k0, k1))
    w = torch.from_numpy(out["wrench"])
    if world > 1:
        if device is not None:
            w = w.to(device)
        dist.all_reduce(w, op=dist.ReduceOp.SUM, group=group)
        w = w.cpu()
    return w.numpy(), out, (k0, k1)


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Best effort: pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers
    allocated afterwards are local to the GPU's PCIe root (matters when 8 ranks stream from host memory at once).
    Returns what was done; never raises."""
    import os
    info = {"bound": False}
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        path = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(os.path.join(path, "numa_node")).read().strip())
        cpulist = open(os.path.join(path, "local_cpulist")).read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        info.update(numa_node=node, cpus=len(cpus))
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
    except Exception as e:  # noqa: BLE001
        info["error"] = str(e)
    return info


def bind_disjoint_cpus(local_rank: int, local_world: int) -> dict:
    """Fallback when the box exposes no NUMA topology (numa_node = -1 in a VM): give every rank of the box its own
    contiguous slice of the allowed CPUs, so that the ranks' host-side work (submission threads, pinned-buffer first
    touch) does not migrate over each other.  Returns what was done; never raises."""
    import os
    info = {"bound": False}
    try:
        allowed = sorted(os.sched_getaffinity(0))
        if local_world > 1 and len(allowed) >= local_world:
            k = len(allowed) // local_world
            mine = allowed[local_rank * k:(local_rank + 1) * k]
            os.sched_setaffinity(0, mine)
            info.update(bound=True, cpus=len(mine), first_cpu=mine[0], policy="disjoint slice per rank")
    except Exception as e:  # noqa: BLE001
        info["error"] = str(e)
    return info
