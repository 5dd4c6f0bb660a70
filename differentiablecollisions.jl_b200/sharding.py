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
