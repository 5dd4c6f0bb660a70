"""CPU, world_size 2 over gloo: the N>1 host path (index sharding + optional all-gather of result rows).
The per-rank compute is stood in for by the oracle (test infrastructure) — on GPUs each rank calls the C ABI."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dcol_b200 as dc
import util


def test_shard_bounds():
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = dc.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dc.shard_bounds(5, 0)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, 0, seed=3)
    b, e = dc.shard(n, rank, world)
    o = orc.query(2, t.arrays(), s1[b:e], s2[b:e], r1[b:e], a1[b:e], r2[b:e], a2[b:e], 0, 1e-6, nthreads=2)
    alpha = dc.gather_rows(torch.from_numpy(o["alpha"]), n, rank, world)
    J = dc.gather_rows(torch.from_numpy(o["J"]), n, rank, world)
    st = dc.gather_rows(torch.from_numpy(o["status"]), n, rank, world)
    # max-over-ranks reduction used by bench.py for the device time
    tmax = torch.tensor([float(rank + 1)]); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((alpha.numpy(), J.numpy(), st.numpy(), float(tmax.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 1])
def test_two_rank_shard_and_gather_matches_single(oracle, n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    alpha, J, st, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t, s1, s2, r1, a1, r2, a2 = util.mixed_batch(n, 0, seed=3)
    ref = oracle.query(2, t.arrays(), s1, s2, r1, a1, r2, a2, 0, 1e-6)
    assert np.array_equal(alpha, ref["alpha"]) and np.array_equal(J, ref["J"]) and np.array_equal(st, ref["status"])
    assert tmax == 2.0


class _OracleScene:
    """Stand-in for a rank's dc.Context on CPU: the slice's gradients from the oracle, accumulated as the example does."""
    def __init__(self, table):
        self.table = table

    def scene_wrench(self, shape, r, att, lam, att_kind, pdip_tol, per_pair=True, pair_range=None):
        import itertools
        from oracle import oracle as orc
        from test_gpu_operators import wrench_reference
        nb = len(shape)
        ij = np.array(list(itertools.combinations(range(nb), 2)))[pair_range[0]:pair_range[1]]
        o = orc.query(1, self.table.arrays(), shape[ij[:, 0]], shape[ij[:, 1]], r[ij[:, 0]], att[ij[:, 0]], r[ij[:, 1]], att[ij[:, 1]],
                      att_kind, pdip_tol, nthreads=2)
        return {"wrench": wrench_reference(o["grad"], lam, att, ij, att_kind), "alpha": o["alpha"], "status": o["status"], "iters": o["iters"]}


def _scene_worker(rank, world, port, nb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from test_gpu_operators import scene_bodies
    t, shape, r, att = scene_bodies(nb, 0, spread=2.0)
    lam = util.rng(11).uniform(0.0, 2.0, nb * (nb - 1) // 2)
    w, mine, (k0, k1) = dc.scene_wrench_sharded(_OracleScene(t), rank, world, shape, r, att, lam)
    alpha = dc.gather_rows(torch.from_numpy(mine["alpha"]), len(lam), rank, world)
    q.put((rank, w, alpha.numpy(), k0, k1))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_scene_wrench_allreduce(oracle):
    """SURVEY.md 8e, all-pairs scene: body arrays on every rank, contiguous slices of the i<j enumeration, one all-reduce
    of the n_bodies x 6 partial sums (the only exchange step on any path of this library)."""
    nb, world = 23, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_scene_worker, args=(r, world, port, nb, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=180) for _ in range(world)], key=lambda g: g[0])
    for p in procs:
        p.join(timeout=60)
    from test_gpu_operators import scene_bodies
    t, shape, r, att = scene_bodies(nb, 0, spread=2.0)
    n = nb * (nb - 1) // 2
    lam = util.rng(11).uniform(0.0, 2.0, n)
    ref = _OracleScene(t).scene_wrench(shape, r, att, lam, 0, 1e-6, pair_range=(0, n))
    assert (got[0][3], got[0][4], got[1][3], got[1][4]) == (0, (n + 1) // 2, (n + 1) // 2, n)
    for rank, w, alpha, _, _ in got:
        np.testing.assert_allclose(w, ref["wrench"], rtol=0, atol=1e-12 * np.abs(ref["wrench"]).max())
        np.testing.assert_array_equal(alpha, ref["alpha"])
