"""N > 1 host logic on CPU (gloo, world_size 2): member sharding and the ensemble mean/spread reduction used by
bench.py / SpeedyEns across GPUs -- sum and sum of squares all-reduced, no data-path collective in the step."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def shard(m_total, world, rank):
    return m_total // world + (1 if rank < m_total % world else 0)


def _worker(rank, world, port, m_total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m_local = shard(m_total, world, rank)
    start = sum(shard(m_total, world, r) for r in range(rank))
    rng = np.random.default_rng(99)
    full = rng.normal(280.0, 3.0, size=(m_total, 500))
    mine = full[start:start + m_local]
    sums = torch.from_numpy(np.concatenate([mine.sum(0), (mine ** 2).sum(0)]))
    dist.all_reduce(sums)
    s = sums.numpy()
    mean = s[:500] / m_total
    spread = np.sqrt(np.maximum(s[500:] / m_total - mean ** 2, 0.0))
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max-over-ranks timing reduction
    if rank == 0:
        out.put((mean, spread, float(t.item()), full.mean(0), full.std(0)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_covers_all_members():
    for m, w in [(4096, 1), (4096, 2), (4096, 8), (10, 4), (3, 8)]:
        assert sum(shard(m, w, r) for r in range(w)) == m
        assert max(shard(m, w, r) for r in range(w)) - min(shard(m, w, r) for r in range(w)) <= 1


def test_mean_spread_allreduce_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    for p in procs:
        p.start()
    mean, spread, tmax, ref_mean, ref_std = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.allclose(mean, ref_mean, rtol=1e-12)
    assert np.allclose(spread, ref_std, rtol=1e-6)
    assert tmax == 2.0
