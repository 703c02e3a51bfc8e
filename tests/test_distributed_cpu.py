"""N > 1 host logic of pyspeedy_b200.distributed on CPU, world size 2: the member sharding, the file rendezvous that carries
the NCCL id from rank 0 to the other ranks, and the mean / spread finalisation fed by a REAL all-reduce (gloo -- the
library's own collective is ncclAllReduce, which needs GPUs; tests/test_multigpu_gpu.py covers it on the box)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyspeedy_b200 import distributed as D


def test_shard_covers_all_members_contiguously():
    for m, w in [(4096, 1), (4096, 2), (4096, 8), (10, 4), (3, 8), (37, 2)]:
        blocks = [D.shard(m, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == m
        for (a0, ac), (b0, _) in zip(blocks, blocks[1:]):
            assert b0 == a0 + ac  # contiguous, in rank order
        counts = [c for _, c in blocks]
        assert max(counts) - min(counts) <= 1 and counts == sorted(counts, reverse=True)
    assert D.Comm(1, 2, 1).shard(4096) == (2048, 2048)
    import pytest

    with pytest.raises(ValueError):
        D.shard(8, 2, 2)


def test_mean_spread_from_sums_matches_numpy():
    x = np.random.default_rng(5).normal(280.0, 3.0, size=(37, 500))
    mean, spread = D.mean_spread_from_sums(x.sum(0), (x ** 2).sum(0), 37)
    assert np.allclose(mean, x.mean(0), rtol=1e-13) and np.allclose(spread, x.std(0), rtol=1e-7)
    shift = x[0]
    mean, spread = D.mean_spread_from_sums(x.sum(0), ((x - shift) ** 2).sum(0), 37, shift=shift)
    assert np.allclose(spread, x.std(0), rtol=1e-11)  # the shifted form loses no digits


def _worker(rank, world, port, rdv, m_total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    # 1. the rendezvous that distributed.init() uses for the 128-byte NCCL id
    blob = D.exchange(rank, world, lambda: bytes(range(128)), 128, path=rdv)
    assert blob == bytes(range(128))
    # 2. this rank's block of the ensemble, local sums, a real all-reduce, the package's finalisation
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = D.Comm(rank, world).shard(m_total)
    full = np.random.default_rng(99).normal(280.0, 3.0, size=(m_total, 500))
    mine = full[first:first + count]
    sums = torch.from_numpy(np.concatenate([mine.sum(0), (mine ** 2).sum(0)]))
    dist.all_reduce(sums)
    mean, spread = D.mean_spread_from_sums(sums.numpy()[:500], sums.numpy()[500:], m_total)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max-over-ranks timing reduction
    if rank == 0:
        out.put((mean, spread, float(t.item()), full.mean(0), full.std(0)))
    dist.barrier()
    dist.destroy_process_group()


def test_rendezvous_and_mean_spread_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    rdv = str(tmp_path / "rdv")
    open(rdv, "wb").write(b"stale")  # a left-over of a crashed run with another length must not be taken for the id
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, rdv, 37, q)) for r in range(2)]
    for p in procs:
        p.start()
    mean, spread, tmax, ref_mean, ref_std = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.allclose(mean, ref_mean, rtol=1e-12)
    assert np.allclose(spread, ref_std, rtol=1e-6)
    assert tmax == 2.0
    assert not os.path.exists(rdv) and not os.path.exists(rdv + ".1")  # rank 0 cleaned up
