"""One process per GPU: member sharding and the NCCL communicator of libspeedy_b200.so.

The reference's ensemble driver treats members as independent units (``parallel_step``,
registry/templates/speedy_driver.f90.j2:58-79); the only cross-member arithmetic in the whole product is the ensemble
mean / spread of the outputs (examples/Ensemble_forecast.ipynb cells 12, 16).  So an ensemble of ``n_total`` members
runs as ``world`` processes, rank ``r`` owning the contiguous block :func:`shard` gives it, with **no communication
inside a time step**; ``SpeedyEns.mean_and_spread`` / ``callbacks.EnsembleStatistics`` add one ``ncclAllReduce`` per
output time, issued by the library itself on its own stream (csrc/ensemble.cu).  No torch anywhere.

Launch exactly like any ``torchrun`` program (``python -m torch.distributed.run --nproc-per-node N prog.py``): the
launcher only provides ``RANK`` / ``WORLD_SIZE`` / ``LOCAL_RANK`` / ``MASTER_PORT``.  The 128-byte NCCL id made by
rank 0 reaches the other ranks of the node through a file rendezvous (:func:`exchange`), keyed by the launcher's
process id and master port.

    from pyspeedy_b200 import SpeedyEns, distributed
    comm = distributed.init()                       # no-op communicator when WORLD_SIZE is 1 / unset
    ens = SpeedyEns(4096, comm=comm)                # this rank creates only its shard: comm.count members
    ens.set_bc(perturb_sigma=0.01)
    ens.run(callbacks=[EnsembleStatistics()])       # mean / spread over all 4096 members on every rank
"""
import os
import sys
import tempfile
import time

import numpy as np


def shard(n_total, rank, world):
    """Contiguous block of members owned by ``rank``: (first member id, number of members); the first
    ``n_total % world`` ranks hold one member more (blocks of ceil(N/G) member slots per GPU)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(int(n_total), int(world))
    return rank * base + min(rank, extra), base + (1 if rank < extra else 0)


def mean_spread_from_sums(s1, s2, n_total, shift=None):
    """Ensemble mean and spread (std, ddof=0) from the all-reduced sums ``s1 = sum x`` and ``s2 = sum (x - shift)^2``
    (``shift`` = 0 if omitted); the host-side counterpart of ``k_mean_spread``."""
    n = float(n_total)
    mean = np.asarray(s1) / n
    d = mean if shift is None else mean - np.asarray(shift)
    return mean, np.sqrt(np.maximum(np.asarray(s2) / n - d * d, 0.0))


def _rendezvous_path(tag):
    key = "%s_%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "none"), os.getppid())
    return os.path.join(os.environ.get("SPDY_RENDEZVOUS_DIR", tempfile.gettempdir()), f"spdy_b200_{tag}_{key}")


def exchange(rank, world, make_blob, nbytes, tag="ncclid", timeout=300.0, path=None):
    """Rank 0 calls ``make_blob()`` (``nbytes`` bytes) and publishes it; every rank returns the same bytes.  Single node:
    a file written atomically next to the launcher's other temporaries; the name carries the launcher's pid, the master
    port and the run id, so concurrent jobs do not meet.  Each reader acknowledges with ``<file>.<rank>``; rank 0 removes
    everything once all have read."""
    path = path or _rendezvous_path(tag)
    if rank == 0:
        for r in range(world):
            for p in ([path] if r == 0 else [f"{path}.{r}"]):
                if os.path.exists(p):
                    os.remove(p)
        blob = bytes(make_blob())
        assert len(blob) == nbytes
        tmp = f"{path}.tmp{os.getpid()}"
        with open(tmp, "wb") as fp:
            fp.write(blob)
        os.replace(tmp, path)
        t0 = time.time()
        while not all(os.path.exists(f"{path}.{r}") for r in range(1, world)):
            if time.time() - t0 > timeout:
                raise TimeoutError(f"rendezvous {path}: not every rank picked the id up within {timeout} s")
            time.sleep(0.005)
        for r in range(1, world):
            os.remove(f"{path}.{r}")
        os.remove(path)
        return blob
    t0 = time.time()
    while True:
        try:
            with open(path, "rb") as fp:
                blob = fp.read()
            if len(blob) == nbytes:
                break
        except FileNotFoundError:
            pass
        if time.time() - t0 > timeout:
            raise TimeoutError(f"rendezvous {path}: rank 0 never published the id")
        time.sleep(0.005)
    with open(f"{path}.{rank}", "wb") as fp:
        fp.write(b"ok")
    return blob


class Comm:
    """The communicator of this process: rank / world, this rank's member block, and the few collectives the host side
    needs (barrier, max / sum of a small vector).  With ``world == 1`` everything is local and NCCL is never loaded."""

    def __init__(self, rank=0, world=1, local_rank=0):
        self.rank, self.world, self.local_rank = int(rank), int(world), int(local_rank)

    def shard(self, n_total):
        return shard(n_total, self.rank, self.world)

    def barrier(self):
        from pyspeedy_b200 import _driver

        _driver.lib().spdy_comm_barrier()

    def allreduce(self, values, op="sum"):
        """In-place semantics on a copy: element-wise sum or max over ranks of up to 64 doubles."""
        from pyspeedy_b200 import _driver

        a = np.ascontiguousarray(np.atleast_1d(values), dtype=np.float64).copy()
        if self.world > 1:
            rc = _driver.lib().spdy_comm_allreduce(_driver._ptr(a), a.shape[0], {"sum": 0, "max": 1}[op])
            if rc != 0:
                raise RuntimeError(f"spdy_comm_allreduce failed: {rc}")
        return a

    def max(self, x):
        return float(self.allreduce([x], "max")[0])

    def destroy(self):
        from pyspeedy_b200 import _driver

        if self.world > 1:
            _driver.lib().spdy_comm_destroy()


def init(rank=None, world=None, local_rank=None):
    """Select this rank's GPU and bring the NCCL communicator of the library up (once, before the first model state is
    created).  Arguments default to the launcher's environment (``RANK``, ``WORLD_SIZE``, ``LOCAL_RANK``)."""
    import ctypes as C

    from pyspeedy_b200 import _driver

    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank))) if local_rank is None else local_rank
    lib = _driver.lib()
    lib.spdy_set_device(local_rank)
    if world > 1:
        def make():
            buf = C.create_string_buffer(128)
            if lib.spdy_comm_unique_id(buf) != 0:
                raise RuntimeError("NCCL is not available (libnccl.so.2)")
            return buf.raw

        blob = exchange(rank, world, make, 128)
        # NCCL announces itself ("NCCL version ...") on STDOUT when the first communicator comes up; programs whose
        # stdout is a protocol (bench.py prints one JSON line) must not see that: route fd 1 to stderr meanwhile
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            rc = lib.spdy_comm_init(rank, world, C.create_string_buffer(blob, 128))
            if rc == 0:
                lib.spdy_comm_barrier()  # first collective: connection set-up happens here, not inside a timed region
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        if rc != 0:
            raise RuntimeError(f"spdy_comm_init failed: {rc}")
    return Comm(rank, world, local_rank)
