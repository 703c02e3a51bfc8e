"""Two processes, two GPUs: the sharded ensemble of pyspeedy_b200.distributed (file rendezvous of the NCCL id, one
communicator inside libspeedy_b200.so, ncclAllReduce issued by the library).  Skipped on a one-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["SPDY_ROOT"])
from datetime import datetime
from pyspeedy_b200 import SpeedyEns, _speedy, distributed
from pyspeedy_b200.callbacks import EnsembleStatistics
comm = distributed.init()
ens = SpeedyEns(48 + 5, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2), comm=comm)   # ragged: 27 + 26
ens.set_bc(perturb_sigma=0.2, seed=77)
assert [m.member_id for m in ens] == list(range(*[comm.shard(53)[0], sum(comm.shard(53))]))
st = EnsembleStatistics(interval=36)
ens.run(callbacks=[st], steps_per_call=36)
s, _ = ens.handles()
out = {"mean_t": st.mean["t_grid"][0], "spread_t": st.spread["t_grid"][0], "mean_ps": st.mean["ps_grid"][0],
       "spread_ps": st.spread["ps_grid"][0], "t": _speedy.ensemble_get(s, "t_grid"), "ps": _speedy.ensemble_get(s, "ps_grid"),
       "tmax": comm.max(float(comm.rank + 1)), "first": comm.shard(53)[0]}
comm.barrier()
np.savez(os.path.join(os.environ["SPDY_OUT"], f"rank{comm.rank}.npz"), **out)
comm.destroy()
print("worker ok", comm.rank)
'''


def test_two_rank_ensemble_statistics(drv, tmp_path):
    if drv.lib().spdy_device_count() < 2:
        pytest.skip("needs two GPUs")
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT="29533",
                   SPDY_ROOT=ROOT, SPDY_OUT=str(tmp_path), SPDY_RENDEZVOUS_DIR=str(tmp_path))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, cwd=ROOT, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "worker ok" in o, o
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in range(2))
    assert r0["first"] == 0 and r1["first"] == 27 and r0["t"].shape[0] == 27 and r1["t"].shape[0] == 26
    assert r0["tmax"] == 2.0 and r1["tmax"] == 2.0
    for k in ("mean_t", "spread_t", "mean_ps", "spread_ps"):
        assert np.array_equal(r0[k], r1[k]), k  # every rank holds the statistics of the WHOLE ensemble
    t = np.concatenate([r0["t"], r1["t"]])
    ps = np.concatenate([r0["ps"], r1["ps"]])
    assert np.abs(r0["mean_t"] - t.mean(axis=0).T).max() < 1e-11 * np.abs(t).max()
    assert np.abs(r0["spread_t"] - t.std(axis=0).T).max() < 1e-6 * t.std(axis=0).max()
    assert np.abs(r0["mean_ps"] - ps.mean(axis=0).T).max() < 1e-11 * np.abs(ps).max()
    assert np.abs(r0["spread_ps"] - ps.std(axis=0).T).max() < 1e-6 * ps.std(axis=0).max()
    assert not np.array_equal(r0["t"][0], r1["t"][0])  # the two shards hold different members
