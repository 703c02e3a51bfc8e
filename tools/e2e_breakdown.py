#!/usr/bin/env python
"""Where the end-to-end time of bench.py goes: per-step driver call vs the once-a-day output path, piece by piece.

  python tools/e2e_breakdown.py --members 4096
"""
import argparse
import ctypes as C
import os
import sys
import time
from datetime import datetime

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import CudaArray  # noqa: E402
from pyspeedy_b200 import DEFAULT_OUTPUT_VARS, SpeedyEns, _driver, _speedy  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=4096)
ap.add_argument("--steps", type=int, default=12)
a = ap.parse_args()
lib = _driver.lib()
lib.spdy_reserve(a.members)
ens = SpeedyEns(a.members, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 11))
ens.set_bc(perturb_sigma=0.01)
s, c = ens.handles()
assert (_speedy.run_steps(s, c, 3) == 0).all()


def timed(f, reps=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = f()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps, r


ms, _ = timed(lambda: _speedy.run_steps(s, c, a.steps))
print(f"run_steps            {ms / a.steps:8.3f} ms/step (device {lib.spdy_last_elapsed_ms() / a.steps:.3f})")
ms, _ = timed(lambda: _speedy.parallel_step(s, c), a.steps)
print(f"parallel_step        {ms:8.3f} ms/step (device {lib.spdy_last_elapsed_ms():.3f})")
for rep in range(2):
    ms, _ = timed(lambda: _speedy.batch_spectral2grid(s))
    print(f"batch_spectral2grid  {ms:8.3f} ms")
    for v in DEFAULT_OUTPUT_VARS:
        e = _driver.REGISTRY[_driver.VAR_ID[v]]
        dev, ne = C.c_void_p(), C.c_size_t()
        ms1, _ = timed(lambda: lib.spdy_ensemble_sums_device(_driver._ptr(s), len(s), e["id"], None, C.byref(dev), C.byref(ne)))
        t = torch.as_tensor(CudaArray(dev.value, 2 * ne.value), device="cuda")
        ms2, host = timed(lambda: t.cpu().numpy())
        print(f"  {v:10s} sums {ms1:8.3f} ms   D2H of {host.nbytes} B {ms2:8.3f} ms")
