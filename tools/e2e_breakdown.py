#!/usr/bin/env python
"""Where the end-to-end time of bench.py goes: per-step driver call vs the once-a-day output path, piece by piece.

  python tools/e2e_breakdown.py --members 4096
"""
import argparse
import os
import sys
import time
from datetime import datetime, timedelta

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyspeedy_b200 import SpeedyEns, _driver, _speedy  # noqa: E402
from pyspeedy_b200.callbacks import DiagnosticCheck, EnsembleStatistics  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=4096)
ap.add_argument("--steps", type=int, default=36)
a = ap.parse_args()
lib = _driver.lib()
lib.spdy_reserve(a.members)
end = datetime(1982, 1, 11)
ens = SpeedyEns(a.members, start_date=datetime(1982, 1, 1), end_date=end)
ens.set_bc(perturb_sigma=0.01)
s, c = ens.handles()
assert (_speedy.run_steps(s, c, 3) == 0).all()


def timed(f, reps=1):
    lib.spdy_synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = f()
    lib.spdy_synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps, r


ms, _ = timed(lambda: _speedy.run_steps(s, c, a.steps))
print(f"run_steps                     {ms / a.steps:8.3f} ms/step (device {lib.spdy_last_elapsed_ms() / a.steps:.3f})")
ms, _ = timed(lambda: _speedy.parallel_step(s, c), a.steps)
print(f"parallel_step                 {ms:8.3f} ms/step (device {lib.spdy_last_elapsed_ms():.3f})")
import numpy as np  # noqa: E402

us = np.zeros(4)
lib.spdy_last_call_host_us(_driver._ptr(us))
print("  host phases of the last parallel_step call (us): prologue %.0f, launches %.0f, wait %.0f, epilogue %.0f" % tuple(us))
ens.mean_and_spread(), ens.check()  # first-call allocations of the output path
for label, cbs in (("ens.run, no callbacks", []), ("ens.run + DiagnosticCheck + EnsembleStatistics", [DiagnosticCheck(36), EnsembleStatistics(36)]),
                   ("ens.run + DiagnosticCheck + EnsembleStatistics", [DiagnosticCheck(36), EnsembleStatistics(36)])):
    ens.current_date = end - a.steps * timedelta(seconds=2400)
    ms, _ = timed(lambda: ens.run(callbacks=cbs))
    print(f"{label:48s} {ms / a.steps:8.3f} ms/step")
for rep in range(2):
    ms, _ = timed(lambda: _speedy.batch_spectral2grid(s))
    print(f"batch_spectral2grid           {ms:8.3f} ms")
    ms, _ = timed(lambda: _speedy.ensemble_mean_spread(s))
    print(f"ensemble_mean_spread (fused)  {ms:8.3f} ms   (spectral2grid + sums + D2H of 3 MB)")
    ms, _ = timed(lambda: _speedy.batch_check(s))
    print(f"batch_check                   {ms:8.3f} ms")
    ms, r = timed(lambda: _speedy.ensemble_get(s[:64], "t_grid", dtype="float32"))
    print(f"ensemble_get(64 x t_grid f32) {ms:8.3f} ms   ({r.nbytes} B)")
