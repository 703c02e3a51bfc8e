#!/usr/bin/env python
"""Small driver for ncu: set up an ensemble, then run a few model steps inside a cudaProfilerStart/Stop bracket.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
      python tools/profile_step.py --members 512 --steps 3
"""
import argparse
import os
import sys
from datetime import datetime

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyspeedy_b200 import SpeedyEns, _driver, _speedy  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=512)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
ens = SpeedyEns(a.members, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 11))
ens.set_bc(perturb_sigma=0.01)
s, c = ens.handles()
assert (_speedy.run_steps(s, c, a.warmup) == 0).all()
lib = _driver.lib()
lib.spdy_profiler_start()
err = _speedy.run_steps(s, c, a.steps)
lib.spdy_profiler_stop()
assert (err == 0).all()
print("ok", a.members, a.steps, lib.spdy_last_elapsed_ms())
