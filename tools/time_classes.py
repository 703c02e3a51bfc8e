"""Per-kernel-class milliseconds of one model step (CUDA events inside the library) for a cloned ensemble.
Usage: python tools/time_classes.py [members] [intermediate=0|1]   (honours SPDY_LIB / SPDY_FUSED)
intermediate=1: the step as a multi-step driver call runs its intermediate steps."""
import sys
from datetime import datetime

sys.path.insert(0, ".")
from pyspeedy_b200 import SpeedyEns, _speedy  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
inter = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
ens = SpeedyEns(n, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
ens.set_bc()
s, c = ens.handles()
_speedy.profile_step(s, c, inter)
acc = None
for _ in range(3):
    p, _err = _speedy.profile_step(s, c, inter)
    acc = p if acc is None else {k: acc[k] + p[k] for k in p}
print({k: round(v / 3, 4) for k, v in acc.items()})
