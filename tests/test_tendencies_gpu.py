"""Per-tendency parity (BASELINE north_star: 1e-12 relative): the five spectral tendencies returned by
get_tendencies (tendencies.f90:11-39: grid-point dynamics + physics + spectral terms + implicit correction) of a
spun-up state, GPU vs oracle (the config-2 acceptance criterion lives in test_config2_gpu.py)."""
import ctypes as C
from datetime import datetime

import numpy as np
import pytest

from util import ptr, relerr

pytestmark = pytest.mark.gpu


def _pair(oracle, steps):
    from pyspeedy_b200 import Speedy, _speedy

    st = oracle.State(n_months=1)
    ctl = oracle.Control((1982, 1, 1, 0, 0), (1982, 1, 2, 0, 0))
    oracle.load_default_bc(st)
    assert st.init(ctl) == 0
    m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    m.set_bc()
    for _ in range(steps):
        assert st.step(ctl) == 0
        assert _speedy.step(m._state_cnt, m._control_cnt) == 0
    return st, ctl, m


@pytest.mark.parametrize("steps", [0, 7])
def test_tendencies(oracle, drv, steps):
    st, ctl, m = _pair(oracle, steps)
    # identical inputs: copy the oracle's prognostics into the GPU member (removes the ~1e-14 drift of earlier steps)
    for v in ("vor", "div", "t", "ps", "tr"):
        m[v] = st[v]
    sw = (steps % 3 == 0)
    st["compute_shortwave"] = int(sw)
    m["compute_shortwave"] = int(sw)
    ref = st.tendencies(j2=2)
    outs = {k: np.zeros(v.shape, dtype=np.complex128, order="F") for k, v in ref.items()}
    rc = drv.lib().spdy_debug_tendencies(m._state_cnt, 2, ptr(outs["vordt"]), ptr(outs["divdt"]), ptr(outs["tdt"]),
                                         ptr(outs["psdt"]), ptr(outs["trdt"]))
    assert rc == 0
    errs = {k: relerr(outs[k], ref[k]) for k in ref}
    print(errs)
    for k, e in errs.items():
        assert e < 1e-12, (k, e)
    # bit-exact mask handling: the tendencies are exactly zero where the oracle's are (rows outside the nsh2 mask)
    for k in ref:
        assert np.all(outs[k][ref[k] == 0] == 0), k
