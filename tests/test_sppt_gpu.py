"""SPPT (sppt.f90:40-146, physics.f90:233-248): the GPU kernels of csrc/sppt.cu against the oracle's
restatement (oracle/physics.cpp: gen_sppt) with the generator stream injected -- both sides draw their Gaussian noise from
the same counter-based generator keyed by (seed, member slot, call count)."""
import ctypes as C
from datetime import datetime

import numpy as np
import pytest

from util import ptr, relerr

pytestmark = pytest.mark.gpu
PROG = ("vor", "div", "t", "ps", "tr")


@pytest.fixture()
def sppt_switch():
    from pyspeedy_b200 import set_sppt

    yield set_sppt
    set_sppt(False)  # the switch is process-wide, like the reference's compile-time constant


def _gpu_sppt(lib, m):
    spec = np.zeros((31, 32, 8), dtype=np.complex128, order="F")
    grid = np.zeros((96, 48, 8), order="F")
    calls = C.c_longlong(-1)
    assert lib.spdy_debug_get_sppt(m._state_cnt, ptr(spec), ptr(grid), C.byref(calls)) == 0
    return spec, grid, calls.value


def test_sppt_matches_oracle(oracle, drv, sppt_switch):
    from pyspeedy_b200 import Speedy, _speedy

    lib = drv.lib()
    st = oracle.State(n_months=1)
    ctl = oracle.Control((1982, 1, 1, 0, 0), (1982, 1, 3, 0, 0))
    oracle.load_default_bc(st)
    assert st.init(ctl) == 0
    m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 3))
    m.set_bc()
    plain = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 3))
    plain.set_bc()
    for _ in range(4):  # spin both up a little, SPPT off
        assert st.step(ctl) == 0 and _speedy.step(m._state_cnt, m._control_cnt) == 0
        assert _speedy.step(plain._state_cnt, plain._control_cnt) == 0
    for v in PROG:
        m[v] = st[v]
    seed = 20240607
    sppt_switch(True, seed)
    st.set_sppt(1, seed=seed, member=m._state_cnt - 1)  # the GPU keys the generator by the arena slot = handle - 1
    for step in range(6):
        assert st.step(ctl) == 0 and _speedy.step(m._state_cnt, m._control_cnt) == 0
        spec, grid, calls = _gpu_sppt(lib, m)
        ospec, ogrid = st.sppt()
        assert calls == step + 1
        assert relerr(spec, ospec) < 1e-12, (step, relerr(spec, ospec))
        assert relerr(grid, ogrid) < 1e-12, (step, relerr(grid, ogrid))
        assert np.abs(grid).max() <= 1.0 and grid.std() > 0.1  # clipped to +-1, a pattern of O(0.2)
        for v in PROG:
            assert relerr(m[v], st[v]) < 1e-11, (step, v, relerr(m[v], st[v]))
    # the perturbation is real: the SPPT member has left the unperturbed trajectory ...
    sppt_switch(False)
    for _ in range(6):
        assert _speedy.step(plain._state_cnt, plain._control_cnt) == 0
    assert np.abs(m["t"] - plain["t"]).max() > 1e-4
    # ... and with the switch off again both members follow the unperturbed equations (oracle without SPPT from here)
    st.set_sppt(0)
    for v in PROG:
        m[v] = st[v]
    for _ in range(3):
        assert st.step(ctl) == 0 and _speedy.step(m._state_cnt, m._control_cnt) == 0
    for v in PROG:
        assert relerr(m[v], st[v]) < 1e-11, v


def test_sppt_members_differ_and_runs_reproduce(sppt_switch):
    from pyspeedy_b200 import SpeedyEns, _driver, _speedy

    lib = _driver.lib()

    def run(seed):
        ens = SpeedyEns(3, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
        ens.set_bc()  # identical members: any difference between them comes from their SPPT patterns
        sppt_switch(True, seed)
        s, c = ens.handles()
        assert (_speedy.run_steps(s, c, 12) == 0).all()
        out = _speedy.ensemble_get(np.sort(s), "t")
        sppt_switch(False)
        return np.sort(s), out

    s1, a = run(7)
    assert np.abs(a[0] - a[1]).max() > 1e-6 and np.abs(a[1] - a[2]).max() > 1e-6  # every member has its own pattern
    import gc

    gc.collect()
    s2, b = run(7)
    s3, c = run(8)
    assert np.array_equal(s1, s2) and np.array_equal(a, b)  # same seed, same slots: bit-identical
    assert not np.array_equal(a, c)
