"""Full model on the GPU vs the oracle: BASELINE config 1 (1 member, 1 day from the default boundary conditions),
instance isolation and the ensemble path.  Everything goes through the Speedy / SpeedyEns classes, i.e. through
the C ABI."""
from datetime import datetime

import numpy as np
import pytest

from util import relerr

pytestmark = pytest.mark.gpu

PROG = ["vor", "div", "t", "ps", "tr", "phi"]
DIAG = ["precnv", "precls", "cbmf", "tsr", "ssrd", "ssr", "slrd", "slr", "olr", "slru", "ustr", "vstr", "shf", "evap",
        "hfluxn", "tt_rsw", "rad_tau2", "land_temp", "sst_am", "stl_lm", "tice_om", "sst_om", "snowc", "alb_surface",
        "flux_solar_in", "flux_ozone_lower", "zenit_correction", "forog", "phis0", "fmask_land", "soilw12", "cdsea"]


def oracle_member(O, start=(1982, 1, 1, 0, 0), end=(1982, 1, 2, 0, 0)):
    st = O.State(n_months=1)
    ctl = O.Control(start, end)
    O.load_default_bc(st)
    assert st.init(ctl) == 0
    return st, ctl


def test_init_matches_oracle(oracle):
    from pyspeedy_b200 import Speedy

    st, ctl = oracle_member(oracle)
    m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    m.set_bc()
    for v in PROG + DIAG:
        a, b = m[v], st[v]
        assert a.shape == b.shape
        assert relerr(a, b) < 1e-11, (v, relerr(a, b))
    for v in ("lon", "lat", "lev"):
        assert np.array_equal(m[v], st[v]), v
    tc, qc = st.corh()
    import ctypes as C
    from pyspeedy_b200 import _driver
    t2 = np.zeros((31, 32), dtype=np.complex128, order="F")
    q2 = np.zeros_like(t2)
    _driver.lib().spdy_debug_get_corh(m._state_cnt, t2.ctypes.data_as(C.c_void_p), q2.ctypes.data_as(C.c_void_p))
    assert relerr(t2, tc) < 1e-12 and relerr(q2, qc) < 1e-11


def test_one_day_matches_oracle(oracle):
    """36 steps: per-step max relative error of the prognostics stays at rounding level (chaotic growth is ~1e-3/day
    in this model, so 1 day keeps ~1e-10), and the final float32 outputs agree with the oracle's."""
    from pyspeedy_b200 import Speedy

    st, ctl = oracle_member(oracle)
    m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    m.set_bc()
    worst = {}
    from pyspeedy_b200 import _speedy
    for step in range(36):
        assert st.step(ctl) == 0
        assert _speedy.step(m._state_cnt, m._control_cnt) == 0
        if step in (0, 1, 2, 3, 17, 35):
            for v in PROG:
                worst[(step, v)] = relerr(m[v], st[v])
    print({k: float(f"{x:.2e}") for k, x in worst.items()})
    for (step, v), x in worst.items():
        assert x < (1e-11 if step < 4 else 1e-8), (step, v, x)
    assert m["current_step"] == 36 == st["current_step"]
    assert _speedy.get_datetime(m._model_date) is not None
    for v in DIAG:
        assert relerr(m[v], st[v]) < 1e-7, v
    m.spectral2grid()
    st.spectral2grid()
    for v in ["u_grid", "v_grid", "t_grid", "q_grid", "phi_grid", "ps_grid"]:
        assert relerr(m[v].astype(np.float32), st[v].astype(np.float32)) < 1e-6, v


def test_fixture_coarse(oracle):
    """The reference's own golden file (test_speedy.py:27-50).  The default SST anomaly file is missing from the
    mount, so only coarse agreement is expected: land points far from the sea agree closely."""
    import os
    from pyspeedy_b200 import Speedy

    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "fixture_1982-01-02.npz"))
    m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    m.set_bc()
    m.run()
    ds = m.to_dataframe()
    assert np.array_equal(ds["lat"], fx["lat"]) and np.array_equal(ds["lev"], fx["lev"]) and np.array_equal(ds["lon"], fx["lon"])
    for k in ("u", "v", "t", "q", "phi", "ps"):
        a, b = ds[k], fx[k]
        assert a.shape == b.shape and a.dtype == np.float32
        rms = np.sqrt(np.mean((a - b) ** 2)) / (b.max() - b.min())
        assert rms < 2e-2, (k, rms)


def test_exceptions():
    """test_speedy.py:117-128: zero temperature -> check() raises."""
    from pyspeedy_b200 import Speedy

    m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    with pytest.raises(RuntimeError):
        m.run()  # not initialised
    m.set_bc()
    m.check()
    t = m["t"]
    t[:] = 0
    m["t"] = t
    with pytest.raises(RuntimeError):
        m.check()


def test_ensemble_and_isolation(oracle):
    """test_speedy.py:53-114: members stepped together equal a member stepped alone (bit-exactly: same kernels,
    same lane arithmetic), and untouched members of the same tile are not modified."""
    from pyspeedy_b200 import Speedy, SpeedyEns, _speedy

    single = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    single.set_bc()
    ens = SpeedyEns(3, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    for member in ens:
        member.set_bc()
    bystander = ens.members[2]
    before = {v: bystander[v] for v in PROG + ["land_temp", "precnv"]}
    s = np.array([ens.members[0]._state_cnt, ens.members[1]._state_cnt], dtype=np.int64)
    c = np.array([ens.members[0]._control_cnt, ens.members[1]._control_cnt], dtype=np.int64)
    for _ in range(4):
        assert _speedy.step(single._state_cnt, single._control_cnt) == 0
        assert (_speedy.parallel_step(s, c) == 0).all()
    for v in PROG:
        assert np.array_equal(ens.members[0][v], single[v]), v
        assert np.array_equal(ens.members[1][v], single[v]), v
    for v, a in before.items():
        assert np.array_equal(bystander[v], a), v
    # batched stepping == per-step stepping
    err = _speedy.run_steps(s, c, 5)
    assert (err == 0).all()
    for _ in range(5):
        assert _speedy.step(single._state_cnt, single._control_cnt) == 0
    for v in PROG:
        assert np.array_equal(ens.members[0][v], single[v]), v
    assert ens.members[0]["current_step"] == 9


def test_grid_spectral_services(oracle):
    from pyspeedy_b200 import Speedy

    st, ctl = oracle_member(oracle)
    m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    m.set_bc()
    st.spectral2grid()
    rng = np.random.default_rng(11)
    pert = rng.normal(0.0, 0.01, size=(96, 48, 8))
    tg = st["t_grid"] + pert
    st["t_grid"] = tg
    m["t_grid"] = m["t_grid"] + pert
    st.grid2spectral()
    m.grid2spectral()
    for v in ["vor", "div", "t", "tr", "phi", "ps"]:
        assert relerr(m[v], st[v]) < 1e-10, v
    st.grid_filter()
    m.grid_filter()
    for v in ["u_grid", "t_grid", "q_grid", "ps_grid"]:
        assert relerr(m[v], st[v]) < 1e-10, v


SURF = ["land_temp", "sst_am", "stl_lm", "tice_om", "sst_om", "sice_om", "stlcl_obs", "snowdcl_obs", "soilwcl_obs",
        "sstcl_ob", "sicecl_ob", "ticecl_ob", "snow_depth", "soil_avail_water", "sice_am", "tice_am", "ssti_om",
        "sstan_am", "hfluxn"]


def test_coupler_across_day_boundary_and_host_writes(oracle):
    """The coupler re-uses its interpolated climatology within a day (surface.cu: k_couple).  Across the day
    boundary and after a host write to a boundary field the results must still follow the reference, which
    re-interpolates on every step (speedy.f90:72, sea_model.f90:193-260)."""
    from pyspeedy_b200 import Speedy, _speedy

    st, ctl = oracle_member(oracle, end=(1982, 1, 3, 0, 0))
    m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 3))
    m.set_bc()
    for step in range(40):
        assert st.step(ctl) == 0
        assert _speedy.step(m._state_cnt, m._control_cnt) == 0
        if step in (0, 1, 34, 35, 36, 37, 39):
            for v in SURF:
                assert relerr(m[v], st[v]) < 1e-8, (step, v, relerr(m[v], st[v]))
    # host write in the middle of a day: visible to the very next step on both sides
    sst12 = st["sst12"].copy()
    sst12 += 1.5
    st["sst12"] = sst12
    m["sst12"] = sst12
    before = m["sstcl_ob"].copy()
    for step in range(2):
        assert st.step(ctl) == 0
        assert _speedy.step(m._state_cnt, m._control_cnt) == 0
        for v in SURF:
            assert relerr(m[v], st[v]) < 1e-8, ("after write", step, v, relerr(m[v], st[v]))
    assert np.abs(m["sstcl_ob"] - before).max() > 0.5
