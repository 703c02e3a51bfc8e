"""Edge cases of the ensemble driver that the reference's API allows (speedy_driver.f90.j2:58-79 takes any list of
state/control containers): members of one 32-member tile at different dates (year end, leap day, mid-year month
interpolation), ragged ensembles, a failing member next to healthy ones."""
from datetime import datetime

import numpy as np
import pytest

from util import relerr

pytestmark = pytest.mark.gpu

PROG = ["vor", "div", "t", "ps", "tr"]
SURF = ["land_temp", "sst_am", "stl_lm", "tice_om", "sst_om", "snow_depth", "soil_avail_water", "alb_surface",
        "flux_solar_in", "tsr", "olr", "precnv", "hfluxn"]


def _oracle_member(O, start, end):
    st = O.State(n_months=1)
    ctl = O.Control(start, end)
    O.load_default_bc(st)
    assert st.init(ctl) == 0
    return st, ctl


def test_members_of_one_tile_at_different_dates(oracle):
    """Calendar, daily forcing, short-wave phase and coupler cache are per member: three members that share a tile
    run through a year end, a leap day and a July day in one parallel_step call sequence."""
    from pyspeedy_b200 import Speedy, _speedy

    periods = [((1983, 12, 31, 0, 0), (1984, 1, 2, 0, 0)),
               ((1984, 2, 28, 0, 0), (1984, 3, 1, 0, 0)),
               ((1982, 7, 15, 0, 0), (1982, 7, 17, 0, 0))]
    ref = [_oracle_member(oracle, a, b) for a, b in periods]
    gpu = []
    for a, b in periods:
        m = Speedy(start_date=datetime(*a), end_date=datetime(*b))
        m.set_bc()
        gpu.append(m)
    # different step phases as well: advance member 1 by one step first
    assert ref[1][0].step(ref[1][1]) == 0
    assert _speedy.step(gpu[1]._state_cnt, gpu[1]._control_cnt) == 0
    s = np.array([m._state_cnt for m in gpu], dtype=np.int64)
    c = np.array([m._control_cnt for m in gpu], dtype=np.int64)
    for step in range(74):  # two day boundaries for every member
        for st, ctl in ref:
            assert st.step(ctl) == 0
        assert (_speedy.parallel_step(s, c) == 0).all()
        if step in (0, 35, 36, 73):
            for i, m in enumerate(gpu):
                for v in PROG + SURF:
                    assert relerr(m[v], ref[i][0][v]) < 1e-7, (step, i, v, relerr(m[v], ref[i][0][v]))
    want = [datetime(1984, 1, 2, 1, 20), datetime(1984, 3, 1, 2, 0), datetime(1982, 7, 17, 1, 20)]
    for m, w, (st, ctl) in zip(gpu, want, ref):
        assert m["current_step"] == st["current_step"]
        assert tuple(ctl.date) == (w.year, w.month, w.day, w.hour, w.minute)
        assert _speedy.get_model_datetime(m._state_cnt) == tuple(ctl.date)


def test_ragged_ensemble_subsets(oracle):
    """37 members = one full tile + 5 lanes of a second; stepping arbitrary subsets in arbitrary order."""
    from pyspeedy_b200 import SpeedyEns, _speedy

    ens = SpeedyEns(37, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    ens.set_bc()
    st, ctl = _oracle_member(oracle, (1982, 1, 1, 0, 0), (1982, 1, 2, 0, 0))
    hs, cs = ens.handles()
    odd = np.arange(1, 37, 2)[::-1].copy()  # reversed order, crosses the tile boundary
    assert (_speedy.parallel_step(hs[odd], cs[odd]) == 0).all()
    assert (_speedy.parallel_step(hs[odd], cs[odd]) == 0).all()
    even = np.arange(0, 37, 2)
    assert all(ens.members[i]["current_step"] == 0 for i in even)
    assert (_speedy.parallel_step(hs[even], cs[even]) == 0).all()
    assert (_speedy.parallel_step(hs[even], cs[even]) == 0).all()
    assert (_speedy.parallel_step(hs, cs) == 0).all()
    for _ in range(3):
        assert st.step(ctl) == 0
    for i in (0, 1, 31, 32, 35, 36):
        assert ens.members[i]["current_step"] == 3
        for v in PROG:
            assert relerr(ens.members[i][v], st[v]) < 1e-11, (i, v)
    # identical members stay bit-identical whatever the call pattern was
    for v in PROG:
        assert np.array_equal(ens.members[0][v], ens.members[35][v]), v
        assert np.array_equal(ens.members[0][v], ens.members[36][v]), v


def test_failing_member_is_isolated(oracle):
    """A member whose diagnostics check fails returns -2, keeps its date (speedy.f90:62-66 returns before
    advance_date) and does not disturb its tile neighbours."""
    from pyspeedy_b200 import SpeedyEns, _speedy

    ens = SpeedyEns(4, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    ens.set_bc()
    st, ctl = _oracle_member(oracle, (1982, 1, 1, 0, 0), (1982, 1, 2, 0, 0))
    bad = ens.members[2]
    bad["t"] = bad["t"] * 0.4  # ~100 K: finite everywhere, outside the 180..320 K window of check_diagnostics
    st_bad, ctl_bad = _oracle_member(oracle, (1982, 1, 1, 0, 0), (1982, 1, 2, 0, 0))
    st_bad["t"] = st_bad["t"] * 0.4
    assert st_bad.step(ctl_bad) == -2
    hs, cs = ens.handles()
    err = _speedy.parallel_step(hs, cs)
    assert [int(x) for x in err] == [0, 0, -2, 0]
    assert st.step(ctl) == 0
    for i in (0, 1, 3):
        for v in PROG:
            assert relerr(ens.members[i][v], st[v]) < 1e-11, (i, v)
    # the failed member: step counter incremented, date not advanced
    assert bad["current_step"] == 1 == st_bad["current_step"]
    assert tuple(ctl_bad.date) == (1982, 1, 1, 0, 0)
    assert _speedy.get_model_datetime(bad._state_cnt) == (1982, 1, 1, 0, 0)
    assert _speedy.get_model_datetime(ens.members[0]._state_cnt) == (1982, 1, 1, 0, 40)


def test_sst_anomaly_across_month_boundary(oracle):
    """A non-zero SST anomaly (the reference's default run reads one from sst_anomaly.nc, which is not shipped here)
    interpolated in time across a month boundary: monthly_interp (interpolation.f90:17-36), the month index carried
    by the control parameters (model_control.f90:113-163) and the coupler's per-day cache."""
    from pyspeedy_b200 import Speedy, _speedy

    rng = np.random.default_rng(11)
    ssta = np.asfortranarray(rng.normal(0.0, 1.5, size=(96, 48, 4)))  # Dec 1981 .. Mar 1982
    st = oracle.State(n_months=2)
    ctl = oracle.Control((1982, 1, 30, 0, 0), (1982, 2, 2, 0, 0))
    oracle.load_default_bc(st)
    st["sst_anom"] = ssta
    assert st.init(ctl) == 0
    m = Speedy(start_date=datetime(1982, 1, 30), end_date=datetime(1982, 2, 2))
    m.set_bc(sst_anomaly=ssta)
    checks = ["sstan_am", "sst_am", "sst_om", "tice_om", "ssti_om", "t", "ps", "vor"]
    for v in checks:
        assert relerr(m[v], st[v]) < 1e-9, ("init", v, relerr(m[v], st[v]))
    assert np.abs(m["sstan_am"]).max() > 0.5  # the anomaly is really in use
    for step in range(100):  # 1982-01-30 00:00 + 100 x 40 min = 02-01 18:40
        assert st.step(ctl) == 0
        assert _speedy.step(m._state_cnt, m._control_cnt) == 0
        if step in (0, 35, 36, 71, 72, 73, 99):
            for v in checks:
                assert relerr(m[v], st[v]) < 1e-8, (step, v, relerr(m[v], st[v]))
    assert _speedy.get_model_datetime(m._state_cnt) == tuple(ctl.date) == (1982, 2, 1, 18, 40)


def test_coefficients_outside_the_truncation(oracle):
    """step_field truncates the TENDENCY (trfilt, time_stepping.f90:177-179), not the state: coefficients with
    m + n > 30 that the host stored (or that grid2spectral left in row m + n = 31) keep going through the Robert-Asselin-
    Williams filter with a zero tendency, and row m + n = 31 still enters the inverse transforms and the n +- 1 stencils.
    The CUDA spectral step does not evaluate tendencies there, stores those rows only where the filter changed them, and
    the forward transform / vort2vel skip the rows nobody reads: all of that must be invisible."""
    from pyspeedy_b200 import Speedy, _speedy

    st, ctl = _oracle_member(oracle, (1982, 1, 1, 0, 0), (1982, 1, 2, 0, 0))
    m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    m.set_bc()
    for _ in range(2):  # leave the rest state first
        assert st.step(ctl) == 0
        assert _speedy.step(m._state_cnt, m._control_cnt) == 0
    mm, nn = np.meshgrid(np.arange(31), np.arange(32), indexing="ij")
    outside = (mm + nn) > 30
    rng = np.random.default_rng(5)
    for v, amp in (("t", 1e-3), ("vor", 1e-9), ("div", 1e-9), ("tr", 1e-7), ("ps", 1e-6)):
        x = np.array(st[v])
        assert x.shape[:2] == (31, 32)
        noise = amp * (rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape))
        x = x + noise * outside.reshape((31, 32) + (1,) * (x.ndim - 2))
        st[v] = x
        m[v] = x
    for step in range(5):
        assert st.step(ctl) == 0
        assert _speedy.step(m._state_cnt, m._control_cnt) == 0
        for v in PROG:
            a, b = np.asarray(m[v]), np.asarray(st[v])
            assert relerr(a, b) < 1e-11, (step, v, relerr(a, b))
            assert np.abs(b[outside]).max() > 0  # the stored values are still there, filtered
            assert relerr(a[outside], b[outside]) < 1e-11, (step, v, "outside", relerr(a[outside], b[outside]))
    m.spectral2grid()
    st.spectral2grid()
    for v in ("u_grid", "v_grid", "t_grid", "q_grid", "phi_grid", "ps_grid"):
        assert relerr(m[v], st[v]) < 1e-10, (v, relerr(m[v], st[v]))


def test_outside_coefficients_in_a_multistep_call():
    """A multi-step driver call scans the coefficients outside the truncation once and its spectral steps skip them for tiles
    where the time filter is the identity on all of them (both time levels bit-equal: k_scan_outer, dynamics.cu).  Same
    final state as per-step calls (which never skip) in three cases: untouched members (skipped), a member with noise
    stored in ONE time level in a tile of untouched ones (not skipped), and a stored -0 (the filter turns it into +0)."""
    from pyspeedy_b200 import SpeedyEns, _speedy

    e = SpeedyEns(6, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    e.set_bc()
    s, c = e.handles()
    assert (_speedy.parallel_step(s, c) == 0).all()
    mm, nn = np.meshgrid(np.arange(31), np.arange(32), indexing="ij")
    outside = (mm + nn) > 30
    rng = np.random.default_rng(7)
    x = np.array(e.members[0]["t"])
    assert np.array_equal(x[outside][..., 0], x[outside][..., 1])  # the model itself keeps the two levels equal there
    noisy = x.copy()
    noisy[..., 0] += 1e-3 * (rng.standard_normal(x.shape[:3]) + 1j * rng.standard_normal(x.shape[:3])) * outside[:, :, None]
    e.members[1]["t"] = noisy
    e.members[4]["t"] = noisy
    vz = np.array(e.members[2]["vor"])
    vz[30, 31, 3, 0] = complex(-0.0, 0.0)
    e.members[2]["vor"] = vz
    e.members[5]["vor"] = vz
    # members 0-2: one multi-step call; members 3-5 (same states): per-step calls
    assert (_speedy.run_steps(s[:3], c[:3], 6) == 0).all()
    for _ in range(6):
        assert (_speedy.parallel_step(s[3:], c[3:]) == 0).all()
    for k in range(3):
        for v in PROG:
            a, b = np.asarray(e.members[k][v]), np.asarray(e.members[k + 3][v])
            assert a.tobytes() == b.tobytes(), (k, v)  # bit patterns: -0 vs +0 would show
    assert np.abs(np.asarray(e.members[1]["t"])[outside]).max() > 0
    # a clean ensemble on its own takes the skipping path: compare with the clean member of the mixed tile
    f = SpeedyEns(2, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    f.set_bc()
    sf, cf = f.handles()
    assert (_speedy.parallel_step(sf, cf) == 0).all()
    assert (_speedy.run_steps(sf, cf, 6) == 0).all()
    for v in PROG:
        assert np.array_equal(np.asarray(f.members[1][v]), np.asarray(e.members[0][v])), v


def test_missing_value_markers_do_not_reach_the_model(tmp_path):
    """The reference's example_bc.nc marks missing land / sea values with the netCDF default fill 9.96921e36 (its _FillValue
    attribute is NaN, so xarray -- and pyspeedy_b200.hdf5_reader -- hand those numbers to the model), the packaged .npz marks
    them NaN.  Every marker lies outside the land / sea masks, where initialisation overwrites the fields: a member
    initialised from either file is the same model, bit for bit."""
    import os

    from pyspeedy_b200 import Speedy, _driver, _speedy, example_bc_file

    bc = dict(np.load(example_bc_file()))
    n_nan = 0
    for k, v in bc.items():
        n_nan += int(np.isnan(v).sum())
        bc[k] = np.where(np.isnan(v), np.float32(9.96921e36), v).astype(v.dtype)
    assert n_nan > 1000
    path = os.path.join(tmp_path, "bc_fill.npz")
    np.savez(path, **bc)
    a = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    a.set_bc()
    b = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    b.set_bc(bc_file=path)
    for _ in range(3):
        assert _speedy.step(a._state_cnt, a._control_cnt) == 0 and _speedy.step(b._state_cnt, b._control_cnt) == 0
    for e in _driver.REGISTRY:
        if e["shape"] is None or e["dtype"] not in ("f8", "c16") or e["name"] == "sst_anom":
            continue
        x, y = a[e["name"]], b[e["name"]]
        if e["name"] in [v for v, _ in (("stl12", 0), ("snowd12", 0), ("soil_wc_l1", 0), ("soil_wc_l2", 0), ("soil_wc_l3", 0),
                                        ("sst12", 0), ("sea_ice_frac12", 0))]:
            continue  # the raw input climatologies keep whatever marker the file had
        assert np.array_equal(x, y, equal_nan=True), e["name"]
