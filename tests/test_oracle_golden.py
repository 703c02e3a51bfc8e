"""Pin the oracle against what the reference ships: the fixture coordinates are an exact known-answer
test of geometry.f90:89-110 + initialization.f90:85-87; the fixture fields (day 1 and day 3 of the reference's own
test, pyspeedy/tests/test_speedy.py:27-50) are a coarse end-to-end check because the default SST-anomaly file
(pyspeedy/data/sst_anomaly.nc) is missing from the mount.  `test_sst_anomaly_explains_the_fixture_residual` shows that
the residual of that check has the size, the vertical/latitudinal structure and the land/Antarctica attenuation of the
model's response to a +-1 K SST anomaly -- on both fixture days, for all six variables."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def one_day(oracle):
    st = oracle.State(n_months=1)
    ctl = oracle.Control((1982, 1, 1, 0, 0), (1982, 1, 2, 0, 0))
    oracle.load_default_bc(st)
    assert st.init(ctl) == 0
    for _ in range(36):
        assert st.step(ctl) == 0
    st.spectral2grid()
    return st, ctl


def test_coordinates_bit_exact(one_day):
    st, _ = one_day
    fx = np.load(os.path.join(GOLDEN, "fixture_1982-01-02.npz"))
    assert np.array_equal(st["lat"], fx["lat"])
    assert np.array_equal(st["lon"], fx["lon"])
    assert np.array_equal(st["lev"][::-1], fx["lev"])  # exported with levels increasing with height


def test_calendar(one_day):
    _, ctl = one_day
    assert ctl.date == (1982, 1, 2, 0, 0)
    f = ctl.forcing
    assert f["imont1"] == 1 and f["month_idx"] == 1
    assert f["tmonth"] == float(np.float32(1.5) / np.float32(31.0))
    assert f["tyear"] == float(np.float32(1.5) / np.float32(365.0))


def test_fixture_fields_coarse(one_day):
    st, _ = one_day
    fx = np.load(os.path.join(GOLDEN, "fixture_1982-01-02.npz"))
    bc = np.load(os.path.join(os.path.dirname(GOLDEN), "..", "pyspeedy_b200", "data", "example_bc.npz"))
    land = bc["lsm"].T > 0.5
    names = {"u": "u_grid", "v": "v_grid", "t": "t_grid", "q": "q_grid", "phi": "phi_grid", "ps": "ps_grid"}
    for k, v in names.items():
        a = st[v].astype(np.float32)
        a = a[:, :, ::-1].transpose(2, 1, 0) if a.ndim == 3 else a.T
        ref = fx[k][0]
        rng = ref.max() - ref.min()
        rms = np.sqrt(np.mean((a - ref) ** 2)) / rng
        assert rms < 1.5e-2, (k, rms)
        if k == "t":  # away from the (unavailable) SST anomalies the agreement is much closer
            d = (a - ref)[:, land]
            assert np.sqrt(np.mean(d ** 2)) < 0.06
            assert np.sqrt(np.mean((a - ref)[:, :6, :] ** 2)) < 0.02  # Antarctica: 6 southernmost rows, all levels


NAMES = {"u": "u_grid", "v": "v_grid", "t": "t_grid", "q": "q_grid", "phi": "phi_grid", "ps": "ps_grid"}


def _export(st):
    """float32, levels increasing with height, (lev, lat, lon): the layout of the reference's exporter."""
    c = st.clone()
    c.spectral2grid()
    out = {}
    for k, v in NAMES.items():
        a = c[v].astype(np.float32)
        out[k] = a[:, :, ::-1].transpose(2, 1, 0) if a.ndim == 3 else a.T
    return out


def _run_days(oracle, sst_anom, snaps=(1, 3)):
    st = oracle.State(n_months=1)
    ctl = oracle.Control((1982, 1, 1, 0, 0), (1982, 1, 4, 0, 0))
    oracle.load_default_bc(st)
    if sst_anom is not None:
        st["sst_anom"] = np.repeat(sst_anom[:, :, None], 3, axis=2)
    assert st.init(ctl) == 0
    out = {}
    for day in range(1, max(snaps) + 1):
        for _ in range(36):
            assert st.step(ctl) == 0
        if day in snaps:
            out[day] = _export(st)
    return out


@pytest.fixture(scope="module")
def sst_runs(oracle):
    base = _run_days(oracle, None)
    anom = np.random.default_rng(7).choice([-1.0, 1.0], size=(96, 48))  # i.i.d. +-1 K at every sea point
    return base, _run_days(oracle, anom)


@pytest.mark.parametrize("day, fixture", [(1, "fixture_1982-01-02.npz"), (3, "fixture_1982-01-04.npz")])
def test_fixture_fields_coarse_both_days(sst_runs, day, fixture):
    """Both golden files of the reference's test (36 and 108 steps from the rest state), zero SST anomaly."""
    base, _ = sst_runs
    fx = np.load(os.path.join(GOLDEN, fixture))
    for k in NAMES:
        ref = fx[k][0]
        rms = np.sqrt(np.mean((base[day][k] - ref) ** 2)) / (ref.max() - ref.min())
        assert rms < 1.5e-2, (day, k, rms)


@pytest.mark.parametrize("day, fixture", [(1, "fixture_1982-01-02.npz"), (3, "fixture_1982-01-04.npz")])
def test_sst_anomaly_explains_the_fixture_residual(sst_runs, day, fixture):
    """The residual (zero-anomaly oracle minus reference fixture) against the oracle's response to a random +-1 K SST
    anomaly (sea_model.f90:218-221,276-296: the anomaly enters sst_am only).  Measured ratios residual/response:
    0.74-1.10 (day 1), 0.87-1.04 (day 3); correlation of the (level, latitude) RMS profiles 0.89-0.98."""
    base, pert = sst_runs
    fx = np.load(os.path.join(GOLDEN, fixture))
    bc = np.load(os.path.join(os.path.dirname(GOLDEN), "..", "pyspeedy_b200", "data", "example_bc.npz"))
    land = bc["lsm"].T > 0.5
    for k in NAMES:
        resp = pert[day][k] - base[day][k]
        res = base[day][k] - fx[k][0]
        ratio = np.sqrt(np.mean(res ** 2)) / np.sqrt(np.mean(resp ** 2))
        assert 0.6 < ratio < 1.6, (day, k, ratio)  # same magnitude for every variable
        ax = 2 if resp.ndim == 3 else 1  # zonal RMS -> (level, latitude) profile: largest where the response is largest
        prof_resp, prof_res = np.sqrt((resp ** 2).mean(axis=ax)).ravel(), np.sqrt((res ** 2).mean(axis=ax)).ravel()
        assert np.corrcoef(prof_resp, prof_res)[0, 1] > 0.85, (day, k)
    resp, res = pert[day]["t"] - base[day]["t"], base[day]["t"] - fx["t"][0]
    rms = lambda x: float(np.sqrt(np.mean(x ** 2)))  # noqa: E731
    # away from the sea the residual is attenuated exactly as the SST response is: land, Antarctica (6 southern rows)
    assert 0.6 < rms(res[:, land]) / rms(resp[:, land]) < 1.6
    assert 0.6 < rms(res[:, :6]) / rms(resp[:, :6]) < 1.6
    assert rms(res[:, land]) < 0.5 * rms(res[0][~land]) and rms(resp[:, land]) < 0.5 * rms(resp[0][~land])


def test_missing_values_only_where_masked():
    """every missing value of the boundary file lies where the binary mask is 0."""
    bc = np.load(os.path.join(os.path.dirname(GOLDEN), "..", "pyspeedy_b200", "data", "example_bc.npz"))
    lsm = bc["lsm"]
    bml, bms = lsm >= np.float32(0.1), (1.0 - lsm.astype(np.float64)) >= float(np.float32(0.1))
    for k in ("stl", "snowd", "swl1", "swl2"):
        assert not np.isnan(bc[k][bml]).any(), k
    for k in ("sst", "icec"):
        assert not np.isnan(bc[k][bms]).any(), k


def test_error_path(oracle):
    """test_speedy.py:117-128: zero temperature -> -2."""
    st = oracle.State(n_months=1)
    ctl = oracle.Control((1982, 1, 1, 0, 0), (1982, 1, 2, 0, 0))
    oracle.load_default_bc(st)
    assert st.step(ctl) == -1  # not initialised
    assert st.init(ctl) == 0
    assert st.check() == 0
    t = st["t"]
    t[:] = 0
    st["t"] = t
    assert st.check() == -2


def test_transform_identities(oracle):
    """Self-consistency of the restated transforms: legendre_dir row n = 32 is zero; linearity; the reference's
    round trip is NOT the identity (measured ~4e-3 in SURVEY 7.1) but must be stable."""
    from util import synth_spec

    x = synth_spec(8, seed=3)
    g = oracle.spec2grid(x, 1)
    y = oracle.grid2spec(g)
    assert np.all(y[:, 31, :] == 0)
    err = np.abs(y - x).max()
    assert 1e-6 < err < 2e-2
    g2 = oracle.spec2grid(2.0 * x, 1)
    assert np.abs(g2 - 2.0 * g).max() < 1e-12 * np.abs(g).max()
