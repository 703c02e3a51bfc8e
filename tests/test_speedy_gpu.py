"""The reference's own test-suite (pyspeedy/tests/test_speedy.py:28-150) re-stated against this package.

Same scenarios, same classes and callbacks; the comparison data set is the oracle run of the same period (the
reference's golden files are compared coarsely in test_model_gpu.py::test_fixture_coarse because the default SST
anomaly file is missing from the mount).  Tolerance: the reference test's rtol = 1e-6 on the float32 output, plus an
absolute floor of 1e-6 x max|field| for zero crossings."""
import os
import tempfile
from datetime import datetime, timedelta

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

start_dates = (
    # twice the same period: library-level globals must not carry over (test_speedy.py:11-17)
    (datetime(1982, 1, 1), datetime(1982, 1, 2)),
    (datetime(1982, 1, 1), datetime(1982, 1, 2)),
    (datetime(1982, 1, 1), datetime(1982, 1, 4)),
)
export_variables = (["u_grid", "v_grid"], ["t_grid", "q_grid"], ["phi_grid", "ps_grid"], ["precnv", "precls"])

_ORACLE_CACHE = {}


def oracle_dataset(O, end_date):
    """Output variables of an oracle run 1982-01-01 -> end_date in the exporter's layout (lev reversed, float32)."""
    from pyspeedy_b200 import DEFAULT_OUTPUT_VARS
    from pyspeedy_b200.speedy import MODEL_STATE_DEF

    if end_date in _ORACLE_CACHE:
        return _ORACLE_CACHE[end_date]
    ndays = (end_date - datetime(1982, 1, 1)).days
    st = O.State(n_months=1)
    ctl = O.Control((1982, 1, 1, 0, 0), (end_date.year, end_date.month, end_date.day, 0, 0))
    O.load_default_bc(st)
    assert st.init(ctl) == 0
    for _ in range(36 * ndays):
        assert st.step(ctl) == 0
    st.spectral2grid()
    out = {}
    for var in DEFAULT_OUTPUT_VARS:
        a = np.asarray(st[var]).astype(np.float32)  # (lon, lat[, lev])
        a = a.transpose(2, 1, 0)[::-1] if a.ndim == 3 else a.T
        out[MODEL_STATE_DEF[var]["alt_name"]] = a
    _ORACLE_CACHE[end_date] = out
    return out


def assert_close(ds, ref, member=None):
    assert set(ds.keys()) == set(ref.keys())
    for name, b in ref.items():
        a = np.asarray(ds[name])
        a = a[0] if member is None else a[0, member]  # time[, ens]
        assert a.shape == b.shape, (name, a.shape, b.shape)
        tol = 1e-6 * np.abs(b) + 1e-6 * np.abs(b).max()
        assert np.all(np.abs(a.astype(np.float64) - b) <= tol), (name, float(np.abs(a - b).max()), float(np.abs(b).max()))


def open_ds(path):
    from pyspeedy_b200.dataset import Dataset

    return Dataset.open_dataset(path)


@pytest.mark.parametrize("start_date, end_date", start_dates)
def test_speedy_run(oracle, start_date, end_date):
    from pyspeedy_b200 import Speedy
    from pyspeedy_b200.callbacks import XarrayExporter

    file_name = end_date.strftime("%Y-%m-%d_%H%M.nc")
    with tempfile.TemporaryDirectory() as tmp:
        model = Speedy(start_date=start_date, end_date=end_date)
        model.set_bc()
        model.run(callbacks=[XarrayExporter(output_dir=tmp)])
        assert_close(open_ds(os.path.join(tmp, file_name)), oracle_dataset(oracle, end_date))


def test_speedy_concurrent(oracle):
    """Two instances advanced alternately, one day at a time (test_speedy.py:53-88)."""
    from pyspeedy_b200 import Speedy
    from pyspeedy_b200.callbacks import XarrayExporter

    start_date, end_date, ndays = datetime(1982, 1, 1), datetime(1982, 1, 4), 3
    file_name = end_date.strftime("%Y-%m-%d_%H%M.nc")
    with tempfile.TemporaryDirectory() as tmp:
        d1, d2 = os.path.join(tmp, "run1"), os.path.join(tmp, "run2")
        model = Speedy(start_date=start_date, end_date=end_date)
        model.set_bc()
        model2 = Speedy(start_date=start_date, end_date=end_date)
        model2.set_bc()
        for day in range(ndays):
            for mdl, d in ((model, d1), (model2, d2)):
                mdl.start_date = start_date + timedelta(days=day)
                mdl.end_date = start_date + timedelta(days=day + 1)
                mdl.run(callbacks=[XarrayExporter(output_dir=d)])
        ref = oracle_dataset(oracle, end_date)
        assert_close(open_ds(os.path.join(d1, file_name)), ref)
        assert_close(open_ds(os.path.join(d2, file_name)), ref)


def test_ens_speedy(oracle):
    """SpeedyEns with identical members: every member equals the single-member run (test_speedy.py:91-114)."""
    from pyspeedy_b200 import SpeedyEns
    from pyspeedy_b200.callbacks import XarrayExporter

    n, start_date, end_date = 3, datetime(1982, 1, 1), datetime(1982, 1, 2)
    file_name = end_date.strftime("%Y-%m-%d_%H%M.nc")
    ref = oracle_dataset(oracle, end_date)
    ens = SpeedyEns(n, start_date=start_date, end_date=end_date)
    for member in ens:
        member.set_bc()
    with tempfile.TemporaryDirectory() as tmp:
        ens.run(callbacks=[XarrayExporter(output_dir=tmp)])
        ds = open_ds(os.path.join(tmp, file_name))
        for m, member in enumerate(ens):
            assert_close(ds, ref, member=m)
            one = member.to_dataframe()
            assert_close(one, ref, member=0)


@pytest.mark.parametrize("variables", export_variables)
def test_speedy_variable_export(variables):
    from pyspeedy_b200 import Speedy
    from pyspeedy_b200.callbacks import XarrayExporter

    start_date, end_date = datetime(1982, 1, 1), datetime(1982, 1, 2)
    file_name = end_date.strftime("%Y-%m-%d_%H%M.nc")
    with tempfile.TemporaryDirectory() as tmp:
        model = Speedy(start_date=start_date, end_date=end_date)
        model.set_bc()
        model.run(callbacks=[XarrayExporter(output_dir=tmp, variables=variables)])
        ds = open_ds(os.path.join(tmp, file_name))
        assert set(v.replace("_grid", "") for v in variables) == set(ds.keys())
