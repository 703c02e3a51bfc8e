"""Host-side logic that needs no GPU: C ABI surface, driver mirror, dataset/NetCDF output, sharding maths."""
import ctypes
import os
import re
from datetime import datetime

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol(drv):
    hdr = open(os.path.join(ROOT, "include", "speedy_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(spdy_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 30
    lib = ctypes.CDLL(drv.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_driver_surface_matches_reference_names(drv):
    """The f2py module exposes get_/set_/is_array_ for all 109 variables and _shape for the 102 arrays."""
    d = drv.speedy_driver
    reg = drv.REGISTRY
    assert len(reg) == 109
    arrays = [e for e in reg if e["shape"] is not None]
    assert len(arrays) == 102
    for e in reg:
        for pre in ("get_", "set_", "is_array_"):
            assert hasattr(d, pre + e["name"]), pre + e["name"]
    for e in arrays:
        assert hasattr(d, f"get_{e['name']}_shape")
        assert getattr(d, f"is_array_{e['name']}")() is True
    for n in ("modelstate_init", "modelstate_init_sst_anom", "modelstate_close", "controlparams_init",
              "controlparams_close", "create_datetime", "get_datetime", "close_datetime", "init", "step",
              "parallel_step", "check", "transform_spectral2grid", "transform_grid2spectral", "apply_grid_filter"):
        assert hasattr(d, n), n
    assert d.is_array_current_step() is False


def test_registry_header_consistent(drv):
    hdr = open(os.path.join(ROOT, "include", "spdy_registry.h")).read()
    for e in drv.REGISTRY:
        assert f"V_{e['name']} = {e['id']}," in hdr


def test_dataset_roundtrip(tmp_path):
    from pyspeedy_b200.dataset import Dataset

    rng = np.random.default_rng(0)
    t = datetime(1982, 1, 2)

    def member(k):
        dv = {"t": (["lon", "lat", "lev", "time", "ens"], rng.standard_normal((96, 48, 8, 1, 1)).astype(np.float32))}
        co = dict(lon=np.arange(96, dtype=np.float32), lat=np.arange(48, dtype=np.float32), lev=np.arange(8, dtype=np.float32),
                  time=[t], ens=[k])
        return Dataset(dv, co).reverse("lev").transpose("time", "ens", "lev", "lat", "lon")

    a, b = member(0), member(1)
    m = Dataset.merge([a, b])
    assert m["t"].shape == (1, 2, 8, 48, 96)
    assert np.array_equal(m.sel_ens(1)["t"], b["t"][:, 0])
    assert m["lev"][0] == 7.0
    p = tmp_path / "x.nc"
    m.to_netcdf(str(p))
    r = Dataset.open_dataset(str(p))
    assert r.dims("t") == ("time", "ens", "lev", "lat", "lon")
    assert np.array_equal(r["t"], m["t"]) and r.coords["time"] == [t]


def test_month_window_arithmetic():
    from pyspeedy_b200.speedy import _add_months

    assert _add_months(datetime(1982, 1, 1), -1) == datetime(1981, 12, 1)
    assert _add_months(datetime(1982, 12, 1), 1) == datetime(1983, 1, 1)
    assert _add_months(datetime(1982, 5, 1), 8) == datetime(1983, 1, 1)


def test_bench_algorithmic_bytes():
    """DESIGN.md section 5 / SURVEY 8(d): the per-unit byte counts the roofline is computed from."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.SPEC_B == 31 * 32 * 16 and b.GRID_B == 96 * 48 * 8 and b.FOUR_B == 62 * 48 * 8
    assert b.ALG_BYTES["legendre_inv"] == 77 * 39680 and b.ALG_BYTES["fft_inv"] == 77 * 60672


def test_fortran_shim_in_sync_with_registry(tmp_path):
    """integration/speedy_driver_b200.f90 (the iso_c_binding host side) is the generator's
    output for the packaged registry: one get/set/shape triple per array, every bind(C) name is declared in the
    header, free-form line length respected."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "integration", "speedy_driver_b200.f90")
    before = open(path).read()
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "gen_fortran_shim.py")], stdout=subprocess.DEVNULL)
    assert open(path).read() == before
    header = open(os.path.join(root, "include", "speedy_b200.h")).read()
    for name in set(re.findall(r'bind\(C, name="(\w+)"\)', before)):
        assert re.search(r"\b%s\(" % name, header), name
    assert max(len(line) for line in before.splitlines()) <= 132
    reg = json.load(open(os.path.join(root, "pyspeedy_b200", "data", "model_state.json")))
    for e in reg:
        assert "subroutine get_%s(" % e["name"] in before and "subroutine set_%s(" % e["name"] in before
        assert "subroutine is_array_%s(" % e["name"] in before
        if e["shape"] is not None:
            assert "subroutine get_%s_shape(" % e["name"] in before


def test_hdf5_reader_on_the_reference_boundary_file():
    """pyspeedy_b200.hdf5_reader (used by Speedy.set_bc(bc_file="*.nc") when xarray / netCDF4 are absent) on the reference's
    own NetCDF-4 file: every variable of pyspeedy/data/example_bc.nc equals the packaged .npz conversion (which marks the
    netCDF default fill value as NaN; the file's _FillValue attribute is NaN, so xarray -- and this reader -- keep the
    9.96921e36 markers, all of which lie outside the land / sea masks: test_oracle_golden.py).  Runs where the reference is
    mounted (this container); the GPU box has no /root/reference."""
    import pytest

    from pyspeedy_b200 import hdf5_reader
    from pyspeedy_b200.speedy import _load_bc

    src = "/root/reference/pyspeedy/data/example_bc.nc"
    if not os.path.isfile(src):
        pytest.skip("reference not mounted")
    d = _load_bc(src)
    ref = np.load(os.path.join(ROOT, "pyspeedy_b200", "data", "example_bc.npz"))
    assert set(ref.files) <= set(d) and {"lon", "lat", "time"} <= set(d)
    fill = np.float32(9.96921e36)
    for k in ref.files:
        a = d[k].copy()
        assert a.dtype == np.float32 and a.shape == ref[k].shape
        a[a == fill] = np.nan
        assert np.array_equal(a, ref[k], equal_nan=True), k
    assert d["lat"].shape == (48,) and abs(d["lat"][0] + 87.159) < 1e-3 and d["lon"][1] == 3.75
    f = hdf5_reader._File(src)
    att = f.attributes(f.links(f.root)["sst"])
    assert bytes(att["long_name"]).startswith(b"sea-sfc. temperature") and np.isnan(att["_FillValue"]).all()
    # the NetCDF-3 fixtures of the reference are not HDF5: a clear error, not garbage
    with pytest.raises(ValueError):
        hdf5_reader.load("/root/reference/pyspeedy/tests/fixtures/1982-01-02_0000.nc")


def test_run_loop_batching_logic(monkeypatch):
    """SpeedyEns.run without a GPU: which driver calls the loop makes for which callbacks (pyspeedy/speedy.py:547-593 calls
    every callback after every step; the stock ones act when step % interval == 0, pyspeedy/callbacks.py:48-58)."""
    from datetime import timedelta

    from pyspeedy_b200 import speedy as sp
    from pyspeedy_b200.callbacks import BaseCallback

    calls, fired = [], []

    class FakeDriver:
        @staticmethod
        def run_steps(s, c, n):
            calls.append(n)
            return np.zeros(len(s), dtype=np.int32)

        @staticmethod
        def parallel_step(s, c):
            calls.append(1)
            return np.zeros(len(s), dtype=np.int32)

        @staticmethod
        def set_datetimes(d, when):
            pass

    class Member:
        _model_date = 1
        end_date = datetime(1982, 1, 3)

    def make(step0=0):
        e = sp.SpeedyEns.__new__(sp.SpeedyEns)
        e.comm, e.n_total, e.n_members, e.members = None, 2, 2, [Member(), Member()]
        e.current_date = datetime(1982, 1, 3) - timedelta(days=2)
        e._step = step0
        e.handles = lambda: (np.array([1, 2], dtype=np.int64), np.array([1, 2], dtype=np.int64))
        e.get_current_step = lambda: e._step + sum(calls)
        return e

    class Every(BaseCallback):
        def __init__(self, interval):
            super().__init__(interval=interval)

        def __call__(self, model):
            if not self.skip_flag(model):
                fired.append((self.interval, model.get_current_step()))

    monkeypatch.setattr(sp, "_speedy", FakeDriver)
    # two stock callbacks: stop exactly where one of them acts, never more than a simulated day per call
    make().run(callbacks=[Every(24), Every(36)])
    assert calls == [24, 12, 12, 24] and fired == [(24, 24), (36, 36), (24, 48), (24, 72), (36, 72)]
    # a start that is not a multiple of the intervals
    calls.clear(), fired.clear()
    make(step0=30).run(callbacks=[Every(36)])
    assert calls == [6, 36, 30] and [s for _, s in fired] == [36, 72]
    # no callbacks: a day per call; an explicit steps_per_call is honoured; 1 = the reference's loop
    calls.clear()
    make().run()
    assert calls == [36, 36]
    calls.clear()
    make().run(callbacks=[Every(36)], steps_per_call=10)
    assert calls == [10] * 7 + [2]
    calls.clear()
    make().run(callbacks=[Every(36)], steps_per_call=1)
    assert calls == [1] * 72
    # a plain callable, or a subclass with its own skip_flag, may act at any step: one driver call per step
    calls.clear()
    make().run(callbacks=[lambda m: None])
    assert calls == [1] * 72

    class Own(Every):
        def skip_flag(self, model):
            return model.get_current_step() % 7 != 0

    calls.clear(), fired.clear()
    make().run(callbacks=[Own(36)])
    assert calls == [1] * 72 and [s for _, s in fired] == list(range(7, 73, 7))
