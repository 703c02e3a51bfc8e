"""The Speedy model classes: host-side mirror of pyspeedy/speedy.py (same names, arguments and error behaviour).

Differences, all forced by this image (no xarray / netCDF4 / HDF5) and documented in DESIGN.md:
  * boundary conditions are read from ``.npz`` files with the reference's variable names (``tools/convert_reference_data.py``
    converts the reference's example_bc.nc); NetCDF input is used when xarray is importable;
  * the default SST anomaly file is not available: a zero anomaly of the right length is used unless
    ``sst_anomaly`` is given (an ``.npz`` with ``ssta`` (lon, lat, time) and ``time`` as ``YYYY-MM`` strings, or an
    array);
  * ``to_dataframe`` returns ``xarray.Dataset`` when xarray is importable, else :class:`pyspeedy_b200.dataset.Dataset`
    (same dims/ordering/dtypes, NetCDF-3 writer).
Ensemble extensions: ``SpeedyEns.run(..., steps_per_call=n)`` advances all members ``n`` steps per driver call;
``SpeedyEns(n, comm=distributed.init())`` shards the members over one process per GPU; ``to_dataframe``, ``check`` and
``mean_and_spread`` of an ensemble are single batched driver calls.
"""
import json
import os
import warnings
from datetime import datetime, timedelta

import numpy as np

from pyspeedy_b200 import (
    _speedy,  # noqa
    example_bc_file,
    example_sst_anomaly_file,
    PACKAGE_DATA_DIR,
    DEFAULT_OUTPUT_VARS,
)
from pyspeedy_b200.dataset import Dataset
from pyspeedy_b200.error_codes import ERROR_CODES

with open(PACKAGE_DATA_DIR / "model_state.json") as fp:
    MODEL_STATE_DEF = {e["name"]: e for e in json.load(fp)}

# boundary-condition file variable -> state variable (pyspeedy/speedy.py:279-296)
_BC_VARS = (
    ("orog", "orog"), ("fmask_orig", "lsm"), ("alb0", "alb"), ("veg_high", "vegh"), ("veg_low", "vegl"),
    ("stl12", "stl"), ("snowd12", "snowd"), ("soil_wc_l1", "swl1"), ("soil_wc_l2", "swl2"), ("soil_wc_l3", "swl3"),
    ("sst12", "sst"), ("sea_ice_frac12", "icec"),
)


def _add_months(date, months):
    y, m = divmod(date.year * 12 + (date.month - 1) + months, 12)
    return date.replace(year=y, month=m + 1)


def _load_bc(bc_file):
    """Boundary-condition fields by name: ``.npz`` (converted files) or NetCDF-4 like the reference's example_bc.nc
    (pyspeedy/speedy.py:277: ``xr.load_dataset``); without xarray + netCDF4 the NetCDF-4 file is read by the package's
    own HDF5 reader, which returns the same values (``_FillValue`` decoding as xarray does it)."""
    if bc_file.endswith(".npz"):
        with np.load(bc_file) as f:
            return {k: f[k] for k in f.files}
    try:
        import xarray as xr

        ds = xr.load_dataset(bc_file, engine="netcdf4")
        return {k: ds[k].values for k in ds.data_vars}
    except ImportError:
        from pyspeedy_b200 import hdf5_reader

        return hdf5_reader.load(bc_file)


def _load_ssta_nc(path):
    """(months as 'YYYY-MM' strings, ssta (lon, lat, time)) of a NetCDF-4 SST anomaly file (variable ``ssta``, CF time
    axis ``<unit> since <date>``), read with the package's HDF5 reader."""
    from pyspeedy_b200 import hdf5_reader

    f = hdf5_reader._File(path)
    links = f.links(f.root)
    if "ssta" not in links or "time" not in links:
        raise RuntimeError(f"{path}: no 'ssta' / 'time' variable")
    units = f.attributes(links["time"]).get("units")
    if units is None:
        raise RuntimeError(f"{path}: the time axis has no (fixed-length string) 'units' attribute")
    unit, _, epoch = bytes(units).decode().strip("\0 ").partition(" since ")
    fmt = "%Y-%m-%d %H:%M:%S" if ":" in epoch else "%Y-%m-%d"
    base = datetime.strptime(epoch.strip()[:19], fmt)
    if unit not in ("days", "hours", "minutes", "seconds"):
        raise NotImplementedError(f"time unit '{unit}'")
    t = f.dataset(links["time"]).astype(np.float64)
    months = [(base + timedelta(**{unit: float(x)})).strftime("%Y-%m") for x in t]
    ssta = hdf5_reader.load(path)["ssta"]
    if ssta.shape[-1] != len(months):  # (time, lat, lon) on file -> (lon, lat, time)
        ssta = np.transpose(ssta, (2, 1, 0))
    return months, ssta


class Speedy:
    """Speedy model (one ensemble member resident on the GPU)."""

    def __init__(self, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2), member=None):
        self._start_date = None
        self._end_date = None
        self._model_date = None
        self._control_cnt = None
        self.member_id = member
        self.is_ensemble_member = self.member_id is not None
        self._state_cnt = _speedy.modelstate_init()
        self.set_params(start_date=start_date, end_date=end_date)
        self._initialized_bc = False
        self._initialized_ssta = False
        self.current_date = self.start_date

    def __del__(self):
        try:
            _speedy.modelstate_close(self._state_cnt)
            _speedy.controlparams_close(self._control_cnt)
            self._dealloc_date(self._start_date)
            self._dealloc_date(self._end_date)
            self._dealloc_date(self._model_date)
        except Exception:  # interpreter shutdown
            pass

    def set_params(self, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2)):
        self.start_date = start_date
        self.end_date = end_date
        if self.start_date > self.end_date:
            raise ValueError("The start date should be lower than the en date.")
        if self._control_cnt is not None:
            _speedy.controlparams_close(self._control_cnt)
        self._control_cnt = _speedy.controlparams_init(self._start_date, self._end_date)
        self.current_date = start_date
        self.n_months = (
            (self.end_date.year - self.start_date.year) * 12 + (self.end_date.month - self.start_date.month) + 1
        )

    @staticmethod
    def _dealloc_date(container):
        if container is not None:
            _speedy.close_datetime(container)

    def __getitem__(self, var_name):
        _getter = getattr(_speedy, f"get_{var_name}", None)
        if _getter is None:
            raise AttributeError(f"The state variable '{var_name}' does not exist.")
        time_dim = MODEL_STATE_DEF[var_name]["time_dim"]
        if time_dim:
            return _getter(self._state_cnt, getattr(self, time_dim))
        return _getter(self._state_cnt)

    def get_shape(self, var_name):
        _getter = getattr(_speedy, f"get_{var_name}_shape", None)
        if _getter is None:
            raise AttributeError(f"The 'get-shape' method for the state variable {var_name}' does not exist.")
        return tuple(int(x) for x in _getter(self._state_cnt))

    def __setitem__(self, var_name, value):
        _setter = getattr(_speedy, f"set_{var_name}", None)
        if _setter is None:
            raise AttributeError(f"The setter for the state variable '{var_name}' does not exist.")
        is_array_func = getattr(_speedy, f"is_array_{var_name}")
        if is_array_func():
            value = np.asarray(value)
            if self.get_shape(var_name) != value.shape:
                raise ValueError("Array shape missmatch")
            value = np.asfortranarray(value)
            time_dim = MODEL_STATE_DEF[var_name]["time_dim"]
            if time_dim:
                return _setter(self._state_cnt, value, getattr(self, time_dim))
            return _setter(self._state_cnt, value)
        return _setter(self._state_cnt, value)

    @staticmethod
    def _get_fortran_date(container):
        return datetime(*_speedy.get_datetime(container))

    @staticmethod
    def _set_fortran_date(container, date_value):
        if not isinstance(date_value, datetime):
            raise TypeError("The input value is not a datetime object.")
        ymdhm = (date_value.year, date_value.month, date_value.day, date_value.hour, date_value.minute)
        if container is not None:  # same container, new value (the reference frees and re-creates it)
            _speedy.set_datetime(container, *ymdhm)
            return container
        return _speedy.create_datetime(*ymdhm)

    def get_current_step(self):
        return self["current_step"]

    @property
    def start_date(self):
        return self._get_fortran_date(self._start_date)

    @start_date.setter
    def start_date(self, value):
        self._start_date = self._set_fortran_date(self._start_date, value)

    @property
    def current_date(self):
        return self._get_fortran_date(self._model_date)

    @current_date.setter
    def current_date(self, value):
        self._model_date = self._set_fortran_date(self._model_date, value)

    @property
    def end_date(self):
        return self._get_fortran_date(self._end_date)

    @end_date.setter
    def end_date(self, value):
        self._end_date = self._set_fortran_date(self._end_date, value)

    def set_bc(self, bc_file=None, sst_anomaly=None):
        """Set the boundary conditions and initialise the model (pyspeedy/speedy.py:217-301)."""
        if self._initialized_bc:
            raise RuntimeError(
                "The model was already initialized. Create a new instance if you need different boundary conditions."
            )
        self._set_sst_anomalies(sst_anomaly=sst_anomaly)
        if bc_file is None:
            bc_file = example_bc_file()
        if not os.path.isfile(bc_file):
            raise RuntimeError("The boundary conditions file does not exist.\n" f"File: {bc_file}")
        ds = _load_bc(bc_file)
        for var, key in _BC_VARS:
            self[var] = np.asarray(ds[key], dtype=np.float64)
        error_code = _speedy.init(self._state_cnt, self._control_cnt)
        if error_code < 0:
            raise RuntimeError(ERROR_CODES[error_code])
        self.spectral2grid()
        self._initialized_bc = True

    def _set_sst_anomalies(self, sst_anomaly=None):
        """Load the SST anomalies for the 3-month window around the run (pyspeedy/speedy.py:303-373)."""
        if self._initialized_ssta:
            raise RuntimeError(
                "The SST anomaly was already initialized."
                " Create a new instance if you need different boundary conditions."
            )
        start_date = _add_months(self.start_date.replace(day=1, hour=0, minute=0, microsecond=0), -1)
        end_date = _add_months(self.end_date.replace(day=1, hour=0, minute=0, microsecond=0), 1)
        expected_months = (end_date.year - start_date.year) * 12 + (end_date.month - start_date.month) + 1
        if sst_anomaly is None and os.path.isfile(example_sst_anomaly_file()):
            sst_anomaly = example_sst_anomaly_file()
        if sst_anomaly is None:
            # documented deviation: the reference raises when its packaged sst_anomaly.nc is missing
            # (pyspeedy/speedy.py:319-326); that file is not redistributable here, so the default is a ZERO anomaly
            warnings.warn(
                "pyspeedy_b200: the default SST anomaly file is not packaged; running with a zero SST anomaly. "
                "Pass sst_anomaly=<array or .npz> to reproduce a reference run that used pyspeedy/data/sst_anomaly.nc.",
                RuntimeWarning, stacklevel=3)
            ssta = np.zeros((96, 48, expected_months))
        elif isinstance(sst_anomaly, str):
            if not os.path.isfile(sst_anomaly):
                raise RuntimeError("The SST anomaly file does not exist.\n" f"File: {sst_anomaly}")
            if sst_anomaly.endswith(".npz"):
                with np.load(sst_anomaly) as f:
                    months, field = [str(m) for m in f["time"]], f["ssta"]
            else:
                months, field = _load_ssta_nc(sst_anomaly)
            wanted = [_add_months(start_date, i).strftime("%Y-%m") for i in range(expected_months)]
            missing = [w for w in wanted if w not in months]
            if missing:
                raise RuntimeError(
                    f"{len(missing)} months are missing in the SST anomalies file for the period: "
                    + start_date.strftime("%Y/%m/%d") + " , " + end_date.strftime("%Y/%m/%d") + ".\n "
                )
            ssta = np.stack([field[:, :, months.index(w)] for w in wanted], axis=-1)
        elif isinstance(sst_anomaly, np.ndarray):
            ssta = sst_anomaly
            if ssta.shape != (96, 48, expected_months):
                raise RuntimeError(f"{expected_months} months of SST anomalies are needed, got shape {ssta.shape}")
        else:
            raise TypeError(f"Unsupported sst_anomaly input: {type(sst_anomaly)}")
        _speedy.modelstate_init_sst_anom(self._state_cnt, expected_months - 2)
        self["sst_anom"] = np.asarray(ssta, dtype=np.float64)
        self._initialized_ssta = True

    def run(self, callbacks=None, steps_per_call=None):
        """Run the model between the start and end dates (pyspeedy/speedy.py:375-405).

        As in :meth:`SpeedyEns.run`: when every callback is a stock ``BaseCallback`` (acts when ``step % interval == 0``)
        the model is advanced by one multi-step driver call up to the next step at which one of them acts;
        ``steps_per_call=1`` is the reference's loop, one ``step`` driver call per model step."""
        if callbacks is None:
            callbacks = list()
        if not self._initialized_bc:
            raise RuntimeError("The SPEEDY model was not initialized. Call the `set_bc` method to initialize the model.")
        self.current_date = self.start_date
        dt_step = timedelta(seconds=3600 * 24 / 36)
        intervals = None
        if steps_per_call is None:
            from pyspeedy_b200.callbacks import BaseCallback

            if all(isinstance(cb, BaseCallback) and type(cb).skip_flag is BaseCallback.skip_flag for cb in callbacks):
                intervals = [max(1, int(cb.interval)) for cb in callbacks]
            else:
                steps_per_call = 1
        step = self.get_current_step() if intervals else 0
        end_date, date = self.end_date, self.current_date
        s, c = np.array([self._state_cnt], dtype=np.int64), np.array([self._control_cnt], dtype=np.int64)
        while date < end_date:
            left = int(round((end_date - date) / dt_step))
            n = min([36] + [i - step % i for i in intervals]) if intervals is not None else steps_per_call
            n = max(1, min(n, left))
            if n > 1:
                error_code = int(_speedy.run_steps(s, c, n)[0])
            else:
                error_code = _speedy.step(self._state_cnt, self._control_cnt)
            if error_code < 0:
                raise RuntimeError(ERROR_CODES[error_code])
            step += n
            date += n * dt_step
            self.current_date = date
            for callback in callbacks:
                callback(self)

    def grid2spectral(self):
        _speedy.transform_grid2spectral(self._state_cnt)

    def spectral2grid(self):
        _speedy.transform_spectral2grid(self._state_cnt)

    def grid_filter(self):
        _speedy.apply_grid_filter(self._state_cnt)

    def to_dataframe(self, variables=None):
        """Dataset with the current model state: dims (time[, ens], lev, lat, lon), float32, levels increasing
        with height (pyspeedy/speedy.py:415-477)."""
        if variables is None:
            variables = DEFAULT_OUTPUT_VARS
        self.spectral2grid()
        data_vars = dict()
        attrs = dict()
        for var in variables:
            dims = list(MODEL_STATE_DEF[var]["nc_dims"]) + ["time"]
            var_data = self[var][..., None].astype("float32")
            if self.is_ensemble_member:
                dims = dims + ["ens"]
                var_data = var_data[..., None]
            name = MODEL_STATE_DEF[var]["alt_name"]
            data_vars[name] = (dims, var_data)
            attrs[name] = dict(long_name=MODEL_STATE_DEF[var]["desc"], standard_name=MODEL_STATE_DEF[var]["std_name"])
            if MODEL_STATE_DEF[var]["units"] is not None:
                attrs[name]["units"] = MODEL_STATE_DEF[var]["units"]
        coords = dict(lon=self["lon"], lat=self["lat"], lev=self["lev"], time=[self.current_date])
        if self.is_ensemble_member:
            coords["ens"] = [self.member_id]
        ds = Dataset(data_vars=data_vars, coords=coords, attrs=attrs)
        sorted_dims = ("time", "ens", "lev", "lat", "lon") if self.is_ensemble_member else ("time", "lev", "lat", "lon")
        ds = ds.reverse("lev").transpose(*sorted_dims)
        return ds.to_xarray() if Dataset.HAVE_XARRAY else ds

    def check(self):
        error_code = _speedy.check(self._state_cnt)
        if error_code < 0:
            raise RuntimeError(ERROR_CODES[error_code])


class SpeedyEns:
    """Ensemble of Speedy model instances, advanced together by one driver call per step.

    ``comm`` (extension, :mod:`pyspeedy_b200.distributed`): with one process per GPU, ``num_of_members`` is the size of
    the WHOLE ensemble and this process creates only its own contiguous block (``member_id`` = global member number);
    nothing is exchanged inside a time step, ``mean_and_spread`` is reduced over all ranks."""

    def __init__(self, num_of_members, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2), comm=None):
        self.comm = comm
        self.n_total = num_of_members
        first, count = comm.shard(num_of_members) if comm is not None else (0, num_of_members)
        self.n_members = count
        self.members = [Speedy(start_date=start_date, end_date=end_date, member=first + n) for n in range(count)]
        self.current_date = self.members[0].current_date

    def __iter__(self):
        return iter(self.members)

    def __len__(self):
        return self.n_members

    def set_params(self, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2)):
        for member in self:
            member.set_params(start_date=start_date, end_date=end_date)
        self.current_date = start_date

    def to_dataframe(self, variables=None):
        """Dataset with the current state of every member: dims (time, ens, lev, lat, lon), float32, levels increasing
        with height -- what merging the members' ``to_dataframe`` gives (pyspeedy/speedy.py:538-545), built from ONE
        batched spectral2grid and ONE device gather + copy per variable (cast to float32 on the device)."""
        if variables is None:
            variables = DEFAULT_OUTPUT_VARS
        s, _ = self.handles()
        _speedy.batch_spectral2grid(s)
        data_vars, attrs = dict(), dict()
        for var in variables:
            e = MODEL_STATE_DEF[var]
            nc_dims = list(e["nc_dims"])  # e.g. (lon, lat, lev): Fortran order == reversed C order of the gather
            a = _speedy.ensemble_get(s, var, dtype=np.float32)  # (ens, lev, lat, lon)
            dims = ["ens"] + nc_dims[::-1]
            if "lev" in dims:
                a = np.flip(a, axis=dims.index("lev"))
            order = [d for d in ("time", "ens", "lev", "lat", "lon") if d in dims or d == "time"]
            a = np.transpose(a[None], [(["time"] + dims).index(d) for d in order])
            name = e["alt_name"]
            data_vars[name] = (order, a)
            attrs[name] = dict(long_name=e["desc"], standard_name=e["std_name"])
            if e["units"] is not None:
                attrs[name]["units"] = e["units"]
        m0 = self.members[0]
        coords = dict(lon=m0["lon"], lat=m0["lat"], lev=m0["lev"][::-1], time=[self.current_date],
                      ens=[m.member_id for m in self])
        ds = Dataset(data_vars=data_vars, coords=coords, attrs=attrs)
        return ds.to_xarray() if Dataset.HAVE_XARRAY else ds

    def check(self):
        """``Speedy.check`` for every member with one driver call; raises like the per-member loop of
        pyspeedy/callbacks.py:103-111 would at the first failing member."""
        s, _ = self.handles()
        codes = _speedy.batch_check(s)
        if (codes < 0).any():
            raise RuntimeError(ERROR_CODES[int(codes[codes < 0][0])])

    # ---- ensemble extensions (no reference counterpart) -----------------------------------------------------
    def handles(self):
        if getattr(self, "_handles", None) is None or len(self._handles[0]) != len(self.members):
            s = np.array([m._state_cnt for m in self], dtype=np.int64)  # noqa
            c = np.array([m._control_cnt for m in self], dtype=np.int64)  # noqa
            self._handles = (s, c)
        s, c = self._handles
        if any(int(c[i]) != m._control_cnt for i, m in ((0, self.members[0]), (-1, self.members[-1]))):  # set_params
            self._handles = None
            return self.handles()
        return s, c

    def set_bc(self, bc_file=None, sst_anomaly=None, perturb_sigma=None, seed=1234):
        """Initialise every member with the same boundary conditions: member 0 runs the full ``Speedy.set_bc`` and
        its device state is cloned into the others (identical to calling ``set_bc`` per member, which is how the
        reference does it, but one initialisation instead of N).  ``perturb_sigma`` adds the i.i.d. N(0, sigma)
        grid-point temperature perturbation of examples/Ensemble_forecast.ipynb to every member (counter-based
        generator keyed by the member's slot, so every rank of a sharded ensemble needs its own ``seed``; the
        default adds the rank)."""
        self.members[0].set_bc(bc_file=bc_file, sst_anomaly=sst_anomaly)
        s, _ = self.handles()
        if len(s) > 1:
            _speedy.clone_state(int(s[0]), s[1:])
        for m in self.members[1:]:
            m._initialized_bc = m._initialized_ssta = True
        if perturb_sigma:
            _speedy.perturb_temperature(s, seed + (self.comm.rank if self.comm is not None else 0), perturb_sigma)

    def mean_and_spread(self, variables=None):
        """Ensemble mean and spread (std, ddof=0) of grid variables over ALL members (all ranks), as
        {var: (mean, spread)}.  The six default outputs come from the fused path (sums in the epilogue of the batched
        spectral2grid, one NCCL all-reduce, one copy); other variables from per-variable device reductions."""
        if variables is None:
            variables = DEFAULT_OUTPUT_VARS
        s, _ = self.handles()
        out = {}
        fused = _speedy.ensemble_mean_spread(s, self.n_total)  # also brings the *_grid variables up to date
        for v in variables:
            if v in fused:
                out[v] = fused[v]
                continue
            from pyspeedy_b200.distributed import mean_spread_from_sums

            s1, s2 = _speedy.ensemble_sums(s, v)
            out[v] = mean_spread_from_sums(s1, s2, self.n_total)
        return out

    def run(self, callbacks=None, steps_per_call=None):
        """Run every member between the start and end dates (pyspeedy/speedy.py:547-593).

        The reference calls every callback after every step; the stock callbacks return at once unless the step counter is
        a multiple of their ``interval`` (``BaseCallback.skip_flag``, pyspeedy/callbacks.py:48-58).  With
        ``steps_per_call=None`` (default) the loop uses that: when every callback derives from ``BaseCallback`` the
        members are advanced by ONE multi-step driver call up to the next step at which some callback can act (at most a
        simulated day per call), which is observably the same sequence of callback actions; any other callable gets the
        reference's one driver call per step.  ``steps_per_call=n`` forces n steps per driver call (1 = the reference's
        loop).  Inside a multi-step call a failing member is frozen at its failing step and reported when the call returns
        (the per-step loop raises right after the failing step)."""
        if callbacks is None:
            callbacks = []
        end_date = self.members[0].end_date
        dt_step = timedelta(seconds=3600 * 24 / 36)
        state_cnts, control_cnts = self.handles()
        date_cnts = np.array([m._model_date for m in self], dtype=np.int64)  # noqa
        intervals = None
        if steps_per_call is None:
            from pyspeedy_b200.callbacks import BaseCallback

            # a subclass that replaces skip_flag may act at any step: only the stock test (step % interval) is batched over
            if all(isinstance(cb, BaseCallback) and type(cb).skip_flag is BaseCallback.skip_flag for cb in callbacks):
                intervals = [max(1, int(cb.interval)) for cb in callbacks]
            else:
                steps_per_call = 1
        step = self.get_current_step() if intervals else 0
        while self.current_date < end_date:
            left = int(round((end_date - self.current_date) / dt_step))
            if intervals is not None:
                n = min([36] + [i - step % i for i in intervals])
            else:
                n = steps_per_call
            n = max(1, min(n, left))
            if n > 1:
                error_codes = _speedy.run_steps(state_cnts, control_cnts, n)
            else:
                error_codes = _speedy.parallel_step(state_cnts, control_cnts)
            step += n
            self.current_date += n * dt_step
            if (error_codes < 0).any():
                msg = ""
                for k, code in enumerate(error_codes):
                    msg += f"Member{k}: {ERROR_CODES[code]}\n"
                raise RuntimeError(msg)
            # "update current date in all members" (pyspeedy/speedy.py:588-590): one driver call, the members' datetime
            # containers are updated in place
            _speedy.set_datetimes(date_cnts, self.current_date)
            for callback in callbacks:
                callback(self)

    def get_current_step(self):
        return self.members[0]["current_step"]
