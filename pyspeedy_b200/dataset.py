"""Minimal labelled dataset used when xarray is not importable (it is not in this image).

Implements only what the reference's output path needs (pyspeedy/speedy.py:415-477,538-545 and
pyspeedy/callbacks.py:115-255): named dimensions, transpose, reversal along a dimension, merge along
``ens`` / ``time`` and a NetCDF-3 classic writer / reader compatible with the reference's fixtures.
"""
from datetime import datetime

import numpy as np

try:  # pragma: no cover - not available in the build image
    import xarray as _xr

    _HAVE_XARRAY = True
except ImportError:
    _xr = None
    _HAVE_XARRAY = False


class Dataset:
    HAVE_XARRAY = _HAVE_XARRAY

    def __init__(self, data_vars, coords, attrs=None):
        self.data_vars = {k: (list(d), np.asarray(a)) for k, (d, a) in data_vars.items()}
        self.coords = {k: (list(v) if k in ("time", "ens") else np.asarray(v)) for k, v in coords.items()}
        self.attrs = attrs or {}

    def keys(self):
        return self.data_vars.keys()

    def __getitem__(self, name):
        if name in self.data_vars:
            return self.data_vars[name][1]
        return np.asarray(self.coords[name])

    def dims(self, name):
        return tuple(self.data_vars[name][0])

    def reverse(self, dim):
        dv = {}
        for k, (d, a) in self.data_vars.items():
            dv[k] = (d, np.flip(a, axis=d.index(dim)) if dim in d else a)
        co = dict(self.coords)
        co[dim] = np.asarray(co[dim])[::-1]
        return Dataset(dv, co, self.attrs)

    def transpose(self, *order):
        dv = {}
        for k, (d, a) in self.data_vars.items():
            o = [x for x in order if x in d]
            dv[k] = (o, np.transpose(a, [d.index(x) for x in o]))
        return Dataset(dv, self.coords, self.attrs)

    @staticmethod
    def merge(datasets):
        """Outer merge along the ``ens`` and/or ``time`` coordinates (xr.merge(..., join="outer"))."""
        datasets = list(datasets)
        if _HAVE_XARRAY and datasets and not isinstance(datasets[0], Dataset):  # pragma: no cover
            return _xr.merge(datasets, join="outer", compat="no_conflicts")
        first = datasets[0]
        out_coords = dict(first.coords)
        for dim in ("time", "ens"):
            if dim in first.coords:
                vals = []
                for ds in datasets:
                    for v in ds.coords[dim]:
                        if v not in vals:
                            vals.append(v)
                out_coords[dim] = sorted(vals)
        dv = {}
        pos = {d: {v: i for i, v in enumerate(out_coords[d])} for d in ("time", "ens") if d in out_coords}
        for name, (dims, a0) in first.data_vars.items():
            shape = [len(out_coords[d]) if d in ("time", "ens") else a0.shape[i] for i, d in enumerate(dims)]
            out = np.full(shape, np.nan, dtype=a0.dtype)
            for ds in datasets:
                if name not in ds.data_vars:
                    continue
                _, a = ds.data_vars[name]
                # one vectorised assignment per dataset: open-mesh index over the merged dims, full ranges elsewhere
                ix = [np.array([pos[d][v] for v in ds.coords[d]]) if d in pos else np.arange(a.shape[i])
                      for i, d in enumerate(dims)]
                out[np.ix_(*ix)] = a
            dv[name] = (dims, out)
        return Dataset(dv, out_coords, first.attrs)

    def sel_ens(self, member):
        k = list(self.coords["ens"]).index(member)
        dv = {n: ([x for x in d if x != "ens"], np.take(a, k, axis=d.index("ens"))) for n, (d, a) in self.data_vars.items()}
        co = {n: v for n, v in self.coords.items() if n != "ens"}
        return Dataset(dv, co, self.attrs)

    def to_xarray(self):  # pragma: no cover
        return _xr.Dataset({k: (d, a) for k, (d, a) in self.data_vars.items()}, coords=self.coords)

    # ---- NetCDF-3 classic I/O (the format of the reference's fixtures) ------------------------------------
    def to_netcdf(self, path, encoding=None):
        from scipy.io import netcdf_file

        with netcdf_file(path, "w") as f:
            t0 = self.coords["time"][0]
            sizes = {}
            for name, (dims, a) in self.data_vars.items():
                for d, n in zip(dims, a.shape):
                    sizes[d] = n
            for d, n in sizes.items():
                f.createDimension(d, n)
            for d in sizes:
                if d == "time":
                    # whole days since t0 as int32 when every output time allows it (the encoding xarray picks for the
                    # reference's daily fixtures), else minutes: sub-daily outputs (ModelCheckpoint(interval < 36))
                    # keep distinct time values
                    secs = [int((t - t0).total_seconds()) for t in self.coords["time"]]
                    unit, div = ("days", 86400) if all(x % 86400 == 0 for x in secs) else ("minutes", 60)
                    v = f.createVariable("time", "i4", ("time",))
                    v[:] = [x // div for x in secs]
                    v.units = f"{unit} since " + t0.strftime("%Y-%m-%d %H:%M:%S")
                    v.calendar = "proleptic_gregorian"
                    v.axis = "T"
                    v.standard_name = "time"
                elif d == "ens":
                    v = f.createVariable("ens", "i4", ("ens",))
                    v[:] = np.asarray(self.coords["ens"], dtype=np.int32)
                else:
                    v = f.createVariable(d, "f4", (d,))
                    v[:] = np.asarray(self.coords[d], dtype=np.float32)
            for name, (dims, a) in self.data_vars.items():
                v = f.createVariable(name, "f4", tuple(dims))
                v[:] = a.astype(np.float32)
                for k, val in self.attrs.get(name, {}).items():
                    setattr(v, k, val)

    @staticmethod
    def open_dataset(path):
        from scipy.io import netcdf_file

        with netcdf_file(path, "r", mmap=False) as f:
            coords, dv = {}, {}
            for name, v in f.variables.items():
                data = np.array(v.data, dtype=v.data.dtype.newbyteorder("="))
                if name in f.dimensions:
                    if name == "time":
                        from datetime import timedelta

                        unit, _, base = v.units.decode().partition(" since ")
                        base = datetime.strptime(base, "%Y-%m-%d %H:%M:%S")
                        coords[name] = [base + timedelta(**{unit: int(x)}) for x in data]
                    elif name == "ens":
                        coords[name] = [int(x) for x in data]
                    else:
                        coords[name] = data
                else:
                    dv[name] = (list(v.dimensions), data)
        return Dataset(dv, coords)
