"""ORACLE (test infrastructure) -- ctypes wrapper around oracle/_build/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module.  The product package
(pyspeedy_b200) never does.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
ROOT = os.path.dirname(HERE)

with open(os.path.join(ROOT, "pyspeedy_b200", "data", "model_state.json")) as _fp:
    REGISTRY = json.load(_fp)
VAR_ID = {e["name"]: e["id"] for e in REGISTRY}
_NP = {"c16": np.complex128, "f8": np.float64, "f4": np.float32, "i4": np.int32, "b1": np.int32}

MX, NX, KX, IX, IL = 31, 32, 8, 96, 48


def build(force=False):
    if force or not os.path.isfile(LIB_PATH):
        subprocess.check_call(["make", "-C", HERE, "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_state_create.restype = C.c_void_p
        L.orc_control_create.restype = C.c_void_p
        L.orc_state_clone.restype = C.c_void_p
        L.orc_state_clone.argtypes = [C.c_void_p]
        L.orc_control_clone.restype = C.c_void_p
        L.orc_control_clone.argtypes = [C.c_void_p]
        L.orc_control_create.argtypes = [C.c_void_p, C.c_void_p]
        for name in ("orc_state_destroy", "orc_control_destroy", "orc_spectral2grid", "orc_grid2spectral",
                     "orc_grid_filter", "orc_state_init_tables", "orc_advance_date"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.orc_alloc_sst_anom.argtypes = [C.c_void_p, C.c_int]
        L.orc_init.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_check.argtypes = [C.c_void_p]
        L.orc_parallel_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_set.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_shape.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_table.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.orc_set_table_dt.argtypes = [C.c_double]
        L.orc_control_date.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_control_forcing.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_corh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_tendencies.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.orc_raw_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_implicit_terms.argtypes = [C.c_void_p] * 4
        L.orc_set_sppt.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong, C.c_ulonglong]
        L.orc_get_sppt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_apply_tendencies.argtypes = [C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 5
        L.orc_set_time_step.argtypes = [C.c_void_p, C.c_double]
        L.orc_set_forcing.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_couple.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_zonal_average_fields.argtypes = [C.c_void_p, C.c_double]
        L.orc_physics_columns.argtypes = [C.c_void_p] * 12
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def table(name, cap=1 << 20):
    buf = np.zeros(cap, dtype=np.float64)
    n = lib().orc_table(name.encode(), _ptr(buf), cap)
    if n < 0:
        raise KeyError(name)
    return buf[:n].copy()


def set_table_dt(dt):
    lib().orc_set_table_dt(float(dt))


# ---- batched stage functions: arrays are C-contiguous (n, <Fortran-order field>) ------------------------
def _f8(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def rfftb(lines):
    a = _f8(lines).copy()
    lib().orc_rfftb(_ptr(a), C.c_int(a.shape[0]))
    return a


def rfftf(lines):
    a = _f8(lines).copy()
    lib().orc_rfftf(_ptr(a), C.c_int(a.shape[0]))
    return a


def legendre_inv(spec):  # (n, 32, 62) real-packed -> (n, 48, 62)
    a = _f8(spec)
    out = np.zeros((a.shape[0], IL, 2 * MX))
    lib().orc_legendre_inv(_ptr(a), _ptr(out), C.c_int(a.shape[0]))
    return out


def legendre_dir(four):  # (n, 48, 62) -> (n, 32, 62)
    a = _f8(four)
    out = np.zeros((a.shape[0], NX, 2 * MX))
    lib().orc_legendre_dir(_ptr(a), _ptr(out), C.c_int(a.shape[0]))
    return out


def fourier_inv(four, kcos=1):  # (n, 48, 62) -> (n, 48, 96)
    a = _f8(four)
    out = np.zeros((a.shape[0], IL, IX))
    lib().orc_fourier_inv(_ptr(a), _ptr(out), C.c_int(kcos), C.c_int(a.shape[0]))
    return out


def fourier_dir(grid):  # (n, 48, 96) -> (n, 48, 62)
    a = _f8(grid)
    out = np.zeros((a.shape[0], IL, 2 * MX))
    lib().orc_fourier_dir(_ptr(a), _ptr(out), C.c_int(a.shape[0]))
    return out


def spec2grid(spec, kcos=1):  # (n, 32, 31) complex -> (n, 48, 96)
    a = np.ascontiguousarray(spec, dtype=np.complex128)
    out = np.zeros((a.shape[0], IL, IX))
    lib().orc_spec2grid(_ptr(a), _ptr(out), C.c_int(kcos), C.c_int(a.shape[0]))
    return out


def grid2spec(grid):  # (n, 48, 96) -> (n, 32, 31) complex
    a = _f8(grid)
    out = np.zeros((a.shape[0], NX, MX), dtype=np.complex128)
    lib().orc_grid2spec(_ptr(a), _ptr(out), C.c_int(a.shape[0]))
    return out


def _spec_pair(fn, a, b):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    o1, o2 = np.zeros_like(a), np.zeros_like(a)
    getattr(lib(), fn)(_ptr(a), _ptr(b), _ptr(o1), _ptr(o2), C.c_int(a.shape[0]))
    return o1, o2


def vort2vel(vor, div):
    return _spec_pair("orc_vort2vel", vor, div)


def vel2vort(u, v):
    return _spec_pair("orc_vel2vort", u, v)


def gradient(psi):
    a = np.ascontiguousarray(psi, dtype=np.complex128)
    o1, o2 = np.zeros_like(a), np.zeros_like(a)
    lib().orc_gradient(_ptr(a), _ptr(o1), _ptr(o2), C.c_int(a.shape[0]))
    return o1, o2


def laplacian(a, inverse=False):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    o = np.zeros_like(a)
    lib().orc_laplacian(_ptr(a), _ptr(o), C.c_int(int(inverse)), C.c_int(a.shape[0]))
    return o


def grid_vel2vort(ug, vg, kcos=2):
    a, b = _f8(ug), _f8(vg)
    o1 = np.zeros((a.shape[0], NX, MX), dtype=np.complex128)
    o2 = np.zeros_like(o1)
    lib().orc_grid_vel2vort(_ptr(a), _ptr(b), _ptr(o1), _ptr(o2), C.c_int(kcos), C.c_int(a.shape[0]))
    return o1, o2


class Control:
    def __init__(self, start, end):
        s = np.array(start, dtype=np.int32)
        e = np.array(end, dtype=np.int32)
        self.h = lib().orc_control_create(_ptr(s), _ptr(e))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_control_destroy(C.c_void_p(self.h))
            self.h = None

    def clone(self):
        c = Control.__new__(Control)
        c.h = lib().orc_control_clone(C.c_void_p(self.h))
        return c

    @property
    def date(self):
        out = np.zeros(5, dtype=np.int32)
        lib().orc_control_date(C.c_void_p(self.h), _ptr(out))
        return tuple(int(x) for x in out)

    @property
    def forcing(self):
        out = np.zeros(4)
        lib().orc_control_forcing(C.c_void_p(self.h), _ptr(out))
        return dict(tmonth=out[0], tyear=out[1], imont1=int(out[2]), month_idx=int(out[3]))

    def advance(self):
        lib().orc_advance_date(C.c_void_p(self.h))


class State:
    """One model instance of the oracle; arrays in/out are Fortran-ordered numpy arrays like f2py's."""

    def __init__(self, n_months=1):
        self.h = lib().orc_state_create()
        lib().orc_alloc_sst_anom(C.c_void_p(self.h), n_months)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_state_destroy(C.c_void_p(self.h))
            self.h = None

    def clone(self):
        s = State.__new__(State)
        s.h = lib().orc_state_clone(C.c_void_p(self.h))
        return s

    def shape(self, name):
        dims = np.zeros(5, dtype=np.int32)
        nd = C.c_int(0)
        lib().orc_shape(C.c_void_p(self.h), VAR_ID[name], _ptr(dims), C.byref(nd))
        return tuple(int(x) for x in dims[: nd.value])

    def __getitem__(self, name):
        e = REGISTRY[VAR_ID[name]]
        dt = _NP[e["dtype"]]
        shp = self.shape(name)
        out = np.zeros(shp, dtype=dt, order="F")
        rc = lib().orc_get(C.c_void_p(self.h), e["id"], _ptr(out), out.nbytes)
        assert rc == 0, (name, rc)
        if not shp:
            return out[()].item() if e["dtype"] != "b1" else bool(out[()])
        return out

    def __setitem__(self, name, value):
        e = REGISTRY[VAR_ID[name]]
        a = np.asarray(value, dtype=_NP[e["dtype"]])
        if a.ndim:
            a = np.asfortranarray(a)
            assert a.shape == self.shape(name), (name, a.shape, self.shape(name))
        rc = lib().orc_set(C.c_void_p(self.h), e["id"], _ptr(a), a.nbytes)
        assert rc == 0, (name, rc)

    def init_tables(self):
        lib().orc_state_init_tables(C.c_void_p(self.h))

    def init(self, ctl):
        return lib().orc_init(C.c_void_p(self.h), C.c_void_p(ctl.h))

    def step(self, ctl):
        return lib().orc_step(C.c_void_p(self.h), C.c_void_p(ctl.h))

    def check(self):
        return lib().orc_check(C.c_void_p(self.h))

    def spectral2grid(self):
        lib().orc_spectral2grid(C.c_void_p(self.h))

    def grid2spectral(self):
        lib().orc_grid2spectral(C.c_void_p(self.h))

    def grid_filter(self):
        lib().orc_grid_filter(C.c_void_p(self.h))

    def corh(self):
        t = np.zeros((MX, NX), dtype=np.complex128, order="F")
        q = np.zeros((MX, NX), dtype=np.complex128, order="F")
        lib().orc_get_corh(C.c_void_p(self.h), _ptr(t), _ptr(q))
        return t, q

    def tendencies(self, j2=2):
        outs = [np.zeros((MX, NX, KX), dtype=np.complex128, order="F") for _ in range(3)]
        psdt = np.zeros((MX, NX), dtype=np.complex128, order="F")
        trdt = np.zeros((MX, NX, KX), dtype=np.complex128, order="F")
        lib().orc_tendencies(C.c_void_p(self.h), j2, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(psdt), _ptr(trdt))
        return dict(vordt=outs[0], divdt=outs[1], tdt=outs[2], psdt=psdt, trdt=trdt)

    def implicit_terms(self, divdt, tdt, psdt):
        """implicit.f90:234-289 applied to copies of the (31,32,8) / (31,32) complex tendencies; returns the corrected ones."""
        out = [np.array(a, dtype=np.complex128, order="F", copy=True) for a in (divdt, tdt, psdt)]
        lib().orc_implicit_terms(C.c_void_p(self.h), *[_ptr(a) for a in out])
        return out

    def apply_tendencies(self, j1, dt, vordt, divdt, tdt, psdt, trdt):
        """Horizontal diffusion + leapfrog / RAW filter of time_stepping.f90:78-144 for given tendencies."""
        t = [np.array(a, dtype=np.complex128, order="F", copy=True) for a in (vordt, divdt, tdt, psdt, trdt)]
        lib().orc_apply_tendencies(C.c_void_p(self.h), j1, float(dt), *[_ptr(a) for a in t])

    def set_sppt(self, on, seed=0, member=0):
        lib().orc_set_sppt(C.c_void_p(self.h), int(on), int(seed), int(member))

    def sppt(self):
        """(spectral AR(1) pattern (31,32,8) complex, grid-point pattern of the last step (96,48,8)), Fortran order."""
        spec = np.zeros((MX, NX, KX), dtype=np.complex128, order="F")
        grid = np.zeros((IX, IL, KX), order="F")
        assert lib().orc_get_sppt(C.c_void_p(self.h), _ptr(spec), _ptr(grid)) == 0
        return spec, grid

    def raw_step(self, j1, j2, dt):
        lib().orc_raw_step(C.c_void_p(self.h), j1, j2, float(dt))

    def set_time_step(self, dt):
        lib().orc_set_time_step(C.c_void_p(self.h), float(dt))

    def set_forcing(self, ctl, imode):
        lib().orc_set_forcing(C.c_void_p(self.h), C.c_void_p(ctl.h), imode)

    def couple(self, ctl, day):
        lib().orc_couple(C.c_void_p(self.h), C.c_void_p(ctl.h), day)

    def zonal_average_fields(self, tyear):
        lib().orc_zonal_average_fields(C.c_void_p(self.h), float(tyear))

    def physics_columns(self, ug, vg, tg, qg, phig, pslg, utend, vtend, ttend, qtend):
        """All (96,48[,8]) Fortran-ordered float64; tendencies and qg are updated in place; returns index dbg."""
        arrs = [ug, vg, tg, qg, phig, pslg, utend, vtend, ttend, qtend]
        for a in arrs:
            assert a.flags["F_CONTIGUOUS"] and a.dtype == np.float64
        dbg = np.zeros((3, IL, IX), dtype=np.int32)
        lib().orc_physics_columns(C.c_void_p(self.h), *[_ptr(a) for a in arrs], _ptr(dbg))
        return dbg


BC_MAP = [("orog", "orog"), ("fmask_orig", "lsm"), ("alb0", "alb"), ("veg_high", "vegh"), ("veg_low", "vegl"),
          ("stl12", "stl"), ("snowd12", "snowd"), ("soil_wc_l1", "swl1"), ("soil_wc_l2", "swl2"),
          ("soil_wc_l3", "swl3"), ("sst12", "sst"), ("sea_ice_frac12", "icec")]  # pyspeedy/speedy.py:279-296


def load_default_bc(state):
    bc = np.load(os.path.join(ROOT, "pyspeedy_b200", "data", "example_bc.npz"))
    for var, key in BC_MAP:
        state[var] = bc[key].astype(np.float64)


def parallel_step(states, ctls, nthreads=0):
    n = len(states)
    sp = (C.c_void_p * n)(*[s.h for s in states])
    cp = (C.c_void_p * n)(*[c.h for c in ctls])
    err = np.zeros(n, dtype=np.int32)
    lib().orc_parallel_step(sp, cp, _ptr(err), n, nthreads)
    return err


def max_threads():
    return lib().orc_max_threads()
