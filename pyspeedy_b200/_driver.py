"""ctypes mirror of the reference's f2py module object ``pyspeedy.speedy_driver.speedy_driver``.

The reference reaches its Fortran through ``from .speedy_driver import speedy_driver as _speedy``
(pyspeedy/__init__.py:14) and calls ``_speedy.<procedure>`` with f2py calling conventions
(registry/templates/speedy_driver.f90.j2:29-334).  This module exposes an object with the same attribute
surface -- ``modelstate_init``, ``init``, ``step``, ``parallel_step``, ``check``, ``transform_*``,
``get_<var>`` / ``set_<var>`` / ``get_<var>_shape`` / ``is_array_<var>`` for the 109 registry variables -- on
top of the C ABI of libspeedy_b200.so (include/speedy_b200.h).  There is no CPU fallback: importing works
anywhere (so host logic can be tested), but the first call that needs the model aborts without a CUDA device.
"""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPDY_LIB", os.path.join(_HERE, "csrc", "libspeedy_b200.so"))

with open(os.path.join(_HERE, "data", "model_state.json")) as _fp:
    REGISTRY = json.load(_fp)
VAR_ID = {e["name"]: e["id"] for e in REGISTRY}
_NP = {"c16": np.complex128, "f8": np.float64, "f4": np.float32, "i4": np.int32, "b1": np.int32}

_lib = None


def lib():
    """Load libspeedy_b200.so (built in-tree by ``make -C pyspeedy_b200/csrc`` / ``__graft_entry__.build``)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA library first (python -c 'import __graft_entry__ as g; "
                "g.build()').  pyspeedy_b200 has no CPU path."
            )
        # ctypes.CDLL releases the GIL during calls; every entry point of the library takes one process-wide lock
        # (engine.cu: API_LOCK), so Python threads driving different Speedy instances are safe (and serialised)
        L = C.CDLL(LIB_PATH)
        i64, vp, ci = C.c_int64, C.c_void_p, C.c_int
        L.spdy_modelstate_init.restype = i64
        L.spdy_modelstate_init_sst_anom.argtypes = [i64, ci]
        L.spdy_modelstate_close.argtypes = [i64]
        L.spdy_create_datetime.restype = i64
        L.spdy_create_datetime.argtypes = [ci] * 5
        L.spdy_get_datetime.argtypes = [i64, vp]
        L.spdy_get_model_datetime.argtypes = [i64, vp]
        L.spdy_get_model_datetime.restype = ci
        L.spdy_close_datetime.argtypes = [i64]
        L.spdy_controlparams_init.restype = i64
        L.spdy_controlparams_init.argtypes = [i64, i64]
        L.spdy_controlparams_close.argtypes = [i64]
        L.spdy_init.argtypes = [i64, i64]
        L.spdy_step.argtypes = [i64, i64]
        L.spdy_parallel_step.argtypes = [vp, vp, vp, ci]
        L.spdy_run_steps.argtypes = [vp, vp, ci, ci, vp]
        L.spdy_check.argtypes = [i64]
        for n in ("spdy_transform_spectral2grid", "spdy_transform_grid2spectral", "spdy_apply_grid_filter"):
            getattr(L, n).argtypes = [i64]
            getattr(L, n).restype = None
        L.spdy_get.argtypes = [i64, ci, vp, C.c_size_t]
        L.spdy_set.argtypes = [i64, ci, vp, C.c_size_t]
        L.spdy_shape.argtypes = [i64, ci, vp, vp]
        L.spdy_reserve.argtypes = [ci]
        L.spdy_set_sppt.argtypes = [ci, C.c_ulonglong]
        L.spdy_debug_get_sppt.argtypes = [i64, vp, vp, vp]
        L.spdy_set_device.argtypes = [ci]
        L.spdy_last_elapsed_ms.restype = C.c_float
        L.spdy_kernel_launches.restype = C.c_longlong
        L.spdy_ensemble_sums.argtypes = [vp, ci, ci, vp, vp, vp]
        L.spdy_ensemble_sums_device.argtypes = [vp, ci, ci, vp, vp, vp]
        L.spdy_ensemble_mean_spread.argtypes = [vp, ci, C.c_longlong, vp]
        L.spdy_ensemble_stats_device_ptr.restype = vp
        L.spdy_ensemble_get.argtypes = [vp, ci, ci, vp, C.c_size_t, ci]
        L.spdy_batch_check.argtypes = [vp, ci, vp]
        L.spdy_set_datetime.argtypes = [i64] + [ci] * 5
        L.spdy_set_datetimes.argtypes = [vp, ci, vp]
        L.spdy_comm_unique_id.argtypes = [vp]
        L.spdy_comm_init.argtypes = [ci, ci, vp]
        L.spdy_comm_allreduce.argtypes = [vp, ci, ci]
        L.spdy_table.argtypes = [C.c_char_p, vp, ci]
        L.spdy_batch_spec2grid.argtypes = [vp, vp, ci, ci]
        L.spdy_batch_grid2spec.argtypes = [vp, vp, ci]
        L.spdy_batch_legendre_inv.argtypes = [vp, vp, ci]
        L.spdy_batch_legendre_dir.argtypes = [vp, vp, ci]
        L.spdy_batch_fourier_inv.argtypes = [vp, vp, ci, ci]
        L.spdy_batch_fourier_dir.argtypes = [vp, vp, ci]
        L.spdy_bench_roundtrip.argtypes = [vp, vp, ci, ci, vp, vp]
        L.spdy_bench_spectral_chain.argtypes = [vp, vp, ci, ci, vp]
        L.spdy_bench_physics.argtypes = [vp, ci, vp, ci, ci, vp]
        for n in ("spdy_batch_vort2vel", "spdy_batch_vel2vort"):
            getattr(L, n).argtypes = [vp, vp, vp, vp, ci]
        L.spdy_batch_gradient.argtypes = [vp, vp, vp, ci]
        L.spdy_batch_laplacian.argtypes = [vp, vp, ci, ci]
        L.spdy_batch_grid_vel2vort.argtypes = [vp, vp, vp, vp, ci, ci]
        L.spdy_debug_tendencies_stage.argtypes = [i64, ci, ci, vp, vp, vp, vp, vp]
        L.spdy_clone_state.argtypes = [i64, vp, ci]
        L.spdy_perturb_temperature.argtypes = [vp, ci, C.c_ulonglong, C.c_double]
        L.spdy_batch_spectral2grid.argtypes = [vp, ci]
        L.spdy_profile_step.argtypes = [vp, vp, ci, vp, vp]
        L.spdy_profile_intermediate.argtypes = [ci]
        L.spdy_debug_physics.argtypes = [i64] + [vp] * 11
        L.spdy_debug_raw_step.argtypes = [i64, ci, ci, ci]
        L.spdy_debug_get_corh.argtypes = [i64, vp, vp]
        L.spdy_debug_tendencies.argtypes = [i64, ci, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class _SpeedyDriver:
    """Attribute-compatible stand-in for the f2py module object (same names, argument order and returns)."""

    # ---- model state / control containers (speedy_driver.f90.j2:131-258)
    @staticmethod
    def modelstate_init():
        return int(lib().spdy_modelstate_init())

    @staticmethod
    def modelstate_init_sst_anom(state, n_months):
        lib().spdy_modelstate_init_sst_anom(int(state), int(n_months))

    @staticmethod
    def modelstate_close(state):
        if _lib is not None:
            _lib.spdy_modelstate_close(int(state))

    @staticmethod
    def create_datetime(year, month, day, hour, minute):
        return int(lib().spdy_create_datetime(int(year), int(month), int(day), int(hour), int(minute)))

    @staticmethod
    def get_datetime(container):
        out = np.zeros(5, dtype=np.int32)
        lib().spdy_get_datetime(int(container), _ptr(out))
        return tuple(int(x) for x in out)

    @staticmethod
    def get_model_datetime(state_cnt):
        """Extension: the member's date on the device calendar."""
        out = np.zeros(5, dtype=np.int32)
        if lib().spdy_get_model_datetime(int(state_cnt), _ptr(out)) != 0:
            raise ValueError("unknown state container")
        return tuple(int(x) for x in out)

    @staticmethod
    def set_datetime(container, year, month, day, hour, minute):
        """Extension: update a datetime container in place (the reference frees and re-creates it)."""
        if lib().spdy_set_datetime(int(container), int(year), int(month), int(day), int(hour), int(minute)) != 0:
            raise ValueError("unknown datetime container")

    @staticmethod
    def set_datetimes(containers, date_value):
        """Extension: store one date in every listed container with a single driver call."""
        c = np.ascontiguousarray(containers, dtype=np.int64)
        d = np.array([date_value.year, date_value.month, date_value.day, date_value.hour, date_value.minute], dtype=np.int32)
        if lib().spdy_set_datetimes(_ptr(c), c.shape[0], _ptr(d)) != 0:
            raise ValueError("unknown datetime container")

    @staticmethod
    def close_datetime(container):
        if _lib is not None:
            _lib.spdy_close_datetime(int(container))

    @staticmethod
    def controlparams_init(start_datetime_cnt, end_datetime_cnt):
        return int(lib().spdy_controlparams_init(int(start_datetime_cnt), int(end_datetime_cnt)))

    @staticmethod
    def controlparams_close(control):
        if _lib is not None:
            _lib.spdy_controlparams_close(int(control))

    # ---- model (speedy_driver.f90.j2:29-125)
    @staticmethod
    def init(state, control):
        return int(lib().spdy_init(int(state), int(control)))

    @staticmethod
    def step(state, control):
        return int(lib().spdy_step(int(state), int(control)))

    @staticmethod
    def parallel_step(state_containers, control_containers):
        s = np.ascontiguousarray(state_containers, dtype=np.int64)
        c = np.ascontiguousarray(control_containers, dtype=np.int64)
        err = np.zeros(s.shape[0], dtype=np.int32)
        lib().spdy_parallel_step(_ptr(s), _ptr(c), _ptr(err), s.shape[0])
        return err

    @staticmethod
    def run_steps(state_containers, control_containers, nsteps):
        """Ensemble extension: ``nsteps`` parallel steps without a host round trip per step."""
        s = np.ascontiguousarray(state_containers, dtype=np.int64)
        c = np.ascontiguousarray(control_containers, dtype=np.int64)
        err = np.zeros(s.shape[0], dtype=np.int32)
        lib().spdy_run_steps(_ptr(s), _ptr(c), s.shape[0], int(nsteps), _ptr(err))
        return err

    @staticmethod
    def set_sppt(on, seed=0):
        """SPPT switch (the reference's compile-time ``sppt_on``, params.f90:44); applies to every member stepped from
        now on.  Pattern generator keyed by (seed, member slot, patterns generated so far)."""
        return int(lib().spdy_set_sppt(int(bool(on)), int(seed)))

    # ---- ensemble extensions (no reference counterpart)
    @staticmethod
    def clone_state(src_state, dst_states):
        d = np.ascontiguousarray(dst_states, dtype=np.int64)
        return int(lib().spdy_clone_state(int(src_state), _ptr(d), d.shape[0]))

    @staticmethod
    def perturb_temperature(states, seed, sigma):
        s = np.ascontiguousarray(states, dtype=np.int64)
        return int(lib().spdy_perturb_temperature(_ptr(s), s.shape[0], int(seed), float(sigma)))

    @staticmethod
    def batch_spectral2grid(states):
        s = np.ascontiguousarray(states, dtype=np.int64)
        return int(lib().spdy_batch_spectral2grid(_ptr(s), s.shape[0]))

    @staticmethod
    def ensemble_sums(states, var_name, shift=None):
        """Per-element sum and sum of squared deviations from ``shift`` over the listed members (device reduction)."""
        s = np.ascontiguousarray(states, dtype=np.int64)
        e = REGISTRY[VAR_ID[var_name]]
        n = int(np.prod(e["shape"])) * (2 if e["dtype"] == "c16" else 1)
        out1, out2 = np.zeros(n), np.zeros(n)
        sh = None if shift is None else np.ascontiguousarray(np.asarray(shift, dtype=np.float64).ravel(order="F"))
        rc = lib().spdy_ensemble_sums(_ptr(s), s.shape[0], e["id"], None if sh is None else _ptr(sh), _ptr(out1), _ptr(out2))
        if rc != 0:
            raise RuntimeError(f"spdy_ensemble_sums({var_name}) failed: {rc}")
        shp = tuple(e["shape"])
        return out1.reshape(shp, order="F"), out2.reshape(shp, order="F")

    # the six default outputs in the order of spdy_ensemble_mean_spread
    STATS_VARS = ("u_grid", "v_grid", "t_grid", "q_grid", "phi_grid", "ps_grid")

    @staticmethod
    def ensemble_mean_spread(states, n_total=None):
        """spectral2grid of the listed members + ensemble mean and spread (std, ddof=0) of the six default outputs over
        all ``n_total`` members of all ranks (default: the listed members): partial sums in the transform's epilogue, one
        NCCL all-reduce when a communicator is up, one device-to-host copy.  Returns {var: (mean, spread)}."""
        s = np.ascontiguousarray(states, dtype=np.int64)
        ntot = 5 * 96 * 48 * 8 + 96 * 48
        out = np.empty(2 * ntot)
        rc = lib().spdy_ensemble_mean_spread(_ptr(s), s.shape[0], int(n_total or s.shape[0]), _ptr(out))
        if rc != 0:
            raise RuntimeError(f"spdy_ensemble_mean_spread failed: {rc}")
        res, o = {}, 0
        for v in _SpeedyDriver.STATS_VARS:
            shp = tuple(REGISTRY[VAR_ID[v]]["shape"])
            n = int(np.prod(shp))
            res[v] = (out[o:o + n].reshape(shp, order="F"), out[ntot + o:ntot + o + n].reshape(shp, order="F"))
            o += n
        return res

    @staticmethod
    def ensemble_get(states, var_name, dtype=np.float64):
        """``get_<var>`` of every listed member in one driver call: array of shape (n_members,) + reversed(var shape),
        C-ordered, i.e. ``out[i].T`` is what ``get_<var>(states[i])`` returns.  ``dtype=np.float32`` casts on the device."""
        s = np.ascontiguousarray(states, dtype=np.int64)
        e = REGISTRY[VAR_ID[var_name]]
        if e["shape"] is None or e["dtype"] not in ("f8", "c16") or var_name == "sst_anom":
            raise ValueError(f"ensemble_get supports the float64 / complex128 array variables, not {var_name}")
        f32 = np.dtype(dtype) == np.float32
        shp = tuple(reversed(e["shape"]))
        if e["dtype"] == "c16":
            if f32:
                raise ValueError("complex variables are returned as complex128")
            out = np.empty((s.shape[0],) + shp, dtype=np.complex128)
        else:
            out = np.empty((s.shape[0],) + shp, dtype=np.float32 if f32 else np.float64)
        rc = lib().spdy_ensemble_get(_ptr(s), s.shape[0], e["id"], _ptr(out), out.nbytes, 1 if f32 else 0)
        if rc != 0:
            raise RuntimeError(f"spdy_ensemble_get({var_name}) failed: {rc}")
        return out

    @staticmethod
    def batch_check(states):
        """``check`` of every listed member in one driver call; int32 error codes."""
        s = np.ascontiguousarray(states, dtype=np.int64)
        err = np.zeros(s.shape[0], dtype=np.int32)
        lib().spdy_batch_check(_ptr(s), s.shape[0], _ptr(err))
        return err

    @staticmethod
    def profile_step(state_containers, control_containers, intermediate=False):
        """One instrumented step: ms per kernel class.  ``intermediate``: timed the way a multi-step call runs its
        intermediate steps (column-physics outputs that nothing reads before the next step are not stored)."""
        lib().spdy_profile_intermediate(1 if intermediate else 0)
        s = np.ascontiguousarray(state_containers, dtype=np.int64)
        c = np.ascontiguousarray(control_containers, dtype=np.int64)
        ms = np.zeros(10, dtype=np.float32)
        err = np.zeros(s.shape[0], dtype=np.int32)
        lib().spdy_profile_step(_ptr(s), _ptr(c), s.shape[0], _ptr(ms), _ptr(err))
        lib().spdy_profile_intermediate(0)
        names = ["forcing", "pre_ops", "legendre_inv", "fft_inv", "grid_dyn", "physics", "fft_fwd", "legendre_dir",
                 "spec_step", "post"]
        return dict(zip(names, (float(x) for x in ms))), err

    @staticmethod
    def check(state):
        return int(lib().spdy_check(int(state)))

    @staticmethod
    def transform_spectral2grid(state):
        lib().spdy_transform_spectral2grid(int(state))

    @staticmethod
    def transform_grid2spectral(state):
        lib().spdy_transform_grid2spectral(int(state))

    @staticmethod
    def apply_grid_filter(state):
        lib().spdy_apply_grid_filter(int(state))


def _shape(state, vid):
    dims = np.zeros(5, dtype=np.int32)
    nd = C.c_int(0)
    lib().spdy_shape(int(state), vid, _ptr(dims), C.byref(nd))
    return tuple(int(x) for x in dims[: nd.value])


def _make_accessors(entry):
    vid, dt, is_arr = entry["id"], _NP[entry["dtype"]], entry["shape"] is not None
    kind = entry["dtype"]

    def getter(state, n_months=None):
        shp = _shape(state, vid)
        out = np.zeros(shp, dtype=dt, order="F")
        rc = lib().spdy_get(int(state), vid, _ptr(out), out.nbytes)
        if rc != 0:
            raise RuntimeError(f"spdy_get({entry['name']}) failed with code {rc}")
        if not is_arr:
            return bool(out[()]) if kind == "b1" else out[()].item()
        return out

    def setter(state, value, n_months=None):
        a = np.asarray(value, dtype=dt)
        if a.ndim:
            a = np.asfortranarray(a)
        rc = lib().spdy_set(int(state), vid, _ptr(a), a.nbytes)
        if rc != 0:
            raise RuntimeError(f"spdy_set({entry['name']}) failed with code {rc} (shape {a.shape})")

    def shape(state):
        return np.array(_shape(state, vid), dtype=np.int32)

    def is_array():
        return is_arr

    return getter, setter, shape, is_array


for _e in REGISTRY:
    _g, _s, _sh, _ia = _make_accessors(_e)
    setattr(_SpeedyDriver, f"get_{_e['name']}", staticmethod(_g))
    setattr(_SpeedyDriver, f"set_{_e['name']}", staticmethod(_s))
    setattr(_SpeedyDriver, f"is_array_{_e['name']}", staticmethod(_ia))
    if _e["shape"] is not None:
        setattr(_SpeedyDriver, f"get_{_e['name']}_shape", staticmethod(_sh))

speedy_driver = _SpeedyDriver()
