"""Operator-level parity on a SPUN-UP state: the spectral operators of spectral.f90:140-296, the
semi-implicit correction (implicit.f90:234-289) and horizontal diffusion + time integration (horizontal_diffusion.f90:131-152,
time_stepping.f90:78-188), each isolated -- GPU (through the C ABI, the kernels of the model step) against the oracle on
the same inputs: the prognostic fields of an oracle run after one simulated day (winds of several m/s, unlike the rest state
of step 0).  Tolerance 1e-12 of max|field| per field."""
import ctypes as C
from datetime import datetime

import numpy as np
import pytest

from util import ptr, relerr, relerr_fields

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def day1(oracle):
    st = oracle.State(n_months=1)
    ctl = oracle.Control((1982, 1, 1, 0, 0), (1982, 1, 3, 0, 0))
    oracle.load_default_bc(st)
    assert st.init(ctl) == 0
    for _ in range(36):
        assert st.step(ctl) == 0
    return st, ctl


def fields(a):
    """(31, 32, 8[, 2]) Fortran complex -> (n, 32, 31) C-ordered batch of fields (all levels, time level 2 then 1)."""
    a = np.asarray(a)
    if a.ndim == 4:
        a = np.concatenate([a[..., 1], a[..., 0]], axis=2)
    return np.ascontiguousarray(a.transpose(2, 1, 0))


def run2(lib, fn, a, b, *extra):
    o1, o2 = np.zeros_like(a), np.zeros_like(a)
    assert getattr(lib, fn)(ptr(a), ptr(b), ptr(o1), ptr(o2), *extra, a.shape[0]) == 0
    return o1, o2


def test_vort2vel_vel2vort_gradient_laplacian(oracle, drv, day1):
    st, _ = day1
    lib = drv.lib()
    vor, div = fields(st["vor"]), fields(st["div"])
    ps = np.concatenate([fields(st["ps"][:, :, None, :]), fields(st["t"])[:3]])
    assert np.abs(vor).max() > 1e-7  # spun up (vorticity in 1/s): the rest state has vor = 0
    u, v = run2(lib, "spdy_batch_vort2vel", vor, div)
    ru, rv = oracle.vort2vel(vor, div)
    assert relerr_fields(u, ru) < TOL and relerr_fields(v, rv) < TOL
    vo, dv = run2(lib, "spdy_batch_vel2vort", ru, rv)
    rvo, rdv = oracle.vel2vort(ru, rv)
    assert relerr_fields(vo, rvo) < TOL and relerr_fields(dv, rdv) < TOL
    dx, dy = np.zeros_like(ps), np.zeros_like(ps)
    assert lib.spdy_batch_gradient(ptr(ps), ptr(dx), ptr(dy), ps.shape[0]) == 0
    rdx, rdy = oracle.gradient(ps)
    assert relerr_fields(dx, rdx) < TOL and relerr_fields(dy, rdy) < TOL
    for inverse in (0, 1):
        out = np.zeros_like(vor)
        assert lib.spdy_batch_laplacian(ptr(vor), ptr(out), inverse, vor.shape[0]) == 0
        assert relerr_fields(out, oracle.laplacian(vor, inverse=bool(inverse))) < 1e-15


@pytest.mark.parametrize("kcos", [2, 1])
def test_grid_vel2vort(oracle, drv, day1, kcos):
    """spectral.f90:218-248 on the day-1 winds: u, v on the grid -> (vor, div), both cos-latitude conventions."""
    st, _ = day1
    c = st.clone()
    c.spectral2grid()
    ug = np.ascontiguousarray(c["u_grid"].transpose(2, 1, 0))  # (8, 48, 96)
    vg = np.ascontiguousarray(c["v_grid"].transpose(2, 1, 0))
    assert np.abs(ug).max() > 3.0
    n = ug.shape[0]
    vo, dv = np.zeros((n, 32, 31), dtype=np.complex128), np.zeros((n, 32, 31), dtype=np.complex128)
    assert drv.lib().spdy_batch_grid_vel2vort(ptr(ug), ptr(vg), ptr(vo), ptr(dv), kcos, n) == 0
    rvo, rdv = oracle.grid_vel2vort(ug, vg, kcos)
    assert relerr_fields(vo, rvo) < TOL and relerr_fields(dv, rdv) < TOL
    assert np.all(vo[rvo == 0] == 0)


def _gpu_member_like(st, steps=36):
    from pyspeedy_b200 import Speedy, _speedy

    m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 3))
    m.set_bc()
    for _ in range(steps):
        assert _speedy.step(m._state_cnt, m._control_cnt) == 0
    for v in ("vor", "div", "t", "ps", "tr"):  # identical inputs: the oracle's prognostics
        m[v] = st[v]
    return m


def _tend(lib, m, stage):
    shp3, shp2 = (31, 32, 8), (31, 32)
    o = {k: np.zeros(shp2 if k == "psdt" else shp3, dtype=np.complex128, order="F") for k in ("vordt", "divdt", "tdt", "psdt", "trdt")}
    assert lib.spdy_debug_tendencies_stage(m._state_cnt, 2, stage, ptr(o["vordt"]), ptr(o["divdt"]), ptr(o["tdt"]),
                                           ptr(o["psdt"]), ptr(o["trdt"])) == 0
    return o


def test_implicit_terms(oracle, drv, day1):
    """implicit.f90:234-289 isolated: the tendencies that ENTER the semi-implicit correction inside the spectral-step
    kernel (stage-1 dump) are given to the oracle's implicit_terms; the result must equal the kernel's own stage-2 output."""
    st, _ = day1
    m = _gpu_member_like(st)
    lib = drv.lib()
    drv.speedy_driver.set_compute_shortwave(m._state_cnt, 1)
    pre = _tend(lib, m, 1)
    post = _tend(lib, m, 2)
    assert not np.array_equal(pre["divdt"], post["divdt"])
    c = st.clone()
    c.set_time_step(2 * 2400.0)
    div, t, ps = c.implicit_terms(pre["divdt"], pre["tdt"], pre["psdt"])
    for got, ref, name in ((post["divdt"], div, "divdt"), (post["tdt"], t, "tdt"), (post["psdt"], ps, "psdt")):
        assert relerr(got, ref) < TOL, (name, relerr(got, ref))  # all (31,32) coefficients, also outside the truncation
    # and the complete tendencies against the oracle's get_tendencies on the same state
    c2 = st.clone()
    c2["compute_shortwave"] = 1
    ref2 = c2.tendencies(j2=2)
    for k in ("vordt", "trdt", "divdt", "tdt", "psdt"):
        assert relerr(post[k], ref2[k]) < TOL, k


def test_horizontal_diffusion_and_time_integration(oracle, drv, day1):
    """horizontal_diffusion.f90:131-152 + time_stepping.f90:78-188 isolated: the GPU's own tendencies are fed to the
    oracle's diffusion / leapfrog / RAW code; the GPU's raw step from the same state must land on the same prognostics."""
    st, _ = day1
    m = _gpu_member_like(st)
    lib = drv.lib()
    drv.speedy_driver.set_compute_shortwave(m._state_cnt, 1)
    tend = _tend(lib, m, 2)  # leaves the prognostics untouched
    c = st.clone()
    c.set_time_step(2 * 2400.0)
    c.apply_tendencies(2, 2 * 2400.0, tend["vordt"], tend["divdt"], tend["tdt"], tend["psdt"], tend["trdt"])
    assert lib.spdy_debug_raw_step(m._state_cnt, 2, 2, 2) == 0
    for v in ("vor", "div", "t", "ps", "tr"):
        a, b = m[v], c[v]
        assert relerr(a, b) < 1e-13, (v, relerr(a, b))
        assert not np.array_equal(b, st[v])
