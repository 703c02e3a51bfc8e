"""CUDA spectral transforms vs the oracle (FP64, tolerance 1e-12 relative to max|field| of EVERY field separately --
the fields of a batch are given amplitudes spread over six decades; BASELINE north_star)."""
import numpy as np
import pytest

from util import IL, IX, MX, NX, field_scales, ptr, relerr_fields as relerr, synth_spec

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _run(lib, fn, a, out_shape, *extra, dtype=np.float64):
    a = np.ascontiguousarray(a)
    out = np.zeros(out_shape, dtype=dtype)
    rc = getattr(lib, fn)(ptr(a), ptr(out), *extra, a.shape[0])
    assert rc == 0
    return out


@pytest.mark.parametrize("n", [1, 5, 32, 77])
def test_legendre_inv(oracle, drv, n):
    x = synth_spec(n, seed=1) * field_scales(n, 11)[:, None, None]
    ref = oracle.legendre_inv(x.view(np.float64).reshape(n, NX, 2 * MX))
    got = _run(drv.lib(), "spdy_batch_legendre_inv", x, (n, IL, 2 * MX))
    assert relerr(got, ref) < TOL


@pytest.mark.parametrize("kcos", [1, 2])
def test_fourier_inv(oracle, drv, kcos):
    n = 40
    four = np.random.default_rng(2).standard_normal((n, IL, 2 * MX)) * field_scales(n, 12)[:, None, None]
    ref = oracle.fourier_inv(four, kcos)
    got = _run(drv.lib(), "spdy_batch_fourier_inv", four, (n, IL, IX), kcos)
    assert relerr(got, ref) < TOL


def test_fourier_dir(oracle, drv):
    n = 40
    grid = np.random.default_rng(3).standard_normal((n, IL, IX)) * field_scales(n, 13)[:, None, None]
    ref = oracle.fourier_dir(grid)
    got = _run(drv.lib(), "spdy_batch_fourier_dir", grid, (n, IL, 2 * MX))
    assert relerr(got, ref) < TOL
    assert np.all(got[:, :, 1] == 0.0)  # Im(m=0) := 0 exactly (fourier.f90:117)


def test_legendre_dir(oracle, drv):
    n = 33
    four = np.random.default_rng(4).standard_normal((n, IL, 2 * MX)) * field_scales(n, 14)[:, None, None]
    ref = oracle.legendre_dir(four)
    got = _run(drv.lib(), "spdy_batch_legendre_dir", four, (n, NX, 2 * MX))
    assert relerr(got, ref) < TOL
    # bit-exact mask handling: entries outside the nsh2 mask and the whole row n = 32 are exactly zero
    assert np.array_equal(got == 0.0, ref == 0.0) or np.all(got[ref == 0.0] == 0.0)
    assert np.all(got[:, NX - 1, :] == 0.0)


@pytest.mark.parametrize("kcos", [1, 2])
def test_spec2grid(oracle, drv, kcos):
    n = 64
    x = synth_spec(n, seed=5) * field_scales(n, 15)[:, None, None]
    ref = oracle.spec2grid(x, kcos)
    got = _run(drv.lib(), "spdy_batch_spec2grid", x, (n, IL, IX), kcos)
    assert relerr(got, ref) < TOL


def test_grid2spec_and_roundtrip(oracle, drv):
    n = 64
    x = synth_spec(n, seed=6) * field_scales(n, 16)[:, None, None]
    g = oracle.spec2grid(x, 1)
    ref = oracle.grid2spec(g)
    got = _run(drv.lib(), "spdy_batch_grid2spec", g, (n, NX, MX), dtype=np.complex128)
    assert relerr(got, ref) < TOL
    # the reference's own round trip is NOT the identity (non-Gaussian latitudes, float-valued FFT constants,
    # SURVEY 7.1): only check it equals the oracle's round trip
    rt = _run(drv.lib(), "spdy_batch_spec2grid", got, (n, IL, IX), 1)
    assert relerr(rt, oracle.spec2grid(ref, 1)) < TOL


def test_full_size_properties(drv):
    """BASELINE config 4 size (16,384 fields): linearity of the transform pair, a size-independent property."""
    n = 16384
    a, b = synth_spec(n, seed=7), synth_spec(n, seed=8)
    lib = drv.lib()
    ga = _run(lib, "spdy_batch_spec2grid", a, (n, IL, IX), 1)
    gb = _run(lib, "spdy_batch_spec2grid", b, (n, IL, IX), 1)
    gab = _run(lib, "spdy_batch_spec2grid", a + 2.0 * b, (n, IL, IX), 1)
    assert relerr(gab, ga + 2.0 * gb) < 1e-13
    sa = _run(lib, "spdy_batch_grid2spec", ga, (n, NX, MX), dtype=np.complex128)
    # idempotence of the projection: g2s(s2g(g2s(s2g(x)))) == g2s(s2g(x)) only approximately in the reference
    # (4e-3) -- so test determinism instead: same input, same bits
    sa2 = _run(lib, "spdy_batch_grid2spec", ga, (n, NX, MX), dtype=np.complex128)
    assert np.array_equal(sa, sa2)


@pytest.mark.parametrize("mode", ["0"])
def test_unfused_transform_path(mode):
    """SPDY_FUSED=0 (separate Legendre and FFT kernels both ways, csrc/transforms.cu) against the oracle, in a fresh
    process because the switch is read when the library initialises; also 3 model steps.  The default -- the fused DMMA
    kernels k_spec2grid_mma3 / k_grid2spec_mma2 -- is what every other test in this suite runs."""
    import os
    import subprocess
    import sys

    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import synth_spec, ptr, relerr as relerr_var, relerr_fields as relerr
from oracle import oracle as O
from pyspeedy_b200 import _driver, Speedy, _speedy
from datetime import datetime
lib = _driver.lib()
n = 70
x = synth_spec(n, seed=21)
for kcos in (1, 2):
    g = np.zeros((n, 48, 96)); lib.spdy_batch_spec2grid(ptr(x), ptr(g), kcos, n)
    assert relerr(g, O.spec2grid(x, kcos)) < 1e-12
g = O.spec2grid(x, 1)
s = np.zeros((n, 32, 31), dtype=np.complex128); lib.spdy_batch_grid2spec(ptr(g), ptr(s), n)
ref = O.grid2spec(g)
assert relerr(s, ref) < 1e-12 and np.all(s[:, 31, :] == 0) and np.all(s[ref == 0] == 0)
st = O.State(n_months=1); ctl = O.Control((1982, 1, 1, 0, 0), (1982, 1, 2, 0, 0)); O.load_default_bc(st); assert st.init(ctl) == 0
m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2)); m.set_bc()
for _ in range(3):
    assert st.step(ctl) == 0 and _speedy.step(m._state_cnt, m._control_cnt) == 0
for v in ("vor", "div", "t", "ps", "tr"):
    assert relerr_var(m[v], st[v]) < 1e-11, v
print("fused ok")
'''
    env = dict(os.environ, SPDY_FUSED=mode)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "fused ok" in out.stdout, out.stdout + out.stderr
