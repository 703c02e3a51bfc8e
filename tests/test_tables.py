"""Host table generator of the product vs the oracle's tables: bit-exact (both evaluate the reference's mixed
REAL(4)/REAL(8) expressions with glibc).  No GPU needed: spdy_table is host-only."""
import ctypes as C

import numpy as np
import pytest

NAMES = ["hsg", "dhs", "fsg", "dhsr", "fsgr", "radang", "coriol", "sia", "coa", "cosgr", "cosgr2", "sigl", "sigh",
         "grdsig", "grdscp", "wvi", "wt", "cpol", "el2", "elm2", "trfilt", "gradx", "gradym", "gradyp", "uvdx",
         "uvdym", "uvdyp", "vddym", "vddyp", "dmp", "dmpd", "dmps", "tcorv", "qcorv", "tref", "tref2", "tref3",
         "xgeop1", "xgeop2", "fband"]
IMPL = ["dmp1", "dmp1d", "dmp1s", "elz", "xc", "xd", "xj", "dhsx"]


def product_table(drv, name, cap=1 << 20):
    buf = np.zeros(cap)
    n = drv.lib().spdy_table(name.encode(), buf.ctypes.data_as(C.c_void_p), cap)
    assert n > 0, name
    return buf[:n].copy()


@pytest.mark.parametrize("name", NAMES)
def test_table_bit_exact(oracle, drv, name):
    a, b = oracle.table(name), product_table(drv, name)
    assert a.shape == b.shape
    assert np.array_equal(a, b), (name, np.abs(a - b).max())


def test_fft_twiddles(oracle, drv):
    a, b = oracle.table("wa"), product_table(drv, "wa")
    assert np.array_equal(a[:93], b[:93])


@pytest.mark.parametrize("kind,dt", [(0, 1200.0), (1, 2400.0), (2, 4800.0)])
def test_implicit_tables_bit_exact(oracle, drv, kind, dt):
    oracle.set_table_dt(dt)
    try:
        for name in IMPL:
            a, b = oracle.table(name), product_table(drv, f"{name}@{kind}")
            assert np.array_equal(a, b), (name, kind, np.abs(a - b).max())
    finally:
        oracle.set_table_dt(4800.0)


def test_oracle_self_consistency(oracle):
    """nsh2 = [62,62,60,...,4,2], ifac = [96,4,2,4,4,3], sum(wt) = 1."""
    nsh2 = oracle.table("nsh2")
    assert nsh2.tolist() == [2 * min(31, 33 - n) for n in range(1, 33)] and nsh2.sum() == 1054
    assert oracle.table("ifac").tolist() == [96, 4, 2, 4, 4, 3]
    assert abs(oracle.table("wt").sum() - 1.0) < 1e-14
