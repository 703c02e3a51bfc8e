"""Shared helpers of the parity tests."""
import ctypes as C

import numpy as np

MX, NX, KX, IX, IL = 31, 32, 8, 96, 48


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


from pyspeedy_b200.synthetic import synth_spec  # noqa: E402,F401  (BASELINE config 4 fields)


def relerr(a, b):
    """max |a-b| / max |b| over ONE variable -- the norm the parity tolerances are stated in."""
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def relerr_fields(a, b):
    """Batched transforms: axis 0 runs over independent fields; every field is measured against ITS OWN max|b_f| and the
    worst field is returned, so a low-amplitude field cannot hide behind a large one."""
    a, b = np.asarray(a), np.asarray(b)
    ax = tuple(range(1, b.ndim))
    den = np.abs(b).max(axis=ax)
    return float((np.abs(a - b).max(axis=ax) / np.where(den > 0, den, 1.0)).max())


def field_scales(n, seed):
    """Per-field amplitudes spread over six decades (makes the per-field norm bite)."""
    return 10.0 ** np.random.default_rng(seed).uniform(-3, 3, size=n)
