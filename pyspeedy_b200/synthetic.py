"""Synthetic inputs of the two microbenchmark configurations of BASELINE.json, shared by bench.py and the
parity tests: spectral fields (config 4) and physics columns (config 5)."""
import numpy as np

MX, NX, KX, IX, IL = 31, 32, 8, 96, 48


def synth_spec(n, seed=2024, scale=1.0):
    """Config 4: spectral coefficients N(0,1) * (1 + l)^-1 inside l <= 30, zero outside, Im(m = 0) = 0.
    Shape (n, 32, 31) complex, C order = Fortran (31, 32) per field."""
    rng = np.random.default_rng(seed)
    m = np.arange(MX)[None, :]
    nn = np.arange(NX)[:, None]
    l = m + nn
    amp = np.where(l <= 30, 1.0 / (1.0 + l), 0.0)
    x = (rng.standard_normal((n, NX, MX)) + 1j * rng.standard_normal((n, NX, MX))) * amp * scale
    x[:, :, 0] = x[:, :, 0].real
    return np.ascontiguousarray(x)


FSG = np.array([0.025, 0.095, 0.2, 0.34, 0.51, 0.685, 0.835, 0.95])
HSG = np.array([0.0, 0.05, 0.14, 0.26, 0.42, 0.6, 0.77, 0.9, 1.0])


def synth_columns(seed=7):
    """Config 5: one (96, 48) sheet of independent 8-level columns.  T = reference profile (prognostics.f90:63-81) +
    N(0, 5 K); ps/p0 ~ U(0.5, 1.05); rh ~ U(0, 1.1) -> q = rh * qsat (humidity.f90:44-78); u, v ~ N(0, 10); hydrostatic phi;
    surface fields: land fraction in {0, 1, 0.37}, phis0 ~ max(0, N(0, 5e3)), sst ~ U(271, 303), land temperature
    ~ U(230, 310), albedos U(0.07, 0.6), soil water and snow cover U(0, 1).
    Returns ((ug, vg, tg, qg, phig, pslg) Fortran-ordered, dict of surface fields)."""
    rng = np.random.default_rng(seed)
    tref = 288.0 * np.maximum(0.2, FSG) ** (287.0 * 0.006 / 9.81)
    tg = tref[None, None, :] + rng.normal(0, 5.0, size=(IX, IL, KX))
    psa = rng.uniform(0.5, 1.05, size=(IX, IL))
    pslg = np.log(psa)
    e0, c1, c2, t0, t1, t2 = 6.108e-3, 17.269, 21.875, 273.16, 35.86, 7.66
    qs = np.where(tg >= t0, e0 * np.exp(c1 * (tg - t0) / (tg - t1)), e0 * np.exp(c2 * (tg - t0) / (tg - t2)))
    qs = 622.0 * qs / (FSG[None, None, :] * psa[:, :, None] - 0.378 * qs)
    qg = rng.uniform(0, 1.1, size=(IX, IL, KX)) * qs
    ug = rng.normal(0, 10, size=(IX, IL, KX))
    vg = rng.normal(0, 10, size=(IX, IL, KX))
    phis0 = np.maximum(0.0, rng.normal(0, 5e3, size=(IX, IL)))
    phig = np.zeros((IX, IL, KX))
    phig[:, :, 7] = phis0 + 287.0 * np.log(HSG[8] / FSG[7]) * tg[:, :, 7]
    for k in range(6, -1, -1):
        phig[:, :, k] = phig[:, :, k + 1] + 287.0 * np.log(FSG[k + 1] / FSG[k]) * 0.5 * (tg[:, :, k] + tg[:, :, k + 1])
    surf = dict(phis0=phis0, fmask_land=rng.choice([0.0, 1.0, 0.37], size=(IX, IL)), forog=rng.uniform(1, 1.5, (IX, IL)),
                sst_am=rng.uniform(271, 303, (IX, IL)), land_temp=rng.uniform(230, 310, (IX, IL)),
                alb_land=rng.uniform(0.07, 0.6, (IX, IL)), alb_sea=rng.uniform(0.07, 0.6, (IX, IL)),
                alb_surface=rng.uniform(0.07, 0.6, (IX, IL)), soil_avail_water=rng.uniform(0, 1, (IX, IL)),
                snowc=rng.uniform(0, 1, (IX, IL)), ssrd=rng.uniform(0, 300, (IX, IL)))
    return [np.asfortranarray(x) for x in (ug, vg, tg, qg, phig, pslg)], surf


def pack_column_set(ug, vg, tg, qg, phig, pslg, tend_seed=3):
    """One column set in the layout of ``spdy_bench_physics``: ug8, vg8, pslg, utend8, vtend8, tg, qg, phig, ttend, qtend."""
    rng = np.random.default_rng(tend_seed)
    tend = [rng.normal(0, 1e-5, (IX, IL, KX)) for _ in range(4)]
    parts = [ug[:, :, 7], vg[:, :, 7], pslg, tend[0][:, :, 7], tend[1][:, :, 7], tg, qg, phig, tend[2], tend[3]]
    return np.concatenate([np.asarray(p).ravel(order="F") for p in parts])
