"""Column physics kernel vs the oracle on synthetic columns (BASELINE config 5 distributions), both short-wave
phases; integer diagnostics (itop, icnv, icltop) must agree exactly."""
import ctypes as C
from datetime import datetime

import numpy as np
import pytest

from pyspeedy_b200.synthetic import synth_columns
from util import ptr, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sw", [True, False])
def test_physics_columns(oracle, drv, sw):
    from pyspeedy_b200 import Speedy

    st = oracle.State(n_months=1)
    st.init_tables()
    st.zonal_average_fields(0.03)
    m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    (ug, vg, tg, qg, phig, pslg), surf = synth_columns(seed=7)
    for k, v in surf.items():
        st[k] = v
    rng = np.random.default_rng(3)
    for name in ("flux_solar_in", "flux_ozone_lower", "flux_ozone_upper", "zenit_correction", "stratospheric_correction"):
        surf[name] = st[name]
    if not sw:  # persisted short-wave state from a previous step
        surf["rad_tau2"] = rng.uniform(0.3, 1.0, (96, 48, 8, 4))
        surf["tt_rsw"] = rng.uniform(0, 1e-5, (96, 48, 8))
        surf["rad_strat_corr"] = rng.uniform(0, 5, (96, 48, 2))
        for k in ("rad_tau2", "tt_rsw", "rad_strat_corr"):
            st[k] = surf[k]
    for k, v in surf.items():
        m[k] = v
    st["compute_shortwave"] = int(sw)
    m["compute_shortwave"] = int(sw)
    tend = [np.asfortranarray(rng.normal(0, 1e-5, (96, 48, 8))) for _ in range(4)]
    o_t = [a.copy(order="F") for a in tend]
    qg_o = qg.copy(order="F")
    dbg_o = st.physics_columns(ug, vg, tg, qg_o, phig, pslg, *o_t)
    u8, v8 = np.ascontiguousarray(ug[:, :, 7].T).T.copy(order="F"), vg[:, :, 7].copy(order="F")
    g_u8, g_v8 = tend[0][:, :, 7].copy(order="F"), tend[1][:, :, 7].copy(order="F")
    g_t, g_q = tend[2].copy(order="F"), tend[3].copy(order="F")
    dbg = np.zeros((3, 48, 96), dtype=np.int32)
    rc = drv.lib().spdy_debug_physics(m._state_cnt, ptr(u8), ptr(v8), ptr(tg), ptr(qg), ptr(phig), ptr(pslg),
                                      ptr(g_u8), ptr(g_v8), ptr(g_t), ptr(g_q), ptr(dbg))
    assert rc == 0
    assert np.array_equal(dbg[0], dbg_o[0]), "itop"
    assert np.array_equal(dbg[1], dbg_o[1]), "icnv"
    if sw:
        assert np.array_equal(dbg[2], dbg_o[2]), "icltop"
    assert relerr(g_u8, o_t[0][:, :, 7]) < 1e-12
    assert relerr(g_v8, o_t[1][:, :, 7]) < 1e-12
    assert relerr(g_t, o_t[2]) < 1e-12
    assert relerr(g_q, o_t[3]) < 1e-12
    # levels 1..7 of the wind tendencies are untouched by the physics (physics.f90:214-221)
    assert np.array_equal(o_t[0][:, :, :7], tend[0][:, :, :7])
    for v in ["precnv", "precls", "cbmf", "slrd", "slr", "olr", "slru", "ustr", "vstr", "shf", "evap", "hfluxn",
              "rad_flux", "rad_st4a"] + (["tsr", "ssrd", "ssr", "tt_rsw", "rad_tau2", "rad_strat_corr", "qcloud_equiv"] if sw else []):
        a, b = m[v], st[v]
        if v == "hfluxn":
            a, b = a[:, :, :2], b[:, :, :2]
        assert relerr(a, b) < 1e-12, (v, relerr(a, b))


def test_full_size_column_independence(oracle):
    """BASELINE config 5 size (1M columns = 224 members x 4608): the physics and every other kernel treat columns /
    members independently, so a 224-member ensemble of identical members must stay bit-identical member by member
    (the device-side sum of (x - x_member0)^2 over the members is exactly zero), while member 0 follows the oracle."""
    from pyspeedy_b200 import SpeedyEns, _speedy

    n = 224
    ens = SpeedyEns(n, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    ens.set_bc()
    hs, cs = ens.handles()
    st = oracle.State(n_months=1)
    ctl = oracle.Control((1982, 1, 1, 0, 0), (1982, 1, 2, 0, 0))
    oracle.load_default_bc(st)
    assert st.init(ctl) == 0
    for _ in range(4):  # steps 0..3: two short-wave steps, two long-wave-only steps
        assert (_speedy.parallel_step(hs, cs) == 0).all()
        assert st.step(ctl) == 0
    for v in ["t", "tr", "vor", "div", "ps"]:
        assert relerr(ens.members[0][v], st[v]) < 1e-11, v
    for v in ["precnv", "precls", "olr", "tsr", "slrd", "shf", "evap", "ustr", "hfluxn", "tt_rsw", "rad_tau2"]:
        assert relerr(ens.members[0][v], st[v]) < 1e-9, v
        _, s2 = _speedy.ensemble_sums(hs, v, shift=ens.members[0][v])  # sum over members of (x - x_member0)^2
        assert not s2.any(), v
    for i in (1, 31, 32, 100, 223):
        for v in ["t", "precnv", "olr"]:
            assert np.array_equal(ens.members[i][v], ens.members[0][v]), (i, v)


def test_fused_dynamics_physics_path():
    """SPDY_FUSE_PHYS=1 (grid-point dynamics and column physics of a column in one kernel, tendencies handed over in
    registers) against the oracle: three model steps, in a fresh process because the switch is read once."""
    import os
    import subprocess
    import sys

    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import relerr
from oracle import oracle as O
from pyspeedy_b200 import Speedy, _speedy
from datetime import datetime
st = O.State(n_months=1); ctl = O.Control((1982, 1, 1, 0, 0), (1982, 1, 2, 0, 0)); O.load_default_bc(st); assert st.init(ctl) == 0
m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2)); m.set_bc()
for _ in range(3):
    assert st.step(ctl) == 0 and _speedy.step(m._state_cnt, m._control_cnt) == 0
for v in ("vor", "div", "t", "ps", "tr"):
    assert relerr(m[v], st[v]) < 1e-11, v
print("fused physics ok")
'''
    env = dict(os.environ, SPDY_FUSE_PHYS="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "fused physics ok" in out.stdout, out.stdout + out.stderr
