"""Ensemble-scale services of the library (csrc/ensemble.cu) through the C ABI: batched getters / checks / output, the
fused spectral2grid + ensemble mean / spread (with the library's own NCCL communicator, world size 1 on this box), the
reference-shaped run loop with callbacks, and sticky error codes in multi-step calls."""
import ctypes as C
from datetime import datetime, timedelta

import numpy as np
import pytest

from util import relerr

pytestmark = pytest.mark.gpu

OUT = ("u_grid", "v_grid", "t_grid", "q_grid", "phi_grid", "ps_grid")


@pytest.fixture(scope="module")
def ens40():
    """40 perturbed members (two tiles, the second ragged), six steps in."""
    from pyspeedy_b200 import SpeedyEns, _speedy

    ens = SpeedyEns(40, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 3))
    ens.set_bc(perturb_sigma=0.5, seed=11)
    s, c = ens.handles()
    assert (_speedy.run_steps(s, c, 6) == 0).all()
    return ens


def test_ensemble_get_equals_member_getters(ens40):
    from pyspeedy_b200 import _speedy

    s, _ = ens40.handles()
    _speedy.batch_spectral2grid(s)
    pick = s[[0, 7, 31, 32, 39, 3]]  # any order, across tiles
    for var in ("t_grid", "ps_grid", "vor", "precnv", "hfluxn"):
        got = _speedy.ensemble_get(pick, var)
        for i, h in enumerate(pick):
            one = getattr(_speedy, f"get_{var}")(int(h))
            assert np.array_equal(got[i].T, one), var  # bit-exact: same doubles, another route
    got32 = _speedy.ensemble_get(pick, "t_grid", dtype=np.float32)
    assert got32.dtype == np.float32 and np.array_equal(got32, _speedy.ensemble_get(pick, "t_grid").astype(np.float32))
    with pytest.raises(ValueError):
        _speedy.ensemble_get(pick, "lon")


def test_batched_to_dataframe_equals_merged_members(ens40):
    """SpeedyEns.to_dataframe (one gather per variable) == merging the members' datasets, which is how the reference
    builds it (pyspeedy/speedy.py:538-545)."""
    from pyspeedy_b200.dataset import Dataset

    fast = ens40.to_dataframe()
    slow = Dataset.merge([m.to_dataframe() for m in ens40])
    assert set(fast.keys()) == set(slow.keys()) == {"u", "v", "t", "q", "phi", "ps"}
    for k in fast.keys():
        assert fast.dims(k) == slow.dims(k)
        assert fast[k].dtype == np.float32 and np.array_equal(fast[k], slow[k]), k
    assert list(fast.coords["ens"]) == list(range(40)) and fast.coords["time"] == slow.coords["time"]
    assert np.array_equal(fast["lev"], slow["lev"]) and fast["lev"][0] > fast["lev"][-1]


def test_fused_mean_spread_with_library_communicator(ens40):
    """spectral2grid + mean / spread in one pass, all-reduced by the library's own NCCL communicator (one rank here; the
    all-reduce of a single rank must leave the sums unchanged)."""
    from pyspeedy_b200 import _driver, _speedy

    lib = _driver.lib()
    s, _ = ens40.handles()
    ref = {}
    _speedy.batch_spectral2grid(s)
    for v in OUT:
        a = _speedy.ensemble_get(s, v).astype(np.float64)
        ref[v] = (a.mean(axis=0).T, a.std(axis=0).T)
    before = ens40.mean_and_spread()
    uid = C.create_string_buffer(128)
    assert lib.spdy_comm_unique_id(uid) == 0
    assert lib.spdy_comm_init(0, 1, uid) == 0 and lib.spdy_comm_world() == 1
    try:
        after = ens40.mean_and_spread()
        x = np.array([3.0, -1.0])
        assert lib.spdy_comm_allreduce(_driver._ptr(x), 2, 1) == 0 and list(x) == [3.0, -1.0]
        assert lib.spdy_comm_barrier() == 0
    finally:
        assert lib.spdy_comm_destroy() == 0
    for v in OUT:
        for got in (before, after):
            assert relerr(got[v][0], ref[v][0]) < 1e-13, v
            # spread of O(1e-2..1) next to values of O(1e2..1e5): E[x^2] - mean^2 keeps ~8 digits
            assert np.abs(got[v][1] - ref[v][1]).max() < 1e-6 * max(ref[v][1].max(), 1e-30), v
        assert np.array_equal(before[v][0], after[v][0]) and np.array_equal(before[v][1], after[v][1])
    # a subset (ragged tiles) and a non-default variable (unfused per-variable reduction)
    sub = _speedy.ensemble_mean_spread(s[5:38])
    a = _speedy.ensemble_get(s[5:38], "t_grid")
    assert relerr(sub["t_grid"][0], a.mean(axis=0).T) < 1e-13
    p = ens40.mean_and_spread(["precnv"])["precnv"]
    a = _speedy.ensemble_get(s, "precnv")
    assert relerr(p[0], a.mean(axis=0).T) < 1e-13 and np.abs(p[1] - a.std(axis=0).T).max() < 1e-9 * max(a.std(axis=0).max(), 1e-30)


def test_batch_check_and_ensemble_check():
    from pyspeedy_b200 import SpeedyEns, _speedy

    ens = SpeedyEns(35, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    ens.set_bc()
    s, _ = ens.handles()
    assert (_speedy.batch_check(s) == 0).all()
    ens.check()
    bad = ens.members[33]
    t = bad["t"]
    t[:] = 0
    bad["t"] = t
    codes = _speedy.batch_check(s)
    assert codes[33] == -2 and (np.delete(codes, 33) == 0).all()
    assert [_speedy.check(int(h)) for h in s[[0, 33]]] == [0, -2]
    with pytest.raises(RuntimeError):
        ens.check()


def test_run_loop_with_callbacks_matches_run_steps():
    """ens.run(callbacks=[DiagnosticCheck, ModelCheckpoint, EnsembleStatistics]) -- the reference-shaped loop: one
    parallel_step, one batched date update and the callbacks per step (the next test shows it ends in exactly the
    state of run_steps)."""
    from pyspeedy_b200 import SpeedyEns
    from pyspeedy_b200.callbacks import DiagnosticCheck, EnsembleStatistics, ModelCheckpoint

    start, end = datetime(1982, 1, 1), datetime(1982, 1, 3)
    a = SpeedyEns(34, start_date=start, end_date=end)
    a.set_bc(perturb_sigma=0.1, seed=3)
    ck, st = ModelCheckpoint(interval=36, variables=["t_grid", "ps_grid"]), EnsembleStatistics(interval=36)
    a.run(callbacks=[DiagnosticCheck(interval=36), ck, st])
    assert a.current_date == end and all(m.current_date == end for m in (a.members[0], a.members[33]))
    assert a.get_current_step() == 72
    assert ck.dataframe["t"].shape == (2, 34, 8, 48, 96) and ck.dataframe["ps"].shape == (2, 34, 48, 96)
    assert ck.dataframe.coords["time"] == [start + timedelta(days=1), end] == st.times
    # the checkpointed members and the statistics describe the same ensemble
    t = ck.dataframe["t"][1].astype(np.float64)  # (ens, lev, lat, lon), lev reversed
    mean = st.mean["t_grid"][1].T[::-1]
    assert np.abs(t.mean(axis=0) - mean).max() < 1e-4 and mean.shape == (8, 48, 96)
    assert np.abs(t.std(axis=0) - st.spread["t_grid"][1].T[::-1]).max() < 1e-3


def test_multistep_call_leaves_the_complete_state():
    """Intermediate steps of a multi-step driver call do not store the column-physics outputs that the next step overwrites
    (Ctx::diag_out, physics.cu).  After the call EVERY registry variable must be what step-by-step driver calls leave --
    including the short-wave diagnostics (tsr, ssr, qcloud_equiv) that only every third step writes: 8 steps end on a
    long-wave-only step, so those come from an intermediate step of the call."""
    from pyspeedy_b200 import MODEL_STATE_DEF, SpeedyEns, _speedy

    e = SpeedyEns(4, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    e.set_bc()  # four identical members: 0, 1 take one multi-step call, 2, 3 eight per-step calls
    s, c = e.handles()
    assert (_speedy.run_steps(s[:2], c[:2], 8) == 0).all()
    for _ in range(8):
        assert (_speedy.parallel_step(s[2:], c[2:]) == 0).all()
    checked = 0
    for v in MODEL_STATE_DEF:
        x, y = e.members[1][v], e.members[3][v]
        assert np.array_equal(np.asarray(x), np.asarray(y), equal_nan=True), v
        checked += 1
    assert checked >= 109
    assert np.abs(e.members[1]["olr"]).max() > 0 and np.abs(e.members[1]["tsr"]).max() > 0


def test_run_loop_batches_steps_between_callbacks():
    """SpeedyEns.run keeps the time loop on the device up to the next step at which a callback can act (all callbacks
    derive from BaseCallback: `interval`), and falls back to one driver call per step for a plain callable."""
    from pyspeedy_b200 import SpeedyEns, _speedy
    from pyspeedy_b200.callbacks import DiagnosticCheck, EnsembleStatistics

    start, end = datetime(1982, 1, 1), datetime(1982, 1, 3)
    calls = []
    orig_run, orig_par = _speedy.run_steps, _speedy.parallel_step
    _speedy.run_steps = lambda s, c, n: (calls.append(n), orig_run(s, c, n))[1]
    _speedy.parallel_step = lambda s, c: (calls.append(1), orig_par(s, c))[1]
    try:
        a = SpeedyEns(3, start_date=start, end_date=end)
        a.set_bc()
        st = EnsembleStatistics(interval=36)
        a.run(callbacks=[DiagnosticCheck(interval=24), st])
        assert calls == [24, 12, 12, 24] and st.times == [start + timedelta(days=1), end]
        calls.clear()
        b = SpeedyEns(3, start_date=start, end_date=end)
        b.set_bc()
        seen = []
        b.run(callbacks=[lambda m: seen.append(m.current_date)])
        assert calls == [1] * 72 and len(seen) == 72
    finally:
        _speedy.run_steps, _speedy.parallel_step = orig_run, orig_par
    for v in ("vor", "t", "tr", "olr"):
        assert np.array_equal(a.members[2][v], b.members[2][v]), v


def test_seeded_ensembles_are_reproducible_per_slot():
    """perturb_temperature is a counter-based generator keyed by (seed, arena slot): the same slots with the same seed
    evolve bit-identically whether stepped by run_steps or by the run loop."""
    from pyspeedy_b200 import SpeedyEns, _speedy

    def make():
        e = SpeedyEns(33, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
        e.set_bc(perturb_sigma=0.1, seed=3)
        return e

    a = make()
    a.run()
    sa = np.sort(a.handles()[0])  # handle = arena slot + 1
    ref = {v: _speedy.ensemble_get(sa, v) for v in ("vor", "t", "ps")}
    assert not np.array_equal(ref["t"][0], ref["t"][1])  # the members do differ
    del a
    import gc

    gc.collect()
    b = make()  # takes the same set of freed slots (in another order)
    s, c = b.handles()
    assert np.array_equal(np.sort(s), sa)
    assert (_speedy.run_steps(s, c, 36) == 0).all()
    for v in ref:
        assert np.array_equal(ref[v], _speedy.ensemble_get(np.sort(s), v)), v


def test_sticky_error_in_multistep_call(oracle):
    """a member that fails at step 1 of run_steps(n = 36) must be reported at the end of the call even
    though the codes are read back once per simulated day -- and it is frozen at its failing step (date not advanced,
    speedy.f90:62-69) while its neighbours in the tile run on undisturbed."""
    from pyspeedy_b200 import SpeedyEns, _speedy

    start, end = datetime(1982, 1, 1), datetime(1982, 1, 3)
    ens = SpeedyEns(5, start_date=start, end_date=end)
    ens.set_bc(perturb_sigma=0.05, seed=9)
    # poison the second time level of T of member 2: the leapfrog step moves it into level 1, which the check reads
    t = ens.members[2]["t"]
    t[0, 0, :, 1] = 1.0e4
    ens.members[2]["t"] = t
    s, c = ens.handles()
    err = _speedy.run_steps(s, c, 36)
    assert list(err) == [0, 0, -2, 0, 0]
    assert ens.members[2]["current_step"] == 1  # counted, as in the reference (speedy.f90:59-61), then frozen
    assert _speedy.get_model_datetime(int(s[2])) == (1982, 1, 1, 0, 0)
    assert _speedy.get_model_datetime(int(s[1])) == (1982, 1, 2, 0, 0)
    assert np.isfinite(ens.members[2]["t"]).all()  # nothing stepped it into NaNs after the failure
    # the next call runs the healthy members on; the failed one is stepped again and fails again (the oracle's sequence
    # for this state: -2 at steps 1, 2 and 3, NaNs -- which pass the check -- from step 4)
    err = _speedy.run_steps(s, c, 36)
    assert err[2] == -2 and (np.delete(err, 2) == 0).all()
    assert _speedy.get_model_datetime(int(s[0])) == (1982, 1, 3, 0, 0)
    # per-step driver call: same code, date of the failed member not advanced on the host mirror either
    e1 = _speedy.parallel_step(s, c)
    assert e1[2] == -2 and e1[0] == 0


def test_check_before_the_first_multistep_call():
    """Regression (round 2): the batched check and the step driver grow the same per-tile buffers; a check made BEFORE
    the first multi-step call left the outer-coefficient flags unallocated.  Fresh process, so the order is what it says."""
    import os
    import subprocess
    import sys

    code = (
        "from datetime import datetime\n"
        "import numpy as np\n"
        "from pyspeedy_b200 import SpeedyEns, _speedy\n"
        "e = SpeedyEns(70, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))\n"
        "e.set_bc()\n"
        "e.check()\n"
        "s, c = e.handles()\n"
        "assert (_speedy.run_steps(s, c, 6) == 0).all()\n"
        "assert (_speedy.run_steps(s[:3], c[:3], 5) == 0).all()\n"
        "e.check()\n"
        "assert e.members[69]['current_step'] == 6 and e.members[1]['current_step'] == 11\n"
        "print('ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_multistep_calls_across_day_boundaries():
    """80 steps as two multi-step calls (37 + 43: both cross a daily forcing / coupler boundary and end off the short-wave
    phase) against 80 per-step calls, on perturbed members of one tile: every registry variable bit-identical."""
    from pyspeedy_b200 import MODEL_STATE_DEF, SpeedyEns, _speedy

    e = SpeedyEns(6, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 4))
    e.set_bc(perturb_sigma=0.05, seed=11)
    s, c = e.handles()
    _speedy.clone_state(int(s[0]), s[3:4]), _speedy.clone_state(int(s[1]), s[4:5]), _speedy.clone_state(int(s[2]), s[5:6])
    assert (_speedy.run_steps(s[:3], c[:3], 37) == 0).all()
    assert (_speedy.run_steps(s[:3], c[:3], 43) == 0).all()
    for _ in range(80):
        assert (_speedy.parallel_step(s[3:], c[3:]) == 0).all()
    for k in range(3):
        assert _speedy.get_model_datetime(int(s[k])) == _speedy.get_model_datetime(int(s[k + 3]))
        for v in MODEL_STATE_DEF:
            x, y = np.asarray(e.members[k][v]), np.asarray(e.members[k + 3][v])
            assert np.array_equal(x, y, equal_nan=True), (k, v)
    assert not np.array_equal(np.asarray(e.members[0]["t"]), np.asarray(e.members[1]["t"]))
