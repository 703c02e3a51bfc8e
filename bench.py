#!/usr/bin/env python
"""Benchmark of the SPEEDY hot path: simulated member-days per wall-second (T30L8).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--members M] [--impl b200|reference] [--config 3|1|2|4|5]

Default workload (BASELINE.json configs[2], --config 3): an M = 4096 member T30L8 perturbed-IC ensemble, members sharded
across the N GPUs (pyspeedy_b200.distributed: one process per GPU, no communication inside a time step, one ncclAllReduce
issued by libspeedy_b200.so for the once-a-day ensemble mean/spread).  Synthetic set-up: member 0 is initialised from the
packaged boundary conditions, cloned on the device and every member's temperature is perturbed by i.i.d. N(0, 0.01 K)
grid-point noise (examples/Ensemble_forecast.ipynb cell 8).  A "step" is one model time step (40 min) of all members;
36 steps = 1 member-day per member.  No torch: the launcher (torchrun) only provides RANK / WORLD_SIZE / LOCAL_RANK.

  value : whole-job throughput, state resident in HBM, K steps by one spdy_run_steps call per rank
  e2e   : the same K steps through the package's public run loop -- SpeedyEns.run(callbacks=[DiagnosticCheck,
          EnsembleStatistics]) (pyspeedy/speedy.py:547-593): the loop advances the members to the next step at which a
          callback acts (one multi-step driver call with host handle arrays in and per-member error codes out, then the
          batched date update); once per simulated day -- the region is aligned so that its last step is one of them -- the
          batched diagnostics check and the ensemble mean/spread of the 6 default outputs (spectral2grid of every member
          with the partial sums in its epilogue, NCCL all-reduce, one device-to-host copy)
  roofline     : dominant kernel class of one step AS THE TIMED REGION RUNS IT (an intermediate step of a multi-step call:
                 the column physics does not store the 39 doubles per column that nothing reads before the next step
                 overwrites them), algorithmic bytes / CUDA-event time vs the measured HBM peak, on a 512-member launch
                 from an idle GPU (burst peak) and in situ (full shard, hot clocks); `last_step` = the same for the step
                 that stores everything (every step of a per-step driver call)
  cpu_baseline : the oracle (C++ restatement of the reference Fortran, which cannot be built in this image) driven like
                 parallel_step with OpenMP over members on all host cores, bounded sample
--impl reference times that CPU path as the reference arm.  --config 1 / 2 / 4 / 5 run the other BASELINE configurations
(single member; 64 members x 30 days; 16,384-field spectral chain; 1M-column physics) and print one JSON line each.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from datetime import datetime, timedelta

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "simulated member-days per wall-second (T30L8)"
UNIT = "member-days/s"
NSTEPS_DAY = 36
DT_STEP = timedelta(seconds=86400 / NSTEPS_DAY)

# algorithmic bytes per member-step of each kernel class (SURVEY.md 8d per-unit figures)
SPEC_B, FOUR_B, GRID_B = 15872, 23808, 36864
SPEC_MASK_B = 527 * 16  # rows n <= 31 - m of a spectral field (the nsh2 mask): what the fused transforms read / write
ALG_BYTES = {
    "legendre_inv": 77 * (SPEC_B + FOUR_B),
    "fft_inv": 77 * (FOUR_B + GRID_B),
    # 73 Fourier outputs; inputs counted once per launch: 33 plain fields, and for the fused product loaders the
    # distinct operand fields (u, v, T') + (u, v, q) + (u, v) = 64 grid fields feeding 40 transforms
    "fft_fwd": 33 * (GRID_B + FOUR_B) + 64 * GRID_B + 40 * FOUR_B,
    "legendre_dir": 73 * (FOUR_B + SPEC_B),
    "grid_dyn": 4608 * (50 + 33) * 8,
    # 158 doubles per column on a long-wave-only step, 167 on a short-wave step (every 3rd): average of 3 steps
    "physics": 4608 * (2 * 158 + 167) * 8 // 3,
    # 496 coefficients inside the triangular truncation: 73 forward outputs + 8 phi + 2 x 33 state rows + 2 correction rows
    # read, 2 x 33 state rows written; the 496 outside it: their 2 x 33 state rows read (zero tendency, stored only if changed)
    "spec_step": 496 * 16 * ((73 + 8 + 2 * 33 + 2 + 2 * 33) + 2 * 33),
}
# default path: each direction is ONE fused kernel, the Fourier array stays on chip; both touch only the rows of a spectral
# field inside the nsh2 mask (the inverse reads them, the forward writes them for the spectral step)
ALG_FUSED = {"fft_inv": 77 * (SPEC_MASK_B + GRID_B), "legendre_inv": 0,
             "fft_fwd": (33 + 64) * GRID_B + 73 * SPEC_MASK_B, "legendre_dir": 0}
PHYS_SW_B, PHYS_LW_B = 4608 * 167 * 8, 4608 * 158 * 8  # per member, short-wave / long-wave-only step
# intermediate steps of a multi-step driver call (physics.cu, Ctx::diag_out): 39 doubles per column are not stored (cbmf,
# precnv, precls, ustr/vstr/slru x3, shf/evap land + mean, slrd, slr, olr, rad_flux x4, rad_st4a x16)
PHYS_LAZY_B = 4608 * 39 * 8
# ... and (dynamics.cu, k_scan_outer) the 465 coefficients with m + n >= 32 of the 33 prognostic fields are not read
SPEC_SKIP_B = 465 * 16 * 2 * 33


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fp:
            return json.load(fp), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def ncu_traffic(n_per_launch):
    """DRAM bytes per launch of each kernel class from the committed `ncu --set full` capture of this round (per launch
    of `members_per_launch` members; profiles/README.md says how it was taken).  None when the launch size differs."""
    for name in ("ncu_traffic_r2.json", "ncu_traffic_r1.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.isfile(p):
            with open(p) as fp:
                tj = json.load(fp)
            if tj.get("members_per_launch") == n_per_launch:
                return tj["bytes"], name
    return {}, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        """Number of samples so far: the samples between two marks were taken during the region in between."""
        return len(self.rows)

    def stop(self, first=0, last=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[first:(last + 1 if last is not None else None)] or self.rows[-1:]
        for r in rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_reference_run(steps, warmup, members=None):
    """The oracle driven like parallel_step (speedy_driver.f90.j2:58-79) on all host threads."""
    from oracle import oracle as O

    # all host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its ranks, which would leave the CPU
    # arm on one core, so the thread count is taken from the affinity mask and passed explicitly
    cores = host_cores()
    members = members or 4 * cores
    st0 = O.State(n_months=1)
    ctl0 = O.Control((1982, 1, 1, 0, 0), (1982, 1, 11, 0, 0))
    O.load_default_bc(st0)
    assert st0.init(ctl0) == 0
    states = [st0] + [st0.clone() for _ in range(members - 1)]
    ctls = [ctl0] + [ctl0.clone() for _ in range(members - 1)]
    rng = np.random.default_rng(1234)
    for s in states:  # perturbed initial conditions (small, spectral): every member follows its own trajectory
        t = s["t"]
        t[:, :, :, 0] += 1e-4 * rng.standard_normal(t[:, :, :, 0].shape) * (np.abs(t[:, :, :, 0]) > 0)
        s["t"] = t
    for _ in range(warmup):
        assert (O.parallel_step(states, ctls, cores) == 0).all()
    t0 = time.perf_counter()
    for _ in range(steps):
        assert (O.parallel_step(states, ctls, cores) == 0).all()
    dt = time.perf_counter() - t0
    return dict(value=members * steps / NSTEPS_DAY / dt, seconds=dt, cores=cores, members=members, steps=steps)


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "T30L8 perturbed-IC ensemble (BASELINE configs[2]), CPU sample", "members": r["members"]},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": f"{r['members']} members x {args.steps} steps, OpenMP dynamic over members"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
def roofline_of(prof, n_launch, peaks, peak_kind, label, intermediate=False):
    """Per-class algorithmic GB/s of one instrumented step (class times `prof` in ms, `n_launch` members per launch)."""
    alg = dict(ALG_BYTES)
    if prof["legendre_inv"] == 0.0 and prof["legendre_dir"] == 0.0:
        alg.update(ALG_FUSED)
    if intermediate:
        alg["physics"] -= PHYS_LAZY_B
        alg["spec_step"] -= SPEC_SKIP_B
    cls = max(alg, key=lambda k: prof[k])
    total = sum(prof.values())
    achieved = alg[cls] * n_launch / (prof[cls] * 1e-3) / 1e9
    return {
        "kernel": cls, "achieved": achieved, "frac": achieved / peaks["hbm_gbs"], "when": label,
        "algorithmic_bytes_per_launch": alg[cls] * n_launch, "members_per_launch": n_launch,
        "share_of_step": prof[cls] / total,
        "per_class_ms": {k: round(v, 4) for k, v in prof.items()},
        "per_class_gbs": {k: round(alg[k] * n_launch / (prof[k] * 1e-3) / 1e9, 1) for k in alg if prof[k] > 0},
        "step_algorithmic_gbs": sum(alg.values()) * n_launch / (total * 1e-3) / 1e9,
        "_alg": alg,
    }


def profile_mean(_speedy, s, c, reps=3, intermediate=False):
    """Three consecutive instrumented steps = one short-wave step + two long-wave-only steps, averaged per class."""
    _speedy.profile_step(s, c, intermediate)
    prof = None
    for _ in range(reps):
        p1, _ = _speedy.profile_step(s, c, intermediate)
        prof = p1 if prof is None else {k: prof[k] + p1[k] for k in p1}
    return {k: v / reps for k, v in prof.items()}


def run_config3(args):
    from pyspeedy_b200 import SpeedyEns, _driver, _speedy, distributed
    from pyspeedy_b200.callbacks import DiagnosticCheck, EnsembleStatistics

    comm = distributed.init()
    rank, world = comm.rank, comm.world
    lib = _driver.lib()
    m_total = args.members
    _, m_local = comm.shard(m_total)
    lib.spdy_reserve(m_local)
    start, end = datetime(1982, 1, 1), datetime(1982, 1, 11)
    ens = SpeedyEns(m_total, start_date=start, end_date=end, comm=comm)
    ens.set_bc(perturb_sigma=0.01, seed=1234)
    s, c = ens.handles()

    # ---- warm-up, then the device-resident timed region --------------------------------------------------------
    # the clock sampler (nvidia-smi -lms 100 needs a few hundred ms to deliver its first line) starts before the warm-up;
    # the samples taken between the two marks below are those of the timed region
    sampler = ClockSampler(comm.local_rank)
    if rank == 0:
        sampler.start()
    warm = max(args.warmup, 3)
    err = _speedy.run_steps(s, c, warm)
    assert (err == 0).all(), err
    comm.barrier()
    l0 = lib.spdy_kernel_launches()
    mark0 = sampler.mark()
    lib.spdy_profiler_start()  # cudaProfilerStart/Stop: `ncu --profile-from-start off` lists exactly the timed launches
    t0 = time.perf_counter()
    err = _speedy.run_steps(s, c, args.steps)  # returns after the stream has drained (device events bracket the steps)
    t_local = time.perf_counter() - t0
    lib.spdy_profiler_stop()
    dev_ms = float(lib.spdy_last_elapsed_ms())
    mark1 = sampler.mark()
    comm.barrier()
    clocks = sampler.stop(mark0, mark1) if rank == 0 else None
    launches = lib.spdy_kernel_launches() - l0
    assert (err == 0).all(), err
    t_max, dev_ms_max = comm.max(t_local), comm.max(dev_ms)
    value = m_total * args.steps / NSTEPS_DAY / t_max

    # ---- end-to-end: the package's run loop with callbacks, exactly K steps -----------------------------------
    class DailyOutput(EnsembleStatistics):
        """EnsembleStatistics(interval=36) with the time spent in it; keeps only the latest statistics."""

        def __init__(self):
            super().__init__(interval=NSTEPS_DAY)
            self.seconds, self.calls = 0.0, 0

        def __call__(self, model):
            if self.skip_flag(model):
                return
            t1 = time.perf_counter()
            super().__call__(model)
            self.times.clear()  # this is a benchmark, not a forecast archive
            for v in self.variables:
                del self.mean[v][:-1], self.spread[v][:-1]
            self.seconds += time.perf_counter() - t1
            self.calls += 1

    # align the region: its LAST step is a multiple of 36 steps, so the daily callbacks act ceil(K / 36) times inside it
    pad = (-(ens.get_current_step() + args.steps)) % NSTEPS_DAY
    if pad:
        assert (_speedy.run_steps(s, c, pad) == 0).all()
    ens.mean_and_spread(), ens.check()  # first-call allocations of the output path (sum buffers, pinned copies)
    out = DailyOutput()
    ens.current_date = end - args.steps * DT_STEP  # the loop variable of SpeedyEns.run: K steps to go
    comm.barrier()
    t0 = time.perf_counter()
    ens.run(callbacks=[DiagnosticCheck(interval=NSTEPS_DAY), out])
    lib.spdy_synchronize()
    t_e2e = comm.max(time.perf_counter() - t0)
    comm.barrier()
    e2e_value = m_total * args.steps / NSTEPS_DAY / t_e2e
    # ---- the reference's strict loop for comparison: one parallel_step driver call per step, every output of every step
    #      stored (no multi-step call anywhere), host handle arrays in and error codes out per step
    k_strict = min(args.steps, 12)
    comm.barrier()
    t0 = time.perf_counter()
    for _ in range(k_strict):
        assert (_speedy.parallel_step(s, c) == 0).all()
    lib.spdy_synchronize()
    t_strict = comm.max(time.perf_counter() - t0)
    comm.barrier()
    strict_value = m_total * k_strict / NSTEPS_DAY / t_strict
    stats_bytes = 2 * (5 * 96 * 48 * 8 + 96 * 48) * 8
    n_calls = -(-args.steps // NSTEPS_DAY)  # driver calls of the run loop: one per callback interval
    # per driver call: state / control handle arrays in, error codes out; per output time: date containers + date in,
    # statistics and the codes of the batched check out
    h2d = int(((s.nbytes + c.nbytes) * n_calls + (s.nbytes + 20) * n_calls) / args.steps)
    d2h = int((4 * m_local * n_calls + out.calls * (stats_bytes + 4 * m_local)) / args.steps)

    # ---- roofline of the dominant kernel class ------------------------------------------------------------------
    # (a) in situ: one instrumented step of the whole shard right after the timed regions (power-capped clocks, launches
    #     of up to 2048 members); (b) one 512-member launch from an idle GPU, the figure comparable with the burst HBM
    #     number of MEASURED_PEAKS.json
    peaks, peak_kind = measured_peaks()
    chunk = min(m_local, 2048)
    prof_hot = profile_mean(_speedy, s, c, intermediate=True)
    hot = roofline_of({k: v * chunk / m_local for k, v in prof_hot.items()}, chunk, peaks, peak_kind, "in situ (hot clocks)", True)
    time.sleep(2.0)
    n_prof = min(m_local, 512)
    idle = roofline_of(profile_mean(_speedy, s[:n_prof], c[:n_prof], intermediate=True), n_prof, peaks, peak_kind,
                       "idle GPU, one launch", True)
    last = roofline_of(profile_mean(_speedy, s[:n_prof], c[:n_prof]), n_prof, peaks, peak_kind, "idle GPU, one launch")
    traffic, tfile = ncu_traffic(n_prof)
    alg = idle.pop("_alg")
    alg_last = last.pop("_alg")
    hot.pop("_alg")
    per_class_traffic = {k: traffic[k] for k in alg if k in traffic and idle["per_class_ms"].get(k, 0) > 0}
    # what the timed region moved per step (intermediate steps; the spectral step also skips the rows k_scan_outer cleared)
    timed_alg = sum(alg.values()) + (PHYS_LAZY_B / args.steps)
    roofline = {
        "bound": "hbm", "kernel": idle["kernel"], "achieved": idle["achieved"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": idle["frac"], "traffic": traffic.get(idle["kernel"]), "peak_kind": peak_kind,
        "variant": "intermediate step of a multi-step driver call (what the timed region runs)", **idle,
        "per_class_traffic_bytes": per_class_traffic, "traffic_source": tfile,
        "traffic_over_algorithmic": {k: round(v / (alg[k] * n_prof), 3) for k, v in per_class_traffic.items() if alg[k]},
        "last_step": {"variant": "step that stores every output (last step of a call; every per-step driver call)",
                      "kernel": last["kernel"], "achieved": last["achieved"], "frac": last["frac"],
                      "algorithmic_bytes_per_launch": last["algorithmic_bytes_per_launch"],
                      "ms": last["per_class_ms"][last["kernel"]], "traffic": traffic.get(last["kernel"] + "_last_step"),
                      "step_algorithmic_gbs": last["step_algorithmic_gbs"]},
        "in_situ": {k: hot[k] for k in ("kernel", "achieved", "frac", "members_per_launch", "per_class_ms", "per_class_gbs",
                                        "step_algorithmic_gbs", "when")},
        "timed_region_algorithmic_gbs": timed_alg * m_local / (dev_ms_max / args.steps * 1e-3) / 1e9,
    }
    del alg_last

    if rank != 0:
        comm.destroy()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(steps=72, warmup=1, members=8 * host_cores())
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"{r['members']} members x {r['steps']} steps in {r['seconds']:.1f} s, oracle (C++ restatement; "
                         "the reference Fortran cannot be built in this image), OpenMP dynamic over members"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": 1e3 * t_max / args.steps, "device_ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "T30L8 4096-member perturbed-IC ensemble, members sharded across GPUs (BASELINE configs[2])"
                   if m_total == 4096 else f"T30L8 {m_total}-member perturbed-IC ensemble",
                   "members": m_total, "members_per_gpu": m_local, "grid": "96x48x8, T30", "steps_per_day": 36,
                   "l2": "state (11.7 MiB/member) + scratch far exceed the 126 MB L2: no flush needed",
                   "collective": "ncclAllReduce(sum, f64) of 377,856 doubles per output time, issued by libspeedy_b200.so",
                   "multi_step_call": "value = ONE spdy_run_steps call of K steps: intermediate steps do not store the 39 "
                                      "doubles per column of physics outputs that nothing reads before the next step "
                                      "overwrites them and skip spectral rows the time filter cannot change; the state "
                                      "after the call is bit-identical to K per-step calls (tests/test_ensemble_gpu.py); "
                                      "e2e.per_step_calls is the loop that stores everything on every step"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * t_e2e / args.steps, "daily_output_ms": 1e3 * out.seconds / max(out.calls, 1),
                "daily_outputs": out.calls,
                "path": "SpeedyEns.run(callbacks=[DiagnosticCheck(36), EnsembleStatistics(36)]): one multi-step driver "
                        "call per callback interval",
                "per_step_calls": {"value": strict_value, "unit": UNIT, "steps": k_strict,
                                   "path": "parallel_step once per step: every step stores every output"}},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    comm.destroy()


# ---------------------------------------------------------------------------------------------------------------------
def base_line(value, unit, steps, warmup, ms_per_step, workload, **config):
    return {"metric": METRIC if unit == UNIT else config.pop("metric"), "value": value, "unit": unit, "n_gpus": 1,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": workload, **config}}


def run_config1(args):
    """configs[0]: one member, one simulated day from the packaged boundary conditions (launch-bound: CUDA graphs)."""
    from pyspeedy_b200 import Speedy, _driver, _speedy
    from pyspeedy_b200.callbacks import ModelCheckpoint

    lib = _driver.lib()
    m = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 3))
    m.set_bc()
    s, c = np.array([m._state_cnt], dtype=np.int64), np.array([m._control_cnt], dtype=np.int64)
    _speedy.run_steps(s, c, NSTEPS_DAY)  # day 1 is the warm-up (graph capture, first daily forcing)
    l0 = lib.spdy_kernel_launches()
    t0 = time.perf_counter()
    assert (_speedy.run_steps(s, c, NSTEPS_DAY) == 0).all()
    dt = time.perf_counter() - t0
    launches = lib.spdy_kernel_launches() - l0
    m2 = Speedy(start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    m2.set_bc()
    ck = ModelCheckpoint(interval=NSTEPS_DAY)
    t0 = time.perf_counter()
    m2.run(callbacks=[ck])
    dt2 = time.perf_counter() - t0
    line = base_line(1.0 / dt, UNIT, NSTEPS_DAY, NSTEPS_DAY, 1e3 * dt / NSTEPS_DAY,
                     "T30L8 single-member 1-day forecast (BASELINE configs[0])", members=1)
    line["e2e"] = {"value": 1.0 / dt2, "unit": UNIT, "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 4 + 41 * 4608 * 8 // NSTEPS_DAY,
                   "path": "Speedy.run(callbacks=[ModelCheckpoint(36)]): one multi-step driver call per callback interval"}
    line["gpu_launches"] = int(launches)
    print(json.dumps(line))


def run_config2(args):
    """configs[1]: 64 members, 30 simulated days (1080 steps), daily ModelCheckpoint of all members."""
    from pyspeedy_b200 import SpeedyEns, _driver, _speedy
    from pyspeedy_b200.callbacks import DiagnosticCheck, ModelCheckpoint

    lib = _driver.lib()
    n, days = 64, 30

    def make():
        e = SpeedyEns(n, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 1) + timedelta(days=days))
        e.set_bc(perturb_sigma=0.01, seed=1234)
        return e

    ens = make()
    s, c = ens.handles()
    _speedy.run_steps(s, c, 3)
    l0 = lib.spdy_kernel_launches()
    t0 = time.perf_counter()
    for _ in range(days):
        assert (_speedy.run_steps(s, c, NSTEPS_DAY) == 0).all()
    dt = time.perf_counter() - t0
    launches = lib.spdy_kernel_launches() - l0
    del ens
    ens = make()
    ck = ModelCheckpoint(interval=NSTEPS_DAY)
    t0 = time.perf_counter()
    ens.run(callbacks=[DiagnosticCheck(interval=NSTEPS_DAY), ck])
    dt2 = time.perf_counter() - t0
    assert ck.dataframe["t"].shape == (days, n, 8, 48, 96)
    line = base_line(n * days / dt, UNIT, days * NSTEPS_DAY, 3, 1e3 * dt / (days * NSTEPS_DAY),
                     "T30L8 64-member perturbed-IC ensemble, 30-day forecast (BASELINE configs[1])", members=n, days=days)
    line["e2e"] = {"value": n * days / dt2, "unit": UNIT, "h2d_bytes_per_step": 3 * 8 * n + 20,
                   "d2h_bytes_per_step": 4 * n + n * 41 * 4608 * 4 // NSTEPS_DAY,
                   "path": "SpeedyEns.run(callbacks=[DiagnosticCheck(36), ModelCheckpoint(36)]): float32 outputs of all members per day"}
    line["gpu_launches"] = int(launches)
    print(json.dumps(line))


def run_config4(args):
    """configs[3]: 16,384 synthetic T30 fields through the spectral chain (vort2vel -> spec2grid -> grid2spec -> vel2vort -> gradient)."""
    from pyspeedy_b200 import _driver
    from pyspeedy_b200.synthetic import synth_spec

    lib = _driver.lib()
    npairs = 8192
    vor, div = synth_spec(npairs, seed=2024, scale=1e-5), synth_spec(npairs, seed=2025, scale=1e-5)
    vor[:, 0, 0] = 0
    div[:, 0, 0] = 0
    ms = np.zeros(6, dtype=np.float32)
    l0 = lib.spdy_kernel_launches()
    reps = max(args.steps, 5)
    assert lib.spdy_bench_spectral_chain(_driver._ptr(vor), _driver._ptr(div), npairs, reps, _driver._ptr(ms)) == 0
    launches = lib.spdy_kernel_launches() - l0
    nf = 2 * npairs
    stage_bytes = {"vort2vel": npairs * 4 * SPEC_B, "spec2grid": nf * (SPEC_MASK_B + GRID_B), "grid2spec": nf * (GRID_B + SPEC_B),
                   "vel2vort": npairs * 4 * SPEC_B, "gradient": npairs * 3 * SPEC_B}
    peaks, peak_kind = measured_peaks()
    per = {k: {"ms": round(float(ms[i + 1]), 4), "gbs": round(b / (float(ms[i + 1]) * 1e-3) / 1e9, 1)} for i, (k, b) in enumerate(stage_bytes.items())}
    top = max(stage_bytes, key=lambda k: per[k]["ms"])
    line = base_line(nf / (float(ms[0]) * 1e-3), "fields/s", reps, 3, float(ms[0]),
                     "spectral transform microbench: 16,384 synthetic T30 fields, vort2vel -> spec2grid(kcos=2) -> grid2spec(cos loader) "
                     "-> vel2vort -> gradient, resident in HBM (BASELINE configs[3])",
                     metric="grid<->spectral round trips per second (T30 fields, with the grad/uvspec/vdspec chain)", fields=nf,
                     l2="2.6 GB of fields per rep: far larger than the 126 MB L2")
    line["roofline"] = {"bound": "hbm", "kernel": top, "achieved": per[top]["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": per[top]["gbs"] / peaks["hbm_gbs"], "traffic": None, "peak_kind": peak_kind, "per_stage": per,
                        "chain_gbs": round(sum(stage_bytes.values()) / (float(ms[0]) * 1e-3) / 1e9, 1)}
    line["gpu_launches"] = int(launches)
    print(json.dumps(line))


def run_config5(args):
    """configs[4]: >= 2^20 independent synthetic columns through the column physics, both short-wave phases."""
    from pyspeedy_b200 import SpeedyEns, _driver
    from pyspeedy_b200.synthetic import pack_column_set, synth_columns

    lib = _driver.lib()
    n, nsets = 256, 32  # 256 members x 4608 columns = 1,179,648 columns; 32 distinct column sheets (one per lane of a warp)
    ens = SpeedyEns(n, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 2))
    ens.set_bc()
    sets = []
    for k in range(nsets):
        cols, surf = synth_columns(seed=7 + k)
        sets.append(pack_column_set(*cols, tend_seed=100 + k))
        if k == 0:
            for name, val in surf.items():  # surface fields of the state: one synthetic sheet, cloned to every member
                ens.members[0][name] = val
    s, _ = ens.handles()
    from pyspeedy_b200 import _speedy

    _speedy.clone_state(int(s[0]), s[1:])
    sets = np.ascontiguousarray(np.stack(sets))
    ms = np.zeros(2, dtype=np.float32)
    reps = max(args.steps, 5)
    l0 = lib.spdy_kernel_launches()
    assert lib.spdy_bench_physics(_driver._ptr(s), n, _driver._ptr(sets), nsets, reps, _driver._ptr(ms)) == 0
    launches = lib.spdy_kernel_launches() - l0
    ncol = n * 4608
    mean_ms = (float(ms[0]) + 2 * float(ms[1])) / 3  # one short-wave step in three (nstrad = 3)
    peaks, peak_kind = measured_peaks()
    gbs = {"sw_step": PHYS_SW_B * n / (float(ms[0]) * 1e-3) / 1e9, "lw_only_step": PHYS_LW_B * n / (float(ms[1]) * 1e-3) / 1e9}
    line = base_line(ncol / (mean_ms * 1e-3), "columns/s", reps, 2, mean_ms,
                     "grid-point physics microbench: 1,179,648 synthetic kx=8 columns (256 members x 4608, 32 distinct column sheets: "
                     "every lane of a warp works on different columns), convection, condensation, SW/LW radiation, surface fluxes, "
                     "vertical diffusion (BASELINE configs[4])",
                     metric="physics columns per second (kx=8; mean of one short-wave and two long-wave-only steps)", columns=ncol,
                     l2="1.8 GB per launch: far larger than the 126 MB L2")
    line["roofline"] = {"bound": "hbm", "kernel": "k_physics", "achieved": (PHYS_SW_B + 2 * PHYS_LW_B) / 3 * n / (mean_ms * 1e-3) / 1e9,
                        "peak": peaks["hbm_gbs"], "unit": "GB/s", "traffic": None, "peak_kind": peak_kind,
                        "ms": {"sw_step": round(float(ms[0]), 4), "lw_only_step": round(float(ms[1]), 4)},
                        "gbs": {k: round(v, 1) for k, v in gbs.items()}}
    line["roofline"]["frac"] = line["roofline"]["achieved"] / peaks["hbm_gbs"]
    line["gpu_launches"] = int(launches)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=36)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--members", type=int, default=4096)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5],
                    help="BASELINE configuration (1-based; 3 = the 4096-member ensemble the metric is quoted on)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        {1: run_config1, 2: run_config2, 3: run_config3, 4: run_config4, 5: run_config5}[args.config](args)


if __name__ == "__main__":
    main()
