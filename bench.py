#!/usr/bin/env python
"""Benchmark of the SPEEDY hot path: simulated member-days per wall-second (T30L8).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--members M] [--impl b200|reference]

Workload (BASELINE.json configs[2]): an M = 4096 member T30L8 perturbed-IC ensemble (members sharded across the N
GPUs, no communication inside a time step; NCCL only for the once-a-day ensemble mean/spread), synthetic set-up:
member 0 is initialised from the packaged boundary conditions, cloned on the device and every member's temperature
is perturbed by i.i.d. N(0, 0.01 K) grid-point noise (examples/Ensemble_forecast.ipynb cell 8).
A "step" is one model time step (40 min) of all members; 36 steps = 1 member-day per member.

  value : whole-job throughput, state resident in HBM, K steps by one spdy_run_steps call per rank
  e2e   : same metric through the reference-facing per-step driver call (parallel_step: host handle arrays in,
          per-member error codes copied back every step) plus the once-a-day output path (spectral2grid of every
          member, ensemble mean/spread of the 6 default outputs reduced on the device, NCCL all-reduce, D2H)
  roofline     : dominant kernel class of one step, algorithmic bytes / CUDA-event time vs the measured HBM peak
  cpu_baseline : the oracle (C++ restatement of the reference Fortran, which cannot be built in this image) driven
                 like parallel_step with OpenMP over members on all host cores, bounded sample
--impl reference times that CPU path as the reference arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from datetime import datetime

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "simulated member-days per wall-second (T30L8)"
UNIT = "member-days/s"
NSTEPS_DAY = 36

# algorithmic bytes per member-step of each kernel class (SURVEY.md 8d per-unit figures)
SPEC_B, FOUR_B, GRID_B = 15872, 23808, 36864
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
SPEC_MASK_B = 527 * 16  # rows n <= 31 - m of a spectral field: what the fused forward kernel writes for the spectral step


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fp:
            return json.load(fp), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


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

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(steps, warmup, members=None):
    """The oracle driven like parallel_step (speedy_driver.f90.j2:58-79) on all host threads."""
    from oracle import oracle as O

    # all host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its ranks, which would leave the CPU
    # arm on one core, so the thread count is taken from the affinity mask and passed explicitly
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else O.max_threads()
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
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


class CudaArray:
    """Zero-copy view of a device buffer for torch.as_tensor (used for the NCCL all-reduce of ensemble sums)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def run_b200(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from pyspeedy_b200 import DEFAULT_OUTPUT_VARS, SpeedyEns, _driver, _speedy

    lib = _driver.lib()
    lib.spdy_set_device(local)
    m_total = args.members
    m_local = m_total // world + (1 if rank < m_total % world else 0)
    lib.spdy_reserve(m_local)
    ens = SpeedyEns(m_local, start_date=datetime(1982, 1, 1), end_date=datetime(1982, 1, 11))
    ens.set_bc(perturb_sigma=0.01, seed=1234 + rank)
    s, c = ens.handles()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out_buf = {}

    def daily_output():
        """spectral2grid of every member + ensemble mean/spread of the default outputs; returns D2H bytes.
        The per-variable sums (reduced over this rank's members by the library) are gathered in one device buffer: one
        NCCL all-reduce and one device-to-host copy per simulated day."""
        _speedy.batch_spectral2grid(s)
        parts, off = [], 0
        for v in DEFAULT_OUTPUT_VARS:
            e = _driver.REGISTRY[_driver.VAR_ID[v]]
            dev, ne = C.c_void_p(), C.c_size_t()
            lib.spdy_ensemble_sums_device(_driver._ptr(s), len(s), e["id"], None, C.byref(dev), C.byref(ne))
            n2 = 2 * ne.value
            if "t" not in out_buf or out_buf["t"].numel() < off + n2:
                grown = torch.empty(max(off + n2, 2 * 2 * 5 * 96 * 48 * 8 + 2 * 96 * 48), dtype=torch.float64, device="cuda")
                if "t" in out_buf:
                    grown[:off].copy_(out_buf["t"][:off])
                out_buf["t"] = grown
            out_buf["t"][off:off + n2].copy_(torch.as_tensor(CudaArray(dev.value, n2), device="cuda"))
            torch.cuda.current_stream().synchronize()  # the library reuses its sum buffer for the next variable
            parts.append((off, ne.value))
            off += n2
        t = out_buf["t"][:off]
        if world > 1:
            dist.all_reduce(t)
        host = t.cpu().numpy()
        for o, n in parts:
            mean = host[o:o + n] / m_total
            spread = np.sqrt(np.maximum(host[o + n:o + 2 * n] / m_total - mean * mean, 0.0))
            del mean, spread
        return host.nbytes

    # ---- warm-up, then the device-resident timed region --------------------------------------------------------
    err = _speedy.run_steps(s, c, max(args.warmup, 3))
    assert (err == 0).all(), err
    sampler = ClockSampler(local)
    barrier()
    l0 = lib.spdy_kernel_launches()
    if rank == 0:
        sampler.start()
    lib.spdy_profiler_start()  # cudaProfilerStart/Stop: `ncu --profile-from-start off` lists exactly the timed launches
    t0 = time.perf_counter()
    err = _speedy.run_steps(s, c, args.steps)
    torch.cuda.synchronize()
    t_local = time.perf_counter() - t0
    lib.spdy_profiler_stop()
    dev_ms = float(lib.spdy_last_elapsed_ms())
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.spdy_kernel_launches() - l0
    assert (err == 0).all(), err
    t_max = max_over_ranks(t_local)
    dev_ms_max = max_over_ranks(dev_ms)
    value = m_total * args.steps / NSTEPS_DAY / t_max

    # ---- end-to-end: per-step driver calls with host buffers + the once-a-day output path ----------------------
    daily_output()  # warm-up of the output path (first-call allocations of the sum buffers, NCCL communicator, pinned copies)
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    t_out = 0.0
    for k in range(args.steps):
        e = _speedy.parallel_step(s, c)
        d2h += e.nbytes
        if (k + 1) % NSTEPS_DAY == 0 or k == args.steps - 1:
            t1 = time.perf_counter()
            d2h += daily_output()
            t_out += time.perf_counter() - t1
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = m_total * args.steps / NSTEPS_DAY / t_e2e

    # ---- roofline of the dominant kernel class (one instrumented step on one 512-member chunk) -----------------
    # the roofline compares one kernel timed alone with the burst HBM figure of MEASURED_PEAKS.json, so the profiled steps
    # start from an idle GPU: the two timed regions above leave the chip at its power cap (SM clock down to ~1.8 GHz)
    time.sleep(2.0)
    n_prof = min(m_local, 512)
    _speedy.profile_step(s[:n_prof], c[:n_prof])
    prof = None  # three consecutive steps = one short-wave step + two long-wave-only steps, averaged
    for _ in range(3):
        p1, _ = _speedy.profile_step(s[:n_prof], c[:n_prof])
        prof = p1 if prof is None else {k: prof[k] + p1[k] for k in p1}
    prof = {k: v / 3.0 for k, v in prof.items()}
    peaks, peak_kind = measured_peaks()
    alg = dict(ALG_BYTES)
    if prof["legendre_inv"] == 0.0:  # default path: spec -> grid is ONE fused kernel, the Fourier array stays on chip
        alg["fft_inv"], alg["legendre_inv"] = 77 * (SPEC_B + GRID_B), 0
    if prof["legendre_dir"] == 0.0:  # default path: grid -> spec is one fused kernel per loader mode as well
        alg["fft_fwd"], alg["legendre_dir"] = (33 + 64) * GRID_B + 73 * SPEC_MASK_B, 0
    cls = max(alg, key=lambda k: prof[k])
    achieved = alg[cls] * n_prof / (prof[cls] * 1e-3) / 1e9
    total_prof = sum(prof.values())
    traffic = None  # DRAM bytes per launch of that kernel from the committed ncu --set full capture (same launch size)
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")
    if os.path.isfile(tpath):
        with open(tpath) as fp:
            tj = json.load(fp)
        if tj["members_per_launch"] == n_prof:
            traffic = tj["bytes"].get(cls)
    roofline = {
        "bound": "hbm", "kernel": cls, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "algorithmic_bytes_per_launch": alg[cls] * n_prof,
        "peak_kind": peak_kind,
        "share_of_step": prof[cls] / total_prof,
        "per_class_ms": {k: round(v, 4) for k, v in prof.items()},
        "per_class_gbs": {k: round(alg[k] * n_prof / (prof[k] * 1e-3) / 1e9, 1) for k in alg if prof[k] > 0},
        "step_algorithmic_gbs": sum(alg.values()) * n_prof / (total_prof * 1e-3) / 1e9,
        "members_profiled": n_prof,
    }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(steps=72, warmup=1, members=8 * len(os.sched_getaffinity(0)))
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"{r['members']} members x {r['steps']} steps in {r['seconds']:.1f} s, oracle (C++ restatement; "
                         "the reference Fortran cannot be built in this image), OpenMP dynamic over members"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * t_max / args.steps, "device_ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "T30L8 4096-member perturbed-IC ensemble, members sharded across GPUs (BASELINE configs[2])"
                   if m_total == 4096 else f"T30L8 {m_total}-member perturbed-IC ensemble",
                   "members": m_total, "members_per_gpu": m_local, "grid": "96x48x8, T30", "steps_per_day": 36,
                   "l2": "state (11.7 MiB/member) + scratch far exceed the 126 MB L2: no flush needed"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(s.nbytes + c.nbytes),
                "d2h_bytes_per_step": int(d2h / args.steps), "ms_per_step": 1e3 * t_e2e / args.steps,
                "daily_output_ms": 1e3 * t_out},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=36)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--members", type=int, default=4096)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
