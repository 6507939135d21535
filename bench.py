#!/usr/bin/env python
"""bench.py -- photons/s of the photon-tracing hot path (computeRadiativeTransfer incl. local-estimate radiances).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (the CPU arm: the oracle port on all host threads)

A "step" is one batch: one pass of the hot path over `--photons` photons of synthetic illumination on the named
workload (default: the I3RC Landsat cloud, BASELINE.json configs[2], the configuration the north-star target is
quoted on), including normalisation and the batch-moment update.  Ranks are independent (batches are sharded like
Example-Drivers/monteCarloDriver.f95:264-274, weak scaling); the only collective is ONE all-reduce of the packed
moment buffer after the last step.

    python bench.py --scaling strong --gpus N ...            (fixed total work: --total-batches batches shared by the
                                                              ranks; domain setup and the all-reduce inside the clock)

Roofline: the gathers of the cell-crossing loop come out of L2 on C1-C4 (fields <= 40 MB), so the
ceiling is the L2 random-gather rate MEASURED in this run (i3rc_measure_gather_rate over 8 MB), HBM only when the gather
field does not fit L2; the issue-slot ceiling (i3rc_measure_issue_rate) is reported beside it, and one launch of the
transport kernel is re-run under `ncu` by this script (a separate short process, after the timed region) for the DRAM
traffic, executed instructions and hit rates of the very build that was timed.

Timing: every step is bracketed by CUDA events on the stream the kernels are launched on; L2 is flushed (256 MiB
memset) between steps outside the brackets; a barrier + synchronize surrounds the K steps; the slowest rank counts.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "photons/sec (incl. local-estimate radiances)"
UNIT = "photons/s"


# ---- workloads ------------------------------------------------------------------------------
def make_workload(name):
    from i3rc_monte_carlo_model_b200 import fields
    rr = dict(useRayTracing=True, useRussianRoulette=True, useRussianRouletteForIntensity=True, zetaMin=0.3,
              minInverseTableSize=10001, minForwardTableSize=10001)
    dirs3 = dict(intensityMus=[1.0, 0.5, 0.5], intensityPhis=[0.0, 0.0, 180.0])  # monteCarloDriver.nml
    src = dict(solarMu=0.5, solarAzimuth=0.0)
    if name == "landsat":
        return dict(workload="i3rcLandsatCloud 128x128x119, HG g=0.85 (299 moments), ssa=1, mu0=0.5, 3 radiance directions, "
                             "Russian roulette for intensity (zetaMin 0.3)",
                    domain=lambda: fields.landsat_cloud(1.0), params=dict(surfaceAlbedo=0.0, **dirs3, **rr), source=src,
                    photons=16_000_000, cpu_photons=60_000)
    if name == "step":
        return dict(workload="i3rcStepCloud 32x1x32, HG g=0.85 (64 moments), ssa=0.99, mu0=0.5, 3 radiance directions, RR",
                    domain=lambda: fields.step_cloud(0.99), params=dict(surfaceAlbedo=0.0, **dirs3, **rr), source=src,
                    photons=16_000_000, cpu_photons=200_000)
    if name == "planeparallel":
        return dict(workload="planeParallel.nml 1x1x1 tau=1 HG g=0.85, 3 radiance directions, plain local estimate",
                    domain=lambda: fields.plane_parallel(),
                    params=dict(surfaceAlbedo=0.0, useRussianRouletteForIntensity=False, **dirs3), source=src,
                    photons=8_000_000, cpu_photons=1_000_000)
    if name == "radar":
        return dict(workload="i3rcRadarCloud 640x1x54, Deirmendjian C1 (tabulated), ssa=0.99, mu0=0.5, 3 radiance directions, RR",
                    domain=lambda: fields.radar_cloud(0.99, "C1"), params=dict(surfaceAlbedo=0.0, **dirs3, **rr), source=src,
                    photons=8_000_000, cpu_photons=40_000)
    if name in ("les", "les-small"):
        n = (512, 512, 256) if name == "les" else (128, 128, 64)
        mus = [1.0, 0.8, 0.6, 0.4]
        return dict(workload=f"synthetic LES {n[0]}x{n[1]}x{n[2]}, 2 components (27-entry cloud table + absorbing gas), "
                             "16 radiance directions, RR",
                    domain=lambda: fields.synthetic_les(nx=n[0], ny=n[1], nz=n[2]),
                    params=dict(surfaceAlbedo=0.05, intensityMus=[m for m in mus for _ in range(4)],
                                intensityPhis=[p for _ in mus for p in (0.0, 90.0, 180.0, 270.0)], **rr),
                    source=dict(solarMu=0.5, solarAzimuth=30.0), photons=4_000_000, cpu_photons=20_000)
    raise SystemExit(f"unknown workload {name}")


def algorithmic_bytes(c, nc):
    """4 B per cell crossing (totalExt) + (4 nC + 16) B per collision (cumExt, ssa, pfIdx, 2 inverse-table
    entries) + 32 B per absorption (two read-modify-writes) + 24 B per local-estimate contribution (2 forward-table
    entries + one read-modify-write) + 8 B per photon exit (one read-modify-write).  A cell crossing is the reference
    algorithm's unit: cells the kernel crosses without gathering them (runs of uniform layers taken in one step) count."""
    cells = (c["crossings_photon"] + c["crossings_intensity"] + c.get("cells_skipped", 0) + c.get("cells_skipped_intensity", 0))
    return (4 * cells + (4 * nc + 16) * c["collisions"] + 32 * c["absorptions"]
            + 24 * c["contributions"] + 8 * (c["exits_top"] + c["surface_hits"]))


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                p = [t.strip() for t in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                    power.append(float(p[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return None
        busy = [s for s, w in zip(sm, power) if w > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": float(max(power))}


def timing_of(be, I):
    ms, nl, no = C.c_double(), C.c_int64(), C.c_int64()
    be.get_timing(I.handle, C.byref(ms), C.byref(nl), C.byref(no))
    return ms.value, nl.value, no.value


# ---- the CPU arm ------------------------------------------------------------------------------------------------
def cpu_port_rate(wl, seconds, threads=0, nbatches=None, fast=True):
    """Photons/s of the oracle port (OpenMP threads stand in for MPI ranks, one batch per thread at a time) on a
    bounded sample of the workload; returns (rate, cores, description, elapsed).  fast: the -O3 -march=native build of
    the same source, compiled on this machine (BASELINE.md section 3's flags for the CPU arm); else the parity build
    (-O2 -ffp-contract=off)."""
    from oracle.binding import fast_oracle_backend, oracle_backend, run_batches
    from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
    from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import (computeRadiativeTransfer, new_Integrator,
                                                                        specifyParameters)
    from i3rc_monte_carlo_model_b200.RandomNumbers import new_RandomNumberSequence
    be = fast_oracle_backend() if fast else oracle_backend()
    cores = threads or os.cpu_count() or 1
    I = new_Integrator(wl["domain"](), backend=be)
    specifyParameters(I, **wl["params"])
    # 1-photon warm-up builds the tables outside the timed region (monteCarloDriver.f95:240-253)
    computeRadiativeTransfer(I, new_RandomNumberSequence([10, 0]), new_PhotonStream(numberOfPhotons=1, **wl["source"]))
    nph = wl["cpu_photons"]
    t0 = time.perf_counter()
    run_batches(I, new_PhotonStream(numberOfPhotons=max(nph // 8, 100), **wl["source"]), 10, cores, nThreads=cores)
    probe = time.perf_counter() - t0
    rate0 = max(nph // 8, 100) * cores / probe
    if nbatches is None:
        nph = int(max(1000, min(nph * 8, rate0 * seconds / cores)))
        nbatches = cores
    t0 = time.perf_counter()
    run_batches(I, new_PhotonStream(numberOfPhotons=nph, **wl["source"]), 10, nbatches, batchBegin=1, nThreads=cores)
    el = time.perf_counter() - t0
    return nph * nbatches / el, cores, f"{nbatches} batches x {nph} photons of the same workload", el


def run_reference(args, wl, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    rates, els = [], []
    per_step = max(2.0, min(20.0, 150.0 / max(args.steps + args.warmup, 1)))
    per_step = float(os.environ.get("I3RC_BENCH_CPU_SECONDS", per_step))  # (the test-suite shortens the sample)
    sample = ""
    for i in range(args.warmup + args.steps):
        r, cores, sample, el = cpu_port_rate(wl, per_step)
        if i >= args.warmup:
            rates.append(r)
            els.append(el)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(els) * 1e3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic illumination on the workload's field",
        "config": {"workload": wl["workload"], "note": "the reference is Fortran 95 and cannot be compiled in this image (no "
                   "Fortran compiler, MPI or netCDF): this arm times the C restatement of its algorithm (oracle/, OpenMP "
                   "threads standing in for MPI ranks) on all host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample + " per step",
                         "build": "gcc -O3 -march=native -fopenmp, compiled on this machine (oracle/Makefile target `fast`)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---- one transport launch under ncu (a separate short process, after the timed region) -----------------------------
NCU_METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum",
               "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
               "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
               "lts__t_sectors_srcunit_tex_op_read.sum"]
NCU_EXTRA = ["l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
             "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]


def probe_launch(args, wl):
    """`bench.py --probe-launch`: set the workload up and trace two batches (the second one is what ncu captures)."""
    from i3rc_monte_carlo_model_b200._lib import backend
    from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
    from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import new_Integrator, specifyParameters
    be = backend()
    assert be.set_device(0) == 0
    I = new_Integrator(wl["domain"](), backend=be)
    specifyParameters(I, **wl["params"])
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        assert be.set_tuning(I.handle, k.encode(), int(v)) == 0, kv
    src = new_PhotonStream(numberOfPhotons=args.photons or wl["photons"], **wl["source"]).as_c()
    assert be.stats_reset(I.handle, 0) == 0
    for b in range(2):
        assert be.run_batches(I.handle, C.byref(src), 10, 0, 1_000_000 + b, 1) == 0, I._msg()
    be.synchronize(I.handle)
    return 0


def ncu_one_launch(args, nph, timeout=240):
    """Metrics of ONE k_transport launch of this workload at this size, measured now with ncu on this GPU.  Returns a
    dict (metric -> value, plus 'command') or None when ncu is not available / not permitted here."""
    import csv
    import shutil
    if shutil.which("ncu") is None:
        return None
    cmd_tail = [sys.executable, os.path.abspath(__file__), "--probe-launch", "--workload", args.workload, "--photons", str(nph)]
    if args.tune:
        cmd_tail += ["--tune", args.tune]
    for metrics in (NCU_METRICS + NCU_EXTRA, NCU_METRICS):
        with tempfile.NamedTemporaryFile("r", suffix=".csv") as f:
            cmd = ["ncu", "--metrics", ",".join(metrics), "--clock-control", "none", "-k", "regex:k_transport", "-s", "1", "-c", "1",
                   "--csv", "--log-file", f.name] + cmd_tail  # (-s 1: the probe's first batch warms up, the second is captured)
            try:
                r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=timeout,
                                   env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0")))
            except Exception:
                return None
            out = {}
            try:
                rows = [ln for ln in open(f.name) if ln.startswith('"')]
                for row in csv.DictReader(rows):
                    try:
                        out[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
                    except (KeyError, ValueError):
                        pass
                    out["kernel"] = row.get("Kernel Name", "")
            except Exception:
                out = {}
            if r.returncode == 0 and "dram__bytes_read.sum" in out:
                out["command"] = " ".join(cmd[:cmd.index("--log-file")] + cmd_tail[1:])
                return out
    return None


# ---- the CUDA arm -------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="landsat")
    ap.add_argument("--photons", type=int, default=0, help="photons per step (batch) and GPU")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --steps batches IN TOTAL are shared by the ranks (monteCarloDriver.f95:264-274); domain setup, "
                         "table building and the all-reduce are inside the clock")
    ap.add_argument("--c5-leg", default="auto", choices=["auto", "on", "off"],
                    help="after the main measurement, a STRONG-scaling leg on BASELINE config 5 (synthetic 512x512x256, 2 components, "
                         "16 directions) with reportVolumeAbsorption: 8 batches of 1 M photons in total, set-up and the 1.15 GB "
                         "all-reduce inside the clock; reported as `c5_strong`.  auto = when more than one GPU takes part")
    ap.add_argument("--report-volume", action="store_true", help="keep batch moments of volumeAbsorption too "
                    "(reportVolumeAbsorption: the all-reduce payload grows by 2 x 8 B per cell)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ncu", action="store_true", help="skip the ncu pass over one transport launch")
    ap.add_argument("--probe-launch", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--tune", default="", help="key=value,... passed to i3rc_set_tuning")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON: anything libraries print there (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    wl = make_workload(args.workload)
    if args.probe_launch:
        return probe_launch(args, wl)
    if args.impl == "reference":
        return run_reference(args, wl, emit)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    from i3rc_monte_carlo_model_b200 import _abi
    from i3rc_monte_carlo_model_b200._lib import backend
    from i3rc_monte_carlo_model_b200.driver import allreduce_device_stats, device_stats_report, partition_batches
    from i3rc_monte_carlo_model_b200.monteCarloIllumination import new_PhotonStream
    from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import (computeRadiativeTransfer, getCounters, new_Integrator,
                                                                        reportResults, specifyParameters)
    from i3rc_monte_carlo_model_b200.RandomNumbers import new_RandomNumberSequence

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    be = backend()
    assert be.set_device(local) == 0
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        warm = torch.zeros(1, device="cuda", dtype=torch.float64)
        dist.all_reduce(warm)  # communicator set-up (connections, buffers) is not part of any timed all-reduce below

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    strong = args.scaling == "strong"
    nph = args.photons or wl["photons"]
    K, W = args.steps, args.warmup
    domain = wl["domain"]()  # (building the synthetic field on the host is input preparation, outside every clock)
    barrier()
    t_setup0 = time.perf_counter()
    I = new_Integrator(domain, backend=be)
    assert I.handle, "new_Integrator failed"
    specifyParameters(I, **wl["params"])
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        assert be.set_tuning(I.handle, k.encode(), int(v)) == 0, kv
    assert be.tabulate(I.handle) == 0, I._msg()
    be.synchronize(I.handle)
    setup_ms = (time.perf_counter() - t_setup0) * 1e3  # upload of the domain, gather field, phase-function tables
    src = new_PhotonStream(numberOfPhotons=nph, **wl["source"]).as_c()
    stream = torch.cuda.ExternalStream(be.stream(I.handle), device=torch.device("cuda", local))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    with_volume = 1 if args.report_volume else 0
    if strong:   # K batches in total, rank r takes its block of them
        nB_total, mine = partition_batches(K, world, rank)
    else:        # K batches per rank
        nB_total, mine = partition_batches(K * world, world, rank)
    mine = list(mine)

    # ---- device-resident arm: photons generated, traced, normalised and folded into the moments on the GPU ----
    assert be.stats_reset(I.handle, with_volume) == 0
    for w in range(W):
        assert be.run_batches(I.handle, C.byref(src), 10, 0, 1_000_000 + w, 1) == 0, I._msg()
    assert be.stats_reset(I.handle, with_volume) == 0
    be.reset_timing(I.handle)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    counters = dict.fromkeys(_abi.COUNTER_FIELDS, 0)
    step_ms = []
    wall0 = time.perf_counter()
    for b in mine:
        flush.zero_()  # L2 flush between timed steps, outside the event brackets
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rc = be.run_batches(I.handle, C.byref(src), 10, 0, b, 1)
        e1.record(stream)
        assert rc == 0, I._msg()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        for k, v in getCounters(I).items():
            counters[k] += v
    # the ONE collective of the job, timed with CUDA events on torch's stream (where NCCL runs it)
    be.synchronize(I.handle)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    allreduce_device_stats(I, dist if world > 1 else None)
    a1.record()
    a1.synchronize()
    allreduce_ms = a0.elapsed_time(a1) if world > 1 else 0.0
    stats_bytes = 0
    if world > 1:
        ptr, n = C.c_void_p(), C.c_int64()
        be.stats_device_buffer(I.handle, C.byref(ptr), C.byref(n))
        stats_bytes = 8 * n.value
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    trace_ms, trace_launches, other_launches = timing_of(be, I)
    dev_ms = float(sum(step_ms)) + allreduce_ms + (setup_ms if strong else 0.0)
    if world > 1:
        t = torch.tensor([dev_ms, wall_ms, trace_ms, setup_ms, allreduce_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, trace_ms_max, setup_ms, allreduce_ms = t.tolist()
        cs = torch.tensor([counters[k] for k in _abi.COUNTER_FIELDS], device="cuda", dtype=torch.int64)
        dist.all_reduce(cs)
        counters_all = dict(zip(_abi.COUNTER_FIELDS, cs.tolist()))
    else:
        counters_all = counters
    total_batches = nB_total if strong else K * world
    total_photons = nph * total_batches
    value = total_photons / (dev_ms * 1e-3)
    stats = device_stats_report(I, 1.0, nB_total)

    # ---- end-to-end arm: the reference's own call sequence per batch with HOST buffers ----
    # Every batch: the photon arrays of type(photonStream) are FILLED ON THE HOST (2 N uniform deviates, what the reference's
    # new_PhotonStream does per batch, monteCarloDriver.f95:281 -> monteCarloIllumination.f95:91-95) -- by a pool of host
    # threads, for batch b+1 while the GPU traces batch b (two sets of pinned buffers) --, handed to
    # i3rc_computeRadiativeTransfer (host -> device copy inside), and the results are read back into host arrays.
    # Generation, copies, kernel and read-back are all inside the clock.
    e2e = None
    if not args.no_e2e and not strong:
        from concurrent.futures import ThreadPoolExecutor
        nthr = max(1, min(16, (os.cpu_count() or 2) // max(world, 1)))
        pool = ThreadPoolExecutor(nthr)
        seeds = np.random.SeedSequence(1234 + rank)

        def pinned(n):
            return torch.empty(n, dtype=torch.float32).pin_memory().numpy()

        def new_stream():
            ph = new_PhotonStream(numberOfPhotons=nph, **wl["source"])
            ph.xPosition, ph.yPosition, ph.zPosition, ph.initialMu, ph.initialPhi = (pinned(nph) for _ in range(5))
            ph.zPosition[:] = np.float32(1.0) - np.finfo(np.float32).eps
            ph.initialMu[:] = -abs(wl["source"].get("solarMu", 0.5))
            ph.initialPhi[:] = np.float32(wl["source"].get("solarAzimuth", 0.0) * np.pi / 180.0)
            return ph

        def fill(ph, batch):  # x, y ~ U(0,1): one slice per host thread, independent generator streams
            gens = [np.random.default_rng(s) for s in seeds.spawn(nthr)]
            cuts = np.linspace(0, nph, nthr + 1).astype(np.int64)

            def part(i):
                gens[i].random(out=ph.xPosition[cuts[i]:cuts[i + 1]], dtype=np.float32)
                gens[i].random(out=ph.yPosition[cuts[i]:cuts[i + 1]], dtype=np.float32)
            return [pool.submit(part, i) for i in range(nthr)]

        phs = [new_stream(), new_stream()]
        want = ["meanFluxUp", "meanFluxDown", "meanFluxAbsorbed", "fluxUp", "fluxDown", "fluxAbsorbed", "absorbedProfile",
                "meanIntensity", "intensity"]
        out = {"fluxUp": pinned(I.nx * I.ny).reshape((I.nx, I.ny), order="F"),
               "fluxDown": pinned(I.nx * I.ny).reshape((I.nx, I.ny), order="F"),
               "fluxAbsorbed": pinned(I.nx * I.ny).reshape((I.nx, I.ny), order="F"),
               "intensity": pinned(I.nx * I.ny * I.nDir).reshape((I.nx, I.ny, I.nDir), order="F")}
        d2h = 4 * (3 * I.nx * I.ny + I.nz + I.nx * I.ny * I.nDir + 3 + I.nDir) + C.sizeof(_abi.Counters)
        h2d = 5 * 4 * nph + C.sizeof(_abi.PhotonSource)
        for w in range(2):
            for f in fill(phs[w], w):
                f.result()
            computeRadiativeTransfer(I, new_RandomNumberSequence([11, 2_000_000 + w]), phs[w])
        barrier()
        be.reset_timing(I.handle)
        t0 = time.perf_counter()
        pending = fill(phs[0], mine[0])
        for i, b in enumerate(mine):
            for f in pending:
                f.result()
            if i + 1 < len(mine):
                pending = fill(phs[(i + 1) % 2], mine[i + 1])  # the next batch's photons, while this one is traced
            computeRadiativeTransfer(I, new_RandomNumberSequence([11, b]), phs[i % 2])  # H2D of the photon arrays + kernel
            r = reportResults(I, *want, out=out)                                        # D2H into host arrays
        be.synchronize(I.handle)
        el = time.perf_counter() - t0
        e2e_trace_ms = timing_of(be, I)[0]  # first transport launch to last one of every batch, copies it waits for included
        pool.shutdown()
        if world > 1:
            t = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = t.item()
        e2e = {"value": total_photons / el, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": el * 1e3 / K, "transport_ms_per_step": e2e_trace_ms / max(len(mine), 1),
               "host_threads_filling_photon_arrays": nthr,
               "path": "per batch: photon arrays of type(photonStream) filled on the host (2N uniform deviates, inside the clock, "
               "overlapped with the previous batch's kernel) -> i3rc_computeRadiativeTransfer (host->device copy inside) -> "
               "i3rc_reportResults into host arrays", "meanFluxUp_last": float(r["meanFluxUp"])}
    barrier()

    # ---- strong-scaling leg on BASELINE config 5 with volume absorption (the case whose all-reduce payload is large) ----
    c5 = None
    if args.c5_leg == "on" or (args.c5_leg == "auto" and world > 1 and args.workload == "landsat"):
        from i3rc_monte_carlo_model_b200.monteCarloRadiativeTransfer import finalize_Integrator
        wl5 = make_workload("les")
        nph5, nb5 = 1_000_000, 8
        dom5 = wl5["domain"]()  # (the synthetic field is made on the host outside the clock)
        barrier()
        t0 = time.perf_counter()
        I5 = new_Integrator(dom5, backend=be)
        assert I5.handle, "new_Integrator failed (config 5)"
        specifyParameters(I5, **wl5["params"])
        assert be.tabulate(I5.handle) == 0, I5._msg()
        be.synchronize(I5.handle)
        setup5 = (time.perf_counter() - t0) * 1e3
        src5 = new_PhotonStream(numberOfPhotons=nph5, **wl5["source"]).as_c()
        assert be.stats_reset(I5.handle, 1) == 0
        assert be.run_batches(I5.handle, C.byref(src5), 10, 0, 1_000_000, 1) == 0, I5._msg()  # warm-up
        assert be.stats_reset(I5.handle, 1) == 0
        nB5, mine5 = partition_batches(nb5, world, rank)
        stream5 = torch.cuda.ExternalStream(be.stream(I5.handle), device=torch.device("cuda", local))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream5)
        for b in mine5:
            assert be.run_batches(I5.handle, C.byref(src5), 10, 0, b, 1) == 0, I5._msg()
        e1.record(stream5)
        e1.synchronize()
        trace5 = e0.elapsed_time(e1)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        allreduce_device_stats(I5, dist if world > 1 else None)
        a1.record()
        a1.synchronize()
        ar5 = a0.elapsed_time(a1) if world > 1 else 0.0
        ptr, n = C.c_void_p(), C.c_int64()
        be.stats_device_buffer(I5.handle, C.byref(ptr), C.byref(n))
        t5 = torch.tensor([setup5 + trace5 + ar5, setup5, trace5, ar5], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        tot5, setup5, trace5, ar5 = t5.tolist()
        st5 = device_stats_report(I5, 1.0, nB5)
        c5 = {"workload": wl5["workload"], "scaling": "strong", "total_batches": nB5, "photons_per_batch": nph5,
              "value": nph5 * nB5 / (tot5 * 1e-3), "unit": UNIT, "ms_total": tot5, "setup_ms": setup5, "trace_ms": trace5,
              "allreduce_ms": ar5, "allreduce_bytes": 8 * n.value, "report_volume_absorption": True,
              "meanFluxUp": [float(st5["meanFluxUp"][0]), float(st5["meanFluxUp"][1])],
              "timing": "domain upload + gather field + tables + this rank's batches (CUDA events) + the all-reduce, max over ranks"}
        finalize_Integrator(I5)
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_transport) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    nc = I.nc
    abytes = algorithmic_bytes(counters, nc)  # this rank's launches
    launches = max(trace_launches, 1)
    kernel_s = trace_ms / launches * 1e-3
    achieved = abytes / launches / kernel_s / 1e9
    crossings = counters["crossings_photon"] + counters["crossings_intensity"]  # DDA steps = gathers issued
    cells_crossed = crossings + counters.get("cells_skipped", 0) + counters.get("cells_skipped_intensity", 0)
    ncell = I.nx * I.ny * I.nz
    nzc = be.get_layout(I.handle, 0)
    gather_field_bytes = 4 * I.nx * I.ny * (nzc if nzc > 0 else I.nz)
    l2_resident = gather_field_bytes <= 96 * 2**20  # (126 MB L2; the other fields and tables share it)
    # ceilings measured now, on this GPU: random 4-byte gathers over an L2-resident array and over 1 GiB; issue slots
    g = C.c_double()
    meas = {}
    for name, nbytes in (("l2_8MB", 8 << 20), ("l2_field", max(1 << 20, min(gather_field_bytes, 64 << 20))), ("hbm_1GiB", 1 << 30)):
        meas[name] = g.value if be.measure_gather_rate(nbytes, 3, C.byref(g)) == 0 else None
    issue_peak = g.value if be.measure_issue_rate(3, C.byref(g)) == 0 else None
    torch.cuda.synchronize()
    ncu = None if (args.no_ncu or world > 1) else ncu_one_launch(args, nph)
    traffic, traffic_source = None, None
    if ncu:
        scale = 1.0  # (the probe traces the same number of photons as a bench step)
        traffic = (ncu["dram__bytes_read.sum"] + ncu["dram__bytes_write.sum"]) * scale
        traffic_source = "ncu pass of this run: " + ncu["command"]
    else:
        tpath = os.path.join(ROOT, "profiles", "transport_traffic.json")
        try:
            ent = json.load(open(tpath)).get(args.workload, {})
            traffic, traffic_source = ent.get("dram_bytes_per_launch"), "profiles/transport_traffic.json (" + ent.get("source", "") + ")"
        except Exception:
            pass
    gather_peak = meas["l2_8MB"] if l2_resident else meas["hbm_1GiB"]
    if gather_peak:
        bound, peak, peak_source = ("l2" if l2_resident else "hbm"), gather_peak * 4 / 1e9, (
            "measured in this run: random 4-byte gathers over " + ("8 MB (L2-resident)" if l2_resident else "1 GiB (HBM)") +
            ", useful bytes (i3rc_measure_gather_rate)")
    else:
        bound, peak, peak_source = "hbm", hbm_peak, "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    issue = None
    if issue_peak:
        issue = {"peak_warp_inst_per_s": issue_peak, "peak_source": "measured in this run (independent FMAs, i3rc_measure_issue_rate); "
                 "nominal 148 SMs x 4 schedulers x SM clock"}
        if ncu and "smsp__inst_executed.sum" in ncu:
            ach = ncu["smsp__inst_executed.sum"] / (ncu["gpu__time_duration.sum"] * 1e-9)
            issue.update({"achieved_warp_inst_per_s": ach, "frac": ach / issue_peak,
                          "issue_slots_busy_pct": ncu.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                          "active_threads_per_warp_inst": ncu.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                          "warp_inst_per_cell_crossing": ncu["smsp__inst_executed.sum"] / max(crossings / launches, 1.0)})
    roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_source, "peak_source": peak_source,
                "kernel": "k_transport", "kernel_ms_per_launch": trace_ms / launches, "kernel_share_of_step": trace_ms / max(sum(step_ms), 1e-9),
                "algorithmic_bytes_per_launch": abytes / launches, "cell_crossings_per_s": cells_crossed / (trace_ms * 1e-3), "gathers_issued_per_s": crossings / (trace_ms * 1e-3),
                "bytes_per_crossing_model": "4 B/crossing + (4nC+16) B/collision + 32 B/absorption + 24 B/contribution + 8 B/exit",
                "gather_field_bytes": gather_field_bytes, "layers_stored_in_3d": nzc if nzc > 0 else I.nz,
                "measured_ceilings": {"gathers_per_s": meas, "hbm_copy_gbs": hbm_peak,
                                      "frac_of_hbm_copy_peak": achieved / hbm_peak},
                "issue": issue,
                "ncu": {k: v for k, v in (ncu or {}).items() if k != "command"} or None,
                "note": "the gathers of the cell-crossing loop are served by " + ("L2" if l2_resident else "HBM") +
                        "; which ceiling binds is read from `frac` (gather rate) against `issue.frac` (instruction issue)"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, cores, sample, el = cpu_port_rate(wl, 15.0)
        rate_parity, _, sample_p, _ = cpu_port_rate(wl, 5.0, fast=False)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "build": "gcc -O3 -march=native -fopenmp, compiled on this machine (oracle/Makefile target `fast`)",
               "parity_build_value": rate_parity, "parity_build": "gcc -O2 -ffp-contract=off (the build the parity tests use); " + sample_p,
               "note": "C restatement of the reference (oracle/), OpenMP threads standing in for MPI ranks; the Fortran reference "
                       "cannot be built in this image"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic illumination (photons drawn on the device from Philox streams) on the I3RC field shipped as a fixture"
                if args.workload in ("landsat", "radar") else "synthetic",
        "config": {"workload": wl["workload"], "photons_per_step_per_gpu": nph if not strong else None,
                   "photons_per_batch": nph, "total_batches": total_batches,
                   "parallelism": f"batches sharded over {world} GPU(s), replicated domain, one all-reduce of the moment buffer",
                   "l2": "flushed between steps (256 MiB memset), flush outside the CUDA-event brackets",
                   "timing": ("strong scaling: domain upload + gather field + tables (setup_ms) + sum of per-batch CUDA-event times + "
                              "the all-reduce, max over ranks" if strong else
                              "sum of per-step CUDA-event times on the launching stream + final all-reduce, max over ranks"),
                   "allreduce_ms": allreduce_ms, "allreduce_bytes": stats_bytes, "setup_ms": setup_ms,
                   "report_volume_absorption": bool(with_volume), "tuning": args.tune or "default"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(trace_launches + other_launches),
        "roofline": roofline, "cpu_baseline": cpu, "c5_strong": c5,
        "wall_ms_timed_region": wall_ms, "allreduce_ms": allreduce_ms,
        "counters_per_photon": {k: v / total_photons for k, v in counters_all.items() if v},
        "results": {"meanFluxUp": [float(stats["meanFluxUp"][0]), float(stats["meanFluxUp"][1])],
                    "meanFluxDown": [float(stats["meanFluxDown"][0]), float(stats["meanFluxDown"][1])],
                    "meanRadiance": [[float(m), float(e)] for m, e in zip(np.ravel(stats["meanRadiance"][0]), np.ravel(stats["meanRadiance"][1]))]
                    if "meanRadiance" in stats else None},
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
