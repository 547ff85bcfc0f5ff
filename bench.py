#!/usr/bin/env python
"""Benchmark of the OBSGRID objective-analysis hot path on B200.

    python bench.py --gpus N --steps K --warmup W          # this repo (CUDA, C ABI)
    python bench.py --impl reference ...                   # CPU: the oracle port of the reference

One step = one analysis time of the workload: the proc_qc pass (error_max + buddy check over
every level x variable slab) followed by the proc_oa pass (multi-scan Cressman over every
slab).  Metric (BASELINE.json): Cressman grid-point x observation contributing updates per
second (pairs with d2 < R2, oa_module.F90:553/616); buddy-check obs/s is reported beside it.

  value  : inputs resident in HBM (og_dev_qc_time + og_dev_oa_time), CUDA events on the
           library's stream, max over ranks; updates per step / step time.
  e2e    : the same step through the asynchronous per-time C ABI with pinned HOST arrays
           (og_time_upload / _compute / _download / _wait): every step uploads its inputs and
           reads the analysed fields + flags back; up to 3 time periods are in flight.  The
           blocking one-call form (og_qc_oa_time) is timed beside it (e2e.blocking_call_ms).
  N > 1  : one process per GPU (torchrun); every rank analyses its own time of the same
           configuration (times are independent) -> weak scaling, no collective on
           the data path; torch.distributed only provides the barrier and the max-reduce.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from obsgrid_b200 import synth  # noqa: E402
from obsgrid_b200._capi import OgTimeDesc  # noqa: E402

METRIC = "cressman_gridpoint_x_obs_updates_per_s"
UNIT = "updates/s"
# SURVEY 8d, algorithmic FP32 operations per contributing (grid point, ob) update, the divide counted
# once: circle 11; ellipse +9 (stream coordinates, division by the elongation); banana about +25
# (atan2, sqrt, angle wrap).  The kernels count the updates per shape (og_stats).
ALGO_FLOP_PER_UPDATE = 11
ALGO_FLOP_ELLIPSE = 20
ALGO_FLOP_BANANA = 36
FIELD_NAMES = ("t", "u", "v", "rh", "slp")
FLAG_NAMES = ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp")
TABLE_NAMES = ("xob", "yob", "lon", "elev", "ob_u", "ob_v", "ob_t", "ob_rh", "ob_slp")


def measured_traffic(cfg: int):
    """dram bytes per scan launch from the newest committed ncu capture of this config."""
    best = None
    for f in sorted(ROOT.glob(f"profiles/r*/traffic_cfg{cfg}.json")):
        best = f
    if best is None:
        return None, None
    try:
        return float(json.loads(best.read_text())["bytes_per_scan"]), str(best.relative_to(ROOT))
    except Exception:      # noqa: BLE001
        return None, None


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_case(args, rank: int):
    cfg = args.config
    # weak scaling: every rank analyses its own time period; strong: all ranks share time 0
    case = synth.make_case(cfg, time_index=0 if args.scaling == "strong" else rank)
    if args.scale_rh:
        case.oa_params.scale_cressman_rh_decreases = 1
    return case


def workload_config(case, args, world: int) -> dict:
    """The `config` object of the JSON line: what is analysed, identical for both arms."""
    return {"workload": case.label, "times_per_step": 1 if args.scaling == "strong" else world,
            "grid": [case.iew, case.jns], "levels": int(case.kbu), "reports": case.nrep,
            "scans": [int(r) for r in case.oa_params.radius_influence if r >= 0],
            "scale_cressman_rh_decreases": bool(case.oa_params.scale_cressman_rh_decreases)}


def qc_like_the_gpu_arm(O, case, threads: int):
    """proc_qc of every level before the timed proc_oa, so that the CPU arm analyses exactly the
    observations the GPU arm does (obs that fail QC are not used, proc_oa.F90:462).  Untimed, through
    the oracle's binned buddy check (bit-identical to the N^2 station loops)."""
    O.set_fast_paths(True, 1)
    try:
        O.qc_time(case, threads=threads)
    finally:
        O.set_fast_paths(False, 1)


# ------------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's CPU algorithm (oracle port; the Fortran original
    cannot be built in this image: no Fortran compiler, no netCDF -- profiles/r2/toolchain_probe.txt)
    on all host threads.  One step = proc_oa of one analysis time after the same proc_qc the GPU arm
    runs (untimed here: the metric is Cressman updates per second); bounded by --ref-levels."""
    if rank != 0:
        return
    from oracle import ogoracle as O
    case0 = build_case(args, 0)
    threads = os.cpu_count() or 1
    k_end = min(case0.kbu, args.ref_levels)
    sample = (f"proc_oa over levels [0,{k_end}) of {case0.kbu} of {case0.label} "
              f"(all 5 surface slabs + {4 * (k_end - 1)} upper slabs) after proc_qc of every level, "
              f"{threads} pthreads (slabs dealt to threads; grid rows inside the large slabs)")
    times, upd = [], 0
    qc_done = case0.copy()
    qc_like_the_gpu_arm(O, qc_done, threads)
    for it in range(args.warmup + args.steps):
        c = qc_done.copy()
        O.reset_stats()
        t0 = time.perf_counter()
        O.oa_time(c, 0, k_end, threads=threads)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
            upd = O.get_stats()["cressman_updates"]
    ms = 1e3 * float(np.mean(times))
    val = upd / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(case0, args, world),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": sample, "updates_per_step": upd},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(args, case0):
    """Oracle timed on the host cores on a bounded sample of the same workload (rank 0, N=1):
    proc_oa on all threads and on one thread (the reference is serial), the as-written N^2 buddy
    check on the upper levels."""
    from oracle import ogoracle as O
    threads = os.cpu_count() or 1
    k_end = min(case0.kbu, args.ref_levels)
    qc_done = case0.copy()
    qc_like_the_gpu_arm(O, qc_done, threads)
    c = qc_done.copy()
    O.reset_stats()
    t0 = time.perf_counter()
    O.oa_time(c, 0, k_end, threads=threads)
    t_oa = time.perf_counter() - t0
    st = O.get_stats()
    # one thread, like the reference itself: a few upper levels (about a tenth of the time)
    k1a, k1b = min(1, case0.kbu - 1), min(case0.kbu, 1 + args.ref_single_levels)
    c = qc_done.copy()
    O.reset_stats()
    t0 = time.perf_counter()
    O.oa_time(c, k1a, k1b, threads=1)
    t_one = time.perf_counter() - t0
    s1 = O.get_stats()
    # buddy check as written: upper levels only (5 k obs each); the 50 k-ob surface slabs are
    # N^2 = 2.5e9 pair tests apiece on one thread and would blow the bound
    c = case0.copy()
    O.reset_stats()
    kq0, kq1 = 1, min(case0.kbu, 1 + args.ref_qc_levels)
    t0 = time.perf_counter()
    O.qc_time(c, kq0, kq1, run_consistency=False, threads=threads)
    t_qc = time.perf_counter() - t0
    sq = O.get_stats()
    n_obs_qc = 0
    for k in range(kq0, kq1):
        for name in ("ob_u", "ob_v", "ob_t", "ob_rh"):
            n_obs_qc += int(np.count_nonzero(np.abs(getattr(case0, name)[k] + 888888.0) > 1))
        n_obs_qc += int(np.count_nonzero((np.abs(case0.ob_rh[k] + 888888.0) > 1) &
                                         (np.abs(case0.ob_t[k] + 888888.0) > 1)))
    return {"value": st["cressman_updates"] / t_oa, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"proc_oa levels [0,{k_end}) of {case0.kbu} after proc_qc of every level ({t_oa:.1f} s, "
                       f"{st['cressman_updates']:.3e} updates, {st['cressman_candidates']:.3e} "
                       f"window iterations); buddy: proc_qc levels [{kq0},{kq1}) as written "
                       f"({t_qc:.1f} s, {sq['buddy_pair_tests']:.3e} pair tests)"),
            "updates_per_step": st["cressman_updates"],
            "single_thread": {"value": s1["cressman_updates"] / t_one, "unit": UNIT, "cores": 1,
                              "sample": f"proc_oa levels [{k1a},{k1b}) on one thread ({t_one:.1f} s, "
                                        f"{s1['cressman_updates']:.3e} updates)"},
            "buddy_obs_per_s": n_obs_qc / t_qc, "buddy_pair_tests_per_s": sq["buddy_pair_tests"] / t_qc,
            "oa_seconds": t_oa, "qc_seconds": t_qc}


# ------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank: int) -> str:
    """Pin this process to the CPUs next to its GPU (pinned staging memory then lands on the
    GPU's NUMA node).  Best effort: returns what was done for the JSON line."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i",
                              str(local_rank)], capture_output=True, text=True, timeout=20).stdout.strip()
        bus = out.lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        cpus = Path(f"/sys/bus/pci/devices/{bus}/local_cpulist").read_text().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return f"cpus {cpus}"
    except Exception as e:          # noqa: BLE001
        return f"unbound ({type(e).__name__})"
    return "unbound"


def strong_block(args, ctx, cfg: int, rank: int, world: int, dist, dev, lib_stream):
    """Strong scaling of ONE analysis time of configuration `cfg` over the `world` ranks, resident:
    (level, variable) slabs dealt out by cost (obsgrid_b200/shard.py), the surface flag rows
    OR-combined with one NCCL all-reduce between QC and OA.  Rank 0 also times the whole time period
    on its own GPU, so that the line carries its own 1-GPU denominator."""
    import torch
    from obsgrid_b200 import shard
    case = synth.make_case(cfg, time_index=0)
    names = FIELD_NAMES + FLAG_NAMES + TABLE_NAMES
    pristine = {nm: torch.from_numpy(getattr(case, nm)).to(dev) for nm in names}
    work = {nm: (pristine[nm].clone() if nm in FIELD_NAMES + FLAG_NAMES else pristine[nm]) for nm in names}
    d = case.desc()
    for nm in names:
        setattr(d, nm, work[nm].data_ptr())
    plan = shard.plan_slabs(case, world)[rank]
    lev = ([0] if plan["sfc_mask"] else []) + list(range(plan["k0"], plan["k1"]))
    msk = ([plan["sfc_mask"]] if plan["sfc_mask"] else []) + [0] * (plan["k1"] - plan["k0"])
    sfc_oa = plan["sfc_mask"] & 0x1f
    oa_lev = ([0] if sfc_oa else []) + list(range(plan["k0"], plan["k1"]))
    oa_msk = ([sfc_oa] if sfc_oa else []) + [0] * (plan["k1"] - plan["k0"])

    def restore():
        torch.cuda.synchronize(dev)
        for nm in FIELD_NAMES + FLAG_NAMES:
            work[nm].copy_(pristine[nm])
        torch.cuda.synchronize(dev)

    ev_x = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def step_sharded():
        ctx.dev_proc_qc_items(d, case.qc_params, lev, msk, True)
        with torch.cuda.stream(lib_stream):
            ev_x[0].record(lib_stream)
            rows = torch.stack([work["qc_u"][0], work["qc_v"][0], work["qc_t"][0], work["qc_rh"][0], work["qc_slp"]])
            shard.allreduce_or(dist, rows)
            for i_, nm_ in enumerate(("qc_u", "qc_v", "qc_t", "qc_rh")):
                work[nm_][0].copy_(rows[i_])
            work["qc_slp"].copy_(rows[4])
            ev_x[1].record(lib_stream)
        ctx.dev_qc_consistency(d, 0, 1)
        ctx.dev_proc_oa_items(d, case.oa_params, oa_lev, oa_msk)

    # second cut: QC of EVERY slab on every rank with the buddy check's cells dealt to the ranks
    # (og_set_exchange; every rank ends with every flag, no OR step), OA slabs dealt by OA cost alone
    plan_c = shard.plan_slabs(case, world, oa_only=True)[rank]
    sfc_c = plan_c["sfc_mask"] & 0x1f
    oa_lev_c = ([0] if sfc_c else []) + list(range(plan_c["k0"], plan_c["k1"]))
    oa_msk_c = ([sfc_c] if sfc_c else []) + [0] * (plan_c["k1"] - plan_c["k0"])
    xcount = {}
    xfn = shard.make_exchange(dist, torch, dev, counter=xcount)

    def step_cells():
        ctx.set_exchange(rank, world, xfn)
        ctx.dev_proc_qc(d, case.qc_params, 0, case.kbu, True)
        ctx.set_exchange(0, 1, None)          # the OA slabs below differ from rank to rank: no exchange
        if oa_lev_c:
            ctx.dev_proc_oa_items(d, case.oa_params, oa_lev_c, oa_msk_c)

    # third cut: only the surface level -- every report, the one level worth more than a rank's share --
    # has its QC cut inside the slabs; upper levels stay whole (QC + OA) on one rank each, so the
    # O(N) work of their QC (selection, transposes, lists) is not replicated
    # The units are dealt on MEASURED costs -- what a run over many analysis times has from the
    # previous time (the network hardly changes from one time to the next); here a calibration pass,
    # the units spread over the ranks, each timed alone on its GPU.
    def unit_ms(fn):
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(); ctx.sync()
        a_.record(lib_stream); fn(); b_.record(lib_stream); ctx.sync()
        return a_.elapsed_time(b_)

    sfc_bits = [b_ for _, b_ in shard.SFC_VARS if b_ & 0x1f]
    unit_list = [("s", b_) for b_ in sfc_bits] + [("k", k_) for k_ in range(1, case.kbu)]
    costs = torch.zeros(len(unit_list), device=dev, dtype=torch.float64)
    restore()
    mine_u = [i_ for i_ in range(len(unit_list)) if i_ % world == rank]
    if any(unit_list[i_][0] == "s" for i_ in mine_u):
        ctx.dev_proc_qc_items(d, case.qc_params, [0], [0], True)     # surface OA runs on checked flags
    for i_ in mine_u:
        kind, ident = unit_list[i_]
        if kind == "s":
            costs[i_] = unit_ms(lambda: ctx.dev_proc_oa_items(d, case.oa_params, [0], [ident]))
        else:
            def one_level(k_=ident):
                ctx.dev_proc_qc_items(d, case.qc_params, [k_], [0], True)
                ctx.dev_proc_oa_items(d, case.oa_params, [k_], [0])
            costs[i_] = unit_ms(one_level)
    dist.all_reduce(costs, op=dist.ReduceOp.SUM)
    costs = [float(c_) for c_ in costs]
    plans_h = shard.plan_units_measured({b_: costs[i_] for i_, b_ in enumerate(sfc_bits)}, costs[len(sfc_bits):], world)
    plan_h = plans_h[rank]
    sfc_h = plan_h["sfc_mask"] & 0x1f
    up_h = plan_h["levels"]
    oa_lev_h = ([0] if sfc_h else []) + up_h
    oa_msk_h = ([sfc_h] if sfc_h else []) + [0] * len(up_h)
    ev_h = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step_hybrid():
        ev_h[0].record(lib_stream)
        ctx.set_exchange(rank, world, xfn)
        ctx.dev_proc_qc_items(d, case.qc_params, [0], [0], True)
        ctx.set_exchange(0, 1, None)
        ev_h[1].record(lib_stream)
        if up_h:
            ctx.dev_proc_qc_items(d, case.qc_params, up_h, [0] * len(up_h), True)
        ev_h[2].record(lib_stream)
        if oa_lev_h:
            ctx.dev_proc_oa_items(d, case.oa_params, oa_lev_h, oa_msk_h)
        ev_h[3].record(lib_stream)

    # fourth cut: the surface level entirely inside its slabs -- QC as above and the Cressman cells of its
    # five slabs dealt to the ranks too (the slab summed over the ranks after every scan); upper levels
    # dealt whole on their measured costs
    plans_g = shard.plan_units_measured({}, costs[len(sfc_bits):], world)
    up_g = plans_g[rank]["levels"]
    ev_g = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step_surface_inside():
        ev_g[0].record(lib_stream)
        ctx.set_exchange(rank, world, xfn)
        ctx.dev_proc_qc_items(d, case.qc_params, [0], [0], True)
        ev_g[1].record(lib_stream)
        ctx.dev_proc_oa_items(d, case.oa_params, [0], [0])
        ctx.set_exchange(0, 1, None)
        ev_g[2].record(lib_stream)
        if up_g:
            ctx.dev_proc_qc_items(d, case.qc_params, up_g, [0] * len(up_g), True)
            ctx.dev_proc_oa_items(d, case.oa_params, up_g, [0] * len(up_g))
        ev_g[3].record(lib_stream)

    def step_alone():
        ctx.dev_proc_qc(d, case.qc_params, 0, case.kbu, True)
        ctx.dev_proc_oa(d, case.oa_params, 0, case.kbu)

    phases_h = [0.0, 0.0, 0.0]
    phases_g = [0.0, 0.0, 0.0]

    def timed(step, n, collective=True):
        tot, xch = 0.0, 0.0
        for _ in range(n):
            restore()
            if collective:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(lib_stream); step(); b.record(lib_stream)
            ctx.sync()
            tot += a.elapsed_time(b)
            if step is step_sharded:
                xch += ev_x[0].elapsed_time(ev_x[1])
            if step is step_hybrid:
                for i_ in range(3):
                    phases_h[i_] += ev_h[i_].elapsed_time(ev_h[i_ + 1]) / n
            if step is step_surface_inside:
                for i_ in range(3):
                    phases_g[i_] += ev_g[i_].elapsed_time(ev_g[i_ + 1]) / n
        return tot / n, xch / n

    n = max(2, min(args.steps, 3))
    for _ in range(2):
        restore(); step_sharded(); ctx.sync()
    ms_n, ms_x = timed(step_sharded, n)
    for _ in range(2):
        restore(); step_cells(); ctx.sync()
    xcount.clear()
    ms_c, _ = timed(step_cells, n)
    xcells = dict(xcount)
    for _ in range(2):
        restore(); step_hybrid(); ctx.sync()
    xcount.clear()
    ms_h, _ = timed(step_hybrid, n)
    xhyb = dict(xcount)
    for _ in range(2):
        restore(); step_surface_inside(); ctx.sync()
    xcount.clear()
    ms_g, _ = timed(step_surface_inside, n)
    v = torch.tensor([ms_n, ms_x, ms_c, ms_h, ms_g], device=dev, dtype=torch.float64)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    ms_n, ms_x, ms_c, ms_h, ms_g = (float(x_) for x_ in v)

    def per_rank(vals):
        t_ = torch.tensor(vals, device=dev, dtype=torch.float64)
        all_ = [torch.zeros_like(t_) for _ in range(world)]
        dist.all_gather(all_, t_)
        return [[round(float(x), 3) for x in q_] for q_ in all_]
    ph_all = per_rank(phases_h)
    pg_all = per_rank(phases_g)
    ms_1 = 0.0
    if rank == 0:
        for _ in range(2):
            restore(); step_alone(); ctx.sync()
        ms_1, _ = timed(step_alone, n, collective=False)
    dist.barrier()
    del work, pristine
    torch.cuda.empty_cache()
    best = min(ms_n, ms_c, ms_h, ms_g)
    return {"workload": case.label, "n_gpus": world, "ms_per_time": best, "ms_per_time_1gpu": ms_1,
            "speedup": (ms_1 / best) if best > 0 else None,
            "efficiency": (ms_1 / (world * best)) if best > 0 else None,
            "steps": n,
            "cuts": {
                "slabs": {"ms_per_time": ms_n, "or_exchange_ms": ms_x,
                          "what": "(level, variable) slabs of one analysis time dealt to the ranks by cost; surface "
                                  "flag rows OR-combined by one NCCL all-reduce between QC and OA"},
                "cells": {"ms_per_time": ms_c, "exchanges_per_time": xcells.get("calls", 0) / n,
                          "exchange_bytes_per_time": xcells.get("bytes", 0) / n,
                          "what": "QC of every slab on every rank with the buddy check's cells dealt to the ranks "
                                  "(og_set_exchange: relaxation mask per fix-point pass and the flags combined by NCCL "
                                  "all-reduce MAX on the library's stream); OA slabs dealt by OA cost"},
                "surface_inside": {"ms_per_time": ms_g, "exchanges_per_time": xcount.get("calls", 0) / n,
                                   "exchange_bytes_per_time": xcount.get("bytes", 0) / n,
                                   "phase_ms_per_rank": {"qc_surface_shared": [p_[0] for p_ in pg_all],
                                                         "oa_surface_shared": [p_[1] for p_ in pg_all],
                                                         "upper_levels_qc_oa": [p_[2] for p_ in pg_all]},
                                   "what": "surface level: buddy-check cells AND Cressman cells of its slabs dealt to the "
                                           "ranks (slab summed over the ranks as int32 after every scan); upper levels "
                                           "whole on one rank each, dealt on measured costs"},
                "surface_cells": {"ms_per_time": ms_h, "exchanges_per_time": xhyb.get("calls", 0) / n,
                                  "exchange_bytes_per_time": xhyb.get("bytes", 0) / n,
                                  "phase_ms_per_rank": {"qc_surface_shared": [p_[0] for p_ in ph_all],
                                                        "qc_upper_levels": [p_[1] for p_ in ph_all],
                                                        "oa": [p_[2] for p_ in ph_all]},
                                  "planned_load_ms_per_rank": [round(p_["load"], 3) for p_ in plans_h],
                                  "what": "surface level: QC of its slabs on every rank with the buddy check's cells dealt "
                                          "to the ranks, OA slabs dealt out; upper levels whole (QC + OA) on one rank "
                                          "each; units dealt longest-first on costs measured in a calibration pass (a "
                                          "production run has them from the previous analysis time)"}}}


def run_gpu(args, rank: int, world: int, local_rank: int):
    import torch
    from obsgrid_b200.api import Context

    # stdout carries exactly one JSON line; NCCL and friends print banners there
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    affinity = bind_to_gpu_numa_node(local_rank) if world > 1 else "not needed (1 process)"

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = Context(local_rank)
    ctx.set_arithmetic(args.arith)
    lib_stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    case = build_case(args, rank)
    K = case.kbu
    KR = 1 if case.oa_params.surface_only else K      # FDDA surface analyses touch level 0 only
    K0 = 0
    slab_plan = None
    if args.scaling == "strong" and world > 1:
        # one time period, its levels cut into contiguous ranges balanced on a cost estimate
        # (obsgrid_b200/shard.py); no exchange between ranks, the host writer gathers afterwards
        from obsgrid_b200 import shard
        rng_ = shard.plan_levels(shard.level_costs(case)[:KR], world)
        K0, KR = rng_[rank]
        if args.shard == "slabs" and not case.oa_params.surface_only:
            # resident path: the surface level's variable slabs are dealt out on their own
            # (shard.plan_slabs); the flags of the surface level are combined between the phases
            slab_plan = shard.plan_slabs(case, world)[rank]
            item_levels = ([0] if slab_plan["sfc_mask"] else []) + list(range(slab_plan["k0"], slab_plan["k1"]))
            item_masks = ([slab_plan["sfc_mask"]] if slab_plan["sfc_mask"] else []) + [0] * (slab_plan["k1"] - slab_plan["k0"])
            # OA has no DEWPOINT slab (bit 6): a rank that owns only that surface unit analyses no surface slab
            sfc_oa = slab_plan["sfc_mask"] & 0x1f
            oa_levels = ([0] if sfc_oa else []) + list(range(slab_plan["k0"], slab_plan["k1"]))
            oa_masks = ([sfc_oa] if sfc_oa else []) + [0] * (slab_plan["k1"] - slab_plan["k0"])

    # ---- resident copy: pristine inputs + working set in HBM ---------------------------------
    pristine, work = {}, {}
    for nm in FIELD_NAMES + FLAG_NAMES + TABLE_NAMES:
        a = getattr(case, nm)
        pristine[nm] = torch.from_numpy(a).to(dev)
        work[nm] = pristine[nm].clone() if nm in FIELD_NAMES + FLAG_NAMES else pristine[nm]
    ddesc = case.desc()
    for nm in FIELD_NAMES + FLAG_NAMES + TABLE_NAMES:
        setattr(ddesc, nm, work[nm].data_ptr())

    def restore_resident():
        torch.cuda.synchronize(dev)      # the library's stream may still be working on `work`
        for nm in FIELD_NAMES + FLAG_NAMES:
            work[nm].copy_(pristine[nm])
        torch.cuda.synchronize(dev)

    def resident_qc():
        if slab_plan is None:
            ctx.dev_proc_qc(ddesc, case.qc_params, K0, KR, True)
            return
        # one batch: the rank's surface slabs and its upper levels share every launch
        ctx.dev_proc_qc_items(ddesc, case.qc_params, item_levels, item_masks, True)
        # exchange: every rank started from the same flags and QC only sets bits -> bitwise OR of the
        # surface flag rows over the ranks (NCCL has no OR: the three QC bits travel as 0/1 planes
        # under MAX, shard.allreduce_or), on the library's stream; then the consistency pass of the
        # surface level (qc0_module.F90:371-404) on the merged rows
        with torch.cuda.stream(lib_stream):
            rows = torch.stack([work["qc_u"][0], work["qc_v"][0], work["qc_t"][0], work["qc_rh"][0], work["qc_slp"]])
            shard.allreduce_or(dist, rows)
            for i_, nm_ in enumerate(("qc_u", "qc_v", "qc_t", "qc_rh")):
                work[nm_][0].copy_(rows[i_])
            work["qc_slp"].copy_(rows[4])
        ctx.dev_qc_consistency(ddesc, 0, 1)

    def resident_oa():
        if slab_plan is None:
            ctx.dev_proc_oa(ddesc, case.oa_params, K0, KR)
            return
        ctx.dev_proc_oa_items(ddesc, case.oa_params, oa_levels, oa_masks)

    def step_resident():
        resident_qc()
        resident_oa()

    # ---- e2e: pinned host arrays through the host-pointer C ABI -----------------------------
    do_e2e = not args.no_e2e
    if do_e2e:
        host0 = {nm: getattr(case, nm).copy() for nm in FIELD_NAMES + FLAG_NAMES}
        pin = {}
        for nm in FIELD_NAMES + FLAG_NAMES + TABLE_NAMES:
            t = torch.from_numpy(getattr(case, nm).copy()).pin_memory()
            pin[nm] = t
        hcase = case.copy()
        for nm in FIELD_NAMES + FLAG_NAMES + TABLE_NAMES:
            setattr(hcase, nm, pin[nm].numpy())

    def restore_host():
        for nm in FIELD_NAMES + FLAG_NAMES:
            np.copyto(getattr(hcase, nm), host0[nm])

    def step_e2e():
        # the reference's driver calls proc_qc then proc_oa for a time period; og_qc_oa_time is
        # that pair in one call, levels pipelined through the copy engines
        ctx.proc_qc_oa(hcase, K0, KR, pipeline_groups=args.pipeline_groups)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- one counted, untimed pass: updates per step, parity spot check vs the e2e path ------
    restore_resident()
    ctx.reset_stats(); ctx.set_counting(True)
    step_resident(); ctx.sync()
    st = ctx.stats()
    ctx.set_counting(False)
    updates, candidates = st["cressman_updates"], st["cressman_candidates"]
    upd_ell, upd_ban = st["cressman_updates_ellipse"], st["cressman_updates_banana"]
    buddy_obs, pair_tests = st["buddy_obs"], st["buddy_pair_tests"]
    if do_e2e:
        restore_host(); step_e2e()
        ref_out = {nm: getattr(hcase, nm).copy() for nm in FIELD_NAMES + FLAG_NAMES}
        if slab_plan is None:
            same = all(np.array_equal(work[nm].cpu().numpy(), ref_out[nm]) for nm in FIELD_NAMES + FLAG_NAMES)
            assert same, "resident and host-pointer paths disagree"
        else:
            # the resident path analysed this rank's (level, variable) slabs, the e2e path its level range:
            # compare where both did the work -- the upper levels the two plans share
            ka, kb = max(K0, slab_plan["k0"]), min(KR, slab_plan["k1"])
            same = all(np.array_equal(work[nm].cpu().numpy()[ka:kb], ref_out[nm][ka:kb])
                       for nm in ("t", "u", "v", "rh", "qc_u", "qc_v", "qc_t", "qc_rh"))
            assert same, "slab-sharded resident path and level-sharded host-pointer path disagree"

    # ---- warm-up -------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        restore_resident(); step_resident()
    ctx.sync()

    # ---- timed: resident -------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ctx.set_timing(True)
    ctx.reset_stats()
    barrier()
    for i in range(args.steps):
        restore_resident()
        ev[i][0].record(lib_stream)
        resident_qc()
        ev[i][1].record(lib_stream)
        resident_oa()
        ev[i][2].record(lib_stream)
    ctx.sync()
    barrier()
    st_timed = ctx.stats()
    launches = st_timed["kernel_launches"]
    pairs_examined_step = st_timed["buddy_pairs_examined"] / max(args.steps, 1)
    scan_ms, buddy_ms = ctx.kernel_ms()
    ctx.set_timing(False)
    qc_ms = sum(a.elapsed_time(b) for a, b, _ in ev) / args.steps
    oa_ms = sum(b.elapsed_time(c) for _, b, c in ev) / args.steps
    step_ms = qc_ms + oa_ms

    # ---- the same steps in the other arithmetic mode of the scan kernels (og_set_arithmetic), so that
    # ---- one line carries both: 'contracted' keeps the summation order and stays within the
    # ---- north-star tolerance of the fields (tests/test_gpu_parity.py::test_contracted_*), flags are
    # ---- bit-identical in both modes
    other_mode = None
    if world == 1:
        other = "contracted" if args.arith == "exact" else "exact"
        ctx.set_arithmetic(other)
        for _ in range(2):
            restore_resident(); step_resident()
        ctx.sync()
        ctx.set_timing(True); ctx.reset_stats()
        ev_o = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            restore_resident()
            ev_o[i][0].record(lib_stream); step_resident(); ev_o[i][1].record(lib_stream)
        ctx.sync()
        scan_o, _ = ctx.kernel_ms()
        ctx.set_timing(False)
        ctx.set_arithmetic(args.arith)
        restore_resident(); step_resident(); ctx.sync()      # `work` holds the headline mode's result again
        other_mode = {"scan_arithmetic": other, "ms_per_step": sum(a.elapsed_time(b) for a, b in ev_o) / args.steps,
                      "cressman_scan_kernels_ms": scan_o / args.steps}

    e2e_ms, e2e_blocking_ms, st_e = 0.0, None, {"h2d_bytes": 0, "d2h_bytes": 0}
    if do_e2e:
        # ---- timed: e2e --------------------------------------------------------------------------------
        # The public asynchronous per-time API (og_time_upload / _compute / _download / _wait): every
        # step uploads the time period's inputs from pinned host arrays, runs proc_qc + proc_oa and
        # reads the analysed fields and flags back into pinned host arrays; up to NS time periods are
        # in flight, so the copies of neighbouring steps overlap the compute of the current one.
        # Two contexts driven from two host threads ("lanes") when the case is small enough for two
        # sets of device mirrors: while one lane waits on a list-size round trip the other lane's
        # kernels keep the SMs busy.  Lane 0 is `ctx`; every lane has NS slots of its own.
        case_bytes = sum(getattr(case, nm).nbytes for nm in FIELD_NAMES + FLAG_NAMES + TABLE_NAMES)
        n_lanes = args.e2e_lanes if args.e2e_lanes > 0 else (2 if case_bytes < (4 << 30) else 1)
        NS = 3 if (n_lanes == 1 or case_bytes < (1 << 30)) else 2
        lanes = [ctx] + [Context(local_rank) for _ in range(n_lanes - 1)]
        for c_ in lanes[1:]:
            c_.set_arithmetic(args.arith)
        pinned = lambda a: torch.from_numpy(np.ascontiguousarray(a).copy()).pin_memory().numpy()
        ocase, ocsr = [], []
        for s_ in range(NS * n_lanes):
            oc = case.copy()
            for nm in FIELD_NAMES + FLAG_NAMES + TABLE_NAMES:
                setattr(oc, nm, pinned(getattr(case, nm)))
            ocase.append(oc)
            ocsr.append(synth.CsrObs(case, alloc=pinned) if not args.e2e_dense else None)
        # the observation rows travel in sparse form (og_obs_csr: per level the reports that have a
        # measurement there) unless --e2e-dense: upper-air rows are ~90 % "missing"
        hcsr = synth.CsrObs(case, alloc=pinned) if not args.e2e_dense else None
        restore_host()

        lane_errors = []

        def lane_run(l_, nsteps):
            try:
                torch.cuda.set_device(local_rank)    # a new host thread starts on device 0
                lane_steps(l_, nsteps)
            except BaseException as e_:      # noqa: BLE001  (re-raised on the main thread)
                lane_errors.append(e_)

        def lane_steps(l_, nsteps):
            c_, outs = lanes[l_], ocase[l_ * NS:(l_ + 1) * NS]
            if nsteps <= 0:
                return
            csrs = ocsr[l_ * NS:(l_ + 1) * NS]
            up = (lambda slot: c_.time_upload(hcase, slot, K0, KR)) if hcsr is None else \
                 (lambda slot: c_.time_upload_csr(hcase, hcsr, slot, K0, KR))
            down = (lambda slot: c_.time_download(outs[slot], slot, K0, KR)) if hcsr is None else \
                   (lambda slot: c_.time_download_csr(outs[slot], csrs[slot], slot, K0, KR))
            up(0)
            for i in range(nsteps):
                s_ = i % NS
                if i + 1 < nsteps:
                    up((i + 1) % NS)
                c_.time_compute(hcase, s_, K0, KR)
                down(s_)
            for s_ in range(NS):
                c_.time_wait(s_)
            c_.sync()

        def run_e2e(nsteps):
            share = [nsteps // n_lanes + (1 if l_ < nsteps % n_lanes else 0) for l_ in range(n_lanes)]
            if n_lanes == 1:
                lane_steps(0, nsteps)
                return
            th = [threading.Thread(target=lane_run, args=(l_, share[l_])) for l_ in range(n_lanes)]
            for t_ in th:
                t_.start()
            for t_ in th:
                t_.join()
            if lane_errors:
                raise lane_errors[0]

        blocking = []
        for _ in range(3):
            restore_host()
            t0 = time.perf_counter(); step_e2e(); ctx.sync()
            blocking.append((time.perf_counter() - t0) * 1e3)
        e2e_blocking_ms = float(np.mean(blocking[1:]))
        restore_host()
        run_e2e((NS + 1) * n_lanes)   # touches every slot of every lane: no allocation is left for the timed steps
        if hcsr is not None:
            for oc_, cs_ in zip(ocase, ocsr):
                cs_.scatter_flags(oc_)      # (outside the timed region: the compact flag rows are the result)
        same = all(np.array_equal(ref_out[nm], getattr(oc_, nm)) for nm in FIELD_NAMES + FLAG_NAMES for oc_ in ocase)
        assert same, "blocking and asynchronous host-pointer paths disagree"
        for c_ in lanes:
            c_.reset_stats()
        barrier()
        t0 = time.perf_counter()
        run_e2e(args.steps)
        ctx.sync()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps      # wall clock: host staging counts
        barrier()
        st_e = {k_: sum(c_.stats()[k_] for c_ in lanes) for k_ in ("h2d_bytes", "d2h_bytes")}
        for c_ in lanes[1:]:
            c_.close()
    # ---- the per-slab drop-in path: what an unmodified proc_qc / proc_oa costs -----------------------
    per_slab = None
    if args.per_slab and world == 1:
        from obsgrid_b200 import perslab
        pc = case.copy()
        perslab.proc_qc_per_slab(ctx, pc); perslab.proc_oa_per_slab(ctx, pc)        # warm-up, and parity
        same = all(np.array_equal(getattr(pc, nm), work[nm].cpu().numpy()) for nm in FIELD_NAMES + FLAG_NAMES)
        assert same, "per-slab drop-ins and the batched path disagree"
        pc = case.copy()
        t0 = time.perf_counter()
        n_calls = perslab.proc_qc_per_slab(ctx, pc) + perslab.proc_oa_per_slab(ctx, pc)
        ps_ms = (time.perf_counter() - t0) * 1e3
        per_slab = {"value": updates / (ps_ms * 1e-3), "unit": UNIT, "ms_per_step": ps_ms, "library_calls": n_calls,
                    "api": "the reference's slab loops over og_error_max / og_buddy_check / og_innovation / og_cressman, "
                           "host arrays, one blocking call per slab and stage (obsgrid_b200/perslab.py)"}
    # ---- BASELINE configs[3] as the pipeline it is: the surface-FDDA chain, resident -----------------
    fdda_line = None
    if args.config == 4 and world == 1 and args.fdda_times > 1:
        from obsgrid_b200 import fdda
        cases3d, hourly = fdda.synthetic_chain(4, args.fdda_times, 6)
        be = fdda.DeviceBackend(ctx, torch, dev)
        for c_ in cases3d + hourly:
            be.upload(c_)
        keep = {id(c_): {nm: be.tensors(c_)[nm].clone() for nm in FIELD_NAMES + FLAG_NAMES} for c_ in cases3d + hourly}

        def chain_once():
            torch.cuda.synchronize(dev)
            for c_ in cases3d + hourly:
                for nm, src in keep[id(c_)].items():
                    be.tensors(c_)[nm].copy_(src)
            torch.cuda.synchronize(dev)
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record(lib_stream)
            fdda.run_chain(be, cases3d, hourly, 6)
            b_.record(lib_stream)
            ctx.sync()
            return a_.elapsed_time(b_)

        chain_once()
        ms_chain = float(np.mean([chain_once() for _ in range(2)]))
        fdda_line = {"traditional_times": len(cases3d), "fdda_times": len(hourly), "ms_per_chain": ms_chain,
                     "ms_per_analysis_time": ms_chain / (len(cases3d) + len(hourly)),
                     "what": "3-D analyses (27 levels) of the traditional times, store_fa, temporal_interp, proc_qc with "
                             "ob_density / obs_distance and surface-only proc_oa of every FDDA time in between; every "
                             "field resident in HBM (obsgrid_b200/fdda.py, driver.F90:107-113,306-341,561-563)"}
        del keep, be
    clocks = sampler.stop()

    # ---- strong scaling of single analysis times, inside the same line (N > 1, weak runs) ---------
    strong = None
    if dist is not None and args.scaling == "weak" and args.strong_configs:
        strong = []
        for cfg_ in [int(x) for x in args.strong_configs.split(",") if x]:
            r_ = strong_block(args, ctx, cfg_, rank, world, dist, dev, lib_stream)
            strong.append(r_)

    # ---- reduce over ranks: max time, summed work ------------------------------------------------
    vec = torch.tensor([step_ms, e2e_ms, qc_ms, oa_ms, scan_ms / args.steps], device=dev, dtype=torch.float64)
    tot = torch.tensor([updates, buddy_obs, pair_tests, candidates, launches, upd_ell, upd_ban], device=dev,
                       dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    step_ms, e2e_ms, qc_ms, oa_ms, scan_ms_step = [float(x) for x in vec.tolist()]
    g_updates, g_bobs, g_pairs, g_cands, g_launch, g_ell, g_ban = [float(x) for x in tot.tolist()]

    if rank == 0:
        hbm_peak, peak_src = peaks()
        fp32_peak_ops = ctx.fp32_peak()                     # lane-instructions/s, FADD/FMUL no FMA
        traffic_bytes, traffic_src = measured_traffic(args.config)
        n_scan_launch = sum(1 for r in case.oa_params.radius_influence if r >= 0)
        per_launch_ms = scan_ms_step / max(n_scan_launch, 1)
        # rank 0's own kernels against rank 0's own peak
        algo_flops = (ALGO_FLOP_PER_UPDATE * (updates - upd_ell - upd_ban) + ALGO_FLOP_ELLIPSE * upd_ell +
                      ALGO_FLOP_BANANA * upd_ban)
        flops_per_launch = algo_flops / max(n_scan_launch, 1)
        achieved_tf = flops_per_launch / (per_launch_ms * 1e-3) / 1e12 if per_launch_ms > 0 else 0.0
        if slab_plan is None:
            n_oa_slabs = (5 if K0 == 0 else 0) + 4 * (KR - max(K0, 1))
        else:
            n_oa_slabs = bin(slab_plan["sfc_mask"] & 0x1f).count("1") + 4 * (slab_plan["k1"] - slab_plan["k0"])
        hbm_bytes_launch = n_oa_slabs * 8.0 * case.iew * case.jns
        roofline = {
            "kernel": "cressman_scan_kernel", "bound": "fp32",
            "achieved": achieved_tf, "peak": fp32_peak_ops / 1e12, "unit": "TFLOP/s",
            "frac": achieved_tf / (fp32_peak_ops / 1e12) if fp32_peak_ops else None,
            "peak_source": "measured in this run: og_fp32_peak (FMUL+FADD chains, no FMA: the "
                           "parity build has -fmad=false, 1 lane-instruction = 1 FLOP)",
            "algorithmic_flop_per_update": {"circle": ALGO_FLOP_PER_UPDATE, "ellipse": ALGO_FLOP_ELLIPSE,
                                            "banana": ALGO_FLOP_BANANA,
                                            "mean_over_this_workload": algo_flops / max(updates, 1)},
            "updates_by_shape": {"circle": updates - upd_ell - upd_ban, "ellipse": upd_ell, "banana": upd_ban},
            "frac_if_every_update_were_a_circle": (ALGO_FLOP_PER_UPDATE * updates / max(n_scan_launch, 1) /
                                                   (per_launch_ms * 1e-3) / fp32_peak_ops) if per_launch_ms > 0 and fp32_peak_ops else None,
            "avg_launch_ms": per_launch_ms, "launches_per_step": n_scan_launch,
            "traffic": traffic_bytes, "traffic_source": traffic_src,
            "hbm": {"achieved": hbm_bytes_launch / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0,
                    "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": hbm_bytes_launch},
            "scan_kernel_share_of_oa": scan_ms_step / oa_ms if oa_ms > 0 else None,
        }
        roofline["hbm"]["frac"] = roofline["hbm"]["achieved"] / hbm_peak
        if other_mode is not None and other_mode["cressman_scan_kernels_ms"] > 0:
            # same algorithmic work, same peak: the fraction scales with the kernel time
            other_mode["frac"] = roofline["frac"] * scan_ms_step / other_mode["cressman_scan_kernels_ms"]
            roofline["other_arithmetic"] = other_mode
        # the buddy check's all-pairs kernel against the same lane-instruction peak: 12 lane operations
        # per examined pair as the reference writes them (2 sub, 2 mul, add, 2 compares, and, count,
        # select, add + the load; the kernel issues fewer: two candidates share packed instructions),
        # and how much of the QC stage is that kernel (the rest builds lists)
        b_ms = buddy_ms / max(args.steps, 1)
        roofline["buddy"] = {
            "kernel": "buddy_pass1_kernel", "bound": "fp32", "instructions_per_examined_pair": 12,
            "pairs_examined_per_step": pairs_examined_step, "avg_launch_ms": b_ms,
            "achieved": 12.0 * pairs_examined_step / (b_ms * 1e-3) / 1e12 if b_ms > 0 else 0.0,
            "peak": fp32_peak_ops / 1e12, "unit": "T lane-instr/s",
            "frac": (12.0 * pairs_examined_step / (b_ms * 1e-3) / fp32_peak_ops) if b_ms > 0 and fp32_peak_ops else None,
            "share_of_qc_stage": b_ms / qc_ms if qc_ms > 0 else None,
            "reference_pair_tests_per_step": pair_tests,
        }
        cpu = cpu_baseline(args, case) if (world == 1 and not args.no_cpu) else None
        line = {
            "metric": METRIC, "value": g_updates / (step_ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(case, args, world),
            "run": {"scan_arithmetic": ctx.arithmetic(),
                    "l2": "inputs larger than L2 (4 fields x levels x grid) and restored from a "
                          "pristine copy before every step",
                    "parallelism": ((f"1 analysis time, (level, variable) slabs sharded over {world} ranks, surface "
                                     f"flags OR-combined with one NCCL all-reduce between QC and OA; e2e shards whole levels"
                                     if slab_plan is not None else
                                     f"1 analysis time, levels sharded over {world} ranks") if args.scaling == "strong"
                                    else f"{world} x independent analysis time"),
                    "cpu_affinity": affinity},
            "e2e": ({"value": g_updates / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                     "api": (f"og_time_upload{'' if args.e2e_dense else '_csr'}/_compute/_download{'' if args.e2e_dense else '_csr'}/_wait, "
                             f"pinned host arrays, observation rows {'dense' if args.e2e_dense else 'sparse (og_obs_csr)'}, {n_lanes} context(s) x "
                             f"{NS} time periods in flight" + (", one host thread per context" if n_lanes > 1 else "")),
                     "blocking_call_ms": e2e_blocking_ms,
                     "h2d_bytes_per_step": st_e["h2d_bytes"] / args.steps,
                     "d2h_bytes_per_step": st_e["d2h_bytes"] / args.steps,
                     # what one rank moved over PCIe, both directions together, while the e2e steps ran; the
                     # ceiling of the box (profiles/r1/pcie_diag_8gpu.txt): ~49 GB/s per direction for one
                     # GPU alone, 8-11 GB/s per GPU and direction when all eight move data both ways
                     "pcie_gbs_per_gpu": (st_e["h2d_bytes"] + st_e["d2h_bytes"]) / args.steps / (e2e_ms * 1e-3) / 1e9}
                    if do_e2e else None),
            "e2e_per_slab": per_slab,
            "fdda_chain": fdda_line,
            "strong": strong,
            "gpu_launches": int(g_launch),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "breakdown": {"qc_ms": qc_ms, "oa_ms": oa_ms, "cressman_scan_kernels_ms": scan_ms_step,
                          "buddy_allpairs_kernel_ms": buddy_ms / args.steps,
                          "cressman_updates_per_step": g_updates, "cressman_candidates_per_step": g_cands,
                          "cressman_updates_per_s_oa_only": g_updates / (oa_ms * 1e-3),
                          "buddy_obs_per_step": g_bobs, "buddy_pair_tests_per_step": g_pairs,
                          "buddy_obs_per_s": g_bobs / (qc_ms * 1e-3),
                          "buddy_pair_tests_per_s": g_pairs / (qc_ms * 1e-3)},
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json configs[config-1]")
    ap.add_argument("--scale-rh", action="store_true")
    ap.add_argument("--arith", default="exact", choices=["exact", "contracted"],
                    help="scan-kernel arithmetic (og_set_arithmetic); exact = bit-identical to the reference's")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-per-slab", dest="per_slab", action="store_false",
                    help="skip the e2e_per_slab leg (the per-slab drop-ins called slab by slab)")
    ap.add_argument("--no-e2e", action="store_true",
                    help="skip the host-pointer leg (e2e: null); for auxiliary scaling runs of the largest "
                         "configuration, whose pinned host copies would not fit eight times into one box")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: one analysis time per rank; strong: one time, levels sharded over the ranks")
    ap.add_argument("--shard", default="slabs", choices=["slabs", "levels"],
                    help="strong scaling: deal out the surface level's variable slabs (slabs) or whole levels only")
    ap.add_argument("--strong-configs", default="3,5",
                    help="N > 1: configurations whose single analysis time is also strong-scaled over the ranks "
                         "(the `strong` block of the line); empty string: none")
    ap.add_argument("--fdda-times", type=int, default=5,
                    help="--config 4: traditional (6-hourly, 3-D) analysis times of the surface-FDDA chain; "
                         "5 of them bracket 20 hourly surface analyses (0 or 1: skip the chain)")
    ap.add_argument("--e2e-dense", action="store_true",
                    help="e2e leg: dense [kbu*nrep] observation rows (og_time_upload) instead of the sparse form")
    ap.add_argument("--e2e-lanes", type=int, default=0,
                    help="contexts (each on its own host thread) of the e2e leg; 0: two when the case is small")
    ap.add_argument("--pipeline-groups", type=int, default=0,
                    help="level groups of the pipelined e2e call (0: library default)")
    ap.add_argument("--ref-levels", type=int, default=64,
                    help="levels in the CPU sample of proc_oa (bounds the CPU time)")
    ap.add_argument("--ref-qc-levels", type=int, default=8)
    ap.add_argument("--ref-single-levels", type=int, default=4,
                    help="upper levels in the single-thread CPU sample of proc_oa")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
