"""Sharding one analysis time over several processes (one per GPU).

The hot path has no exchange step: every (level, variable) slab reads only its
own field plus the pre-analysis winds of its level, DEWPOINT reads the RH flags of its own
level, and `qc_consistency` is per level -- so a time period is cut into contiguous level
ranges `[k_begin, k_end)` (the unit `og_qc_time_levels` / `og_oa_time` take), each rank runs its
range, and the analysed levels and flags are gathered to the rank that owns the host writer.
The surface pseudo-level carries every report (upper levels only the soundings), so ranges are
balanced on a cost estimate, not on level counts.

`run_time_sharded` is backend-agnostic: the compute callables are the GPU context's
`proc_qc` / `proc_oa` in production and anything with the same signature in the CPU tests
(`gloo`, world size 2, tests/test_shard.py).
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

from ._capi import OG_MISSING_R

FIELDS3 = ("t", "u", "v", "rh")
FLAGS3 = ("qc_u", "qc_v", "qc_t", "qc_rh")


# Relative cost of the two kernels that dominate a level, calibrated on B200 (profiles/exp_levels.py,
# bench breakdowns of cfg2 / cfg3): a Cressman (grid point, ob) update costs ~4.3e-9 ms, a buddy
# pair examined by the binned kernel ~5e-10 ms, and every level carries ~0.05 ms of fixed work
# (transposes, selection, list building, launches).
PAIR_PER_UPDATE = 0.12
LEVEL_FIXED_UPDATES = 1.2e7


def level_costs(case, oa_only: bool = False) -> np.ndarray:
    """Relative cost of each level, in units of one Cressman update: obs x Cressman reach, plus
    the pairs the binned buddy check examines (candidates within a 3 x 3 block of range-wide
    cells), plus the per-level fixed work.  oa_only: without the buddy pairs (the plan of the OA
    phase when the QC phase is cut inside the slabs, og_set_exchange)."""
    radii = [int(r) for r in case.oa_params.radius_influence if r >= 0]
    r2 = float(sum(r * r for r in radii)) or 1.0
    cost = np.zeros(case.kbu, np.float64)
    area = float(case.iew * case.jns)
    for k in range(case.kbu):
        n = 0
        for nm in ("ob_u", "ob_v", "ob_t", "ob_rh"):
            n += int(np.count_nonzero(np.abs(getattr(case, nm)[k] - OG_MISSING_R) > 1.0))
        if k == 0:
            n += int(np.count_nonzero(np.abs(case.ob_slp - OG_MISSING_R) > 1.0))
        p = float(case.pressure[k])
        rng_km = 650.0 if p <= 201.0 else (300.0 if p <= 1000.1 else 500.0)
        rc = rng_km / float(case.dx_km)
        per_var = n / 4.0
        pairs = 4.0 * per_var * per_var * min(9.0 * rc * rc / area, 1.0)
        cost[k] = n * np.pi * r2 + (0.0 if oa_only else PAIR_PER_UPDATE * pairs) + LEVEL_FIXED_UPDATES
    return cost


def _bottleneck(c: np.ndarray, parts: int) -> float:
    """Smallest possible heaviest-part cost when c is cut into `parts` contiguous parts."""
    K = len(c)
    parts = min(parts, K)
    pre = np.concatenate([[0.0], np.cumsum(c)])
    best = np.full((parts + 1, K + 1), np.inf)
    best[0, 0] = 0.0
    for p in range(1, parts + 1):
        for k in range(p, K + 1):
            for j in range(p - 1, k):
                v = max(best[p - 1, j], pre[k] - pre[j])
                if v < best[p, k]:
                    best[p, k] = v
    return float(best[parts, K])


def plan_levels(costs: Sequence[float], world: int) -> list[tuple[int, int]]:
    """Contiguous level ranges, one per rank.  The heaviest range is minimal (linear-partition DP),
    and so is, in turn, the heaviest of the ranges after the first, and so on: a rank that cannot
    be helped (the surface level on its own) does not leave the others lumped together.  Ranks
    beyond the number of levels get empty ranges."""
    c = np.asarray(costs, np.float64)
    K = len(c)
    world = max(int(world), 1)
    parts = min(world, K)
    ranges, start = [], 0
    for r in range(parts):
        left = parts - r
        if left == 1:
            end = K
        else:
            limit = _bottleneck(c[start:], left) * (1.0 + 1e-12)
            end, acc = start, 0.0
            # as many levels as fit under the optimal bottleneck, leaving one per remaining rank
            while end < K - (left - 1) and (end == start or acc + c[end] <= limit):
                acc += c[end]
                end += 1
        ranges.append((start, end))
        start = end
    ranges += [(K, K)] * (world - parts)
    return ranges


def _gather_rows(dist, arr: np.ndarray, ranges, rank: int, dst: int, device=None) -> None:
    """Gather the level rows [k0,k1) each rank owns into `arr` on rank `dst` (in place)."""
    import torch

    world = len(ranges)
    rows = max(k1 - k0 for k0, k1 in ranges)
    if rows == 0:
        return
    k0, k1 = ranges[rank]
    per = int(np.prod(arr.shape[1:])) if arr.ndim > 1 else 1
    buf = np.zeros((rows, per), arr.dtype)
    buf[: k1 - k0] = arr[k0:k1].reshape(k1 - k0, per)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, out, dst=dst)
    if rank == dst:
        for r, (a, b) in enumerate(ranges):
            if b > a and r != dst:
                arr[a:b] = out[r][: b - a].cpu().numpy().reshape((b - a,) + arr.shape[1:])


def run_time_sharded(case, rank: int, world: int, proc_qc: Callable, proc_oa: Callable,
                     dist=None, dst: int = 0, device=None, ranges=None):
    """QC then OA of one time period, levels sharded over `world` ranks.

    proc_qc(case, k_begin, k_end, run_consistency) and proc_oa(case, k_begin, k_end) update
    `case` in place for their level range.  Between the two stages nothing is exchanged: OA of a
    level reads only that level's flags.  Returns the ranges; on rank `dst` `case` then holds
    every level (fields and flags), elsewhere only the rank's own range is valid.
    """
    ranges = plan_levels(level_costs(case), world) if ranges is None else ranges
    k0, k1 = ranges[rank]
    if k1 > k0:
        proc_qc(case, k0, k1, True)
        proc_oa(case, k0, k1)
    if world > 1:
        assert dist is not None, "world > 1 needs torch.distributed"
        for nm in FIELDS3 + FLAGS3:
            _gather_rows(dist, getattr(case, nm), ranges, rank, dst, device)
        # slp / qc_slp belong to the surface level: ship them from its owner
        import torch
        owner = next(r for r, (a, b) in enumerate(ranges) if a <= 0 < b)
        if owner != dst:
            for nm in ("slp", "qc_slp"):
                a = getattr(case, nm)
                t = torch.from_numpy(a)
                if device is not None:
                    t = t.to(device)
                if rank == owner:
                    dist.send(t, dst=dst)
                elif rank == dst:
                    dist.recv(t, src=owner)
                    a[...] = t.cpu().numpy()
    return ranges


# ---- level x variable sharding --------------------------------------------------------------
# The surface pseudo-level carries every report and is the one level worth more than a rank's
# fair share (a third of cfg3, a quarter of cfg2), so its slabs -- UU VV TT RH PMSL for QC and
# OA, DEWPOINT for QC only -- are dealt out on their own; the upper levels stay whole and
# contiguous.  The price is one exchange step: qc_consistency couples the flags of U and V and of
# T and RH (qc0_module.F90:371-404), so between the QC and the OA phase the ranks combine the
# surface flag rows and run the consistency pass on the merged rows.  Every rank starts from the
# same incoming flags and QC only sets bits (qc0_module.F90:27-31, :55-72), so the merge is a
# bitwise OR -- RH and DEWPOINT, which share a flag row, may sit on different ranks.
SFC_VARS = (("u", 1 << 0), ("v", 1 << 1), ("t", 1 << 2), ("rh", 1 << 3), ("slp", 1 << 4), ("dew", 1 << 6))
ANISO_COST = 1.6          # a UU / VV / RH update against a TT / PMSL one (measured, profiles/exp_levels.py)
SFC_FLAGS = ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp")
QC_FLAG_BITS = (14, 16, 17)   # no_buddies, fails_error_max, fails_buddy_check (missing.inc:21-43)


def allreduce_or(dist, t):
    """In-place bitwise OR of an int32 tensor over the ranks.  gloo has the operator; NCCL does not,
    so there the three bits QC can set travel as 0/1 planes under MAX (every other bit is the
    incoming flag, identical on all ranks)."""
    import torch

    if dist.get_backend() == "gloo":
        dist.all_reduce(t, op=dist.ReduceOp.BOR)
        return t
    planes = torch.stack([(t >> b) & 1 for b in QC_FLAG_BITS]).to(torch.int8)
    dist.all_reduce(planes, op=dist.ReduceOp.MAX)
    for i, b in enumerate(QC_FLAG_BITS):
        t |= planes[i].to(t.dtype) << b
    return t


def surface_unit_costs(case, oa_only: bool = False) -> list[float]:
    """Cost of each surface unit of SFC_VARS, same unit as level_costs."""
    radii = [int(r) for r in case.oa_params.radius_influence if r >= 0]
    r2 = float(sum(r * r for r in radii)) or 1.0
    area = float(case.iew * case.jns)
    rc = 500.0 / float(case.dx_km)
    out = []
    for nm, _ in SFC_VARS:
        if nm == "dew":     # QC only; obs that carry both T and RH
            ok = (np.abs(case.ob_rh[0] - OG_MISSING_R) > 1.0) & (np.abs(case.ob_t[0] - OG_MISSING_R) > 1.0)
            n = float(np.count_nonzero(ok))
        else:
            ob = case.ob_slp if nm == "slp" else getattr(case, "ob_" + nm)[0]
            n = float(np.count_nonzero(np.abs(ob - OG_MISSING_R) > 1.0))
        pairs = n * n * min(9.0 * rc * rc / area, 1.0)
        shape = ANISO_COST if nm in ("u", "v", "rh") else 1.0
        oa = 0.0 if nm == "dew" else n * np.pi * r2 * shape
        if oa_only:
            out.append(0.0 if nm == "dew" else oa + LEVEL_FIXED_UPDATES / 6.0)
        else:
            out.append(oa + PAIR_PER_UPDATE * pairs + LEVEL_FIXED_UPDATES / 6.0)
    return out


def plan_slabs(case, world: int, oa_only: bool = False, sfc_qc_shared: bool = False) -> list[dict]:
    """One entry per rank: {'sfc_mask': variable bits (bit v-1 of variable v; bit 6 = DEWPOINT, QC
    only) of the surface slabs it owns (0: none), 'k0', 'k1': its contiguous range of upper
    levels (k0 >= 1)}."""
    world = max(int(world), 1)
    K = case.kbu
    if world == 1:
        return [{"sfc_mask": sum(b for _, b in SFC_VARS), "k0": 1, "k1": K}]
    up = level_costs(case, oa_only)[1:]
    loads = [0.0] * world
    masks = [0] * world
    # sfc_qc_shared: the surface level's QC is cut inside its slabs over all ranks (og_set_exchange),
    # an equal share for everyone -- only its OA slabs are dealt here; upper levels keep QC + OA
    units = sorted(zip(surface_unit_costs(case, oa_only or sfc_qc_shared), (b for _, b in SFC_VARS)), reverse=True)
    for cost, bits in units:                        # longest processing time first
        r = int(np.argmin(loads))
        loads[r] += cost
        masks[r] |= bits
    # upper levels: contiguous runs, lightest rank first, each filled up to the mean of what is left
    order = sorted(range(world), key=lambda r: loads[r])
    ranges = {r: (K, K) for r in range(world)}
    k = 1
    for pos, r in enumerate(order):
        rest = order[pos:]
        remaining = float(up[k - 1:].sum())
        target = (remaining + sum(loads[q] for q in rest)) / len(rest)
        k0 = k
        if pos == world - 1:
            k = K
        else:
            while k < K and loads[r] + up[k - 1] <= target * 1.02:
                loads[r] += up[k - 1]
                k += 1
        ranges[r] = (k0, k)
    return [{"sfc_mask": masks[r], "k0": ranges[r][0], "k1": ranges[r][1]} for r in range(world)]


def run_time_sharded_slabs(case, rank: int, world: int, proc_qc: Callable, proc_oa: Callable,
                           consistency: Callable, dist=None, dst: int = 0, device=None, plan=None):
    """QC then OA of one time period sharded by (level, variable).

    proc_qc(case, k0, k1, run_consistency, var_mask), proc_oa(case, k0, k1, var_mask) and
    consistency(case, k0, k1) update `case` in place.  Phases: QC of the rank's surface slabs
    (no consistency) and of its upper levels (with); exchange of the surface flag rows (OR);
    consistency of the surface level; OA of the same slabs; gather to `dst`.
    """
    import torch

    plan = plan_slabs(case, world) if plan is None else plan
    me = plan[rank]
    if me["sfc_mask"]:
        proc_qc(case, 0, 1, False, me["sfc_mask"])
    if me["k1"] > me["k0"]:
        proc_qc(case, me["k0"], me["k1"], True, 0)
    if world > 1:
        assert dist is not None, "world > 1 needs torch.distributed"
        rows = np.stack([case.qc_u[0], case.qc_v[0], case.qc_t[0], case.qc_rh[0], case.qc_slp])
        t = torch.from_numpy(rows)
        if device is not None:
            t = t.to(device)
        allreduce_or(dist, t)
        rows = t.cpu().numpy()
        case.qc_u[0], case.qc_v[0], case.qc_t[0], case.qc_rh[0] = rows[0], rows[1], rows[2], rows[3]
        case.qc_slp[...] = rows[4]
    consistency(case, 0, 1)
    if me["sfc_mask"] & 0x1f:       # DEWPOINT (bit 6) has no analysis
        proc_oa(case, 0, 1, me["sfc_mask"] & 0x1f)
    if me["k1"] > me["k0"]:
        proc_oa(case, me["k0"], me["k1"], 0)
    if world > 1:
        ranges = [(p["k0"], p["k1"]) for p in plan]
        for nm in FIELDS3 + FLAGS3:
            _gather_rows(dist, getattr(case, nm), ranges, rank, dst, device)
        # surface fields: each comes from the rank that analysed it
        for nm, bits in SFC_VARS:
            if not bits & 0x1f:
                continue
            owner = next(r for r, p in enumerate(plan) if p["sfc_mask"] & bits)
            if owner == dst:
                continue
            a = case.slp if nm == "slp" else getattr(case, nm)[0]
            t = torch.from_numpy(np.ascontiguousarray(a))
            if device is not None:
                t = t.to(device)
            if rank == owner:
                dist.send(t, dst=dst)
            elif rank == dst:
                dist.recv(t, src=owner)
                a[...] = t.cpu().numpy()
    return plan


def plan_units_measured(sfc_costs: dict, upper_costs: Sequence[float], world: int) -> list[dict]:
    """Plan from MEASURED unit costs (milliseconds of the previous analysis time, or of a calibration
    pass): sfc_costs maps a variable bit of SFC_VARS to the cost of that surface slab's unit,
    upper_costs[k - 1] is the cost of upper level k.  Longest processing time first; a rank's upper
    levels need not be contiguous (og_dev_*_time_items take any increasing list).
    Returns per rank {'sfc_mask', 'levels'}."""
    world = max(int(world), 1)
    units = [(float(c), 0, int(b)) for b, c in sfc_costs.items()] + \
            [(float(c), 1, k + 1) for k, c in enumerate(upper_costs)]
    units.sort(key=lambda u: (-u[0], u[1], u[2]))
    loads = [0.0] * world
    out = [{"sfc_mask": 0, "levels": []} for _ in range(world)]
    for cost, kind, ident in units:
        r = min(range(world), key=lambda q: (loads[q], q))
        loads[r] += cost
        if kind == 0:
            out[r]["sfc_mask"] |= ident
        else:
            out[r]["levels"].append(ident)
    for o, l in zip(out, loads):
        o["levels"].sort()
        o["load"] = l
    return out


# ---- inside one slab: the buddy check's cells over the ranks ---------------------------------
# The buddy check of a dense surface network (cfg4: 2 M reports, ~27 of the 38 ms of its time
# period) is one slab, so level / variable sharding cannot cut it.  The library deals its cells to
# the ranks and combines the relaxation mask of every fix-point pass and the final flags through an
# exchange callback (include/obsgrid_gpu.h: og_set_exchange); this is that callback over
# torch.distributed.  The all-reduce is queued on the library's stream, so the call stays
# stream-ordered: nothing synchronises with the host here.
class _RawBuffer:
    """A device (or host) buffer the library hands to the callback, as something torch can wrap."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3}


def make_exchange(dist, torch, device=None, group=None, counter=None):
    """An og_exchange_fn (ctypes callback object; keep it referenced while the context uses it).

    device None: the buffers are host memory (the CPU tests drive the callback directly with
    numpy buffers over gloo); otherwise a torch cuda device, the buffers are device pointers.
    counter: optional dict, gets 'calls' and 'bytes' incremented (bench.py reports them)."""
    import ctypes as C

    from ._capi import OG_EXCHANGE_FN, OG_XCHG_MAX, OG_XCHG_UINT8

    def fn(_user, buf, count, dtype, op, stream):
        try:
            tdt, npdt, ts, size = ((torch.uint8, np.uint8, "|u1", 1) if dtype == OG_XCHG_UINT8
                                   else (torch.int32, np.int32, "<i4", 4))
            rop = dist.ReduceOp.MAX if op == OG_XCHG_MAX else dist.ReduceOp.SUM
            if counter is not None:
                counter["calls"] = counter.get("calls", 0) + 1
                counter["bytes"] = counter.get("bytes", 0) + count * size
            if device is None:
                a = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint8 if size == 1 else C.c_int32)), shape=(count,))
                dist.all_reduce(torch.from_numpy(a), op=rop, group=group)
                return 0
            t = torch.as_tensor(_RawBuffer(buf, count, ts), device=device)
            with torch.cuda.stream(torch.cuda.ExternalStream(stream or 0, device=device)):
                dist.all_reduce(t, op=rop, group=group)
            return 0
        except Exception as e:          # an exception must not unwind through the C frames
            import sys
            print(f"obsgrid_b200.shard exchange failed: {e!r}", file=sys.stderr)
            return 1

    return OG_EXCHANGE_FN(fn)
