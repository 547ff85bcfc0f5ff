"""Multi-process path on CPU: world size 2 over gloo.  The sharding logic is backend-agnostic;
here the compute callables are the CPU oracle (test infrastructure), on the B200 they are the
GPU context's proc_qc / proc_oa (tests/test_gpu_parity.py::test_time_level_sharding_is_exact)."""
import os
import socket

import numpy as np
import pytest

from obsgrid_b200 import shard, synth

FIELDS = ("t", "u", "v", "rh", "slp", "qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp")


def test_plan_levels_balances_the_heavy_surface():
    costs = [100.0] + [10.0] * 20
    r = shard.plan_levels(costs, 2)
    assert r[0][0] == 0 and r[-1][1] == 21 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    loads = [sum(costs[a:b]) for a, b in r]
    assert max(loads) <= 160.0                      # 300 total: surface + a few upper vs the rest
    assert shard.plan_levels(costs, 1) == [(0, 21)]
    # a rank that cannot be helped (the surface on its own) must not leave the others lumped
    r4 = shard.plan_levels(costs, 4)
    loads4 = [sum(costs[a:b]) for a, b in r4]
    assert r4[0] == (0, 1) and max(loads4[1:]) <= 70.0 and sum(loads4) == 300.0
    r8 = shard.plan_levels([1.0, 1.0, 1.0], 8)
    assert r8[:3] == [(0, 1), (1, 2), (2, 3)] and all(a == b for a, b in r8[3:])


def test_level_costs_surface_dominates():
    case = synth.small_case(iew=40, jns=36, kbu=5, n_sfc=150, n_snd=40, seed=3)
    c = shard.level_costs(case)
    assert c.shape == (5,) and c[0] > c[1:].max()
    # the part that scales with the observations (the fixed per-level work taken out) is several
    # times larger at the surface, which carries every report
    v = c - shard.LEVEL_FIXED_UPDATES
    assert v[0] > 2.0 * v[1:].max()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    from oracle import ogoracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = synth.small_case(iew=40, jns=36, kbu=6, n_sfc=150, n_snd=40, seed=11, radii=(5, 3))

    def qc(c, k0, k1, cons):
        ogoracle.qc_time(c, k0, k1, run_consistency=cons)

    def oa(c, k0, k1):
        ogoracle.oa_time(c, k0, k1)

    ranges = shard.run_time_sharded(case, rank, world, qc, oa, dist=dist, dst=0)
    if rank == 0:
        np.savez(tmp, ranges=np.array(ranges), **{n: getattr(case, n) for n in FIELDS})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_over_gloo_match_one_process(tmp_path, oracle):
    import torch.multiprocessing as mp

    out = str(tmp_path / "sharded.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    assert got["ranges"].shape == (2, 2) and got["ranges"][0, 0] == 0 and got["ranges"][1, 1] == 6
    ref = synth.small_case(iew=40, jns=36, kbu=6, n_sfc=150, n_snd=40, seed=11, radii=(5, 3))
    oracle.qc_time(ref)
    oracle.oa_time(ref)
    for n in FIELDS:
        assert np.array_equal(got[n], getattr(ref, n)), n


# ---- level x variable sharding ------------------------------------------------------------------
def _oracle_callables():
    from oracle import ogoracle

    def qc(c, k0, k1, cons, mask):
        c.qc_params.var_mask = mask
        ogoracle.qc_time(c, k0, k1, run_consistency=cons)
        c.qc_params.var_mask = 0

    def oa(c, k0, k1, mask):
        c.oa_params.var_mask = mask
        ogoracle.oa_time(c, k0, k1)
        c.oa_params.var_mask = 0

    return qc, oa, ogoracle.qc_consistency


def test_plan_slabs_covers_every_slab_once():
    case = synth.small_case(iew=40, jns=36, kbu=9, n_sfc=400, n_snd=60, seed=3)
    allbits = sum(b for _, b in shard.SFC_VARS)
    for world in (1, 2, 3, 4, 8, 16):
        plan = shard.plan_slabs(case, world)
        assert len(plan) == world
        seen = 0
        for p in plan:
            assert seen & p["sfc_mask"] == 0
            seen |= p["sfc_mask"]
        assert seen == allbits
        ks = sorted((p["k0"], p["k1"]) for p in plan if p["k1"] > p["k0"])
        assert ks[0][0] == 1 and ks[-1][1] == case.kbu and all(a[1] == b[0] for a, b in zip(ks, ks[1:]))
        if world >= 2:      # the surface is shared out: no rank holds all five of its slabs
            assert all(p["sfc_mask"] != allbits for p in plan)


def test_slab_sharding_sequential_equals_one_process(oracle):
    """The phases of run_time_sharded_slabs, executed rank after rank in one process on separate
    copies and merged as the collectives would, reproduce the unsharded time exactly."""
    base = synth.small_case(iew=40, jns=36, kbu=6, n_sfc=150, n_snd=40, seed=11, radii=(5, 3))
    ref = base.copy()
    oracle.qc_time(ref); oracle.oa_time(ref)
    qc, oa, cons = _oracle_callables()
    world = 7          # one more rank than surface units: RH and DEWPOINT land on different ranks
    plan = shard.plan_slabs(base, world)
    assert not any((p["sfc_mask"] & (1 << 3)) and (p["sfc_mask"] & (1 << 6)) for p in plan)
    cs = [base.copy() for _ in range(world)]
    for r, p in enumerate(plan):
        if p["sfc_mask"]:
            qc(cs[r], 0, 1, False, p["sfc_mask"])
        if p["k1"] > p["k0"]:
            qc(cs[r], p["k0"], p["k1"], True, 0)
    rows = [np.bitwise_or.reduce([getattr(c, nm)[0] if nm != "qc_slp" else c.qc_slp for c in cs], axis=0)
            for nm in shard.SFC_FLAGS]
    out = base.copy()
    for r, p in enumerate(plan):
        c = cs[r]
        c.qc_u[0], c.qc_v[0], c.qc_t[0], c.qc_rh[0] = rows[0], rows[1], rows[2], rows[3]
        c.qc_slp[...] = rows[4]
        cons(c, 0, 1)
        if p["sfc_mask"] & 0x1f:
            oa(c, 0, 1, p["sfc_mask"] & 0x1f)
        if p["k1"] > p["k0"]:
            oa(c, p["k0"], p["k1"], 0)
        for nm in ("t", "u", "v", "rh", "qc_u", "qc_v", "qc_t", "qc_rh"):
            getattr(out, nm)[p["k0"]:p["k1"]] = getattr(c, nm)[p["k0"]:p["k1"]]
        for nm, bits in shard.SFC_VARS:
            if p["sfc_mask"] & bits & 0x1f:
                if nm == "slp":
                    out.slp[...] = c.slp
                else:
                    getattr(out, nm)[0] = getattr(c, nm)[0]
    for i, nm in enumerate(("qc_u", "qc_v", "qc_t", "qc_rh")):
        getattr(out, nm)[0] = getattr(cs[0], nm)[0]
    out.qc_slp[...] = cs[0].qc_slp
    for n in FIELDS:
        assert np.array_equal(getattr(out, n), getattr(ref, n)), n


def _worker_slabs(rank, world, port, tmp):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = synth.small_case(iew=40, jns=36, kbu=6, n_sfc=150, n_snd=40, seed=11, radii=(5, 3))
    qc, oa, cons = _oracle_callables()
    plan = shard.run_time_sharded_slabs(case, rank, world, qc, oa, cons, dist=dist, dst=0)
    if rank == 0:
        np.savez(tmp, masks=np.array([p["sfc_mask"] for p in plan]), **{n: getattr(case, n) for n in FIELDS})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_slab_sharding_over_gloo(tmp_path, oracle):
    import torch.multiprocessing as mp

    out = str(tmp_path / "sharded_slabs.npz")
    mp.spawn(_worker_slabs, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    assert np.all(got["masks"] != 0)                    # both ranks own surface slabs
    ref = synth.small_case(iew=40, jns=36, kbu=6, n_sfc=150, n_snd=40, seed=11, radii=(5, 3))
    oracle.qc_time(ref)
    oracle.oa_time(ref)
    for n in FIELDS:
        assert np.array_equal(got[n], getattr(ref, n)), n


def test_allreduce_or_bit_planes_equal_bitwise_or():
    """NCCL has no OR reduction: the three bits QC can set travel as 0/1 planes under MAX.  A stub
    collective stands in for the second rank."""
    import torch

    rng = np.random.default_rng(8)
    q0 = rng.integers(0, 1 << 13, 5000).astype(np.int32) | (rng.integers(0, 2, 5000).astype(np.int32) << 18)
    bits = np.array([1 << b for b in shard.QC_FLAG_BITS], np.int32)
    a = q0 | np.bitwise_or.reduce(bits[:, None] * rng.integers(0, 2, (3, 5000)).astype(np.int32), axis=0)
    b = q0 | np.bitwise_or.reduce(bits[:, None] * rng.integers(0, 2, (3, 5000)).astype(np.int32), axis=0)

    class ReduceOp:
        MAX = "max"

    class Stub:
        def __init__(self, other):
            self.other, self.ReduceOp = other, ReduceOp

        def get_backend(self):
            return "nccl"

        def all_reduce(self, planes, op):
            assert op == "max" and planes.dtype == torch.int8
            o = torch.stack([(self.other >> k) & 1 for k in shard.QC_FLAG_BITS]).to(torch.int8)
            planes.copy_(torch.maximum(planes, o))

    t = torch.from_numpy(a.copy())
    shard.allreduce_or(Stub(torch.from_numpy(b)), t)
    np.testing.assert_array_equal(t.numpy(), a | b)


def _worker_exchange(rank, world, port, tmp):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from obsgrid_b200 import _capi

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counter = {}
    fn = shard.make_exchange(dist, torch, device=None, counter=counter)
    mask = np.zeros(1000, np.uint8); mask[rank::3] = 1
    flags = np.full(1000, 1 << 13, np.int32); flags[rank::2] += 1 << 17
    rc1 = fn(None, mask.ctypes.data, mask.size, _capi.OG_XCHG_UINT8, _capi.OG_XCHG_MAX, None)
    rc2 = fn(None, flags.ctypes.data, flags.size, _capi.OG_XCHG_INT32, _capi.OG_XCHG_MAX, None)
    np.savez(f"{tmp}.{rank}.npz", mask=mask, flags=flags, rc=np.array([rc1, rc2]), calls=counter["calls"], nbytes=counter["bytes"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_exchange_callback_over_gloo(tmp_path):
    """The og_exchange_fn of shard.make_exchange, driven the way the library drives it (raw pointer,
    count, dtype, op), on host buffers over gloo: in-place MAX over two ranks."""
    import torch.multiprocessing as mp

    out = str(tmp_path / "xchg")
    mp.spawn(_worker_exchange, args=(2, _free_port(), out), nprocs=2, join=True)
    want_mask = np.zeros(1000, np.uint8); want_mask[0::3] = 1; want_mask[1::3] = 1
    want_flags = np.full(1000, (1 << 13) + (1 << 17), np.int32)
    for r in range(2):
        got = np.load(f"{out}.{r}.npz")
        assert list(got["rc"]) == [0, 0]
        assert np.array_equal(got["mask"], want_mask) and np.array_equal(got["flags"], want_flags)
        assert int(got["calls"]) == 2 and int(got["nbytes"]) == 1000 + 4000


def test_plan_units_measured_deals_every_unit_once_and_balances():
    rng = np.random.default_rng(5)
    sfc = {b: float(c) for b, c in zip((1, 2, 4, 8, 16), rng.uniform(20, 45, 5))}
    up = rng.uniform(3, 12, 49)
    for world in (1, 2, 3, 8):
        plan = shard.plan_units_measured(sfc, up, world)
        assert len(plan) == world
        masks = [p["sfc_mask"] for p in plan]
        assert sum(masks) == 31 and all(a & b == 0 for i, a in enumerate(masks) for b in masks[i + 1:])
        levels = sorted(k for p in plan for k in p["levels"])
        assert levels == list(range(1, 50))
        assert all(p["levels"] == sorted(p["levels"]) for p in plan)
        loads = [p["load"] for p in plan]
        assert abs(sum(loads) - (sum(sfc.values()) + up.sum())) < 1e-9
        assert max(loads) - min(loads) <= max(max(sfc.values()), up.max())      # the LPT bound
