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
