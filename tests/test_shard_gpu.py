"""One slab over two real GPUs: the buddy check's cells dealt to the ranks, the relaxation mask and
the flags combined by NCCL all-reduces on the library's stream (og_set_exchange,
obsgrid_b200/shard.py:make_exchange).  Launches itself under torch.distributed.run; needs two GPUs."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

pytestmark = pytest.mark.gpu


def _case():
    from obsgrid_b200 import synth
    return synth.small_case(iew=120, jns=100, kbu=4, n_sfc=6000, n_snd=400, seed=9, radii=(6, 3))


def _rank_main(out):
    import torch
    import torch.distributed as dist
    from obsgrid_b200 import shard
    from obsgrid_b200.api import Context
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    base = _case()
    ref = base.copy()
    ctx.proc_qc_oa(ref)                                   # single GPU, no exchange
    counter = {}
    ctx.set_exchange(rank, world, shard.make_exchange(dist, torch, torch.device("cuda", local), counter=counter))
    ctx.reset_stats()
    got = base.copy()
    ctx.proc_qc_oa(got)
    ctx.sync()
    pairs = torch.tensor([ctx.stats()["buddy_pairs_examined"]], dtype=torch.int64, device=f"cuda:{local}")
    allp = [torch.zeros_like(pairs) for _ in range(world)]
    dist.all_gather(allp, pairs)
    bad = [nm for nm in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp", "t", "u", "v", "rh", "slp")
           if not np.array_equal(getattr(got, nm), getattr(ref, nm))]
    res = {"rank": rank, "bad": bad, "calls": counter.get("calls", 0), "pairs": [int(p.item()) for p in allp]}
    with open(f"{out}.{rank}", "w") as f:
        json.dump(res, f)
    ctx.set_exchange(0, 1, None)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpus_share_the_buddy_check_of_a_slab(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = str(tmp_path / "res")
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", __file__, out]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=500)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    for r in range(2):
        res = json.load(open(f"{out}.{r}"))
        assert res["bad"] == [], res
        assert res["calls"] > 0
        assert min(res["pairs"]) > 0 and max(res["pairs"]) < sum(res["pairs"])


if __name__ == "__main__":
    sys.path.insert(0, str(ROOT))
    _rank_main(sys.argv[1])
