#!/usr/bin/env python
"""Lean driver for ncu: W warm-up steps + 1 step of the resident QC+OA path of one config.

    python profiles/prof_step.py --config 2 --warmup 1 [--oa-only]

Nothing is timed here (numbers taken under a profiler are never bench values); bench.py
measures.  Kernel launches per OA step: 2 + 2 binning + per scan (prep, expand, 2 scan kernels).
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

from obsgrid_b200 import synth  # noqa: E402
from obsgrid_b200.api import Context  # noqa: E402

FIELD_NAMES = ("t", "u", "v", "rh", "slp")
FLAG_NAMES = ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp")
TABLE_NAMES = ("xob", "yob", "lon", "elev", "ob_u", "ob_v", "ob_t", "ob_rh", "ob_slp")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--oa-only", action="store_true")
    ap.add_argument("--qc-only", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = Context(0)
    case = synth.make_case(args.config)
    pristine, work = {}, {}
    for nm in FIELD_NAMES + FLAG_NAMES + TABLE_NAMES:
        pristine[nm] = torch.from_numpy(getattr(case, nm)).to(dev)
        work[nm] = pristine[nm].clone() if nm in FIELD_NAMES + FLAG_NAMES else pristine[nm]
    d = case.desc()
    for nm in FIELD_NAMES + FLAG_NAMES + TABLE_NAMES:
        setattr(d, nm, work[nm].data_ptr())
    for it in range(args.warmup + 1):
        for nm in FIELD_NAMES + FLAG_NAMES:
            work[nm].copy_(pristine[nm])
        torch.cuda.synchronize(dev)
        if not args.oa_only:
            ctx.dev_proc_qc(d, case.qc_params, 0, case.kbu, True)
        if not args.qc_only:
            ctx.dev_proc_oa(d, case.oa_params, 0, case.kbu)
        ctx.sync()
    print("launches", ctx.stats()["kernel_launches"])
    ctx.close()


if __name__ == "__main__":
    main()
