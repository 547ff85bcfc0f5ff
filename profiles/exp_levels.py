"""Scan-kernel time of the surface level vs the upper levels (experiment; not a bench)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from obsgrid_b200 import synth
from obsgrid_b200.api import Context
FIELD_NAMES = ("t", "u", "v", "rh", "slp"); FLAG_NAMES = ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp")
TABLE_NAMES = ("xob", "yob", "lon", "elev", "ob_u", "ob_v", "ob_t", "ob_rh", "ob_slp")
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0); ctx = Context(0); case = synth.make_case(cfg)
pr, wk = {}, {}
for nm in FIELD_NAMES + FLAG_NAMES + TABLE_NAMES:
    pr[nm] = torch.from_numpy(getattr(case, nm)).to(dev)
    wk[nm] = pr[nm].clone() if nm in FIELD_NAMES + FLAG_NAMES else pr[nm]
d = case.desc()
for nm in FIELD_NAMES + FLAG_NAMES + TABLE_NAMES: setattr(d, nm, wk[nm].data_ptr())
ctx.dev_proc_qc(d, case.qc_params, 0, case.kbu, True); ctx.sync()
flags = {nm: wk[nm].clone() for nm in FLAG_NAMES}
for (k0, k1, mask) in ((0, 1, 0), (1, case.kbu, 0), (0, case.kbu, 0), (0, 1, 0b00100 | 0b10000), (0, 1, 0b01011), (1, case.kbu, 0b00100), (1, case.kbu, 0b01011)):
    case.oa_params.var_mask = mask
    for it in range(3):
        torch.cuda.synchronize()
        for nm in FIELD_NAMES: wk[nm].copy_(pr[nm])
        torch.cuda.synchronize()
        ctx.set_timing(True); ctx.reset_stats(); ctx.set_counting(it == 0)
        ctx.dev_proc_oa(d, case.oa_params, k0, k1); ctx.sync()
        ms, _ = ctx.kernel_ms(); st = ctx.stats()
        if it == 0: upd, cand = st["cressman_updates"], st["cressman_candidates"]
    print(f"levels [{k0},{k1}) var_mask {mask:05b}: scan kernels {ms:.3f} ms, updates {upd:.3e}, candidates {cand:.3e}, {upd/ms/1e6:.1f} G upd/s")
