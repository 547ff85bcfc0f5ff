"""Exact vs contracted arithmetic of the Cressman scan kernels: scan-kernel time and field deviation.
(experiment; bench.py measures)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
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
K = 1 if case.oa_params.surface_only else case.kbu
ctx.dev_proc_qc(d, case.qc_params, 0, K, True); ctx.sync()
out = {}
for mode in ("exact", "contracted", "exact", "contracted"):
    ctx.set_arithmetic(mode)
    best = 1e9
    for it in range(4):
        torch.cuda.synchronize()
        for nm in FIELD_NAMES: wk[nm].copy_(pr[nm])
        torch.cuda.synchronize()
        ctx.set_timing(True); ctx.reset_stats()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        ls = torch.cuda.ExternalStream(ctx.stream(), device=dev)
        e0.record(ls); ctx.dev_proc_oa(d, case.oa_params, 0, K); e1.record(ls); ctx.sync()
        ms, _ = ctx.kernel_ms()
        best = min(best, ms); oa = e0.elapsed_time(e1)
    out[mode] = {nm: wk[nm].cpu().numpy().copy() for nm in FIELD_NAMES}
    print(f"cfg{cfg} {mode:10s}: scan kernels {best:.3f} ms (best of 4), proc_oa {oa:.3f} ms", flush=True)
for nm in FIELD_NAMES:
    a, b = out["exact"][nm].astype(np.float64), out["contracted"][nm].astype(np.float64)
    err = np.abs(a - b)
    tol = 1e-3 + 1e-5 * np.abs(a)
    print(f"  {nm:4s} max |diff| {err.max():.3e}  max diff/tol {np.max(err / tol):.3e}  differing {np.mean(err > 0):.3f}")
