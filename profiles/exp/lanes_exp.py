"""Experiment: L independent contexts ("lanes"), each pipelining its own share of analysis times
through og_time_upload/_compute/_download from its own host thread.  Measures e2e ms per time."""
import sys, time, threading, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from obsgrid_b200.api import Context
from obsgrid_b200 import synth
import bench

class A: pass
args = A(); args.config = int(sys.argv[1]) if len(sys.argv) > 1 else 2; args.scale_rh = False; args.scaling = "weak"
case = bench.build_case(args, 0)
NAMES = bench.FIELD_NAMES + bench.FLAG_NAMES + bench.TABLE_NAMES
K = case.kbu
def pinned_copy():
    oc = case.copy()
    for nm in NAMES:
        setattr(oc, nm, torch.from_numpy(getattr(case, nm).copy()).pin_memory().numpy())
    return oc
NT = int(sys.argv[2]) if len(sys.argv) > 2 else 16
for L in (1, 2, 3):
    NS = 2 if L > 1 else 3
    ctxs = [Context(0) for _ in range(L)]
    hin = pinned_copy()
    outs = [[pinned_copy() for _ in range(NS)] for _ in range(L)]
    def lane(l, nsteps):
        ctx = ctxs[l]
        ctx.time_upload(hin, 0, 0, K)
        for i in range(nsteps):
            s_ = i % NS
            if i + 1 < nsteps:
                ctx.time_upload(hin, (i + 1) % NS, 0, K)
            ctx.time_compute(hin, s_, 0, K)
            ctx.time_download(outs[l][s_], s_, 0, K)
        for s_ in range(NS):
            ctx.time_wait(s_)
        ctx.sync()
    def run(nt):
        th = [threading.Thread(target=lane, args=(l, nt // L)) for l in range(L)]
        for t in th: t.start()
        for t in th: t.join()
    run(L * (NS + 1))
    torch.cuda.synchronize()
    t0 = time.perf_counter(); run(NT - NT % L); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / (NT - NT % L)
    ref = outs[0][0]
    same = all(np.array_equal(getattr(outs[l][s], nm), getattr(ref, nm)) for l in range(L) for s in range(NS) for nm in bench.FIELD_NAMES + bench.FLAG_NAMES)
    print(f"cfg{args.config} lanes={L} slots/lane={NS} times={NT - NT % L}: {ms:.3f} ms per time  identical_outputs={same}", flush=True)
    for c in ctxs: c.close()
