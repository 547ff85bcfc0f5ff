#!/usr/bin/env python
"""Where does the asynchronous per-time pipeline (og_time_upload/_compute/_download) spend a step?
Times, over N steps on cfg2: copies only (no compute), compute only (slots filled once), everything;
one context or two, dense or sparse observation rows.  Wall clock; nothing here is a bench value."""
import sys, time, threading
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
from obsgrid_b200 import synth
from obsgrid_b200.api import Context

case = synth.make_case(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
N = 20
pinned = lambda a: torch.from_numpy(np.ascontiguousarray(a).copy()).pin_memory().numpy()
NAMES = ("t", "u", "v", "rh", "slp", "qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp", "xob", "yob", "lon", "elev", "ob_u", "ob_v", "ob_t", "ob_rh", "ob_slp")
def pinned_case():
    c = case.copy()
    for nm in NAMES: setattr(c, nm, pinned(getattr(case, nm)))
    return c

def run(n_lanes, NS, sparse, do_copy, do_compute):
    lanes = [Context(0) for _ in range(n_lanes)]
    hc = pinned_case(); hcsr = synth.CsrObs(case, alloc=pinned)
    outs = [[pinned_case() for _ in range(NS)] for _ in range(n_lanes)]
    ocsr = [[synth.CsrObs(case, alloc=pinned) for _ in range(NS)] for _ in range(n_lanes)]
    def lane(l, nsteps):
        torch.cuda.set_device(0)
        c = lanes[l]
        up = (lambda s: c.time_upload_csr(hc, hcsr, s)) if sparse else (lambda s: c.time_upload(hc, s))
        down = (lambda s: c.time_download_csr(outs[l][s], ocsr[l][s], s)) if sparse else (lambda s: c.time_download(outs[l][s], s))
        if not do_copy:
            for s in range(NS): up(s)
            c.sync()
        else:
            up(0)
        for i in range(nsteps):
            s = i % NS
            if do_copy and i + 1 < nsteps: up((i + 1) % NS)
            if do_compute: c.time_compute(hc, s)
            elif do_copy: c.time_compute(hc, s, 0, 0)      # empty level range: only the events
            if do_copy: down(s)
        for s in range(NS): c.time_wait(s)
        c.sync()
    def go(nsteps):
        share = [nsteps // n_lanes + (1 if l < nsteps % n_lanes else 0) for l in range(n_lanes)]
        th = [threading.Thread(target=lane, args=(l, share[l])) for l in range(n_lanes)]
        [t.start() for t in th]; [t.join() for t in th]
    go(2 * NS * n_lanes)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); go(N); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / N
    for c in lanes: c.close()
    return ms

for n_lanes, NS in ((1, 2), (1, 3), (1, 4), (2, 2), (2, 3)):
    for sparse in (False, True):
        r = {k: run(n_lanes, NS, sparse, *v) for k, v in (("copies", (True, False)), ("compute", (False, True)), ("all", (True, True)))}
        print(f"lanes {n_lanes} slots {NS} {'sparse' if sparse else 'dense '}: " + "  ".join(f"{k} {v:6.2f} ms" for k, v in r.items()), flush=True)
