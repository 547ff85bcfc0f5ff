"""Where does the pipelined e2e step spend its time?  (experiment; not a bench)"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from obsgrid_b200 import synth
from obsgrid_b200.api import Context
ALL = ("t","u","v","rh","slp","qc_u","qc_v","qc_t","qc_rh","qc_slp","xob","yob","lon","elev","ob_u","ob_v","ob_t","ob_rh","ob_slp")
ctx = Context(0)
case = synth.make_case(2)
def pinned(c):
    o = c.copy()
    for nm in ALL: setattr(o, nm, torch.from_numpy(getattr(c, nm).copy()).pin_memory().numpy())
    return o
h = pinned(case); outs = [pinned(case) for _ in range(3)]
st = torch.cuda.ExternalStream(ctx.stream())
def run(n, up=True, down=True, NS=3):
    if up: ctx.time_upload(h, 0)
    ev = []
    for i in range(n):
        s = i % NS
        if up and i + 1 < n: ctx.time_upload(h, (i + 1) % NS)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record(st); ctx.time_compute(h, s); b.record(st)
        t1 = time.perf_counter()
        if down: ctx.time_download(outs[s], s)
        ev.append((a, b, t1 - t0))
    for s in range(NS): ctx.time_wait(s)
    ctx.sync()
    return np.mean([a.elapsed_time(b) for a, b, _ in ev[1:]]), 1e3 * np.mean([t for _, _, t in ev[1:]])
run(3)
for up, down in ((False, False), (True, False), (False, True), (True, True)):
    if not up:
        for s in range(3): ctx.time_upload(h, s)
        ctx.sync(); torch.cuda.synchronize()
    t0 = time.perf_counter(); dev_ms, host_ms = run(8, up, down); wall = (time.perf_counter() - t0) * 1e3 / 8
    print(f"upload={up} download={down}: wall/step {wall:.2f} ms, compute (events) {dev_ms:.2f} ms, compute call host time {host_ms:.2f} ms")
