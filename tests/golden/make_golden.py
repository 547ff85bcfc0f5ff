#!/usr/bin/env python
"""Generate the golden input/output vectors under tests/golden/ from the CPU oracle.

    python tests/golden/make_golden.py

The reference ships no fixtures and cannot be built in this image (Fortran 90 + netCDF), so
these vectors are outputs of oracle/og_oracle.c -- the line-by-line restatement pinned by the
hand-derived known-answer tests in tests/test_oracle_kat.py -- frozen at generation time.  They
(a) pin the oracle itself against compiler / flag / refactoring drift (tests/test_golden.py,
CPU) and (b) let the GPU path be checked on a box where nothing under oracle/ is built
(tests/test_golden.py -m gpu).  Inputs are stored alongside the outputs, so the files do not
depend on any PRNG.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from obsgrid_b200 import synth  # noqa: E402
from oracle import ogoracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent
CASE_ARRAYS = ("pressure", "t", "u", "v", "rh", "slp", "xob", "yob", "lon", "elev", "ob_u", "ob_v",
               "ob_t", "ob_rh", "ob_slp", "qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp")


def wavy(iew, jns, a=10.0, b=280.0, seed=0):
    i = np.arange(iew)[None, :]; j = np.arange(jns)[:, None]
    return (b + a * np.sin(i / 7.0 + seed) * np.cos(j / 5.0)).astype(np.float32)


def wind_fields(iew, jns, jet=40.0, curl=True):
    i = np.arange(1, iew + 1, dtype=np.float64)[None, :]
    j = np.arange(1, jns + 1, dtype=np.float64)[:, None]
    yc = jns / 2 + 0.15 * jns * np.sin(2 * np.pi * i / iew * 1.5)
    core = np.exp(-((j - yc) / (0.12 * jns)) ** 2)
    u = 3.0 + jet * core
    v = jet * core * (0.15 * jns * 2 * np.pi * 1.5 / iew) * np.cos(2 * np.pi * i / iew * 1.5) if curl else 0 * u
    return u.astype(np.float32), (v + 0 * u).astype(np.float32)


def obs_xy(rng, n, iew, jns, lo=1.0):
    return (rng.uniform(lo, iew - 1.0, n).astype(np.float32),
            rng.uniform(lo, jns - 1.0, n).astype(np.float32))


def main():
    rng = np.random.default_rng(20261019)

    # 1. isotropic Cressman scan (TT), with and without first guess
    iew, jns, n, R = 64, 48, 300, 5
    g = wavy(iew, jns); x, y = obs_xy(rng, n, iew, jns); d = rng.normal(0, 3, n).astype(np.float32)
    np.savez_compressed(OUT / "cressman_circle.npz", g=g, x=x, y=y, d=d, var=3, R=R, dxd=12.0,
                        out=O.cressman(d, x, y, g, 3, R, 12.0),
                        out_nofg=O.cressman(d, x, y, g, 3, R, 12.0, use_first_guess=False))

    # 2. anisotropic scans: straight jet (circles + ellipses, bit-exact on the GPU) and
    #    meandering jet (bananas: atan2 -> tolerance), RH with the scale factor
    iew, jns, n, R = 72, 60, 500, 5
    g = wavy(iew, jns, 20.0, 50.0); x, y = obs_xy(rng, n, iew, jns)
    d = rng.normal(0, 5, n).astype(np.float32)
    fg = (np.abs(rng.normal(50, 20, n)) + 1).astype(np.float32)
    ub, vb = wind_fields(iew, jns, jet=45.0, curl=False)
    O.reset_stats()
    out_e = O.cressman(d, x, y, g, 1, R, 12.0, ub, vb, 300.0)
    st_e = O.get_stats()
    ub2, vb2 = wind_fields(iew, jns)
    O.reset_stats()
    out_b = O.cressman(d, x, y, g, 4, R, 12.0, ub2, vb2, 850.0, fg_value=fg, scale_rh=True)
    st_b = O.get_stats()
    assert st_e["n_banana"] == 0 and st_e["n_ellipse"] > 20 and st_b["n_banana"] > 0
    np.savez_compressed(OUT / "cressman_aniso.npz", g=g, x=x, y=y, d=d, fg=fg, R=R, dxd=12.0,
                        ub=ub, vb=vb, out_ellipse=out_e, ub2=ub2, vb2=vb2, out_banana=out_b,
                        shapes_ellipse=[st_e["n_circle"], st_e["n_ellipse"], st_e["n_banana"]],
                        shapes_banana=[st_b["n_circle"], st_b["n_ellipse"], st_b["n_banana"]])

    # 3. one slab of proc_oa: innovations, three scans, re-interpolated innovations between them
    iew, jns, n = 56, 44, 250
    g = wavy(iew, jns); x, y = obs_xy(rng, n, iew, jns, lo=2.0)
    ob = (280 + rng.normal(0, 4, n)).astype(np.float32)
    out, rc, ns = O.oa_slab(g, 3, 700.0, ob, x, y, (6, 4, 2), 12.0)
    np.savez_compressed(OUT / "oa_slab.npz", g=g, x=x, y=y, ob=ob, radii=[6, 4, 2], dxd=12.0,
                        pressure=700.0, out=out, rc=rc, nscan=ns)

    # 4. error_max + buddy_check of one slab (UU at 850 hPa, 9 km grid)
    iew, jns, n = 120, 90, 700
    g = wavy(iew, jns, 8.0, 10.0); x, y = obs_xy(rng, n, iew, jns, lo=2.0)
    ob = (10 + rng.normal(0, 3, n)).astype(np.float32)
    bad = rng.random(n) < 0.05
    ob[bad] += rng.choice([-1, 1], bad.sum()).astype(np.float32) * rng.uniform(15, 40, bad.sum()).astype(np.float32)
    elev = rng.uniform(0, 1500, n).astype(np.float32)
    lt = (rng.integers(0, 24, n) * 100).astype(np.int32)
    q0 = np.zeros(n, np.int32)
    q1, aob, err, thr = O.error_max(ob, x, y, elev, q0, g, 1, 10.0, 850.0, lt)
    q2, aob2, errb, dm, bnf = O.buddy_check(ob, x, y, q1, elev, g, 1, 850.0, 9.0, 1.0, 8.0)
    np.savez_compressed(OUT / "qc_slab.npz", g=g, x=x, y=y, ob=ob, elev=elev, lt=lt, var=1,
                        pressure=850.0, max_error=10.0, max_buddy=8.0, dx_km=9.0,
                        qc_after_error_max=q1, err_max=err, thr_max=thr,
                        qc_after_buddy=q2, err_buddy=errb, difmax=dm, buddy_num_final=bnf)

    # 5. one whole (small) analysis time: proc_qc then proc_oa over every level and variable
    case = synth.small_case(iew=40, jns=36, kbu=5, n_sfc=150, n_snd=40, seed=7, radii=(5, 3))
    inputs = {"in_" + n_: getattr(case, n_).copy() for n_ in CASE_ARRAYS}
    O.qc_time(case)
    flags = {n_: getattr(case, n_).copy() for n_ in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp")}
    O.oa_time(case)
    fields = {n_: getattr(case, n_).copy() for n_ in ("t", "u", "v", "rh", "slp")}
    np.savez_compressed(OUT / "time_small.npz", **inputs, **flags, **fields,
                        args=dict(iew=40, jns=36, kbu=5, n_sfc=150, n_snd=40, seed=7, radii=(5, 3)))
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)


def epilogue():
    """6. the rows widened in SURVEY 8f ranks 2-4: surface-FDDA QC with density / distance,
    temporal_interp, mixing_ratio + sfcprs, wind increments onto the C grid, surface PRES.
    Own generator: adding it leaves files 1-5 untouched (python make_golden.py --epilogue)."""
    rng = np.random.default_rng(20261020)
    kbu, iew, jns = 4, 30, 26
    mk = lambda lo, hi: rng.uniform(lo, hi, (kbu, iew, jns)).astype(np.float32)
    u, uA, uC, v, vA, vC = (mk(-30, 30) for _ in range(6))
    gu, gv = O.wind_increment_to_c_grid(u, uA, uC, v, vA, vC)
    pres = rng.uniform(60000, 103000, (iew, jns)).astype(np.float32)
    slpc = rng.uniform(98000, 104000, (iew, jns)).astype(np.float32)
    slpx = (slpc + rng.normal(0, 300, (iew, jns))).astype(np.float32)
    a = rng.normal(280, 5, (kbu, iew, jns)).astype(np.float32); b = rng.normal(282, 5, (kbu, iew, jns)).astype(np.float32)
    press = np.array([1001, 1000, 925, 850, 700, 500, 300], np.float32)
    k7 = len(press)
    t = np.empty((k7, iew, jns), np.float32); h = np.empty_like(t)
    for k, p in enumerate(press):
        t[k] = 288.0 * (min(p, 1000.0) / 1000.0) ** 0.19 + rng.normal(0, 1.5, (iew, jns))
        h[k] = 44330.0 * (1 - (min(p, 1000.0) / 1013.25) ** 0.1903) + rng.normal(0, 10, (iew, jns))
    rh = rng.uniform(-5, 110, (k7, iew, jns)).astype(np.float32)
    terrain = rng.uniform(0, 5200, (iew, jns)).astype(np.float32)
    slp = rng.uniform(99000, 103000, (iew, jns)).astype(np.float32)
    rc, rh1, qv, psfc = O.mixing_ratio(rh, t, h, terrain, slp, press)
    assert rc == 0
    fargs = dict(iew=60, jns=50, kbu=4, n_sfc=50, n_snd=8, seed=29, radii=(6, 4), dx_m=60000.0)
    case = synth.small_case(**fargs)
    inputs = {"in_" + n_: getattr(case, n_).copy() for n_ in CASE_ARRAYS}
    tb0 = rng.integers(0, 3, (case.iew, case.jns)).astype(np.float32)
    tb, od = O.qc_time_fdda(case, tb0)
    np.savez_compressed(OUT / "epilogue.npz", u=u, uA=uA, uC=uC, v=v, vA=vA, vC=vC, out_u=gu, out_v=gv,
                        pres=pres, slpc=slpc, slpx=slpx, out_pres=O.pres_sfc_from_slp(pres, slpc, slpx),
                        ta=a, tb=b, tinterp=[2, 0, 6], out_tinterp_cross=O.temporal_interp(a, b, 2, 0, 6, True)[1],
                        out_tinterp_full=O.temporal_interp(a, b, 2, 0, 6, False)[1],
                        press=press, t=t, h=h, rh=rh, terrain=terrain, slp=slp, out_rh=rh1, out_qv=qv, out_psfc=psfc,
                        fdda_args=fargs, fdda_tb0=tb0, fdda_tb=tb, fdda_od=od,
                        **{"fdda_" + k_: v_ for k_, v_ in inputs.items()},
                        **{"fdda_" + n_: getattr(case, n_).copy() for n_ in ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp")})
    f = OUT / "epilogue.npz"
    print(f.name, f.stat().st_size)


if __name__ == "__main__":
    epilogue() if "--epilogue" in sys.argv else (main(), epilogue())
