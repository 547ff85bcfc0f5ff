"""GPU vs oracle on every BASELINE.json configuration (north star: "reproduces reference analyses
and QC flags on all 5 configs").

cfg1-cfg3: one whole analysis time, every level and variable.  cfg4: the whole surface-FDDA
analysis time (its 6 QC and 5 OA slabs).  cfg5 (4000 x 4000, 2 M reports, 50 levels): the surface
level and three upper levels with exactly the values the full case has there
(synth.make_case(levels=...)), UU / TT / RH (+ DEWPOINT in QC) -- all levels and variables are
independent slabs, so a sample of slabs is a sample of the full result.

The oracle runs its fast paths here (spatial bins for the buddy check, rows dealt to threads),
which tests/test_oracle_fast_paths.py holds bit for bit to the as-written loops.  Bars: every QC
flag bit-identical; t and slp bit-identical; u, v, rh within 1e-5 relative / 1e-3 absolute (banana
pairs: atan2), each side analysing from its own flags.
"""
import os
import time

import numpy as np
import pytest

from obsgrid_b200 import synth

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-3
FLAGS = ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp")


def _run_both(ctx, oracle, base, qc_mask=0, oa_mask=0, exact=("t", "slp"), tol=("u", "v", "rh")):
    threads = os.cpu_count() or 1
    cpu, gpu = base.copy(), base.copy()
    for c in (cpu, gpu):
        c.qc_params.var_mask = qc_mask
        c.oa_params.var_mask = oa_mask
    oracle.set_fast_paths(True, 1)
    try:
        t0 = time.perf_counter()
        oracle.qc_time(cpu, threads=threads)
        t1 = time.perf_counter()
        ctx.proc_qc(gpu)
        t2 = time.perf_counter()
        for nm in FLAGS:
            np.testing.assert_array_equal(getattr(cpu, nm), getattr(gpu, nm), err_msg=nm)
        nflag = {nm: int(np.count_nonzero(getattr(cpu, nm))) for nm in FLAGS}
        t3 = time.perf_counter()
        rc_c = oracle.oa_time(cpu, threads=threads)
        t4 = time.perf_counter()
        rc_g = ctx.proc_oa(gpu)
        t5 = time.perf_counter()
    finally:
        oracle.set_fast_paths(False, 1)
    assert rc_c == rc_g == 0
    for nm in exact:
        np.testing.assert_array_equal(getattr(cpu, nm), getattr(gpu, nm), err_msg=nm)
    worst = {}
    for nm in tol:
        a, b = getattr(gpu, nm), getattr(cpu, nm)
        np.testing.assert_allclose(a, b, rtol=RTOL, atol=ATOL, err_msg=nm)
        worst[nm] = float(np.max(np.abs(a.astype(np.float64) - b)))
    print(f"\n{base.label}: oracle qc {t1 - t0:.1f} s / oa {t4 - t3:.1f} s on {threads} threads; "
          f"gpu (host arrays, blocking) qc {t2 - t1:.2f} s / oa {t5 - t4:.2f} s; flagged {nflag}; "
          f"max |gpu - oracle| {worst}")
    return cpu, gpu, nflag


def test_cfg2_full_time(ctx, oracle):
    """BASELINE configs[1]: 500 x 500, 38 levels, 50 k reports -- the benchmark workload."""
    base = synth.make_case(2)
    cpu, gpu, nflag = _run_both(ctx, oracle, base)
    assert all(nflag[nm] > 0 for nm in FLAGS)
    assert np.any(cpu.t != base.t) and np.any(cpu.u[20] != base.u[20])


def test_cfg3_full_time(ctx, oracle):
    """BASELINE configs[2]: 1000 x 1000, 38 levels, 200 k reports, one of the 16 times."""
    base = synth.make_case(3)
    cpu, gpu, nflag = _run_both(ctx, oracle, base)
    assert all(nflag[nm] > 0 for nm in FLAGS)


def test_cfg3_another_time(ctx, oracle):
    """A later analysis time of cfg3 (time 7: other phase of the first guess, other stations),
    surface + 4 levels."""
    base = synth.make_case(3, time_index=7, levels=[0, 6, 14, 27, 37])
    _run_both(ctx, oracle, base)


def test_cfg4_surface_fdda_time(ctx, oracle):
    """BASELINE configs[3]: 2000 x 2000, 500 k surface reports, surface-only analysis."""
    base = synth.make_case(4, levels=[0])
    assert base.oa_params.surface_only == 1
    cpu, gpu, nflag = _run_both(ctx, oracle, base)
    assert nflag["qc_t"] > 0 and nflag["qc_slp"] > 0


def test_cfg5_sampled_slabs(ctx, oracle):
    """BASELINE configs[4]: 4000 x 4000, 2 M reports, scans 12/9/6/4/2.  Surface + 850 / 500 / 100 hPa;
    UU, TT, RH (+ DEWPOINT in QC)."""
    full_p = synth.pressure_levels(50)
    lev = [0] + [int(np.flatnonzero(full_p == p)[0]) for p in (850.0, 500.0, 100.0)]
    base = synth.make_case(5, levels=lev)
    assert base.nrep == 2_000_000 and base.iew == 4000
    qc_mask = (1 << 0) | (1 << 2) | (1 << 3) | (1 << 6)
    oa_mask = (1 << 0) | (1 << 2) | (1 << 3)
    cpu, gpu, nflag = _run_both(ctx, oracle, base, qc_mask, oa_mask, exact=("t",), tol=("u", "rh"))
    assert nflag["qc_u"] > 0 and nflag["qc_t"] > 0 and nflag["qc_rh"] > 0
    np.testing.assert_array_equal(gpu.v, base.v)          # masked out on both sides
    np.testing.assert_array_equal(cpu.v, base.v)
