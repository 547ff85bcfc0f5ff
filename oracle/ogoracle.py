"""ctypes wrapper of the CPU oracle (oracle/og_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the obsgrid_b200 product package.
PARITY UNPINNED by the reference (no fixtures, Fortran unbuildable here); see og_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from obsgrid_b200._capi import (OgOaParams, OgQcParams, OgTimeDesc, OG_MAX_SCAN, f32, i32, ptr)

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libogoracle.so"
_vp = C.c_void_p


class OgoStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("cressman_updates", "cressman_candidates",
                                         "buddy_pair_tests", "n_circle", "n_ellipse", "n_banana",
                                         "buddy_pairs_examined")]


def build(force: bool = False) -> Path:
    src = [HERE / "og_oracle.c", HERE / "og_oracle_reports.c", HERE / "og_oracle.h",
           HERE.parent / "include" / "obsgrid_gpu.h", HERE.parent / "include" / "obsgrid_reports.h"]
    if force or not LIB.exists() or any(s.stat().st_mtime > LIB.stat().st_mtime for s in src):
        subprocess.check_call(["make", "-C", str(HERE), "-s"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(str(LIB))
        l.ogo_dewpoint.restype = C.c_float
        l.ogo_dewpoint.argtypes = [C.c_float, C.c_float]
        l.ogo_fg_t_at_point.restype = C.c_float
        l.ogo_fg_t_at_point.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_float,
                                        C.c_float, C.c_float]
        l.ogo_ob_density.argtypes = [_vp, _vp, C.c_float, C.c_int, _vp, C.c_int, C.c_int]
        l.ogo_obs_distance.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_float]
        l.ogo_innovation.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_float,
                                     C.c_int, _vp, _vp]
        l.ogo_cressman.argtypes = [_vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_float, _vp, _vp, C.c_float, C.c_int,
                                   C.c_int, C.c_int, _vp, C.c_int]
        l.ogo_oa_slab.restype = C.c_int
        l.ogo_oa_slab.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_float, _vp, _vp, _vp,
                                  C.c_int, C.c_float, _vp, C.c_float, C.c_float, _vp, _vp,
                                  C.c_int, C.c_int, C.c_int, C.c_int, _vp]
        l.ogo_smooth_5.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int]
        l.ogo_smoother_desmoother.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int]
        l.ogo_clean_rh.argtypes = [_vp, C.c_int, C.c_int, C.c_float, C.c_float]
        l.ogo_modify_qc_flag.argtypes = [_vp, C.c_int, _vp, _vp, C.c_int]
        l.ogo_add_to_qc_flag.argtypes = [_vp, C.c_int]
        l.ogo_contains_2n.restype = C.c_int
        l.ogo_contains_2n.argtypes = [C.c_int, C.c_int]
        l.ogo_local_time.argtypes = [C.c_int, _vp, C.c_int, _vp]
        l.ogo_error_max.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int,
                                    C.c_int, C.c_float, C.c_float, _vp, _vp, _vp, _vp]
        l.ogo_buddy_check.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int,
                                      C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _vp,
                                      _vp, _vp, _vp]
        for name in ("ogo_qc_time",):
            getattr(l, name).restype = C.c_int
        l.ogo_qc_time.argtypes = [C.POINTER(OgTimeDesc), C.POINTER(OgQcParams), C.c_int, C.c_int,
                                  C.c_int]
        l.ogo_oa_time.restype = C.c_int
        l.ogo_oa_time.argtypes = [C.POINTER(OgTimeDesc), C.POINTER(OgOaParams), C.c_int, C.c_int]
        l.ogo_oa_time_mt.restype = C.c_int
        l.ogo_oa_time_mt.argtypes = [C.POINTER(OgTimeDesc), C.POINTER(OgOaParams), C.c_int,
                                     C.c_int, C.c_int]
        l.ogo_qc_time_mt.restype = C.c_int
        l.ogo_qc_time_mt.argtypes = [C.POINTER(OgTimeDesc), C.POINTER(OgQcParams), C.c_int,
                                     C.c_int, C.c_int, C.c_int]
        l.ogo_get_stats.argtypes = [C.POINTER(OgoStats)]
        _lib = l
    return _lib


def set_fast_paths(buddy_binned: bool = False, inner_threads: int = 1):
    """Bit-identical re-schedulings of the as-written loops for the large configurations."""
    lib().ogo_set_fast_paths(int(buddy_binned), int(inner_threads))


def reset_stats():
    lib().ogo_reset_stats()


def get_stats() -> dict:
    s = OgoStats()
    lib().ogo_get_stats(C.byref(s))
    return {n: getattr(s, n) for n, _ in s._fields_}


# ---- array-level wrappers; 2-D slabs are numpy arrays of shape (jns, iew) (C order), i.e.
# ---- exactly the memory layout of Fortran's gridded(iew,jns).
def innovation(gridded, xob, yob, obs_value, scale=1.0, use_first_guess=True):
    g = f32(gridded); jns, iew = g.shape
    xob, yob, obs_value = f32(xob), f32(yob), f32(obs_value)
    n = xob.size
    diff = np.zeros(n, np.float32); fg = np.full(n, -9999999.0, np.float32)
    lib().ogo_innovation(ptr(g), iew, jns, ptr(xob), ptr(yob), ptr(obs_value), n, scale,
                         int(use_first_guess), ptr(diff), ptr(fg))
    return diff, fg


def cressman(diff, xob, yob, gridded, var, radius, dxd_km, u_banana=None, v_banana=None,
             pressure=1001.0, passes=0, smooth_type=1, use_first_guess=True, fg_value=None,
             scale_rh=False, crsdot=1):
    g = f32(gridded).copy(); jns, iew = g.shape
    diff, xob, yob = f32(diff), f32(xob), f32(yob)
    ub = f32(u_banana) if u_banana is not None else np.zeros_like(g)
    vb = f32(v_banana) if v_banana is not None else np.zeros_like(g)
    fg = f32(fg_value) if fg_value is not None else np.zeros(diff.size, np.float32)
    lib().ogo_cressman(ptr(diff), ptr(xob), ptr(yob), diff.size, ptr(g), iew, jns, crsdot, var,
                       int(radius), dxd_km, ptr(ub), ptr(vb), pressure, passes, smooth_type,
                       int(use_first_guess), ptr(fg), int(scale_rh))
    return g


def oa_slab(gridded, var, pressure, obs_value, xob, yob, radii, dxd_km, u_banana=None,
            v_banana=None, scale=1.0, sfc_mult=1.0, passes=0, smooth_type=1,
            use_first_guess=True, scale_rh=False):
    g = f32(gridded).copy(); jns, iew = g.shape
    obs_value, xob, yob = f32(obs_value), f32(xob), f32(yob)
    ub = f32(u_banana) if u_banana is not None else np.zeros_like(g)
    vb = f32(v_banana) if v_banana is not None else np.zeros_like(g)
    rad = np.full(OG_MAX_SCAN, -1, np.int32); rad[:len(radii)] = radii
    ns = C.c_int(0)
    rc = lib().ogo_oa_slab(ptr(g), iew, jns, var, pressure, ptr(obs_value), ptr(xob), ptr(yob),
                           xob.size, scale, ptr(rad), sfc_mult, dxd_km, ptr(ub), ptr(vb), passes,
                           smooth_type, int(use_first_guess), int(scale_rh), C.byref(ns))
    return g, rc, ns.value


def local_time(time_hhmm, lonob):
    lon = f32(lonob); out = np.zeros(lon.size, np.int32)
    lib().ogo_local_time(int(time_hhmm), ptr(lon), lon.size, ptr(out))
    return out


def error_max(obs, xob, yob, elev, qc_flag, gridded, var, max_difference, pressure, local_t):
    g = f32(gridded); jns, iew = g.shape
    obs, xob, yob, elev = f32(obs), f32(xob), f32(yob), f32(elev)
    qc = i32(qc_flag).copy(); lt = i32(local_t); n = obs.size
    aob = np.zeros(n, np.float32); err = np.zeros(n, np.float32); thr = np.zeros(n, np.float32)
    lib().ogo_error_max(ptr(obs), ptr(xob), ptr(yob), ptr(elev), ptr(qc), n, ptr(g), iew, jns,
                        var, max_difference, pressure, ptr(lt), ptr(aob), ptr(err), ptr(thr))
    return qc, aob, err, thr


def buddy_check(obs, xob, yob, qc_flag, elev, gridded, var, pressure, dx_km, buddy_weight,
                buddy_difference):
    g = f32(gridded); jns, iew = g.shape
    obs, xob, yob, elev = f32(obs), f32(xob), f32(yob), f32(elev)
    qc = i32(qc_flag).copy(); n = obs.size
    aob = np.zeros(n, np.float32); err = np.zeros(n, np.float32); dm = np.zeros(n, np.float32)
    bnf = np.zeros(n, np.int32)
    lib().ogo_buddy_check(ptr(obs), n, ptr(xob), ptr(yob), ptr(qc), ptr(elev), ptr(g), iew, jns,
                          var, pressure, dx_km, buddy_weight, buddy_difference, ptr(aob),
                          ptr(err), ptr(dm), ptr(bnf))
    return qc, aob, err, dm, bnf


def smooth_5(field, passes, crsdot=1):
    g = f32(field).copy(); jns, iew = g.shape
    lib().ogo_smooth_5(ptr(g), iew, jns, passes, crsdot)
    return g


def smoother_desmoother(field, passes, crsdot=1):
    g = f32(field).copy(); jns, iew = g.shape
    lib().ogo_smoother_desmoother(ptr(g), iew, jns, passes, crsdot)
    return g


def modify_qc_flag(qc, err, thr, test):
    qc = i32(qc).copy(); err = f32(err); thr = f32(thr)
    lib().ogo_modify_qc_flag(ptr(qc), qc.size, ptr(err), ptr(thr), test)
    return qc


def add_to_qc_flag(q, add):
    v = C.c_int(int(q))
    lib().ogo_add_to_qc_flag(C.byref(v), add)
    return v.value


def ob_density(xob, yob, grid_dist, tobbox):
    """tobbox: (iew, jns) C order == Fortran tobbox(jns,iew); returns the accumulated copy."""
    tb = f32(tobbox).copy(); iew, jns = tb.shape
    xob, yob = f32(xob), f32(yob)
    lib().ogo_ob_density(ptr(xob), ptr(yob), C.c_float(grid_dist), xob.size, ptr(tb), iew, jns)
    return tb


def obs_distance(tobbox, dx):
    tb = f32(tobbox); iew, jns = tb.shape
    out = np.zeros_like(tb)
    lib().ogo_obs_distance(ptr(tb), ptr(out), iew, jns, C.c_float(dx))
    return out


def wind_increment_to_c_grid(u, uA, uC, v, vA, vC):
    """3-D arrays (kbu, iew, jns) C order == Fortran (jns,iew,kbu)."""
    u, v = f32(u).copy(), f32(v).copy()
    uA, uC, vA, vC = f32(uA), f32(uC), f32(vA), f32(vC)
    kbu, iew, jns = u.shape
    l = lib()
    l.ogo_wind_increment_u.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int]
    l.ogo_wind_increment_v.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int]
    l.ogo_wind_increment_u(ptr(u), ptr(uA), ptr(uC), jns, iew, kbu)
    l.ogo_wind_increment_v(ptr(v), ptr(vA), ptr(vC), jns, iew, kbu)
    return u, v


def pres_sfc_from_slp(pres_sfc, slp_C, slp_x):
    p, c, x = f32(pres_sfc), f32(slp_C), f32(slp_x)
    iew, jns = p.shape
    out = np.zeros_like(p)
    l = lib()
    l.ogo_pres_sfc_from_slp.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_int]
    l.ogo_pres_sfc_from_slp(ptr(p), ptr(c), ptr(x), ptr(out), jns, iew)
    return out


def temporal_interp(first, second, icount_fdda, icount_1, icount_2, cross):
    a, b = f32(first), f32(second)
    nlev = a.shape[0] if a.ndim == 3 else 1
    iew, jns = a.shape[-2], a.shape[-1]
    out = np.zeros_like(a)
    l = lib()
    l.ogo_temporal_interp.restype = C.c_int
    l.ogo_temporal_interp.argtypes = [_vp, _vp, _vp] + [C.c_int] * 7
    rc = l.ogo_temporal_interp(ptr(a), ptr(b), ptr(out), jns, iew, nlev, icount_fdda, icount_1,
                               icount_2, int(cross))
    return rc, out


def mixing_ratio(rh, t, h, terrain, slp_x_pa, pressure):
    rh = f32(rh).copy(); t, h = f32(t), f32(h)
    terrain, slp, pressure = f32(terrain), f32(slp_x_pa), f32(pressure)
    kbu, iew, jns = rh.shape
    qv = np.zeros_like(rh); psfc = np.zeros((iew, jns), np.float32)
    l = lib()
    l.ogo_mixing_ratio.restype = C.c_int
    l.ogo_mixing_ratio.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]
    rc = l.ogo_mixing_ratio(ptr(rh), ptr(t), ptr(h), ptr(terrain), ptr(slp), ptr(pressure), iew, jns,
                            kbu, ptr(qv), ptr(psfc))
    return rc, rh, qv, psfc


def dewpoint(t, rh):
    return lib().ogo_dewpoint(float(t), float(rh))


def logf(x):
    """The host libm's logf over an array (gfortran's LOG)."""
    x = f32(x); out = np.zeros_like(x)
    l = lib()
    l.ogo_logf_array.argtypes = [_vp, C.c_int, _vp]
    l.ogo_logf_array(ptr(x), x.size, ptr(out))
    return out


def qc_time(case, k_begin=0, k_end=None, run_consistency=True, threads=1):
    """Run the oracle's proc_qc mirror on a synth.TimeCase (flags updated in place)."""
    d = case.desc()
    k_end = case.kbu if k_end is None else k_end
    return lib().ogo_qc_time_mt(C.byref(d), C.byref(case.qc_params), k_begin, k_end,
                                int(run_consistency), threads)


def qc_time_fdda(case, tobbox):
    """proc_qc of a surface-FDDA time; returns (tobbox, odis), (iew, jns) C order."""
    d = case.desc()
    tb = f32(tobbox).copy(); od = np.zeros_like(tb)
    l = lib()
    l.ogo_qc_time_fdda.restype = C.c_int
    l.ogo_qc_time_fdda.argtypes = [C.POINTER(OgTimeDesc), C.POINTER(OgQcParams), _vp, _vp]
    l.ogo_qc_time_fdda(C.byref(d), C.byref(case.qc_params), ptr(tb), ptr(od))
    return tb, od


def qc_consistency(case, k_begin=0, k_end=None):
    d = case.desc()
    l = lib()
    l.ogo_qc_consistency.restype = None
    l.ogo_qc_consistency.argtypes = [C.POINTER(OgTimeDesc), C.c_int, C.c_int]
    l.ogo_qc_consistency(C.byref(d), k_begin, case.kbu if k_end is None else k_end)


def oa_time(case, k_begin=0, k_end=None, threads=1):
    d = case.desc()
    k_end = case.kbu if k_end is None else k_end
    return lib().ogo_oa_time_mt(C.byref(d), C.byref(case.oa_params), k_begin, k_end, threads)


# ---- the reference's observation access as written, over a report list (og_oracle_reports.c) ----
def qc_time_reports(case, rl, date, time, request_p_diff=1):
    """proc_qc over the report list `rl` (obsgrid_b200.reports.ReportList) with the first guess of
    `case`: 'get' / error_max / buddy_check / 'put' per slab, then qc_consistency; modifies rl."""
    d = case.desc()
    l = lib()
    l.ogo_qc_time_reports.restype = C.c_int
    l.ogo_qc_time_reports.argtypes = [C.POINTER(OgTimeDesc), C.POINTER(OgQcParams), _vp, C.c_int, _vp,
                                      C.c_int, C.c_int, C.c_int]
    return l.ogo_qc_time_reports(C.byref(d), C.byref(case.qc_params), rl.reports, rl.nrep, rl.meas_ptr(),
                                 date, time, request_p_diff)


def get_obs(rl, is_get, var, level, date, time, request_p_diff, request_qc_max, iew, jns):
    """proc_ob_access 'get' (is_get) / 'use' of one slab; returns dict of the returned arrays."""
    n = rl.nrep
    val, x, y, lon, elev = (np.zeros(n, np.float32) for _ in range(5))
    idx, qc = np.zeros(n, np.int32), np.zeros(n, np.int32)
    l = lib()
    l.ogo_get_obs.restype = C.c_int
    l.ogo_get_obs.argtypes = [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_float] + [C.c_int] * 6 + [_vp] * 7
    m = l.ogo_get_obs(rl.reports, n, rl.meas_ptr(), int(is_get), var, level, date, time, request_p_diff,
                      request_qc_max, iew, jns, ptr(val), ptr(x), ptr(y), ptr(lon), ptr(elev), ptr(idx), ptr(qc))
    return dict(n=m, value=val[:m], x=x[:m], y=y[:m], lon=lon[:m], elev=elev[:m], index=idx[:m], qc=qc[:m])
