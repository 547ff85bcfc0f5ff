"""proc_qc / proc_oa as the reference runs them -- one blocking call per (level, variable) slab
and stage -- over the per-slab drop-ins of the C ABI (og_error_max, og_buddy_check, og_innovation,
og_cressman), with host arrays.

This is the path an unmodified proc_qc.F90 / proc_oa.F90 takes once cressman, error_max and
buddy_check are replaced by the wrappers of fortran/obsgrid_gpu_iface.F90: the level x variable
loops (src/proc_qc.F90:243-408, src/proc_oa.F90:356-849), the observation selection of
proc_ob_access on the flat table, the radius logic of the multi_scan loop (:641-707) and the
innovations between scans (:748-780) all stay on the host, exactly as in the reference.  bench.py
times it as `e2e_per_slab` (what "no change to proc_qc / proc_oa" costs against the batched
og_time_* path); tests/test_gpu_parity.py holds it bit-identical to the batched path.
"""
from __future__ import annotations

import numpy as np

from ._capi import OG_MISSING_R, FAILS_BUDDY_CHECK, FAILS_ERROR_MAX

F = np.float32


def _not_missing(a):
    return np.abs(a - F(OG_MISSING_R)) > F(1.0)


def _in_domain(case):
    x, y = case.xob, case.yob
    return (x >= 2) & (x <= case.iew - 1) & (y >= 2) & (y <= case.jns - 1)


def _dewpoint(ctx, t, rh):
    """ob_access_module.F90:476-486 / proc_qc.F90:315-319 in REAL(4); LOG through the library's glibc logf."""
    t = np.asarray(t, F); rh = np.asarray(rh, F)
    arg = np.where(rh < F(0.0001), F(0.0001) / F(100.0), rh / F(100.0)).astype(F)
    lg = ctx.selftest_logf(arg.ravel()).reshape(arg.shape) if arg.size else arg
    rv_over_l = F(1.0) / F(5418.12)
    return (F(1.0) / (F(1.0) / t - rv_over_l * lg)).astype(F)


def scan_radii(radius_influence, sfc_mult, pressure):
    """proc_oa.F90:641-707; returns (radii, stop) -- stop: the reference STOPs (scan 1 below 4.5 cells)."""
    out, stop = [], False
    is_sfc = abs(F(pressure) - F(1001.0)) < F(0.0001)
    for n, r in enumerate(radius_influence):
        if r < 0:
            break
        sf = F(sfc_mult) if is_sfc else F(1.0)
        r_scaled, r_scan1 = F(r) * sf, F(radius_influence[0]) * sf
        if F(sfc_mult) < F(0.99999) and is_sfc:
            if r_scan1 > F(100.0):
                r_scaled = r_scaled * (F(100.0) / r_scan1)
            if r_scaled < F(4.5):
                stop = n == 0
                break
        out.append(int(np.floor(np.abs(r_scaled) + F(0.5)) * np.sign(r_scaled)))     # NINT
    return out, stop


def proc_qc_per_slab(ctx, case):
    """Flags of `case` updated in place; returns the number of library calls."""
    p = case.qc_params
    dom = _in_domain(case)
    bogus = case.bogus if case.bogus is not None else np.zeros(case.nrep, np.int32)
    calls = 0
    for k in range(case.kbu if (p.qc_test_error_max or p.qc_test_buddy) else 0):     # any_checks_to_do, proc_qc.F90:237
        pk = float(case.pressure[k])
        for ivar in (1, 2, 3, 4, 5, 7):
            if k > 0 and ivar == 5:
                continue
            if ivar == 7 and not p.qc_dewpoint:
                continue
            if p.var_mask and not (p.var_mask & (1 << (ivar - 1))):
                continue
            if ivar == 5:
                ob, qc, g = case.ob_slp, case.qc_slp, (case.slab_xy("slp") * F(100.0)).astype(F)
                me, mb = p.max_error_p, p.max_buddy_p
            else:
                ob = {1: case.ob_u, 2: case.ob_v, 3: case.ob_t, 4: case.ob_rh, 7: case.ob_rh}[ivar][k]
                qc = {1: case.qc_u, 2: case.qc_v, 3: case.qc_t, 4: case.qc_rh, 7: case.qc_rh}[ivar][k]
                name = {1: "u", 2: "v", 3: "t", 4: "rh", 7: "rh"}[ivar]
                g = case.slab_xy(name, k)
                me, mb = {1: (p.max_error_uv, p.max_buddy_uv), 2: (p.max_error_uv, p.max_buddy_uv),
                          3: (p.max_error_t, p.max_buddy_t), 4: (p.max_error_rh, p.max_buddy_rh),
                          7: (p.max_error_dewpoint, p.max_buddy_dewpoint)}[ivar]
                if ivar == 4:
                    me = me * 2.0 if pk == 300.0 else (me * 3.0 if pk < 300.0 else me)     # proc_qc.F90:282-288
            sel = _not_missing(ob) & (qc < (1 << 20)) & dom & (bogus == 0)
            if ivar == 7:
                sel &= _not_missing(case.ob_t[k]) & (case.qc_t[k] < (1 << 20))
            idx = np.flatnonzero(sel)
            val = ob[idx]
            if ivar == 7:
                g = _dewpoint(ctx, case.slab_xy("t", k), g)
                val = _dewpoint(ctx, case.ob_t[k][idx], val)
                calls += 2
            x, y, elev = case.xob[idx], case.yob[idx], case.elev[idx]
            lt = ctx.local_time(p.time_hhmmss // 100, case.lon[idx]) if idx.size else np.zeros(0, np.int32)
            q = qc[idx].copy()
            if idx.size and p.qc_test_error_max:
                q = ctx.error_max(val, x, y, elev, q, g, ivar, me, pk, lt)[0]
                calls += 1
            if idx.size and p.qc_test_buddy:
                q = ctx.buddy_check(val, x, y, q, elev, g, ivar, pk, float(case.dx_km), p.buddy_weight, mb)[0]
                calls += 1
            qc[idx] = q                                                    # 'put'
            if ivar == 7 and case.ob_td is not None and case.qc_td is not None:
                case.ob_td[k][idx] = val
                case.qc_td[k][idx] = q
    ctx.qc_consistency(case)
    return calls + 1


def proc_oa_per_slab(ctx, case):
    """Fields of `case` analysed in place; returns the number of library calls."""
    p = case.oa_params
    dom = _in_domain(case)
    qc_max = min(FAILS_ERROR_MAX, FAILS_BUDDY_CHECK)                       # proc_oa.F90:462
    radius_influence = [int(r) for r in p.radius_influence]
    calls = 0
    if p.clamp_fg_rh:                                                      # mixing_ratio's clamp, final_analysis_module.F90:1055
        case.rh[:, :-1, :-1] = np.minimum(np.maximum(case.rh[:, :-1, :-1], F(1.0)), F(100.0))
    for k in range(case.kbu):
        if p.surface_only and k > 0:
            break
        pk = float(case.pressure[k])
        ub, vb = case.slab_xy("u", k), case.slab_xy("v", k)                # pre-analysis winds, proc_oa.F90:376-386
        for ivar in (1, 2, 3, 4, 5):
            if ivar == 5 and k > 0:
                continue
            if p.var_mask and not (p.var_mask & (1 << (ivar - 1))):
                continue
            name = {1: "u", 2: "v", 3: "t", 4: "rh", 5: "slp"}[ivar]
            ob = case.ob_slp if ivar == 5 else getattr(case, "ob_" + name)[k]
            qc = case.qc_slp if ivar == 5 else getattr(case, "qc_" + name)[k]
            idx = np.flatnonzero(_not_missing(ob) & (qc < qc_max) & dom)   # 'use'
            val, x, y = ob[idx], case.xob[idx], case.yob[idx]
            g = case.slab_xy(name, k)
            scale = 100.0 if ivar == 5 else 1.0
            radii, _ = scan_radii(radius_influence, p.radius_influence_sfc_mult, pk)
            passes = [[p.smooth_sfc_wind, p.smooth_sfc_wind, p.smooth_sfc_temp, p.smooth_sfc_rh, p.smooth_sfc_slp],
                      [p.smooth_upper_wind, p.smooth_upper_wind, p.smooth_upper_temp, p.smooth_upper_rh, 0]][k > 0][ivar - 1]
            fg = None
            if idx.size:
                diff, fg = ctx.innovation(g, x, y, val, scale, bool(p.use_first_guess)); calls += 1
            for R in radii:
                if idx.size == 0:
                    diff = np.zeros(0, F)
                g = ctx.cressman(diff, x, y, g, ivar, R, float(case.dx_km), ub, vb, pk, passes, p.smooth_type,
                                 bool(p.use_first_guess), fg, bool(p.scale_cressman_rh_decreases))
                calls += 1
                if idx.size:
                    diff, _ = ctx.innovation(g, x, y, val, scale, bool(p.use_first_guess)); calls += 1
            if ivar == 4:                                                  # clean_rh, proc_oa.F90:803
                g = np.minimum(np.maximum(g, F(0.0)), F(100.0))
            if ivar == 5:
                case.slp[...] = g.T
            else:
                getattr(case, name)[k] = g.T
    return calls
