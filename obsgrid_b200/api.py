"""Host-side mirror of the reference's interface for the hot path, over the C ABI.

Function names, argument meaning and error behaviour follow the reference procedures
(`cressman` oa_module.F90:25, `error_max` qc1_module.F90:9, `buddy_check` qc2_module.F90:9,
`local` qc0_module.F90:236, `proc_qc` / `proc_oa` level x variable loops); 2-D slabs are
numpy arrays of shape (jns, iew) in C order, i.e. the memory of Fortran's gridded(iew,jns).
Everything runs on the GPU through libobsgrid_gpu.so; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import OG_MAX_SCAN, OgStats, check, f32, i32, ptr


class Context:
    """An og_ctx: one CUDA device + stream.  Raises OgError without an sm_100 GPU."""

    def __init__(self, device: int = 0):
        self._lib = _capi.lib()
        h = C.c_void_p()
        check(self._lib.og_init(C.byref(h), device), "og_init")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.og_finalize(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- bookkeeping -------------------------------------------------------------------
    def stats(self) -> dict:
        s = OgStats()
        check(self._lib.og_get_stats(self._h, C.byref(s)), "og_get_stats")
        return {n: getattr(s, n) for n, _ in s._fields_}

    def reset_stats(self):
        check(self._lib.og_reset_stats(self._h), "og_reset_stats")

    def set_counting(self, on: bool):
        check(self._lib.og_set_counting(self._h, int(on)), "og_set_counting")

    def set_arithmetic(self, mode: str):
        """'exact' (reference arithmetic, bit-identical) or 'contracted' (FMA + hardware reciprocal)."""
        check(self._lib.og_set_arithmetic(self._h, {"exact": 0, "contracted": 1}[mode]), "og_set_arithmetic")

    def arithmetic(self) -> str:
        return ("exact", "contracted")[self._lib.og_get_arithmetic(self._h)]

    def set_timing(self, on: bool):
        check(self._lib.og_set_timing(self._h, int(on)), "og_set_timing")

    def kernel_ms(self):
        a, b = C.c_float(0), C.c_float(0)
        check(self._lib.og_last_kernel_ms(self._h, C.byref(a), C.byref(b)), "og_last_kernel_ms")
        return a.value, b.value

    def sync(self):
        check(self._lib.og_sync(self._h), "og_sync")

    def stream(self) -> int:
        return int(self._lib.og_stream(self._h) or 0)

    def fp32_peak(self) -> float:
        v = C.c_double(0)
        check(self._lib.og_fp32_peak(self._h, C.byref(v)), "og_fp32_peak")
        return v.value

    def selftest_div(self, n=1 << 22, seed=1234, rmax=100) -> int:
        bad = C.c_longlong(-1)
        check(self._lib.og_selftest_div(self._h, n, seed, rmax, C.byref(bad)), "og_selftest_div")
        return bad.value

    def selftest_logf(self, x):
        """The device's LOG (glibc's logf algorithm, csrc/og_logf.h) of every element of x."""
        x = f32(x); out = np.zeros_like(x)
        check(self._lib.og_selftest_logf(self._h, ptr(x), x.size, ptr(out)), "og_selftest_logf")
        return out

    # ---- per-slab drop-ins ---------------------------------------------------------------
    def innovation(self, gridded, xob, yob, obs_value, scale=1.0, use_first_guess=True):
        g = f32(gridded); jns, iew = g.shape
        xob, yob, obs_value = f32(xob), f32(yob), f32(obs_value)
        n = xob.size
        diff = np.zeros(n, np.float32); fg = np.full(n, -9999999.0, np.float32)
        check(self._lib.og_innovation(self._h, ptr(g), iew, jns, ptr(xob), ptr(yob),
                                      ptr(obs_value), n, scale, int(use_first_guess), ptr(diff),
                                      ptr(fg)), "og_innovation")
        return diff, fg

    def cressman(self, diff, xob, yob, gridded, var, radius, dxd_km, u_banana=None,
                 v_banana=None, pressure=1001.0, passes=0, smooth_type=1, use_first_guess=True,
                 fg_value=None, scale_rh=False, crsdot=1):
        g = f32(gridded).copy(); jns, iew = g.shape
        diff, xob, yob = f32(diff), f32(xob), f32(yob)
        ub = f32(u_banana) if u_banana is not None else None
        vb = f32(v_banana) if v_banana is not None else None
        fg = f32(fg_value) if fg_value is not None else None
        check(self._lib.og_cressman(self._h, ptr(diff), ptr(xob), ptr(yob), diff.size, ptr(g), iew,
                                    jns, crsdot, var, int(radius), dxd_km, ptr(ub), ptr(vb),
                                    pressure, passes, smooth_type, int(use_first_guess), ptr(fg),
                                    int(scale_rh)), "og_cressman")
        return g

    def oa_slab(self, gridded, var, pressure, obs_value, xob, yob, radii, dxd_km, u_banana=None,
                v_banana=None, scale=1.0, sfc_mult=1.0, passes=0, smooth_type=1,
                use_first_guess=True, scale_rh=False):
        g = f32(gridded).copy(); jns, iew = g.shape
        obs_value, xob, yob = f32(obs_value), f32(xob), f32(yob)
        ub = f32(u_banana) if u_banana is not None else None
        vb = f32(v_banana) if v_banana is not None else None
        rad = np.full(OG_MAX_SCAN, -1, np.int32); rad[:len(radii)] = radii
        ns = C.c_int(0)
        rc = self._lib.og_oa_slab(self._h, ptr(g), iew, jns, var, pressure, ptr(obs_value), ptr(xob),
                                  ptr(yob), xob.size, scale, ptr(rad), sfc_mult, dxd_km, ptr(ub),
                                  ptr(vb), passes, smooth_type, int(use_first_guess),
                                  int(scale_rh), C.byref(ns))
        if rc not in (0, -5):
            check(rc, "og_oa_slab")
        return g, rc, ns.value

    def local_time(self, time_hhmm, lonob):
        lon = f32(lonob); out = np.zeros(lon.size, np.int32)
        check(self._lib.og_local_time(self._h, int(time_hhmm), ptr(lon), lon.size, ptr(out)),
              "og_local_time")
        return out

    def fg_t_at_point(self, t3d, h3d, x, y, height):
        """GET_FG_T_AT_POINT batched; t3d/h3d: (kbu, iew, jns) C order == Fortran (jns,iew,kbu)."""
        t3d, h3d = f32(t3d), f32(h3d)
        kbu, iew, jns = t3d.shape
        x, y, height = f32(x), f32(y), f32(height)
        out = np.zeros(x.size, np.float32)
        check(self._lib.og_fg_t_at_point(self._h, ptr(t3d), ptr(h3d), jns, iew, kbu, ptr(x), ptr(y),
                                         ptr(height), x.size, ptr(out)), "og_fg_t_at_point")
        return out

    def ob_density(self, xob, yob, grid_dist, tobbox):
        """tobbox: (iew, jns) C order == Fortran tobbox(jns,iew); returns the accumulated copy."""
        tb = f32(tobbox).copy(); iew, jns = tb.shape
        xob, yob = f32(xob), f32(yob)
        check(self._lib.og_ob_density(self._h, ptr(xob), ptr(yob), grid_dist, xob.size, ptr(tb), iew, jns),
              "og_ob_density")
        return tb

    def obs_distance(self, tobbox, dx):
        tb = f32(tobbox); iew, jns = tb.shape
        out = np.zeros_like(tb)
        check(self._lib.og_obs_distance(self._h, ptr(tb), ptr(out), iew, jns, dx), "og_obs_distance")
        return out

    # ---- post-analysis epilogue (3-D arrays: (kbu, iew, jns) C order == Fortran (jns,iew,kbu)) ----
    def wind_increment_to_c_grid(self, u, uA, uC, v, vA, vC):
        """write_analysis :316-339: u = uC + crs2dot_pertU(u - uA), v = vC + crs2dot_pertV(v - vA)."""
        u, v = f32(u).copy(), f32(v).copy()
        uA, uC, vA, vC = f32(uA), f32(uC), f32(vA), f32(vC)
        kbu, iew, jns = u.shape
        check(self._lib.og_wind_increment_to_c_grid(self._h, ptr(u), ptr(uA), ptr(uC), ptr(v), ptr(vA),
                                                    ptr(vC), jns, iew, kbu), "og_wind_increment_to_c_grid")
        return u, v

    def temporal_interp(self, first, second, icount_fdda, icount_1, icount_2, cross):
        """tinterp_module.F90:53-215 for one field; arrays (nlev, iew, jns) or (iew, jns)."""
        a, b = f32(first), f32(second)
        shp = a.shape
        nlev = shp[0] if a.ndim == 3 else 1
        iew, jns = shp[-2], shp[-1]
        out = np.zeros_like(a)
        check(self._lib.og_temporal_interp(self._h, ptr(a), ptr(b), ptr(out), jns, iew, nlev,
                                           icount_fdda, icount_1, icount_2, int(cross)), "og_temporal_interp")
        return out

    def pres_sfc_from_slp(self, pres_sfc, slp_C, slp_x):
        p, c, x = f32(pres_sfc), f32(slp_C), f32(slp_x)
        iew, jns = p.shape
        out = np.zeros_like(p)
        check(self._lib.og_pres_sfc_from_slp(self._h, ptr(p), ptr(c), ptr(x), ptr(out), jns, iew),
              "og_pres_sfc_from_slp")
        return out

    def mixing_ratio(self, rh, t, h, terrain, slp_x_pa, pressure):
        """mixing_ratio + sfcprs; returns (rh clamped, qv, psfc)."""
        rh = f32(rh).copy(); t, h = f32(t), f32(h)
        terrain, slp, pressure = f32(terrain), f32(slp_x_pa), f32(pressure)
        kbu, iew, jns = rh.shape
        qv = np.zeros_like(rh); psfc = np.zeros((iew, jns), np.float32)
        check(self._lib.og_mixing_ratio(self._h, ptr(rh), ptr(t), ptr(h), ptr(terrain), ptr(slp),
                                        ptr(pressure), iew, jns, kbu, ptr(qv), ptr(psfc)), "og_mixing_ratio")
        return rh, qv, psfc

    def error_max(self, obs, xob, yob, elev, qc_flag, gridded, var, max_difference, pressure,
                  local_t):
        g = f32(gridded); jns, iew = g.shape
        obs, xob, yob, elev = f32(obs), f32(xob), f32(yob), f32(elev)
        qc = i32(qc_flag).copy(); lt = i32(local_t); n = obs.size
        aob = np.zeros(n, np.float32); err = np.zeros(n, np.float32); thr = np.zeros(n, np.float32)
        check(self._lib.og_error_max(self._h, ptr(obs), ptr(xob), ptr(yob), ptr(elev), ptr(qc), n,
                                     ptr(g), iew, jns, var, max_difference, pressure, ptr(lt),
                                     ptr(aob), ptr(err), ptr(thr)), "og_error_max")
        return qc, aob, err, thr

    def buddy_check(self, obs, xob, yob, qc_flag, elev, gridded, var, pressure, dx_km,
                    buddy_weight, buddy_difference):
        g = f32(gridded); jns, iew = g.shape
        obs, xob, yob, elev = f32(obs), f32(xob), f32(yob), f32(elev)
        qc = i32(qc_flag).copy(); n = obs.size
        aob = np.zeros(n, np.float32); err = np.zeros(n, np.float32); dm = np.zeros(n, np.float32)
        bnf = np.zeros(n, np.int32)
        check(self._lib.og_buddy_check(self._h, ptr(obs), n, ptr(xob), ptr(yob), ptr(qc), ptr(elev),
                                       ptr(g), iew, jns, var, pressure, dx_km, buddy_weight,
                                       buddy_difference, ptr(aob), ptr(err), ptr(dm), ptr(bnf)),
              "og_buddy_check")
        return qc, aob, err, dm, bnf

    def qc_slab(self, gridded, var, pressure, obs, xob, yob, elev, local_t, qc_flag, max_error,
                max_buddy, buddy_weight, dx_km, do_error_max=True, do_buddy=True):
        g = f32(gridded); jns, iew = g.shape
        obs, xob, yob, elev = f32(obs), f32(xob), f32(yob), f32(elev)
        qc = i32(qc_flag).copy(); lt = i32(local_t)
        check(self._lib.og_qc_slab(self._h, int(do_error_max), int(do_buddy), ptr(g), iew, jns, var,
                                   pressure, ptr(obs), ptr(xob), ptr(yob), ptr(elev), ptr(lt),
                                   obs.size, max_error, max_buddy, buddy_weight, dx_km, ptr(qc)),
              "og_qc_slab")
        return qc

    # ---- per-time drivers (synth.TimeCase; arrays updated in place) ------------------------
    def proc_qc(self, case, k_begin=0, k_end=None, run_consistency=True):
        d = case.desc()
        k_end = case.kbu if k_end is None else k_end
        check(self._lib.og_qc_time_levels(self._h, C.byref(d), C.byref(case.qc_params), k_begin,
                                          k_end, int(run_consistency)), "og_qc_time_levels")

    def proc_qc_fdda(self, case, tobbox):
        """proc_qc of a surface-FDDA time: flags in place; returns (tobbox, odis), (iew, jns) C order."""
        d = case.desc()
        tb = f32(tobbox).copy(); od = np.zeros_like(tb)
        check(self._lib.og_qc_time_fdda(self._h, C.byref(d), C.byref(case.qc_params), ptr(tb), ptr(od)),
              "og_qc_time_fdda")
        return tb, od

    def qc_consistency(self, case, k_begin=0, k_end=None):
        d = case.desc()
        check(self._lib.og_qc_consistency(self._h, C.byref(d), k_begin,
                                          case.kbu if k_end is None else k_end), "og_qc_consistency")

    def dev_qc_consistency(self, desc, k_begin, k_end):
        check(self._lib.og_dev_qc_consistency(self._h, C.byref(desc), k_begin, k_end), "og_dev_qc_consistency")

    def proc_oa(self, case, k_begin=0, k_end=None):
        d = case.desc()
        k_end = case.kbu if k_end is None else k_end
        rc = self._lib.og_oa_time(self._h, C.byref(d), C.byref(case.oa_params), k_begin, k_end)
        if rc not in (0, -5):
            check(rc, "og_oa_time")
        return rc

    def proc_qc_oa(self, case, k_begin=0, k_end=None, pipeline_groups=0):
        """proc_qc then proc_oa of one time period in one pipelined call (og_qc_oa_time)."""
        d = case.desc()
        k_end = case.kbu if k_end is None else k_end
        rc = self._lib.og_qc_oa_time(self._h, C.byref(d), C.byref(case.qc_params),
                                     C.byref(case.oa_params), k_begin, k_end, pipeline_groups)
        if rc not in (0, -5):
            check(rc, "og_qc_oa_time")
        return rc

    # asynchronous per-time API: several time periods in flight (og_time_*)
    def time_upload(self, case, slot, k_begin=0, k_end=None):
        d = case.desc()
        check(self._lib.og_time_upload(self._h, C.byref(d), k_begin,
                                       case.kbu if k_end is None else k_end, slot), "og_time_upload")

    def time_compute(self, case, slot, k_begin=0, k_end=None):
        d = case.desc()
        rc = self._lib.og_time_compute(self._h, C.byref(d), C.byref(case.qc_params),
                                       C.byref(case.oa_params), k_begin,
                                       case.kbu if k_end is None else k_end, slot)
        if rc not in (0, -5):
            check(rc, "og_time_compute")
        return rc

    def time_download(self, case_out, slot, k_begin=0, k_end=None):
        d = case_out.desc()
        check(self._lib.og_time_download(self._h, C.byref(d), k_begin,
                                         case_out.kbu if k_end is None else k_end, slot),
              "og_time_download")

    def time_upload_csr(self, case, csr, slot, k_begin=0, k_end=None):
        """og_time_upload with the observation rows in sparse form (synth.CsrObs)."""
        d, o = case.desc(), csr.struct()
        check(self._lib.og_time_upload_csr(self._h, C.byref(d), C.byref(o), k_begin,
                                           case.kbu if k_end is None else k_end, slot), "og_time_upload_csr")

    def time_download_csr(self, case_out, csr_out, slot, k_begin=0, k_end=None):
        d, o = case_out.desc(), csr_out.struct()
        check(self._lib.og_time_download_csr(self._h, C.byref(d), C.byref(o), k_begin,
                                             case_out.kbu if k_end is None else k_end, slot), "og_time_download_csr")

    def time_wait(self, slot):
        check(self._lib.og_time_wait(self._h, slot), "og_time_wait")

    # resident variants: `desc` holds device pointers (see bench.py)
    def dev_proc_qc(self, desc, qc_params, k_begin, k_end, run_consistency=True):
        check(self._lib.og_dev_qc_time(self._h, C.byref(desc), C.byref(qc_params), k_begin, k_end,
                                       int(run_consistency)), "og_dev_qc_time")

    def dev_proc_oa(self, desc, oa_params, k_begin, k_end):
        rc = self._lib.og_dev_oa_time(self._h, C.byref(desc), C.byref(oa_params), k_begin, k_end)
        if rc not in (0, -5):
            check(rc, "og_dev_oa_time")
        return rc

    def dev_temporal_interp(self, first_ptr, second_ptr, out_ptr, jns, iew, nlev, icount_fdda, icount_1, icount_2, cross):
        check(self._lib.og_dev_temporal_interp(self._h, first_ptr, second_ptr, out_ptr, jns, iew, nlev,
                                               icount_fdda, icount_1, icount_2, int(cross)), "og_dev_temporal_interp")

    def set_exchange(self, rank, world, fn=None):
        """One slab over several GPUs (og_set_exchange): `fn` is a _capi.OG_EXCHANGE_FN (shard.make_exchange
        builds one over torch.distributed); the context keeps it alive.  fn None / world 1: off."""
        self._exchange_fn = fn
        check(self._lib.og_set_exchange(self._h, rank, world, C.cast(fn, C.c_void_p) if fn is not None else None, None),
              "og_set_exchange")

    def dev_scale_field(self, ptr_, jns, iew, factor, interior_only=False):
        check(self._lib.og_dev_scale_field(self._h, ptr_, jns, iew, factor, int(interior_only)), "og_dev_scale_field")

    def dev_proc_qc_fdda(self, desc, qc_params, tobbox_ptr, odis_ptr):
        check(self._lib.og_dev_qc_time_fdda(self._h, C.byref(desc), C.byref(qc_params), tobbox_ptr, odis_ptr),
              "og_dev_qc_time_fdda")

    def dev_proc_qc_items(self, desc, qc_params, levels, masks, run_consistency=True):
        lv, mk = i32(levels), i32(masks)
        check(self._lib.og_dev_qc_time_items(self._h, C.byref(desc), C.byref(qc_params), ptr(lv), ptr(mk),
                                             lv.size, int(run_consistency)), "og_dev_qc_time_items")

    def dev_proc_oa_items(self, desc, oa_params, levels, masks):
        lv, mk = i32(levels), i32(masks)
        rc = self._lib.og_dev_oa_time_items(self._h, C.byref(desc), C.byref(oa_params), ptr(lv), ptr(mk), lv.size)
        if rc not in (0, -5):
            check(rc, "og_dev_oa_time_items")
        return rc
