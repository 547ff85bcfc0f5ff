"""ctypes binding of libobsgrid_gpu.so (include/obsgrid_gpu.h).

This module only declares the C ABI; it contains no numerics and no fallback.  If the
shared library has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)
importing :func:`lib` raises -- the product path never routes through a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libobsgrid_gpu.so"

OG_MAX_SCAN = 10
OG_MISSING_R = np.float32(-888888.0)
NO_BUDDIES = 1 << 14
FAILS_ERROR_MAX = 1 << 16
FAILS_BUDDY_CHECK = 1 << 17

VAR_UU, VAR_VV, VAR_TT, VAR_RH, VAR_PMSL, VAR_PSFC, VAR_DEWPOINT = 1, 2, 3, 4, 5, 6, 7
VAR_PMSLPSFC = 6
VAR_NAMES = {1: "UU", 2: "VV", 3: "TT", 4: "RH", 5: "PMSL", 6: "PSFC", 7: "DEWPOINT"}

fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)


class OgStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "kernel_launches", "cressman_updates", "cressman_candidates", "buddy_pair_tests",
        "buddy_obs", "h2d_bytes", "d2h_bytes", "buddy_fixpoint_iters",
        "buddy_pairs_examined", "cressman_updates_ellipse", "cressman_updates_banana")]


class OgQcParams(C.Structure):
    _fields_ = [("qc_test_error_max", C.c_int), ("qc_test_buddy", C.c_int),
                ("max_error_t", C.c_float), ("max_error_uv", C.c_float),
                ("max_error_rh", C.c_float), ("max_error_dewpoint", C.c_float),
                ("max_error_p", C.c_float),
                ("max_buddy_t", C.c_float), ("max_buddy_uv", C.c_float),
                ("max_buddy_rh", C.c_float), ("max_buddy_dewpoint", C.c_float),
                ("max_buddy_p", C.c_float),
                ("buddy_weight", C.c_float), ("time_hhmmss", C.c_int),
                ("qc_dewpoint", C.c_int), ("var_mask", C.c_int)]


class OgOaParams(C.Structure):
    _fields_ = [("radius_influence", C.c_int * OG_MAX_SCAN),
                ("radius_influence_sfc_mult", C.c_float),
                ("use_first_guess", C.c_int), ("scale_cressman_rh_decreases", C.c_int),
                ("smooth_type", C.c_int),
                ("smooth_sfc_wind", C.c_int), ("smooth_sfc_temp", C.c_int),
                ("smooth_sfc_rh", C.c_int), ("smooth_sfc_slp", C.c_int),
                ("smooth_upper_wind", C.c_int), ("smooth_upper_temp", C.c_int),
                ("smooth_upper_rh", C.c_int),
                ("surface_only", C.c_int), ("clamp_fg_rh", C.c_int), ("var_mask", C.c_int)]


class OgTimeDesc(C.Structure):
    _fields_ = [("iew", C.c_int), ("jns", C.c_int), ("kbu", C.c_int), ("dx_km", C.c_float),
                ("pressure", C.c_void_p),
                ("t", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("rh", C.c_void_p),
                ("slp", C.c_void_p),
                ("nrep", C.c_int),
                ("xob", C.c_void_p), ("yob", C.c_void_p), ("lon", C.c_void_p),
                ("elev", C.c_void_p),
                ("ob_u", C.c_void_p), ("ob_v", C.c_void_p), ("ob_t", C.c_void_p),
                ("ob_rh", C.c_void_p), ("ob_slp", C.c_void_p),
                ("qc_u", C.c_void_p), ("qc_v", C.c_void_p), ("qc_t", C.c_void_p),
                ("qc_rh", C.c_void_p), ("qc_slp", C.c_void_p),
                ("ob_td", C.c_void_p), ("qc_td", C.c_void_p), ("bogus", C.c_void_p)]


# og_exchange_fn(user, buf, count, dtype, op, stream)
OG_EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p)
OG_XCHG_INT32, OG_XCHG_UINT8 = 0, 1
OG_XCHG_SUM, OG_XCHG_MAX = 0, 1


class OgObsCsr(C.Structure):
    _fields_ = [("nnz", C.c_int), ("lev_start", C.c_void_p), ("rep", C.c_void_p),
                ("u", C.c_void_p), ("v", C.c_void_p), ("t", C.c_void_p), ("rh", C.c_void_p), ("td", C.c_void_p),
                ("qc_u", C.c_void_p), ("qc_v", C.c_void_p), ("qc_t", C.c_void_p), ("qc_rh", C.c_void_p),
                ("qc_td", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/obsgrid_gpu.h declares.
_vp = C.c_void_p
SIGNATURES = {
    "og_abi_version": (C.c_int, []),
    "og_init": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "og_finalize": (C.c_int, [_vp]),
    "og_last_error": (C.c_char_p, []),
    "og_get_stats": (C.c_int, [_vp, C.POINTER(OgStats)]),
    "og_reset_stats": (C.c_int, [_vp]),
    "og_set_counting": (C.c_int, [_vp, C.c_int]),
    "og_set_arithmetic": (C.c_int, [_vp, C.c_int]),
    "og_get_arithmetic": (C.c_int, [_vp]),
    "og_set_timing": (C.c_int, [_vp, C.c_int]),
    "og_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    "og_host_free": (C.c_int, [_vp]),
    "og_selftest_div": (C.c_int, [_vp, C.c_int, C.c_uint, C.c_int, C.POINTER(C.c_longlong)]),
    "og_selftest_logf": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "og_innovation": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_float,
                                C.c_int, _vp, _vp]),
    "og_cressman": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int,
                              C.c_int, C.c_int, C.c_float, _vp, _vp, C.c_float, C.c_int,
                              C.c_int, C.c_int, _vp, C.c_int]),
    "og_oa_slab": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_float, _vp, _vp, _vp,
                             C.c_int, C.c_float, _vp, C.c_float, C.c_float, _vp, _vp, C.c_int,
                             C.c_int, C.c_int, C.c_int, _vp]),
    "og_local_time": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp]),
    "og_error_max": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int,
                               C.c_int, C.c_float, C.c_float, _vp, _vp, _vp, _vp]),
    "og_buddy_check": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int,
                                 C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _vp, _vp,
                                 _vp, _vp]),
    "og_qc_slab": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_float,
                             _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_float, C.c_float, C.c_float,
                             C.c_float, _vp]),
    "og_fg_t_at_point": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp]),
    "og_ob_density": (C.c_int, [_vp, _vp, _vp, C.c_float, C.c_int, _vp, C.c_int, C.c_int]),
    "og_obs_distance": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_float]),
    "og_wind_increment_to_c_grid": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int]),
    "og_pres_sfc_from_slp": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int]),
    "og_temporal_interp": (C.c_int, [_vp, _vp, _vp, _vp] + [C.c_int] * 7),
    "og_mixing_ratio": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "og_qc_time": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.POINTER(OgQcParams)]),
    "og_qc_time_fdda": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.POINTER(OgQcParams), _vp, _vp]),
    "og_qc_consistency": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.c_int, C.c_int]),
    "og_dev_qc_time_items": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.POINTER(OgQcParams), _vp, _vp, C.c_int, C.c_int]),
    "og_dev_oa_time_items": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.POINTER(OgOaParams), _vp, _vp, C.c_int]),
    "og_dev_qc_consistency": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.c_int, C.c_int]),
    "og_oa_time": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.POINTER(OgOaParams), C.c_int, C.c_int]),
    "og_qc_oa_time": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.POINTER(OgQcParams), C.POINTER(OgOaParams),
                                C.c_int, C.c_int, C.c_int]),
    "og_time_upload": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.c_int, C.c_int, C.c_int]),
    "og_time_compute": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.POINTER(OgQcParams), C.POINTER(OgOaParams),
                                  C.c_int, C.c_int, C.c_int]),
    "og_time_download": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.c_int, C.c_int, C.c_int]),
    "og_time_wait": (C.c_int, [_vp, C.c_int]),
    "og_time_upload_csr": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.POINTER(OgObsCsr), C.c_int, C.c_int, C.c_int]),
    "og_time_download_csr": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.POINTER(OgObsCsr), C.c_int, C.c_int, C.c_int]),
    "og_qc_time_levels": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.POINTER(OgQcParams), C.c_int,
                                    C.c_int, C.c_int]),
    "og_dev_qc_time": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.POINTER(OgQcParams), C.c_int,
                                 C.c_int, C.c_int]),
    "og_dev_oa_time": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.POINTER(OgOaParams), C.c_int,
                                 C.c_int]),
    "og_reports_to_table": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(OgTimeDesc), _vp]),
    "og_table_to_reports": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.POINTER(OgTimeDesc), _vp]),
    "og_table_shared_levels": (C.c_int, [_vp, C.c_int, C.c_int]),
    "og_dev_temporal_interp": (C.c_int, [_vp, _vp, _vp, _vp] + [C.c_int] * 7),
    "og_dev_qc_time_fdda": (C.c_int, [_vp, C.POINTER(OgTimeDesc), C.POINTER(OgQcParams), _vp, _vp]),
    "og_dev_scale_field": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_float, C.c_int]),
    "og_set_exchange": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
    "og_stream": (_vp, [_vp]),
    "og_sync": (C.c_int, [_vp]),
    "og_last_kernel_ms": (C.c_int, [_vp, fp, fp]),
    "og_fp32_peak": (C.c_int, [_vp, C.POINTER(C.c_double)]),
}

_lib = None


class OgError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load libobsgrid_gpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        path = os.environ.get("OBSGRID_GPU_LIB", str(LIB_PATH))
        if not Path(path).exists():
            raise OgError(
                f"{path} not found: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        l = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr(a):
    """void* of a numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().og_last_error()
        raise OgError(f"{what} failed with og_status {rc}: {msg.decode() if msg else ''}")
