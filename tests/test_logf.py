"""LOG of the DEWPOINT quantities (proc_qc.F90:315-319, ob_access_module.F90:476-486).

The reference build evaluates Fortran's LOG with glibc's logf; QC flags are compared bit for bit,
so the device restates that routine (obsgrid_b200/csrc/og_logf.h: glibc >= 2.28 / ARM optimized
routines logf).  CPU: the host build of the very same header against the running libm.  GPU: the
device build against the running libm, on every float of the argument range the path produces
(rh/100 with rh clamped below at 1e-4, i.e. [1e-6, ~1.2]) and well beyond it.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

HERE = Path(__file__).resolve().parent


@pytest.fixture(scope="module")
def hostlog(tmp_path_factory):
    out = tmp_path_factory.mktemp("logf") / "libhostlogf.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-mfma", "-o", str(out),
                           str(HERE / "host_logf_check.cpp"), "-lm"])
    l = C.CDLL(str(out))
    l.check_logf_sweep.restype = C.c_longlong
    l.check_logf_sweep.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_uint)]
    return l


def bits(x):
    return int(np.float32(x).view(np.uint32))


def test_host_build_equals_libm_on_the_dewpoint_range(hostlog):
    """Every float in [1e-7, 4): rh/100 (>= 1e-6), and room around it."""
    first = C.c_uint(0)
    bad = hostlog.check_logf_sweep(bits(1e-7), bits(4.0), 1, C.byref(first))
    assert bad == 0, hex(first.value)


def test_host_build_equals_libm_over_all_binades(hostlog):
    """A stride through every positive normal float (all exponents, k from -126 to 127)."""
    first = C.c_uint(0)
    bad = hostlog.check_logf_sweep(0x00800000, 0x7f800000, 61, C.byref(first))
    assert bad == 0, hex(first.value)


def test_oracle_dewpoint_uses_libm(oracle):
    x = np.array([1e-6, 0.37, 0.5, 0.999, 1.0], np.float32)
    np.testing.assert_allclose(oracle.logf(x), np.log(x.astype(np.float64)), rtol=2e-7, atol=1e-9)
    t, rh = np.float32(281.5), np.float32(37.0)
    want = np.float32(1) / (np.float32(1) / t - (np.float32(1) / np.float32(5418.12)) * oracle.logf([rh / np.float32(100)])[0])
    assert oracle.dewpoint(t, rh) == want


@pytest.mark.gpu
def test_device_logf_equals_libm(ctx, oracle):
    lo, hi = bits(1e-7), bits(4.0)
    step = 1 << 24
    for b in range(lo, hi, step):
        x = np.arange(b, min(b + step, hi), dtype=np.uint32).view(np.float32)
        a = oracle.logf(x)
        d = ctx.selftest_logf(x)
        bad = np.flatnonzero(a.view(np.uint32) != d.view(np.uint32))
        assert bad.size == 0, (x[bad[:4]], a[bad[:4]], d[bad[:4]])
    # all binades
    x = np.arange(0x00800000, 0x7f800000, 127, dtype=np.uint32).view(np.float32)
    np.testing.assert_array_equal(oracle.logf(x).view(np.uint32), ctx.selftest_logf(x).view(np.uint32))
