"""The drop-in boundary without a GPU: libobsgrid_gpu.so loads, exports every function
include/obsgrid_gpu.h declares, the ctypes binding covers all of them, and the product path
fails loudly (no CPU fallback) when there is no sm_100 device."""
import ctypes as C
import re
from pathlib import Path

import pytest

from obsgrid_b200 import _capi

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "obsgrid_gpu.h").read_text()
HEADER_REPORTS = (ROOT / "include" / "obsgrid_reports.h").read_text()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER + HEADER_REPORTS, flags=re.S)
    body = re.sub(r"typedef\s+struct[^{;]*\{.*?\}\s*\w+\s*;", "", body, flags=re.S)
    return sorted(set(re.findall(r"\b(og_[a-z0-9_]+)\s*\(", body)))


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_declares_the_documented_surface():
    names = declared_functions()
    for must in ("og_init", "og_finalize", "og_last_error", "og_innovation", "og_cressman",
                 "og_oa_slab", "og_local_time", "og_error_max", "og_buddy_check", "og_qc_slab",
                 "og_qc_time", "og_qc_time_levels", "og_oa_time", "og_dev_qc_time", "og_dev_oa_time"):
        assert must in names, must


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(str(_capi.LIB_PATH))
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.og_abi_version() == 3


def test_ctypes_binding_covers_the_header():
    assert set(declared_functions()) <= set(_capi.SIGNATURES), \
        sorted(set(declared_functions()) - set(_capi.SIGNATURES))


def test_fortran_shim_binds_existing_symbols():
    shim = (ROOT / "fortran" / "obsgrid_gpu_iface.F90").read_text()
    bound = set(re.findall(r"NAME\s*=\s*'(og_[a-z_]+)'", shim))
    assert bound and bound <= set(declared_functions()), bound - set(declared_functions())


@pytest.mark.skipif(has_gpu(), reason="needs a box without a GPU")
def test_no_cpu_fallback():
    lib = _capi.lib()
    h = C.c_void_p()
    rc = lib.og_init(C.byref(h), 0)
    assert rc == -1 and not h.value                      # OG_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.og_last_error() or b"sm_100" in lib.og_last_error()
    from obsgrid_b200.api import Context
    with pytest.raises(_capi.OgError):
        Context(0)


def test_null_context_is_rejected_not_dereferenced():
    lib = _capi.lib()
    assert lib.og_sync(None) == -2 and lib.og_set_counting(None, 1) == -2     # OG_ERR_INVALID


def _c_struct_members(name):
    m = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (name, name), HEADER, flags=re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(int|float)\b", "", decl)
        for part in decl.split(","):
            out.append(re.sub(r"[\s\*]|\bconst\b|\bint\b|\bfloat\b|\[.*?\]", "", part))
    return [x for x in out if x]


def _fortran_type_members(name):
    shim = (ROOT / "fortran" / "obsgrid_gpu_iface.F90").read_text()
    m = re.search(r"TYPE\s*,\s*BIND\s*\(\s*C\s*\)\s*::\s*%s\b(.*?)END TYPE %s" % (name, name), shim, flags=re.S)
    out = []
    for line in m.group(1).splitlines():
        line = line.split("!")[0]
        if "::" not in line:
            continue
        for part in line.split("::", 1)[1].split(","):
            part = re.sub(r"\(.*?\)", "", part).strip()
            if part:
                out.append(part)
    return out


@pytest.mark.parametrize("name", ["og_qc_params", "og_oa_params", "og_time_desc"])
def test_fortran_derived_types_mirror_the_c_structs(name):
    """Same members in the same order (types are all 4-byte or pointers; BIND(C) pads as C does)."""
    assert _fortran_type_members(name) == _c_struct_members(name)
