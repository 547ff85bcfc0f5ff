"""Flat binary case files for the compiled host driver (host/obsgrid_host.cpp) and its build.

The C++ driver is the compiled host side above the C ABI (the reference is compiled Fortran,
whose toolchain this image lacks); Python only writes its input and reads its output."""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
HOST_SRC = ROOT / "host" / "obsgrid_host.cpp"
HOST_BIN = PKG / "lib" / "obsgrid_host"
FIELDS = ("t", "u", "v", "rh", "slp")
FLAGS = ("qc_u", "qc_v", "qc_t", "qc_rh", "qc_slp")
TABLES = ("xob", "yob", "lon", "elev", "ob_u", "ob_v", "ob_t", "ob_rh", "ob_slp")


def build_host(force: bool = False) -> Path:
    """g++ -std=c++17 host/obsgrid_host.cpp against libobsgrid_gpu.so (rpath $ORIGIN)."""
    deps = [HOST_SRC, ROOT / "include" / "obsgrid_gpu.hpp", ROOT / "include" / "obsgrid_gpu.h"]
    if not force and HOST_BIN.exists() and all(d.stat().st_mtime <= HOST_BIN.stat().st_mtime for d in deps):
        return HOST_BIN
    from . import build as b
    b.build()
    gxx = shutil.which("g++") or "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-I", str(ROOT / "include"), str(HOST_SRC), "-o",
           str(HOST_BIN), "-L", str(PKG / "lib"), "-lobsgrid_gpu", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed building obsgrid_host:\n" + r.stdout + r.stderr)
    return HOST_BIN


def write_case(case, path) -> None:
    with open(path, "wb") as f:
        f.write(b"OGC1")
        f.write(np.array([case.iew, case.jns, case.kbu, case.nrep], np.int32).tobytes())
        f.write(np.float32(case.dx_km).tobytes())
        f.write(bytes(case.qc_params))
        f.write(bytes(case.oa_params))
        f.write(np.ascontiguousarray(case.pressure, np.float32).tobytes())
        for nm in FIELDS + TABLES:
            f.write(np.ascontiguousarray(getattr(case, nm), np.float32).tobytes())
        for nm in FLAGS:
            f.write(np.ascontiguousarray(getattr(case, nm), np.int32).tobytes())


def read_result(path, case):
    """Returns a copy of `case` holding the fields and flags of a result file."""
    out = case.copy()
    with open(path, "rb") as f:
        for nm in FIELDS:
            a = getattr(out, nm)
            a[...] = np.frombuffer(f.read(a.nbytes), np.float32).reshape(a.shape)
        for nm in FLAGS:
            a = getattr(out, nm)
            a[...] = np.frombuffer(f.read(a.nbytes), np.int32).reshape(a.shape)
    return out


def run_host(case, workdir, *options):
    """Run the compiled driver on `case`; returns (exit code, result case or None, stderr)."""
    exe = build_host()
    workdir = Path(workdir)
    cin, cout = workdir / "case.bin", workdir / "result.bin"
    write_case(case, cin)
    r = subprocess.run([str(exe), str(cin), str(cout), *options], capture_output=True, text=True)
    res = read_result(cout, case) if r.returncode in (0, 5) and cout.exists() else None
    return r.returncode, res, r.stderr
