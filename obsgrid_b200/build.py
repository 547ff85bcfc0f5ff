"""Build libobsgrid_gpu.so in-tree with nvcc for sm_100a (no torch, no CMake).

    python -m obsgrid_b200.build [--force]

Flags: -fmad=false -prec-div=true -prec-sqrt=true -ftz=false keep every FP32 operation a
single IEEE-754 round-to-nearest operation in source order, like the reference's gfortran
build (arch/configure.defaults:44-46); -lineinfo keeps the ncu source page usable.
"""
from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = PKG / "lib" / "libobsgrid_gpu.so"
SOURCES = ["og_capi.cu", "og_time.cu", "og_oa.cu", "og_qc.cu", "og_bin.cu", "og_density.cu", "og_final.cu", "og_misc.cu", "og_reports.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math", "-Xcompiler", "-ffp-contract=off",
    "-shared", "-cudart", "static",
]


def needs_build() -> bool:
    if not OUT.exists():
        return True
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [
        PKG.parent / "include" / "obsgrid_gpu.h", PKG.parent / "include" / "obsgrid_reports.h"]
    t = OUT.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    OUT.parent.mkdir(parents=True, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-o", str(OUT)] + [str(CSRC / s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libobsgrid_gpu.so")
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
