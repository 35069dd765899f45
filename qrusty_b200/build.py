#!/usr/bin/env python
"""Builds qrusty_b200/lib/libqrusty_cuda.so in-tree with nvcc for sm_100a.

Run as a script (`python qrusty_b200/build.py`): importing the package itself needs
the library to exist already.  nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
SRC = PKG / "csrc" / "qrusty_cuda.cu"
OUT = PKG / "lib" / "libqrusty_cuda.so"
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-diag-suppress", "186",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-shared",
]


def newest_source_mtime():
    files = list((PKG / "csrc").glob("*.cu*")) + [PKG.parent / "include" / "qrusty_cuda.h"]
    return max(f.stat().st_mtime for f in files)


def build(force=False, verbose=False):
    if not force and OUT.exists() and OUT.stat().st_mtime >= newest_source_mtime():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        if OUT.exists():        # GPU box without a toolkit on PATH: use the shipped .so
            return OUT
        raise RuntimeError("nvcc not found and no prebuilt libqrusty_cuda.so")
    OUT.parent.mkdir(exist_ok=True)
    cmd = [nvcc, *FLAGS, "-o", str(OUT), str(SRC), "-ldl"]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=str(PKG / "csrc"))
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p, os.path.getsize(p), "bytes")
