"""Build libcuclark_b200.so (sm_100a) and the host tools in-tree with nvcc/g++.

Run as ``python -m cuclark_b200.build``; ``__graft_entry__.build()`` calls
``build_all()``. Outputs go to cuclark_b200/lib/ and cuclark_b200/bin/ (git-ignored,
but they travel to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
BINDIR = os.path.join(HERE, "bin")
LIB = os.path.join(LIBDIR, "libcuclark_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CU_SOURCES = ["table.cu", "classify.cu", "route.cu", "capi.cu", "textpipe.cu", "stream.cu", "dbbuild.cu"]
HEADERS = ["common.cuh", "internal.h", "hits.cuh", "kmerwin.cuh", "synth.cuh", "textpipe.cuh", "fmt_g.h", os.path.join("..", "..", "include", "cuclark_b200.h")]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS]
    if force or _newer(LIB, deps):
        cmd = [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-O3", "-shared",
               "-ccbin", "/usr/bin/g++", "-o", LIB, *srcs]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        subprocess.check_call(cmd)
    return LIB


def build_variant(name: str, defines: list[str], verbose: bool = False) -> str:
    """The same library with other compile-time choices (-DCUCLARK_...=...), for A/B timing on the GPU box:
    cuclark_b200/lib/variants/<name>.so, picked up through CUCLARK_LIB."""
    out_dir = os.path.join(LIBDIR, "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, name + ".so")
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    cmd = [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-O3", "-shared",
           "-ccbin", "/usr/bin/g++", "-o", out, *[f"-D{d}" for d in defines], *srcs]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    subprocess.check_call(cmd)
    return out


def build_all(verbose: bool = False, force: bool = False) -> None:
    build_lib(verbose, force)
    try:
        from . import hostbuild  # C++ host tools (CLI); optional until present
    except ImportError:
        return
    hostbuild.build(force=force)


if __name__ == "__main__":
    build_all(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(LIB)
