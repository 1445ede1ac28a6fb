"""Build the C++ command-line executables cuCLARK / cuCLARK-l (csrc/cli_main.cc) against
libcuclark_b200.so. Like the reference (src/Makefile:26-34) the two variants are the same source
compiled with a different HTSIZE; outputs go to cuclark_b200/bin/ (git-ignored, shipped by gpurun).
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "cli_main.cc")
BINDIR = os.path.join(HERE, "bin")
LIBDIR = os.path.join(HERE, "lib")
CXX = os.environ.get("CXX", "/usr/bin/g++")
TARGETS = {"cuCLARK": [], "cuCLARK-l": ["-DCUCLARK_LIGHT"]}


def build(force: bool = False) -> list[str]:
    os.makedirs(BINDIR, exist_ok=True)
    deps = [SRC, os.path.join(HERE, "..", "include", "cuclark_b200.h"), os.path.join(LIBDIR, "libcuclark_b200.so")]
    outs = []
    for name, defs in TARGETS.items():
        out = os.path.join(BINDIR, name)
        outs.append(out)
        if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
            continue
        subprocess.check_call([CXX, "-O2", "-std=c++17", "-Wall", *defs, "-o", out, SRC, "-L" + LIBDIR, "-lcuclark_b200",
                               "-Wl,-rpath,$ORIGIN/../lib", "-lpthread"])
    return outs


if __name__ == "__main__":
    print("\n".join(build(force=True)))
