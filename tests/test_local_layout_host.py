"""CPU: the LOCAL table layout's address arithmetic (cuclark_b200/csrc/common.cuh) on the host.

The functions are __host__ __device__; tests/shims/local_layout_shim.cu compiles them with nvcc for the host
and checks (a) that a k-mer's (line, key) identifies it exactly (local_rebuild inverts local_locate, 37-bit keys)
and (b) that the classify kernel's way to the home sector — minimizer found in READ orientation, leftmost or
rightmost smallest hash depending on the strand — lands where the table builder put the k-mer.
"""
import ctypes as C
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not found")
    so = str(tmp_path_factory.mktemp("shim") / "local_layout_shim.so")
    subprocess.check_call([NVCC, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "shims", "local_layout_shim.cu")])
    lib = C.CDLL(so)
    lib.shim_local_check.restype = C.c_long
    lib.shim_local_check.argtypes = [C.c_int, C.c_uint64, C.c_long, C.POINTER(C.c_int)]
    return lib


@pytest.mark.parametrize("k", [19, 21, 24, 27, 31, 32])
def test_locate_rebuild_and_kernel_route(shim, k):
    m = k - 8 + 1
    min_lines = 1 << max(0, 2 * m - 19)          # key width: mix(minimizer) div NL must fit 19 bits
    for nl in (max(min_lines, 1024), max(min_lines, 1000) | 1, (max(min_lines, 1000) * 3 // 2) | 1):   # power of two, odd, odd
        bad = C.c_int(-1)
        ties = shim.shim_local_check(k, nl, 300_000, C.byref(bad))
        assert bad.value == 0, f"k={k} NL={nl}: check {bad.value} failed at k-mer #{ties}"
        assert ties > 1000          # the periodic inputs do exercise the tie rule
