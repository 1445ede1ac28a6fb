"""CPU: the product's integer-only "%g" (cuclark_b200/csrc/fmt_g.h, shared by host and device code)
against glibc's printf("%g"), which is what the reference's CSV writer uses
(src/CuCLARK_hh.hh:2132-2135)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "fmt_g_shim.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "shims", "fmt_g_shim.cc")])
    lib = C.CDLL(so)
    lib.shim_fmt_g.argtypes = [C.c_double, C.c_char_p]
    lib.shim_gamma.argtypes = [C.c_uint, C.c_uint, C.c_int, C.c_char_p]
    lib.shim_conf.argtypes = [C.c_uint, C.c_uint, C.c_char_p]
    lib.shim_sweep.restype = C.c_long
    lib.shim_sweep.argtypes = [C.c_uint] * 4 + [C.POINTER(C.c_uint)] * 2
    return lib


def g(lib, d):
    buf = C.create_string_buffer(32)
    lib.shim_fmt_g(d, buf)
    return buf.value.decode()


def test_known_values(shim):
    for d, s in [(0.0, "0"), (-0.0, "-0"), (1.0, "1"), (0.5, "0.5"), (1 / 3, "0.333333"), (2 / 3, "0.666667"),
                 (100000.0, "100000"), (999999.5, "1e+06"), (1e-5, "1e-05"), (0.0001, "0.0001"),
                 (0.00012345678, "0.000123457"), (13 / 128, "0.101562"), (1 / 512, "0.00195312"),
                 (40001 / 400000, "%g" % (40001 / 400000)), (65535.0, "65535"), (123456.7, "123457"),
                 (1 / 4e9, "2.5e-10"), (float("inf"), "inf"), (float("-inf"), "-inf")]:
        assert g(shim, d) == s, (d, g(shim, d), s)
    assert g(shim, float("nan")) in ("nan", "-nan")


def test_gamma_and_confidence_special_cases(shim):
    buf = C.create_string_buffer(32)
    shim.shim_gamma(0, 26, 27, buf)          # Length == k-1: 0/0 -> "-nan" as the x86 host prints it
    assert buf.value == b"-nan"
    shim.shim_gamma(0, 10, 27, buf)          # shorter: 0/negative -> "-0"
    assert buf.value == b"-0"
    shim.shim_gamma(0, 27, 27, buf)
    assert buf.value == b"0"
    shim.shim_gamma(74, 100, 27, buf)
    assert buf.value == b"1"
    shim.shim_conf(0, 0, buf)
    assert buf.value == b"0"
    shim.shim_conf(14, 14, buf)
    assert buf.value == b"0.5"


def test_sweep_all_small_ratios(shim):
    """Every a/b the CSV can print for reads up to 1500 k-mers, plus long-read denominators."""
    ba, bb = C.c_uint(), C.c_uint()
    assert shim.shim_sweep(0, 1501, 1, 1501, C.byref(ba), C.byref(bb)) == 0, (ba.value, bb.value)
    assert shim.shim_sweep(1, 400, 399_000, 401_000, C.byref(ba), C.byref(bb)) == 0, (ba.value, bb.value)
    assert shim.shim_sweep(39_990, 40_010, 399_990, 400_010, C.byref(ba), C.byref(bb)) == 0, (ba.value, bb.value)
    assert shim.shim_sweep(65_000, 65_536, 1, 3000, C.byref(ba), C.byref(bb)) == 0, (ba.value, bb.value)
    assert shim.shim_sweep(1, 50, 4_294_000_000, 4_294_001_000, C.byref(ba), C.byref(bb)) == 0, (ba.value, bb.value)


def test_random_doubles(shim):
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.random(20000), rng.random(20000) * 1e-6, rng.random(20000) * 65535,
                         10.0 ** rng.uniform(-12, 8, 20000)])
    for d in xs:
        assert g(shim, float(d)) == "%g" % d, d
