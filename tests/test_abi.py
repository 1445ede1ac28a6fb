"""CPU: the C-ABI library builds, loads and exports every symbol include/cuclark_b200.h declares."""
import ctypes
import os
import re

import pytest

from cuclark_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cuclark_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cuclark_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cuclark_b200 import build
    build.build_lib()
    lib = ctypes.CDLL(api.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cuclark_b200.h but not exported"
    assert set(names) == set(api.ABI_SYMBOLS)


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (it never routes to the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.CuclarkError) as e:
        api.CuClarkDB(31, 10)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_argument_validation():
    lib = api.load_library()
    h = ctypes.c_void_p()
    cfg = api.Config(33, api.HTSIZE_FULL, 0, 10, 0, 0, 0, 1, 0.0, 0)
    assert lib.cuclark_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"[2,32]" in lib.cuclark_last_error()
    cfg = api.Config(31, api.HTSIZE_FULL, 0, 0, 0, 0, 0, 1, 0.0, 0)
    assert lib.cuclark_create(ctypes.byref(cfg), ctypes.byref(h)) == -1


def test_product_does_not_import_oracle():
    """Nothing under cuclark_b200/ may import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "cuclark_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", ".hh")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "from oracle" not in text and "import oracle" not in text, f
                assert "cuclark_oracle" not in text, f


def test_key_width_dispatch():
    # src/main.cc:278-316
    assert api.key_bytes_for(31, api.HTSIZE_FULL) == 4
    assert api.key_bytes_for(23, api.HTSIZE_FULL) == 2
    assert api.key_bytes_for(32, api.HTSIZE_FULL) == 8
    assert api.key_bytes_for(27, api.HTSIZE_LIGHT) == 4
    assert api.key_bytes_for(20, api.HTSIZE_LIGHT) == 2


def test_plan_table_picks_the_layout_without_a_device(monkeypatch):
    """cuclark_plan_table: which device layout the loader takes for a database of a given size (no GPU needed)."""
    monkeypatch.delenv("CUCLARK_LAYOUT", raising=False)
    monkeypatch.delenv("CUCLARK_NO_LOCAL", raising=False)
    # BASELINE configs[1]: k=31, 5.72 G entries on one device -> minimizer lines at their minimum count (2^29 + 1 lines)
    p = api.plan_table(31, 5_719_957_086)
    assert p["layout"] == 3 and p["n_buckets"] == 4 * ((1 << 29) + 1) and p["home_bytes"] == p["n_buckets"] * 32
    # the 2,000-target reading (8 G entries): still LOCAL, 12 entries per line
    p = api.plan_table(31, 7_999_939_980)
    assert p["layout"] == 3 and 80e9 < p["home_bytes"] < 90e9
    # a shard of a table-partitioned run, and a small database, stay on the hashed sectors
    p = api.plan_table(31, 5_719_957_086, shard=(3, 8))
    assert p["layout"] == 1 and abs(p["n_local_buckets"] - p["n_buckets"] / 8) <= 1
    assert api.plan_table(27, 185_000, htsize=api.HTSIZE_LIGHT)["layout"] == 1
    assert api.plan_table(31, 1_000_000)["layout"] == 2          # k=31 small: 64-bit keys
    # forcing LOCAL where its minimum table (4^(k-7)/2^19 lines) would be mostly empty falls back to the hashed layouts;
    # at k=27 the minimum is 2^21 (+1) lines (268 MB) and is accepted
    assert api.plan_table(31, 1_000_000, layout=3)["layout"] == 2
    p = api.plan_table(27, 185_000, htsize=api.HTSIZE_LIGHT, layout=3)
    assert p["layout"] == 3 and p["n_buckets"] == 4 * ((1 << 21) + 1)
    # the automatic choice can be switched off
    monkeypatch.setenv("CUCLARK_NO_LOCAL", "1")
    assert api.plan_table(31, 5_719_957_086)["layout"] == 1


def test_ctypes_structs_match_the_header(tmp_path):
    """Every struct that crosses the C ABI has the same size in include/cuclark_b200.h (compiled with gcc as plain C)
    and in the ctypes mirror the tests and the bench use."""
    import subprocess
    pairs = [("cuclark_config", api.Config), ("cuclark_stats", api.Stats), ("cuclark_table_plan", api.TablePlan),
             ("cuclark_build_opts", api.BuildOpts), ("cuclark_build_stats", api.BuildStats),
             ("cuclark_text_opts", api.TextOpts), ("cuclark_text_stats", api.TextStats), ("cuclark_text_arrays", api.TextArrays),
             ("cuclark_route_stats", api.RouteStats)]
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "cuclark_b200.h"\nint main(void) {\n' +
                   "".join(f'  printf("{c} %zu\\n", sizeof({c}));\n' for c, _ in pairs) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    sizes = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, ct in pairs:
        assert int(sizes[cname]) == ctypes.sizeof(ct), (cname, sizes[cname], ctypes.sizeof(ct))
