"""Boundary (b), proven by compiling it: the reference's OWN orchestrator (src/main.cc + src/CuCLARK_hh.hh, unmodified,
compiled where they lie) linked against libcuclark_b200.so through integration/CuClarkDB_b200.cc, which implements the
reference's `CuClarkDB<HKMERr>` class (src/CuClarkDB.cuh:98-150) over the C ABI. oracle/Makefile `make adapter`
builds oracle/_ref/cuCLARK[-l]_b200adapter in this container (test infrastructure; the binaries travel to the GPU box).

Here the reference indexes and packs the reads, calls malloc / readyBatch / queryBatch / waitForBatch and prints the
CSV itself; only the device side is ours. The CSV must equal the unmodified reference binary's (tests/golden/*.csv.gz).
"""
import gzip
import os
import subprocess

import pytest

from conftest import GOLDEN, has_gpu
from test_cli import setup_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def adapter(light: bool) -> str:
    p = os.path.join(REF, "cuCLARK-l_b200adapter" if light else "cuCLARK_b200adapter")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref adapter binaries not built (need /root/reference at build time)")
    return p


def run(cmd, cwd=None, env=None):
    return subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=900, env=env)


def test_adapter_links_and_fails_loudly_without_gpu(light_small, tmp_path):
    exe = adapter(True)
    p = run([exe, "--version"])
    assert p.returncode == 0 and p.stdout.startswith("Version: 1.1")
    if has_gpu():
        return
    reads = setup_case(light_small, str(tmp_path))
    p = run([exe, "-T", "targets.txt", "-D", "db/", "-O", os.path.basename(reads), "-R", "out"], cwd=str(tmp_path))
    assert p.returncode == 1 and "Not enough CUDA devices found" in p.stderr
    assert not os.path.exists(tmp_path / "out.csv")


@pytest.mark.gpu
@pytest.mark.parametrize("case_name,args", [("light_small", ["-n", "1"]), ("light_small", ["-n", "4", "-b", "7"]),
                                            ("light_c1", ["-n", "8"]), ("full_small", ["-n", "3"])])
def test_reference_orchestrator_over_the_library(request, tmp_path, case_name, args):
    case = request.getfixturevalue(case_name)
    exe = adapter(case.light)
    reads = setup_case(case, str(tmp_path))
    p = run([exe, "-T", "targets.txt", "-D", "db/", "-O", os.path.basename(reads), "-R", "out", *args], cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr[-2000:]
    assert "through libcuclark_b200" in p.stderr
    ref = gzip.open(os.path.join(GOLDEN, case_name + ".csv.gz")).read()
    assert (tmp_path / "out.csv").read_bytes() == ref
    # (the reference's own speed line reports "0 objects": clearReadData() zeroes m_nbObjects before printSpeedStats,
    #  src/CuCLARK_hh.hh:319-321, 1787, 1942 — the orchestrator is unmodified, so the adapter build prints the same)
    assert " objects/min. (" in p.stdout and " - Results stored in out.csv" in p.stdout


@pytest.mark.gpu
def test_reference_orchestrator_extended(light_small, tmp_path):
    """--extended: the reference's writer expands our sparse rows into one column per target (src/CuCLARK_hh.hh:2014-2031);
    the product's own CLI must print the same bytes."""
    exe = adapter(True)
    reads = setup_case(light_small, str(tmp_path))
    cmd = ["-T", "targets.txt", "-D", "db/", "-O", os.path.basename(reads), "--extended"]
    p = run([exe, *cmd, "-R", "ref_orch", "-n", "2"], cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr[-2000:]
    ours = os.path.join(ROOT, "cuclark_b200", "bin", "cuCLARK-l")
    q = run([ours, *cmd, "-R", "ours"], cwd=str(tmp_path))
    assert q.returncode == 0, q.stderr[-2000:]
    assert (tmp_path / "ref_orch.csv").read_bytes() == (tmp_path / "ours.csv").read_bytes()
