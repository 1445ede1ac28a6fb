"""CPU: the oracle's result CSV against the CSV the UNMODIFIED reference binary wrote on a B200.

tests/golden/<case>.csv.gz were produced by `tests/golden/make_golden.py --stage csv` under gpurun
(reference cuCLARK / cuCLARK-l from oracle/_ref, run on the GPU box) and committed unchanged.
Byte-identical CSVs pin stages 1-4 of the oracle (index, pack, extract, lookup, histogram, top-2,
gamma/confidence formatting) to the reference's own output.
"""
import gzip
import os

import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def oracle_csv(oracle, case, tmp_path, n_batches, extended=False):
    sz, ky, lb = case.arrays
    odb = oracle.db_from_arrays(case.htsize, case.k, sz, ky, lb)
    ix, buf = oracle.index(case.reads_bytes, n_batches)
    ptr, cont = oracle.pack(ix, buf, case.k)
    final, rows, _ = oracle.classify(odb, ptr, cont, case.n_targets, case.maxhits, threads=4)
    out = str(tmp_path / "out.csv")
    oracle.write_csv(out, ix, buf, case.k, False, case.names, final, rows if extended else None, case.maxhits)
    oracle.free_index(ix)
    return open(out, "rb").read()


@pytest.mark.parametrize("name,n_batches", [("light_small", 1), ("light_small", 7), ("light_c1", 4)])
def test_csv_equals_reference_gpu_output(oracle, request, tmp_path, name, n_batches):
    case = request.getfixturevalue(name)
    ref = gzip.open(os.path.join(GOLDEN, f"{name}.csv.gz")).read()
    assert oracle_csv(oracle, case, tmp_path, n_batches) == ref


@pytest.mark.slow
def test_csv_equals_reference_gpu_output_full(oracle, full_small, tmp_path):
    ref = gzip.open(os.path.join(GOLDEN, "full_small.csv.gz")).read()
    assert oracle_csv(oracle, full_small, tmp_path, 4) == ref
