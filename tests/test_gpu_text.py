"""GPU parity of the device-side text stages (through the C ABI): raw FASTA/FASTQ bytes -> read index ->
2-bit containers -> classification -> CSV text, against the oracle's index/pack/classify/CSV and against the
CSVs the UNMODIFIED reference binaries wrote on a B200 (tests/golden/*.csv.gz)."""
import gzip
import os
import sys

import numpy as np
import pytest

from cuclark_b200.api import CuClarkDB, CuclarkError

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_gpu(case, **kw):
    g = CuClarkDB(case.k, case.n_targets, htsize=case.htsize, **kw)
    sz, ky, lb = case.arrays
    g.load_arrays(sz, ky, lb)
    return g


def oracle_all(oracle, odb, case, data, tmp_path, paired=False, extended=False, names=None, row_pairs=None):
    from oracle.binding import ReadIndex
    rp = row_pairs or case.maxhits
    ix, buf = oracle.index(data, 1)
    rix = ReadIndex(ix)
    ptr, cont = oracle.pack(ix, buf, case.k)
    final, rows, lookups = oracle.classify(odb, ptr, cont, case.n_targets, rp, threads=4)
    out = str(tmp_path / "oracle.csv")
    oracle.write_csv(out, ix, buf, case.k, paired, names or case.names, final, rows if extended else None, rp)
    oracle.free_index(ix)
    return rix, ptr, cont, final, rows, lookups, open(out, "rb").read()


def check_arrays(got, rix, ptr, cont, final, rows=None):
    for key in ("name_s", "name_e", "seq_s", "seq_e", "len"):
        assert np.array_equal(got[key], getattr(rix, key).astype(np.uint64)), key
    assert np.array_equal(got["reads_ptr"], ptr)
    assert np.array_equal(got["containers"], cont)
    assert np.array_equal(got["final5"], final)
    if rows is not None:
        assert np.array_equal(got["rows"], rows)


@pytest.mark.parametrize("chunk", [0, 1 << 16])
def test_light_small_fasta_index_pack_csv(oracle, light_small, tmp_path, chunk):
    """FASTA, one chunk and many small chunks: every intermediate array and the CSV bytes."""
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    data = c.reads_bytes
    rix, ptr, cont, final, rows, lookups, csv = oracle_all(oracle, odb, c, data, tmp_path)
    ref = gzip.open(os.path.join(GOLDEN, "light_small.csv.gz")).read()
    assert csv == ref
    with make_gpu(c) as g:
        got, st = g.text_debug(data, rix.n + 10, cont.size + 10, chunk_bytes=chunk, want_rows=True)
        assert st["n_reads"] == rix.n and st["lookups"] == lookups
        if chunk:
            assert st["n_chunks"] > 10
        check_arrays(got, rix, ptr, cont, final, rows)
        out, st2 = g.classify_text(data, names=c.names, chunk_bytes=chunk)
        assert out == ref                                   # byte-identical to the reference binary's CSV
        assert st2["n_reads"] == rix.n and st2["csv_bytes"] == len(ref)


def test_config1_csv_equals_reference_binary(light_c1):
    """BASELINE.json configs[0]: 100k x 100 bp reads, CSV == the reference cuCLARK-l output on B200."""
    c = light_c1
    ref = gzip.open(os.path.join(GOLDEN, "light_c1.csv.gz")).read()
    with make_gpu(c) as g:
        out, st = g.classify_text(c.reads_bytes, names=c.names, chunk_bytes=1 << 20, n_slots=3)
        assert out == ref
        assert st["n_reads"] == ref.count(b"\n") - 1 and st["lookups"] == 7380778
        out2, _ = g.classify_text(c.reads_bytes, names=c.names)          # default chunking
        assert out2 == ref


@pytest.mark.slow
def test_full_fastq_csv_equals_reference_binary(full_small, tmp_path):
    """cuCLARK (full, k=31) on FASTQ input; also through cuclark_classify_file."""
    c = full_small
    ref = gzip.open(os.path.join(GOLDEN, "full_small.csv.gz")).read()
    with make_gpu(c) as g:
        for chunk in (0, 1 << 17):
            out, st = g.classify_text(c.reads_bytes, names=c.names, chunk_bytes=chunk)
            assert out == ref
        src = tmp_path / "reads.fq"
        src.write_bytes(c.reads_bytes)
        st = g.classify_file(str(src), str(tmp_path / "res.csv"), names=c.names, chunk_bytes=1 << 18)
        assert (tmp_path / "res.csv").read_bytes() == ref
        with pytest.raises(CuclarkError) as e:
            g.classify_file(str(tmp_path / "missing.fq"), str(tmp_path / "x.csv"), names=c.names)
        assert e.value.code == -4 and "Failed to open" in str(e.value)


def edge_records(c):
    k = c.k
    asc = np.frombuffer(b"ACGT", np.uint8)
    mg = sys.modules["make_golden"]
    g0 = asc[mg.target_codes(c.case, 0)].tobytes()
    g3 = asc[mg.target_codes(c.case, 3)].tobytes()
    return [
        b">short\nACGT\n",
        b">empty\n\n",
        b">nothing\n",
        b">allN\n" + b"N" * 80 + b"\n",
        b">exactk\n" + g0[:k] + b"\n",
        b">kminus1\n" + g0[:k - 1] + b"\n",
        b">multi line\textra words\n" + g0[0:60] + b"\n" + g0[60:130] + b"\n" + g0[130:216] + b"\n",
        b">parts\n" + g0[0:40] + b"N" + g0[41:60] + b"NN" + g3[108:216] + b"n" + g3[216:220] + b"\n",
        b">lower\n" + g0[0:216].lower() + b"\n",
        b">rna\n" + g0[0:216].replace(b"T", b"U") + b"\n",
        b">cr\n" + g0[0:108] + b"\r\n" + g0[108:216] + b"\r\n",
        b">two_targets\n" + g0[0:216] + g3[0:324] + b"\n",
        b">a_very_long_read_name_that_exceeds_the_thirty_nine_characters_limit x\n" + g3[0:150] + b"\n",
        b">gt_inside>name\n" + g3[100:250] + b"\n",
        b">blank_lines\n\n" + g0[300:400] + b"\n\n" + g0[400:450] + b"\n\n",
        b">long\n" + b"\n".join(g3[i:i + 70] for i in range(1000, 9000, 70)) + b"\n",
        b">tail_no_newline\n" + g3[0:300],
    ]


@pytest.mark.parametrize("chunk", [0, 4096])
def test_fasta_edge_cases(oracle, light_small, tmp_path, chunk):
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    recs = edge_records(c)
    for data in (b"".join(recs), b"".join(recs[:-1]), recs[2], recs[0] + recs[2]):
        rix, ptr, cont, final, rows, lookups, csv = oracle_all(oracle, odb, c, data, tmp_path)
        if chunk and len(data) > chunk:
            chunk_eff = 1 << 14          # the longest record (8.2 kB) must fit one chunk
        else:
            chunk_eff = chunk
        with make_gpu(c) as g:
            got, st = g.text_debug(data, rix.n + 4, cont.size + 16, chunk_bytes=chunk_eff, want_rows=True)
            assert st["n_reads"] == rix.n
            check_arrays(got, rix, ptr, cont, final, rows)
            out, _ = g.classify_text(data, names=c.names, chunk_bytes=chunk_eff)
            assert out == csv
            ext_csv = oracle_all(oracle, odb, c, data, tmp_path, extended=True)[6]
            out, _ = g.classify_text(data, names=c.names, chunk_bytes=chunk_eff, extended=True)
            assert out == ext_csv
            pair_csv = oracle_all(oracle, odb, c, data, tmp_path, paired=True)[6]
            out, _ = g.classify_text(data, names=c.names, chunk_bytes=chunk_eff, paired=True)
            assert out == pair_csv


def fastq_from(recs_fa, qual_at=False):
    """FASTA-style (name, seq) -> 4-line FASTQ; quality lines may start with '@' to stress the boundary search."""
    out = []
    for i, (name, seq) in enumerate(recs_fa):
        q = (b"@" if qual_at and i % 2 == 0 else b"I") + b"I" * max(0, len(seq) - 1) if seq else b""
        out.append(b"@" + name + b"\n" + seq + b"\n+\n" + q + b"\n")
    return b"".join(out)


@pytest.mark.parametrize("chunk", [0, 4096])
def test_fastq_edge_cases(oracle, light_small, tmp_path, chunk):
    c = light_small
    k = c.k
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    asc = np.frombuffer(b"ACGT", np.uint8)
    mg = sys.modules["make_golden"]
    g1 = asc[mg.target_codes(c.case, 1)].tobytes()
    g6 = asc[mg.target_codes(c.case, 6)].tobytes()
    recs = [(b"r%d/1 extra" % i, (g1 if i % 3 else g6)[i * 11:i * 11 + 60 + (i * 37) % 200]) for i in range(300)]
    recs += [(b"withN", g1[0:50] + b"N" + g1[51:120]), (b"short", b"ACGT"), (b"emptyseq", b""),
             (b"exactk", g6[0:k]), (b"lower", g6[500:700].lower())]
    base = fastq_from(recs, qual_at=True)
    variants = [base, base[:-1], base + b"\n", base + b"@partial\n" + g1[0:100], base + b"@p2\n" + g1[0:100] + b"\n+\n"]
    for data in variants:
        rix, ptr, cont, final, rows, lookups, csv = oracle_all(oracle, odb, c, data, tmp_path)
        with make_gpu(c) as g:
            got, st = g.text_debug(data, rix.n + 4, cont.size + 16, chunk_bytes=chunk, want_rows=True)
            assert st["n_reads"] == rix.n
            if chunk:
                assert st["n_chunks"] > 5
            check_arrays(got, rix, ptr, cont, final, rows)
            out, _ = g.classify_text(data, names=c.names, chunk_bytes=chunk)
            assert out == csv


def test_text_errors(light_small):
    c = light_small
    with make_gpu(c) as g:
        with pytest.raises(CuclarkError) as e:
            g.classify_text(b"ACGT\nACGT\n", names=c.names)
        assert e.value.code == -8 and "Failed to recognize the format" in str(e.value)
        # a record larger than chunk_bytes is taken whole (the slot grows; src/CuCLARK_hh.hh:1377-1389 takes any size)
        big = b">r\n" + b"ACGT" * 4096 + b"\n>r2\nACGT\n"
        out, st = g.classify_text(big, names=c.names, chunk_bytes=4096)
        assert st["n_reads"] == 2 and out.split(b"\n")[1].startswith(b"r,16384,") and out.endswith(b"r2,4,-0,NA,0,NA,0,0\n")
        # the handle stays usable after an error
        out, st = g.classify_text(b">r\nACGT\n", names=c.names)
        assert out.endswith(b"r,4,-0,NA,0,NA,0,0\n") and st["n_reads"] == 1


def test_many_targets_extended_and_long_names(oracle, tmp_path):
    """> MAXHITS targets in a read (rows truncated as the oracle defines), long target names, extended columns."""
    from cuclark_b200 import synth
    from cuclark_b200.api import HTSIZE_LIGHT
    from oracle import dbtools
    k, T, G = 27, 300, 2000
    targets = [synth.genome_codes(77, t, 0, G) for t in range(T)]
    kmers, labels = dbtools.build_entries(targets, k, 0)
    sz, ky, lb = dbtools.entries_to_arrays(kmers, labels, HTSIZE_LIGHT, 4)
    odb = oracle.db_from_arrays(HTSIZE_LIGHT, k, sz, ky, lb)
    asc = np.frombuffer(b"ACGT", np.uint8)
    seg = lambda t, n: asc[targets[t][100:100 + n]].tobytes()
    reads = [b">r70\n" + b"".join(seg(t, 40) for t in range(0, 70)) + b"\n",
             b">r20\n" + b"".join(seg(t, 45) for t in range(10, 30)) + b"\n",
             b">r1\n" + seg(5, 200) + b"\n", b">none\n" + b"ACGT" * 20 + b"\n"]
    data = b"".join(reads)
    names = ["target_with_a_rather_long_name_%05d" % t for t in range(T)]

    class Case:
        pass
    c = Case()
    c.k, c.n_targets, c.maxhits, c.names = k, T, 23, names
    for ext in (False, True):
        csv = oracle_all(oracle, odb, c, data, tmp_path, extended=ext, names=names)[6]
        with CuClarkDB(k, T, htsize=HTSIZE_LIGHT) as g:
            g.load_arrays(sz, ky, lb)
            out, st = g.classify_text(data, names=names, extended=ext)
            assert out == csv
