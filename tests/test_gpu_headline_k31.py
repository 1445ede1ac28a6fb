"""GPU parity of the kernel instantiations the headline number is quoted on: k=31 through the NARROW
and the LOCAL table (kshift 2, 48-bit m-mers, nl_sh 16, 19-bit quotient), against the oracle.

At k=31 a small database gets the WIDE layout (test_gpu_parity.py::test_full_variant_k31_wide); NARROW
needs M > 2^30 buckets (34 GB) and LOCAL 2^29 lines (68.7 GB) whatever they hold. Here both are FORCED on
the small golden database (`layout=` + CUCLARK_ALLOW_SPARSE_TABLE=1, csrc/table.cu choose_geometry): the
tables are mostly empty, the address arithmetic, key packing and probe path are those of the
bacterial-scale run. Also: k=32 / 8-byte `.ky` keys (the T64 branch, src/main.cc:278-316), k=30 with
8-byte keys on the light HTSIZE, and parts of 65,536 nt or more (deviation Q8).
"""
import numpy as np
import pytest

from cuclark_b200 import synth
from cuclark_b200.api import CuClarkDB, HTSIZE_FULL, HTSIZE_LIGHT
from oracle import dbtools
from oracle.binding import key_bytes_for

from test_gpu_local_layout import low_complexity_reads
from test_gpu_parity import make_gpu, oracle_expect, pack_kmers_as_reads
from test_gpu_text import oracle_all

pytestmark = pytest.mark.gpu


@pytest.mark.slow
@pytest.mark.parametrize("layout", [1, 3])
def test_k31_headline_instantiations(oracle, full_small, layout, monkeypatch):
    """full_small (golden, k=31) + low-complexity reads + every DB k-mer in both strands + random k-mers."""
    monkeypatch.setenv("CUCLARK_ALLOW_SPARSE_TABLE", "1")
    c = full_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    rng = np.random.default_rng(31)
    # reads of the golden case, then periodic reads (tie k-mers), then a long two-strand read
    import sys
    mg = sys.modules["make_golden"]
    asc = np.frombuffer(b"ACGT", np.uint8)
    t2 = mg.target_codes(c.case, 2)
    t3 = mg.target_codes(c.case, 3)
    long_codes = np.concatenate([t2[:3000], 3 - t3[::-1][:2500]])
    fasta = low_complexity_reads(rng, 400, 150) + b">long\n" + asc[long_codes].tobytes() + b"\n"
    p1, c1, f1, r1, l1 = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits)
    p2, c2, f2, r2, l2 = oracle_expect(oracle, odb, c.k, fasta, c.n_targets, c.maxhits)
    probe = np.concatenate([c.kmers, dbtools.revcomp_codes(c.kmers, c.k),
                            rng.integers(0, 1 << 62, 100_000, dtype=np.uint64)])
    expect, _ = odb.query(probe)
    p3, c3 = pack_kmers_as_reads(probe, c.k)
    with make_gpu(c, layout=layout) as g:
        st = g.stats()
        assert st["layout"] == layout and st["n_entries"] == c.kmers.size
        if layout == 3:
            assert st["n_buckets"] >= 4 << 29          # 2^29 lines of 4 sectors: the bacterial-scale geometry
        else:
            assert st["n_buckets"] > 1 << 30
        gf, gr = g.classify(p1, c1, want_rows=True)
        assert g.stats()["lookups"] == l1
        assert np.array_equal(gf, f1) and np.array_equal(gr, r1)
        gf, gr = g.classify(p2, c2, want_rows=True)
        assert g.stats()["lookups"] == l2
        assert np.array_equal(gf, f2) and np.array_equal(gr, r2)
        gf, _ = g.classify(p3, c3)
    got = np.where(gf[:, 2] > 0, gf[:, 1].astype(np.int32) - 1, -1)
    assert np.array_equal(got, expect)
    assert (f1[:, 2] > 100).sum() > 10000 and (expect >= 0).sum() == 2 * c.kmers.size


def _small_case(k, htsize, T=6, G=30_000, seed=17):
    targets = [synth.genome_codes(seed, t, 0, G) for t in range(T)]
    kmers, labels = dbtools.build_entries(targets, k, 0)
    kb = key_bytes_for(k, htsize)
    return targets, kmers, labels, dbtools.entries_to_arrays(kmers, labels, htsize, kb), kb


def _reads_from(targets, rng, n, k):
    out = []
    T, G = len(targets), targets[0].size
    for i in range(n):
        t = int(rng.integers(0, T))
        L = int(rng.integers(k - 2, 320))
        pos = int(rng.integers(0, G - L))
        codes = targets[t][pos:pos + L].copy()
        if rng.random() < 0.5:
            codes = 3 - codes[::-1]
        seq = bytearray(b"ACGT"[c] for c in codes)
        if rng.random() < 0.3:
            seq[int(rng.integers(0, L))] = ord("N")
        if rng.random() < 0.3:
            j = int(rng.integers(0, L))
            if seq[j] != ord("N"):
                seq[j] = b"ACGT"[(b"ACGT".index(seq[j]) + 1) % 4]
        out.append(b">r%d\n" % i + bytes(seq) + b"\n")
    return b"".join(out)


@pytest.mark.slow
@pytest.mark.parametrize("k,htsize,layouts", [(32, HTSIZE_FULL, (0, 2)), (30, HTSIZE_LIGHT, (0, 1, 2, 3)),
                                              (29, HTSIZE_LIGHT, (0, 3))])
def test_uint64_keys(oracle, k, htsize, layouts, monkeypatch):
    """8-byte .ky elements: k=32 on the full HTSIZE is the reference's T64 instantiation
    (CuClarkDB<uint64_t>, src/main.cc:278-316); k=29/30 on the light HTSIZE have 8-byte keys too.
    (LOCAL at k=30 is a 17 GB table of 2^27 lines whatever it holds: forced as in the k=31 test.)"""
    monkeypatch.setenv("CUCLARK_ALLOW_SPARSE_TABLE", "1")
    targets, kmers, labels, (sz, ky, lb), kb = _small_case(k, htsize)
    assert kb == 8 and ky.dtype == np.uint64
    odb = oracle.db_from_arrays(htsize, k, sz, ky, lb)
    rng = np.random.default_rng(k)
    data = _reads_from(targets, rng, 1200, k) + low_complexity_reads(rng, 100, 150)
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, k, data, len(targets), 15)
    hi = (1 << 64) - 1 if k == 32 else (1 << (2 * k)) - 1
    probe = np.concatenate([kmers, dbtools.revcomp_codes(kmers, k),
                            rng.integers(0, hi, 50_000, dtype=np.uint64, endpoint=True)])
    expect, _ = odb.query(probe)
    pp, pc = pack_kmers_as_reads(probe, k)
    for layout in layouts:
        with CuClarkDB(k, len(targets), htsize=htsize, layout=layout, row_pairs=15) as g:
            g.load_arrays(sz, ky, lb)
            st = g.stats()
            assert st["n_entries"] == kmers.size
            if layout:
                assert st["layout"] == layout
            gf, gr = g.classify(ptr, cont, want_rows=True)
            assert g.stats()["lookups"] == lookups
            assert np.array_equal(gf, final), f"k={k} layout {layout}"
            assert np.array_equal(gr, rows), f"k={k} layout {layout}"
            gf, _ = g.classify(pp, pc)
        got = np.where(gf[:, 2] > 0, gf[:, 1].astype(np.int32) - 1, -1)
        assert np.array_equal(got, expect), f"k={k} layout {layout}"
    assert (final[:, 1] > 0).sum() > 1000


def test_k32_database_files_roundtrip(oracle, tmp_path):
    """k=32: .ky written with 8-byte keys, loaded from files (key width chosen as src/main.cc:278-316)."""
    k = 32
    targets, kmers, labels, (sz, ky, lb), kb = _small_case(k, HTSIZE_LIGHT, T=4, G=8000)
    base = str(tmp_path / "db")
    dbtools.write_db_files(base, sz, ky, lb)
    odb = oracle.db_load(base, HTSIZE_LIGHT, k)
    rng = np.random.default_rng(5)
    data = _reads_from(targets, rng, 300, k)
    ptr, cont, final, rows, _ = oracle_expect(oracle, odb, k, data, 4, 23)
    with CuClarkDB(k, 4, htsize=HTSIZE_LIGHT) as g:
        assert g.read(base) is True
        gf, gr = g.classify(ptr, cont, want_rows=True)
    assert np.array_equal(gf, final) and np.array_equal(gr, rows)


@pytest.mark.parametrize("layout", [0, 3])
def test_parts_of_65536_nt_and_more(oracle, light_small, layout):
    """An N-free record of 70,000 / 140,000 nt: the part header is a uint16, the packer (oracle and device alike)
    cuts such a run into parts of <= 65,535 nt overlapping by k-1 nt, so every k-mer is looked up exactly once
    (deviation Q8: the reference's header wraps and its kernel reads data as headers). Text path and packed path."""
    c = light_small
    k = c.k
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, k, sz, ky, lb)
    import sys
    mg = sys.modules["make_golden"]
    asc = np.frombuffer(b"ACGT", np.uint8)
    g1, g2, g5 = (asc[mg.target_codes(c.case, t)].tobytes() for t in (1, 2, 5))
    wrap = lambda s, w: b"\n".join(s[i:i + w] for i in range(0, len(s), w))
    recs = [b">short\n" + g1[:300] + b"\n",
            b">n70000\n" + wrap(g1[100:70_100], 70) + b"\n",
            b">exact65535\n" + g2[:65_535] + b"\n",
            b">exact65536\n" + g2[:65_536] + b"\n",
            b">exact65535pluskminus1\n" + g2[7:7 + 65_535 + k - 1] + b"\n",
            b">n140000_two_splits\n" + wrap(g5[:140_000], 61) + b"N" + g5[150_000:150_100] + b"\n",
            b">tail\n" + g5[:200] + b"\n"]
    data = b"".join(recs)
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, k, data, c.n_targets, c.maxhits)
    exp_lookups = sum(L - k + 1 for L in (300, 70_000, 65_535, 65_536, 65_535 + k - 1, 140_000, 100, 200))
    assert lookups == exp_lookups                            # every window exactly once
    hdrs = cont[ptr[1]], cont[ptr[1] + 1 + 8192]
    assert hdrs == (65_535, 70_000 - 65_535 + k - 1)
    with make_gpu(c, layout=layout) as g:
        gf, gr = g.classify(ptr, cont, want_rows=True)
        assert g.stats()["lookups"] == lookups
        assert np.array_equal(gf, final) and np.array_equal(gr, rows)
        arr, st = g.text_debug(data, 16, cont.size + 64, classify=True, want_rows=True)
        assert st["n_reads"] == len(recs) and st["lookups"] == lookups
        assert np.array_equal(arr["reads_ptr"], ptr) and np.array_equal(arr["containers"], cont)
        assert np.array_equal(arr["final5"], final) and np.array_equal(arr["rows"], rows)
        # a header that lies (the reference's own packer wraps it at 65,536): results are undefined as in the
        # reference, but the kernel must stay inside the read's containers
        bad = cont.copy()
        bad[ptr[1]] = 4464                                   # 70,000 & 0xFFFF
        g.classify(ptr, bad, want_rows=True)
        bad[ptr[6]] = 65_535                                 # header far larger than the read
        g.classify(ptr, bad, want_rows=True)
        gf, _ = g.classify(ptr, cont)
        assert np.array_equal(gf, final)


def test_record_larger_than_chunk(oracle, light_small, tmp_path):
    """A record longer than chunk_bytes is taken whole (the slot grows) instead of failing
    (src/CuCLARK_hh.hh:1377-1389 takes records of any size)."""
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    import sys
    mg = sys.modules["make_golden"]
    asc = np.frombuffer(b"ACGT", np.uint8)
    g0, g4 = (asc[mg.target_codes(c.case, t)].tobytes() for t in (0, 4))
    wrap = lambda s, w: b"\n".join(s[i:i + w] for i in range(0, len(s), w))
    for fastq in (False, True):
        recs = []
        for i, (src, L) in enumerate([(g0, 150), (g4, 30_000), (g0, 150), (g0, 9_000), (g4, 20_000), (g4, 120)]):
            seq = src[i * 11:i * 11 + L]
            if fastq:
                recs.append(b"@q%d\n" % i + seq + b"\n+\n" + b"@" * L + b"\n")      # quality lines starting with '@'
            else:
                recs.append(b">f%d\n" % i + wrap(seq, 80) + b"\n")
        data = b"".join(recs)
        _, ptr, cont, final, rows, lookups, expect = oracle_all(oracle, odb, c, data, tmp_path)
        with make_gpu(c) as g:
            arr, st = g.text_debug(data, 16, cont.size + 64, chunk_bytes=8192, n_slots=2, classify=True)
            assert st["n_reads"] == 6 and st["lookups"] == lookups
            assert np.array_equal(arr["reads_ptr"], ptr) and np.array_equal(arr["containers"], cont)
            assert np.array_equal(arr["final5"], final)
            csv, st2 = g.classify_text(data, names=c.names, chunk_bytes=8192, n_slots=2)
            assert st2["n_reads"] == 6
            assert csv == expect
