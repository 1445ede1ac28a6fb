"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit-exact.

Every test drives libcuclark_b200.so through cuclark_b200.api.CuClarkDB (the
mirror of the reference's CuClarkDB class) and compares per-read final results
(sum, best, hits, second, hits) and sparse rows with oracle/cuclark_oracle.c on
the same seeded inputs.
"""
import numpy as np
import pytest

from cuclark_b200 import synth
from cuclark_b200.api import CuClarkDB, HTSIZE_LIGHT
from oracle import dbtools

pytestmark = pytest.mark.gpu


def oracle_expect(oracle, case_or_db, k, reads_bytes, n_targets, row_pairs, n_batches=1, part=None):
    ix, buf = oracle.index(reads_bytes, n_batches)
    ptr, cont = oracle.pack(ix, buf, k)
    final, rows, lookups = oracle.classify(case_or_db, ptr, cont, n_targets, row_pairs, part=part, threads=4)
    oracle.free_index(ix)
    return ptr, cont, final, rows, lookups


def make_gpu(case, **kw):
    g = CuClarkDB(case.k, case.n_targets, htsize=case.htsize, **kw)
    sz, ky, lb = case.arrays
    g.load_arrays(sz, ky, lb)
    return g


def test_light_small_final_and_rows(oracle, light_small):
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits)
    with make_gpu(c) as g:
        st = g.stats()
        assert st["n_entries"] == c.kmers.size and st["layout"] == 1
        gf, gr = g.classify(ptr, cont, want_rows=True)
        assert np.array_equal(gf, final)
        assert np.array_equal(gr, rows)
        assert g.stats()["lookups"] == lookups
        gf2, _ = g.classify(ptr, cont, want_rows=False)
        assert np.array_equal(gf2, final)
    assert (final[:, 1] > 0).sum() > 1000          # the case really classifies reads


def test_config1_light_c1(oracle, light_c1):
    """BASELINE.json configs[0]: CuCLARK-l k=27, 20 x 1 Mbp targets, 100k x 100 bp reads."""
    c = light_c1
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits, 8)
    with make_gpu(c) as g:
        gf, gr = g.classify(ptr, cont, want_rows=True)
    assert np.array_equal(gf, final) and np.array_equal(gr, rows)
    assert lookups == 7380778


@pytest.mark.parametrize("layout", [1, 2])
@pytest.mark.parametrize("load", [0.0, 4.2])
def test_layouts_and_spills(oracle, light_small, layout, load):
    """Both bucket layouts, and a tight geometry that forces spilled entries."""
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    # random + DB k-mers as one long list of single-k-mer "reads"
    rng = np.random.default_rng(3)
    kmers = np.concatenate([c.kmers, dbtools.revcomp_codes(c.kmers, c.k),
                            rng.integers(0, 1 << (2 * c.k), 50_000, dtype=np.uint64)])
    expect, _ = odb.query(kmers)
    ptr, cont = pack_kmers_as_reads(kmers, c.k)
    # a dense table: few buckets -> many full buckets -> spills (narrow needs M > 4^k/2^32, so only wide gets tight)
    with CuClarkDB(c.k, c.n_targets, htsize=c.htsize, layout=layout, bucket_load=load if layout == 2 else 0.0) as g:
        g.load_arrays(sz, ky, lb)
        st = g.stats()
        assert st["layout"] == layout
        if layout == 2 and load:
            assert st["n_spilled"] > 100 and st["n_spill_buckets"] > 100
        gf, _ = g.classify(ptr, cont)
    got = np.where(gf[:, 2] > 0, gf[:, 1].astype(np.int32) - 1, -1)
    assert np.array_equal(got, expect)


def pack_kmers_as_reads(kmers, k):
    """One part per read holding exactly one k-mer (R form is what is stored in containers)."""
    n = kmers.size
    nc = (k + 7) // 8
    cont = np.zeros((n, 1 + nc), np.uint16)
    cont[:, 0] = k
    x = kmers.astype(np.uint64) << np.uint64(64 - 2 * k)       # left-align
    for j in range(nc):
        cont[:, 1 + j] = ((x >> np.uint64(48 - 16 * j)) & np.uint64(0xFFFF)).astype(np.uint16)
    ptr = (np.arange(n + 1, dtype=np.uint64) * (1 + nc)).astype(np.uint32)
    return ptr, cont.reshape(-1)


def test_load_from_files_and_missing_file(oracle, light_small, tmp_path):
    c = light_small
    sz, ky, lb = c.arrays
    base = str(tmp_path / "db_central")
    dbtools.write_db_files(base, sz, ky, lb)
    odb = oracle.db_load(base, c.htsize, c.k)
    ptr, cont, final, rows, _ = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits)
    with CuClarkDB(c.k, c.n_targets, htsize=c.htsize) as g:
        assert g.read(str(tmp_path / "nope")) is False        # CuClarkDB::read contract
        assert g.read(base) is True
        gf, gr = g.classify(ptr, cont, want_rows=True)
    assert np.array_equal(gf, final) and np.array_equal(gr, rows)


@pytest.mark.parametrize("sfactor", [2, 5])
def test_sampling_factor(oracle, light_small, sfactor):
    """-s: keep every s-th non-empty bucket (src/CuClarkDB.cu:511-524)."""
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb, sfactor=sfactor)
    ptr, cont, final, rows, _ = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits)
    with CuClarkDB(c.k, c.n_targets, htsize=c.htsize) as g:
        g.load_arrays(sz, ky, lb, mod_collision=sfactor)
        assert g.stats()["n_entries"] == odb.size
        gf, gr = g.classify(ptr, cont, want_rows=True)
    assert np.array_equal(gf, final) and np.array_equal(gr, rows)


def test_batches_api(oracle, light_small):
    """malloc / readyBatch / queryBatch / waitForBatch with the library's pinned buffers."""
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    nb = 3
    ix, buf = oracle.index(c.reads_bytes, nb)
    from oracle.binding import ReadIndex
    rix = ReadIndex(ix)
    bf = rix.batch_first
    packed = [oracle.pack(ix, buf, c.k, int(bf[b]), int(bf[b + 1] - bf[b])) for b in range(nb)]
    max_reads = int(np.diff(bf).max())
    max_cont = max(p[1].size for p in packed)
    with make_gpu(c) as g:
        views = g.malloc(nb, max_reads, max_cont, is_extended=True)
        for b, (ptr, cont) in enumerate(packed):
            vp, vc, _, _ = views[b]
            vp[:ptr.size] = ptr
            vc[:cont.size] = cont
            assert g.readyBatch(b, ptr.size - 1, cont.size)
            assert g.queryBatch(b, True)
        for b, (ptr, cont) in enumerate(packed):
            assert g.waitForBatch(b)
            final, rows, _ = oracle.classify(odb, ptr, cont, c.n_targets, c.maxhits)
            n = ptr.size - 1
            assert np.array_equal(views[b][2][:n], final)
            assert np.array_equal(views[b][3][:n], rows)
        g.freeBatchMemory()
    oracle.free_index(ix)


def test_edge_cases(oracle, light_small):
    """Empty input, reads shorter than k, all-N reads, many parts, lower case / U, Length == k."""
    c = light_small
    k = c.k
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    asc = np.frombuffer(b"ACGT", np.uint8)
    import sys
    g0 = asc[sys.modules["make_golden"].target_codes(c.case, 0)].tobytes()
    g3 = asc[sys.modules["make_golden"].target_codes(c.case, 3)].tobytes()
    recs = [
        b">short\nACGT\n",
        b">empty\n\n",
        b">allN\n" + b"N" * 80 + b"\n",
        b">exactk\n" + g0[:k] + b"\n",
        b">kminus1\n" + g0[:k - 1] + b"\n",
        b">multi line\n" + g0[0:60] + b"\n" + g0[60:130] + b"\n" + g0[130:216] + b"\n",
        b">parts\n" + g0[0:40] + b"N" + g0[41:60] + b"NN" + g3[108:216] + b"n" + g3[216:220] + b"\n",
        b">lower\n" + g0[0:216].lower() + b"\n",
        b">rna\n" + g0[0:216].replace(b"T", b"U") + b"\n",
        b">cr\n" + g0[0:108] + b"\r\n" + g0[108:216] + b"\r\n",
        b">two_targets\n" + g0[0:216] + g3[0:324] + b"\n",
        b">tail_no_newline\n" + g3[0:300],
    ]
    data = b"".join(recs)
    ptr, cont, final, rows, _ = oracle_expect(oracle, odb, k, data, c.n_targets, c.maxhits)
    assert final[10, 1] == 4 and final[10, 3] == 1          # target 3 first (3 hits), target 0 second (2 hits)
    with make_gpu(c) as g:
        gf, gr = g.classify(ptr, cont, want_rows=True)
        assert np.array_equal(gf, final) and np.array_equal(gr, rows)
        # zero reads
        e_f, e_r = g.classify(np.zeros(1, np.uint32), np.zeros(0, np.uint16), want_rows=True)
        assert e_f.shape == (0, 5) and e_r.shape[0] == 0


def test_long_reads_cross_chunks(oracle, light_small):
    """Parts longer than one 992-k-mer chunk, and lengths around the chunk edges."""
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    import sys
    asc = np.frombuffer(b"ACGT", np.uint8)
    g5 = asc[sys.modules["make_golden"].target_codes(c.case, 5)].tobytes()
    recs = []
    for i, L in enumerate([991 + 26, 992 + 26, 993 + 26, 1024, 1984 + 26, 1985 + 26, 5000, 30000, 65535]):
        recs.append(b">long%d\n" % i + g5[i * 7: i * 7 + L] + b"\n")
    data = b"".join(recs)
    ptr, cont, final, rows, _ = oracle_expect(oracle, odb, c.k, data, c.n_targets, c.maxhits)
    assert final[:, 2].min() >= 9
    with make_gpu(c) as g:
        gf, gr = g.classify(ptr, cont, want_rows=True)
    assert np.array_equal(gf, final) and np.array_equal(gr, rows)


def test_many_targets_dense_fallback(oracle):
    """> 64 distinct targets in one read: exact via the dense fallback; rows keep the first MAXHITS targets."""
    k, T, G = 27, 300, 2000
    targets = [synth.genome_codes(77, t, 0, G) for t in range(T)]
    kmers, labels = dbtools.build_entries(targets, k, 0)
    sz, ky, lb = dbtools.entries_to_arrays(kmers, labels, HTSIZE_LIGHT, 4)
    odb = oracle.db_from_arrays(HTSIZE_LIGHT, k, sz, ky, lb)
    asc = np.frombuffer(b"ACGT", np.uint8)
    seg = lambda t, n: asc[targets[t][100:100 + n]].tobytes()
    reads = [
        b">r70\n" + b"".join(seg(t, 40) for t in range(0, 70)) + b"\n",          # 70 targets
        b">r200\n" + b"".join(seg(t, 30 + (t % 7)) for t in range(299, 99, -1)) + b"\n",
        b">r64\n" + b"".join(seg(t, 40) for t in range(100, 164)) + b"\n",        # exactly 64: still the warp table
        b">r20\n" + b"".join(seg(t, 45) for t in range(10, 30)) + b"\n",          # > MAXHITS(15) but <= 64
        b">r1\n" + seg(5, 200) + b"\n",
    ]
    data = b"".join(reads)
    for rp in (15, 23):
        ptr, cont, final, rows, _ = oracle_expect(oracle, odb, k, data, T, rp)
        assert rows[0, 0] == 70 and rows[1, 0] == 200 and rows[2, 0] == 64 and rows[3, 0] == 20
        with CuClarkDB(k, T, htsize=HTSIZE_LIGHT, row_pairs=rp) as g:
            g.load_arrays(sz, ky, lb)
            gf, gr = g.classify(ptr, cont, want_rows=True)
            st = g.stats()
            assert np.array_equal(gf, final)
            assert np.array_equal(gr, rows)
            assert st["dense_reads"] == 2
            assert st["truncated_rows"] == (4 if rp == 15 else 3)
            gf2, _ = g.classify(ptr, cont, want_rows=False)
            assert np.array_equal(gf2, final)


def test_tie_breaking(oracle):
    """Equal counts: lowest target index is first, the next equal one second (resultKernel's strict '>')."""
    k, T, G = 27, 12, 3000
    targets = [synth.genome_codes(91, t, 0, G) for t in range(T)]
    kmers, labels = dbtools.build_entries(targets, k, 0)
    sz, ky, lb = dbtools.entries_to_arrays(kmers, labels, HTSIZE_LIGHT, 4)
    odb = oracle.db_from_arrays(HTSIZE_LIGHT, k, sz, ky, lb)
    asc = np.frombuffer(b"ACGT", np.uint8)
    seg = lambda t, n: asc[targets[t][200:200 + n]].tobytes()
    reads = [
        # segments are joined by N so that no k-mer spans a junction: counts are exact
        b">tie3\n" + seg(9, 40) + b"N" + seg(2, 40) + b"N" + seg(7, 40) + b"\n",        # 14,14,14 -> 2 then 7
        b">tie_second\n" + seg(4, 60) + b"N" + seg(11, 40) + b"N" + seg(1, 40) + b"\n",  # 34; 14,14 -> 4 then 1
        b">desc\n" + seg(10, 50) + b"N" + seg(6, 45) + b"N" + seg(3, 40) + b"\n",
    ]
    data = b"".join(reads)
    ptr, cont, final, rows, _ = oracle_expect(oracle, odb, k, data, T, 15)
    assert list(final[0]) == [42, 3, 14, 8, 14]
    assert list(final[1]) == [62, 5, 34, 2, 14]
    with CuClarkDB(k, T, htsize=HTSIZE_LIGHT, row_pairs=15) as g:
        g.load_arrays(sz, ky, lb)
        gf, gr = g.classify(ptr, cont, want_rows=True)
    assert np.array_equal(gf, final) and np.array_equal(gr, rows)


@pytest.mark.slow
def test_full_variant_k31_wide(oracle, full_small):
    """cuCLARK (full) k=31: 1.6 G-bucket file layout, small table -> wide device layout."""
    c = full_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ptr, cont, final, rows, _ = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits)
    with make_gpu(c) as g:
        assert g.stats()["layout"] == 2
        gf, gr = g.classify(ptr, cont, want_rows=True)
    assert np.array_equal(gf, final) and np.array_equal(gr, rows)
    assert (final[:, 2] > 100).sum() > 10000


def test_table_partitioned_merge(oracle, light_small):
    """Table-partitioned mode: every shard classifies every read, rows are merged (mergeKernel+resultKernel)."""
    import torch
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ptr, cont, final, rows, _ = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits)
    n = ptr.size - 1
    G = 4
    parts = torch.zeros((G, n, 2 * c.maxhits + 2), dtype=torch.int16, device="cuda")
    d_ptr = torch.from_numpy(ptr.astype(np.int32)).cuda()
    d_cont = torch.from_numpy(cont.astype(np.int16)).cuda()
    torch.cuda.synchronize()          # the library works on its own stream: finish torch's fills first
    entries = 0
    shards = []
    for s in range(G):
        g = CuClarkDB(c.k, c.n_targets, htsize=c.htsize, shard=(s, G))
        g.load_arrays(sz, ky, lb)
        entries += g.stats()["n_entries"]
        g.classify_device(d_ptr.data_ptr(), d_cont.data_ptr(), n, 0, parts[s].data_ptr())
        g.stats(sync=True)
        shards.append(g)
    assert entries == c.kmers.size
    torch.cuda.synchronize()
    out_rows = torch.zeros((n, 2 * c.maxhits + 2), dtype=torch.int16, device="cuda")
    out_final = torch.zeros((n, 5), dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    shards[0].merge_rows_device(parts.data_ptr(), G, n, out_rows.data_ptr(), out_final.data_ptr())
    shards[0].stats(sync=True)
    torch.cuda.synchronize()
    assert np.array_equal(out_final.cpu().numpy().view(np.uint16), final)
    assert np.array_equal(out_rows.cpu().numpy().view(np.uint16), rows)
    # per-shard rows equal the oracle restricted to nothing (sum of shards == whole): check one pairwise merge
    a = parts[0].cpu().numpy().view(np.uint16)
    b = parts[1].cpu().numpy().view(np.uint16)
    ab = oracle.merge_rows(a, b, c.maxhits)
    two = torch.stack([parts[0], parts[1]]).contiguous()
    shards[0].merge_rows_device(two.data_ptr(), 2, n, out_rows.data_ptr(), 0)
    shards[0].stats(sync=True)
    torch.cuda.synchronize()
    assert np.array_equal(out_rows.cpu().numpy().view(np.uint16), ab)
    for g in shards:
        g.close()


def test_synthetic_device_generators_match_numpy(oracle):
    """Device-built synthetic DB + device-generated packed reads == numpy twin + oracle."""
    import torch
    k, T, G, seed, rseed = 27, 16, 50_000, 5, 9
    n, L = 4000, 150
    targets = [synth.genome_codes(seed, t, 0, G) for t in range(T)]
    for gap in (0, 4):
        kmers, labels = dbtools.build_entries(targets, k, gap)
        sz, ky, lb = dbtools.entries_to_arrays(kmers, labels, HTSIZE_LIGHT, 4)
        odb = oracle.db_from_arrays(HTSIZE_LIGHT, k, sz, ky, lb)
        codes, *_ = synth.read_codes(rseed, n, L, T, G, seed, pct_random=10, sub_per_10k=100)
        data = synth.reads_fasta(codes)
        ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, k, data, T, 23)
        with CuClarkDB(k, T, htsize=HTSIZE_LIGHT) as g:
            g.build_synthetic(seed, T, G, light_gap=gap)
            assert g.stats()["n_entries"] == kmers.size
            per = 1 + (L + 7) // 8
            d_ptr = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
            d_cont = torch.zeros(n * per, dtype=torch.int16, device="cuda")
            torch.cuda.synchronize()      # the library works on its own stream: finish torch's fills first
            g.synth_reads_device(rseed, seed, T, G, 0, n, L, 10, 100, d_ptr.data_ptr(), d_cont.data_ptr())
            g.stats(sync=True)
            torch.cuda.synchronize()
            assert np.array_equal(d_ptr.cpu().numpy().view(np.uint32), ptr)
            assert np.array_equal(d_cont.cpu().numpy().view(np.uint16), cont)
            gf, gr = g.classify(ptr, cont, want_rows=True)
            assert np.array_equal(gf, final) and np.array_equal(gr, rows)
            assert g.stats()["lookups"] == lookups


def test_synthetic_db_removes_common_kmers(oracle):
    """Synthetic builder dedupe: k-mers present in two targets are removed, repeats within one kept once."""
    # tiny genomes with k small enough that collisions between random targets are frequent
    k, T, G, seed = 9, 6, 3000, 3
    targets = [synth.genome_codes(seed, t, 0, G) for t in range(T)]
    kmers, labels = dbtools.build_entries(targets, k, 0)
    assert 0 < kmers.size < T * (G - k + 1) * 0.95          # collisions exist
    with CuClarkDB(k, T, htsize=HTSIZE_LIGHT) as g:
        g.build_synthetic(seed, T, G, light_gap=0)
        assert g.stats()["n_entries"] == kmers.size
        allk = np.arange(1 << (2 * k), dtype=np.uint64)
        ptr, cont = pack_kmers_as_reads(allk, k)
        gf, _ = g.classify(ptr, cont)
    got = np.where(gf[:, 2] > 0, gf[:, 1].astype(np.int32) - 1, -1)
    expect = np.full(allk.size, -1, np.int32)
    can = dbtools.canonical(allk, k)
    pos = np.searchsorted(kmers, can)
    pos[pos >= kmers.size] = 0
    hit = kmers[pos] == can
    expect[hit] = labels[pos[hit]]
    assert np.array_equal(got, expect)
