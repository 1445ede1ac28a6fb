"""GPU parity of the LOCAL table layout (minimizer-addressed 128-byte lines, csrc/common.cuh).

The layout changes only WHERE a k-mer lives and how the kernel finds it (rolling window minimum
across lanes); every result must stay bit-identical to the oracle, including k-mers that spill to
the overflow table, k-mers homed in other shards, ties between equal m-mers inside one k-mer
(low-complexity reads) and the device-built synthetic database.
"""
import numpy as np
import pytest

from cuclark_b200 import synth
from cuclark_b200.api import CuClarkDB, HTSIZE_LIGHT
from oracle import dbtools
from oracle.binding import key_bytes_for

from test_gpu_parity import make_gpu, oracle_expect, pack_kmers_as_reads

pytestmark = pytest.mark.gpu
LOCAL = 3


def test_light_small_final_and_rows_local(oracle, light_small):
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits)
    with make_gpu(c, layout=LOCAL) as g:
        st = g.stats()
        assert st["layout"] == LOCAL and st["n_entries"] == c.kmers.size
        gf, gr = g.classify(ptr, cont, want_rows=True)
        assert np.array_equal(gf, final)
        assert np.array_equal(gr, rows)
        assert g.stats()["lookups"] == lookups
        gf2, _ = g.classify(ptr, cont, want_rows=False)
        assert np.array_equal(gf2, final)


def small_k_case(k, n_targets, genome_len, seed):
    """All overlapping k-mers of seeded genomes (full-variant scanner): consecutive k-mers share
    minimizers, so the lines fill in clumps as they do at bacterial scale."""
    targets = [synth.genome_codes(seed, t, 0, genome_len) for t in range(n_targets)]
    kmers, labels = dbtools.build_entries(targets, k, 0)
    kb = key_bytes_for(k, HTSIZE_LIGHT)
    return targets, kmers, labels, dbtools.entries_to_arrays(kmers, labels, HTSIZE_LIGHT, kb), kb


@pytest.mark.parametrize("rescue", [True, False])
@pytest.mark.parametrize("k,load", [(21, 0.0), (21, 3.8), (19, 3.7), (24, 3.5)])
def test_local_spills_every_kmer_and_random(oracle, k, load, rescue, monkeypatch):
    """Each DB k-mer (both strands) and random k-mers as single-k-mer reads; tight loads force spills. With the
    builder's rescue pass (entries whose two sectors are full displace a neighbour, csrc/table.cu k_local_rescue)
    nearly all of them are housed in the lines after all; without it (CUCLARK_NO_RESCUE) they exercise the overflow
    table. Every k-mer must be found either way."""
    if not rescue:
        monkeypatch.setenv("CUCLARK_NO_RESCUE", "1")
    T, G = 8, 40_000
    _, kmers, labels, (sz, ky, lb), kb = small_k_case(k, T, G, 11)
    odb = oracle.db_from_arrays(HTSIZE_LIGHT, k, sz, ky, lb)
    rng = np.random.default_rng(4)
    probe = np.concatenate([kmers, dbtools.revcomp_codes(kmers, k),
                            rng.integers(0, 1 << (2 * k), 60_000, dtype=np.uint64)])
    expect, _ = odb.query(probe)
    ptr, cont = pack_kmers_as_reads(probe, k)
    with CuClarkDB(k, T, htsize=HTSIZE_LIGHT, layout=LOCAL, bucket_load=load) as g:
        g.load_arrays(sz, ky, lb)
        st = g.stats()
        assert st["layout"] == LOCAL and st["n_entries"] == kmers.size
        if load and not rescue:   # with two candidate lines only a nearly full table spills
            assert st["n_spilled"] > 100 and st["n_spill_buckets"] > 100
        if load and rescue:       # 3.5-3.8 entries per 4-slot sector (88-95 % full): the displacement still houses most of them
            assert st["n_spilled"] < 0.10 * kmers.size
        gf, _ = g.classify(ptr, cont)
    got = np.where(gf[:, 2] > 0, gf[:, 1].astype(np.int32) - 1, -1)
    assert np.array_equal(got, expect)


def low_complexity_reads(rng, n, length):
    """Periodic sequences (period 1..6) with a few point changes: equal m-mers inside one k-mer."""
    out = []
    for i in range(n):
        period = int(rng.integers(1, 7))
        unit = rng.integers(0, 4, period)
        codes = np.resize(unit, length).copy()
        for _ in range(int(rng.integers(0, 4))):
            codes[int(rng.integers(0, length))] = rng.integers(0, 4)
        out.append(b">lc%d\n" % i + bytes(b"ACGT"[c] for c in codes) + b"\n")
    return b"".join(out)


@pytest.mark.parametrize("k", [21, 27])
def test_local_reads_with_ties_and_long_parts(oracle, k):
    """Whole reads: sampled from the targets (both strands, substitutions, N), low-complexity reads whose
    k-mers hold the same m-mer several times (also present in the DB), and parts longer than one chunk."""
    T, G = 6, 30_000
    rng = np.random.default_rng(21)
    targets = [synth.genome_codes(13, t, 0, G) for t in range(T)]
    # two targets carry low-complexity stretches so that tie k-mers are IN the database
    for t in (1, 4):
        for _ in range(30):
            period = int(rng.integers(1, 6))
            pos = int(rng.integers(0, G - 200))
            targets[t][pos:pos + 120] = np.resize(rng.integers(0, 4, period), 120)
    kmers, labels = dbtools.build_entries(targets, k, 0)
    kb = key_bytes_for(k, HTSIZE_LIGHT)
    sz, ky, lb = dbtools.entries_to_arrays(kmers, labels, HTSIZE_LIGHT, kb)
    odb = oracle.db_from_arrays(HTSIZE_LIGHT, k, sz, ky, lb)
    reads = []
    for i in range(1500):
        t = int(rng.integers(0, T))
        L = int(rng.integers(k - 2, 400))
        pos = int(rng.integers(0, G - L))
        codes = targets[t][pos:pos + L].copy()
        if rng.random() < 0.5:
            codes = 3 - codes[::-1]
        seq = bytearray(b"ACGT"[c] for c in codes)
        if rng.random() < 0.3:
            seq[int(rng.integers(0, L))] = ord("N")
        if rng.random() < 0.3:
            j = int(rng.integers(0, L))
            seq[j] = b"ACGT"[(b"ACGT".index(seq[j]) + 1) % 4] if seq[j] != ord("N") else seq[j]
        reads.append(b">r%d\n" % i + bytes(seq) + b"\n")
    # reads straight out of the low-complexity stretches of the targets, and synthetic periodic ones
    data = b"".join(reads) + low_complexity_reads(rng, 300, 150)
    long_codes = np.concatenate([targets[2][:3000], 3 - targets[3][::-1][:2500]])
    data += b">long\n" + bytes(b"ACGT"[c] for c in long_codes) + b"\n"
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, k, data, T, 23)
    for load in (0.0, 3.7):
        with CuClarkDB(k, T, htsize=HTSIZE_LIGHT, layout=LOCAL, bucket_load=load) as g:
            g.load_arrays(sz, ky, lb)
            assert g.stats()["layout"] == LOCAL
            gf, gr = g.classify(ptr, cont, want_rows=True)
            assert g.stats()["lookups"] == lookups
        assert np.array_equal(gf, final), f"load {load}"
        assert np.array_equal(gr, rows), f"load {load}"
    assert (final[:, 1] > 0).sum() > 1200


def test_local_table_partitioned(oracle, light_small):
    """Shards hold whole lines; rows of the shards merge to the single-table result."""
    import torch
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ptr, cont, final, rows, _ = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits)
    n = ptr.size - 1
    G = 3
    parts = torch.zeros((G, n, 2 * c.maxhits + 2), dtype=torch.int16, device="cuda")
    d_ptr = torch.from_numpy(ptr.astype(np.int32)).cuda()
    d_cont = torch.from_numpy(cont.astype(np.int16)).cuda()
    torch.cuda.synchronize()          # the library works on its own stream: finish torch's fills first
    entries, shards = 0, []
    for s in range(G):
        g = CuClarkDB(c.k, c.n_targets, htsize=c.htsize, shard=(s, G), layout=LOCAL)
        g.load_arrays(sz, ky, lb)
        st = g.stats()
        assert st["layout"] == LOCAL
        entries += st["n_entries"]
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
    for g in shards:
        g.close()


@pytest.mark.parametrize("k,gap", [(27, 0), (27, 4), (21, 0)])
def test_local_synthetic_builder_and_cache(oracle, tmp_path, k, gap):
    """Device-built synthetic DB (incl. RemoveCommon at k=21, where random genomes share k-mers) in the
    LOCAL layout, and its table cache round trip."""
    T, G, seed, rseed = 12, 50_000, 5, 9
    n, L = 3000, 150
    targets = [synth.genome_codes(seed, t, 0, G) for t in range(T)]
    kmers, labels = dbtools.build_entries(targets, k, gap)
    kb = key_bytes_for(k, HTSIZE_LIGHT)
    sz, ky, lb = dbtools.entries_to_arrays(kmers, labels, HTSIZE_LIGHT, kb)
    odb = oracle.db_from_arrays(HTSIZE_LIGHT, k, sz, ky, lb)
    codes, *_ = synth.read_codes(rseed, n, L, T, G, seed, pct_random=10, sub_per_10k=100)
    data = synth.reads_fasta(codes)
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, k, data, T, 23)
    path = str(tmp_path / "t.b200")
    with CuClarkDB(k, T, htsize=HTSIZE_LIGHT, layout=LOCAL, bucket_load=3.6 if k == 21 else 0.0) as g:
        g.build_synthetic(seed, T, G, light_gap=gap)
        st = g.stats()
        assert st["layout"] == LOCAL and st["n_entries"] == kmers.size
        gf, gr = g.classify(ptr, cont, want_rows=True)
        assert np.array_equal(gf, final) and np.array_equal(gr, rows)
        g.save_table(path)
    with CuClarkDB(k, T, htsize=HTSIZE_LIGHT, layout=LOCAL) as g2:
        g2.load_table(path)
        st2 = g2.stats()
        for key in ("n_entries", "n_buckets", "n_local_buckets", "table_bytes", "n_spilled", "n_spill_buckets", "layout"):
            assert st2[key] == st[key], key
        gf2, gr2 = g2.classify(ptr, cont, want_rows=True)
    assert np.array_equal(gf2, final) and np.array_equal(gr2, rows)


def test_local_many_targets_dense_fallback(oracle):
    """> 64 distinct targets in one read: the exact dense fallback probes the LOCAL table through table_lookup()
    (brute-force minimizer, both candidate sectors, overflow table), at a load that spills entries."""
    k, T, G = 27, 300, 2000
    targets = [synth.genome_codes(77, t, 0, G) for t in range(T)]
    kmers, labels = dbtools.build_entries(targets, k, 0)
    sz, ky, lb = dbtools.entries_to_arrays(kmers, labels, HTSIZE_LIGHT, 4)
    odb = oracle.db_from_arrays(HTSIZE_LIGHT, k, sz, ky, lb)
    asc = np.frombuffer(b"ACGT", np.uint8)
    seg = lambda t, n: asc[targets[t][100:100 + n]].tobytes()
    reads = [
        b">r70\n" + b"".join(seg(t, 40) for t in range(0, 70)) + b"\n",
        b">r200\n" + b"".join(seg(t, 30 + (t % 7)) for t in range(299, 99, -1)) + b"\n",
        b">r64\n" + b"".join(seg(t, 40) for t in range(100, 164)) + b"\n",
        b">r20\n" + b"".join(seg(t, 45) for t in range(10, 30)) + b"\n",
        b">r1\n" + seg(5, 200) + b"\n",
    ]
    data = b"".join(reads)
    ptr, cont, final, rows, _ = oracle_expect(oracle, odb, k, data, T, 23)
    for load in (0.0, 3.7):
        with CuClarkDB(k, T, htsize=HTSIZE_LIGHT, row_pairs=23, layout=LOCAL, bucket_load=load) as g:
            g.load_arrays(sz, ky, lb)
            gf, gr = g.classify(ptr, cont, want_rows=True)
            st = g.stats()
        assert st["layout"] == LOCAL and st["dense_reads"] == 2
        assert np.array_equal(gf, final) and np.array_equal(gr, rows)


@pytest.mark.parametrize("sfactor", [2, 5])
def test_local_sampling_factor_and_batches(oracle, light_small, sfactor):
    """-s sampling at load time and the CuClarkDB batch API (malloc / readyBatch / queryBatch / waitForBatch) on a LOCAL table."""
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb, sfactor=sfactor)
    ptr, cont, final, rows, _ = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits)
    n = ptr.size - 1
    with CuClarkDB(c.k, c.n_targets, htsize=c.htsize, layout=LOCAL) as g:
        g.load_arrays(sz, ky, lb, mod_collision=sfactor)
        assert g.stats()["layout"] == LOCAL and g.stats()["n_entries"] == odb.size
        gf, gr = g.classify(ptr, cont, want_rows=True)
        assert np.array_equal(gf, final) and np.array_equal(gr, rows)
        nb = 3
        bounds = [n * b // nb for b in range(nb + 1)]
        max_reads = max(bounds[b + 1] - bounds[b] for b in range(nb))
        max_cont = max(int(ptr[bounds[b + 1]] - ptr[bounds[b]]) for b in range(nb))
        views = g.malloc(nb, max_reads, max_cont, is_extended=True)
        for b in range(nb):
            lo, hi = bounds[b], bounds[b + 1]
            vp, vc, _, _ = views[b]
            vp[:hi - lo + 1] = ptr[lo:hi + 1] - ptr[lo]
            vc[:int(ptr[hi] - ptr[lo])] = cont[ptr[lo]:ptr[hi]]
            g.readyBatch(b, hi - lo, int(ptr[hi] - ptr[lo]))
            g.queryBatch(b, True)
        for b in range(nb):
            g.waitForBatch(b)
            lo, hi = bounds[b], bounds[b + 1]
            assert np.array_equal(views[b][2][:hi - lo], final[lo:hi])
            assert np.array_equal(views[b][3][:hi - lo], rows[lo:hi])
        g.freeBatchMemory()


def test_local_shards_ties_through_the_dense_path(oracle):
    """Sharded LOCAL table + tie k-mers (periodic sequence: the smallest minimizer hash at two offsets, possibly homed in
    two shards) + reads that hit more than 64 targets on SOME shards only (dense fallback there, fast path on the
    others): exactly one shard must answer for every k-mer whichever path the read takes on each shard. Checked on the
    per-shard hit totals (final[0]), which are exact on both paths: their sum over the shards is the oracle's total.
    (The merged rows of such reads are cut at 63 pairs per shard and flagged in truncated_rows — see k_merge_rows.)"""
    import torch
    k, T, G = 21, 160, 1500
    rng = np.random.default_rng(77)
    targets = [synth.genome_codes(31, t, 0, G) for t in range(T)]
    for t in range(T):                                           # every target carries a periodic stretch of its own
        period = 2 + (t % 5)
        unit = rng.integers(0, 4, period)
        while len(set(unit.tolist())) < 2:
            unit = rng.integers(0, 4, period)
        targets[t][300:300 + 90] = np.resize(np.concatenate([unit, rng.integers(0, 4, 1 + t % 3)]), 90)
    kmers, labels = dbtools.build_entries(targets, k, 0)
    kb = key_bytes_for(k, HTSIZE_LIGHT)
    sz, ky, lb = dbtools.entries_to_arrays(kmers, labels, HTSIZE_LIGHT, kb)
    odb = oracle.db_from_arrays(HTSIZE_LIGHT, k, sz, ky, lb)
    asc = np.frombuffer(b"ACGT", np.uint8)

    def seg(t, a, n, rc):
        c = targets[t][a:a + n]
        return asc[(3 - c[::-1]) if rc else c].tobytes()
    reads = []
    for r in range(120):                                         # 60..110 targets per read: around and above the 64-slot limit
        ts = rng.permutation(T)[:int(rng.integers(60, 111))]
        reads.append(b">r%d\n" % r + b"".join(seg(int(t), 290, 110, bool(r & 1)) for t in ts) + b"\n")
    for r in range(40):                                          # few targets: fast path everywhere
        ts = rng.permutation(T)[:3]
        reads.append(b">s%d\n" % r + b"N".join(seg(int(t), 280, 130, bool(r & 1)) for t in ts) + b"\n")
    data = b"".join(reads)
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, k, data, T, 23)
    n = ptr.size - 1
    d_ptr = torch.from_numpy(ptr.astype(np.int32)).cuda()
    d_cont = torch.from_numpy(cont.astype(np.int16)).cuda()
    mixed = any_dense = 0
    for n_shards in (1, 3, 5):
        total = np.zeros(n, np.int64)
        dense_per_shard = []
        for s in range(n_shards):
            with CuClarkDB(k, T, htsize=HTSIZE_LIGHT, shard=(s, n_shards), layout=LOCAL) as g:
                g.load_arrays(sz, ky, lb)
                assert g.stats()["layout"] == LOCAL
                d_final = torch.zeros((n, 5), dtype=torch.int16, device="cuda")
                torch.cuda.synchronize()
                g.classify_device(d_ptr.data_ptr(), d_cont.data_ptr(), n, d_final.data_ptr(), 0)
                dense_per_shard.append(g.stats(sync=True)["dense_reads"])
                torch.cuda.synchronize()
                total += d_final.cpu().numpy().view(np.uint16)[:, 0].astype(np.int64)
        bad = np.nonzero((total & 0xFFFF) != final[:, 0])[0]
        assert bad.size == 0, (n_shards, bad[:5], total[bad[:5]], final[bad[:5], 0])
        if n_shards > 1:
            any_dense += sum(dense_per_shard)
            mixed += len(set(dense_per_shard)) > 1               # some shards took more reads to the dense path than others
    assert any_dense > 0 and mixed >= 1
    assert (final[:, 0] > 500).sum() >= 100
