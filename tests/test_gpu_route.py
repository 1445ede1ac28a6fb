"""GPU parity of table-partitioned classification by k-mer routing (csrc/route.cu) against the oracle.

Rank g of N holds shard g of the table and ITS share of the reads; canonical k-mers are bucketed by owner
shard (scatter), every shard probes the k-mers addressed to it out of the peers' memory and stores the labels
back (probe), the owner of the reads counts them (gather). The N ranks may share one device (plain pointers,
what these tests do on the one-GPU box) or sit on N devices (peer access over NVLink: the *_multi_gpu tests,
which skip below 2 GPUs) or in N processes (CUDA IPC: bench.py under torchrun).
The result must equal the single-table result — final rows and sparse rows — bit for bit.
"""
import numpy as np
import pytest

from cuclark_b200 import api, synth
from cuclark_b200.api import CuClarkDB, HTSIZE_LIGHT
from oracle import dbtools

from test_gpu_parity import oracle_expect

pytestmark = pytest.mark.gpu


def n_gpus() -> int:
    import torch
    return torch.cuda.device_count()


def split_reads(ptr, cont, n_ranks):
    """Contiguous read ranges per rank, each with its own containers (offsets rebased)."""
    n = ptr.size - 1
    out = []
    for r in range(n_ranks):
        lo, hi = r * n // n_ranks, (r + 1) * n // n_ranks
        p = (ptr[lo:hi + 1] - ptr[lo]).astype(np.uint32)
        out.append((lo, hi, p, cont[ptr[lo]:ptr[hi]].copy()))
    return out


def run_routed(case_k, n_targets, htsize, arrays, ptr, cont, n_ranks, row_pairs, devices=None, layout=0, want_rows=True):
    """N shard handles in this process -> (final, rows, per-rank route stats)."""
    import torch
    sz, ky, lb = arrays
    devices = devices or [0] * n_ranks
    shards, bufs = [], []
    n = ptr.size - 1
    pitch = 2 * row_pairs + 2
    parts = split_reads(ptr, cont, n_ranks)
    max_cont = max(max(c.size for _, _, _, c in parts), 1)
    entries = 0
    for r in range(n_ranks):
        g = CuClarkDB(case_k, n_targets, htsize=htsize, shard=(r, n_ranks), device=devices[r], row_pairs=row_pairs,
                      layout=layout)
        g.load_arrays(sz, ky, lb)
        entries += g.stats()["n_entries"]
        g.route_alloc(n_ranks, max_cont)
        shards.append(g)
    api.route_connect(shards)
    for r, (lo, hi, p, c) in enumerate(parts):
        dev = torch.device("cuda", devices[r])
        d_ptr = torch.from_numpy(p.astype(np.int32)).to(dev)
        d_cont = torch.from_numpy(c.astype(np.int16)).to(dev) if c.size else torch.zeros(1, dtype=torch.int16, device=dev)
        d_final = torch.zeros(((hi - lo) + 1, 5), dtype=torch.int16, device=dev)
        d_rows = torch.zeros(((hi - lo) + 1, pitch), dtype=torch.int16, device=dev)
        bufs.append((d_ptr, d_cont, d_final, d_rows))
    for d in set(devices):
        torch.cuda.synchronize(d)
    api.classify_routed_device(shards, [b[0].data_ptr() for b in bufs], [b[1].data_ptr() for b in bufs],
                               [hi - lo for lo, hi, _, _ in parts], [c.size for _, _, _, c in parts],
                               [b[2].data_ptr() for b in bufs], [b[3].data_ptr() for b in bufs] if want_rows else None)
    final = np.concatenate([b[2][:hi - lo].cpu().numpy().view(np.uint16) for b, (lo, hi, _, _) in zip(bufs, parts)])
    rows = np.concatenate([b[3][:hi - lo].cpu().numpy().view(np.uint16) for b, (lo, hi, _, _) in zip(bufs, parts)])
    stats = [g.route_stats() for g in shards]
    gstats = [g.stats() for g in shards]
    for g in shards:
        g.close()
    assert final.shape[0] == n
    return final, rows, stats, gstats, entries


@pytest.mark.parametrize("n_ranks,layout", [(1, 0), (2, 0), (3, 2), (4, 1), (8, 0), (1, 3), (2, 3), (5, 3)])
def test_routed_equals_oracle_one_device(oracle, light_small, n_ranks, layout):
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits)
    gf, gr, st, gst, entries = run_routed(c.k, c.n_targets, c.htsize, c.arrays, ptr, cont, n_ranks, c.maxhits, layout=layout)
    assert entries >= c.kmers.size if layout == 3 else entries == c.kmers.size
    assert np.array_equal(gf, final)
    assert np.array_equal(gr, rows)
    assert sum(s["lookups"] for s in st) == lookups            # every k-mer scattered once ...
    assert sum(s["probed"] for s in st) == lookups             # ... and probed once, by exactly one shard
    assert all(s["err"] == 0 for s in st)
    if n_ranks > 1:
        assert all(s["blocks_remote"] > 0 for s in st)


def test_routed_edge_cases_and_dense_fallback(oracle):
    """Empty ranks, reads shorter than k, many parts, > 64 targets in one read (dense fallback through the labels),
    > MAXHITS targets (truncated rows), long parts crossing 992-k-mer chunks."""
    k, T, G = 27, 300, 2000
    targets = [synth.genome_codes(77, t, 0, G) for t in range(T)]
    kmers, labels = dbtools.build_entries(targets, k, 0)
    arrays = dbtools.entries_to_arrays(kmers, labels, HTSIZE_LIGHT, 4)
    odb = oracle.db_from_arrays(HTSIZE_LIGHT, k, *arrays)
    asc = np.frombuffer(b"ACGT", np.uint8)
    seg = lambda t, n: asc[targets[t][100:100 + n]].tobytes()
    reads = [
        b">r70\n" + b"".join(seg(t, 40) for t in range(0, 70)) + b"\n",
        b">short\nACGT\n",
        b">r200\n" + b"".join(seg(t, 30 + (t % 7)) for t in range(299, 99, -1)) + b"\n",
        b">r64\n" + b"".join(seg(t, 40) for t in range(100, 164)) + b"\n",
        b">empty\n\n",
        b">r20\n" + b"".join(seg(t, 45) for t in range(10, 30)) + b"\n",
        b">parts\n" + seg(3, 60) + b"N" + seg(4, 20) + b"NN" + seg(5, 90) + b"\n",
        b">r1\n" + seg(5, 200) + b"\n",
        b">long\n" + asc[np.concatenate([targets[t][:1900] for t in range(7, 12)])].tobytes() + b"\n",
    ]
    data = b"".join(reads)
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, k, data, T, 15)
    for n_ranks, layout in ((2, 0), (16, 0), (3, 3)):          # 16 ranks over 9 reads: some ranks have no read; LOCAL shards
        gf, gr, st, gst, _ = run_routed(k, T, HTSIZE_LIGHT, arrays, ptr, cont, n_ranks, 15, layout=layout)
        assert np.array_equal(gf, final), n_ranks
        assert np.array_equal(gr, rows), n_ranks
        assert sum(s["lookups"] for s in st) == lookups == sum(s["probed"] for s in st)
        assert sum(g["dense_reads"] for g in gst) == 2         # r70 and r200 hit more than 64 targets
    gf, _, _, _, _ = run_routed(k, T, HTSIZE_LIGHT, arrays, ptr, cont, 3, 15, want_rows=False)
    assert np.array_equal(gf, final)


def test_routed_full_variant_k31(oracle, full_small):
    """cuCLARK (full) k=31 sharded 4 ways: NARROW needs M > 2^30, a small sharded table is WIDE."""
    c = full_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits)
    gf, gr, st, _, _ = run_routed(c.k, c.n_targets, c.htsize, c.arrays, ptr, cont, 4, c.maxhits)
    assert np.array_equal(gf, final) and np.array_equal(gr, rows)
    assert sum(s["probed"] for s in st) == lookups


def test_routed_refuses_what_it_cannot_do(light_small):
    c = light_small
    sz, ky, lb = c.arrays
    with CuClarkDB(c.k, c.n_targets, htsize=c.htsize, shard=(0, 2)) as g:
        with pytest.raises(api.CuclarkError):
            g.route_alloc(2, 1000)                             # no database loaded
        g.load_arrays(sz, ky, lb)
        with pytest.raises(api.CuclarkError):
            g.route_alloc(3, 1000)                             # the handle is shard 0 of 2
        g.route_alloc(2, 1000)
        with pytest.raises(api.CuclarkError):
            g.route_probe()                                    # rank 1 is not connected
        with pytest.raises(api.CuclarkError):
            g.route_scatter(0, 0, 5, 2000)                     # more containers than allocated


def test_routed_local_shards_low_complexity(oracle):
    """LOCAL shards: tie k-mers (periodic sequence, two possible homes that may lie in two shards), both strands, parts
    longer than a chunk: the scatter kernel picks the home as the single-table kernel does, exactly one shard answers."""
    from test_gpu_local_layout import low_complexity_reads
    from oracle.binding import key_bytes_for
    k, T, G = 21, 6, 30_000
    rng = np.random.default_rng(5)
    targets = [synth.genome_codes(13, t, 0, G) for t in range(T)]
    for t in (1, 4):
        for _ in range(30):
            period = int(rng.integers(1, 6))
            pos = int(rng.integers(0, G - 200))
            targets[t][pos:pos + 120] = np.resize(rng.integers(0, 4, period), 120)
    kmers, labels = dbtools.build_entries(targets, k, 0)
    arrays = dbtools.entries_to_arrays(kmers, labels, HTSIZE_LIGHT, key_bytes_for(k, HTSIZE_LIGHT))
    odb = oracle.db_from_arrays(HTSIZE_LIGHT, k, *arrays)
    asc = np.frombuffer(b"ACGT", np.uint8)
    reads = []
    for i in range(600):
        t = int(rng.integers(0, T)); L = int(rng.integers(k - 2, 400)); pos = int(rng.integers(0, G - L))
        codes = targets[t][pos:pos + L]
        if i & 1:
            codes = 3 - codes[::-1]
        reads.append(b">r%d\n" % i + asc[codes].tobytes() + b"\n")
    long_codes = np.concatenate([targets[2][:3000], 3 - targets[3][::-1][:2500]])
    data = b"".join(reads) + low_complexity_reads(rng, 300, 150) + b">long\n" + asc[long_codes].tobytes() + b"\n"
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, k, data, T, 23)
    for n_ranks in (2, 3, 7):
        gf, gr, st, _, entries = run_routed(k, T, HTSIZE_LIGHT, arrays, ptr, cont, n_ranks, 23, layout=3)
        assert entries >= kmers.size                              # tie k-mers live in the overflow table of both homes' shards
        assert np.array_equal(gf, final) and np.array_equal(gr, rows), n_ranks
        assert sum(s["lookups"] for s in st) == lookups == sum(s["probed"] for s in st)


@pytest.mark.parametrize("n_ranks,layout", [(2, 0), (4, 0), (8, 0), (2, 3), (8, 3)])
def test_routed_multi_gpu(oracle, light_c1, n_ranks, layout):
    """One process, one shard per DEVICE: k-mers and labels cross NVLink through peer pointers."""
    if n_gpus() < n_ranks:
        pytest.skip(f"needs {n_ranks} GPUs")
    c = light_c1
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits, 8)
    gf, gr, st, _, _ = run_routed(c.k, c.n_targets, c.htsize, c.arrays, ptr, cont, n_ranks, c.maxhits,
                                  devices=list(range(n_ranks)), layout=layout)
    assert np.array_equal(gf, final) and np.array_equal(gr, rows)
    assert sum(s["probed"] for s in st) == lookups
