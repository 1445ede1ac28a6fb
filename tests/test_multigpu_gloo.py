"""CPU, world_size 2, gloo: the host-side logic of both multi-GPU modes.

The per-shard "kernel" is played by the oracle restricted to a range of reference buckets (any partition
of the canonical k-mer space gives bit-identical merged results: each k-mer lives in exactly one shard)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import conftest
        from cuclark_b200 import multigpu
        from oracle.binding import Oracle
        c = conftest.get_case("light_small")
        orc = Oracle()
        sz, ky, lb = c.arrays
        odb = orc.db_from_arrays(c.htsize, c.k, sz, ky, lb)
        ix, buf = orc.index(c.reads_bytes, 1)
        n_total = ix.n - ix.n % world            # equal slices
        ptr, cont = orc.pack(ix, buf, c.k, 0, n_total)
        full_final, full_rows, _ = orc.classify(odb, ptr, cont, c.n_targets, c.maxhits)

        # --- mode A: read-partitioned, no collective on the data path
        lo, hi = multigpu.read_range(n_total, rank, world)
        p2, c2 = orc.pack(ix, buf, c.k, lo, hi - lo)
        fa, ra, _ = orc.classify(odb, p2, c2, c.n_targets, c.maxhits)
        ok_a = np.array_equal(fa, full_final[lo:hi]) and np.array_equal(ra, full_rows[lo:hi])

        # --- mode B: table-partitioned. This shard = reference buckets [r_lo, r_hi)
        r_lo, r_hi = c.htsize * rank // world, c.htsize * (rank + 1) // world
        _, rows_shard, _ = orc.classify(odb, ptr, cont, c.n_targets, c.maxhits, part=(r_lo, r_hi))
        t = torch.from_numpy(rows_shard.astype(np.int16))
        parts = multigpu.exchange_rows(t, world).numpy().view(np.uint16)     # [world, n, pitch]
        merged = parts[0]
        for g in range(1, world):
            merged = orc.merge_rows(np.ascontiguousarray(merged), np.ascontiguousarray(parts[g]), c.maxhits)
        final_b = orc.result_from_rows(np.ascontiguousarray(merged), c.maxhits)
        ok_b = np.array_equal(merged, full_rows[lo:hi]) and np.array_equal(final_b, full_final[lo:hi])

        # --- gather of fixed-size packed reads
        local = torch.arange(rank * 10, rank * 10 + 10, dtype=torch.int16)
        allr = multigpu.gather_fixed_reads(local, world)
        ok_g = allr.tolist() == list(range(0, 10 * world))
        q.put((rank, ok_a, ok_b, ok_g, int((full_final[lo:hi, 1] > 0).sum())))
    finally:
        dist.destroy_process_group()


def test_read_range_covers_everything():
    from cuclark_b200 import multigpu
    for n, w in [(10, 3), (7, 8), (100, 4), (0, 2)]:
        spans = [multigpu.read_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_two_rank_gloo_both_modes():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0, "a rank failed (see its traceback above)"
    res = [q.get(timeout=10) for _ in range(world)]
    for rank, ok_a, ok_b, ok_g, classified in res:
        assert ok_a, f"rank {rank}: read-partitioned slice differs"
        assert ok_b, f"rank {rank}: table-partitioned merge differs"
        assert ok_g and classified > 100
