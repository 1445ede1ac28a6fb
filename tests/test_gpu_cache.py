"""Table cache: cuclark_save_table / cuclark_load_table.

A table streamed back from its cache must classify exactly as the table built from
.sz/.ky/.lb does (and as the oracle does); a cache that is corrupt, truncated, foreign or
stale must be refused so that the caller falls back to the database files.
"""
import os

import numpy as np
import pytest

from cuclark_b200.api import CuClarkDB
from test_gpu_parity import make_gpu, oracle_expect

pytestmark = pytest.mark.gpu


def write_db_files(case, base):
    sz, ky, lb = case.arrays
    sz.tofile(base + ".sz"); ky.tofile(base + ".ky"); lb.tofile(base + ".lb")


@pytest.mark.parametrize("layout,load", [(1, 0.0), (2, 0.0), (2, 4.2)])
def test_cache_round_trip_is_bit_exact(oracle, light_small, tmp_path, layout, load):
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ptr, cont, final, rows, lookups = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits)
    path = str(tmp_path / "t.b200")
    with CuClarkDB(c.k, c.n_targets, htsize=c.htsize, layout=layout, bucket_load=load) as g:
        g.load_arrays(sz, ky, lb)
        st0 = g.stats()
        if load:
            assert st0["n_spilled"] > 0          # the overflow table travels too
        g.save_table(path)
    assert os.path.getsize(path) == 192 + st0["table_bytes"]
    with CuClarkDB(c.k, c.n_targets, htsize=c.htsize) as g:
        assert g.load_table(path)
        st1 = g.stats()
        for key in ("n_entries", "n_buckets", "n_local_buckets", "table_bytes", "n_spilled", "n_spill_buckets", "layout"):
            assert st1[key] == st0[key], key
        gf, gr = g.classify(ptr, cont, want_rows=True)
        assert np.array_equal(gf, final) and np.array_equal(gr, rows)
        assert g.stats()["lookups"] == lookups
        # a second save of the loaded table is byte-identical
        g.save_table(path + "2")
    assert open(path, "rb").read() == open(path + "2", "rb").read()


def test_cache_refuses_corrupt_foreign_and_stale_files(light_small, tmp_path):
    c = light_small
    base = str(tmp_path / "db")
    write_db_files(c, base)
    path = base + ".b200"
    with CuClarkDB(c.k, c.n_targets, htsize=c.htsize) as g:
        assert g.read(base)
        g.save_table(path)
        blob = open(path, "rb").read()
        assert g.load_table(path, src_base=base)                       # matches its source files
        assert not g.load_table(str(tmp_path / "missing.b200"))        # CUCLARK_ERR_IO
        # flipped payload bit -> checksum
        bad = bytearray(blob); bad[192 + len(blob) // 2] ^= 0x10
        open(path, "wb").write(bad)
        assert not g.load_table(path)
        # truncated
        open(path, "wb").write(blob[:-32])
        assert not g.load_table(path)
        # not a cache at all
        open(path, "wb").write(b"Object_ID,Length\n" * 100)
        assert not g.load_table(path)
        # stale: a database rebuilt in place with as many k-mers keeps its sizes; the modification time gives it away
        open(path, "wb").write(blob)
        assert g.load_table(path, src_base=base)
        import os
        st_ky = os.stat(base + ".ky")
        os.utime(base + ".ky", ns=(st_ky.st_atime_ns, st_ky.st_mtime_ns + 2_000_000_000))
        assert not g.load_table(path, src_base=base)
        os.utime(base + ".ky", ns=(st_ky.st_atime_ns, st_ky.st_mtime_ns))
        assert g.load_table(path, src_base=base)
        # stale: the source files changed size
        with open(base + ".lb", "ab") as f:
            f.write(b"\0\0")
        assert not g.load_table(path, src_base=base)
        assert g.load_table(path)                                      # without the source check it is still a valid cache
        # other sampling factor
        assert not g.load_table(path, mod_collision=2)
        # after a refused load the handle holds no table: classify must fail loudly, not return zeros
        assert not g.load_table(str(tmp_path / "missing.b200"))
        with pytest.raises(Exception):
            g.classify(np.zeros(2, np.uint32), np.zeros(4, np.uint16))
    # other k / other number of targets
    with CuClarkDB(c.k, c.n_targets + 1, htsize=c.htsize) as g:
        assert not g.load_table(path)
    with CuClarkDB(c.k - 2, c.n_targets, htsize=c.htsize) as g:
        assert not g.load_table(path)


def test_cache_of_a_shard(oracle, light_small, tmp_path):
    """Table-partitioned mode: every shard caches its own slice."""
    c = light_small
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ptr, cont, final, rows, _ = oracle_expect(oracle, odb, c.k, c.reads_bytes, c.n_targets, c.maxhits)
    n = len(final)
    import torch
    parts = []
    for i in range(2):
        path = str(tmp_path / f"shard{i}.b200")
        with CuClarkDB(c.k, c.n_targets, htsize=c.htsize, shard=(i, 2)) as g:
            g.load_arrays(sz, ky, lb)
            g.save_table(path)
        with CuClarkDB(c.k, c.n_targets, htsize=c.htsize, shard=(1 - i, 2)) as g:
            assert not g.load_table(path)                              # the other shard's file
        with CuClarkDB(c.k, c.n_targets, htsize=c.htsize, shard=(i, 2)) as g:
            assert g.load_table(path)
            _, gr = g.classify(ptr, cont, want_rows=True)
            parts.append(gr)
    with make_gpu(c) as g:
        d_parts = torch.from_numpy(np.stack(parts).view(np.int16)).cuda()
        d_final = torch.empty((n, 5), dtype=torch.int16, device="cuda")
        d_rows = torch.empty((n, g.row_size), dtype=torch.int16, device="cuda")
        g.merge_rows_device(d_parts.data_ptr(), 2, n, d_rows.data_ptr(), d_final.data_ptr())
        g.stats(sync=True)
        torch.cuda.synchronize()
    assert np.array_equal(d_final.cpu().numpy().view(np.uint16), final)
    assert np.array_equal(d_rows.cpu().numpy().view(np.uint16), rows)
