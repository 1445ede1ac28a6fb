#!/usr/bin/env python
"""Table cache timing on one B200: build a synthetic k=31 table on the device, write it with
cuclark_save_table, stream it back with cuclark_load_table, classify with both and compare."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--targets", type=int, default=300)
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--dir", default="/dev/shm" if os.path.isdir("/dev/shm") else None)
    a = ap.parse_args()
    import torch
    from cuclark_b200.api import CuClarkDB, HTSIZE_FULL
    K, G, L = 31, 4_000_000, 150
    per = 1 + (L + 7) // 8
    n = a.reads
    d_ptr = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    d_cont = torch.empty(n * per, dtype=torch.int16, device="cuda")
    f0 = torch.empty(n * 5, dtype=torch.int16, device="cuda")
    f1 = torch.empty(n * 5, dtype=torch.int16, device="cuda")
    path = os.path.join(tempfile.mkdtemp(prefix="cache_", dir=a.dir), "t.b200")
    out = {}
    try:
        with CuClarkDB(K, a.targets, htsize=HTSIZE_FULL) as g:
            t0 = time.time(); g.build_synthetic(1, a.targets, G, 0); out["build_s"] = time.time() - t0
            st = g.stats()
            g.synth_reads_device(2, 1, a.targets, G, 0, n, L, 10, 0, d_ptr.data_ptr(), d_cont.data_ptr(), 0)
            g.classify_device(d_ptr.data_ptr(), d_cont.data_ptr(), n, f0.data_ptr(), 0, 0)
            g.stats(sync=True)
            t0 = time.time(); g.save_table(path); out["save_s"] = time.time() - t0
        size = os.path.getsize(path)
        with CuClarkDB(K, a.targets, htsize=HTSIZE_FULL) as g:
            t0 = time.time(); ok = g.load_table(path); out["load_s"] = time.time() - t0
            assert ok
            t0 = time.time(); ok = g.load_table(path); out["load_again_s"] = time.time() - t0
            g.classify_device(d_ptr.data_ptr(), d_cont.data_ptr(), n, f1.data_ptr(), 0, 0)
            g.stats(sync=True)
            st1 = g.stats()
        torch.cuda.synchronize()
        out.update(entries=st["n_entries"], table_bytes=st["table_bytes"], file_bytes=size, where=os.path.dirname(path),
                   save_GBps=size / out["save_s"] / 1e9, load_GBps=size / out["load_s"] / 1e9,
                   load_again_GBps=size / out["load_again_s"] / 1e9,
                   results_identical=bool(torch.equal(f0, f1)), stats_identical=all(st[k] == st1[k] for k in ("n_entries", "n_spilled", "table_bytes")))
        print(json.dumps(out), flush=True)
    finally:
        try:
            os.remove(path); os.rmdir(os.path.dirname(path))
        except OSError:
            pass


if __name__ == "__main__":
    main()
