#!/usr/bin/env python
"""The product's own multi-GPU path in ONE process (what `cuCLARK -d N` runs), strong scaling + start-up.

BASELINE configs[3] is "read-partitioned replicated-DB throughput sweep 1/2/4/8 B200" over a FIXED read set. bench.py
measures N independent torchrun ranks (weak scaling); this tool drives the in-process path the command line uses:
one host process, N handles, the table loaded ONCE and cloned device to device (`cuclark_clone_table`), the chunks of
one pinned FASTQ buffer dealt over the devices by `cuclark_classify_text_buffer` with N handles, CSV back in file order.

Reports, for N in --gpus: reads/s and lookups/s end to end (wall clock of the call, pinned host text -> pinned host
CSV), the SHA-1 of the CSV (must not depend on N), and the start-up: table build, table cache save/load
(`cuclark_save_table/load_table`), clone time per replica -> time to first read from a cache for N GPUs.

    python tools/inprocess_multi.py --reads 40000000 --gpus 1,2,4,8 > gpurun_out/inprocess.json
"""
import argparse
import hashlib
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
K, G, L = 31, 4_000_000, 150


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--targets", type=int, default=1430)
    ap.add_argument("--reads", type=int, default=40_000_000)
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--chunk-mb", type=int, default=64)
    ap.add_argument("--slots", type=int, default=4)
    ap.add_argument("--no-cache", action="store_true")
    a = ap.parse_args()
    import torch
    from cuclark_b200.api import CuClarkDB, HTSIZE_FULL
    n_dev = torch.cuda.device_count()
    ns = [n for n in (int(x) for x in a.gpus.split(",")) if n <= n_dev]
    T, n = a.targets, a.reads
    out = {"workload": f"k=31, {T} x 4 Mbp targets, {n} x {L} bp FASTQ reads (one pinned buffer), in-process -d N path",
           "devices_present": n_dev, "host_cores": os.cpu_count()}

    # ---- start-up: build once (or load from a cache), clone to the other devices
    g0 = CuClarkDB(K, T, htsize=HTSIZE_FULL, device=0)
    t0 = time.time(); g0.build_synthetic(1, T, G, 0); out["build_synthetic_s"] = time.time() - t0
    st = g0.stats()
    out["table"] = {"bytes": st["table_bytes"], "entries": st["n_entries"], "layout": st["layout"], "overflow_entries": st["n_spilled"]}
    if not a.no_cache:
        d = tempfile.mkdtemp(prefix="cuclark_cache_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        path = os.path.join(d, "t.b200")
        try:
            t0 = time.time(); g0.save_table(path); out["cache_save_s"] = time.time() - t0
            size = os.path.getsize(path)
            g0.close()
            g0 = CuClarkDB(K, T, htsize=HTSIZE_FULL, device=0)
            t0 = time.time(); ok = g0.load_table(path); out["cache_load_s"] = time.time() - t0
            assert ok
            out["cache_bytes"] = size
            out["cache_load_GBps"] = size / out["cache_load_s"] / 1e9
            out["cache_save_GBps"] = size / out["cache_save_s"] / 1e9
        finally:
            try:
                os.remove(path); os.rmdir(d)
            except OSError:
                pass
    handles = [g0]
    clone_s = []
    for dev in range(1, max(ns)):
        h = CuClarkDB(K, T, htsize=HTSIZE_FULL, device=dev)
        t0 = time.time(); h.clone_table_from(g0); clone_s.append(time.time() - t0)
        assert h.stats()["n_entries"] == st["n_entries"]
        handles.append(h)
    out["clone_s_per_replica"] = clone_s
    out["clone_GBps"] = [st["table_bytes"] / s / 1e9 for s in clone_s]
    if "cache_load_s" in out:
        out["time_to_first_read_from_cache_s"] = {str(k): out["cache_load_s"] + sum(clone_s[:k - 1]) for k in ns}

    # ---- one pinned FASTQ buffer, generated on device 0 in pieces
    rec = 16 + 2 * L
    t0 = time.time()
    h_text = torch.empty(n * rec, dtype=torch.uint8, pin_memory=True)
    step = 4_000_000
    torch.cuda.set_device(0)
    d_buf = torch.empty(step * rec, dtype=torch.uint8, device="cuda:0")
    torch.cuda.synchronize()
    for lo in range(0, n, step):
        m = min(step, n - lo)
        g0.synth_fastq_device(2, 1, T, G, lo, m, L, 10, 0, d_buf.data_ptr())
        g0.stats(sync=True)
        h_text[lo * rec:(lo + m) * rec].copy_(d_buf[:m * rec])
    torch.cuda.synchronize()
    del d_buf
    cap = n * 40 + 4096
    h_csv = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    out["reads_ready_s"] = time.time() - t0
    names = [f"T{t:05d}" for t in range(T)]

    runs = {}
    for k in ns:
        def one():
            return handles[0].classify_text_buffer(h_text.data_ptr(), n * rec, h_csv.data_ptr(), cap, names=names,
                                                   chunk_bytes=a.chunk_mb << 20, n_slots=a.slots, peers=handles[1:k])
        one()                                   # slots of every device get allocated
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            ln, ts = one()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        sha = hashlib.sha1(h_csv[:ln].numpy().tobytes()).hexdigest()
        runs[str(k)] = {"s": best, "reads_per_s": n / best, "lookups_per_s": ts["lookups"] / best, "h2d_GBps": n * rec / best / 1e9,
                        "csv_sha1": sha, "csv_bytes": ln, "n_reads": ts["n_reads"]}
    out["runs"] = runs
    base = runs[str(ns[0])]["reads_per_s"] / ns[0]
    out["strong_scaling_efficiency"] = {k: v["reads_per_s"] / (base * int(k)) for k, v in runs.items()}
    out["csv_identical_for_every_n"] = len({v["csv_sha1"] for v in runs.values()}) == 1
    print(json.dumps(out), flush=True)
    for h in handles:
        h.close()


if __name__ == "__main__":
    main()
