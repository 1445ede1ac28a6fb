#!/usr/bin/env python
"""bench.py — k-mer lookups/s and reads/s of the classification hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
cuCLARK full variant, k=31, synthetic bacterial-scale database of 1,430 targets x
4 Mbp = 5.72 G target-specific 31-mers (~36 GB in the reference's .sz/.ky/.lb
form; ~77 GB as 32-byte sector buckets in HBM), 10 M single-end 150 bp reads
(90 % sampled from the targets, 10 % random). A step = one pass of the hot path
(k-mer extraction, table probe, per-read histogram, top-2) over the 10 M reads.
Database and reads are generated ON the device from seeded counter-based hashes
(cuclark_b200/synth.py is the numpy twin, checked in tests/).

N > 1 (torchrun, one rank per GPU): read-partitioned mode against a replicated
table, the reads are split over the ranks, no collective on the data path
("scaling": "weak": every rank classifies --reads reads).

Output: ONE JSON line on rank 0. `value` = lookups/s with inputs resident in HBM;
`e2e` = the same through the public text call with pinned HOST buffers (FASTQ bytes in,
CSV bytes out; H2D and D2H inside the timed region), `e2e_packed` through the batch API
with packed reads; `roofline` for the classify kernel (32 algorithmic bytes per lookup);
`cpu_baseline` = the oracle port of the whole path on the host cores on a bounded sample,
with the reference's own hTable::find beside it as `reference_lookup` (N=1 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 31
GENOME_LEN = 4_000_000
READ_LEN = 150
DB_SEED, READ_SEED = 1, 2
BYTES_PER_LOOKUP = 32          # one sector bucket; SURVEY.md section 8(d), DESIGN.md


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--targets", type=int, default=int(os.environ.get("CUCLARK_BENCH_TARGETS", 1430)))
    ap.add_argument("--reads", type=int, default=int(os.environ.get("CUCLARK_BENCH_READS", 10_000_000)))
    ap.add_argument("--pct-random", type=int, default=10)
    ap.add_argument("--sub-per-10k", type=int, default=0,
                    help="substitution errors per 10,000 bases of the sampled reads (configs[2] uses 100 = 1%%)")
    ap.add_argument("--cpu-targets", type=int, default=8, help="targets of the CPU baseline's (smaller) database")
    ap.add_argument("--cpu-reads", type=int, default=400_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ref-lookup", action="store_true",
                    help="skip the second host baseline: the reference's own hTable::find (oracle/_ref) on the CPU sample (~45 s)")
    ap.add_argument("--chunk-mb", type=int, default=64, help="text pipeline: MiB of text per chunk")
    ap.add_argument("--slots", type=int, default=4, help="text pipeline: chunks in flight")
    ap.add_argument("--layout", type=int, default=int(os.environ.get("CUCLARK_BENCH_LAYOUT", 0)),
                    help="device table layout: 0 auto, 1 narrow, 2 wide, 3 local (minimizer-addressed lines)")
    ap.add_argument("--load", type=float, default=0.0, help="entries per bucket (0 = the layout's default)")
    ap.add_argument("--extended-reads", type=int, default=500_000, help="reads of the --extended e2e leg (0 = skip)")
    ap.add_argument("--no-table-mode", action="store_true", help="N > 1: skip the table-partitioned sub-records")
    ap.add_argument("--c5-targets", type=int, default=11000,
                    help="N > 1: targets of the configs[4] table (44 G entries, ~600 GB) measured when the shards fit; 0 = skip")
    ap.add_argument("--mode", default="read", choices=["read", "table"],
                    help="multi-GPU mode: read-partitioned/replicated table, or table-partitioned + NCCL row exchange")
    return ap.parse_args()


# ---------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clocks and throttle reasons through NVML during the timed region."""

    def __init__(self, index: int, period_s: float = 0.05):
        self.index, self.period, self.samples, self.reasons = index, period_s, [], set()
        self.max_mhz, self._stop, self._thr, self.err = None, threading.Event(), None, None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML indices follow PCI order; honour CUDA_VISIBLE_DEVICES when it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception as e:      # pragma: no cover
            self.err = repr(e)
            return
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}

        def loop():
            while not self._stop.is_set():
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for nm, bit in names.items():
                        if r & bit:
                            self.reasons.add(nm)
                except Exception as e:   # pragma: no cover
                    self.err = repr(e)
                    return
                time.sleep(self.period)

        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1.0)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.err:
            out["error"] = self.err
        return out


def bind_to_gpu_numa(local: int) -> dict:
    """Pin this rank's host threads (and with them its first-touch pinned buffers) to the NUMA node of its GPU, so
    that N ranks do not all stage their input out of node 0. Best effort: cgroup cpusets and VMs may hide the node."""
    import torch
    out = {"node": None, "cpus": None}
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read())
        out["node"] = node
        if node < 0:
            return out
        cpus = set()
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            for part in fh.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
            out["cpus"] = len(use)
        else:
            out["cpus"] = 0          # the node's CPUs are outside this container's cpuset: left as it is
    except Exception as e:           # pragma: no cover
        out["error"] = repr(e)
    return out


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json: device copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload_key: str, lookups: int):
    """DRAM bytes per classify launch: bytes/lookup from the committed `ncu --set full` capture (profiles/traffic.json,
    written by tools/refresh_traffic.py) x the lookups of one launch. The capture names the SHA-1 of the kernel sources
    it was taken from: if the sources differ (the kernel changed since) or no capture exists the figure is None, with
    the reason beside it."""
    import hashlib
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f)[workload_key]
        h = hashlib.sha1()
        for src in ("classify.cu", "common.cuh", "hits.cuh", "kmerwin.cuh"):
            h.update(open(os.path.join(ROOT, "cuclark_b200", "csrc", src), "rb").read())
        if e.get("kernel_source_sha1") != h.hexdigest():
            return None, "stale: the kernel sources changed since the ncu capture (tools/refresh_traffic.py)"
        return e["dram_bytes_per_lookup"] * lookups, f"ncu --set full, {e.get('report')}, {e['dram_bytes_per_lookup']:.1f} B per lookup"
    except Exception as ex:
        return None, f"no capture: {ex!r}"


# ---------------------------------------------------------------- CPU legs (oracle: checker/baseline only)
def cpu_sample(args, threads: int):
    """Reads + port database for the CPU legs: a bounded sample of the same workload."""
    from cuclark_b200 import synth
    from oracle.binding import HTSIZE_FULL, Oracle
    orc = Oracle()
    t0 = time.time()
    db = orc.db_build_synth(DB_SEED, args.cpu_targets, GENOME_LEN, K, HTSIZE_FULL, 0, threads)
    build_s = time.time() - t0
    codes, *_ = synth.read_codes(READ_SEED, args.cpu_reads, READ_LEN, args.cpu_targets, GENOME_LEN, DB_SEED,
                                 pct_random=args.pct_random)
    data = synth.reads_fastq(codes)
    return orc, db, data, build_s


def cpu_step(orc, db, data, n_targets, threads):
    """index + 2-bit pack + extract + lookup + histogram + top-2 on the host; returns (seconds, lookups, reads)."""
    t0 = time.perf_counter()
    ix, buf = orc.index(data, threads)
    ptr, cont = orc.pack(ix, buf, K)
    final, _, lookups = orc.classify(db, ptr, cont, n_targets, 15, want_rows=False, threads=threads)
    dt = time.perf_counter() - t0
    n = ptr.size - 1
    orc.free_index(ix)
    return dt, lookups, n


def reference_lookup_leg(orc, db, data, threads: int):
    """Host baseline B2: the reference's OWN host table — hTable::read + hTable::find
    (src/hashTable_hh.hh:476-513, 666-946) compiled from /root/reference into oracle/_ref/libref_lookup_full.so —
    over every k-mer of the sample reads, OpenMP over k-mers. Writing the 1.6 GB .sz file and filling the reference's
    25.8 GB empty table take most of a minute (--no-ref-lookup skips the leg)."""
    import shutil
    import tempfile
    from oracle import dbtools
    from oracle.binding import HTSIZE_FULL, RefLookup
    km, lb = db.entries()
    sz, ky, lbl = dbtools.entries_to_arrays(km, lb, HTSIZE_FULL, 4)
    d = tempfile.mkdtemp(prefix="cuclark_b2_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        base = os.path.join(d, "db")
        dbtools.write_db_files(base, sz, ky, lbl)
        del sz, ky, lbl
        t0 = time.time()
        ref = RefLookup(light=False).open(base, K, 1, threads)
        load_s = time.time() - t0
        ix, buf = orc.index(data, threads)
        ptr, cont = orc.pack(ix, buf, K)
        kmers = orc.extract(ptr, cont, K)
        orc.free_index(ix)
        ref.query(kmers[:200_000], threads)
        t0 = time.perf_counter()
        out, hits = ref.query(kmers, threads)
        dt = time.perf_counter() - t0
        exp, _ = db.query(kmers, threads)
        same = bool(np.array_equal(out, exp))
        ref.close()
        return {"value": kmers.size / dt, "unit": "lookups/s", "cores": threads, "kind": "reference",
                "what": "hTable::find of the reference (oracle/_ref/libref_lookup_full.so), stage 3 only, OpenMP over k-mers",
                "lookups": int(kmers.size), "hits": int(hits), "table_load_s": load_s, "equals_port": same}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def reference_host_path_leg(threads: int):
    """Host baseline B1: the UNMODIFIED reference binary (oracle/_ref/cuCLARK-l, compiled from
    /root/reference) on this box. (i) its database build — makeSpecificTargetSets/addElement/RemoveCommon/write, serial
    host code, 1 core — on 8 x 1 Mbp targets; (ii) its classification of 400,000 x 100 bp reads with -n <cores>: the
    "Assignment time" it prints covers its OpenMP index + 2-bit pack loops (src/CuCLARK_hh.hh:1340-1727), its own GPU
    kernels and its serial CSV writer. Bounded to a few seconds; None if the binary did not travel."""
    import re
    import shutil
    import tempfile
    from cuclark_b200 import synth
    exe = os.path.join(ROOT, "oracle", "_ref", "cuCLARK-l")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/cuCLARK-l not built (needs /root/reference at build time)"}
    n_t, glen, n_reads, rlen = 8, 1_000_000, 400_000, 100
    d = tempfile.mkdtemp(prefix="cuclark_b1_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        os.makedirs(os.path.join(d, "tg")); os.makedirs(os.path.join(d, "db"))
        asc = np.frombuffer(b"ACGT", np.uint8)
        with open(os.path.join(d, "targets.txt"), "w") as tf:
            for t in range(n_t):
                p = os.path.join(d, "tg", f"T{t}.fa")
                synth.write_fasta(p, f"T{t}", asc[synth.genome_codes(DB_SEED, t, 0, glen)].tobytes())
                tf.write(f"{p} T{t}\n")
        codes, *_ = synth.read_codes(READ_SEED, n_reads, rlen, n_t, glen, DB_SEED, pct_random=10)
        with open(os.path.join(d, "reads.fa"), "wb") as f:
            f.write(synth.reads_fasta(codes))
        cmd = [exe, "-T", "targets.txt", "-D", "db/", "-O", "reads.fa", "-R", "out", "-n", str(threads), "-b", str(threads)]
        walls, assigns = [], []
        for _ in range(2):                 # first run builds the database on the host, second finds it
            t0 = time.time()
            p = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=600)
            walls.append(time.time() - t0)
            m = re.search(r"Assignment time: ([0-9.eE+-]+) s", p.stdout)
            assigns.append(float(m.group(1)) if m else None)
        m = re.search(r"(\d+) \d+-mers successfully stored", p.stderr) or re.search(r"(\d+) 27-mers", p.stderr)
        kept = n_t * (glen // 27) // 4
        out = {"kind": "reference", "binary": "oracle/_ref/cuCLARK-l (unmodified reference, k=27, -g 4)",
               "db_build_s": walls[0] - walls[1], "db_build_cores": 1, "db_build_nt_per_s": n_t * glen / max(walls[0] - walls[1], 1e-9),
               "db_entries": kept, "classify_wall_s": walls[1], "assignment_s": assigns[1], "cores": threads,
               "reads_per_s": n_reads / assigns[1] if assigns[1] else None,
               "lookups_per_s": n_reads * (rlen - 27 + 1) / assigns[1] if assigns[1] else None,
               "sample": f"{n_t} x 1 Mbp targets, {n_reads} x {rlen} bp FASTA reads"}
        if assigns[1] is None:
            # on sm_100 the light reference binary finishes ("Done.", CSV complete) and then dies in its teardown
            # (CUERR 'invalid argument' at CuClarkDB.cu:292, freeBatchMemory) before it prints its speed line: the rate
            # is then taken from the process wall clock, which includes its CUDA start-up and database load
            out["exit_code"] = p.returncode
            out["stderr_tail"] = p.stderr[-160:]
            if os.path.exists(os.path.join(d, "out.csv")) and "Done." in p.stderr:
                out["reads_per_s"] = n_reads / walls[1]
                out["lookups_per_s"] = n_reads * (rlen - 27 + 1) / walls[1]
                out["rate_from"] = "process wall clock (speed line not printed)"
        return out
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_reference(args, rank: int):
    """--impl reference: the reference algorithm on the host cores (the oracle port)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    orc, db, data, build_s = cpu_sample(args, threads)
    for _ in range(min(args.warmup, 1)):
        cpu_step(orc, db, data, args.cpu_targets, threads)
    tot_t = tot_l = tot_r = 0
    for _ in range(args.steps):
        dt, lk, n = cpu_step(orc, db, data, args.cpu_targets, threads)
        tot_t += dt; tot_l += lk; tot_r += n
    val = tot_l / tot_t
    sample = (f"{args.cpu_reads} x {READ_LEN} bp reads per step against a {args.cpu_targets} x 4 Mbp "
              f"({db.size / 1e6:.0f} M 31-mer) database in the reference's table layout (HTSIZE 1610612741); "
              f"index+pack+extract+lookup+histogram+top-2, OpenMP")
    line = {
        "impl": "reference", "metric": "kmer_lookups_per_s", "value": val, "unit": "lookups/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "reads_per_s": tot_r / tot_t,
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": val, "unit": "lookups/s", "cores": threads, "kind": "port", "sample": sample,
                         "db_build_s": build_s},
        "e2e": {"value": val, "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"config2: cuCLARK k=31, synthetic DB {args.targets} targets x 4 Mbp "
                        f"(~{args.targets * (GENOME_LEN - K + 1) / 1e9:.2f} G 31-mers), "
                        f"{args.reads} x {READ_LEN} bp single-end reads per GPU, {args.pct_random}% random"
                        + (f", {args.sub_per_10k / 100:g}% substitutions" if args.sub_per_10k else ""),
            "k": K, "targets": args.targets, "reads_per_gpu": args.reads, "read_len": READ_LEN,
            "mode": ("single GPU" if world == 1 else
                     "table-partitioned, rows exchanged over NCCL all-to-all" if args.mode == "table"
                     else "read-partitioned, replicated table"),
            "l2_policy": "inputs larger than L2 (packed reads 400 MB, table >> 126 MB); no flush needed"}


def run_table_mode(args, rank, world, local, targets, d_ptr, d_cont, n, ts, ref_final=None, steps=None, layout=0):
    """Table-partitioned run of the SAME reads: rank r holds shard r of the table (hashed sectors) and its own reads;
    k-mers travel to their shard and labels come back over NVLink inside the probe kernel (csrc/route.cu), with two
    stream-ordered NCCL all-reduces of one float as the barriers between scatter | probe | gather."""
    import torch
    import torch.distributed as dist
    from cuclark_b200 import synth
    from cuclark_b200.api import CuClarkDB, HTSIZE_FULL
    steps = steps or args.steps
    per = 1 + (READ_LEN + 7) // 8
    stream = ts.cuda_stream
    g = CuClarkDB(K, targets, htsize=HTSIZE_FULL, device=local, shard=(rank, world), layout=layout)
    t0 = time.time()
    g.build_synthetic(DB_SEED, targets, GENOME_LEN, 0)
    build_s = time.time() - t0
    st = g.stats()
    g.route_alloc(world, n * per)
    mine = g.route_export()
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    for r in range(world):
        if r != rank:
            g.route_import(r, handles[r])
    d_final = torch.zeros(n * 5, dtype=torch.int16, device="cuda")
    tok = torch.zeros(1, device="cuda")
    torch.cuda.synchronize()
    dist.barrier()

    def step(marks=None):
        def mark():
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(ts)
                marks.append(e)
        with torch.cuda.stream(ts):
            mark()
            g.route_scatter(d_ptr.data_ptr(), d_cont.data_ptr(), n, n * per, stream)
            mark()
            dist.all_reduce(tok)                      # every rank's k-mers are in its arena
            mark()
            g.route_probe(stream)
            mark()
            dist.all_reduce(tok)                      # every label is back with the rank that asked
            mark()
            g.route_gather(d_ptr.data_ptr(), d_cont.data_ptr(), n, n * per, d_final.data_ptr(), 0, stream)
            mark()

    for _ in range(3):
        step()
    torch.cuda.synchronize(); dist.barrier()
    from cuclark_b200 import api as _api
    l0 = _api.kernel_launches()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record(ts)
    for _ in range(steps):
        step()
    ev[1].record(ts)
    launches = _api.kernel_launches() - l0
    torch.cuda.synchronize(); dist.barrier()
    ms = ev[0].elapsed_time(ev[1]) / steps
    # one more step with an event after every phase (this rank's view; the waits include the slower ranks)
    marks = []
    step(marks)
    torch.cuda.synchronize(); dist.barrier()
    names = ["scatter", "wait_all_scattered", "probe", "wait_all_probed", "gather"]
    phase_ms = {nm: marks[i].elapsed_time(marks[i + 1]) for i, nm in enumerate(names)}
    g.stats(sync_stream=stream, sync=True)
    rs = g.route_stats()
    f = d_final.view(n, 5).cpu().numpy().view(np.uint16)
    same = bool(np.array_equal(f, ref_final)) if ref_final is not None else None
    # ground truth from the generator (as for the read-partitioned rows)
    n_chk = n if not args.sub_per_10k else min(n, 1_000_000)
    tgt, clean = synth.read_truth(READ_SEED, rank * n, n_chk, READ_LEN, targets, args.pct_random, args.sub_per_10k, K)
    exp = np.zeros((n_chk, 5), np.uint16)
    smp = (tgt >= 0) & (clean > 0)
    exp[smp, 0] = clean[smp]; exp[smp, 1] = tgt[smp] + 1; exp[smp, 2] = clean[smp]
    bad = int(((f[:n_chk] != exp).any(axis=1)).sum())
    n_entries_all = float(targets) * (GENOME_LEN - K + 1)
    bound = int(10 + n_chk * READ_LEN * 4 * n_entries_all / 4.0 ** K + n_chk * 120 * 2 * n_entries_all / 4.0 ** K)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor([float(rs["lookups"]), float(rs["probed"]), float(rs["blocks_remote"]), float(st["table_bytes"]),
                      float(bad), float(0 if same in (True, None) else 1), float(rs["err"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    lookups_all, probed_all, blocks_remote, table_bytes, bad_all, differ, err = c.tolist()
    g.close()
    del d_final
    torch.cuda.empty_cache()
    ms_max = t.item()
    return {"value": lookups_all / (ms_max * 1e-3), "unit": "lookups/s", "ms_per_step": ms_max, "steps": steps,
            "targets": targets, "table_bytes_all_shards": table_bytes, "table_bytes_per_gpu": table_bytes / world,
            "layout": {1: "narrow", 2: "wide", 3: "local"}.get(st["layout"]), "shard_build_s": build_s,
            "lookups_per_step": lookups_all, "probed_per_step": probed_all,
            "nvlink_bytes_per_lookup": blocks_remote * 256 * (14 if st["layout"] == 3 else 10) / max(lookups_all, 1.0),
            "routing_buffers_bytes_per_gpu": rs["region_bytes"] + rs["map_bytes"],
            "rows_equal_read_partitioned": (differ == 0) if ref_final is not None else None,
            "rows_equal_ground_truth": bool(bad_all <= bound * world), "ground_truth_mismatches": int(bad_all),
            "gpu_launches_per_rank": launches, "route_err": int(err), "phase_ms_rank0": phase_ms,
            "overflow_entries_this_shard": st["n_spilled"],
            "path": "reads partitioned, table partitioned by " + ("line range (LOCAL shards: minimizer lines, entries = sector + key, 12 B)"
                                                                   if st["layout"] == 3 else "bucket range (hashed shards, entries = k-mer, 8 B)") +
                    "; k_route_scatter | all-reduce | k_route_probe_tma (TMA pulls of the peers' blocks, peer stores of labels "
                    "over NVLink) | all-reduce | k_route_gather"}


# ---------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from cuclark_b200.api import CuClarkDB, HTSIZE_FULL

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the host baseline)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        # NCCL prints its version banner on stdout at some debug levels; stdout carries the JSON line only
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ.pop("NCCL_DEBUG", None)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n, T = args.reads, args.targets
    table_mode = args.mode == "table" and world > 1
    g = CuClarkDB(K, T, htsize=HTSIZE_FULL, device=local, shard=(rank, world) if table_mode else (0, 1),
                  layout=args.layout, bucket_load=args.load)
    t0 = time.time()
    g.build_synthetic(DB_SEED, T, GENOME_LEN, 0)
    build_s = time.time() - t0
    st = g.stats()

    per = 1 + (READ_LEN + 7) // 8
    d_ptr = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    d_cont = torch.empty(n * per, dtype=torch.int16, device="cuda")
    d_final = torch.empty(n * 5, dtype=torch.int16, device="cuda")
    ts = torch.cuda.Stream()
    stream = ts.cuda_stream
    # rank r classifies reads [r*n, (r+1)*n) of the global read set
    g.synth_reads_device(READ_SEED, DB_SEED, T, GENOME_LEN, rank * n, n, READ_LEN, args.pct_random, args.sub_per_10k,
                         d_ptr.data_ptr(), d_cont.data_ptr(), stream)
    g.stats(sync_stream=stream, sync=True)

    if table_mode:
        # every rank probes ITS shard for ALL reads; rows of rank r's reads travel to rank r
        from cuclark_b200 import multigpu
        pitch = g.row_size
        if world * n * per >= 2 ** 32:
            raise SystemExit("table mode: world*reads*containers exceeds 2^32; lower --reads")
        all_ptr = (torch.arange(world * n + 1, dtype=torch.int64, device="cuda") * per).to(torch.int32)
        rows_all = torch.empty((world * n, pitch), dtype=torch.int16, device="cuda")

        def step():
            with torch.cuda.stream(ts):
                all_cont = multigpu.gather_fixed_reads(d_cont, world)          # reads to every GPU
                g.classify_device(all_ptr.data_ptr(), all_cont.data_ptr(), world * n, 0, rows_all.data_ptr(), stream)
                parts = multigpu.exchange_rows(rows_all, world)                # one all-to-all of sparse rows
                g.merge_rows_device(parts.data_ptr(), world, n, 0, d_final.data_ptr(), stream)
                return parts
    else:
        def step():
            g.classify_device(d_ptr.data_ptr(), d_cont.data_ptr(), n, d_final.data_ptr(), 0, stream)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    from cuclark_b200 import api as _api
    launches0 = _api.kernel_launches()
    ev[0].record(ts)
    for i in range(args.steps):
        step()
        ev[i + 1].record(ts)
    launches = _api.kernel_launches() - launches0       # counted by the library at its launch sites
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    lookups = g.stats(sync_stream=stream, sync=True)["lookups"]
    if table_mode:
        lookups //= world          # the kernel saw every rank's reads; count each read's k-mers once
        args.no_e2e = True

    # ---- e2e: pinned host buffers -> batch API -> host results --------------------
    e2e = None
    if not args.no_e2e:
        nb = 16
        bn = (n + nb - 1) // nb
        h_ptr = d_ptr.cpu().numpy().view(np.uint32)
        h_cont = d_cont.cpu().numpy().view(np.uint16)
        # the caller's packed reads live in the library's pinned buffers (as the reference's pack
        # loop writes them, src/CuCLARK_hh.hh:1616-1708); filled once, outside the timed region
        views = g.malloc(nb, bn, bn * per, is_extended=False)
        spans = []
        for b in range(nb):
            lo, hi = b * bn, min(n, (b + 1) * bn)
            m = hi - lo
            vp, vc, _, _ = views[b]
            np.subtract(h_ptr[lo:hi + 1], h_ptr[lo], out=vp[:m + 1])
            vc[:m * per] = h_cont[lo * per:hi * per]
            spans.append((lo, hi))

        def e2e_pass():
            for b, (lo, hi) in enumerate(spans):
                g.readyBatch(b, hi - lo, (hi - lo) * per)
                g.queryBatch(b)                      # async: H2D + kernels + D2H on the batch's stream
            for b in range(nb):
                g.waitForBatch(b)                    # results are now in pinned host memory

        e2e_pass()
        barrier()
        t0 = time.perf_counter()
        reps = max(2, min(args.steps, 4))
        for _ in range(reps):
            e2e_pass()
        barrier()
        e2e_s = (time.perf_counter() - t0) / reps
        out = np.concatenate([views[b][2][:hi - lo] for b, (lo, hi) in enumerate(spans)])
        same = bool(np.array_equal(out, d_final.view(n, 5).cpu().numpy().view(np.uint16)))
        g.freeBatchMemory()
        e2e = {"s": e2e_s, "same_as_device_path": same,
               "h2d": int((n + nb) * 4 + n * per * 2), "d2h": int(n * 10 + nb * 32)}

    # ---- e2e, text: pinned host FASTQ bytes -> cuclark_classify_text_buffer -> pinned host CSV bytes ----
    # (device-side indexing, 2-bit packing, classification and CSV formatting; the call the CLI makes)
    e2e_text = None
    if not args.no_e2e:
        rec = 16 + 2 * READ_LEN
        d_text = torch.empty(n * rec, dtype=torch.uint8, device="cuda")
        g.synth_fastq_device(READ_SEED, DB_SEED, T, GENOME_LEN, rank * n, n, READ_LEN, args.pct_random, args.sub_per_10k,
                             d_text.data_ptr(), stream)
        g.stats(sync_stream=stream, sync=True)
        h_text = torch.empty(n * rec, dtype=torch.uint8, pin_memory=True)
        h_text.copy_(d_text)
        del d_text
        cap = n * 72 + 4096
        h_csv = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        names = [f"T{t:05d}" for t in range(T)]

        def text_pass():
            return g.classify_text_buffer(h_text.data_ptr(), n * rec, h_csv.data_ptr(), cap, names=names,
                                          chunk_bytes=args.chunk_mb << 20, n_slots=args.slots)

        text_pass()
        barrier()
        t0 = time.perf_counter()
        reps = max(2, min(args.steps, 4))
        for _ in range(reps):
            csv_len, tst = text_pass()
        barrier()
        text_s = (time.perf_counter() - t0) / reps
        # parity at full size: every CSV line against the device-resident path's result rows
        csv = bytes(h_csv[:csv_len].numpy())
        lines = csv.split(b"\n")
        ref5 = d_final.view(n, 5).cpu().numpy().view(np.uint16)
        ok = len(lines) == n + 2 and tst["n_reads"] == n and tst["lookups"] == lookups
        step_chk = max(1, n // 200_000)
        for i in range(0, n, step_chk):
            f = lines[1 + i].split(b",")
            s5 = ref5[i]
            exp_name1 = names[s5[1] - 1].encode() if s5[1] else b"NA"
            exp_name2 = names[s5[3] - 1].encode() if s5[3] else b"NA"
            if not (f[0] == b"r%09d" % ((rank * n + i) % 1_000_000_000) and int(f[1]) == READ_LEN and f[3] == exp_name1
                    and int(f[4]) == s5[2] and f[5] == exp_name2 and int(f[6]) == s5[4]
                    and f[2] == (b"%g" % (s5[0] / (READ_LEN - K + 1.0)))):
                ok = False
                break
        # --extended (one hit-count column per target, src/CuCLARK_hh.hh:2014-2031): sparse rows travel D2H as text;
        # measured on a prefix of the reads (the CSV grows to ~2.9 kB per read with 1,430 targets)
        e2e_ext = None
        n_ext = min(n, args.extended_reads)
        if n_ext:
            cap_x = n_ext * (2 * T + 96) + (8 * T + 4096)
            h_csvx = torch.empty(cap_x, dtype=torch.uint8, pin_memory=True)

            def ext_pass():
                return g.classify_text_buffer(h_text.data_ptr(), n_ext * rec, h_csvx.data_ptr(), cap_x, names=names, extended=True,
                                              chunk_bytes=16 << 20, n_slots=args.slots)
            ext_pass()
            t0 = time.perf_counter()
            x_len, x_st = ext_pass()
            x_s = time.perf_counter() - t0
            xl = bytes(h_csvx[:x_len].numpy()).split(b"\n")
            okx = len(xl) == n_ext + 2 and xl[0].count(b",") == T + 7
            for i in range(0, n_ext, max(1, n_ext // 2000)):
                fcol = xl[1 + i].split(b",")
                s5 = ref5[i]
                cols = np.array(fcol[1:1 + T], dtype=np.int64)
                okx = okx and len(fcol) == T + 8 and int(cols.sum()) == s5[0] and (s5[1] == 0 or cols[s5[1] - 1] == s5[2])
            e2e_ext = {"reads": n_ext, "s": x_s, "csv_bytes": int(x_len), "ok": bool(okx)}
            del h_csvx
        # the platform's ceiling for this path: the same pinned bytes moved by plain cudaMemcpyAsync, all ranks at once
        d_sink = torch.empty(n * rec, dtype=torch.uint8, device="cuda")
        with torch.cuda.stream(ts):
            d_sink.copy_(h_text, non_blocking=True)
        barrier()
        ce = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ce[0].record(ts)
        with torch.cuda.stream(ts):
            for _ in range(3):
                d_sink.copy_(h_text, non_blocking=True)
        ce[1].record(ts)
        barrier()
        h2d_s = ce[0].elapsed_time(ce[1]) * 1e-3 / 3
        del d_sink
        e2e_text = {"s": text_s, "h2d": int(n * rec), "d2h": int(csv_len), "ok": bool(ok), "chunks": tst["n_chunks"],
                    "h2d_s": h2d_s, "ext": e2e_ext}
        del h_text, h_csv
    clocks = sampler.stop()

    # ---- roofline denominator measured on the same table ---------------------------
    gather_ms = min(g.gather_bench(1 << 28, 32, ilp, 3) for ilp in (1, 4, 8))
    random_gbs = (1 << 28) * 32 / gather_ms / 1e6

    # ---- size-independent property check at full size ------------------------------
    f = d_final.view(n, 5).cpu().numpy().view(np.uint16)
    classified = float((f[:, 1] > 0).mean())
    full_hits = float((f[:, 2] == READ_LEN - K + 1).mean())
    import hashlib
    final_sha1 = hashlib.sha1(np.ascontiguousarray(f).tobytes()).hexdigest()   # same for every table layout

    # ---- every result row against the generator's ground truth (numpy twin, cuclark_b200/synth.py) ----
    # A sampled read must come back as [clean, target+1, clean, 0, 0] (clean = k-mer windows without a substituted
    # base = 120 for clean reads), a random read as zeros. The only legitimate exceptions are reads touching one of the
    # few k-mers common to two random genomes (RemoveCommon took them out of the table: ~n_entries^2 / 4^k of them) and
    # random k-mers hitting the table by chance (~lookups x n_entries / 4^k): a handful per 10 M reads, bounded below.
    from cuclark_b200 import synth
    n_chk = n if not args.sub_per_10k else min(n, 1_000_000)      # the substitution twin costs ~5 s per million reads
    tgt, clean = synth.read_truth(READ_SEED, rank * n, n_chk, READ_LEN, T, args.pct_random, args.sub_per_10k, K)
    exp = np.zeros((n_chk, 5), np.uint16)
    smp = (tgt >= 0) & (clean > 0)
    exp[smp, 0] = clean[smp]; exp[smp, 1] = tgt[smp] + 1; exp[smp, 2] = clean[smp]
    bad = np.nonzero((f[:n_chk] != exp).any(axis=1))[0]
    n_entries_all = float(T) * (GENOME_LEN - K + 1)
    gt_bound = int(10 + n_chk * READ_LEN * 4 * n_entries_all / 4.0 ** K + n_chk * 120 * 2 * n_entries_all / 4.0 ** K)
    gt = {"rows_equal_ground_truth": bool(bad.size <= gt_bound), "ground_truth_rows_checked": int(n_chk),
          "ground_truth_mismatches": int(bad.size), "ground_truth_mismatch_bound": gt_bound,
          "ground_truth_mismatch_examples": [{"read": int(rank * n + i), "got": f[i].tolist(), "expected": exp[i].tolist()}
                                             for i in bad[:4]]}

    # ---- reduce over ranks --------------------------------------------------------
    t = torch.tensor([total_ms, e2e["s"] if e2e else 0.0, e2e_text["s"] if e2e_text else 0.0,
                      e2e_text["h2d_s"] if e2e_text else 0.0], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(lookups), float(n)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    total_ms_max, e2e_s_max, text_s_max, h2d_s_max = t.tolist()
    lookups_all, reads_all = cnt.tolist()

    # ---- the same reads through the table-partitioned path (N > 1) -------------------
    table_records = {}
    if world > 1 and not table_mode and not args.no_table_mode:
        g.close()
        torch.cuda.empty_cache()
        # config 2's table as hashed shards (any table size; the faster of the two) and as LOCAL shards (tables of up to
        # 2^30 lines: measured slower here, the probe is bound by the pulls over NVLink and LOCAL entries are 12 B, not 8)
        table_records["table_mode"] = run_table_mode(args, rank, world, local, T, d_ptr, d_cont, n, ts, ref_final=f)
        table_records["table_mode_local_shards"] = run_table_mode(args, rank, world, local, T, d_ptr, d_cont, n, ts, ref_final=f,
                                                                  steps=max(2, args.steps // 2), layout=3)
        # BASELINE configs[4]: a table that exceeds one GPU (11,000 targets = 44 G entries, ~600 GB over the shards)
        free_b = torch.cuda.mem_get_info()[0]
        c5_targets = args.c5_targets
        need_per_gpu = c5_targets * (GENOME_LEN - K + 1) / 2.6 * 32 * 1.06 / world + 30e9
        if c5_targets and need_per_gpu < free_b:
            n5 = min(n, 5_000_000)
            gen = CuClarkDB(K, c5_targets, htsize=HTSIZE_FULL, device=local)          # a handle without a table: generator only
            gen.synth_reads_device(READ_SEED, DB_SEED, c5_targets, GENOME_LEN, rank * n5, n5, READ_LEN, args.pct_random,
                                   args.sub_per_10k, d_ptr.data_ptr(), d_cont.data_ptr(), stream)
            gen.stats(sync_stream=stream, sync=True)
            gen.close()
            table_records["table_mode_config5"] = run_table_mode(args, rank, world, local, c5_targets, d_ptr, d_cont, n5, ts,
                                                                 steps=max(2, args.steps // 2))

    if rank == 0:
        peak, peak_src = measured_peaks()
        kernel_ms = float(np.mean(step_ms))
        traffic, traffic_src = ncu_traffic("config2" if st["layout"] == 1 else "config2_layout%d" % st["layout"], lookups)
        achieved = lookups * BYTES_PER_LOOKUP / (kernel_ms * 1e-3) / 1e9
        value = lookups_all * args.steps / (total_ms_max * 1e-3)
        line = {
            "metric": "kmer_lookups_per_s", "value": value, "unit": "lookups/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "reads_per_s": reads_all * args.steps / (total_ms_max * 1e-3),
            "config": workload_config(args, world),
            "table": {"entries": st["n_entries"], "buckets": st["n_buckets"], "bytes": st["table_bytes"],
                      "layout": {1: "narrow", 2: "wide", 3: "local"}.get(st["layout"], str(st["layout"])), "overflow_entries": st["n_spilled"],
                      "overflowed_buckets": st["n_spill_buckets"], "build_s": build_s},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": lookups * BYTES_PER_LOOKUP, "peak_source": peak_src,
                         "kernel": "k_classify<%s,false>" % {1: "NARROW", 2: "WIDE", 3: "LOCAL"}.get(st["layout"], "?"), "kernel_ms": kernel_ms,
                         "bytes_per_lookup": BYTES_PER_LOOKUP, "lookups_per_launch": lookups,
                         "random_access_peak": random_gbs, "frac_random_access": achieved / random_gbs,
                         "random_access_peak_source": "measured live: 2^28 random 32 B sector loads over the same table"},
            "clocks": clocks,
            "numa": numa,
            "gpu_launches": launches,         # this rank's kernels in the timed region, counted by the library at its launch sites
            "parity_properties": {"classified_frac": classified,
                                  "expected_classified_frac": (1 - args.pct_random / 100) if not args.sub_per_10k else None,
                                  "reads_with_all_kmers_hit_frac": full_hits,
                                  "final_rows_sha1_rank0": final_sha1, **gt},
        }
        line.update(table_records)
        if e2e_text:
            line["e2e"] = {"value": lookups_all / text_s_max, "unit": "lookups/s", "reads_per_s": reads_all / text_s_max,
                           "h2d_bytes_per_step": e2e_text["h2d"], "d2h_bytes_per_step": e2e_text["d2h"],
                           "ms_per_step": text_s_max * 1e3, "results_equal_device_path": e2e_text["ok"],
                           "chunks_per_step": e2e_text["chunks"],
                           # plain cudaMemcpyAsync of the same pinned input, all ranks at once (max over ranks): what the
                           # host-to-device link gives this box; the e2e call moves the same bytes plus the CSV back
                           "h2d_ceiling_gbs_aggregate": world * e2e_text["h2d"] / h2d_s_max / 1e9,
                           "h2d_achieved_gbs_aggregate": world * e2e_text["h2d"] / text_s_max / 1e9,
                           "frac_of_h2d_ceiling": h2d_s_max / text_s_max,
                           "path": "pinned host FASTQ text -> cuclark_classify_text_buffer (H2D, device index + 2-bit pack + "
                                   "classify + CSV format, D2H) -> pinned host CSV text; wall clock of the call"}
        if e2e_text and e2e_text.get("ext"):
            x = e2e_text["ext"]
            line["e2e_extended"] = {"reads_per_s_rank0": x["reads"] / x["s"], "lookups_per_s_rank0": x["reads"] * (READ_LEN - K + 1) / x["s"],
                                    "reads": x["reads"], "csv_bytes": x["csv_bytes"], "columns": T + 8,
                                    "results_equal_device_path": x["ok"],
                                    "path": "as e2e with --extended: per-target hit columns formatted on the device from the sparse rows"}
        if e2e:
            line["e2e_packed"] = {"value": lookups_all / e2e_s_max, "unit": "lookups/s", "reads_per_s": reads_all / e2e_s_max,
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "ms_per_step": e2e_s_max * 1e3, "results_equal_device_path": e2e["same_as_device_path"],
                           "path": "pinned host packed reads -> cuclark_batch_query (H2D, kernels, D2H) -> host results"}
        if world == 1 and not args.no_cpu_baseline:
            # the host legs run with the GPU handed back (the reference binary of the B1 leg sizes its batches by the free
            # device memory and needs the device to itself)
            g.close()
            d_ptr = d_cont = d_final = None
            torch.cuda.empty_cache()
            threads = os.cpu_count() or 1
            orc, db, data, cpu_build_s = cpu_sample(args, threads)
            dt, lk, nr = cpu_step(orc, db, data, args.cpu_targets, threads)
            line["cpu_baseline"] = {
                "value": lk / dt, "unit": "lookups/s", "cores": threads, "kind": "port",
                "reads_per_s": nr / dt, "db_build_s": cpu_build_s,
                "sample": f"{args.cpu_reads} x {READ_LEN} bp reads against a {args.cpu_targets} x 4 Mbp "
                          f"({db.size / 1e6:.0f} M 31-mer) database in the reference's table layout; "
                          f"index+pack+extract+lookup+histogram+top-2 (oracle port, OpenMP)"}
            if not args.no_ref_lookup and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_lookup_full.so")):
                try:
                    line["cpu_baseline"]["reference_lookup"] = reference_lookup_leg(orc, db, data, threads)
                except Exception as e:      # oracle/_ref is built where /root/reference exists
                    line["cpu_baseline"]["reference_lookup"] = {"unavailable": repr(e)}
            try:
                line["cpu_baseline"]["reference_host_path"] = reference_host_path_leg(threads)
            except Exception as e:
                line["cpu_baseline"]["reference_host_path"] = {"unavailable": repr(e)}
        print(json.dumps(line), flush=True)
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a, int(os.environ.get("RANK", 0)))
    else:
        run_b200(a)
