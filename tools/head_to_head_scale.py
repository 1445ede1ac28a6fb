#!/usr/bin/env python
"""Drop-in head to head at BASELINE scale on one B200 box: the UNMODIFIED reference GPU binary (oracle/_ref/cuCLARK,
built by oracle/Makefile from /root/reference/src) against cuclark_b200/bin/cuCLARK — same command line, same
target FASTA files, same database FILES, same FASTQ reads; CSVs compared byte for byte.

    python tools/head_to_head_scale.py --targets 1430 --reads 10000000 [--paired] > gpurun_out/h2h.json

Steps (BASELINE.json configs[1] and, with --paired, configs[2]):
  1. --targets seeded genomes of 4 Mbp are written as FASTA target files (the numpy twin of the bench's generator);
  2. OUR executable finds no database and builds .sz/.ky/.lb on the GPU (cuclark_build_database; the reference's host
     builder needs ~146 GB of RAM and hours at this scale, README.md:93) — then classifies;
  3. the reference executable runs on the same folder: it loads the same files (32-bit bucket pointers force it into
     >= 2 swap cycles above 2^32 entries), indexes and packs on the host and classifies;
  4. both CSVs are compared byte for byte; assignment time (the executables' own "Assignment time" line: reads file
     -> last CSV line, database load excluded) and process wall time are recorded.
The folder lives in --dir (default /dev/shm); before anything is written the free space and the free RAM are checked
and the target count is scaled down (and reported) if the box cannot hold the run.
"""
import argparse
import json
import multiprocessing as mp
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
K, G, L = 31, 4_000_000, 150
HTSIZE_FULL = 1610612741
DB_SEED, READ_SEED = 1, 2


def write_target(job):
    folder, t = job
    from cuclark_b200 import synth
    asc = np.frombuffer(b"ACGT", np.uint8)
    p = os.path.join(folder, "tg", f"T{t:05d}.fa")
    with open(p, "wb") as f:
        f.write(b">T%05d\n" % t)
        f.write(asc[synth.genome_codes(DB_SEED, t, 0, G)].tobytes())
        f.write(b"\n")
    return p


def mem_available_gb():
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable:"):
            return int(line.split()[1]) / 1e6
    return 0.0


def run_exe(exe, folder, args, timeout, env=None):
    t0 = time.time()
    try:
        p = subprocess.run([exe, *args], cwd=folder, capture_output=True, text=True, timeout=timeout, env=env)
    except subprocess.TimeoutExpired:
        return {"error": f"timeout after {timeout} s", "wall_s": time.time() - t0}
    wall = time.time() - t0
    out = {"wall_s": wall, "rc": p.returncode}
    m = re.search(r"Assignment time: ([0-9.eE+-]+) s\. Speed: (\d+) objects/min\. \((\d+) objects\)", p.stdout)
    if m:
        out["assignment_s"] = float(m.group(1))
        out["objects_line"] = int(m.group(3))
    else:
        out["error"] = (p.stdout + p.stderr)[-1500:]
    out["stderr_tail"] = p.stderr.strip().splitlines()[-12:]
    out["library_timing"] = [l for l in p.stderr.splitlines() if l.startswith("[cuclark timing]")]
    return out


def same_file(a, b):
    if not (os.path.exists(a) and os.path.exists(b)):
        return None
    if os.path.getsize(a) != os.path.getsize(b):
        return False
    with open(a, "rb") as fa, open(b, "rb") as fb:
        while True:
            x, y = fa.read(1 << 24), fb.read(1 << 24)
            if x != y:
                return False
            if not x:
                return True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--targets", type=int, default=1430)
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--paired", action="store_true", help="also BASELINE configs[2]: 2 x 150 bp mates with 1%% substitutions")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--dir", default="/dev/shm" if os.path.isdir("/dev/shm") else None)
    ap.add_argument("--ref-timeout", type=int, default=1500)
    ap.add_argument("--skip-reference", action="store_true")
    a = ap.parse_args()

    import torch
    from cuclark_b200 import build
    from cuclark_b200.api import CuClarkDB
    build.build_all()
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "cuCLARK")
    our_exe = os.path.join(ROOT, "cuclark_b200", "bin", "cuCLARK")

    # ---- does the box hold the run? files: targets + database + reads + CSVs; RAM: the reference keeps the database
    # in pinned host memory on top (README.md:95: ~40 GB at bacterial scale)
    def need_gb(T):
        entries = T * (G - K + 1)
        db = (HTSIZE_FULL + entries * 6) / 1e9
        files = T * G / 1e9 + db + a.reads * 316 / 1e9 * (3 if a.paired else 1) + a.reads * 80 / 1e9 * (4 if a.paired else 2)
        return files, db
    free_disk = shutil.disk_usage(a.dir or tempfile.gettempdir()).free / 1e9
    ram = mem_available_gb()
    T = a.targets
    while T > 8:
        files, db = need_gb(T)
        in_ram = files if (a.dir or "").startswith("/dev/shm") else 0.0
        if files * 1.05 < free_disk and in_ram + db * 1.4 + 20 < ram:
            break
        T = int(T * 0.8)
    out = {"requested_targets": a.targets, "targets": T, "reads": a.reads, "k": K, "free_disk_gb": free_disk, "mem_available_gb": ram,
           "threads": a.threads, "dir": a.dir, "gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count()}
    folder = tempfile.mkdtemp(prefix="h2h_scale_", dir=a.dir)
    try:
        os.makedirs(os.path.join(folder, "tg")); os.makedirs(os.path.join(folder, "db"))
        t0 = time.time()
        with mp.Pool(min(32, a.threads)) as pool:
            paths = pool.map(write_target, [(folder, t) for t in range(T)], chunksize=4)
        with open(os.path.join(folder, "targets.txt"), "w") as tf:
            for t, p in enumerate(paths):
                tf.write(f"{p} T{t:05d}\n")
        out["targets_written_s"] = time.time() - t0

        # ---- reads from the device generator (the bench's), written as FASTQ files
        t0 = time.time()
        rec = 16 + 2 * L
        step = 2_000_000
        with CuClarkDB(K, T) as gen:
            d = torch.empty(step * rec, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            files = [("reads.fq", 0, 0)] + ([("mates_1.fq", 1, 100), ("mates_2.fq", 2, 100)] if a.paired else [])
            for name, mate, subs in files:
                with open(os.path.join(folder, name), "wb") as f:
                    for lo in range(0, a.reads, step):
                        m = min(step, a.reads - lo)
                        if mate:
                            gen.synth_fastq_pair_device(READ_SEED, DB_SEED, T, G, lo, m, L, 10, subs, mate, d.data_ptr())
                        else:
                            gen.synth_fastq_device(READ_SEED, DB_SEED, T, G, lo, m, L, 10, subs, d.data_ptr())
                        gen.stats(sync=True)
                        f.write(d[:m * rec].cpu().numpy().tobytes())
            del d
        torch.cuda.empty_cache()
        out["reads_written_s"] = time.time() - t0

        common = ["-k", str(K), "-T", "targets.txt", "-D", "db/", "-n", str(a.threads), "-d", "1"]
        env = dict(os.environ, CUCLARK_TIMING="1")
        runs = {}
        # ---- ours: first run builds the database on the GPU, second run finds it
        runs["b200_build_and_classify"] = run_exe(our_exe, folder, common + ["-O", "reads.fq", "-R", "out_b200"], 1800, env)
        dbfiles = sorted(f for f in os.listdir(os.path.join(folder, "db")) if f.endswith((".sz", ".ky", ".lb")))
        out["db_files"] = {f: os.path.getsize(os.path.join(folder, "db", f)) for f in dbfiles}
        if len(dbfiles) == 3:
            runs["b200_classify"] = run_exe(our_exe, folder, common + ["-O", "reads.fq", "-R", "out_b200"], 1800, env)
            if a.paired:
                runs["b200_paired"] = run_exe(our_exe, folder, common + ["-P", "mates_1.fq", "mates_2.fq", "-R", "outp_b200"], 1800, env)
            if not a.skip_reference and os.path.exists(ref_exe):
                runs["reference_classify"] = run_exe(ref_exe, folder, common + ["-b", str(a.threads), "-O", "reads.fq", "-R", "out_ref"],
                                                     a.ref_timeout)
                if a.paired:
                    runs["reference_paired"] = run_exe(ref_exe, folder, common + ["-b", str(a.threads), "-P", "mates_1.fq", "mates_2.fq",
                                                                                  "-R", "outp_ref"], a.ref_timeout)
        out["runs"] = runs
        out["csv_identical_single_end"] = same_file(os.path.join(folder, "out_b200.csv"), os.path.join(folder, "out_ref.csv"))
        if a.paired:
            out["csv_identical_paired"] = same_file(os.path.join(folder, "outp_b200.csv"), os.path.join(folder, "outp_ref.csv"))
        for nm in ("out_b200.csv", "out_ref.csv", "outp_b200.csv", "outp_ref.csv"):
            p = os.path.join(folder, nm)
            if os.path.exists(p):
                out.setdefault("csv_bytes", {})[nm] = os.path.getsize(p)
        lookups = a.reads * (L - K + 1)

        def rate(r, mult=1):
            return mult * lookups / r["assignment_s"] if r and "assignment_s" in r else None
        out["lookups_per_s"] = {k: rate(v, 2 if "paired" in k else 1) for k, v in runs.items()}
        b, r = runs.get("b200_classify"), runs.get("reference_classify")
        if b and r and "assignment_s" in b and "assignment_s" in r:
            out["speedup_assignment_single_end"] = r["assignment_s"] / b["assignment_s"]
            out["speedup_wall_single_end"] = r["wall_s"] / b["wall_s"]
        b, r = runs.get("b200_paired"), runs.get("reference_paired")
        if b and r and "assignment_s" in b and "assignment_s" in r:
            out["speedup_assignment_paired"] = r["assignment_s"] / b["assignment_s"]
            out["speedup_wall_paired"] = r["wall_s"] / b["wall_s"]
        print(json.dumps(out), flush=True)
    finally:
        shutil.rmtree(folder, ignore_errors=True)


if __name__ == "__main__":
    main()
