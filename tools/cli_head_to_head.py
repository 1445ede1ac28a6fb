#!/usr/bin/env python
"""Drop-in head to head on one B200: the UNMODIFIED reference binary (oracle/_ref/cuCLARK, built by
oracle/Makefile from /root/reference/src) against cuclark_b200/bin/cuCLARK, same command line, same
database files, same FASTQ file; prints one JSON line.

Both executables print " - Assignment time: <s> s. Speed: <n> objects/min." (src/CuCLARK_hh.hh:1938-1945):
the time from the mmap of the reads file to the last CSV line, database load excluded. The wall time of
the whole process (database load included) is recorded next to it. The two CSV files are compared byte
for byte.

The database is a bounded slice of bench.py's workload (same seeded genomes, fewer targets) because the
reference needs the .sz/.ky/.lb FILES, which only a host builder can write here (the oracle port's);
this is test tooling and the only place outside tests/ and bench.py's CPU legs that touches oracle/.
"""
import argparse
import json
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

K, GENOME_LEN, READ_LEN = 31, 4_000_000, 150


def run(exe, folder, reads, out, threads, extra=()):
    cmd = [exe, "-k", str(K), "-T", "targets.txt", "-D", "db/", "-O", reads, "-R", out, "-n", str(threads), *extra]
    t0 = time.time()
    p = subprocess.run(cmd, cwd=folder, capture_output=True, text=True)
    wall = time.time() - t0
    m = re.search(r"Assignment time: ([0-9.eE+-]+) s\. Speed: (\d+) objects/min\. \((\d+) objects\)", p.stdout)
    if not m:
        return {"error": (p.stdout + p.stderr)[-800:], "rc": p.returncode, "wall_s": wall}
    r = {"assignment_s": float(m.group(1)), "objects": int(m.group(3)), "wall_s": wall}
    timing = [l for l in p.stderr.splitlines() if l.startswith("[cuclark timing]")]
    if timing:
        r["library_timing"] = timing
    return r


def config1(a):
    """BASELINE.json configs[0] from scratch: cuCLARK-l k=27, 20 x 1 Mbp target FASTA files, 100k x 100 bp
    reads, NO database on disk: each executable builds .sz/.ky/.lb first (the reference on the host,
    serially; ours on the GPU) and then classifies; a second run finds the database."""
    import hashlib
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    case = make_golden.CASES["light_c1"]
    out = {"workload": "BASELINE configs[0]: cuCLARK-l k=27 -g 4, 20 x 1 Mbp targets (FASTA), 100000 x 100 bp reads; "
                       f"database built by each executable; -n {a.threads}"}
    shas, csvs = {}, {}
    for name, exe in (("reference", os.path.join(ROOT, "oracle", "_ref", "cuCLARK-l")),
                      ("b200", os.path.join(ROOT, "cuclark_b200", "bin", "cuCLARK-l"))):
        folder = tempfile.mkdtemp(prefix="h2h_c1_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            reads = make_golden.write_inputs(case, folder)
            cmd = [exe, "-T", "targets.txt", "-D", "db/", "-O", os.path.basename(reads), "-R", "out", "-n", str(a.threads)]
            res = {}
            dbdir = os.path.join(folder, "db")
            for label in ("build_and_classify", "classify_only"):
                walls, assigns = [], []
                for _ in range(a.repeats):
                    if label == "build_and_classify":
                        for f in os.listdir(dbdir):
                            os.remove(os.path.join(dbdir, f))
                    if os.path.exists(os.path.join(folder, "out.csv")):
                        os.remove(os.path.join(folder, "out.csv"))
                    t0 = time.time()
                    p = subprocess.run(cmd, cwd=folder, capture_output=True, text=True)
                    walls.append(round(time.time() - t0, 3))
                    m = re.search(r"Assignment time: ([0-9.eE+-]+) s", p.stdout)
                    if m:
                        assigns.append(float(m.group(1)))
                    if not os.path.exists(os.path.join(folder, "out.csv")):
                        res["error"] = (p.stdout + p.stderr)[-600:]
                        break
                    if p.returncode != 0:      # the reference dies in its teardown on sm_100 (CUERR at CuClarkDB.cu:292) after "Done."
                        res["exit_code"] = p.returncode
                        res["exit_note"] = p.stderr.strip().splitlines()[-1][-120:]
                    timing = [l for l in p.stderr.splitlines() if l.startswith("[cuclark timing]")]
                    if timing:
                        res[label + "_library_timing"] = timing
                res[label + "_wall_s"] = min(walls)
                res[label + "_wall_s_all"] = walls
                if assigns:
                    res[label + "_assignment_s"] = min(assigns)
            shas[name] = {f.rsplit(".", 1)[1]: hashlib.sha256(open(os.path.join(dbdir, f), "rb").read()).hexdigest()
                          for f in sorted(os.listdir(dbdir)) if f.endswith((".sz", ".ky", ".lb"))}
            csvs[name] = open(os.path.join(folder, "out.csv"), "rb").read() if os.path.exists(os.path.join(folder, "out.csv")) else None
            out[name] = res
        finally:
            shutil.rmtree(folder, ignore_errors=True)
    out["db_files_identical"] = bool(shas.get("reference")) and shas.get("reference") == shas.get("b200")
    out["csv_identical"] = csvs.get("reference") is not None and csvs.get("reference") == csvs.get("b200")
    golden = os.path.join(ROOT, "tests", "golden", "light_c1.csv.gz")
    if os.path.exists(golden) and csvs.get("b200") is not None:
        import gzip
        out["csv_equals_committed_golden"] = gzip.open(golden).read() == csvs["b200"]
    r, b = out.get("reference", {}), out.get("b200", {})
    if "error" not in r and "error" not in b:
        out["speedup_from_scratch_wall"] = r["build_and_classify_wall_s"] / b["build_and_classify_wall_s"]
        out["speedup_classify_only_wall"] = r["classify_only_wall_s"] / b["classify_only_wall_s"]
        if "classify_only_assignment_s" in r and "classify_only_assignment_s" in b:
            out["speedup_assignment"] = r["classify_only_assignment_s"] / b["classify_only_assignment_s"]
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config1", action="store_true", help="BASELINE configs[0] from scratch (database build included)")
    ap.add_argument("--targets", type=int, default=8)
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    if a.config1:
        from cuclark_b200 import build
        build.build_all()
        return config1(a)

    from cuclark_b200 import build, synth
    from oracle.binding import HTSIZE_FULL, Oracle
    build.build_all()
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "cuCLARK")
    our_exe = os.path.join(ROOT, "cuclark_b200", "bin", "cuCLARK")
    folder = tempfile.mkdtemp(prefix="h2h_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        os.makedirs(os.path.join(folder, "tg")); os.makedirs(os.path.join(folder, "db"))
        with open(os.path.join(folder, "targets.txt"), "w") as tf:
            for t in range(a.targets):
                p = os.path.join(folder, "tg", f"T{t:05d}.fa")
                open(p, "w").write(f">T{t:05d}\n")
                tf.write(f"{p} T{t:05d}\n")
        base = os.path.join(folder, "db", f"db_central_k{K}_t{a.targets}_s{HTSIZE_FULL}_m0.tsk")
        t0 = time.time()
        orc = Oracle()
        db = orc.db_build_synth(1, a.targets, GENOME_LEN, K, HTSIZE_FULL, 0, a.threads, write_base=base)
        n_entries = db.size
        del db
        t_db = time.time() - t0
        t0 = time.time()
        with open(os.path.join(folder, "reads.fq"), "wb") as f:
            step = 200_000
            for lo in range(0, a.reads, step):
                codes, *_ = synth.read_codes(2, min(step, a.reads - lo), READ_LEN, a.targets, GENOME_LEN, 1,
                                             pct_random=10, first=lo)
                f.write(synth.reads_fastq(codes, first=lo))
        t_reads = time.time() - t0
        lookups = a.reads * (READ_LEN - K + 1)
        out = {"workload": f"cuCLARK k=31, {a.targets} x 4 Mbp targets ({n_entries / 1e6:.1f} M 31-mers, HTSIZE {HTSIZE_FULL}), "
                           f"{a.reads} x {READ_LEN} bp FASTQ reads, 10% random; -n {a.threads}",
               "lookups": lookups, "host_db_build_s": t_db, "reads_gen_s": t_reads}
        for name, exe in (("reference", ref_exe), ("b200", our_exe)):
            if not os.path.exists(exe):
                out[name] = {"error": f"{exe} missing"}
                continue
            runs = [run(exe, folder, "reads.fq", f"out_{name}", a.threads) for _ in range(a.repeats)]
            ok = [r for r in runs if "error" not in r]
            best = min(ok, key=lambda r: r["assignment_s"]) if ok else runs[-1]
            best["runs"] = len(runs)
            if ok:
                best["lookups_per_s"] = lookups / best["assignment_s"]
                best["reads_per_s"] = a.reads / best["assignment_s"]
                best["wall_s_all"] = [round(r["wall_s"], 3) for r in runs]
            out[name] = best
        fa, fb = (os.path.join(folder, f"out_{n}.csv") for n in ("reference", "b200"))
        if os.path.exists(fa) and os.path.exists(fb):
            A, B = open(fa, "rb").read(), open(fb, "rb").read()
            out["csv_bytes"] = len(A)
            out["csv_identical"] = A == B
            if A != B:
                la, lb = A.split(b"\n"), B.split(b"\n")
                diff = [i for i in range(min(len(la), len(lb))) if la[i] != lb[i]]
                out["csv_lines_differing"] = len(diff) + abs(len(la) - len(lb))
                out["first_diff"] = [la[diff[0]].decode(), lb[diff[0]].decode()] if diff else None
        if "assignment_s" in out.get("reference", {}) and "assignment_s" in out.get("b200", {}):
            out["speedup_assignment"] = out["reference"]["assignment_s"] / out["b200"]["assignment_s"]
            out["speedup_wall"] = out["reference"]["wall_s"] / out["b200"]["wall_s"]
        print(json.dumps(out), flush=True)
    finally:
        if not a.keep:
            shutil.rmtree(folder, ignore_errors=True)


if __name__ == "__main__":
    main()
