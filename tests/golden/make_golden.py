#!/usr/bin/env python
"""Generate the golden fixtures from the UNMODIFIED reference binaries.

The reference ships no tests or golden vectors, so the pins
are made by running its own code (oracle/_ref/cuCLARK, cuCLARK-l; built by
oracle/Makefile from /root/reference/src) on seeded synthetic inputs:

  --stage db    (CPU, this container)  reference DB build -> sha256 of .sz/.ky/.lb
                -> tests/golden/db_<case>.json
  --stage csv   (GPU box, via gpurun)  full reference run -> result CSV
                -> gpurun_out/golden_csv/<case>.csv.gz (copy into tests/golden/)

Inputs are regenerated from (seed, shape) by cuclark_b200.synth, so only digests
and small CSVs are committed. Cases are defined in CASES below.
"""
from __future__ import annotations

import argparse
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cuclark_b200 import synth  # noqa: E402
from oracle import dbtools  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(ROOT, "oracle", "_ref")

# name -> dict(light, k, n_targets, genome_len, seed, gap, shared, reads...)
CASES = {
    # small light DB with a segment shared between targets 0 and 1 (RemoveCommon)
    "light_small": dict(light=True, k=27, n_targets=8, genome_len=200_000, seed=11, gap=4, shared=5400,
                        n_reads=20_000, read_len=100, read_seed=21, pct_random=2, pct_n=3, fmt="fasta"),
    # BASELINE.json configs[0]: CuCLARK-l k=27, 20 x 1 Mbp, 100k x 100 bp
    "light_c1": dict(light=True, k=27, n_targets=20, genome_len=1_000_000, seed=1, gap=4, shared=0,
                     n_reads=100_000, read_len=100, read_seed=1, pct_random=1, pct_n=1, fmt="fasta"),
    # full variant, k=31 (reference needs ~26 GB RAM and ~90 s for this)
    "full_small": dict(light=False, k=31, n_targets=4, genome_len=40_000, seed=12, gap=0, shared=3000,
                       n_reads=20_000, read_len=150, read_seed=22, pct_random=10, pct_n=2, fmt="fastq"),
}


def target_codes(case: dict, t: int) -> np.ndarray:
    codes = synth.genome_codes(case["seed"], t, 0, case["genome_len"])
    if case["shared"] and t == 1:
        # copy a stretch of target 0 into target 1: those k-mers are common to both
        src = synth.genome_codes(case["seed"], 0, 10_800, case["shared"])
        codes = codes.copy()
        codes[21_600:21_600 + case["shared"]] = src
    return codes


def target_name(t: int) -> str:
    return f"T{t:05d}"


def write_inputs(case: dict, folder: str):
    os.makedirs(os.path.join(folder, "tg"), exist_ok=True)
    os.makedirs(os.path.join(folder, "db"), exist_ok=True)
    asc = np.frombuffer(b"ACGT", np.uint8)
    with open(os.path.join(folder, "targets.txt"), "w") as tf:
        for t in range(case["n_targets"]):
            p = os.path.join(folder, "tg", f"{target_name(t)}.fa")
            synth.write_fasta(p, target_name(t), asc[target_codes(case, t)].tobytes())
            tf.write(f"{p} {target_name(t)}\n")
    reads = make_reads(case)
    path = os.path.join(folder, "reads.fa" if case["fmt"] == "fasta" else "reads.fq")
    with open(path, "wb") as f:
        f.write(reads)
    return path


def make_reads(case: dict) -> bytes:
    codes, *_ = synth.read_codes(case["read_seed"], case["n_reads"], case["read_len"], case["n_targets"],
                                 case["genome_len"], case["seed"], pct_random=case["pct_random"], pct_n=case["pct_n"])
    return synth.reads_fasta(codes) if case["fmt"] == "fasta" else synth.reads_fastq(codes)


def run_reference(case: dict, folder: str, reads_path: str, threads: int = 4):
    exe = os.path.join(REF, "cuCLARK-l" if case["light"] else "cuCLARK")
    cmd = [exe, "-T", "targets.txt", "-D", "db/", "-O", os.path.basename(reads_path), "-R", "out", "-n", str(threads)]
    if not case["light"]:
        cmd += ["-k", str(case["k"])]
    elif case["gap"] != 4:
        cmd += ["-g", str(case["gap"])]
    p = subprocess.run(cmd, cwd=folder, capture_output=True, text=True)
    return p


def stage_db(names):
    for name in names:
        case = CASES[name]
        folder = tempfile.mkdtemp(prefix=f"golden_{name}_")
        try:
            reads = write_inputs(case, folder)
            p = run_reference(case, folder, reads)
            htsize = 57777779 if case["light"] else 1610612741
            base = dbtools.db_name(os.path.join(folder, "db"), case["k"], case["n_targets"], htsize, 0, case["gap"])
            files = {ext: base + ext for ext in (".sz", ".ky", ".lb")}
            if not all(os.path.exists(f) for f in files.values()):
                raise RuntimeError(f"reference did not write the DB:\n{p.stdout}\n{p.stderr}")
            out = dict(case=case, db_basename=os.path.basename(base),
                       sha256={ext: dbtools.sha256_file(f) for ext, f in files.items()},
                       bytes={ext: os.path.getsize(f) for ext, f in files.items()},
                       reference_stderr_tail=p.stderr[-600:])
            with open(os.path.join(HERE, f"db_{name}.json"), "w") as f:
                json.dump(out, f, indent=1)
            print(name, out["bytes"])
        finally:
            shutil.rmtree(folder, ignore_errors=True)


def stage_csv(names, outdir):
    os.makedirs(outdir, exist_ok=True)
    for name in names:
        case = CASES[name]
        folder = tempfile.mkdtemp(prefix=f"golden_{name}_")
        try:
            reads = write_inputs(case, folder)
            p = run_reference(case, folder, reads)
            csv = os.path.join(folder, "out.csv")
            with open(os.path.join(outdir, f"{name}.log"), "w") as f:
                f.write(p.stdout + "\n----\n" + p.stderr)
            if not os.path.exists(csv):
                print(name, "reference produced no CSV; see log")
                continue
            with open(csv, "rb") as f, open(os.path.join(outdir, f"{name}.csv.gz"), "wb") as raw, \
                    gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as g:
                shutil.copyfileobj(f, g)
            print(name, "csv bytes", os.path.getsize(csv))
        finally:
            shutil.rmtree(folder, ignore_errors=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", choices=["db", "csv"], required=True)
    ap.add_argument("--cases", nargs="*", default=list(CASES))
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden_csv"))
    a = ap.parse_args()
    if a.stage == "db":
        stage_db(a.cases)
    else:
        stage_csv(a.cases, a.out)
