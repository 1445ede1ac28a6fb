#!/usr/bin/env python
"""Refresh profiles/traffic.json from an `ncu --set full` capture of the classify kernel.

    python tools/refresh_traffic.py gpurun_out/<capture>.ncu-rep --key config2_layout3 --lookups 360000000

Reads dram__bytes_read.sum + dram__bytes_write.sum of the first `k_classify` launch in the report (ncu must be on
PATH; it reads reports without a GPU), divides by the lookups of that launch and stores the figure together with the
SHA-1 of the kernel's sources. bench.py reports `roofline.traffic` only while that SHA-1 still matches the sources it
runs: a changed kernel with a stale capture yields null, not an old number."""
import argparse
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_SOURCES = ["classify.cu", "common.cuh", "hits.cuh", "kmerwin.cuh"]


def kernel_sha1() -> str:
    h = hashlib.sha1()
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(ROOT, "cuclark_b200", "csrc", f), "rb").read())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--key", required=True, help="workload key in traffic.json (config2 = NARROW, config2_layout3 = LOCAL)")
    ap.add_argument("--lookups", type=int, required=True, help="lookups of the captured launch (bench line: lookups_per_launch)")
    a = ap.parse_args()
    raw = subprocess.check_output(["ncu", "-i", a.report, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    row = next(r for r in rows[2:] if "k_classify<" in r[hdr.index("Kernel Name")] and "dense" not in r[hdr.index("Kernel Name")])

    def metric(name):
        i = hdr.index(name)
        v = float(row[i])
        u = units[i].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

    total = metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum")
    path = os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[a.key] = {"dram_bytes_per_lookup": total / a.lookups, "dram_bytes": total, "lookups": a.lookups,
                   "kernel_ms": metric("gpu__time_duration.sum") if "gpu__time_duration.sum" in hdr else None,
                   "report": os.path.basename(a.report), "kernel_source_sha1": kernel_sha1(),
                   "kernel": row[hdr.index("Kernel Name")]}
    json.dump(data, open(path, "w"), indent=1)
    print(json.dumps(data[a.key], indent=1))


if __name__ == "__main__":
    sys.exit(main())
