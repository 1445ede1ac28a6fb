"""Random 32-byte sector gather rate against the footprint of the table (B200, one GPU).

Builds the bench database (NARROW layout) at several bucket loads, i.e. table sizes, and times
cuclark_gather_bench over each: shows where the random-access ceiling falls off with the footprint.
    python tools/footprint_probe.py [targets]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuclark_b200.api import CuClarkDB, HTSIZE_FULL  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1430
out = []
for load in (4.0, 3.2, 2.6, 2.2, 1.9, 1.6):
    with CuClarkDB(31, T, htsize=HTSIZE_FULL, layout=1, bucket_load=load) as g:
        g.build_synthetic(1, T, 4_000_000, 0)
        st = g.stats()
        ms = min(g.gather_bench(1 << 27, 32, ilp, 3) for ilp in (4, 8))
        row = {"bucket_load": load, "table_gb": st["table_bytes"] / 1e9, "home_gb": st["n_local_buckets"] * 32 / 1e9,
               "gsectors_per_s": (1 << 27) / ms / 1e6}
        print(json.dumps(row), flush=True)
        out.append(row)
