#!/usr/bin/env python
"""GPU probe: random-sector gather ceiling vs table footprint, table build time, first classify rates.
Writes gpurun_out/probe.json. Not a bench (bench.py is); this informs design choices."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cuclark_b200.api import CuClarkDB, HTSIZE_FULL  # noqa: E402

out = {"host": {"cpus": os.cpu_count()}}
try:
    out["host"]["mem_gb"] = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 1e9
    out["nvidia_smi"] = subprocess.run(["nvidia-smi", "--query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem",
                                        "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
except Exception as e:  # pragma: no cover
    out["host"]["err"] = str(e)

targets_list = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "40,350,1430").split(",")]
n_reads, L, k = 4_000_000, 150, 31
results = []
for T in targets_list:
    r = {"targets": T}
    g = CuClarkDB(k, max(T, 1), htsize=HTSIZE_FULL)
    t0 = time.time()
    g.build_synthetic(1, T, 4_000_000, 0)
    r["build_s"] = time.time() - t0
    st = g.stats()
    r.update({kk: st[kk] for kk in ("n_entries", "n_buckets", "table_bytes", "n_spilled", "n_spill_buckets", "layout")})
    # gather ceiling
    probes = 1 << 28
    for B in ((32, 64, 128) if os.environ.get("PROBE_ALL") else (32,)):
        for ilp in (1, 4, 8):
            ms = g.gather_bench(probes, B, ilp, 3)
            r[f"gather_{B}B_ilp{ilp}_Gps"] = probes / ms / 1e6
            r[f"gather_{B}B_ilp{ilp}_GBs"] = probes * B / ms / 1e6
    # classify
    per = 1 + (L + 7) // 8
    d_ptr = torch.zeros(n_reads + 1, dtype=torch.int32, device="cuda")
    d_cont = torch.zeros(n_reads * per, dtype=torch.int16, device="cuda")
    d_final = torch.zeros(n_reads * 5, dtype=torch.int16, device="cuda")
    d_rows = torch.zeros(n_reads * 32, dtype=torch.int16, device="cuda")
    for pct_random in (10, 100):
        g.synth_reads_device(2, 1, T, 4_000_000, 0, n_reads, L, pct_random, 0, d_ptr.data_ptr(), d_cont.data_ptr())
        g.stats(sync=True)
        ts = torch.cuda.Stream()          # a real handle: 0/NULL would mean "library stream"
        torch.cuda.synchronize()
        stream = ts.cuda_stream
        for rows in (0, 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for it in range(4):
                if it == 1:
                    e0.record(ts)
                g.classify_device(d_ptr.data_ptr(), d_cont.data_ptr(), n_reads, d_final.data_ptr(),
                                  d_rows.data_ptr() if rows else 0, stream)
            e1.record(ts)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            s2 = g.stats(sync_stream=stream, sync=True)
            r[f"classify_rand{pct_random}_rows{rows}_ms"] = ms
            r[f"classify_rand{pct_random}_rows{rows}_Glookups_s"] = s2["lookups"] / ms / 1e6
            r[f"classify_rand{pct_random}_rows{rows}_dense"] = s2["dense_reads"]
        f = d_final.view(n_reads, 5).cpu().numpy().view("uint16")
        r[f"rand{pct_random}_classified_frac"] = float((f[:, 1] > 0).mean())
        r[f"rand{pct_random}_mean_h1"] = float(f[:, 2].mean())
    del d_ptr, d_cont, d_final, d_rows
    g.close()
    torch.cuda.empty_cache()
    results.append(r)
    print(json.dumps(r), flush=True)
out["results"] = results
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
    json.dump(out, f, indent=1)
