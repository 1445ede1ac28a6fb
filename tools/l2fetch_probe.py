#!/usr/bin/env python
"""Effect of cudaLimitMaxL2FetchGranularity on random 32 B probes and on the classify kernel."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from cuclark_b200.api import CuClarkDB, HTSIZE_FULL
    T = int(os.environ.get("T", 350)); n = 4_000_000; L = 150; per = 20
    g = CuClarkDB(31, T, htsize=HTSIZE_FULL)
    g.build_synthetic(1, T, 4_000_000, 0)
    r = {"gran": os.environ.get("CUCLARK_L2_FETCH"), "T": T}
    for ilp in (1, 4, 8):
        ms = g.gather_bench(1 << 28, 32, ilp, 3); r[f"gather32_ilp{ilp}_Gps"] = (1 << 28) / ms / 1e6
    d_ptr = torch.zeros(n + 1, dtype=torch.int32, device="cuda"); d_cont = torch.zeros(n * per, dtype=torch.int16, device="cuda")
    d_final = torch.zeros(n * 5, dtype=torch.int16, device="cuda")
    ts = torch.cuda.Stream(); st = ts.cuda_stream
    g.synth_reads_device(2, 1, T, 4_000_000, 0, n, L, 10, 0, d_ptr.data_ptr(), d_cont.data_ptr(), st); g.stats(sync_stream=st, sync=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(5):
        if it == 2: e0.record(ts)
        g.classify_device(d_ptr.data_ptr(), d_cont.data_ptr(), n, d_final.data_ptr(), 0, st)
    e1.record(ts); torch.cuda.synchronize()
    r["classify_Glps"] = n * 120 / (e0.elapsed_time(e1) / 3) / 1e6
    print(json.dumps(r), flush=True)
else:
    for gran in ("128", "64", "32"):
        env = dict(os.environ, CUCLARK_L2_FETCH=gran)
        subprocess.run([sys.executable, __file__, "child"], env=env)
