#!/usr/bin/env python
"""Cold vs warm cost of cuclark_classify_file (what the CLI calls once per process) for several
chunk sizes / slot counts. CUCLARK_TIMING=1 makes the library print its phase times to stderr."""
import argparse, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--targets", type=int, default=8)
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--configs", default="16x64,16x16,8x16,4x16,16x8,8x8,4x64")
    a = ap.parse_args()
    os.environ["CUCLARK_TIMING"] = "1"
    import torch
    from cuclark_b200.api import CuClarkDB, HTSIZE_FULL
    K, G, L = 31, 4_000_000, 150
    rec = 16 + 2 * L
    d = tempfile.mkdtemp(prefix="probe_", dir="/dev/shm")
    fq, csv = os.path.join(d, "reads.fq"), os.path.join(d, "out.csv")
    with CuClarkDB(K, a.targets, htsize=HTSIZE_FULL) as g:
        g.build_synthetic(1, a.targets, G, 0)
        t = torch.empty(a.reads * rec, dtype=torch.uint8, device="cuda")
        g.synth_fastq_device(2, 1, a.targets, G, 0, a.reads, L, 10, 0, t.data_ptr(), 0)
        g.stats(sync=True)
        torch.cuda.synchronize()
        t.cpu().numpy().tofile(fq)
        del t
        g.save_table(os.path.join(d, "t.b200"))
    ref = None
    for cfg in a.configs.split(","):
        slots, mb = (int(x) for x in cfg.split("x"))
        with CuClarkDB(K, a.targets, htsize=HTSIZE_FULL) as g:
            assert g.load_table(os.path.join(d, "t.b200"))
            res = {"slots": slots, "chunk_mb": mb}
            for label in ("cold", "warm", "warm2"):
                sys.stderr.write(f"--- {cfg} {label}\n"); sys.stderr.flush()
                t0 = time.perf_counter()
                st = g.classify_file(fq, csv, chunk_bytes=mb << 20, n_slots=slots)
                res[label + "_s"] = round(time.perf_counter() - t0, 4)
            blob = open(csv, "rb").read()
            if ref is None:
                ref = blob
            res["same_csv"] = blob == ref
            res["reads"] = st["n_reads"]
            print(json.dumps(res), flush=True)
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
    os.rmdir(d)


if __name__ == "__main__":
    main()
