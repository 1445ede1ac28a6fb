"""CPU: the oracle's packer on runs of 65,536 nt and more (deviation Q8, oracle/cuclark_oracle.c orc_pack)."""
import numpy as np

from cuclark_b200 import synth


def direct_kmers(codes, k):
    """R-form integers of every window: complement code, first base in the high bits."""
    rc = (3 - codes).astype(np.uint64)
    out = np.zeros(codes.size - k + 1, np.uint64)
    for j in range(k):
        out = (out << np.uint64(2)) | rc[j:j + out.size]
    return out


def test_long_runs_are_split_with_overlap(oracle):
    k = 31
    asc = np.frombuffer(b"ACGT", np.uint8)
    a = synth.genome_codes(3, 0, 0, 200_000)
    wrap = lambda s, w: b"\n".join(s[i:i + w] for i in range(0, len(s), w))
    lens = [65_535, 65_536, 65_535 + k - 1, 70_000, 140_000, 150]
    recs = [b">r%d\n" % i + wrap(asc[a[i:i + L]].tobytes(), 70 + i) + b"\n" for i, L in enumerate(lens)]
    ix, buf = oracle.index(b"".join(recs), 1)
    ptr, cont = oracle.pack(ix, buf, k)
    oracle.free_index(ix)
    # headers: <= 65,535, split parts overlap by k-1
    h = []
    for r in range(len(lens)):
        p, parts = int(ptr[r]), []
        while p < ptr[r + 1]:
            L = int(cont[p]); parts.append(L); p += 1 + (L + 7) // 8
        assert p == ptr[r + 1]
        h.append(parts)
    assert h[0] == [65_535] and h[1] == [65_535, k] and h[2] == [65_535, 2 * k - 2]
    assert h[3] == [65_535, 70_000 - 65_535 + k - 1]
    assert h[4] == [65_535, 65_535, 140_000 - 2 * (65_535 - (k - 1))] and h[5] == [150]
    km = oracle.extract(ptr, cont, k)
    exp = np.concatenate([direct_kmers(a[i:i + L], k) for i, L in enumerate(lens)])
    assert np.array_equal(km, exp)                   # every window exactly once, in order
