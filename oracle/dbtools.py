"""Numpy restatement of the reference's database builder and file writer.

TEST INFRASTRUCTURE ONLY (oracle/). Used to make databases for the tests
without the reference binary, and pinned against it: tests/golden/*.json hold
sha256 digests of the .sz/.ky/.lb files the UNMODIFIED reference binaries wrote
for the same seeded inputs (tests/golden/make_golden.py), and
tests/test_oracle_golden.py checks that this module reproduces them bit for bit.

Follows (paths relative to /root/reference):
  full  scanner   src/CuCLARK_hh.hh:896-975   every overlapping k-mer
  light scanner   src/CuCLARK_hh.hh:694-767   every g-th NON-overlapping k-mer
  addElement      src/HashTableStorage_hh.hh:484-523   canonical = min(F, R)
  RemoveCommon    src/HashTableStorage_hh.hh:242-292   keep k-mers of exactly one target
  write           src/hashTable_hh.hh:591-663          .sz/.ky/.lb, bucket-major, ascending
"""
from __future__ import annotations

import hashlib

import numpy as np

U64 = np.uint64


def revcomp_codes(x: np.ndarray, k: int) -> np.ndarray:
    """Reverse the 2-bit groups, complement, shift down (src/kmersConversion.cc:39-47)."""
    x = x.astype(U64)
    x = ((x >> U64(2)) & U64(0x3333333333333333)) | ((x & U64(0x3333333333333333)) << U64(2))
    x = ((x >> U64(4)) & U64(0x0F0F0F0F0F0F0F0F)) | ((x & U64(0x0F0F0F0F0F0F0F0F)) << U64(4))
    x = ((x >> U64(8)) & U64(0x00FF00FF00FF00FF)) | ((x & U64(0x00FF00FF00FF00FF)) << U64(8))
    x = ((x >> U64(16)) & U64(0x0000FFFF0000FFFF)) | ((x & U64(0x0000FFFF0000FFFF)) << U64(16))
    x = (x >> U64(32)) | (x << U64(32))
    return (~x) >> U64(64 - 2 * k)


def canonical(x: np.ndarray, k: int) -> np.ndarray:
    return np.minimum(x.astype(U64), revcomp_codes(x, k))


def kmers_of_codes(codes: np.ndarray, k: int, starts: np.ndarray) -> np.ndarray:
    """R-form integer (first base in the high bits, complement code) of the
    windows starting at ``starts`` of an ACGT code array (0..3)."""
    out = np.zeros(starts.size, U64)
    comp = (3 - codes).astype(U64)
    for j in range(k):
        out = (out << U64(2)) | comp[starts + j]
    return out


def target_kmers(codes: np.ndarray, k: int, light_gap: int = 0) -> np.ndarray:
    """Canonical k-mers the reference adds for one all-ACGT target sequence."""
    n = codes.size
    if n < k:
        return np.zeros(0, U64)
    if light_gap:
        nk = n // k
        starts = np.arange(0, nk, light_gap, dtype=np.int64) * k
    else:
        starts = np.arange(0, n - k + 1, dtype=np.int64)
    return canonical(kmers_of_codes(codes, k, starts), k)


def build_entries(targets: list[np.ndarray], k: int, light_gap: int = 0):
    """(canonical k-mers, labels) of the k-mers seen in exactly ONE target, any multiplicity within it."""
    ks, ls = [], []
    for t, codes in enumerate(targets):
        u = np.unique(target_kmers(codes, k, light_gap))
        ks.append(u)
        ls.append(np.full(u.size, t, np.uint16))
    allk = np.concatenate(ks) if ks else np.zeros(0, U64)
    alll = np.concatenate(ls) if ls else np.zeros(0, np.uint16)
    order = np.argsort(allk, kind="stable")
    allk, alll = allk[order], alll[order]
    if allk.size == 0:
        return allk, alll
    first = np.ones(allk.size, bool)
    first[1:] = allk[1:] != allk[:-1]
    last = np.ones(allk.size, bool)
    last[:-1] = allk[1:] != allk[:-1]
    single = first & last
    return allk[single], alll[single]


def key_dtype(key_bytes: int):
    return {2: np.uint16, 4: np.uint32, 8: np.uint64}[key_bytes]


def entries_to_arrays(kmers: np.ndarray, labels: np.ndarray, htsize: int, key_bytes: int):
    """(.sz, .ky, .lb) arrays: bucket r = c mod HTSIZE, key = c div HTSIZE, ascending key in a bucket."""
    r = (kmers % U64(htsize)).astype(np.int64)
    q = kmers // U64(htsize)
    order = np.lexsort((q, r))
    r, q, lab = r[order], q[order], labels[order]
    sz = np.bincount(r, minlength=htsize)
    if sz.max(initial=0) >= 256:
        raise ValueError("This table can not be stored on disk: Some bucket list size exceeds 255.")
    return sz.astype(np.uint8), q.astype(key_dtype(key_bytes)), lab.astype(np.uint16)


def write_db_files(base: str, sz: np.ndarray, ky: np.ndarray, lb: np.ndarray) -> None:
    sz.tofile(base + ".sz")
    ky.tofile(base + ".ky")
    lb.tofile(base + ".lb")


def read_db_files(base: str, htsize: int, key_bytes: int):
    sz = np.fromfile(base + ".sz", np.uint8)
    assert sz.size == htsize, (sz.size, htsize)
    return sz, np.fromfile(base + ".ky", key_dtype(key_bytes)), np.fromfile(base + ".lb", np.uint16)


def arrays_to_entries(sz, ky, lb, htsize: int):
    r = np.repeat(np.arange(htsize, dtype=U64), sz)
    return ky.astype(U64) * U64(htsize) + r, lb


def sha256_file(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 24), b""):
            h.update(chunk)
    return h.hexdigest()


def sha256_array(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def db_name(folder: str, k: int, n_labels: int, htsize: int, min_count: int = 0, light_gap: int = 0) -> str:
    """getdbName (src/CuCLARK_hh.hh:580-591)."""
    folder = folder if folder.endswith("/") else folder + "/"
    base = f"{folder}/db_central_k{k}_t{n_labels}_s{htsize}_m{min_count}"
    return base + (f"_light_{light_gap}.tsk" if light_gap else ".tsk")
