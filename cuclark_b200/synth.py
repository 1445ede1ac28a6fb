"""Deterministic synthetic genomes and reads (counter-based, no RNG state).

Every base is a pure function of (seed, target, position) and every read is a
pure function of (seed, read index), so the numpy code here and the CUDA
generators in ``csrc/synth.cu`` produce identical data at any scale: the tests
use this module at sizes the oracle can check, ``bench.py`` uses the device
twin at BASELINE.json's full sizes.

Shapes follow SURVEY.md section 8(d): i.i.d. uniform ACGT targets, reads
sampled uniformly from the targets (half of them reverse-complemented), a
fraction of non-target random reads, optional single ``N`` and per-base
substitutions.
"""
from __future__ import annotations

import numpy as np

U64 = np.uint64
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)

# stream tags (keep in sync with csrc/synth.cuh)
TAG_GENOME = 0x47
TAG_READ = 0x52
TAG_RBASE = 0x62
TAG_SUB = 0x73


def mix64(x):
    """splitmix64 finaliser on uint64 arrays (wraps mod 2**64)."""
    x = np.asarray(x, dtype=U64)
    with np.errstate(over="ignore"):
        x = (x ^ (x >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> U64(27))) * U64(0x94D049BB133111EB)
        return x ^ (x >> U64(31))


def _key(tag: int, seed: int, a, b):
    """Two rounds of mixing over (tag, seed, a, b); a < 2**32, b < 2**40."""
    a = np.asarray(a, dtype=U64)
    b = np.asarray(b, dtype=U64)
    with np.errstate(over="ignore"):
        h = mix64((U64(tag) << U64(56)) ^ (U64(seed & 0xFFFF) << U64(40)) ^ b)
        return mix64(h ^ (a * U64(0x9E3779B97F4A7C15)))


def genome_codes(seed: int, target: int, start: int, n: int) -> np.ndarray:
    """Codes 0..3 (A,C,G,T) of ``n`` bases of ``target`` from ``start``.

    One 64-bit hash yields 32 bases: base p lives in bits 2*(p%32) of
    hash(seed, target, p//32).
    """
    p = np.arange(start, start + n, dtype=np.int64)
    w = (p >> 5).astype(U64)
    w0 = int(w[0]) if n else 0
    words = _key(TAG_GENOME, seed, U64(target), np.arange(w0, (int(w[-1]) if n else 0) + 1, dtype=U64))
    h = words[(w - U64(w0)).astype(np.int64)]
    return ((h >> (U64(2) * (p & 31).astype(U64))) & U64(3)).astype(np.uint8)


def genome_ascii(seed: int, target: int, length: int) -> bytes:
    return _BASES[genome_codes(seed, target, 0, length)].tobytes()


def write_fasta(path: str, name: str, seq: bytes, width: int = 70) -> None:
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        for i in range(0, len(seq), width):
            f.write(seq[i:i + width] + b"\n")


def read_codes(seed: int, n_reads: int, read_len: int, n_targets: int, genome_len: int,
               genome_seed: int, pct_random: int = 0, pct_n: int = 0, sub_per_10k: int = 0,
               first: int = 0):
    """Reads ``first .. first+n_reads`` as a (n_reads, read_len) uint8 matrix.

    Codes 0..3 = ACGT, 4 = N.  Returns (codes, target, pos, is_rc, is_random);
    target is -1 for random reads.
    """
    i = np.arange(first, first + n_reads, dtype=U64)
    h1 = _key(TAG_READ, seed, i, U64(0))
    h2 = _key(TAG_READ, seed, i, U64(1))
    h3 = _key(TAG_READ, seed, i, U64(2))
    is_random = (h1 % U64(100)) < U64(pct_random)
    target = ((h1 >> U64(8)) % U64(n_targets)).astype(np.int64)
    is_rc = ((h1 >> U64(40)) & U64(1)).astype(bool)
    pos = (h2 % U64(genome_len - read_len + 1)).astype(np.int64)

    j = np.arange(read_len, dtype=np.int64)
    codes = np.empty((n_reads, read_len), dtype=np.uint8)
    # target-derived reads: gather genome bases position by position
    gp = pos[:, None] + j[None, :]
    gw = (gp >> 5).astype(U64)
    gh = _key(TAG_GENOME, genome_seed, target.astype(U64)[:, None], gw)
    fwd = ((gh >> (U64(2) * (gp & 31).astype(U64))) & U64(3)).astype(np.uint8)
    rc = (3 - fwd)[:, ::-1]
    codes[:] = np.where(is_rc[:, None], rc, fwd)
    # random reads: independent base stream
    if pct_random:
        rh = _key(TAG_RBASE, seed, i[:, None], (j >> 5).astype(U64)[None, :])
        rnd = ((rh >> (U64(2) * (j & 31).astype(U64))[None, :]) & U64(3)).astype(np.uint8)
        codes[:] = np.where(is_random[:, None], rnd, codes)
    if sub_per_10k:
        sh = _key(TAG_SUB, seed, i[:, None], j.astype(U64)[None, :])
        hit = (sh % U64(10000)) < U64(sub_per_10k)
        delta = (1 + ((sh >> U64(20)) % U64(3))).astype(np.uint8)
        codes[:] = np.where(hit, (codes + delta) & 3, codes)
    if pct_n:
        has_n = ((h3 % U64(100)) < U64(pct_n))
        npos = ((h3 >> U64(8)) % U64(read_len)).astype(np.int64)
        rows = np.nonzero(has_n)[0]
        codes[rows, npos[rows]] = 4
    target = np.where(is_random, -1, target)
    return codes, target, pos, is_rc, is_random


_ASCII5 = np.frombuffer(b"ACGTN", dtype=np.uint8)


def reads_fasta(codes: np.ndarray, prefix: str = "r", first: int = 0) -> bytes:
    out = bytearray()
    asc = _ASCII5[codes]
    for r in range(codes.shape[0]):
        out += b">" + prefix.encode() + str(first + r).encode() + b"\n" + asc[r].tobytes() + b"\n"
    return bytes(out)


def reads_fastq(codes: np.ndarray, prefix: str = "r", first: int = 0, suffix: str = "") -> bytes:
    out = bytearray()
    asc = _ASCII5[codes]
    qual = b"I" * codes.shape[1]
    for r in range(codes.shape[0]):
        out += (b"@" + prefix.encode() + str(first + r).encode() + suffix.encode() + b"\n"
                + asc[r].tobytes() + b"\n+\n" + qual + b"\n")
    return bytes(out)


def read_truth(seed: int, first: int, n_reads: int, read_len: int, n_targets: int, pct_random: int,
               sub_per_10k: int = 0, k: int = 31, chunk: int = 1 << 20):
    """Ground truth of reads ``first .. first+n_reads`` as the generators define them, without touching a database:
    (target, clean) with target = source target (-1 for a random read) and clean = number of k-mer windows of the
    read that hold no substituted base (= read_len-k+1 without substitutions). A target-specific k-mer database of
    the same genomes classifies a sampled read as [clean, target+1, clean, 0, 0] and a random read as all zeros,
    except for the handful of reads that touch a k-mer common to two targets or hit one by chance
    (expected counts: bench.py). Works in chunks so that 10 M reads stay within a few hundred MB."""
    target = np.empty(n_reads, np.int64)
    clean = np.empty(n_reads, np.int64)
    nk = read_len - k + 1
    j = np.arange(read_len, dtype=U64)
    for lo in range(0, n_reads, chunk):
        hi = min(n_reads, lo + chunk)
        i = np.arange(first + lo, first + hi, dtype=U64)
        h1 = _key(TAG_READ, seed, i, U64(0))
        is_random = (h1 % U64(100)) < U64(pct_random)
        t = ((h1 >> U64(8)) % U64(n_targets)).astype(np.int64)
        target[lo:hi] = np.where(is_random, -1, t)
        if not sub_per_10k:
            clean[lo:hi] = nk
            continue
        for a in range(lo, hi, 1 << 17):                    # (reads x read_len) hash matrix: smaller pieces
            b = min(hi, a + (1 << 17))
            ii = np.arange(first + a, first + b, dtype=U64)
            sh = _key(TAG_SUB, seed, ii[:, None], j[None, :])
            hit = ((sh % U64(10000)) < U64(sub_per_10k)).astype(np.int32)
            cs = np.concatenate([np.zeros((b - a, 1), np.int32), np.cumsum(hit, axis=1, dtype=np.int32)], axis=1)
            clean[a:b] = ((cs[:, k:] - cs[:, :nk]) == 0).sum(axis=1)
    return target, clean
