"""CPU: the oracle against the pins made from the reference's own code.

tests/golden/db_*.json hold sha256 digests of the .sz/.ky/.lb files written by
the unmodified reference binaries (tests/golden/make_golden.py --stage db).
"""
import os

import numpy as np
import pytest

from oracle import dbtools
from oracle.binding import HTSIZE_LIGHT, RefLookup


@pytest.mark.parametrize("name", ["light_small", "light_c1"])
def test_db_files_match_reference_digests(name, request):
    case = request.getfixturevalue(name)
    sz, ky, lb = case.arrays
    g = case.golden
    assert sz.size == g["bytes"][".sz"] and ky.nbytes == g["bytes"][".ky"] and lb.nbytes == g["bytes"][".lb"]
    assert dbtools.sha256_array(sz) == g["sha256"][".sz"]
    assert dbtools.sha256_array(ky) == g["sha256"][".ky"]
    assert dbtools.sha256_array(lb) == g["sha256"][".lb"]


@pytest.mark.slow
def test_db_files_match_reference_digests_full(full_small):
    sz, ky, lb = full_small.arrays
    g = full_small.golden
    assert dbtools.sha256_array(ky) == g["sha256"][".ky"]
    assert dbtools.sha256_array(lb) == g["sha256"][".lb"]
    assert dbtools.sha256_array(sz) == g["sha256"][".sz"]


def test_removecommon_exercised(light_small):
    # 50 sampled 27-mers are shared between targets 0 and 1 and must be gone
    assert "14766 27-mers stored" in light_small.golden["reference_stderr_tail"]
    assert light_small.kmers.size == 14716


def test_port_roundtrip_entries(oracle, light_small):
    sz, ky, lb = light_small.arrays
    db = oracle.db_from_arrays(light_small.htsize, light_small.k, sz, ky, lb)
    km, lab = db.entries()
    order = np.argsort(km)
    assert np.array_equal(km[order], light_small.kmers)
    assert np.array_equal(lab[order], light_small.labels)
    got, hits = db.query(light_small.kmers)
    assert hits == light_small.kmers.size and np.array_equal(got, light_small.labels.astype(np.int32))


def test_port_lookup_equals_reference_find(oracle, light_small, tmp_path):
    """Stage 3 ground truth: the reference's hTable::read + find on the same files."""
    try:
        ref = RefLookup(light=True)
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    sz, ky, lb = light_small.arrays
    base = str(tmp_path / "db")
    dbtools.write_db_files(base, sz, ky, lb)
    ref.open(base, light_small.k)
    db = oracle.db_load(base, light_small.htsize, light_small.k)
    ix, buf = oracle.index(light_small.reads_bytes, 4)
    ptr, cont = oracle.pack(ix, buf, light_small.k)
    kmers = oracle.extract(ptr, cont, light_small.k)
    rng = np.random.default_rng(5)
    probe = np.concatenate([kmers, rng.integers(0, 1 << 54, 200_000, dtype=np.uint64), light_small.kmers,
                            dbtools.revcomp_codes(light_small.kmers, light_small.k)])
    a, ha = db.query(probe, 4)
    b, hb = ref.query(probe, 4)
    assert ha == hb and np.array_equal(a, b)
    assert ha >= 2 * light_small.kmers.size
    ref.close()
    oracle.free_index(ix)


@pytest.mark.parametrize("sfactor", [2, 3])
def test_port_sampling_equals_reference(oracle, light_small, tmp_path, sfactor):
    try:
        ref = RefLookup(light=True)
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built")
    sz, ky, lb = light_small.arrays
    base = str(tmp_path / "db")
    dbtools.write_db_files(base, sz, ky, lb)
    ref.open(base, light_small.k, sfactor=sfactor)
    db = oracle.db_load(base, light_small.htsize, light_small.k, sfactor=sfactor)
    a, ha = db.query(light_small.kmers)
    b, hb = ref.query(light_small.kmers)
    assert np.array_equal(a, b) and 0 < ha < light_small.kmers.size
    ref.close()


@pytest.mark.parametrize("k", [19, 30, 32])
def test_port_lookup_equals_reference_find_other_key_widths(oracle, tmp_path, k):
    """uint16 (k=19) and uint64 (k=30, 32) `.ky` elements: EHashtable<uint16_t/uint64_t, rElement> of the
    reference (the T16/T64 branches of src/main.cc:278-316) against the port, every DB k-mer in both
    orientations + random k-mers."""
    from cuclark_b200 import synth
    from oracle.binding import key_bytes_for
    try:
        ref = RefLookup(light=True)
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    targets = [synth.genome_codes(40 + k, t, 0, 20_000) for t in range(5)]
    kmers, labels = dbtools.build_entries(targets, k, 0)
    kb = key_bytes_for(k, HTSIZE_LIGHT)
    assert kb == (2 if k == 19 else 8)
    sz, ky, lb = dbtools.entries_to_arrays(kmers, labels, HTSIZE_LIGHT, kb)
    base = str(tmp_path / "db")
    dbtools.write_db_files(base, sz, ky, lb)
    ref.open(base, k)
    db = oracle.db_load(base, HTSIZE_LIGHT, k)
    rng = np.random.default_rng(k)
    hi = (1 << 64) - 1 if k == 32 else (1 << (2 * k)) - 1
    probe = np.concatenate([kmers, dbtools.revcomp_codes(kmers, k),
                            rng.integers(0, hi, 100_000, dtype=np.uint64, endpoint=True)])
    a, ha = db.query(probe, 4)
    b, hb = ref.query(probe, 4)
    assert ha == hb and np.array_equal(a, b)
    assert ha >= 2 * kmers.size
    ref.close()
