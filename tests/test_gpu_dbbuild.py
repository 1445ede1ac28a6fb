"""GPU: the device database builder (cuclark_build_database) against
 (a) sha256 digests of the .sz/.ky/.lb files the UNMODIFIED reference binaries built from the same seeded
     targets (tests/golden/db_*.json, made by tests/golden/make_golden.py --stage db), and
 (b) a pure-Python restatement of the reference's target scanners + addElement + RemoveCommon on edge cases
     (src/CuCLARK_hh.hh:705-767 light, :914-975 full; src/HashTableStorage_hh.hh:242-292, 484-523)."""
import os
import sys

import numpy as np
import pytest

from cuclark_b200 import api
from oracle import dbtools

pytestmark = pytest.mark.gpu

CODE = {c: i for i, c in enumerate("ACGT")}
CODE.update({c.lower(): i for c, i in list(CODE.items())})
CODE["U"] = CODE["u"] = 3


def scan_target(data: bytes, k: int, gap: int):
    """R-form k-mers the reference scanner hands to addElement for one FASTA file, in order."""
    out = []
    mask = (1 << (2 * k)) - 1
    i, n = 0, len(data)
    R, cpt, full, it = 0, 0, False, 0
    while i < n:
        ch = chr(data[i])
        if ch in CODE:
            R = ((R << 2) | (3 - CODE[ch])) & mask
            if gap:                                   # light: non-overlapping, every gap-th
                if cpt == k - 1:
                    if it % gap == 0:
                        out.append(R)
                    R, cpt = 0, 0
                    it += 1
                else:
                    cpt += 1
            else:                                     # full: every window
                if full:
                    out.append(R)
                elif cpt == k - 1:
                    full = True
                    out.append(R)
                else:
                    cpt += 1
            i += 1
        elif ch == "\n":
            i += 1
        elif ch == ">":
            R, cpt, full = 0, 0, False
            while i < n and data[i] != 10:
                i += 1
            i += 1
        else:
            R, cpt, full = 0, 0, False
            i += 1
    return out


def expected_db(files, labels, k, gap, htsize, key_bytes, min_count=0):
    seen = {}
    for data, lab in zip(files, labels):
        arr = np.array(scan_target(data, k, gap), dtype=np.uint64)
        for c in dbtools.canonical(arr, k).tolist() if arr.size else []:
            e = seen.get(c)
            if e is None:
                seen[c] = [lab, 1, True]
            else:
                if e[0] != lab:
                    e[2] = False
                if e[1] + 1 < 255:
                    e[1] += 1
    kept = sorted((c, e[0]) for c, e in seen.items() if e[2] and e[1] > min_count)
    kmers = np.array([c for c, _ in kept], dtype=np.uint64)
    labs = np.array([l for _, l in kept], dtype=np.uint16)
    return dbtools.entries_to_arrays(kmers, labs, htsize, key_bytes)


def read_files(base, htsize, key_bytes):
    return dbtools.read_db_files(base, htsize, key_bytes)


@pytest.mark.parametrize("name", ["light_small", "light_c1", pytest.param("full_small", marks=pytest.mark.slow)])
def test_builder_reproduces_reference_db_files(name, tmp_path):
    """Same seeded targets as tests/golden/make_golden.py -> identical .sz/.ky/.lb (sha256) as the reference binary."""
    import make_golden
    from conftest import load_case
    golden = load_case(name)
    case = golden["case"]
    make_golden.write_inputs(case, str(tmp_path))
    files, labels = [], []
    for line in open(tmp_path / "targets.txt"):
        f, lab = line.split()
        files.append(f)
        labels.append(int(lab[1:]))
    htsize = api.HTSIZE_LIGHT if case["light"] else api.HTSIZE_FULL
    base = dbtools.db_name(str(tmp_path / "db"), case["k"], case["n_targets"], htsize, 0, case["gap"])
    st = api.build_database(files, labels, base, case["k"], light=case["light"], light_gap=case["gap"])
    assert st["n_nucleotides"] == case["n_targets"] * case["genome_len"]
    for ext in (".sz", ".ky", ".lb"):
        assert os.path.getsize(base + ext) == golden["bytes"][ext], ext
        assert dbtools.sha256_file(base + ext) == golden["sha256"][ext], ext
    assert st["n_kmers_kept"] == golden["bytes"][".lb"] // 2


def edge_targets():
    rng = np.random.default_rng(17)
    seq = lambda n: "".join("ACGT"[i] for i in rng.integers(0, 4, n))
    a, b, c = seq(5000), seq(4000), seq(3000)
    wrap = lambda s, w=70: "\n".join(s[i:i + w] for i in range(0, len(s), w))
    f0 = f">chr1 first\n{wrap(a)}\n>chr2\n{wrap(b[:1000])}N{wrap(b[1000:2000], 61)}\n>tiny\nACGT\n>empty\n\n"
    # shares a[100:900] with f0 (common k-mers -> removed), lower case, RNA, \r line ends, '>' inside a sequence line
    f1 = (f">other\n{wrap(a[100:900].lower())}\n>rna\n{wrap(c[:1500].replace('T', 'U'))}\n"
          f">crlf\r\n{c[1500:1600]}\r\n{c[1600:1700]}\r\n>midgt\n{c[1700:1800]}>{c[1800:1900]}\n{c[1900:2100]}\n"
          f">nrun\n{c[2100:2200]}NNNNNNNNNN{c[2200:2300]}nn{c[2300:2326]}x{c[2326:2500]}\n")
    # same label as f0 (label 0): duplicates inside one label are kept; repeats of its own sequence
    f2 = f">dup\n{wrap(b[2000:3000])}\n{wrap(b[2000:3000])}\n>dup2\n{wrap(b[2500:3500], 80)}\n{b[3500:]}"     # no final newline
    f3 = f">one_line_genome\n{seq(20000)}\n"
    return [f0.encode(), f1.encode(), f2.encode(), f3.encode()], [0, 1, 0, 2]


@pytest.mark.parametrize("gap,k", [(4, 27), (1, 27), (0, 27), (0, 31), (0, 12), (5, 9)])
@pytest.mark.parametrize("min_count", [0, 1])
def test_builder_edge_cases(tmp_path, gap, k, min_count):
    datas, labels = edge_targets()
    files = []
    for i, d in enumerate(datas):
        p = tmp_path / f"t{i}.fa"
        p.write_bytes(d)
        files.append(str(p))
    hts = api.HTSIZE_LIGHT
    kb = api.key_bytes_for(k, hts)
    base = str(tmp_path / "db")
    st = api.build_database(files, labels, base, k, light=True, light_gap=gap, htsize=hts, min_count=min_count) if gap else \
        api.build_database(files, labels, base, k, light=False, htsize=hts, min_count=min_count)
    sz, ky, lb = expected_db(datas, labels, k, gap, hts, kb, min_count)
    gsz, gky, glb = read_files(base, hts, kb)
    assert st["key_bytes"] == kb
    assert np.array_equal(gsz, sz)
    assert np.array_equal(gky, ky) and np.array_equal(glb, lb)
    assert st["n_kmers_kept"] == ky.size
    if min_count == 0 or gap == 0:
        assert ky.size > 100          # the case is not trivially empty


def fastq_as_fasta(data: bytes) -> bytes:
    """A FASTQ target as the reference scans it (src/CuCLARK_hh.hh:986-1080; light :769-860): header line skipped,
    sequence line scanned, its newline resets the window, three lines skipped. Equivalent FASTA text for scan_target."""
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return b"".join(b">\n" + lines[i] + b"\n" for i in range(1, len(lines), 4))


def fastq_targets():
    rng = np.random.default_rng(99)
    seq = lambda n: rng.choice(np.frombuffer(b"ACGT", np.uint8), n).tobytes()
    recs = lambda name, seqs: b"".join(b"@%s_%d some text\n%s\n+\n%s\n" % (name, i, s, b"I" * len(s)) for i, s in enumerate(seqs))
    shared = seq(400)
    f0 = recs(b"a", [seq(300), seq(26), seq(27), shared, seq(150) + b"N" + seq(90), seq(500).lower()])
    f1 = recs(b"b", [seq(1000), shared[100:300], seq(40) + b"NNNN" + seq(41)])
    f2 = recs(b"c", [seq(2500)])[:-1]                           # no trailing newline
    f3 = (">fa_among_fastq\n%s\n" % seq(700).decode()).encode()
    return [f0, f1, f2, f3], [0, 1, 2, 1]


@pytest.mark.parametrize("gap,k", [(4, 27), (0, 27), (0, 31)])
def test_builder_fastq_targets(tmp_path, gap, k):
    """FASTQ target files (the reference accepts FASTA and FASTQ targets, src/CuCLARK_hh.hh:896-1112)."""
    datas, labels = fastq_targets()
    files = []
    for i, d in enumerate(datas):
        p = tmp_path / f"t{i}.fq"
        p.write_bytes(d)
        files.append(str(p))
    hts = api.HTSIZE_LIGHT
    kb = api.key_bytes_for(k, hts)
    base = str(tmp_path / "db")
    st = api.build_database(files, labels, base, k, light=bool(gap), light_gap=gap, htsize=hts)
    as_fasta = [fastq_as_fasta(d) if d[:1] == b"@" else d for d in datas]
    sz, ky, lb = expected_db(as_fasta, labels, k, gap, hts, kb)
    gsz, gky, glb = read_files(base, hts, kb)
    assert np.array_equal(gsz, sz) and np.array_equal(gky, ky) and np.array_equal(glb, lb)
    assert st["n_kmers_kept"] == ky.size and ky.size > 50


def test_builder_errors(tmp_path):
    p = tmp_path / "spectrum.txt"
    p.write_bytes(b"ACGTACGTACGTACGTACGTACGTACG 5\n")
    with pytest.raises(api.CuclarkError) as e:
        api.build_database([str(p)], [0], str(tmp_path / "db"), 27, light=True)
    assert e.value.code == -8 and "FASTA or FASTQ" in str(e.value)
    with pytest.raises(api.CuclarkError):
        api.build_database([str(p)], [0], str(tmp_path / "db"), 33, light=True)
    # a missing target file is skipped as the reference does ("Failed to open"), an empty one adds nothing
    q = tmp_path / "t.fa"
    q.write_bytes(b">t\n" + b"ACGT" * 50 + b"\n")
    (tmp_path / "empty.fa").write_bytes(b"")
    st = api.build_database([str(tmp_path / "nope.fa"), str(tmp_path / "empty.fa"), str(q)], [0, 1, 2], str(tmp_path / "db"),
                            27, light=True, light_gap=1)
    assert st["n_nucleotides"] == 200 and st["n_kmers_added"] == 7


def test_cli_builds_missing_database_then_classifies(light_small, tmp_path):
    """cuCLARK-l with an empty database directory: builds the files on the GPU (identical to the reference's),
    then classifies; the CSV equals the reference binary's."""
    import gzip
    import subprocess
    import make_golden
    from conftest import GOLDEN, ROOT, load_case
    from cuclark_b200 import build
    build.build_all()
    golden = load_case("light_small")
    reads = make_golden.write_inputs(golden["case"], str(tmp_path))
    exe = os.path.join(ROOT, "cuclark_b200", "bin", "cuCLARK-l")
    p = subprocess.run([exe, "-T", "targets.txt", "-D", "db/", "-O", os.path.basename(reads), "-R", "out"], cwd=str(tmp_path),
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    assert "Starting the creation of the database of targets specific 27-mers from input files..." in p.stderr
    assert "14716 27-mers successfully stored in database." in p.stderr
    base = str(tmp_path / "db" / golden["db_basename"])
    for ext in (".sz", ".ky", ".lb"):
        assert dbtools.sha256_file(base + ext) == golden["sha256"][ext], ext
    assert (tmp_path / "out.csv").read_bytes() == gzip.open(os.path.join(GOLDEN, "light_small.csv.gz")).read()
