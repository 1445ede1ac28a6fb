"""The C++ executables cuCLARK / cuCLARK-l (cuclark_b200/csrc/cli_main.cc): same command line, database
file names and result CSV as the reference (src/main.cc, src/CuCLARK_hh.hh:383-591). CPU tests cover
argument handling and the loud failure without a GPU; GPU tests compare the CSV files with the ones the
unmodified reference binaries wrote on a B200 (tests/golden/*.csv.gz) and with the oracle."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exes():
    from cuclark_b200 import build
    build.build_all()
    full = os.path.join(ROOT, "cuclark_b200", "bin", "cuCLARK")
    light = os.path.join(ROOT, "cuclark_b200", "bin", "cuCLARK-l")
    assert os.path.exists(full) and os.path.exists(light)
    return full, light


def run(cmd, cwd=None):
    return subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=600)


def setup_case(case, folder, write_reads=True):
    """targets.txt + (empty) target files + database files named as getdbName does + the reads file."""
    from oracle import dbtools
    os.makedirs(os.path.join(folder, "tg"))
    os.makedirs(os.path.join(folder, "db"))
    with open(os.path.join(folder, "targets.txt"), "w") as tf:
        for t, name in enumerate(case.names):
            p = os.path.join(folder, "tg", f"{name}.fa")
            open(p, "w").write(f">{name}\n")
            tf.write(f"{p} {name}\n")
    base = dbtools.db_name(os.path.join(folder, "db"), case.k, case.n_targets, case.htsize, 0, case.case["gap"])
    dbtools.write_db_files(base, *case.arrays)
    reads = os.path.join(folder, "reads.fa" if case.case["fmt"] == "fasta" else "reads.fq")
    if write_reads:
        with open(reads, "wb") as f:
            f.write(case.reads_bytes)
    return reads


# ------------------------------------------------------------------ CPU
def test_version_help_and_argument_errors(exes, tmp_path):
    full, light = exes
    p = run([full, "--version"])
    assert p.returncode == 0 and p.stdout.startswith("Version: 1.1 ")
    p = run([light, "--help"])
    assert p.returncode == 0 and "-T <fileTargets>" in p.stdout and "-P <file1> <file2>" in p.stdout
    p = run([full, "-T", "x"])
    assert p.returncode == 255 and "at least four  parameters are necessary" in p.stderr
    t = tmp_path / "t.txt"
    t.write_text("")
    p = run([full, "-T", str(t), "-D", str(tmp_path), "-O", str(t), "-R", "r", "--bogus"])
    assert p.returncode == 1 and "Failed to recognize option: --bogus" in p.stderr
    p = run([full, "-k", "33", "-T", str(t), "-D", str(tmp_path), "-O", str(t), "-R", "r"])
    assert p.returncode == 1 and "The k-mer length should be in [2,32]." in p.stderr
    p = run([full, "-T", str(tmp_path / "none"), "-D", str(tmp_path), "-O", str(t), "-R", "r"])
    assert p.returncode == 1 and "Failed to find/read the file of the targets definition" in p.stderr
    p = run([full, "-s", "31", "-T", str(t), "-D", str(tmp_path), "-O", str(t), "-R", "r"])
    assert p.returncode == 1 and "sampling factor value should be in the interval [2,30]" in p.stderr
    p = run([full, "-n", "4", "-b", "2", "-T", str(t), "-D", str(tmp_path), "-O", str(t), "-R", "r"])
    assert p.returncode == 1 and "number of batches should be higher than the number of threads" in p.stderr
    # a target file that does not exist (src/CuCLARK_hh.hh:1808-1813)
    t.write_text("/nonexistent/genome.fa L1\n")
    p = run([full, "-T", str(t), "-D", str(tmp_path), "-O", str(t), "-R", "r"])
    assert p.returncode == 255 and "Failed to open file: /nonexistent/genome.fa defined in" in p.stderr
    # database absent
    g = tmp_path / "g.fa"
    g.write_text(">g\nACGT\n")
    t.write_text(f"{g} L1\n")
    if not has_gpu():
        # no database files: the builder is called, and without a GPU it refuses loudly (no host builder)
        p = run([light, "-T", str(t), "-D", str(tmp_path), "-O", str(g), "-R", "r", "-g", "5"])
        assert p.returncode == 1
        assert "Starting the creation of the database of targets specific 27-mers" in p.stderr
        assert "Not enough CUDA devices found" in p.stderr and "no CPU fallback" in p.stderr


def test_cli_fails_loudly_without_gpu(exes, light_small, tmp_path):
    """No CPU fallback: with a valid database but no CUDA device the executable exits 1 with a message."""
    if has_gpu():
        pytest.skip("GPU present")
    reads = setup_case(light_small, str(tmp_path))
    p = run([exes[1], "-T", "targets.txt", "-D", "db/", "-O", os.path.basename(reads), "-R", "out"], cwd=str(tmp_path))
    assert p.returncode == 1
    assert "Not enough CUDA devices found" in p.stderr and "no CPU fallback" in p.stderr
    assert not os.path.exists(tmp_path / "out.csv")


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cli_light_fasta_equals_reference_csv(exes, light_small, tmp_path):
    reads = setup_case(light_small, str(tmp_path))
    p = run([exes[1], "-T", "targets.txt", "-D", "db/", "-O", os.path.basename(reads), "-R", "out", "-n", "4"],
            cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    ref = gzip.open(os.path.join(GOLDEN, "light_small.csv.gz")).read()
    assert (tmp_path / "out.csv").read_bytes() == ref
    assert "Processing file 'reads.fa' in 4 batches using 4 CPU thread(s)." in p.stdout
    assert " objects/min. (20000 objects)." in p.stdout and " - Results stored in out.csv" in p.stdout
    assert "Loading database [db//db_central_k27_t8_s57777779_m0_light_4.tsk.*] (s=1)..." in p.stderr
    # --extended, against the oracle's extended CSV; and a list of inputs (results file exists -> list mode)
    p = run([exes[1], "-T", "targets.txt", "-D", "db/", "-O", "reads.fa", "-R", "ext", "--extended"], cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    head = (tmp_path / "ext.csv").read_bytes().split(b"\n", 1)[0]
    assert head == b"Object_ID," + b",".join(n.encode() for n in light_small.names) + \
        b",Length,Gamma,1st_assignment,score1,2nd_assignment,score2,confidence"
    (tmp_path / "inputs.txt").write_text("reads.fa\nreads.fa\n")
    (tmp_path / "outputs.txt").write_text("o1\no2\n")
    p = run([exes[1], "-T", "targets.txt", "-D", "db/", "-O", "inputs.txt", "-R", "outputs.txt"], cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    assert (tmp_path / "o1.csv").read_bytes() == ref and (tmp_path / "o2.csv").read_bytes() == ref


@pytest.mark.gpu
def test_cli_through_the_local_table_layout(exes, light_small, tmp_path):
    """The same command line with the table in the LOCAL layout (minimizer-addressed lines; chosen automatically
    only at bacterial scale, forced here through CUCLARK_LAYOUT): byte-identical CSVs, plain and --extended."""
    reads = setup_case(light_small, str(tmp_path))
    env = dict(os.environ, CUCLARK_LAYOUT="3", CUCLARK_TIMING="1")
    p = subprocess.run([exes[1], "-T", "targets.txt", "-D", "db/", "-O", os.path.basename(reads), "-R", "out", "-n", "2"],
                       cwd=str(tmp_path), capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr
    ref = gzip.open(os.path.join(GOLDEN, "light_small.csv.gz")).read()
    assert (tmp_path / "out.csv").read_bytes() == ref
    assert "(table layout 3)" in p.stderr
    p1 = subprocess.run([exes[1], "-T", "targets.txt", "-D", "db/", "-O", "reads.fa", "-R", "ext_local", "--extended"],
                        cwd=str(tmp_path), capture_output=True, text=True, env=env)
    p2 = subprocess.run([exes[1], "-T", "targets.txt", "-D", "db/", "-O", "reads.fa", "-R", "ext_hashed", "--extended"],
                        cwd=str(tmp_path), capture_output=True, text=True, env=dict(os.environ, CUCLARK_LAYOUT="1"))
    assert p1.returncode == 0 and p2.returncode == 0, p1.stderr + p2.stderr
    assert (tmp_path / "ext_local.csv").read_bytes() == (tmp_path / "ext_hashed.csv").read_bytes()


@pytest.mark.gpu
def test_cli_extended_equals_oracle(exes, oracle, light_small, tmp_path):
    c = light_small
    reads = setup_case(c, str(tmp_path))
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ix, buf = oracle.index(c.reads_bytes, 1)
    ptr, cont = oracle.pack(ix, buf, c.k)
    final, rows, _ = oracle.classify(odb, ptr, cont, c.n_targets, c.maxhits, threads=4)
    oracle.write_csv(str(tmp_path / "oracle_ext.csv"), ix, buf, c.k, False, c.names, final, rows, c.maxhits)
    oracle.free_index(ix)
    p = run([exes[1], "-T", "targets.txt", "-D", "db/", "-O", os.path.basename(reads), "-R", "ext", "--extended"],
            cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    assert (tmp_path / "ext.csv").read_bytes() == (tmp_path / "oracle_ext.csv").read_bytes()


@pytest.mark.gpu
def test_cli_paired_end(exes, oracle, light_small, tmp_path):
    """-P: mates merged as '>id\\n<seq1>N<seq2>' (src/file.cc:205-268), Length column minus the N."""
    c = light_small
    setup_case(c, str(tmp_path), write_reads=False)
    import sys
    asc = np.frombuffer(b"ACGT", np.uint8)
    mg = sys.modules["make_golden"]
    rng = np.random.default_rng(4)
    f1, f2 = [], []
    for i in range(3000):
        t = int(rng.integers(0, c.n_targets))
        g = asc[mg.target_codes(c.case, t)].tobytes()
        pos = int(rng.integers(0, len(g) - 500))
        s1, s2 = g[pos:pos + 100], g[pos + 250:pos + 350]
        f1.append(b"@pair%d/1\n%s\n+\n%s\n" % (i, s1, b"I" * 100))
        f2.append(b"@pair%d/2 extra\n%s\n+\n%s\n" % (i, s2, b"I" * 100))
    (tmp_path / "r1.fq").write_bytes(b"".join(f1))
    (tmp_path / "r2.fq").write_bytes(b"".join(f2))
    assert oracle.merge_paired(str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq"), str(tmp_path / "merged.fa")) == 0
    merged = (tmp_path / "merged.fa").read_bytes()
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ix, buf = oracle.index(merged, 1)
    ptr, cont = oracle.pack(ix, buf, c.k)
    final, rows, _ = oracle.classify(odb, ptr, cont, c.n_targets, c.maxhits, threads=4)
    oracle.write_csv(str(tmp_path / "oracle.csv"), ix, buf, c.k, True, c.names, final, None, c.maxhits)
    oracle.free_index(ix)
    p = run([exes[1], "-T", "targets.txt", "-D", "db/", "-P", "r1.fq", "r2.fq", "-R", "paired"], cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    got = (tmp_path / "paired.csv").read_bytes()
    assert got == (tmp_path / "oracle.csv").read_bytes()
    assert got.split(b"\n")[1].startswith(b"pair0,200,")
    assert not os.path.exists(tmp_path / "r1.fq_ConcatenatedByCLARK.fa")      # deleted as the reference does
    assert "Processing file: 'r1.fq_ConcatenatedByCLARK.fa'" in p.stdout


@pytest.mark.gpu
@pytest.mark.slow
def test_cli_config3_paired_k31_with_errors(exes, oracle, full_small, tmp_path):
    """BASELINE configs[2] in small: cuCLARK k=31, paired-end 2 x 150 bp, mate 2 from the opposite strand, 1 %
    substitution errors; mates merged by the reference rule, Length 300, gamma over 300-31+1 (paired-mode voting)."""
    c = full_small
    setup_case(c, str(tmp_path), write_reads=False)
    import sys
    asc = np.frombuffer(b"ACGT", np.uint8)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    mg = sys.modules["make_golden"]
    rng = np.random.default_rng(31)
    genomes = [asc[mg.target_codes(c.case, t)].tobytes() for t in range(c.n_targets)]

    def with_errors(seq: bytes) -> bytes:
        b = bytearray(seq)
        for j in np.nonzero(rng.random(len(b)) < 0.01)[0]:
            b[j] = b"ACGT"[(b"ACGT".index(b[j]) + 1 + int(rng.integers(0, 3))) % 4]
        return bytes(b)

    f1, f2 = [], []
    for i in range(4000):
        g = genomes[int(rng.integers(0, c.n_targets))]
        pos = int(rng.integers(0, len(g) - 600))
        s1 = with_errors(g[pos:pos + 150])
        s2 = with_errors(g[pos + 350:pos + 500].translate(comp)[::-1])
        f1.append(b"@frag%d/1\n%s\n+\n%s\n" % (i, s1, b"F" * 150))
        f2.append(b"@frag%d/2\n%s\n+\n%s\n" % (i, s2, b"F" * 150))
    (tmp_path / "r1.fq").write_bytes(b"".join(f1))
    (tmp_path / "r2.fq").write_bytes(b"".join(f2))
    assert oracle.merge_paired(str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq"), str(tmp_path / "merged.fa")) == 0
    merged = (tmp_path / "merged.fa").read_bytes()
    sz, ky, lb = c.arrays
    odb = oracle.db_from_arrays(c.htsize, c.k, sz, ky, lb)
    ix, buf = oracle.index(merged, 1)
    ptr, cont = oracle.pack(ix, buf, c.k)
    final, rows, _ = oracle.classify(odb, ptr, cont, c.n_targets, c.maxhits, threads=4)
    oracle.write_csv(str(tmp_path / "oracle.csv"), ix, buf, c.k, True, c.names, final, None, c.maxhits)
    oracle.free_index(ix)
    p = run([exes[0], "-k", "31", "-T", "targets.txt", "-D", "db/", "-P", "r1.fq", "r2.fq", "-R", "paired"], cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    got = (tmp_path / "paired.csv").read_bytes()
    assert got == (tmp_path / "oracle.csv").read_bytes()
    assert got.split(b"\n")[1].startswith(b"frag0,300,")
    assert (final[:, 1] > 0).mean() > 0.95 and (final[:, 0] < 2 * (150 - 31 + 1)).mean() > 0.9    # classified, but errors cost k-mers


@pytest.mark.gpu
@pytest.mark.slow
def test_cli_full_fastq_equals_reference_csv(exes, full_small, tmp_path):
    reads = setup_case(full_small, str(tmp_path))
    p = run([exes[0], "-k", "31", "-T", "targets.txt", "-D", "db/", "-O", os.path.basename(reads), "-R", "out"],
            cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    ref = gzip.open(os.path.join(GOLDEN, "full_small.csv.gz")).read()
    assert (tmp_path / "out.csv").read_bytes() == ref


@pytest.mark.gpu
def test_cli_table_cache(exes, light_small, tmp_path):
    """--cache: the first run writes <db>.b200, the second streams it back; a stale cache is ignored and
    rewritten. The CSV is the reference binary's in all three runs."""
    reads = setup_case(light_small, str(tmp_path))
    ref = gzip.open(os.path.join(GOLDEN, "light_small.csv.gz")).read()
    cmd = [exes[1], "-T", "targets.txt", "-D", "db/", "-O", os.path.basename(reads), "-R", "out", "--cache"]
    cache = tmp_path / "db" / "db_central_k27_t8_s57777779_m0_light_4.tsk.b200"
    p = run(cmd, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    assert "written." in p.stderr and cache.exists()
    assert (tmp_path / "out.csv").read_bytes() == ref
    os.remove(tmp_path / "out.csv")
    p = run(cmd, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    assert "Table cache" in p.stderr and "loaded." in p.stderr and "written." not in p.stderr
    assert (tmp_path / "out.csv").read_bytes() == ref
    # damage the cache: it is refused, the database files are used, the cache is rewritten
    blob = bytearray(cache.read_bytes()); blob[4096] ^= 1
    cache.write_bytes(blob)
    os.remove(tmp_path / "out.csv")
    p = run(cmd, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    assert "Ignoring table cache" in p.stderr and "written." in p.stderr
    assert (tmp_path / "out.csv").read_bytes() == ref
