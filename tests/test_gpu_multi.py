"""The product's own multi-device paths in ONE process (the CLI's -d N), against the reference binary's CSVs:

* read-partitioned — N handles that each hold the whole table (loaded once, cloned device to device), chunks of
  the input dealt over the devices (cuclark_classify_text_buffer / cuclark_classify_file_multi with n_dbs > 1);
* table-partitioned — N shard handles, k-mers routed to their shard (csrc/route.cu) through the text pipeline.

The table-partitioned path also runs with all shards on ONE device (plain pointers instead of NVLink), which is what
the one-GPU test box exercises; the `two_devices` cases skip below 2 GPUs and are run with `gpurun --gpus 2`
(log under profiles/).
"""
import gzip
import os
import subprocess

import numpy as np
import pytest

from cuclark_b200.api import CuClarkDB
from conftest import GOLDEN
from test_cli import exes, run, setup_case          # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def n_gpus() -> int:
    import torch
    return torch.cuda.device_count()


def golden(name):
    return gzip.open(os.path.join(GOLDEN, name + ".csv.gz")).read()


def text_through(handles, data: bytes, names, **kw):
    import torch
    h_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    cap = len(data) + (1 << 20)
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    n, st = handles[0].classify_text_buffer(h_in.data_ptr(), len(data), h_out.data_ptr(), cap, names=names,
                                            peers=handles[1:], **kw)
    return bytes(h_out[:n].numpy()), st


@pytest.mark.parametrize("n_shards", [2, 3])
def test_text_pipeline_table_partitioned_one_device(light_small, n_shards):
    """FASTA bytes -> CSV bytes through N shard handles on device 0: the reference binary's CSV, byte for byte."""
    c = light_small
    sz, ky, lb = c.arrays
    shards = []
    for i in range(n_shards):
        g = CuClarkDB(c.k, c.n_targets, htsize=c.htsize, shard=(i, n_shards))
        g.load_arrays(sz, ky, lb)
        shards.append(g)
    for chunk in (0, 1 << 16):
        csv, st = text_through(shards, c.reads_bytes, c.names, chunk_bytes=chunk)
        assert csv == golden("light_small"), chunk
        assert st["n_reads"] == 20000
    csv, _ = text_through(shards, c.reads_bytes, c.names, chunk_bytes=1 << 18, extended=True)
    with CuClarkDB(c.k, c.n_targets, htsize=c.htsize) as whole:
        whole.load_arrays(sz, ky, lb)
        ref, _ = whole.classify_text(c.reads_bytes, names=c.names, extended=True)
    assert csv == ref
    for g in shards:
        g.close()


def test_text_pipeline_table_partitioned_fastq_k31(full_small):
    c = full_small
    sz, ky, lb = c.arrays
    shards = []
    for i in range(2):
        g = CuClarkDB(c.k, c.n_targets, htsize=c.htsize, shard=(i, 2))
        g.load_arrays(sz, ky, lb)
        shards.append(g)
    csv, st = text_through(shards, c.reads_bytes, c.names, chunk_bytes=1 << 20)
    assert csv == golden("full_small")
    for g in shards:
        g.close()


def test_clone_table_and_read_partitioned_two_devices(light_c1):
    """BASELINE configs[0] through two devices, table loaded once and cloned over NVLink."""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    c = light_c1
    sz, ky, lb = c.arrays
    a = CuClarkDB(c.k, c.n_targets, htsize=c.htsize, device=0)
    a.load_arrays(sz, ky, lb)
    b = CuClarkDB(c.k, c.n_targets, htsize=c.htsize, device=1)
    b.clone_table_from(a)
    assert b.stats()["n_entries"] == a.stats()["n_entries"] == c.kmers.size
    csv, st = text_through([a, b], c.reads_bytes, c.names, chunk_bytes=1 << 18, n_slots=3)
    assert csv == golden("light_c1") and st["n_reads"] == 100000
    csv_b, _ = text_through([b], c.reads_bytes, c.names)                  # the clone alone
    assert csv_b == golden("light_c1")
    a.close(); b.close()


def test_text_pipeline_table_partitioned_two_devices(light_c1):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    c = light_c1
    sz, ky, lb = c.arrays
    shards = []
    for i in range(2):
        g = CuClarkDB(c.k, c.n_targets, htsize=c.htsize, shard=(i, 2), device=i)
        g.load_arrays(sz, ky, lb)
        shards.append(g)
    csv, st = text_through(shards, c.reads_bytes, c.names, chunk_bytes=1 << 20)
    assert csv == golden("light_c1") and st["n_reads"] == 100000
    for g in shards:
        g.close()


def test_cli_default_devices_and_partitioned_table(exes, light_small, tmp_path):
    """cuCLARK-l without -d uses every GPU of the box (src/main.cc:104); CUCLARK_PARTITION_TABLE=1 forces the
    table-partitioned path the CLI takes by itself when the table exceeds one device; -d beyond the box fails
    as the reference does. All CSVs equal the reference binary's."""
    reads = setup_case(light_small, str(tmp_path))
    cmd = [exes[1], "-T", "targets.txt", "-D", "db/", "-O", os.path.basename(reads)]
    ref = golden("light_small")
    p = run(cmd + ["-R", "all"], cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    assert (tmp_path / "all.csv").read_bytes() == ref
    if n_gpus() > 1:
        assert f"Using {n_gpus()} devices: table replicated, reads partitioned." in p.stderr
    p = run(cmd + ["-R", "toomany", "-d", str(n_gpus() + 1)], cwd=str(tmp_path))
    assert p.returncode == 1 and "Not enough CUDA devices found" in p.stderr
    if n_gpus() > 1:
        env = dict(os.environ, CUCLARK_PARTITION_TABLE="1")
        p = subprocess.run(cmd + ["-R", "part", "-d", "2"], cwd=str(tmp_path), capture_output=True, text=True, env=env)
        assert p.returncode == 0, p.stderr
        assert "Using 2 devices: table partitioned by bucket range" in p.stderr
        assert (tmp_path / "part.csv").read_bytes() == ref
        p = subprocess.run(cmd + ["-R", "partx", "-d", "2", "--extended"], cwd=str(tmp_path), capture_output=True, text=True, env=env)
        q = run(cmd + ["-R", "onex", "-d", "1", "--extended"], cwd=str(tmp_path))
        assert p.returncode == 0 and q.returncode == 0, p.stderr + q.stderr
        assert (tmp_path / "partx.csv").read_bytes() == (tmp_path / "onex.csv").read_bytes()
