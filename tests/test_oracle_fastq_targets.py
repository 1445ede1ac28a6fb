"""CPU: FASTQ target files through the UNMODIFIED reference binary (oracle/_ref/cuCLARK-l builds its database on the host
before it looks for a GPU) against the restatement the GPU builder's test uses (tests/test_gpu_dbbuild.py:
fastq_as_fasta + scan_target + expected_db). Pins the FASTQ-target semantics of src/CuCLARK_hh.hh:769-860."""
import os
import subprocess

import numpy as np
import pytest

from oracle import dbtools
from oracle.binding import HTSIZE_LIGHT, key_bytes_for

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "cuCLARK-l")


def test_reference_binary_fastq_targets(tmp_path):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from test_gpu_dbbuild import expected_db, fastq_as_fasta, fastq_targets
    datas, labels = fastq_targets()
    names = ["L0", "L1", "L2"]
    (tmp_path / "db").mkdir()
    with open(tmp_path / "targets.txt", "w") as tf:
        for i, d in enumerate(datas):
            p = tmp_path / f"t{i}.fq"
            p.write_bytes(d)
            tf.write(f"{p} {names[labels[i]]}\n")
    (tmp_path / "reads.fa").write_bytes(b">r\nACGT\n")
    # builds db_central_k27_t3_s57777779_m0_light_4.tsk.{sz,ky,lb}, then stops at the CUDA device check
    subprocess.run([REF, "-T", "targets.txt", "-D", "db/", "-O", "reads.fa", "-R", "out"], cwd=str(tmp_path),
                   capture_output=True, text=True, timeout=600)
    base = dbtools.db_name(str(tmp_path / "db"), 27, 3, HTSIZE_LIGHT, 0, 4)
    assert os.path.exists(base + ".ky"), "the reference did not write a database"
    kb = key_bytes_for(27, HTSIZE_LIGHT)
    gsz, gky, glb = dbtools.read_db_files(base, HTSIZE_LIGHT, kb)
    as_fasta = [fastq_as_fasta(d) if d[:1] == b"@" else d for d in datas]
    sz, ky, lb = expected_db(as_fasta, labels, 27, 4, HTSIZE_LIGHT, kb)
    assert ky.size > 50
    assert np.array_equal(gsz, sz) and np.array_equal(gky, ky) and np.array_equal(glb, lb)
