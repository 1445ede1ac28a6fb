"""Shared fixtures. GPU tests are marked `gpu`; everything else runs on CPU."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import Oracle
    return Oracle()


def load_case(name: str) -> dict:
    with open(os.path.join(GOLDEN, f"db_{name}.json")) as f:
        return json.load(f)


class CaseData:
    """Seeded inputs of a golden case + its database arrays rebuilt with oracle/dbtools."""

    def __init__(self, name: str):
        sys.path.insert(0, GOLDEN)
        import make_golden
        from oracle import dbtools
        from oracle.binding import HTSIZE_FULL, HTSIZE_LIGHT, key_bytes_for
        self.name = name
        self.golden = load_case(name)
        self.case = c = self.golden["case"]
        self.light, self.k, self.n_targets = c["light"], c["k"], c["n_targets"]
        self.htsize = HTSIZE_LIGHT if self.light else HTSIZE_FULL
        self.key_bytes = key_bytes_for(self.k, self.htsize)
        self.maxhits = 23 if self.light else 15
        self.names = [make_golden.target_name(t) for t in range(self.n_targets)]
        targets = [make_golden.target_codes(c, t) for t in range(self.n_targets)]
        self.kmers, self.labels = dbtools.build_entries(targets, self.k, c["gap"])
        self._arrays = None
        self._reads = None
        self._mg = make_golden

    @property
    def arrays(self):
        if self._arrays is None:
            from oracle import dbtools
            self._arrays = dbtools.entries_to_arrays(self.kmers, self.labels, self.htsize, self.key_bytes)
        return self._arrays

    @property
    def reads_bytes(self) -> bytes:
        if self._reads is None:
            self._reads = self._mg.make_reads(self.case)
        return self._reads


_cases = {}


def get_case(name: str) -> CaseData:
    if name not in _cases:
        _cases[name] = CaseData(name)
    return _cases[name]


@pytest.fixture(scope="session")
def light_small():
    return get_case("light_small")


@pytest.fixture(scope="session")
def light_c1():
    return get_case("light_c1")


@pytest.fixture(scope="session")
def full_small():
    return get_case("full_small")
