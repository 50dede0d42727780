import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (C restatement of the reference), built on demand with gcc."""
    from oracle import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def S():
    """The product package.  The CUDA library is normally built by __graft_entry__.build(); if it is missing
    (fresh clone) it is compiled here with nvcc, which cross-compiles for sm_100a without a GPU."""
    import swimmer_ars_b200
    if not os.path.exists(swimmer_ars_b200._lib.LIB_PATH):
        swimmer_ars_b200.build_library()
    return swimmer_ars_b200


def golden(name):
    path = os.path.join(GOLDEN, name)
    return np.load(path, allow_pickle=False)


def rel_err(a, b):
    """norm-relative error with an absolute floor of 1 (near-zero quantities)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))))


def rand_states(rng, n, B, scale=3.0):
    st = np.empty((B, 2 * n + 2))
    st[:, :2] = rng.normal(size=(B, 2))
    st[:, 2::2] = rng.uniform(-4, 4, (B, n))
    st[:, 3::2] = rng.normal(size=(B, n)) * scale
    return st
