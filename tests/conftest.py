"""pytest configuration: the `gpu` marker, fixture loading and the C-ABI library handle."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped, not failed, when no device is visible (CPU container)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ansatz_kats():
    meta = json.load(open(os.path.join(GOLDEN, "ansatz_kats.json")))
    arrs = np.load(os.path.join(GOLDEN, "ansatz_kats.npz"))
    return meta, arrs


@pytest.fixture(scope="session")
def gatelist_kats():
    meta = json.load(open(os.path.join(GOLDEN, "gatelist_kats.json")))
    arrs = np.load(os.path.join(GOLDEN, "gatelist_kats.npz"))
    return meta, arrs


@pytest.fixture(scope="session")
def trials():
    return json.load(open(os.path.join(GOLDEN, "trials.json")))


@pytest.fixture(scope="session")
def lib():
    """The C-ABI CUDA library, built in-tree if it is not there yet (nvcc cross-compiles on CPU)."""
    from cpflow_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def hst(u, v):
    """cost_HST in numpy (matrix_utils.py:35-42): global-phase-invariant distance."""
    n = u.shape[0]
    return 1 - abs(np.sum(u * np.conj(v))) ** 2 / n ** 2
