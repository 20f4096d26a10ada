"""Shared fixtures.  `-m "not gpu"` runs on the CPU-only authoring box; `-m gpu` on a B200."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLDEN, "golden_meta.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_arrays():
    return np.load(os.path.join(GOLDEN, "golden_arrays.npz"))


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): oracle/oracle.py over oracle/libmandel_oracle.so."""
    import oracle as _oracle
    _oracle.lib()
    return _oracle


@pytest.fixture(scope="session")
def native():
    from distributedmandelbrot_b200 import _native
    _native.lib()
    return _native


@pytest.fixture(scope="session")
def gpu(native):
    """Binds device 0.  On a box without CUDA the gpu tests are skipped; on a GPU box any
    failure to load/initialise the CUDA library is an ERROR, never a silent fallback."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device on this box")
    native.init(0)
    return native


def load_fullsize(name: str):
    """tests/golden/fullsize_<name>.json: per-chunk digests of a BASELINE frame at its full size,
    produced by the CPU oracle alone (oracle/gen_fullsize_golden.py)."""
    path = os.path.join(GOLDEN, f"fullsize_{name}.json")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    with open(path) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fullsize():
    return load_fullsize
