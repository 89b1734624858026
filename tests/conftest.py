"""pytest configuration: `-m gpu` tests need a B200 (they call the CUDA library through its
C-ABI and fail loudly if it is missing); everything else runs on CPU."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENARIOS = ["basic", "block", "filter", "waterfall", "weird-edges"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """The reference-parity suites on the shipped scenarios first; then the random-scenario fuzz,
    then the opt-in modes that are not in the reference (mixed-precision PCG) — so that `-x`
    reports a failure of the core suites before anything else."""
    def rank(item):
        module = item.nodeid.split("::")[0]
        return 2 if "mixed" in module else 1 if "fuzz" in module else 0
    items.sort(key=rank)


@pytest.fixture(scope="session")
def known_answers():
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        return json.load(f)


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    if a.dtype == np.float64:
        return a.view(np.uint64)
    return a


def same_bits(a, b):
    """Bit-for-bit equality (NaNs with equal payload compare equal; +0 != -0)."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(bits(a), bits(b))


def load_state(name):
    """Reference state dumped by tests/golden/make_golden.py."""
    return np.load(os.path.join(GOLDEN, "state_%s_f10.npz" % name))
