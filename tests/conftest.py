import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


def grad_from_seed(seed, shape):
    """The upstream gradient make_golden.py drew for the 640x480 cases."""
    return np.random.default_rng(int(seed)).standard_normal(tuple(shape)).astype(np.float32)


# Tolerance of BASELINE.json north_star: 1e-5 relative / 1e-6 absolute, fp32.
RTOL, ATOL = 1e-5, 1e-6


def assert_close(actual, expected, what=""):
    actual, expected = np.asarray(actual), np.asarray(expected)
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    bad = ~(np.abs(actual - expected) <= ATOL + RTOL * np.abs(expected))
    bad &= ~(np.isnan(actual) & np.isnan(expected))
    assert not bad.any(), "%s: %d of %d outside 1e-6+1e-5*|ref| (max abs err %g)" % (
        what, bad.sum(), bad.size, np.nanmax(np.abs(actual - expected)))


def assert_bits(actual, expected, what=""):
    actual, expected = np.asarray(actual), np.asarray(expected)
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    assert actual.dtype == expected.dtype, (what, actual.dtype, expected.dtype)
    same = np.array_equal(actual, expected, equal_nan=actual.dtype.kind == "f")
    if not same:
        diff = actual != expected
        raise AssertionError("%s: %d of %d elements differ" % (what, diff.sum(), diff.size))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()
    return o
