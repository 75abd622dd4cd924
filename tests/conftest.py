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


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def g_distance():
    return load_golden("distance")


@pytest.fixture(scope="session")
def g_synthetic():
    return load_golden("synthetic")


@pytest.fixture(scope="session")
def g_knntest():
    return load_golden("knntest")


@pytest.fixture(scope="session")
def g_library():
    return load_golden("library")


def assert_csr_equal(got, g, prefix, weights_exact=True, check_idx=True):
    """Compare a list of (idx, w) tuples (or a CSR triple) with a golden CSR record."""
    from oracle.prograph_oracle import to_csr
    indptr, idx, w = got if isinstance(got, tuple) and len(got) == 3 and not isinstance(got[0], tuple) else to_csr(got)
    np.testing.assert_array_equal(indptr, g[prefix + "_indptr"])
    if check_idx:
        np.testing.assert_array_equal(idx, g[prefix + "_idx"])
        assert str(idx.dtype) == "int64"
    gw = g[prefix + "_w"]
    if len(gw):
        assert str(np.asarray(w).dtype) == str(g[prefix + "_w_dtype"]), (prefix, np.asarray(w).dtype, g[prefix + "_w_dtype"])
    if weights_exact:
        np.testing.assert_array_equal(np.asarray(w), gw)
