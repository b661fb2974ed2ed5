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


@pytest.fixture(scope="session")
def golden_test1():
    return dict(np.load(os.path.join(GOLDEN, "test1_operations.npz")))


@pytest.fixture(scope="session")
def golden_ref():
    return dict(np.load(os.path.join(GOLDEN, "reference_kernels.npz")))


# the reference's own golden vectors (hard-coded in its tests / docs)
SIX_BY_THREE = dict(  # src/test/cscs_to_csr_test.py:13-25, csc.py:53-87
    m=6, n=3,
    data=np.array([4, 3, 3, 9, 7, 8, 4, 8, 8, 9], dtype=np.float64),
    indices=np.array([0, 1, 3, 1, 2, 4, 5, 2, 3, 4], dtype=np.int32),
    indptr=np.array([0, 3, 7, 10], dtype=np.int32),
    csr_data=np.array([4, 3, 9, 7, 8, 3, 8, 8, 9, 4], dtype=np.float64),
    csr_indices=np.array([0, 0, 1, 1, 2, 0, 2, 1, 2, 1], dtype=np.int32),
    csr_indptr=np.array([0, 1, 3, 5, 7, 9, 10], dtype=np.int32))
CONNECTIVITY = dict(  # docs/connectivity_matrix.rst:23-27, 93-105
    m=3, n=5,
    indptr=np.array([0, 0, 1, 2, 2, 3], dtype=np.int32),
    indices=np.array([0, 1, 2], dtype=np.int32),
    data=np.array([1.0, 1.0, 1.0]),
    p=np.array([10.0, 20.0, 30.0]),
    injections=np.array([0.0, 10.0, 20.0, 0.0, 30.0]))


def sort_columns(n, Cp, Ci, Cx):
    """Canonical form for comparing products whose in-column order is unspecified."""
    Ci, Cx = np.array(Ci), np.array(Cx)
    for j in range(n):
        s, e = Cp[j], Cp[j + 1]
        o = np.argsort(Ci[s:e], kind="stable")
        Ci[s:e] = Ci[s:e][o]
        Cx[s:e] = Cx[s:e][o]
    return Ci, Cx


@pytest.fixture(scope="session")
def golden_helpers():
    return dict(np.load(os.path.join(GOLDEN, "reference_helpers.npz")))
