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
def orc():
    import oracle
    return oracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own compiled code (oracle/_ref); skip where it was never built."""
    import oracle
    if not oracle.Reference.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle.Reference()


@pytest.fixture(scope="session")
def golden_csrmv():
    return np.load(os.path.join(GOLDEN, "merge_csrmv_ref.npz"))


def random_csr(rng, rows, cols, mean_len, empty_frac=0.0, long_rows=0):
    lens = rng.poisson(mean_len, rows)
    lens[rng.random(rows) < empty_frac] = 0
    for _ in range(long_rows):
        lens[rng.integers(rows)] = rng.integers(max(cols // 2, 1), cols + 1)
    lens = np.minimum(lens, cols)
    ro = np.zeros(rows + 1, np.int32)
    ro[1:] = np.cumsum(lens)
    nnz = int(ro[-1])
    col = np.empty(nnz, np.int32)
    for r in range(rows):
        if lens[r]:
            col[ro[r]:ro[r + 1]] = np.sort(rng.choice(cols, lens[r], replace=False))
    return ro, col
