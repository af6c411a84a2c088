import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def random_graph(n, kmin, kmax, seed, zero_frac=0.0, dup_rows=0):
    """Ragged random neighbour graph: rows sorted ascending, no self edges, distinct neighbours per row."""
    rng = np.random.Generator(np.random.PCG64(seed))
    row_ptr = [0]
    cols, dists = [], []
    for i in range(n):
        k = int(rng.integers(kmin, kmax + 1))
        cand = rng.choice(n - 1, size=k, replace=False)
        cand = np.where(cand >= i, cand + 1, cand)
        d = np.sort(rng.gamma(2.0, 1.0, size=k)).astype(np.float32)
        if zero_frac > 0 and rng.random() < zero_frac:
            d[: int(rng.integers(1, k + 1))] = 0.0
        cols.append(cand)
        dists.append(d)
        row_ptr.append(row_ptr[-1] + k)
    for r in range(dup_rows):                    # all-equal rows
        i = int(rng.integers(0, n))
        dists[i][:] = dists[i][0]
    return (np.array(row_ptr, np.uint64), np.concatenate(cols).astype(np.uint32),
            np.concatenate(dists).astype(np.float32))


@pytest.fixture(scope="session")
def small_graph():
    return random_graph(300, 3, 12, seed=7, zero_frac=0.05, dup_rows=5)
