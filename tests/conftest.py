import glob
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle, build_oracle
    build_oracle()
    return Oracle()


def skewed_matrix(rng, nRow, nCol, density_rows):
    """Random sorted duplicate-free COO with empty rows, very long rows and short rows."""
    rows, cols = [], []
    for r in range(nRow):
        kind = rng.random()
        if kind < 0.2:
            n = 0
        elif kind < 0.25:
            n = int(nCol * 0.9)
        else:
            n = int(rng.integers(1, max(2, density_rows)))
        c = np.sort(rng.choice(nCol, size=min(n, nCol), replace=False))
        rows.append(np.full(len(c), r))
        cols.append(c)
    row = np.concatenate(rows).astype(np.int32)
    col = np.concatenate(cols).astype(np.int32)
    val = rng.standard_normal(len(row))
    return row, col, val
