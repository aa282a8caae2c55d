import sys
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """-> dict with the fixture arrays plus ``csr`` (float64) and ``chr_pos``."""
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    out = {k: z[k] for k in z.files}
    out["csr"] = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(int(v) for v in z["shape"]))
    out["chr_pos"] = {str(k): int(v) for k, v in zip(z["chr_names"], z["chr_offsets"])}
    return out


@pytest.fixture(scope="session")
def golden_loader():
    return load_golden


@pytest.fixture()
def full_mock():
    """The reference's seeded 4 cells x 10 genes fixture (/root/reference/tests/conftest.py:61-75)."""
    import pandas as pd

    np.random.seed(0)
    genes = [f"gene{i}" for i in range(1, 11)]
    var = pd.DataFrame(
        {
            "start": [100, 200, 300, 400, 500, 0, 100, 200, 300, 400],
            "end": [199, 299, 399, 499, 599, 99, 199, 299, 399, 499],
            "chromosome": ["chr1"] * 5 + ["chr2"] * 5,
        },
        index=genes,
    )
    X = np.random.randint(low=0, high=50, size=(4, 10))
    return X, var


# /root/reference/tests/conftest.py:98-108
X_RES_ACTUAL = np.array(
    [
        [1.00, 0.00, 0.00, 0.00, 0.00, 1.00],
        [-1.00, 0.00, 0.00, 0.00, 0.00, 0.00],
        [0.00, 1.25, 1.25, 0.00, 0.00, 0.00],
        [0.00, 0.00, 0.00, 0.875, 0.00, 0.00],
    ]
)


@pytest.fixture()
def x_res_actual():
    return X_RES_ACTUAL.copy()


# /root/reference/tests/conftest.py:78-95 (columns gene1..gene10, rows = the 4 cells)
GENE_RES_ACTUAL = np.array(
    [
        [0.75, 0.00, 0.000000, 0.00, -0.75, 0.000000, 0.000000, 0.0, 0.0, 0.75],
        [-1.00, 0.00, 0.000000, 0.00, 0.00, 0.000000, 0.000000, 0.0, 0.0, 0.00],
        [0.00, 0.75, 0.91666667, 1.25, 1.25, 0.000000, 0.000000, 0.0, 0.0, 0.00],
        [0.00, 0.00, 0.000000, 0.00, 0.00, 0.921875, 0.703125, 0.0, 0.0, 0.00],
    ]
)


@pytest.fixture()
def gene_res_actual():
    return GENE_RES_ACTUAL.copy()
