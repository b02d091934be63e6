"""pytest configuration: registers the ``gpu`` marker and puts the product package on sys.path.

``-m "not gpu"`` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbol export.
``-m gpu`` runs on a B200: the parity tests proper, all through the C-ABI library.
"""

import gzip
import json
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
PACKAGE_DIR = REPO / "fvdb-core_b200"
for p in (str(REPO), str(PACKAGE_DIR)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def semantics_golden():
    with gzip.open(REPO / "tests" / "golden" / "semantics_golden.json.gz", "rt") as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def dense_golden():
    data = np.load(REPO / "tests" / "golden" / "dense_golden.npz")
    meta = json.loads(bytes(data["meta"]).decode())
    return data, meta
