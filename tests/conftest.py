import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"  # only present in the build container; never required


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def scenes():
    """name -> (env [n,3,3], robot [m,3,3]) float64 from tests/golden/scenes/*.npz"""
    out = {}
    for f in sorted(os.listdir(os.path.join(GOLDEN, "scenes"))):
        d = np.load(os.path.join(GOLDEN, "scenes", f))
        out[f[:-4]] = (d["env"], d["robot"] if "robot" in d.files else None)
    return out


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "se3_verdicts.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py
