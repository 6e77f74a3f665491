import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["tg12", "sod48", "chcyl", "thermal", "pressure", "slip", "shear"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.build()
    return O


def load_golden(name):
    import numpy as np
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    return z, str(z["deck"]), [int(s) for s in z["steps"]]
