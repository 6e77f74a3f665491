import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["tg12", "sod48", "chcyl", "thermal", "pressure", "slip", "shear"]


AMR_GOLDEN_CASES = ["amr2_tg", "amr2_chcyl", "amr3_chcyl", "amr2_sod"]


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


def load_amr_golden(name):
    """-> (npz, deck text, stored steps, boxes[lev] = [(lo, hi), ...], is_fluid[lev] dense int32)"""
    import numpy as np
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    nlev = int(z["nlev"])
    boxes = [[(list(map(int, b[0])), list(map(int, b[1]))) for b in z[f"boxes_l{l}"]] for l in range(nlev)]
    is_fluid = [z[f"is_fluid_l{l}"].astype(np.int32) for l in range(nlev)]
    return z, str(z["deck"]), [int(s) for s in z["steps"]], boxes, is_fluid
