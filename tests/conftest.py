import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["tg12", "sod48", "chcyl", "thermal", "pressure", "slip", "shear", "slipyz", "touch"]


AMR_GOLDEN_CASES = ["amr2_tg", "amr2_chcyl", "amr3_chcyl", "amr2_sod", "amr2_sod_regrid", "amr2_tg_appear", "amr2_chcyl_appear", "amr2_sod_bc"]


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


def golden_is_fluid(name):
    """is_fluid (component 0) for a single-level golden case: the reference's valid-cell field, except where the body
    crosses a domain face (`touch`): there the 3 ghost layers carry the geometry evaluated beyond the domain, which a
    plotfile does not hold -- taken from the analytic body of the deck (checked against the valid cells)."""
    import numpy as np
    z, deck_text, _ = load_golden(name)
    fl = z["is_fluid"].astype(np.int32)
    if name != "touch":
        return fl
    from marbles_b200.geometry import is_fluid_from_deck
    from marbles_b200.inputs import lbm_inputs, parse_deck
    inp = lbm_inputs(parse_deck(text=deck_text))
    grown = is_fluid_from_deck(inp.deck, inp.n_cell, inp.prob_lo, inp.dx, ng=3)
    assert np.array_equal(grown[3:-3, 3:-3, 3:-3], fl)
    return grown


def amr_boxes_at(z, step, lev):
    """box list of level lev during coarse step `step` of a dynamically regridded golden case (None: unchanged)"""
    key = f"boxes_s{step}_l{lev}"
    if key not in z.files:
        return None
    return [(list(map(int, b[0])), list(map(int, b[1]))) for b in z[key]]


def amr_regrid_actions(z, step, current):
    """What AmrCore::regrid did at the start of coarse step `step` of a dynamically regridded golden case, read off the
    box lists the reference wrote: [(lev, "make" | "remake" | "clear", boxes)] -- levels made (MakeNewLevelFromCoarse)
    or re-made (RemakeLevel) from the coarsest up, then cleared (ClearLevel) from the finest down.
    `current` = {lev: box list now, [] if the level does not exist} is updated."""
    acts, gone = [], []
    for lev in sorted(current):
        nb = amr_boxes_at(z, step, lev)
        if nb is None or nb == current[lev]:
            continue
        if not nb:
            gone.append((lev, "clear", nb))
        else:
            acts.append((lev, "remake" if current[lev] else "make", nb))
        current[lev] = nb
    return acts + gone[::-1]
