"""Host-side geometry and input-deck logic on CPU: EB flags of the shipped body types against the reference's own
is_fluid (golden plotfiles), the eb_boundary flag, and the deck quantities the kernels are configured with."""
import numpy as np
import pytest

from conftest import load_golden
from marbles_b200.geometry import is_fluid_from_deck
from marbles_b200.inputs import lbm_inputs, parse_deck
from marbles_b200.plotfile import eb_boundary


@pytest.mark.parametrize("case", ["chcyl", "pressure", "slip", "tg12", "sod48"])
def test_is_fluid_matches_reference(case):
    z, deck_text, steps = load_golden(case)
    deck = parse_deck(text=deck_text)
    inp = lbm_inputs(deck)
    a = is_fluid_from_deck(deck, inp.n_cell, inp.prob_lo, inp.dx, ng=3)
    assert np.array_equal(a[3:-3, 3:-3, 3:-3], z["is_fluid"].astype(np.int32))
    key = f"s{steps[-1]}_eb_boundary"
    if key in z.files:  # component 1 of m_is_fluid as the reference plots it
        assert np.array_equal(eb_boundary(a, 3), z[key].astype(np.int32))


def test_deck_quantities():
    z, deck_text, _ = load_golden("chcyl")
    inp = lbm_inputs(parse_deck(text=deck_text))
    assert tuple(inp.n_cell) == (32, 12, 4) and tuple(inp.periodic) == (0, 0, 1)
    assert inp.bc_lo[0] == 2 and inp.bc_hi[0] == 5 and inp.bc_lo[1] == 1  # velocity inlet, outflow, no-slip
    assert inp.dx[0] == pytest.approx((inp.prob_hi[0] - inp.prob_lo[0]) / 32)
    ov = parse_deck(text=deck_text, overrides=["amr.n_cell = 64 24 8", "lbm.nu=0.02"])
    inp2 = lbm_inputs(ov)
    assert tuple(inp2.n_cell) == (64, 24, 8) and inp2.nu == pytest.approx(0.02)
