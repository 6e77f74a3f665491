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


def test_slab_is_fluid_takes_neighbour_and_periodic_planes():
    """multi-rank decks: the ghost planes of a z-slab hold the neighbouring ranks' cells and, at the domain ends of a
    periodic direction, the periodic image -- not the body evaluated beyond the domain (m_is_fluid.FillBoundary)"""
    import numpy as np
    from conftest import load_golden
    from marbles_b200.inputs import lbm_inputs, parse_deck
    from marbles_b200.lbm import slab_bounds, slab_is_fluid_from_deck, wrap_periodic
    _, deck_text, _ = load_golden("chcyl")  # cylinder along the periodic z direction: solid cells cross every slab cut
    inp = lbm_inputs(parse_deck(text=deck_text))
    ng, nz = 3, inp.n_cell[2]
    whole = slab_is_fluid_from_deck(inp, 0, nz - 1, ng)
    assert (whole == 0).any()
    valid = whole[ng:-ng, ng:-ng, ng:-ng]
    for world in (2, 4):
        for rank in range(world):
            lo, hi = slab_bounds(nz, rank, world)
            a = slab_is_fluid_from_deck(inp, lo, hi, ng)
            assert a.shape == (hi - lo + 1 + 2 * ng,) + whole.shape[1:]
            for k in range(lo - ng, hi + ng + 1):  # every plane, ghost planes included, is the wrapped global plane
                assert np.array_equal(a[k - lo + ng, ng:-ng, ng:-ng], valid[k % nz]), (world, rank, k)
    # a non-periodic direction keeps the geometry beyond the domain
    w = wrap_periodic(np.arange(5 * 5 * 5).reshape(5, 5, 5), 1, (3, 3, 3), (1, 0, 0))
    assert w[2, 2, 0] == w[2, 2, 3] and w[0, 2, 2] == 2 + 2 * 5


def test_invalid_bc_code_is_refused():
    import pytest
    from conftest import load_golden
    from marbles_b200.inputs import DeckError, lbm_inputs, parse_deck
    _, deck_text, _ = load_golden("chcyl")
    with pytest.raises(DeckError, match="Invalid bc"):
        lbm_inputs(parse_deck(text=deck_text, overrides=["lbm.bc_lo = 4 1 0"]))
