"""The oracle (oracle/marbles_oracle.c) against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  The restatement keeps the reference's operation order,
so the comparison is BIT-EXACT on every plotted field of every stored step."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, golden_is_fluid, load_golden


def make_oracle(O, deck_text, is_fluid=None):
    deck = O.parse_deck(None, deck_text.splitlines())
    o = O.Oracle(O.lbm_setup(deck), is_fluid=is_fluid)
    o.initialize()
    return o


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_bit_exact_vs_reference_golden(oracle_mod, case):
    O = oracle_mod
    z, deck_text, steps = load_golden(case)
    o = make_oracle(O, deck_text, golden_is_fluid(case))
    done = 0
    for s in steps:
        o.step(s - done)
        done = s
        mine = o.fields()
        keys = [k[len(f"s{s}_"):] for k in z.files if k.startswith(f"s{s}_")]
        assert len(keys) >= 73
        for k in keys:
            ref = z[f"s{s}_{k}"]
            assert mine[k].shape == ref.shape
            assert np.array_equal(mine[k], ref), (case, s, k, float(np.abs(mine[k] - ref).max()))


def test_stencil_tables(oracle_mod):
    """check_stencil() invariants (Source/Stencil.cpp:5-62) + mirror tables."""
    O = oracle_mod
    assert O.lib().orc_check_stencil() == 0
    ev, w, opp, mx, my, mz = O.stencil()
    assert abs(w.sum() - 1.0) < 1e-15
    for q in range(27):
        assert (ev[opp[q]] == -ev[q]).all()
        assert (ev[mx[q]] == ev[q] * [-1, 1, 1]).all()
        assert (ev[my[q]] == ev[q] * [1, -1, 1]).all()
        assert (ev[mz[q]] == ev[q] * [1, 1, -1]).all()
    # second moments of the weights: sum w e_a e_b = theta0 delta_ab
    m2 = np.einsum("q,qa,qb->ab", w, ev, ev)
    assert np.allclose(m2, np.eye(3) / 3.0, atol=1e-15)


def test_solid_cells_hold_sentinel(oracle_mod):
    """solid cells are 0 at t=0 and -1 after the first stream."""
    O = oracle_mod
    z, deck_text, _ = load_golden("chcyl")
    fl = z["is_fluid"].astype(np.int32)
    o = make_oracle(O, deck_text, fl)
    assert (o.f_valid[:, fl == 0] == 0.0).all()
    o.step(1)
    assert (o.f_valid[:, fl == 0] == -1.0).all() and (o.g_valid[:, fl == 0] == -1.0).all()
    assert (o.macro_valid[:, fl == 0] == 0.0).all()


def test_mass_and_energy_conserved_periodic(oracle_mod):
    """BGK collision conserves sum f and sum g; periodic streaming permutes them."""
    O = oracle_mod
    _, deck_text, _ = load_golden("tg12")
    o = make_oracle(O, deck_text)
    m0, e0 = o.f_valid.sum(), o.g_valid.sum()
    o.step(5)
    assert abs(o.f_valid.sum() - m0) < 1e-10 * m0
    assert abs(o.g_valid.sum() - e0) < 1e-10 * e0
