"""BASELINE config 2, the thermal Sod shock tube (Tools/sod_test of the reference): density of the f+g lattice run
against the exact solution of the Riemann problem (rho 0.5 | 2.0, T 0.2 | 0.025, gamma = 2, R = 1).  The reference's
arithmetic (oracle, CPU) stays within 1 % in L1 at 400 cells; the CUDA path must reproduce the oracle."""
import numpy as np
import pytest

from conftest import load_golden

N, T = 400, 130


def exact_riemann(rl, ul, pl, rr, ur, pr, g, xi):
    """exact solution of the Riemann problem for a polytropic gas at similarity coordinates xi = (x - x0) / t (Toro)"""
    al, ar = np.sqrt(g*pl/rl), np.sqrt(g*pr/rr)
    def fk(p, rk, pk, ak):
        if p > pk:
            A, B = 2/((g+1)*rk), (g-1)/(g+1)*pk
            return (p-pk)*np.sqrt(A/(p+B)), np.sqrt(A/(p+B))*(1-(p-pk)/(2*(B+p)))
        return 2*ak/(g-1)*((p/pk)**((g-1)/(2*g))-1), 1/(rk*ak)*(p/pk)**(-(g+1)/(2*g))
    p = 0.5*(pl+pr)
    for _ in range(100):
        f1, d1 = fk(p, rl, pl, al); f2, d2 = fk(p, rr, pr, ar)
        dp = (f1+f2+ur-ul)/(d1+d2); p = max(p-dp, 1e-12)
        if abs(dp) < 1e-14*p: break
    f1,_ = fk(p, rl, pl, al); f2,_ = fk(p, rr, pr, ar)
    us = 0.5*(ul+ur)+0.5*(f2-f1)
    rho = np.empty_like(xi)
    for n, s in enumerate(xi):
        if s <= us:  # left of contact
            if p > pl:
                sl = ul - al*np.sqrt((g+1)/(2*g)*p/pl+(g-1)/(2*g))
                rho[n] = rl if s < sl else rl*((p/pl+(g-1)/(g+1))/((g-1)/(g+1)*p/pl+1))
            else:
                shl, stl = ul-al, us-al*(p/pl)**((g-1)/(2*g))
                if s < shl: rho[n] = rl
                elif s > stl: rho[n] = rl*(p/pl)**(1/g)
                else: rho[n] = rl*(2/(g+1)+(g-1)/((g+1)*al)*(ul-s))**(2/(g-1))
        else:
            if p > pr:
                sr = ur + ar*np.sqrt((g+1)/(2*g)*p/pr+(g-1)/(2*g))
                rho[n] = rr if s > sr else rr*((p/pr+(g-1)/(g+1))/((g-1)/(g+1)*p/pr+1))
            else:
                shr, stR = ur+ar, us+ar*(p/pr)**((g-1)/(2*g))
                if s > shr: rho[n] = rr
                elif s < stR: rho[n] = rr*(p/pr)**(1/g)
                else: rho[n] = rr*(2/(g+1)-(g-1)/((g+1)*ar)*(ur-s))**(2/(g-1))
    return rho, p, us



def sod_deck():
    _, deck_text, _ = load_golden("sod48")
    return deck_text, [f"amr.n_cell = {N} 2 2", f"geometry.prob_hi = {N}.0 1.0 1.0", f"ic_sod.x_discontinuity = {N / 2}",
                       "amr.max_grid_size = 1024"]


def l1_error(rho):
    x = np.arange(N) + 0.5
    ex, _, _ = exact_riemann(0.5, 0.0, 0.5 * 0.2, 2.0, 0.0, 2.0 * 0.025, 2.0, (x - N / 2) / T)
    return float(np.abs(rho - ex).sum() / np.abs(ex).sum())


def test_oracle_sod_density_vs_exact_riemann(oracle_mod):
    O = oracle_mod
    deck_text, ov = sod_deck()
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines() + ov)))
    o.initialize()
    o.step(T)
    err = l1_error(o.fields()["rho"][0, 0, :])
    assert err <= 0.01, err


@pytest.mark.gpu
def test_gpu_sod_density_vs_exact_riemann_and_oracle(oracle_mod):
    from parity import compare, scales
    from marbles_b200.inputs import parse_deck
    from marbles_b200.lbm import LBM
    O = oracle_mod
    deck_text, ov = sod_deck()
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines() + ov)))
    o.initialize()
    o.step(T)
    ref = o.fields()
    lbm = LBM(parse_deck(text=deck_text, overrides=ov))
    lbm.init_data()
    lbm.step(T, want_macrodata=True)
    got = lbm.fields()
    err = l1_error(got["rho"][0, 0, :])
    worst, key = compare(got, ref, scales(ref, lbm.inp.R, lbm.inp.gamma, 1.0), T)
    print(f"Sod tube {N} cells, {T} steps: L1 density error vs exact Riemann {err:.4f}; vs oracle {worst:.2e} ({key})")
    assert err <= 0.01
    lbm.close()
