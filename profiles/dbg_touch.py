import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, golden_is_fluid
from marbles_b200.inputs import parse_deck
from marbles_b200.lbm import LBM
from oracle import oracle as O
z, deck_text, steps = load_golden("touch")
fl = golden_is_fluid("touch")
o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines())), is_fluid=fl); o.initialize()
for variant in (0, 5):
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines())), is_fluid=fl); o.initialize()
    lbm = LBM(parse_deck(text=deck_text), is_fluid=fl, variant=variant); lbm.init_data()
    for s in range(1, 7):
        o.step(1); lbm.step(1, want_macrodata=True)
        f, g, m = lbm.get_f(), lbm.get_g(), lbm.get_macrodata()
        df = np.abs(f - o.f_valid); dg = np.abs(g - o.g_valid); dm = np.abs(m - o.macro_valid)
        i = np.unravel_index(np.argmax(df), df.shape); j = np.unravel_index(np.argmax(dm[18]), dm[18].shape)
        print(f"v{variant} step {s}: max|df| {df.max():.3e} at q,k,j,i={i} (fl {fl[3:-3,3:-3,3:-3][i[1:]]}), max|dg| {dg.max():.3e}, "
              f"max|dT| {dm[18].max():.3e} at {j}, cells with |df|>1e-12: {(df.max(axis=0) > 1e-12).sum()}")
        if s in (2, 6):
            bad = np.argwhere(df.max(axis=0) > 1e-12)
            print("   bad cells (k,j,i):", bad[:24].tolist())
    lbm.close()
