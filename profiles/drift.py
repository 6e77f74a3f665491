#!/usr/bin/env python
"""Drift of the CUDA path against the CPU restatement of the reference (oracle, bit-identical to the reference
on the golden decks) over many steps: worst |difference| of every plotfile field relative to its scale
(tests/parity.py) after 1, 10, 30, 100 and 300 steps, for the default step (tile carry) and the two-kernel step, on
  * periodic Taylor-Green 32^3 (omega ~ 1),
  * the thermal-diffusivity wave (`thermal` golden deck: Pr != 1, R != 1, omega ~ 1.994 -- SURVEY App. C warns the
    round-off grows fastest next to the over-relaxation limit),
  * the thermal Sod tube (`sod48` golden deck: gamma = 2, outflow faces).
The stated tolerance is 1e-12 per step.  One JSON line per (deck, variant)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from marbles_b200.inputs import parse_deck  # noqa: E402
from marbles_b200.lbm import LBM  # noqa: E402
from oracle import oracle as O  # noqa: E402
from parity import compare, scales  # noqa: E402

marks = [1, 10, 30, 100, 300]
DECKS = [("Taylor-Green 32^3 periodic", "tg12", ["amr.n_cell = 32 32 32"]),
         ("thermal-diffusivity wave 4x24x2 periodic (nu 8.1e-6, alpha 1e-5, omega ~ 1.994)", "thermal", []),
         ("thermal Sod tube 48x2x2, outflow in x", "sod48", [])]
for title, case, ov in DECKS:
    z = np.load(os.path.join(ROOT, "tests", "golden", f"{case}.npz"))
    deck_text = str(z["deck"])
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines() + ov)))
    o.initialize()
    gpus = {name: LBM(parse_deck(text=deck_text, overrides=ov), variant=v)
            for name, v in (("tile carry (default)", 5), ("two kernels", 0))}
    for g in gpus.values():
        g.init_data()
    rows = {name: {} for name in gpus}
    done = 0
    for n in marks:
        o.step(n - done)
        ref = o.fields()
        for name, g in gpus.items():
            g.step(n - done, want_macrodata=True)
            sc = scales(ref, g.inp.R, g.inp.gamma, 1.0 / g.inp.dx[0])
            worst, key = compare(g.fields(), ref, sc, n, tol_per_step=1.0)  # measure, do not assert
            rows[name][n] = {"worst": worst, "field": key, "per_step": worst / n}
        done = n
    omega = 1.0 / (gpus["two kernels"].inp.nu / (gpus["two kernels"].inp.R * float(np.abs(ref["temperature"]).max())) + 0.5)
    for name, r in rows.items():
        print(json.dumps({"variant": name, "deck": title, "omega_at_Tmax": omega, "errors_vs_oracle": r}))
    for g in gpus.values():
        g.close()
