#!/usr/bin/env python
"""Drift of the CUDA path against the CPU restatement of the reference (oracle, bit-identical to the reference
on the golden decks) over many steps: periodic Taylor-Green 32^3, worst |difference| of every plotfile field
relative to its scale (tests/parity.py) after 1, 10, 30, 100 and 300 steps, for the default step (tile carry)
and the two-kernel step.  The stated tolerance is 1e-12 per step.  One JSON line per variant."""
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

z = np.load(os.path.join(ROOT, "tests", "golden", "tg12.npz"))
ov = ["amr.n_cell = 32 32 32"]
marks = [1, 10, 30, 100, 300]
o = O.Oracle(O.lbm_setup(O.parse_deck(None, str(z["deck"]).splitlines() + ov)))
o.initialize()
gpus = {name: LBM(parse_deck(text=str(z["deck"]), overrides=ov), variant=v) for name, v in (("tile carry (default)", 5), ("two kernels", 0))}
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
for name, r in rows.items():
    print(json.dumps({"variant": name, "deck": "Taylor-Green 32^3 periodic", "errors_vs_oracle": r}))
