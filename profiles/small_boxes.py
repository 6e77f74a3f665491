#!/usr/bin/env python
"""Launch-bound boxes: steps per second of the golden decks (channel + cylinder with ~30 ghost-fill launches per
step, periodic TG 12^3 / 64^3) with mbl_step replaying CUDA graphs of step pairs (default) and without (MBL_GRAPH=0)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marbles_b200.inputs import parse_deck  # noqa: E402
from marbles_b200.lbm import LBM  # noqa: E402

for case, ov in (("chcyl", None), ("pressure", None), ("tg12", None), ("tg12", ["amr.n_cell = 64 64 64"])):
    z = np.load(os.path.join(ROOT, "tests", "golden", case + ".npz"))
    fl = z["is_fluid"].astype(np.int32) if ov is None else None
    row = {"deck": case + ("" if ov is None else " 64^3")}
    for graph in ("1", "0"):
        os.environ["MBL_GRAPH"] = graph
        lbm = LBM(parse_deck(text=str(z["deck"]), overrides=ov), is_fluid=fl)
        lbm.init_data()
        lbm.step(21)
        lbm.sync()
        t0 = time.perf_counter()
        lbm.step(2001)
        lbm.sync()
        dt = time.perf_counter() - t0
        row["cells"] = lbm.ncells
        row["us_per_step_graph" if graph == "1" else "us_per_step_eager"] = dt / 2001 * 1e6
        lbm.close()
    row["speedup"] = row["us_per_step_eager"] / row["us_per_step_graph"]
    print(json.dumps(row))
