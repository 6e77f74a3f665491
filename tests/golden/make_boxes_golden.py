"""Level-0 box lists of the UNMODIFIED reference (oracle/_ref/marbles3d.ex) for a few domain sizes and
amr.max_grid_size values, read from Level_0/Cell_H of its step-0 plotfile -> boxes.json.
tests/test_plotfile.py checks marbles_b200.plotfile.chop_boxes against them.

    python tests/golden/make_boxes_golden.py
"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

z = np.load(os.path.join(HERE, "tg12.npz"))
cases = [((12, 12, 12), 8), ((40, 24, 20), 16), ((48, 18, 6), 32), ((10, 14, 26), 8), ((64, 64, 64), 32), ((36, 20, 12), 16)]
out = []
for n, mgs in cases:
    work = tempfile.mkdtemp(prefix="mbl_box_")
    with open(os.path.join(work, "tg.inp"), "w") as fh:
        fh.write(str(z["deck"]))
    ov = ["max_step=0", "amr.plot_int=1", "lbm.save_streaming=0", "lbm.save_derived=0", f"amr.max_grid_size={mgs}",
          f"amr.n_cell={n[0]} {n[1]} {n[2]}", "amr.blocking_factor=2", "amr.chk_int=-1"]
    subprocess.run([O.REF_SERIAL, "tg.inp"] + ov, cwd=work, check=True, capture_output=True)
    ch = open(os.path.join(work, "plt00000", "Level_0", "Cell_H")).read().split("\n")
    i = next(k for k, l in enumerate(ch) if l.startswith("("))
    nbox = int(ch[i].strip("(").split()[0])
    boxes = []
    for b in range(nbox):
        m = [int(v) for v in re.findall(r"-?\d+", ch[i + 1 + b])]
        boxes.append([m[0:3], m[3:6]])
    out.append({"n_cell": list(n), "max_grid_size": mgs, "boxes": boxes})
    print(n, mgs, nbox, "boxes")
json.dump(out, open(os.path.join(HERE, "boxes.json"), "w"))
