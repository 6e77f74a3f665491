"""Golden plotfile written by the UNMODIFIED reference (oracle/_ref/marbles3d.ex): the Taylor-Green deck of
tg12.npz at 8^3 cells with amr.max_grid_size=4 (8 boxes), step 1, all 82 components.  Stored in
plt_tg8.npz: the text of Header and Level_0/Cell_H, the SHA-256 of Level_0/Cell_D_00000 and the fields read
back with oracle.read_plotfile.  tests/test_plotfile.py checks that marbles_b200.plotfile reproduces the files
byte for byte from the fields.

    python tests/golden/make_plotfile_golden.py        (needs /root/reference built into oracle/_ref)
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

z = np.load(os.path.join(HERE, "tg12.npz"))
work = tempfile.mkdtemp(prefix="mbl_plt_")
with open(os.path.join(work, "tg.inp"), "w") as fh:
    fh.write(str(z["deck"]))
ov = ["max_step=1", "amr.plot_int=1", "lbm.save_streaming=1", "lbm.save_derived=1", "amr.max_grid_size=4",
      "amr.n_cell=8 8 8", "amr.blocking_factor=4", "amr.chk_int=-1"]
subprocess.run([O.REF_SERIAL, "tg.inp"] + ov, cwd=work, check=True, capture_output=True)
plt = os.path.join(work, "plt00001")
pf = O.read_plotfile(plt)
names = pf["__names__"]
np.savez_compressed(
    os.path.join(HERE, "plt_tg8.npz"),
    header=open(os.path.join(plt, "Header")).read(),
    cell_h=open(os.path.join(plt, "Level_0", "Cell_H")).read(),
    cell_d_sha256=hashlib.sha256(open(os.path.join(plt, "Level_0", "Cell_D_00000"), "rb").read()).hexdigest(),
    names=np.array(names), data=np.stack([pf[n] for n in names]), time=pf["__time__"], overrides=np.array(ov))
print("wrote plt_tg8.npz:", len(names), "components")

# ---- multi-level: a 3-level Taylor-Green hierarchy (8^3 cells per level, boxes of 4^3), step 2, the 28 components of
# lbm.save_streaming = 0.  Stored in plt_amr3.npz: Header, Level_k/Cell_H (text), SHA-256 of Level_k/Cell_D_00000, the
# box lists and the fields of every box -- the inputs marbles_b200.plotfile.write_plotfile_levels must turn back into
# the same bytes.
work = tempfile.mkdtemp(prefix="mbl_plt_amr_")
with open(os.path.join(work, "tg.inp"), "w") as fh:
    fh.write(str(z["deck"]))
ov = ["max_step=2", "amr.plot_int=2", "lbm.save_streaming=0", "lbm.save_derived=1", "amr.max_grid_size=4",
      "amr.n_cell=8 8 8", "amr.blocking_factor=4", "amr.chk_int=-1", "amr.max_level=2", "amr.n_error_buf=0",
      "amr.regrid_int=1000000", "tagging.refinement_indicators=a b", "tagging.a.in_box_lo=-0.5 -0.5 -0.5",
      "tagging.a.in_box_hi=0.5 0.5 0.5", "tagging.a.max_level=1", "tagging.b.in_box_lo=-0.25 -0.25 -0.25",
      "tagging.b.in_box_hi=0.25 0.25 0.25", "amrex.fpe_trap_invalid=0", "amrex.fpe_trap_zero=0", "amrex.fpe_trap_overflow=0"]
subprocess.run([O.REF_SERIAL, "tg.inp"] + ov, cwd=work, check=True, capture_output=True)
plt = os.path.join(work, "plt00002")
out = {"header": open(os.path.join(plt, "Header")).read(), "overrides": np.array(ov)}
nlev = 0
while os.path.isdir(os.path.join(plt, f"Level_{nlev}")):
    lev = nlev
    out[f"cell_h_{lev}"] = open(os.path.join(plt, f"Level_{lev}", "Cell_H")).read()
    out[f"cell_d_sha256_{lev}"] = hashlib.sha256(open(os.path.join(plt, f"Level_{lev}", "Cell_D_00000"), "rb").read()).hexdigest()
    boxes = O.read_plotfile_boxes(plt, lev)
    pf = O.read_plotfile(plt, lev)
    names = pf["__names__"]
    out[f"boxes_{lev}"] = np.array(boxes)
    dense = np.stack([pf[n] for n in names])
    for ib, (lo, hi) in enumerate(boxes):
        out[f"fab_{lev}_{ib}"] = dense[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]
    nlev += 1
out["names"] = np.array(names)
out["time"] = pf["__time__"]
out["nlev"] = nlev
np.savez_compressed(os.path.join(HERE, "plt_amr3.npz"), **out)
print("wrote plt_amr3.npz:", nlev, "levels,", [len(out[f"boxes_{l}"]) for l in range(nlev)], "boxes,", len(names), "components")
