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
