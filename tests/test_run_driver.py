"""`marbles_b200.run` (the reference's main/evolve loop around the GPU step) against the UNMODIFIED reference
executable run live on the same deck: same files written, fields within the stated tolerance."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("case,extra", [("tg12", []), ("chcyl", []), ("sod48", ["amr.max_grid_size=16"])])
def test_deck_runs_like_the_reference(tmp_path, case, extra):
    from oracle import oracle as O
    from parity import compare, scales
    from marbles_b200 import plotfile as P
    from marbles_b200.inputs import parse_deck
    from marbles_b200.lbm import LBM
    from marbles_b200.run import evolve
    if not os.path.exists(O.REF_SERIAL):
        pytest.skip("reference executable not built")
    z = np.load(os.path.join(HERE, "golden", case + ".npz"))
    ov = ["max_step=8", "amr.plot_int=4", "amr.chk_int=4", "lbm.save_streaming=1", "lbm.save_derived=1"] + extra
    ref_dir, my_dir = str(tmp_path / "ref"), str(tmp_path / "mine")
    os.makedirs(ref_dir), os.makedirs(my_dir)
    deck_path = os.path.join(ref_dir, "deck.inp")
    with open(deck_path, "w") as fh:
        fh.write(str(z["deck"]))
    subprocess.run([O.REF_SERIAL, "deck.inp"] + ov, cwd=ref_dir, check=True, capture_output=True)
    lbm = LBM(parse_deck(deck_path, overrides=ov))
    written = evolve(lbm, my_dir, log=lambda s: None)
    names = sorted(os.path.basename(p) for p in written)
    ref_names = sorted(d for d in os.listdir(ref_dir) if d.startswith(("plt", "chk")))
    assert names == ref_names == ["chk00000", "chk00004", "chk00008", "plt00000", "plt00004", "plt00008"]
    for step in (0, 4, 8):
        ref = O.read_plotfile(os.path.join(ref_dir, f"plt{step:05d}"))
        got = O.read_plotfile(os.path.join(my_dir, f"plt{step:05d}"))
        assert got["__names__"] == ref["__names__"] and got["__time__"] == ref["__time__"]
        assert np.array_equal(got["is_fluid"], ref["is_fluid"]) and np.array_equal(got["eb_boundary"], ref["eb_boundary"])
        keys = [k for k in ref["__names__"] if k not in ("is_fluid", "eb_boundary")]
        sc = scales(ref, lbm.inp.R, lbm.inp.gamma, 1.0 / lbm.inp.dx[0])
        worst, key = compare({k: got[k] for k in keys}, ref, sc, step)
        print(f"{case} plt{step:05d}: worst {worst:.2e} ({key}) over {len(keys)} fields")
        a, b = P.read_checkpoint(os.path.join(ref_dir, f"chk{step:05d}")), P.read_checkpoint(os.path.join(my_dir, f"chk{step:05d}"))
        assert (a["step"], a["time"], a["dt"]) == (b["step"], b["time"], b["dt"])
        for lat in ("f", "g"):
            scale = max(np.abs(a[lat]).max(), 1e-300)
            assert np.abs(a[lat] - b[lat]).max() <= 1e-12 * max(step, 1) * scale
    # the layout of the files themselves
    for name in ("Header", os.path.join("Level_0", "Cell_H")):
        ra = open(os.path.join(ref_dir, "plt00008", name)).read().split("\n")
        mb = open(os.path.join(my_dir, "plt00008", name)).read().split("\n")
        assert len(ra) == len(mb)
    lbm.close()


def test_forces_file_matches_reference(tmp_path):
    """lbm.compute_forces = 1: one line of EB forces per step, same layout as the reference's forces file"""
    from oracle import oracle as O
    from marbles_b200.inputs import parse_deck
    from marbles_b200.lbm import LBM
    from marbles_b200.run import evolve
    if not os.path.exists(O.REF_SERIAL):
        pytest.skip("reference executable not built")
    z = np.load(os.path.join(HERE, "golden", "chcyl.npz"))
    ov = ["max_step=6", "amr.plot_int=-1", "amr.chk_int=-1", "lbm.compute_forces=1", "lbm.forces_file=forces.txt"]
    ref_dir, my_dir = str(tmp_path / "ref"), str(tmp_path / "mine")
    os.makedirs(ref_dir), os.makedirs(my_dir)
    deck_path = os.path.join(ref_dir, "deck.inp")
    with open(deck_path, "w") as fh:
        fh.write(str(z["deck"]))
    subprocess.run([O.REF_SERIAL, "deck.inp"] + ov, cwd=ref_dir, check=True, capture_output=True)
    lbm = LBM(parse_deck(deck_path, overrides=ov))
    evolve(lbm, my_dir, log=lambda s: None)
    lbm.close()
    ref = open(os.path.join(ref_dir, "forces.txt")).read().split("\n")
    got = open(os.path.join(my_dir, "forces.txt")).read().split("\n")
    assert got[0] == ref[0] and len(got) == len(ref) == 9  # header, step 0 .. 6, trailing newline
    a = np.array([[float(v) for v in l.split()] for l in ref[1:-1]])
    b = np.array([[float(v) for v in l.split()] for l in got[1:-1]])
    assert all(len(l) == 96 for l in got[1:-1])
    assert np.array_equal(a[:, 0], b[:, 0])
    # the force is a sum of O(1e-2) momentum-exchange terms that cancel to O(1e-7): round-off is absolute
    assert np.abs(a[:, 1:] - b[:, 1:]).max() <= 1e-14
    print("forces: worst absolute difference", np.abs(a[:, 1:] - b[:, 1:]).max(), "of", np.abs(a[:, 1:]).max())
