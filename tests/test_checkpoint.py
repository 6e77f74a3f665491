"""Checkpoint compatibility with the UNMODIFIED reference (oracle/_ref/marbles3d.ex, run live): the VisMF files are
byte-identical, the reference restarts from a checkpoint written here and arrives at the same state as its own
uninterrupted run, and a checkpoint written by the reference loads into the CUDA path.  Skipped where the reference
executable has not been built."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

from marbles_b200 import plotfile as P

HERE = os.path.dirname(os.path.abspath(__file__))
OV = ["amr.n_cell=12 12 12", "amr.max_grid_size=8", "amr.plot_int=-1", "lbm.save_streaming=1", "lbm.save_derived=0"]


def _ref():
    from oracle import oracle as O
    if not os.path.exists(O.REF_SERIAL):
        pytest.skip("reference executable not built (python __graft_entry__.py)")
    return O


def _run(O, work, args):
    z = np.load(os.path.join(HERE, "golden", "tg12.npz"))
    os.makedirs(work, exist_ok=True)
    with open(os.path.join(work, "tg.inp"), "w") as fh:
        fh.write(str(z["deck"]))
    subprocess.run([O.REF_SERIAL, "tg.inp"] + OV + args, cwd=work, check=True, capture_output=True)


def test_vismf_files_byte_identical_to_reference(tmp_path):
    O = _ref()
    work = str(tmp_path / "ref")
    _run(O, work, ["max_step=4", "amr.chk_int=4"])
    ref = os.path.join(work, "chk00004")
    c = P.read_checkpoint(ref)
    assert (c["step"], c["dt"], c["time"]) == (4, 1.0, 4.0) and c["f"].shape == (27, 12, 12, 12)
    out = str(tmp_path / "mine" / "chk00004")
    P.write_checkpoint(out, c["f"], c["g"], step=4, dt=1.0, time=4.0, periodic=[1, 1, 1], max_grid_size=8)
    assert open(os.path.join(ref, "Header")).read() == open(os.path.join(out, "Header")).read()
    for name in ("f_00_H", "f_00_D_00000", "g_00_H", "g_00_D_00000"):
        assert filecmp.cmp(os.path.join(ref, "Level_0", name), os.path.join(out, "Level_0", name), shallow=False), name


def test_reference_restarts_from_our_checkpoint(tmp_path):
    """state after 4 steps (here from the oracle; the GPU test below takes it from the CUDA path) -> checkpoint ->
    the reference restarts, runs to step 10 and must match its own uninterrupted 10-step run"""
    O = _ref()
    z = np.load(os.path.join(HERE, "golden", "tg12.npz"))
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, str(z["deck"]).splitlines())))
    o.initialize()
    o.step(4)
    fl = o.fields()
    f = np.stack([fl[f"f_{q:02d}"] for q in range(27)])
    g = np.stack([fl[f"g_{q:02d}"] for q in range(27)])
    work = str(tmp_path / "restart")
    os.makedirs(work)
    P.write_checkpoint(os.path.join(work, "chk00004"), f, g, step=4, dt=1.0, time=4.0, periodic=[1, 1, 1], max_grid_size=8)
    _run(O, work, ["max_step=10", "amr.plot_int=10", "amr.chk_int=-1", "amr.restart=chk00004"])
    got = O.read_plotfile(os.path.join(work, "plt00010"))
    for q in range(27):
        for lat in ("f", "g"):
            k = f"{lat}_{q:02d}"
            assert np.array_equal(got[k], z[f"s10_{k}"]), k  # the oracle is bit-identical to the reference
    assert np.array_equal(got["rho"], z["s10_rho"])


@pytest.mark.gpu
def test_gpu_checkpoint_round_trip_with_reference(tmp_path):
    """CUDA path 4 steps -> checkpoint -> reference restarts to step 10 -> within tolerance of the golden step 10;
    and the reference's own step-4 checkpoint -> CUDA path to step 10"""
    from conftest import load_golden
    from parity import compare, scales
    from marbles_b200.inputs import parse_deck
    from marbles_b200.lbm import LBM
    O = _ref()
    z, deck_text, _ = load_golden("tg12")
    ref10 = {k[len("s10_"):]: z[k] for k in z.files if k.startswith("s10_")}
    lbm = LBM(parse_deck(text=deck_text, overrides=["amr.max_grid_size = 8"]))
    lbm.init_data()
    lbm.step(4)
    work = str(tmp_path / "restart")
    os.makedirs(work)
    path = lbm.write_checkpoint_file(work)
    assert os.path.basename(path) == "chk00004"
    _run(O, work, ["max_step=10", "amr.plot_int=10", "amr.chk_int=-1", "amr.restart=chk00004"])
    got = O.read_plotfile(os.path.join(work, "plt00010"))
    sc = scales(ref10, lbm.inp.R, lbm.inp.gamma, 1.0 / lbm.inp.dx[0])
    keys = [k for k in ref10 if k in got and not k.startswith("__")]
    worst, key = compare({k: got[k] for k in keys}, ref10, sc, 10)
    print(f"reference restarted from the CUDA checkpoint: worst {worst:.2e} ({key})")
    # the other direction
    work2 = str(tmp_path / "ref")
    _run(O, work2, ["max_step=4", "amr.chk_int=4"])
    lbm.read_checkpoint_file(os.path.join(work2, "chk00004"))
    assert lbm.isteps == 4 and lbm.time == 4.0
    lbm.step(6, want_macrodata=True)
    worst, key = compare(lbm.fields(), ref10, sc, 10)
    print(f"CUDA path continued from the reference checkpoint: worst {worst:.2e} ({key})")
    lbm.close()


def _run_amr(O, work, deck_text, args):
    os.makedirs(work, exist_ok=True)
    with open(os.path.join(work, "case.inp"), "w") as fh:
        fh.write(deck_text)
    subprocess.run([O.REF_SERIAL, "case.inp"] + args, cwd=work, check=True, capture_output=True)


def test_multilevel_checkpoint_round_trip_and_reference_restart(tmp_path):
    """A 2-level checkpoint of the reference (channel + cylinder, 3 + 2 boxes) read FAB by FAB and written again is the
    same bytes; the reference restarts from the rewritten one and arrives where its uninterrupted run does."""
    from conftest import load_amr_golden
    O = _ref()
    z, deck_text, steps, boxes, is_fluid = load_amr_golden("amr2_chcyl")
    ref = str(tmp_path / "ref")
    _run_amr(O, ref, deck_text, ["max_step=4", "amr.chk_int=2", "amr.plot_int=4"])
    c = P.read_checkpoint_levels(os.path.join(ref, "chk00002"))
    assert c["isteps"] == [2, 4] and c["dts"] == [1.0, 0.5] and c["ng"] == 3
    assert [[(list(a), list(b)) for a, b in lv[0]] for lv in c["levels"]] == [[(list(a), list(b)) for a, b in bx] for bx in boxes]
    mine = str(tmp_path / "mine")
    os.makedirs(mine)
    out = os.path.join(mine, "chk00002")
    P.write_checkpoint_levels(out, c["levels"], isteps=c["isteps"], dts=c["dts"], times=c["times"])
    assert open(os.path.join(ref, "chk00002", "Header")).read() == open(os.path.join(out, "Header")).read()
    for lev in range(2):
        for name in ("f_00_H", "f_00_D_00000", "g_00_H", "g_00_D_00000"):
            assert filecmp.cmp(os.path.join(ref, "chk00002", f"Level_{lev}", name), os.path.join(out, f"Level_{lev}", name),
                               shallow=False), (lev, name)
    _run_amr(O, mine, deck_text, ["max_step=4", "amr.chk_int=-1", "amr.plot_int=4", "amr.restart=chk00002"])
    for lev in range(2):
        a, b = O.read_plotfile(os.path.join(ref, "plt00004"), lev), O.read_plotfile(os.path.join(mine, "plt00004"), lev)
        for k in ("rho", "vel_x", "two_rho_e", "f_05", "g_11"):
            assert np.array_equal(a[k], b[k], equal_nan=True), (lev, k)


def test_multilevel_checkpoint_written_by_two_ranks_restarts_the_reference(tmp_path):
    """distributed hierarchies: every rank writes the FABs it holds (f_00_D_<rank>, g_00_D_<rank>), rank 0 the headers.
    Read back FAB by FAB it is the original data, and the unmodified reference restarts from it."""
    from conftest import load_amr_golden
    O = _ref()
    z, deck_text, steps, boxes, is_fluid = load_amr_golden("amr2_chcyl")
    ref = str(tmp_path / "ref")
    _run_amr(O, ref, deck_text, ["max_step=4", "amr.chk_int=2", "amr.plot_int=4"])
    c = P.read_checkpoint_levels(os.path.join(ref, "chk00002"))
    world = 2
    owners = [[(ib + lev) % world for ib in range(len(lv[0]))] for lev, lv in enumerate(c["levels"])]
    mine = str(tmp_path / "mine")
    os.makedirs(mine)
    out = os.path.join(mine, "chk00002")
    store, calls = {}, {}

    def write(rank):
        calls[rank] = 0

        def gather(obj):  # the n-th collective of every rank meets the n-th of the others
            store[(calls[rank], rank)] = obj
            parts = [store.get((calls[rank], r)) for r in range(world)]
            calls[rank] += 1
            return parts

        levels = [(bx, [f if owners[lev][ib] == rank else None for ib, f in enumerate(ff)],
                   [g if owners[lev][ib] == rank else None for ib, g in enumerate(gg)])
                  for lev, (bx, ff, gg) in enumerate(c["levels"])]
        P.write_checkpoint_levels(out, levels, isteps=c["isteps"], dts=c["dts"], times=c["times"], rank=rank, owners=owners,
                                  gather=gather)

    write(1)
    write(0)
    assert open(os.path.join(ref, "chk00002", "Header")).read() == open(os.path.join(out, "Header")).read()
    assert sorted(os.listdir(os.path.join(out, "Level_0"))) == ["f_00_D_00000", "f_00_D_00001", "f_00_H", "g_00_D_00000",
                                                                "g_00_D_00001", "g_00_H"]
    back = P.read_checkpoint_levels(out)
    for (bx, ff, gg), (bx2, ff2, gg2) in zip(c["levels"], back["levels"]):
        assert bx == bx2
        assert all(np.array_equal(a, b) for a, b in zip(ff, ff2)) and all(np.array_equal(a, b) for a, b in zip(gg, gg2))
    _run_amr(O, mine, deck_text, ["max_step=4", "amr.chk_int=-1", "amr.plot_int=4", "amr.restart=chk00002"])
    for lev in range(2):
        a, b = O.read_plotfile(os.path.join(ref, "plt00004"), lev), O.read_plotfile(os.path.join(mine, "plt00004"), lev)
        for k in ("rho", "two_rho_e", "f_05", "g_11"):
            assert np.array_equal(a[k], b[k], equal_nan=True), (lev, k)
