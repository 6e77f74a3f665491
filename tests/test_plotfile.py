"""Plotfile emission: marbles_b200.plotfile against a plotfile written by the unmodified
reference (tests/golden/plt_tg8.npz, made by tests/golden/make_plotfile_golden.py) -- byte for byte -- and a
round trip through the oracle's plotfile reader."""
import hashlib
import os

import numpy as np
import pytest

from marbles_b200 import plotfile as P

GOLD = os.path.join(os.path.dirname(__file__), "golden", "plt_tg8.npz")


def test_var_names_match_reference_order():
    z = np.load(GOLD)
    assert P.plot_file_var_names(True, True) == [str(n) for n in z["names"]]
    assert len(P.plot_file_var_names(False, False)) == 19 + 2


def test_plotfile_bytes_match_reference(tmp_path):
    z = np.load(GOLD)
    names = [str(n) for n in z["names"]]
    out = str(tmp_path / "plt00001")
    P.write_plotfile(out, names, z["data"], time=float(z["time"]), step=1, prob_lo=[-1, -1, -1], prob_hi=[1, 1, 1],
                     max_grid_size=4)
    assert open(os.path.join(out, "Header")).read() == str(z["header"])
    assert open(os.path.join(out, "Level_0", "Cell_H")).read() == str(z["cell_h"])
    digest = hashlib.sha256(open(os.path.join(out, "Level_0", "Cell_D_00000"), "rb").read()).hexdigest()
    assert digest == str(z["cell_d_sha256"])


def test_multilevel_plotfile_bytes_match_reference(tmp_path):
    """amrex::WriteMultiLevelPlotfile of a 3-level hierarchy (tests/golden/plt_amr3.npz: written by the unmodified
    reference, 8 boxes per level): Header, every Level_k/Cell_H and every Level_k/Cell_D_00000 byte for byte"""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "plt_amr3.npz"))
    names = [str(n) for n in z["names"]]
    nlev = int(z["nlev"])
    levels = []
    for lev in range(nlev):
        boxes = [(tuple(int(v) for v in b[0]), tuple(int(v) for v in b[1])) for b in z[f"boxes_{lev}"]]
        levels.append((boxes, [z[f"fab_{lev}_{ib}"] for ib in range(len(boxes))]))
    out = str(tmp_path / "plt00002")
    P.write_plotfile_levels(out, names, levels, time=float(z["time"]), level_steps=[2 * 2 ** l for l in range(nlev)],
                            prob_lo=[-1, -1, -1], prob_hi=[1, 1, 1], n_cell=(8, 8, 8))
    assert open(os.path.join(out, "Header")).read() == str(z["header"])
    for lev in range(nlev):
        assert open(os.path.join(out, f"Level_{lev}", "Cell_H")).read() == str(z[f"cell_h_{lev}"]), lev
        digest = hashlib.sha256(open(os.path.join(out, f"Level_{lev}", "Cell_D_00000"), "rb").read()).hexdigest()
        assert digest == str(z[f"cell_d_sha256_{lev}"]), lev
    # and the oracle's reader gets the fields back, level by level
    from oracle import oracle as O
    for lev in range(nlev):
        pf = O.read_plotfile(out, lev)
        boxes, fabs = levels[lev]
        for (lo, hi), fab in zip(boxes, fabs):
            assert np.array_equal(pf["rho"][lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1], fab[0])


def test_multilevel_plotfile_written_by_several_ranks(tmp_path):
    """distributed levels: every rank writes the FABs it holds into Level_k/Cell_D_<rank>, rank 0 the headers (boxes in
    the level's order, FabOnDisk pointing into the owners' files).  Same Header as the reference's, and the oracle's
    reader gets every box back."""
    from oracle import oracle as O
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "plt_amr3.npz"))
    names = [str(n) for n in z["names"]]
    nlev, world = int(z["nlev"]), 3
    boxes = [[(tuple(int(v) for v in b[0]), tuple(int(v) for v in b[1])) for b in z[f"boxes_{lev}"]] for lev in range(nlev)]
    owners = [[(ib + lev) % world for ib in range(len(boxes[lev]))] for lev in range(nlev)]
    out = str(tmp_path / "plt00002")
    contributions = {}

    def write(rank):
        levels = [(boxes[lev], [z[f"fab_{lev}_{ib}"] if owners[lev][ib] == rank else None for ib in range(len(boxes[lev]))])
                  for lev in range(nlev)]

        def gather(obj):
            contributions[rank] = obj
            return [contributions.get(r) for r in range(world)]

        P.write_plotfile_levels(out, names, levels, time=float(z["time"]), level_steps=[2 * 2 ** l for l in range(nlev)],
                                prob_lo=[-1, -1, -1], prob_hi=[1, 1, 1], n_cell=(8, 8, 8), rank=rank, owners=owners,
                                gather=gather)

    for rank in (2, 1, 0):  # rank 0 last: it writes the headers from everybody's offsets and extrema
        write(rank)
    assert open(os.path.join(out, "Header")).read() == str(z["header"])
    assert sorted(os.listdir(os.path.join(out, "Level_1"))) == ["Cell_D_00000", "Cell_D_00001", "Cell_D_00002", "Cell_H"]
    for lev in range(nlev):
        pf = O.read_plotfile(out, lev)
        assert [(tuple(a), tuple(b)) for a, b in O.read_plotfile_boxes(out, lev)] == boxes[lev]
        for ib, (lo, hi) in enumerate(boxes[lev]):
            for c, nm in enumerate(names):
                assert np.array_equal(pf[nm][lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1], z[f"fab_{lev}_{ib}"][c])


@pytest.mark.parametrize("shape,mgs", [((5, 6, 7), 4), ((8, 8, 8), 8), ((3, 16, 4), 32)])
def test_plotfile_round_trip_through_oracle_reader(tmp_path, shape, mgs):
    from oracle import oracle as O
    rng = np.random.default_rng(7)
    names = ["a", "b", "c"]
    nz, ny, nx = shape
    data = rng.standard_normal((3, nz, ny, nx))
    out = str(tmp_path / "plt00042")
    P.write_plotfile(out, names, data, time=1.5, step=42, prob_lo=[0, 0, 0], prob_hi=[nx, ny, nz], max_grid_size=mgs)
    pf = O.read_plotfile(out)
    assert pf["__names__"] == names and pf["__time__"] == 1.5
    for c, n in enumerate(names):
        assert np.array_equal(pf[n], data[c])


def test_eb_boundary_flag():
    a = np.ones((5, 5, 5), dtype=np.int32)
    a[2, 2, 2] = 0
    a[2, 2, 3] = 0
    eb = P.eb_boundary(np.pad(a, 1, mode="edge"), 1)
    assert eb.sum() == 2 and eb[2, 2, 2] == 1 and eb[2, 2, 3] == 1


@pytest.mark.gpu
def test_lbm_write_plot_file_matches_golden(tmp_path):
    """the device state written as a plotfile, read back, against the reference's fields of the same step"""
    from conftest import load_golden
    from oracle import oracle as O
    from parity import compare, scales
    from marbles_b200.inputs import parse_deck
    from marbles_b200.lbm import LBM
    z, deck_text, steps = load_golden("chcyl")
    lbm = LBM(parse_deck(text=deck_text), is_fluid=z["is_fluid"].astype(np.int32))
    lbm.init_data()
    n = steps[-1]
    lbm.step(n, want_macrodata=True)
    path = P.write_lbm_plotfile(lbm, str(tmp_path))
    assert os.path.basename(path) == f"plt{n:05d}"
    pf = O.read_plotfile(path)
    ref = {k[len(f"s{n}_"):]: z[k] for k in z.files if k.startswith(f"s{n}_")}
    assert np.array_equal(pf["is_fluid"], ref["is_fluid"]) and np.array_equal(pf["eb_boundary"], ref["eb_boundary"])
    got = {k: pf[k] for k in ref if k in pf and k not in ("is_fluid", "eb_boundary")}
    sc = scales(ref, lbm.inp.R, lbm.inp.gamma, 1.0 / lbm.inp.dx[0])
    worst, key = compare(got, {k: ref[k] for k in got}, sc, n)
    print(f"plotfile fields: worst {worst:.2e} ({key}), {len(got)} components")
    lbm.close()


def test_box_lists_match_reference():
    """chop_boxes against the level-0 BoxArrays of the unmodified reference (tests/golden/boxes.json, made by
    tests/golden/make_boxes_golden.py): sizes that do and do not divide by max_grid_size, odd half-lengths"""
    import json
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "boxes.json")))
    assert len(cases) >= 6
    for c in cases:
        mine = [[list(lo), list(hi)] for lo, hi in P.chop_boxes(c["n_cell"], c["max_grid_size"])]
        assert mine == c["boxes"], (c["n_cell"], c["max_grid_size"])


def test_fcompare_and_product_reader(tmp_path, capsys):
    """marbles_b200.fcompare on two plotfiles (one written as a single file, one as two rank files)"""
    from marbles_b200 import fcompare as F
    rng = np.random.default_rng(11)
    names = ["rho", "vel_x"]
    data = rng.standard_normal((2, 12, 6, 10))
    a, b, c = str(tmp_path / "plt00001"), str(tmp_path / "two" / "plt00001"), str(tmp_path / "off" / "plt00001")
    P.write_plotfile(a, names, data, time=1.0, step=1, prob_lo=[0, 0, 0], prob_hi=[1, 1, 1], max_grid_size=4)
    # two-file version through the real collective path is covered by test_multirank_gloo; here: same data, other boxes
    P.write_plotfile(b, names, data, time=1.0, step=1, prob_lo=[0, 0, 0], prob_hi=[1, 1, 1], max_grid_size=8)
    pert = data.copy()
    pert[1, 3, 2, 1] += 1e-9
    P.write_plotfile(c, names, pert, time=1.0, step=1, prob_lo=[0, 0, 0], prob_hi=[1, 1, 1], max_grid_size=4)
    pf = P.read_plotfile(b)
    assert pf["__names__"] == names and pf["__step__"] == 1 and np.array_equal(pf["vel_x"], data[1])
    assert F.main([a, b]) == 0 and "PLOTFILES AGREE" in capsys.readouterr().out
    assert F.main([a, c]) == 1
    assert F.main([a, c, "--abs-tol", "1e-8"]) == 0
    rows, _ = F.compare(a, c)
    assert rows[0][1] == 0.0 and rows[1][1] == pytest.approx(1e-9, rel=1e-3)
