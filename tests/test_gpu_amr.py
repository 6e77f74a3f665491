"""Multi-box / multi-level (AMR-exact) mode of the CUDA path through the C ABI (mbl_level_define_boxes, mbl_fillpatch
with coarse-fine interpolation, grown-box mbl_stream, mbl_average_down, mbl_collide) against golden vectors written by
the unmodified reference on 2- and 3-level decks (BASELINE configs 4-5 at reduced size) and against the multi-level
oracle on seeded random states.  Sub-cycling order driven from Python (marbles_b200/amr.py) as LBM::time_step does.
Tolerance: 1e-12 of the field scale per level step (tests/parity.py)."""
import os

import numpy as np
import pytest

from conftest import AMR_GOLDEN_CASES, amr_regrid_actions, load_amr_golden, load_golden
from parity import compare, scales

pytestmark = pytest.mark.gpu


def nan0(d):
    return {k: np.where(np.isnan(v), 0.0, v) for k, v in d.items()}


def golden_level(z, step, lev):
    pre = f"s{step}_l{lev}_"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


def compare_levels(amr, ref_of_level, nsteps_coarse, inp, macro=True, keys=None):
    worst_all = 0.0
    for lev in range(amr.finest + 1):
        ref = ref_of_level(lev)
        got = amr.fields(lev, macro=macro)
        cover = ~np.isnan(ref["f_00"])
        assert (np.isnan(got["f_00"]) == ~cover).all(), f"level {lev}: box coverage differs"
        r0 = nan0(ref)
        if "rho" not in r0:  # step 0 goldens hold f, g and macrodata; oracle dicts always do
            raise AssertionError("reference fields lack macrodata")
        sc = scales(r0, inp.R, inp.gamma, 2 ** lev / inp.dx[0])
        nsteps = nsteps_coarse * 2 ** lev
        worst, key = compare(nan0(got), r0, sc, nsteps, keys=keys)
        print(f"  level {lev}: worst {worst:.2e} ({key}) after {nsteps} level steps")
        worst_all = max(worst_all, worst)
    return worst_all


def new_amr(case):
    from marbles_b200.amr import AmrLBM
    from marbles_b200.inputs import parse_deck
    z, deck_text, steps, boxes, is_fluid = load_amr_golden(case)
    existing = [b for b in boxes if b]  # levels of plt00000 (a level may appear later)
    amr = AmrLBM(parse_deck(text=deck_text), existing, is_fluid[:len(existing)])
    amr.init_data()
    return amr, z, deck_text, steps, boxes, is_fluid


def full_golden_level(z, steps, lev):
    """a stored step that holds f, g and the macrodata of this level (the last one that does)"""
    for s in reversed(steps):
        g = golden_level(z, s, lev)
        if "f_00" in g and "rho" in g:
            return g
    raise AssertionError(f"no full step stored for level {lev}")


@pytest.mark.parametrize("case", AMR_GOLDEN_CASES)
def test_amr_cuda_vs_reference_golden(case):
    amr, z, deck_text, steps, boxes, is_fluid = new_amr(case)
    fg = [f"f_{q:02d}" for q in range(27)] + [f"g_{q:02d}" for q in range(27)]
    done = 0
    current = {lev: boxes[lev] for lev in range(1, len(boxes))}
    seen = set()
    for s in steps:
        if s == 0:
            compare_levels(amr, lambda lev: golden_level(z, 0, lev), 1, amr.inp, macro=False, keys=fg)
            continue
        while done < s:
            # AmrCore::regrid at the start of a coarse step: levels whose box list changed are re-made on the device,
            # new ones interpolated from below, vanished ones dropped
            for lev, what, nb in amr_regrid_actions(z, done + 1, current):
                if what == "make":
                    amr.make_level_from_coarse(lev, nb, is_fluid[lev])
                elif what == "remake":
                    amr.regrid_level(lev, nb, is_fluid[lev])
                else:
                    amr.clear_level(lev)
                seen.add(what)
            amr.step(1, want_macrodata=done + 1 == s)
            done += 1
        print(f"{case} step {s}, {amr.finest + 1} levels")
        assert not golden_level(z, s, amr.finest + 1), "the reference has one more level here"
        for lev in range(amr.finest + 1):
            ref = golden_level(z, s, lev)
            assert ref, f"the reference has no level {lev} at step {s}"
            if "f_00" in ref and "rho" in ref:
                continue  # compared below, with the box coverage
            # mid steps store the 19 macrodata fields only
            got = nan0(amr.fields(lev))
            sc = scales(nan0(full_golden_level(z, steps, lev)), amr.inp.R, amr.inp.gamma, 2 ** lev / amr.inp.dx[0])
            compare(got, nan0(ref), sc, s * 2 ** lev, keys=list(ref.keys()))
        if all("f_00" in golden_level(z, s, lev) for lev in range(amr.finest + 1)):
            compare_levels(amr, lambda lev: golden_level(z, s, lev), s, amr.inp)
    if case.endswith("_appear"):
        assert seen == {"make", "clear"}
    amr.close()


@pytest.mark.parametrize("case", ["amr2_tg", "amr3_chcyl"])
def test_amr_random_state_vs_oracle(oracle_mod, case):
    """seeded random perturbation of f, g on every level, two coarse steps against the multi-level oracle"""
    from oracle import amr_oracle as A
    O = oracle_mod
    amr, z, deck_text, steps, boxes, is_fluid = new_amr(case)
    o = A.AmrOracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines())), boxes, is_fluid)
    o.initialize()
    rng = np.random.default_rng(99)
    for lev, L in enumerate(o.levels):
        for ib, b in enumerate(L.boxes):
            for name, which in (("f", 0), ("g", 1)):
                a = getattr(b, name)
                a[:] = np.where(a > 0, a * (1.0 + 0.03 * rng.standard_normal(a.shape)), a)
        L.fill_boundary("f", 3)
        L.fill_boundary("g", 3)
        for ib, b in enumerate(L.boxes):
            amr.set_box(lev, ib, 0, b.f, ng=3)
            amr.set_box(lev, ib, 1, b.g, ng=3)
    nsteps = 2
    o.step(nsteps)
    amr.step(nsteps, want_macrodata=True)
    amr.compute_derived()  # post_time_step: vorticity of every level
    worst = compare_levels(amr, lambda lev: o.fields(lev), nsteps, amr.inp)
    print(f"{case}: worst {worst:.2e}")
    amr.close()


@pytest.mark.parametrize("case", ["amr2_chcyl", "amr3_chcyl", "amr2_sod_bc"])
def test_fused_advance_is_bit_identical_to_the_operator_sequence(case):
    """mbl_advance fuses the stream and the collide of the finest level into one pass (and drops the FillBoundary in
    between, whose ghost values nothing reads): every cell of every FAB, ghost cells included, and the macrodata must
    equal what mbl_stream; mbl_average_down; mbl_collide leave"""
    a, z, deck_text, steps, boxes, is_fluid = new_amr(case)
    b, *_ = new_amr(case)
    b.advance = b.advance_unfused
    l0 = (a.launches, b.launches)
    a.step(3, want_macrodata=True)
    b.step(3, want_macrodata=True)
    a.compute_derived(), b.compute_derived()
    assert a.launches - l0[0] < b.launches - l0[1]
    for lev in range(a.finest + 1):
        for ib in range(len(a.boxes[lev])):
            for which in (0, 1):
                assert np.array_equal(a.get_box(lev, ib, which, ng=3), b.get_box(lev, ib, which, ng=3)), (lev, ib, which)
        for which in ("macro", "derived"):
            assert np.array_equal(a.dense(lev, which), b.dense(lev, which), equal_nan=True), (lev, which)
    a.close()
    b.close()


def test_single_level_many_boxes_matches_single_box(oracle_mod):
    """a single level cut into many boxes (amr.max_grid_size < domain, as 10 of the 13 shipped decks do): the fused
    multi-box advance against the single-level oracle on the channel + cylinder deck"""
    from marbles_b200.amr import AmrLBM
    from marbles_b200.inputs import parse_deck
    from parity import compare, scales
    O = oracle_mod
    z, deck_text, _ = load_golden("chcyl")
    fl = z["is_fluid"].astype(np.int32)
    deck = parse_deck(text=deck_text)
    n = [int(v) for v in deck["amr.n_cell"]]
    cuts = [[(a, min(a + 7, n[d] - 1)) for a in range(0, n[d], 8)] for d in range(3)]
    boxes = [((x[0], y[0], zz[0]), (x[1], y[1], zz[1])) for zz in cuts[2] for y in cuts[1] for x in cuts[0]]
    amr = AmrLBM(deck, [boxes], [fl])
    amr.init_data()
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines())), fl)
    o.initialize()
    nsteps = 5
    o.step(nsteps)
    amr.step(nsteps, want_macrodata=True)
    got = amr.fields(0, macro=False)
    ref = {f"f_{q:02d}": o.f_valid[q] for q in range(27)} | {f"g_{q:02d}": o.g_valid[q] for q in range(27)}
    worst = 0.0
    for k, v in ref.items():
        worst = max(worst, float(np.abs(got[k] - v).max()) / max(float(np.abs(v).max()), 1e-300))
    print(f"{len(boxes)} boxes: worst {worst:.2e}")
    assert worst <= 1e-12 * nsteps
    amr.close()


def run_ranks(world, body):
    """`world` ranks as threads of this process on one device (marbles_b200.amr_comm.ThreadExchange)"""
    from concurrent.futures import ThreadPoolExecutor
    from marbles_b200.amr_comm import ThreadExchange
    ex = ThreadExchange(world, timeout=60.0)
    with ThreadPoolExecutor(world) as pool:
        futs = [pool.submit(body, r, ex) for r in range(world)]
        return [f.result(timeout=600) for f in futs]


@pytest.mark.parametrize("case,world", [("amr2_tg", 2), ("amr2_chcyl", 2), ("amr3_chcyl", 3), ("amr2_sod_regrid", 2),
                                        ("amr2_tg_appear", 3), ("amr2_sod_bc", 2)])
def test_distributed_levels_match_one_rank(case, world):
    """Boxes of every level spread over `world` ranks (AMReX's DistributionMapping; here round-robin, the worst case for
    the message count): FillBoundary, the coarse patches of the interpolation, both copies of the average-down and the
    old -> new copies of a regrid cross ranks as packed device messages.  Same kernels on the same cells: every FAB of
    every rank must equal the one-rank run bit for bit, through regrids and levels that appear and vanish."""
    from marbles_b200.amr import AmrLBM, merge_dense
    from marbles_b200.inputs import parse_deck
    z, deck_text, steps, boxes, is_fluid = load_amr_golden(case)
    existing = [b for b in boxes if b]
    deck = parse_deck(text=deck_text)
    nsteps = min(steps[-1], 8)
    current = {lev: boxes[lev] for lev in range(1, len(boxes))}
    actions = [amr_regrid_actions(z, done + 1, current) for done in range(nsteps)]  # (the npz reader is not thread-safe)

    def drive(amr):
        for done in range(nsteps):
            for lev, what, nb in actions[done]:
                if what == "make":
                    amr.make_level_from_coarse(lev, nb, is_fluid[lev])
                elif what == "remake":
                    amr.regrid_level(lev, nb, is_fluid[lev])
                else:
                    amr.clear_level(lev)
            amr.step(1, want_macrodata=done + 1 == nsteps)
        amr.compute_derived()
        amr.sync()
        return [{w: amr.dense(lev, w) for w in ("f", "g", "macro", "derived")} for lev in range(amr.finest + 1)]

    one = AmrLBM(deck, existing, is_fluid[:len(existing)])
    one.init_data()
    ref = drive(one)
    one.close()

    def body(rank, ex):
        amr = AmrLBM(deck, existing, is_fluid[:len(existing)], rank=rank, world=world, exchange=ex,
                     owners=lambda lev, bxs: [(i + lev) % world for i in range(len(bxs))])
        amr.init_data()
        out = drive(amr)
        nloc = sum(amr.is_local(l, i) for l in range(amr.finest + 1) for i in range(len(amr.boxes[l])))
        amr.close()
        return out, nloc

    res = run_ranks(world, body)
    assert all(r[1] > 0 for r in res), "every rank must hold boxes"
    for lev in range(len(ref)):
        for w in ("f", "g", "macro", "derived"):
            got = merge_dense([r[0][lev][w] for r in res])
            assert np.array_equal(got, ref[lev][w], equal_nan=True), (case, lev, w)


@pytest.mark.parametrize("case", ["amr2_chcyl", "amr3_chcyl"])
def test_amr_plotfile_from_device_state(oracle_mod, case, tmp_path):
    """LBM::write_plot_file for a hierarchy on the device (plotfile.write_amr_plotfile): read back with the oracle's
    plotfile reader, every component of every level against the reference's plotfile of the same step (the golden),
    is_fluid / eb_boundary and the box lists included"""
    from marbles_b200 import plotfile as P
    from parity import compare, scales
    O = oracle_mod
    amr, z, deck_text, steps, boxes, is_fluid = new_amr(case)
    s = steps[-1]
    amr.step(s, want_macrodata=True)
    amr.compute_derived()
    path = P.write_amr_plotfile(amr, str(tmp_path))
    assert os.path.basename(path) == f"plt{s:05d}"
    for lev in range(amr.finest + 1):
        assert [(list(a), list(b)) for a, b in O.read_plotfile_boxes(path, lev)] == [(list(a), list(b)) for a, b in boxes[lev]]
        pf = O.read_plotfile(path, lev)
        ref = golden_level(z, s, lev)
        sc = scales(nan0(ref), amr.inp.R, amr.inp.gamma, 2 ** lev / amr.inp.dx[0])
        got = {k: pf[k] for k in ref}
        worst, key = compare(nan0(got), nan0(ref), sc, s * 2 ** lev)
        fl = z[f"is_fluid_l{lev}"]
        m = ~np.isnan(pf["is_fluid"])
        assert np.array_equal(pf["is_fluid"][m], fl[m].astype(np.float64))
        print(f"{case} level {lev}: worst {worst:.2e} ({key}), {len(pf['__names__'])} components")
    amr.close()


def test_amr_eb_forces_match_reference_forces_file():
    """compute_eb_forces on a 3-level hierarchy on the device against the forces file the unmodified reference wrote
    (tests/golden/amr3_chcyl_forces.npz: uniform flow onto the cylinder, one line per coarse step)"""
    amr, z, deck_text, steps, boxes, is_fluid = new_amr("amr3_chcyl_forces")
    ref = z["forces"]
    scale = np.abs(ref[:, 1:]).max()
    worst = 0.0
    for s in range(len(ref)):
        got = amr.compute_eb_forces()
        worst = max(worst, float(np.abs(got - ref[s, 1:]).max()) / scale)
        if s + 1 < len(ref):
            amr.step(1)
    print(f"forces: worst {worst:.2e} of {scale:.3f} over {len(ref)} lines")
    assert worst <= 1e-12 * len(ref)
    amr.close()


def test_amr_checkpoint_restart_on_the_device(tmp_path):
    """LBM::write_checkpoint_file / read_checkpoint_file for a hierarchy: the state after 2 coarse steps written as a
    checkpoint, loaded into a new object, both run on -- bit-identical -- and the golden of step 4 is met"""
    from marbles_b200.amr import AmrLBM
    from marbles_b200.inputs import parse_deck
    a, z, deck_text, steps, boxes, is_fluid = new_amr("amr2_chcyl")
    a.step(2)
    path = a.write_checkpoint_file(str(tmp_path))
    assert os.path.basename(path) == "chk00002" and os.path.exists(os.path.join(path, "Level_1", "g_00_D_00000"))
    b = AmrLBM.from_checkpoint(parse_deck(text=deck_text), path, is_fluid)
    assert b.isteps == 2 and b.time == 2.0 and b.boxes == a.boxes
    a.step(2, want_macrodata=True)
    b.step(2, want_macrodata=True)
    a.compute_derived(), b.compute_derived()
    for lev in range(2):
        for w in ("f", "g", "macro"):
            assert np.array_equal(a.dense(lev, w), b.dense(lev, w), equal_nan=True), (lev, w)
    compare_levels(b, lambda lev: golden_level(z, 4, lev), 4, b.inp)
    a.close()
    b.close()


def test_level_bind_is_zero_copy_and_bit_identical():
    """mbl_level_bind: the fine level's boxes live in caller-owned device memory (torch tensors standing in for the
    FABs of an AMReX device-arena MultiFab, 27 comps x 3 ghost cells); every operator leaves its result there"""
    import ctypes as C
    import torch
    from marbles_b200._lib import check
    a, z, deck_text, steps, boxes, is_fluid = new_amr("amr2_chcyl")
    b, *_ = new_amr("amr2_chcyl")
    lev = 1
    fabs = []
    for ib in range(len(boxes[lev])):
        pair = []
        for which in (0, 1):
            host = b.get_box(lev, ib, which, ng=3)
            t = torch.from_numpy(host).cuda()
            check(b.lib.mbl_level_bind(b.ctx, lev, ib, which, C.c_void_p(t.data_ptr())))
            pair.append(t)
        fabs.append(pair)
    a.step(3)
    b.step(3)
    a.sync(), b.sync()
    for l in range(2):
        assert np.array_equal(a.dense(l, "f"), b.dense(l, "f"), equal_nan=True)
        assert np.array_equal(a.dense(l, "g"), b.dense(l, "g"), equal_nan=True)
    for ib, (tf, tg) in enumerate(fabs):  # the caller's memory holds the state, ghost cells included
        assert np.array_equal(tf.cpu().numpy(), a.get_box(lev, ib, 0, ng=3))
        assert np.array_equal(tg.cpu().numpy(), a.get_box(lev, ib, 1, ng=3))
    a.close()
    b.close()


def test_regrid_moves_the_fine_level(oracle_mod):
    """mbl_level_regrid (RemakeLevel) on the periodic Taylor-Green hierarchy: the fine region is shifted by two coarse
    cells and cut differently between two coarse steps, so the new boxes take old fine data, coarse-fine interpolated
    data (new valid cells) and periodic images; the run continues and matches the oracle driven through the same change"""
    from oracle import amr_oracle as A
    O = oracle_mod
    amr, z, deck_text, steps, boxes, is_fluid = new_amr("amr2_tg")
    setup = O.lbm_setup(O.parse_deck(None, deck_text.splitlines()))
    o = A.AmrOracle(setup, boxes, is_fluid)
    o.initialize()
    o.step(2)
    amr.step(2)
    lo = [min(b[0][d] for b in boxes[1]) for d in range(3)]
    hi = [max(b[1][d] for b in boxes[1]) for d in range(3)]
    lo2, hi2 = [lo[0] + 4, lo[1] - 4, lo[2]], [hi[0] + 4, hi[1] - 4, hi[2]]
    xm = (lo2[0] + hi2[0] + 1) // 2 // 2 * 2
    new = [(lo2, [xm - 1, hi2[1], hi2[2]]), ([xm, lo2[1], lo2[2]], hi2)]
    o.regrid_level(1, new, is_fluid[1])
    amr.regrid_level(1, new, is_fluid[1])
    o.step(2)
    amr.step(2, want_macrodata=True)
    amr.compute_derived()
    worst = compare_levels(amr, lambda lev: o.fields(lev), 4, amr.inp)
    print(f"regrid: worst {worst:.2e}")
    amr.close()


def test_fine_box_next_to_a_wall_is_refused():
    """the documented restriction fails loudly instead of interpolating wrongly"""
    from marbles_b200._lib import MarblesError
    from marbles_b200.amr import AmrLBM
    from marbles_b200.inputs import parse_deck
    z, deck_text, steps, boxes, is_fluid = load_amr_golden("amr2_chcyl")
    fine = [([16, 0, 0], [31, 15, 7])]  # touches the no-slip wall at y = 0
    amr = AmrLBM(parse_deck(text=deck_text), [boxes[0], fine], None)
    with pytest.raises(MarblesError, match="non-periodic domain face"):
        amr.init_data()  # the first inter-level operator (average_down after initialisation) builds the tag lists
        amr.step(1)
    amr.close()
