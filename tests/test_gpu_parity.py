"""Parity of the CUDA path (through the C ABI) with the reference: golden vectors written by the
unmodified reference executable, and the oracle on seeded inputs.  fp64, tolerance 1e-12 of the
field scale per step (tests/parity.py)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, golden_is_fluid, load_golden
from parity import compare, scales

pytestmark = pytest.mark.gpu


def golden_fields(z, s):
    pre = f"s{s}_"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


# step variants (mbl_set_variant).  "carry*" = variant 4 with different marching / warp-overlap tunings so
# that small boxes still exercise several z-chunks, ragged chunks, several warps per row and both launch bounds
CARRY_ENV = {
    "carry": (4, {}),
    "carry-ky5-own28": (4, {"MBL_KY": "5", "MBL_OWN": "28", "MBL_MINB": "3"}),
    "carry-ky1": (4, {"MBL_KY": "1"}),
    "tile": (5, {}),
    "tile-6rows-own28": (5, {"MBL_ROWS": "6", "MBL_OWN": "28"}),
    "tile-12rows": (5, {"MBL_ROWS": "12"}),
    "tile-4rows": (5, {"MBL_ROWS": "4"}),
    "lean": (6, {"MBL_MINB": "4"}),
    "pair": (7, {}),
    # variant 9: tile carry step marching through z-chunks (z sums completed on chip, QCorr finished by the collide kernel)
    "zmarch": (9, {}),
    "zmarch-zm3": (9, {"MBL_ZMARCH": "3"}),
    "zmarch-zm2": (9, {"MBL_ZMARCH": "2"}),
    # variant 10 (MBL_EXPERIMENTS): the same with plane k+1 pulled into shared memory while plane k is collided
    "zpipe": (10, {}),
    "zpipe-4rows-zm3": (10, {"MBL_ROWS": "4", "MBL_ZMARCH": "3"}),
    "zpipe-zm2": (10, {"MBL_ZMARCH": "2"}),
    # variant 8: the one-kernel march step (march.cu): default tuning, small CTAs with short ragged marches, and
    # without the early pull of the next plane
    "march": (8, {}),
    "march-4rows-zm3": (8, {"MBL_MROWS": "4", "MBL_ZM": "3"}),
    "march-nopipe-zm5": (8, {"MBL_PIPE": "0", "MBL_ZM": "5"}),
}
TUNING_VARS = ("MBL_KY", "MBL_OWN", "MBL_MINB", "MBL_ROWS", "MBL_MROWS", "MBL_ZM", "MBL_PIPE", "MBL_ZMARCH")
# variants 1-4 are round-1 experiments, compiled only with MBL_EXPERIMENTS=1
import os as _os
EXPERIMENTS = _os.environ.get("MBL_EXPERIMENTS") == "1"


def with_experiments(default, extra):
    return list(default) + (list(extra) if EXPERIMENTS else [])


def variant_of(v):
    """-> (variant number, environment overrides)"""
    if isinstance(v, str):
        return CARRY_ENV[v]
    return v, {}


def new_lbm(deck_text, is_fluid=None, overrides=None, variant=None):
    import os
    from marbles_b200.inputs import parse_deck
    from marbles_b200.lbm import LBM
    deck = parse_deck(text=deck_text, overrides=overrides)
    variant, env = variant_of(variant)
    for k in TUNING_VARS:
        os.environ.pop(k, None)
    os.environ.update(env)  # read by mbl_create
    lbm = LBM(deck, is_fluid=is_fluid, variant=variant)
    lbm.init_data()
    return lbm


# fused: mbl_step with the persistent TMA kernel (variant 1, the default), its two job types as two
# launches (2), or the two plain kernels (0); unfused: the reference-granular operator sequence
@pytest.mark.parametrize("fused", with_experiments([0, "tile", "tile-6rows-own28", "lean", "pair", "zmarch", "zmarch-zm3", None],
                                                   [1, 2, 3, "carry", "carry-ky5-own28", "march", "march-4rows-zm3", "march-nopipe-zm5", "zpipe",
                                                    "zpipe-4rows-zm3"]),
                         ids=lambda v: {0: "twopass-plain", 1: "fused-tma", 2: "twopass-tma", 3: "fused-plain",
                                        None: "unfused"}.get(v, str(v)))
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_cuda_vs_reference_golden(case, fused):
    z, deck_text, steps = load_golden(case)
    lbm = new_lbm(deck_text, golden_is_fluid(case), variant=fused)
    inp = lbm.inp
    done = 0
    for s in steps:
        if s == 0:
            ref = golden_fields(z, 0)
            sc = scales(ref, inp.R, inp.gamma, 1.0 / inp.dx[0])
            mine = {"f": lbm.get_f(), "g": lbm.get_g()}
            got = {f"f_{q:02d}": mine["f"][q] for q in range(27)}
            got.update({f"g_{q:02d}": mine["g"][q] for q in range(27)})
            compare(got, ref, sc, 1)
            continue
        if fused is not None:
            lbm.step(s - done, want_macrodata=True)
        else:
            lbm.evolve(s - done, fused=False)
        done = s
        ref = golden_fields(z, s)
        sc = scales(ref, inp.R, inp.gamma, 1.0 / inp.dx[0])
        got = lbm.fields()
        d = lbm.get_derived()
        got.update({"dQCorrX": d[4], "dQCorrY": d[5], "dQCorrZ": d[6]})
        worst, key = compare(got, ref, sc, s)
        print(f"{case} step {s}: worst {worst:.2e} ({key})")
    lbm.close()


def test_solid_sentinel_and_macrodata_zero():
    z, deck_text, _ = load_golden("chcyl")
    fl = z["is_fluid"].astype(np.int32)
    lbm = new_lbm(deck_text, fl)
    assert (lbm.get_f()[:, fl == 0] == 0.0).all()
    lbm.step(1, want_macrodata=True)
    assert (lbm.get_f()[:, fl == 0] == -1.0).all() and (lbm.get_g()[:, fl == 0] == -1.0).all()
    assert (lbm.get_macrodata()[:, fl == 0] == 0.0).all()
    lbm.close()


def test_geometry_matches_reference_is_fluid():
    """host-mirror EB flags (corner rule) against the reference's is_fluid for the shipped body types"""
    from marbles_b200.geometry import is_fluid_from_deck
    from marbles_b200.inputs import lbm_inputs, parse_deck
    for case in ("chcyl", "pressure", "slip"):
        z, deck_text, _ = load_golden(case)
        deck = parse_deck(text=deck_text)
        inp = lbm_inputs(deck)
        a = is_fluid_from_deck(deck, inp.n_cell, inp.prob_lo, inp.dx, ng=3)[3:-3, 3:-3, 3:-3]
        assert np.array_equal(a, z["is_fluid"].astype(np.int32)), case


@pytest.mark.parametrize("variant", with_experiments([0, "tile", "tile-6rows-own28", "tile-12rows", "tile-4rows", "pair", "zmarch", "zmarch-zm3", "zmarch-zm2"],
                                                     [1, 3, "carry", "carry-ky5-own28", "carry-ky1", "march", "march-4rows-zm3",
                                                      "march-nopipe-zm5", "zpipe", "zpipe-4rows-zm3", "zpipe-zm2"]),
                         ids=lambda v: {0: "twopass-plain", 1: "fused-tma", 3: "fused-plain"}.get(v, str(v)))
@pytest.mark.parametrize("case", ["chcyl", "pressure", "slip", "tg12"])
def test_random_state_vs_oracle(oracle_mod, case, variant):
    """seeded random perturbation of f, g and a random solid mask, 3 steps, all boundary types"""
    O = oracle_mod
    z, deck_text, _ = load_golden(case)
    rng = np.random.default_rng(1234)
    fl = z["is_fluid"].astype(np.int32).copy()
    nzv, nyv, nxv = fl.shape
    # sprinkle solid cells in the interior (away from the non-periodic faces)
    mask = rng.random(fl.shape) < 0.03
    mask[:2], mask[-2:], mask[:, :2], mask[:, -2:], mask[:, :, :2], mask[:, :, -2:] = (False,) * 6
    fl[mask] = 0
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines())), is_fluid=fl)
    o.initialize()
    noise = lambda a: a * (1.0 + 0.05 * rng.standard_normal(a.shape))
    o.f[:] = np.where(o.f > 0, noise(o.f), o.f)
    o.g[:] = np.where(o.g > 0, noise(o.g), o.g)
    lbm = new_lbm(deck_text, fl, variant=variant)
    lbm.set_state(o.f, o.g, ng=3)
    nsteps = 4
    o.step(nsteps)
    lbm.step(nsteps, want_macrodata=True)
    ref = o.fields()
    sc = scales(ref, lbm.inp.R, lbm.inp.gamma, 1.0 / lbm.inp.dx[0])
    got = lbm.fields()
    worst, key = compare(got, ref, sc, nsteps)
    print(f"{case}: worst {worst:.2e} ({key})")
    lbm.close()


def test_eb_forces_and_vorticity_vs_oracle(oracle_mod):
    O = oracle_mod
    z, deck_text, _ = load_golden("chcyl")
    fl = z["is_fluid"].astype(np.int32)
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines())), is_fluid=fl)
    o.initialize()
    lbm = new_lbm(deck_text, fl)
    o.step(4)
    lbm.step(4, want_macrodata=True)
    lbm.compute_derived()
    fo, fm = o.eb_forces(), lbm.compute_eb_forces()
    assert np.abs(fo - fm).max() <= 1e-11 * max(np.abs(fo).max(), 1.0), (fo, fm)
    d = lbm.get_derived()
    cs = np.sqrt(lbm.inp.gamma * lbm.inp.R * 0.03)
    for n in range(4):
        assert np.abs(d[n] - o.derived[n]).max() <= 4e-12 * cs
    lbm.close()


@pytest.mark.parametrize("variant", with_experiments([0, "tile", "pair", "zmarch"], ["carry", "march", "zpipe"]),
                         ids=lambda v: "twopass-plain" if v == 0 else str(v))
def test_tg64_vs_oracle_and_conservation(oracle_mod, variant):
    """BASELINE config 1 (TG 64^3): 3 steps against the oracle, then size-independent properties"""
    O = oracle_mod
    z, deck_text, _ = load_golden("tg12")
    ov = ["amr.n_cell = 64 64 64"]
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines() + ov)))
    o.initialize()
    lbm = new_lbm(deck_text, overrides=ov, variant=variant)
    m0, e0 = lbm.get_f().sum(), lbm.get_g().sum()
    o.step(3)
    lbm.step(3, want_macrodata=True)
    ref = o.fields()
    sc = scales(ref, 1.0, 5.0 / 3.0, 32.0)
    worst, key = compare(lbm.fields(), ref, sc, 3)
    print(f"tg64: worst {worst:.2e} ({key})")
    lbm.step(20)
    f, g = lbm.get_f(), lbm.get_g()
    assert abs(f.sum() - m0) <= 1e-11 * m0 and abs(g.sum() - e0) <= 1e-11 * e0
    # TG symmetry: w == 0 initially and the flow is mirror-symmetric in z -> sum of e_z f vanishes
    from oracle.oracle import stencil
    ev = stencil()[0]
    jz = np.tensordot(ev[:, 2].astype(float), f, axes=1)
    assert abs(jz.sum()) <= 1e-10 * m0
    lbm.close()


@pytest.mark.parametrize("variant", with_experiments([0, "tile", "zmarch"], ["carry", "march", "zpipe"]),
                         ids=lambda v: "twopass-plain" if v == 0 else str(v))
def test_full_size_conservation_256(variant):
    """periodic 256^3 (largest size the test box does in seconds): mass/energy conservation of
    stream+collide and agreement of the fused and un-fused operator sequences"""
    _, deck_text, _ = load_golden("tg12")
    ov = ["amr.n_cell = 256 256 256"]
    a = new_lbm(deck_text, overrides=ov, variant=variant)
    s0 = a.get_f(0).sum(dtype=np.float64), a.get_g(0).sum(dtype=np.float64)
    a.step(10)
    f = a.get_f(0)
    assert abs(f.sum(dtype=np.float64) - s0[0]) <= 1e-11 * s0[0]
    assert np.isfinite(f).all() and f.min() > 0
    g = a.get_g(0)
    assert abs(g.sum(dtype=np.float64) - s0[1]) <= 1e-11 * s0[1]
    a.close()


@pytest.mark.parametrize("case,nz,world", [("tg12", 12, 2), ("tg12", 13, 3), ("sod48", 8, 2), ("chcyl", None, 2),
                                           ("pressure", None, 2)])
@pytest.mark.parametrize("variant", with_experiments([0, "tile-6rows-own28", "zmarch-zm3"], ["carry-ky5-own28", "march-4rows-zm3", "zpipe-4rows-zm3"]),
                         ids=lambda v: "twopass-plain" if v == 0 else str(v))
def test_two_slabs_match_single_box(case, nz, world, variant):
    """the multi-rank scheme (z-slabs, ONE exchange of two ghost planes per step, q-correction of the first
    ghost plane recomputed locally, BC ghosts of neighbour-owned planes) on one device: the assembled slabs
    must reproduce the single-box run to round-off (different kernels launches, same arithmetic per cell:
    the two are expected to be bit-identical)"""
    import os
    import torch
    from marbles_b200.inputs import lbm_inputs, parse_deck
    from marbles_b200.lbm import LBM, slab_bounds
    from marbles_b200.parallel import LocalSlabs
    variant, env = variant_of(variant)
    for k in TUNING_VARS:
        os.environ.pop(k, None)
    os.environ.update(env)
    z, deck_text, _ = load_golden(case)
    fl = z["is_fluid"].astype(np.int32)
    ov = None
    if nz is not None:
        n = lbm_inputs(parse_deck(text=deck_text)).n_cell
        ov = [f"amr.n_cell = {n[0]} {n[1]} {nz}"]
        fl = None if fl.min() == 1 else fl
        assert fl is None
    deck = parse_deck(text=deck_text, overrides=ov)
    single = LBM(deck, is_fluid=fl, variant=0)
    single.init_data()
    nzt = single.n_local[2]

    def make(rank, w):
        lo, hi = slab_bounds(nzt, rank, w)
        if fl is None:
            sub = None
        else:
            # is_fluid of the slab grown by 3: interior neighbours from the full array, domain ends = fluid
            ng = 3
            full = np.ones((nzt + 2 * ng,) + tuple(d + 2 * ng for d in fl.shape[1:]), dtype=np.int32)
            full[ng:-ng, ng:-ng, ng:-ng] = fl
            full = single._wrap_periodic(full, ng, z_local=True)
            sub = np.ascontiguousarray(full[lo:hi + 1 + 2 * ng])
        s = LBM(deck, rank=rank, world=w, comm=None, is_fluid=sub, variant=variant)
        s.init_data()
        return s

    slabs = LocalSlabs(make, world, bool(single.inp.periodic[2]), torch.device("cuda", 0))
    nsteps = 4
    single.step(nsteps, want_macrodata=True)
    slabs.step(nsteps, want_macrodata=True)
    for name, get in (("f", lambda s: s.get_f()), ("g", lambda s: s.get_g()), ("macro", lambda s: s.get_macrodata())):
        a, b = get(single), slabs.gather(get)
        assert a.shape == b.shape
        err = float(np.abs(a - b).max())
        # carry: the moments behind the q-corrections are summed in another order -> round-off, not bit-identical
        assert err <= (1e-13 if variant == 0 else 1e-12 * nsteps) * max(float(np.abs(a).max()), 1.0), (case, name, err)
    slabs.close()
    single.close()


@pytest.mark.parametrize("case,ov,chunk,ng", [
    ("tg12", ["amr.n_cell = 16 12 23"], "3", 0),     # pipelined: many ragged z-chunks, tight host FABs
    ("tg12", ["amr.n_cell = 16 12 23"], "16", 3),    # pipelined: host FABs with the reference's 3 ghost cells
    ("tg12", ["amr.n_cell = 16 12 9"], "1", 0),      # pipelined: one plane per chunk
    ("tg12", ["amr.n_cell = 16 12 23"], "-1", 0),    # pipelining off
    ("chcyl", None, "3", 3),                          # not all-periodic: the general upload / step / download path
])
def test_step_host_matches_device_step(case, ov, chunk, ng):
    """mbl_step_host (host FAB buffers in, host FAB buffers out; z-chunked copies overlapped with the kernels on
    all-periodic boxes) against the device-resident step: same kernels, same planes -> bit-identical"""
    import os
    z, deck_text, _ = load_golden(case)
    fl = z["is_fluid"].astype(np.int32) if ov is None else None
    os.environ["MBL_HOST_CHUNK"] = chunk
    try:
        a = new_lbm(deck_text, fl, overrides=ov, variant=0)
        b = new_lbm(deck_text, fl, overrides=ov, variant=0)
    finally:
        os.environ.pop("MBL_HOST_CHUNK", None)
    a.step(2)
    f, g = a.get_f(ng), a.get_g(ng)
    b.step(2)
    for _ in range(3):
        a.step(1)
        b.step_host(f, g, 1, ng=ng)
    s = (slice(None),) + ((slice(ng, -ng),) * 3 if ng else (slice(None),) * 3)
    assert np.array_equal(f[s], a.get_f()) and np.array_equal(g[s], a.get_g())
    assert np.array_equal(b.get_f(), a.get_f())
    a.close()
    b.close()


@pytest.mark.parametrize("variant", with_experiments([0, 5, None, "zmarch-zm3", "zmarch-zm2"], [8]),
                         ids=lambda v: {0: "twopass", 5: "tile", 8: "march", None: "default"}.get(v, str(v)))
@pytest.mark.parametrize("case,nxy,nz,world", [("tg12", "12 12", 16, 2), ("tg12", "12 12", 27, 3), ("chcyl", "32 12", 16, 2),
                                               ("pressure", "10 10", 17, 2), ("tg12", "12 12", 40, 2)])
def test_overlapped_slab_step_matches_single_box(case, nxy, nz, world, variant):
    """mbl_step_split (ghost fill where the level has walls, boundary planes first, exchange of the written buffers'
    boundary planes, interior planes) in the order LBM._step_overlapped issues it, on one device: bit-identical to the
    single box.  tg12: all periodic (lean halo); chcyl: inlet / outflow / no-slip / periodic z with a cylinder
    through every slab cut; pressure: non-periodic z as well (slabs end at an inlet and a pressure outlet)"""
    import torch
    from marbles_b200.inputs import parse_deck
    from marbles_b200.lbm import LBM, slab_bounds
    from marbles_b200.parallel import LocalSlabs
    z, deck_text, _ = load_golden(case)
    ov = [f"amr.n_cell = {nxy} {nz}"]
    if case == "pressure":
        ov.append(f"geometry.prob_hi = 10.0 10.0 {nz}.0")
    elif case == "chcyl":
        ov.append(f"geometry.prob_hi = 32.0 12.0 {nz}.0")
    deck = parse_deck(text=deck_text, overrides=ov)
    single = LBM(deck, variant=0)
    single.init_data()

    import os
    vnum, env = variant_of(variant)
    for k in TUNING_VARS:
        os.environ.pop(k, None)
    os.environ.update(env)  # read by mbl_create

    def make(rank, w):
        s = LBM(deck, rank=rank, world=w, comm=None, variant=vnum)
        s.init_data()
        return s

    slabs = LocalSlabs(make, world, bool(single.inp.periodic[2]), torch.device("cuda", 0))
    for k in env:
        os.environ.pop(k, None)
    single.step(7)
    slabs.step_overlapped(3)
    slabs.step(2)  # and back to the plain slab step (the carried sums change their chunk layout)
    slabs.step_overlapped(2)
    for get in (lambda s: s.get_f(), lambda s: s.get_g()):
        a, b = get(single), slabs.gather(get)
        if variant == 0:
            assert np.array_equal(a, b)
        else:  # carried moments are summed in another order
            assert np.abs(a - b).max() <= 7e-12 * np.abs(a).max()
    slabs.close()
    single.close()


@pytest.mark.parametrize("case", ["chcyl", "tg12", "pressure"])
def test_lean_collide_is_bit_identical_to_collide(case):
    """k_collide_lean (g through shared memory, 32-bit offsets) does the arithmetic of k_collide in the same order"""
    import os
    z, deck_text, _ = load_golden(case)
    fl = z["is_fluid"].astype(np.int32)
    a = new_lbm(deck_text, fl, variant=0)
    b = new_lbm(deck_text, fl, variant="lean")
    os.environ["MBL_PLAIN_COLLIDE"] = "1"  # variant 0 with the original k_collide
    try:
        a.step(6)
    finally:
        os.environ.pop("MBL_PLAIN_COLLIDE", None)
    b.step(6)
    assert np.array_equal(a.get_f(), b.get_f()) and np.array_equal(a.get_g(), b.get_g())
    a.close()
    b.close()


@pytest.mark.parametrize("nz,world,chunk,ng", [(24, 2, "3", 0), (27, 3, "16", 3)])
def test_step_host_on_slabs_matches_device_slab_step(nz, world, chunk, ng):
    """mbl_step_host_begin / exchange / mbl_step_host_finish (host FABs in and out, pipelined behind the upload
    frontier) against the device-resident slab step: same kernels on the same planes -> bit-identical"""
    import os
    import torch
    from marbles_b200.inputs import parse_deck
    from marbles_b200.lbm import LBM
    from marbles_b200.parallel import LocalSlabs
    z, deck_text, _ = load_golden("tg12")
    deck = parse_deck(text=deck_text, overrides=[f"amr.n_cell = 16 12 {nz}"])

    def make(rank, w):
        s = LBM(deck, rank=rank, world=w, comm=None, variant=0)
        s.init_data()
        return s

    os.environ["MBL_HOST_CHUNK"] = chunk
    try:
        a = LocalSlabs(make, world, True, torch.device("cuda", 0))
        b = LocalSlabs(make, world, True, torch.device("cuda", 0))
    finally:
        os.environ.pop("MBL_HOST_CHUNK", None)
    a.step(2)
    b.step(2)
    fabs = [(s.get_f(ng), s.get_g(ng)) for s in a.slabs]
    for _ in range(3):
        a.step(1)
        b.step_host(fabs, ng)
    inner = (slice(None),) + ((slice(ng, -ng),) * 3 if ng else (slice(None),) * 3)
    for s, (f, g) in zip(a.slabs, fabs):
        assert np.array_equal(f[inner], s.get_f()) and np.array_equal(g[inner], s.get_g())
    assert np.array_equal(a.gather(lambda s: s.get_f()), b.gather(lambda s: s.get_f()))
    a.close()
    b.close()


def test_full_size_conservation_512():
    """BASELINE config 3 at its full size (periodic 512^3, the default step): mass and energy are conserved by
    stream + collide, every population stays finite and f positive"""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 150e9:
        pytest.skip("needs the 180 GB of a B200")
    _, deck_text, _ = load_golden("tg12")
    a = new_lbm(deck_text, overrides=["amr.n_cell = 512 512 512"])
    sums = []
    for it in range(2):
        if it == 1:
            a.step(4)
        tot = []
        for name, get in (("f", a.get_f), ("g", a.get_g)):
            x = get(0)
            if it == 1:  # the energy lattice g may be negative (its rest population is close to zero)
                assert np.isfinite(x).all() and (name == "g" or x.min() > 0)
            tot.append(x.sum(dtype=np.float64))
            del x
        sums.append(tot)
    assert abs(sums[1][0] - sums[0][0]) <= 1e-11 * sums[0][0]
    assert abs(sums[1][1] - sums[0][1]) <= 1e-11 * sums[0][1]
    a.close()


@pytest.mark.parametrize("case", ["chcyl", "tg12", "sod48"])
@pytest.mark.parametrize("variant", [None, 0, "tile"], ids=["default", "twopass", "tile"])
def test_graph_replay_is_bit_identical(case, variant):
    """mbl_step replays pairs of steps as one CUDA graph on small boxes: same kernels, same order.  A step with
    macrodata between replays moves the buffer parity without swapping the two QCorr arrays of the default
    (z-march) step, so the graph must be re-captured for the state it is replayed on."""
    import os
    z, deck_text, _ = load_golden(case)
    fl = z["is_fluid"].astype(np.int32)
    os.environ["MBL_GRAPH"] = "0"
    try:
        a = new_lbm(deck_text, fl, variant=variant)
    finally:
        os.environ.pop("MBL_GRAPH", None)
    b = new_lbm(deck_text, fl, variant=variant)
    for n in (8, 5, 1, 7):
        a.step(n)
        b.step(n)
    a.step(6, want_macrodata=True)
    b.step(6, want_macrodata=True)
    assert np.isfinite(a.get_f()).all()
    assert np.array_equal(a.get_f(), b.get_f()) and np.array_equal(a.get_g(), b.get_g())
    assert np.array_equal(a.get_macrodata(), b.get_macrodata())
    assert a.launches == b.launches
    for n in (7, 1, 6, 9):
        a.step(n)
        b.step(n)
        a.step(1, want_macrodata=True)
        b.step(1, want_macrodata=True)
    assert np.array_equal(a.get_f(), b.get_f()) and np.array_equal(a.get_g(), b.get_g())
    assert np.array_equal(a.get_macrodata(), b.get_macrodata())
    assert a.launches == b.launches
    a.close()
    b.close()


@pytest.mark.parametrize("case,n_cell", [("tg12", "67 45 13"), ("tg12", "31 7 9"), ("sod48", "75 3 5"), ("sod48", "130 2 2")])
@pytest.mark.parametrize("variant", with_experiments([None, 0, "tile", "pair", "zmarch-zm3"], ["carry", "march", "march-4rows-zm3", "zpipe", "zpipe-4rows-zm3"]),
                         ids=lambda v: {None: "default", 0: "twopass"}.get(v, str(v)))
def test_odd_box_sizes_vs_oracle(oracle_mod, case, n_cell, variant):
    """box sizes that are no multiple of the warp strip (30 cells), the CTA height (6 rows) or the march length"""
    O = oracle_mod
    z, deck_text, _ = load_golden(case)
    ov = [f"amr.n_cell = {n_cell}"]
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines() + ov)))
    o.initialize()
    lbm = new_lbm(deck_text, overrides=ov, variant=variant)
    o.step(5)
    lbm.step(5, want_macrodata=True)
    ref = o.fields()
    sc = scales(ref, lbm.inp.R, lbm.inp.gamma, 1.0 / lbm.inp.dx[0])
    worst, key = compare(lbm.fields(), ref, sc, 5)
    print(f"{case} {n_cell}: worst {worst:.2e} ({key})")
    lbm.close()


@pytest.mark.parametrize("case,nz,world", [("tg12", 16, 2), ("chcyl", None, 2)])
def test_slab_vorticity_matches_single_box(case, nz, world):
    """compute_derived on slabs with the neighbours' macrodata planes exchanged: vorticity of the planes next to
    another rank equals the single-box result (periodic TG: the ring closes; channel: non-periodic... z periodic)"""
    import torch
    from marbles_b200.inputs import lbm_inputs, parse_deck
    from marbles_b200.lbm import LBM, slab_bounds
    from marbles_b200.parallel import LocalSlabs
    z, deck_text, _ = load_golden(case)
    fl = z["is_fluid"].astype(np.int32)
    ov = None
    if nz is not None:
        n = lbm_inputs(parse_deck(text=deck_text)).n_cell
        ov, fl = [f"amr.n_cell = {n[0]} {n[1]} {nz}"], None
    deck = parse_deck(text=deck_text, overrides=ov)
    single = LBM(deck, is_fluid=fl, variant=0)
    single.init_data()
    nzt = single.n_local[2]

    def make(rank, w):
        lo, hi = slab_bounds(nzt, rank, w)
        sub = None
        if fl is not None:
            ng = 3
            full = np.ones((nzt + 2 * ng,) + tuple(d + 2 * ng for d in fl.shape[1:]), dtype=np.int32)
            full[ng:-ng, ng:-ng, ng:-ng] = fl
            full = single._wrap_periodic(full, ng, z_local=True)
            sub = np.ascontiguousarray(full[lo:hi + 1 + 2 * ng])
        s = LBM(deck, rank=rank, world=w, comm=None, is_fluid=sub, variant=0)
        s.init_data()
        return s

    slabs = LocalSlabs(make, world, bool(single.inp.periodic[2]), torch.device("cuda", 0))
    single.step(3, want_macrodata=True)
    slabs.step(3, want_macrodata=True)
    single.compute_derived()
    slabs.compute_derived()
    a, b = single.get_derived(), slabs.gather(lambda s: s.get_derived())
    assert np.abs(a[:4] - b[:4]).max() <= 1e-13 * max(np.abs(a[:4]).max(), 1e-30)
    slabs.close()
    single.close()


@pytest.mark.parametrize("variant", with_experiments([None, "tile"], ["march", "zpipe"]), ids=lambda v: "default" if v is None else str(v))
@pytest.mark.parametrize("n", [256, 512])
def test_full_size_values_tiled_tg_vs_oracle(oracle_mod, n, variant):
    """Value-level pin of the benchmarked size (BASELINE config 3, 512^3; and 256^3): the box is initialised with a
    Taylor-Green state of wavelength 64 cells (omega = n/64), i.e. an (n/64)^3 tiling of one 64^3 period.  After 3
    steps every 64^3 tile must equal the 64^3 oracle run of one period (same dx, same coordinates) on the cells
    further than 4 cells from a tile face: the oracle's 64^3 DOMAIN differences QCorr one-sidedly at its edges
    (Utilities.H:292-309) where the big box differences centrally, and that difference travels one cell per
    stream plus one per gradient.  Covers the 32-bit byte offsets of the default kernels over the whole box."""
    import torch
    O = oracle_mod
    need = 2.6 * 54 * 8 * n ** 3
    if torch.cuda.get_device_properties(0).total_memory < need:
        pytest.skip("not enough device memory")
    _, deck_text, _ = load_golden("tg12")
    t = n // 64
    ov_small = ["amr.n_cell = 64 64 64", "geometry.prob_lo = -1.0 -1.0 -1.0",
                f"geometry.prob_hi = {-1.0 + 2.0 / t!r} {-1.0 + 2.0 / t!r} {-1.0 + 2.0 / t!r}",
                f"ic_taylorgreen.omega = {t}.0 {t}.0 {t}.0"]
    ov_big = [f"amr.n_cell = {n} {n} {n}", f"ic_taylorgreen.omega = {t}.0 {t}.0 {t}.0"]
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines() + ov_small)))
    o.initialize()
    nsteps, m = 3, 5
    o.step(nsteps)
    lbm = new_lbm(deck_text, overrides=ov_big, variant=variant)
    assert abs(lbm.inp.dx[0] - o.p.dx[0]) == 0.0
    lbm.step(nsteps)
    worst = 0.0
    for name, get, ref in (("f", lbm.get_f, o.f_valid), ("g", lbm.get_g, o.g_valid)):
        a = get(0)
        scale = float(np.abs(ref).max())
        for q in range(27):
            tiles = a[q].reshape(t, 64, t, 64, t, 64)[:, m:-m, :, m:-m, :, m:-m]
            r = ref[q][m:-m, m:-m, m:-m][None, :, None, :, None, :]
            worst = max(worst, float(np.abs(tiles - r).max()) / scale)
        del a
    print(f"tiled TG {n}^3: worst {worst:.2e} of scale over {t ** 3} tiles")
    assert worst <= 1e-12 * nsteps, worst
    lbm.close()


@pytest.mark.parametrize("variant", [0, 5], ids=["twopass", "tile"])
def test_lean_halo_is_bit_identical(variant):
    """the lean z-halo (27 of 54 plane-components per lattice: what a neighbour's pull can reach) against the full one
    on an all-periodic all-fluid box cut into three slabs: plain and overlapped slab steps, same bits"""
    import os
    import torch
    from marbles_b200.inputs import parse_deck
    from marbles_b200.lbm import LBM
    from marbles_b200.parallel import LocalSlabs
    _, deck_text, _ = load_golden("tg12")
    deck = parse_deck(text=deck_text, overrides=["amr.n_cell = 12 12 27"])

    def run(lean):
        os.environ["MBL_HALO_LEAN"] = "1" if lean else "0"
        try:
            def make(rank, w):
                s = LBM(deck, rank=rank, world=w, comm=None, variant=variant)
                assert s.halo_lean == lean
                s.init_data()
                return s
            slabs = LocalSlabs(make, 3, True, torch.device("cuda", 0))
        finally:
            os.environ.pop("MBL_HALO_LEAN", None)
        n = int(slabs.slabs[0].lib.mbl_halo_doubles(slabs.slabs[0].ctx, 0))
        slabs.step(3)
        slabs.step_overlapped(3)
        slabs.step(1, want_macrodata=True)
        out = [slabs.gather(lambda s: s.get_f()), slabs.gather(lambda s: s.get_g()), slabs.gather(lambda s: s.get_macrodata())]
        slabs.close()
        return n, out

    n_full, full = run(False)
    n_lean, lean = run(True)
    assert 2 * n_lean == n_full
    for a, b in zip(full, lean):
        assert np.array_equal(a, b)


def test_body_crossing_a_domain_face_from_the_deck():
    """`touch` golden case through the deck path (is_fluid = None: the analytic body evaluated by the host mirror, ghost
    layers included): where a body crosses the outlet and a wall, the flags of the out-of-domain ghost cells decide
    between "pull the boundary value" and "bounce back" """
    z, deck_text, steps = load_golden("touch")
    lbm = new_lbm(deck_text, None)
    lbm.step(steps[-1], want_macrodata=True)
    ref = golden_fields(z, steps[-1])
    sc = scales(ref, lbm.inp.R, lbm.inp.gamma, 1.0 / lbm.inp.dx[0])
    worst, key = compare(lbm.fields(), ref, sc, steps[-1])
    print(f"touch (deck geometry): worst {worst:.2e} ({key})")
    lbm.close()


@pytest.mark.parametrize("case", ["chcyl", "pressure", "slip", "touch"])
def test_device_geometry_matches_host_geometry(case):
    """mbl_set_body (analytic body evaluated on the device, ghost layers and periodic images included) against the
    host mirror's is_fluid uploaded through mbl_set_is_fluid: same flags, hence bit-identical runs"""
    import os
    z, deck_text, steps = load_golden(case)
    a = new_lbm(deck_text, None, variant=0)           # device geometry
    os.environ["MBL_HOST_GEOMETRY"] = "1"
    try:
        b = new_lbm(deck_text, None, variant=0)       # host geometry, uploaded
    finally:
        os.environ.pop("MBL_HOST_GEOMETRY", None)
    assert a._is_fluid_cache is None and b._is_fluid_cache is not None
    a.step(5, want_macrodata=True)
    b.step(5, want_macrodata=True)
    assert np.array_equal(a.get_f(), b.get_f()) and np.array_equal(a.get_g(), b.get_g())
    assert np.array_equal(a.get_macrodata(), b.get_macrodata())
    fa, fb = a.compute_eb_forces(), b.compute_eb_forces()  # block sums meet in atomicAdd: their order is not fixed
    assert np.abs(fa - fb).max() <= 1e-13
    a.close()
    b.close()
