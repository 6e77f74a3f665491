"""N > 1 host logic on CPU: world_size 2 over gloo.

What is under test is the product's slab decomposition and ghost-plane exchange schedule
(marbles_b200.lbm.slab_bounds, marbles_b200.parallel.neighbours / exchange_buffers: posting order,
periodic ring with lower == upper, non-periodic ends).  The per-rank arithmetic is done by the CPU
oracle (the checker) in the reference's own order -- ghost planes are swapped wherever the reference
calls FillBoundary (fillpatch, end of stream) -- and the assembled result must be bit-identical to the
single-box oracle run.  The product's own scheme (ONE exchange of two planes per step, q-correction
of the first ghost plane recomputed locally) has no CPU path; it is covered on the GPU by
tests/test_gpu_parity.py::test_two_slabs_match_single_box."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden

GZ = 2  # ghost planes the product exchanges (marbles_b200/csrc/lattice.cuh)


def _worker(rank, world, case, nsteps, initfile, outdir, nz_override):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C

    from marbles_b200.lbm import slab_bounds
    from marbles_b200.parallel import exchange_buffers, neighbours
    from oracle import oracle as O

    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    z, deck_text, _ = load_golden(case)
    lines = deck_text.splitlines()
    deck = O.parse_deck(None, lines)
    n = [int(v) for v in deck["amr.n_cell"]]
    if nz_override:
        n[2] = nz_override
        deck = O.parse_deck(None, lines + [f"amr.n_cell = {n[0]} {n[1]} {n[2]}"])
    zlo, zhi = slab_bounds(n[2], rank, world)
    s = O.lbm_setup(deck, lo=(0, 0, zlo), hi=(n[0] - 1, n[1] - 1, zhi))
    periodic_z = bool(s.params.periodic[2])
    lower, upper = neighbours(rank, world, periodic_z)
    o = O.Oracle(s)  # all-fluid geometry in these cases (is_fluid of a slab needs the caller's geometry)
    L, p = O.lib(), C.byref(o.p)
    ng = o.ng
    nzl = zhi - zlo + 1

    def exchange():
        for a in (o.f, o.g):
            send_lo = torch.from_numpy(np.ascontiguousarray(a[:, ng:ng + GZ]))
            send_hi = torch.from_numpy(np.ascontiguousarray(a[:, ng + nzl - GZ:ng + nzl]))
            recv_lo, recv_hi = torch.empty_like(send_lo), torch.empty_like(send_hi)
            exchange_buffers(send_lo, send_hi, recv_lo, recv_hi, lower, upper)
            if lower is not None:
                a[:, ng - GZ:ng] = recv_lo.numpy()
            if upper is not None:
                a[:, ng + nzl:ng + nzl + GZ] = recv_hi.numpy()

    # initial state: IC on the grown box (position-dependent, no communication needed)
    L.orc_initialize(p, C.byref(s.ic), O._ptr(o.is_fluid, C.c_int), O._ptr(o.f), O._ptr(o.g))
    for _ in range(nsteps):
        exchange()
        for a, en in ((o.f, 0), (o.g, 1)):
            L.orc_prepass(p, O._ptr(a))
            L.orc_fill_periodic(p, O._ptr(a), 27, ng)
        exchange()  # whole padded planes: carries the neighbours' x/y periodic images too
        for a, en in ((o.f, 0), (o.g, 1)):
            L.orc_physbc(p, O._ptr(a), en, C.c_double(0.0))
        L.orc_stream(p, O._ptr(o.is_fluid, C.c_int), O._ptr(o.f), 1)
        L.orc_stream(p, O._ptr(o.is_fluid, C.c_int), O._ptr(o.g), 1)
        exchange()  # the cross-rank part of the FillBoundary that ends LBM::stream (LBM.cpp:603)
        L.orc_collide(p, O._ptr(o.is_fluid, C.c_int), O._ptr(o.f), O._ptr(o.g), O._ptr(o.macro),
                      O._ptr(o.derived), O._ptr(o.eq), O._ptr(o.eq_g), 0)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), f=o.f_valid, g=o.g_valid, zlo=zlo, zhi=zhi)
    dist.destroy_process_group()


@pytest.mark.parametrize("case,nz,world", [("tg12", 12, 2), ("tg12", 13, 2), ("sod48", 8, 2), ("thermal", 6, 3)])
def test_slab_decomposition_matches_single_box(oracle_mod, case, nz, world):
    O = oracle_mod
    nsteps = 3
    with tempfile.TemporaryDirectory() as tmp:
        initfile = os.path.join(tmp, "init")
        mp.spawn(_worker, args=(world, case, nsteps, initfile, tmp, nz), nprocs=world, join=True)
        parts = [np.load(os.path.join(tmp, f"rank{r}.npz")) for r in range(world)]
    _, deck_text, _ = load_golden(case)
    deck = O.parse_deck(None, deck_text.splitlines())
    n = [int(v) for v in deck["amr.n_cell"]]
    deck = O.parse_deck(None, deck_text.splitlines() + [f"amr.n_cell = {n[0]} {n[1]} {nz}"])
    o = O.Oracle(O.lbm_setup(deck))
    o.initialize()
    o.step(nsteps)
    f = np.concatenate([p["f"] for p in parts], axis=1)
    g = np.concatenate([p["g"] for p in parts], axis=1)
    assert f.shape == o.f_valid.shape
    assert [int(p["zlo"]) for p in parts] == sorted(int(p["zlo"]) for p in parts)
    assert np.array_equal(f, o.f_valid), float(np.abs(f - o.f_valid).max())
    assert np.array_equal(g, o.g_valid), float(np.abs(g - o.g_valid).max())


def test_slab_bounds_and_neighbours():
    from marbles_b200.lbm import slab_bounds
    from marbles_b200.parallel import neighbours
    for nz, world in ((512, 8), (13, 2), (7, 3), (4096, 8)):
        b = [slab_bounds(nz, r, world) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == nz - 1
        assert all(b[r + 1][0] == b[r][1] + 1 for r in range(world - 1))
        sizes = [hi - lo + 1 for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
    assert neighbours(0, 2, True) == (1, 1)
    assert neighbours(0, 4, False) == (None, 1) and neighbours(3, 4, False) == (2, None)
    assert neighbours(0, 1, True) == (0, 0)


def _plot_worker(rank, world, initfile, outdir):
    sys.path.insert(0, ROOT)
    from marbles_b200.lbm import slab_bounds
    from marbles_b200.plotfile import write_plotfile_slabs
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    full = np.load(os.path.join(outdir, "full.npy"))
    nz = full.shape[1]
    zlo, zhi = slab_bounds(nz, rank, world)

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    write_plotfile_slabs(os.path.join(outdir, "plt00007"), ["a", "b", "c"], full[:, zlo:zhi + 1], zlo=zlo, nz_total=nz,
                         rank=rank, gather=gather, time=7.0, step=7, prob_lo=[0, 0, 0], prob_hi=[1, 1, 1], max_grid_size=8)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shape,world", [((3, 20, 12, 10), 2), ((3, 13, 6, 18), 3)])
def test_multirank_plotfile_round_trip(oracle_mod, shape, world):
    """every rank writes the FABs of its slab (Cell_D_<rank>), rank 0 the headers: the oracle's plotfile reader
    must see the full array"""
    rng = np.random.default_rng(3)
    full = rng.standard_normal(shape)
    with tempfile.TemporaryDirectory() as tmp:
        np.save(os.path.join(tmp, "full.npy"), full)
        initfile = os.path.join(tmp, "init")
        mp.spawn(_plot_worker, args=(world, initfile, tmp), nprocs=world, join=True)
        pf = oracle_mod.read_plotfile(os.path.join(tmp, "plt00007"))
        assert sorted(os.listdir(os.path.join(tmp, "plt00007", "Level_0"))) == \
            sorted(["Cell_H"] + [f"Cell_D_{r:05d}" for r in range(world)])
        assert pf["__names__"] == ["a", "b", "c"] and pf["__time__"] == 7.0
        for c, name in enumerate(pf["__names__"]):
            assert np.array_equal(pf[name], full[c])


def _chk_worker(rank, world, initfile, outdir):
    sys.path.insert(0, ROOT)
    from marbles_b200.lbm import slab_bounds
    from marbles_b200 import plotfile as P
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    full = np.load(os.path.join(outdir, "full.npz"))
    nz = full["f"].shape[1]
    zlo, zhi = slab_bounds(nz, rank, world)

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    P.write_checkpoint_slabs(os.path.join(outdir, "chk00004"), full["f"][:, zlo:zhi + 1], full["g"][:, zlo:zhi + 1], zlo=zlo,
                             nz_total=nz, rank=rank, gather=gather, step=4, dt=1.0, time=4.0, periodic=[1, 1, 1],
                             max_grid_size=8)
    dist.barrier()
    mine = P.read_checkpoint_slab(os.path.join(outdir, "chk00004"), zlo, zhi)  # and back, slab by slab
    assert np.array_equal(mine["f"], full["f"][:, zlo:zhi + 1]) and np.array_equal(mine["g"], full["g"][:, zlo:zhi + 1])
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_multirank_checkpoint_restarts_the_reference(oracle_mod, world):
    """a checkpoint written by several ranks (one data file per rank): readable as a whole and slab by slab, and
    the unmodified reference restarts from it and reproduces its uninterrupted run"""
    import subprocess
    from marbles_b200 import plotfile as P
    O = oracle_mod
    z, deck_text, _ = load_golden("tg12")
    o = O.Oracle(O.lbm_setup(O.parse_deck(None, deck_text.splitlines())))
    o.initialize()
    o.step(4)
    fl = o.fields()
    f = np.stack([fl[f"f_{q:02d}"] for q in range(27)])
    g = np.stack([fl[f"g_{q:02d}"] for q in range(27)])
    with tempfile.TemporaryDirectory() as tmp:
        np.savez(os.path.join(tmp, "full.npz"), f=f, g=g)
        mp.spawn(_chk_worker, args=(world, os.path.join(tmp, "init"), tmp), nprocs=world, join=True)
        c = P.read_checkpoint(os.path.join(tmp, "chk00004"))
        assert c["step"] == 4 and np.array_equal(c["f"], f) and np.array_equal(c["g"], g)
        if not os.path.exists(O.REF_SERIAL):
            pytest.skip("reference executable not built")
        with open(os.path.join(tmp, "tg.inp"), "w") as fh:
            fh.write(deck_text)
        subprocess.run([O.REF_SERIAL, "tg.inp", "max_step=10", "amr.plot_int=10", "amr.chk_int=-1", "amr.restart=chk00004",
                        "lbm.save_streaming=1"], cwd=tmp, check=True, capture_output=True)
        got = O.read_plotfile(os.path.join(tmp, "plt00010"))
        for q in range(27):
            assert np.array_equal(got[f"f_{q:02d}"], z[f"s10_f_{q:02d}"]) and np.array_equal(got[f"g_{q:02d}"], z[f"s10_g_{q:02d}"])
