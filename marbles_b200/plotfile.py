"""AMReX plotfile emission for the device-resident state.

`LBM::write_plot_file` (Source/LBM.cpp:1677-1690) hands `plot_file_mf()` (Source/LBM.cpp:1629-1675: macrodata,
then f and g if lbm.save_streaming, then the derived fields if lbm.save_derived, then the two is_fluid
components) to `amrex::WriteMultiLevelPlotfile`.  This module writes the same on-disk format -- `Header`,
`Level_0/Cell_H`, `Level_0/Cell_D_00000` with native little-endian doubles -- byte for byte for a single level
(checked against files written by the unmodified reference, tests/test_plotfile.py), so post-processing
(fcompare, yt, the reference's Tools/) keeps working when the step runs on the GPU.

Host side only: the fields come from `LBM.fields()` / `get_derived()` (device -> host copies of the C ABI).
"""
from __future__ import annotations

import os

import numpy as np

MACRO_NAMES = ["rho", "vel_x", "vel_y", "vel_z", "vel_mag", "two_rho_e", "QCorrX", "QCorrY", "QCorrZ", "pxx", "pyy", "pzz",
               "pxy", "pxz", "pyz", "qx", "qy", "qz", "temperature"]  # Source/LBM.cpp:302-340, Constants.H:8-31
DERIVED_NAMES = ["vort_x", "vort_y", "vort_z", "vort_mag", "dQCorrX", "dQCorrY", "dQCorrZ"]  # Constants.H:39-47
IS_FLUID_NAMES = ["is_fluid", "eb_boundary"]
FAB_HEADER = "FAB ((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))"  # IEEE double, little endian


def plot_file_var_names(save_streaming: bool = True, save_derived: bool = True) -> list[str]:
    """LBM::plot_file_var_names (the order plot_file_mf copies the components in)"""
    names = list(MACRO_NAMES)
    if save_streaming:
        names += [f"f_{q:02d}" for q in range(27)] + [f"g_{q:02d}" for q in range(27)]
    if save_derived:
        names += DERIVED_NAMES
    return names + IS_FLUID_NAMES


def _max_size_cuts(n: int, chunk: int):
    """BoxList::maxSize for one direction (AMReX_BoxList.cpp:765-818): halve block size and length together while
    both are even, cut the remaining length into ceil(len / block) pieces whose sizes differ by at most one
    (the longer ones first), scale back.  Returns inclusive (lo, hi) pairs."""
    if n <= chunk:
        return [(0, n - 1)]
    ratio, bs, nlen = 1, chunk, n
    while bs % 2 == 0 and nlen % 2 == 0:
        ratio, bs, nlen = ratio * 2, bs // 2, nlen // 2
    numblk = (nlen + bs - 1) // bs
    sz = nlen // numblk
    extra = nlen - sz * numblk
    cuts, lo = [], 0
    for b in range(numblk):
        ln = (sz + 1 if b < extra else sz) * ratio
        cuts.append((lo, lo + ln - 1))
        lo += ln
    return cuts


def chop_boxes(n_cell, max_grid_size: int):
    """The level-0 BoxArray as AmrMesh::MakeBaseGrids builds it on one process (AMReX_AmrMesh.cpp): the domain is
    coarsened by 2 in every direction with an even number of cells, chopped with BoxList::maxSize(max_grid_size / 2)
    and refined again; boxes are listed x fastest."""
    cuts = []
    for d in range(3):
        n = int(n_cell[d])
        fac = 2 if n % 2 == 0 else 1
        cuts.append([(lo * fac, (hi + 1) * fac - 1) for lo, hi in _max_size_cuts(n // fac, max(1, max_grid_size // fac))])
    boxes = []
    for kz in cuts[2]:
        for jy in cuts[1]:
            for ix in cuts[0]:
                boxes.append(((ix[0], jy[0], kz[0]), (ix[1], jy[1], kz[1])))
    return boxes


def _g17(v: float) -> str:
    """AMReX writes the reals of the Header with setprecision(17), default float format"""
    return "%.17g" % float(v)


def _box(lo, hi) -> str:
    return f"(({lo[0]},{lo[1]},{lo[2]}) ({hi[0]},{hi[1]},{hi[2]}) (0,0,0))"


def write_plotfile(path: str, names: list[str], data: np.ndarray, *, time: float, step: int, prob_lo, prob_hi,
                   max_grid_size: int = 32) -> None:
    """Write a single-level cell-centred plotfile.  data: [ncomp, nz, ny, nx] float64."""
    data = np.asarray(data, dtype=np.float64)
    ncomp, nz, ny, nx = data.shape
    assert ncomp == len(names)
    n = (nx, ny, nz)
    dx = [(float(prob_hi[d]) - float(prob_lo[d])) / n[d] for d in range(3)]
    boxes = chop_boxes(n, max_grid_size)
    lev_dir = os.path.join(path, "Level_0")
    os.makedirs(lev_dir, exist_ok=True)

    # ---- Cell_D_00000 + per-box offsets, minima, maxima --------------------------------------------------
    offsets, mins, maxs = [], [], []
    with open(os.path.join(lev_dir, "Cell_D_00000"), "wb") as fh:
        for lo, hi in boxes:
            offsets.append(fh.tell())
            sub = np.ascontiguousarray(data[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1])
            fh.write(f"{FAB_HEADER}{_box(lo, hi)} {ncomp}\n".encode())
            fh.write(sub.astype("<f8").tobytes())
            flat = sub.reshape(ncomp, -1)
            mins.append(flat.min(axis=1))
            maxs.append(flat.max(axis=1))

    # ---- Cell_H (VisMF header, version 1 = with min/max tables) -------------------------------------------
    with open(os.path.join(lev_dir, "Cell_H"), "w") as fh:
        fh.write(f"1\n1\n{ncomp}\n0\n")
        fh.write(f"({len(boxes)} 0\n")
        for lo, hi in boxes:
            fh.write(_box(lo, hi) + "\n")
        fh.write(")\n")
        fh.write(f"{len(boxes)}\n")
        for off in offsets:
            fh.write(f"FabOnDisk: Cell_D_00000 {off}\n")
        fh.write("\n")
        for table in (mins, maxs):
            fh.write(f"{len(boxes)},{ncomp}\n")
            for row in table:
                fh.write("".join("%.17e," % v for v in row) + "\n")
            fh.write("\n")

    # ---- Header (amrex::WriteGenericPlotfileHeader, AMReX_PlotFileUtil.cpp) --------------------------------
    with open(os.path.join(path, "Header"), "w") as fh:
        fh.write("HyperCLaw-V1.1\n")
        fh.write(f"{ncomp}\n")
        for nm in names:
            fh.write(nm + "\n")
        fh.write("3\n")
        fh.write(_g17(time) + "\n")
        fh.write("0\n")  # finest level
        fh.write(" ".join(_g17(v) for v in prob_lo) + " \n")
        fh.write(" ".join(_g17(v) for v in prob_hi) + " \n")
        fh.write("\n")  # refinement ratios: none for a single level
        fh.write(_box((0, 0, 0), (nx - 1, ny - 1, nz - 1)) + " \n")
        fh.write(f"{step} \n")
        fh.write(" ".join(_g17(v) for v in dx) + " \n")
        fh.write("0\n0\n")  # coordinate system, boundary width
        fh.write(f"0 {len(boxes)} {_g17(time)}\n")
        fh.write(f"{step}\n")
        for lo, hi in boxes:
            for d in range(3):
                fh.write(f"{_g17(prob_lo[d] + lo[d] * dx[d])} {_g17(prob_lo[d] + (hi[d] + 1) * dx[d])}\n")
        fh.write("Level_0/Cell\n")


def write_plotfile_levels(path: str, names: list[str], levels, *, time: float, level_steps, prob_lo, prob_hi, n_cell,
                          ref_ratio: int = 2, rank: int = 0, owners=None, gather=None) -> None:
    """amrex::WriteMultiLevelPlotfile for a hierarchy (Source/LBM.cpp:1677-1690; AMReX_PlotFileUtil.cpp): levels[lev] =
    (boxes, fabs) with boxes = [(lo, hi), ...] in the index space of level lev and fabs[i] = [ncomp, nz, ny, nx] float64 of
    box i (the level's BoxArray order); level_steps[lev] = m_isteps[lev]; n_cell = cells of level 0.  One process:
    every level's FABs go to Level_<lev>/Cell_D_00000.  Distributed levels (owners[lev][i] = rank of box i, fabs[i] only
    for this rank's boxes, `gather(obj) -> [obj of rank 0, obj of rank 1, ...]` the only communication): every rank writes
    its FABs into Level_<lev>/Cell_D_<rank> and rank 0 the headers, the layout VisMF uses when every rank owns a file."""
    ncomp, nlev = len(names), len(levels)
    dx0 = [(float(prob_hi[d]) - float(prob_lo[d])) / n_cell[d] for d in range(3)]
    mine_all = []
    for lev, (boxes, fabs) in enumerate(levels):
        assert len(boxes) == len(fabs)
        lev_dir = os.path.join(path, f"Level_{lev}")
        os.makedirs(lev_dir, exist_ok=True)
        own = owners[lev] if owners is not None else [rank] * len(boxes)
        mine = {}
        if any(o == rank for o in own):
            with open(os.path.join(lev_dir, f"Cell_D_{rank:05d}"), "wb") as fh:
                for ib, ((lo, hi), fab) in enumerate(zip(boxes, fabs)):
                    if own[ib] != rank:
                        continue
                    fab = np.ascontiguousarray(fab, dtype="<f8")
                    assert fab.shape == (ncomp, hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1), (fab.shape, lo, hi)
                    off = fh.tell()
                    fh.write(f"{FAB_HEADER}{_box(lo, hi)} {ncomp}\n".encode())
                    fh.write(fab.tobytes())
                    flat = fab.reshape(ncomp, -1)
                    mine[ib] = (off, flat.min(axis=1).tolist(), flat.max(axis=1).tolist())
        mine_all.append(mine)
    parts = gather(mine_all) if gather is not None else [mine_all]
    if rank != 0:
        return
    for lev, (boxes, _) in enumerate(levels):
        lev_dir = os.path.join(path, f"Level_{lev}")
        own = owners[lev] if owners is not None else [0] * len(boxes)
        rec = [parts[own[ib]][lev][ib] for ib in range(len(boxes))]
        with open(os.path.join(lev_dir, "Cell_H"), "w") as fh:
            fh.write(f"1\n1\n{ncomp}\n0\n")
            fh.write(f"({len(boxes)} 0\n")
            for lo, hi in boxes:
                fh.write(_box(lo, hi) + "\n")
            fh.write(")\n")
            fh.write(f"{len(boxes)}\n")
            for ib, r in enumerate(rec):
                fh.write(f"FabOnDisk: Cell_D_{own[ib]:05d} {r[0]}\n")
            fh.write("\n")
            for col in (1, 2):
                fh.write(f"{len(boxes)},{ncomp}\n")
                for r in rec:
                    fh.write("".join("%.17e," % v for v in r[col]) + "\n")
                fh.write("\n")
    with open(os.path.join(path, "Header"), "w") as fh:
        fh.write("HyperCLaw-V1.1\n")
        fh.write(f"{ncomp}\n")
        for nm in names:
            fh.write(nm + "\n")
        fh.write("3\n")
        fh.write(_g17(time) + "\n")
        fh.write(f"{nlev - 1}\n")
        fh.write(" ".join(_g17(v) for v in prob_lo) + " \n")
        fh.write(" ".join(_g17(v) for v in prob_hi) + " \n")
        fh.write("".join(f"{ref_ratio} " for _ in range(nlev - 1)) + "\n")
        fh.write("".join(_box((0, 0, 0), tuple(n_cell[d] * ref_ratio ** lev - 1 for d in range(3))) + " " for lev in range(nlev)) + "\n")
        fh.write("".join(f"{int(st)} " for st in level_steps) + "\n")
        for lev in range(nlev):
            fh.write(" ".join(_g17(dx0[d] / ref_ratio ** lev) for d in range(3)) + " \n")
        fh.write("0\n0\n")  # coordinate system, boundary width
        for lev, (boxes, _) in enumerate(levels):
            dx = [dx0[d] / ref_ratio ** lev for d in range(3)]
            fh.write(f"{lev} {len(boxes)} {_g17(time)}\n")
            fh.write(f"{int(level_steps[lev])}\n")
            for lo, hi in boxes:
                for d in range(3):
                    fh.write(f"{_g17(prob_lo[d] + lo[d] * dx[d])} {_g17(prob_lo[d] + (hi[d] + 1) * dx[d])}\n")
            fh.write(f"Level_{lev}/Cell\n")


def write_amr_plotfile(amr, directory: str = ".", prefix: str = "plt", save_streaming: bool | None = None,
                       save_derived: bool | None = None, digits: int = 5) -> str:
    """LBM::write_plot_file for a multi-level state (marbles_b200.amr.AmrLBM, all boxes on this rank): the macrodata of
    every level must be current (last step with want_macrodata=True) and compute_derived() called, as post_time_step
    leaves them.  With distributed levels a collective call: every rank writes the FABs it holds."""
    deck = amr.inp.deck

    def deck_int(key: str, default: int) -> int:
        v = deck.get(key, default)
        if isinstance(v, (list, tuple)):
            v = v[0]
        return int(str(v).split()[0])

    if save_streaming is None:
        save_streaming = bool(deck_int("lbm.save_streaming", 1))
    if save_derived is None:
        save_derived = bool(deck_int("lbm.save_derived", 1))
    names = plot_file_var_names(save_streaming, save_derived)
    levels = []
    for lev in range(amr.finest + 1):
        fabs = []
        for ib in range(len(amr.boxes[lev])):
            if not amr.is_local(lev, ib):
                fabs.append(None)  # another rank writes it
                continue
            parts = [amr.get_box_macrodata(lev, ib, derived=False)]
            if save_streaming:
                parts += [amr.get_box(lev, ib, 0), amr.get_box(lev, ib, 1)]
            if save_derived:
                parts.append(amr.get_box_macrodata(lev, ib, derived=True))
            fl = amr.box_is_fluid(lev, ib)  # grown by 3
            parts.append(np.stack([fl[3:-3, 3:-3, 3:-3].astype(np.float64), eb_boundary(fl, 3).astype(np.float64)]))
            fabs.append(np.concatenate(parts, axis=0))
        levels.append((amr.boxes[lev], fabs))
    path = os.path.join(directory, plot_file_name(prefix, amr.isteps, digits))
    gather = None
    if amr.world > 1:  # collective call: every rank writes its FABs, rank 0 the headers
        import torch.distributed as dist

        def gather(obj):
            out = [None] * amr.world
            dist.all_gather_object(out, obj)
            return out

    write_plotfile_levels(path, names, levels, time=amr.time, level_steps=[amr.isteps * 2 ** l for l in range(amr.finest + 1)],
                          prob_lo=amr.inp.prob_lo, prob_hi=amr.inp.prob_hi, n_cell=amr.inp.n_cell, rank=amr.rank,
                          owners=amr.owner if amr.world > 1 else None, gather=gather)
    return path


def plot_file_name(prefix: str, step: int, digits: int = 5) -> str:
    """amrex::Concatenate(plot_file, step, m_file_name_digits) (Source/LBM.cpp:1617-1621; amr.file_name_digits)"""
    return f"{prefix}{step:0{digits}d}"


def write_lbm_plotfile(lbm, directory: str = ".", prefix: str = "plt", max_grid_size: int | None = None,
                       save_streaming: bool | None = None, save_derived: bool | None = None, digits: int = 5) -> str:
    """LBM::write_plot_file: the macrodata must be current (last step taken with want_macrodata=True, as the
    reference's post_time_step leaves it).  On several ranks every rank writes the FABs of its z-slab and rank 0 the
    headers (collective call; compute_derived exchanges the neighbours' macrodata planes first)."""
    deck = lbm.inp.deck

    def deck_int(key: str, default: int) -> int:
        v = deck.get(key, default)  # parse_deck keeps a list of tokens per key
        if isinstance(v, (list, tuple)):
            v = v[0]
        return int(str(v).split()[0])

    if save_streaming is None:
        save_streaming = bool(deck_int("lbm.save_streaming", 1))
    if save_derived is None:
        save_derived = bool(deck_int("lbm.save_derived", 1))
    if max_grid_size is None:
        max_grid_size = deck_int("amr.max_grid_size", 32)
    names = plot_file_var_names(save_streaming, save_derived)
    parts = [lbm.get_macrodata()]
    if save_streaming:
        parts += [lbm.get_f(), lbm.get_g()]
    if save_derived:
        lbm.compute_derived()
        parts.append(lbm.get_derived())
    ng = (lbm._is_fluid.shape[0] - lbm.n_local[2]) // 2
    fl = lbm._is_fluid[ng:-ng, ng:-ng, ng:-ng] if ng else lbm._is_fluid
    parts.append(np.stack([fl.astype(np.float64), eb_boundary(lbm._is_fluid, ng).astype(np.float64)]))
    data = np.concatenate(parts, axis=0)
    path = os.path.join(directory, plot_file_name(prefix, lbm.isteps, digits))
    if lbm.world == 1:
        write_plotfile(path, names, data, time=lbm.time, step=lbm.isteps, prob_lo=lbm.inp.prob_lo,
                       prob_hi=lbm.inp.prob_hi, max_grid_size=max_grid_size)
    else:  # one z-slab per rank: every rank writes its own FABs, rank 0 the two headers
        import torch.distributed as dist

        def gather(obj):
            out = [None] * lbm.world
            dist.all_gather_object(out, obj)
            return out

        write_plotfile_slabs(path, names, data, zlo=lbm.lo[2], nz_total=lbm.inp.n_cell[2], rank=lbm.rank, gather=gather,
                             time=lbm.time, step=lbm.isteps, prob_lo=lbm.inp.prob_lo, prob_hi=lbm.inp.prob_hi,
                             max_grid_size=max_grid_size)
    return path


def eb_boundary(is_fluid_grown: np.ndarray, ng: int) -> np.ndarray:
    """component 1 of m_is_fluid (Source/LBM.cpp:1244-1259): a solid cell with at least one fluid face neighbour"""
    a = is_fluid_grown
    if ng == 0:
        a = np.pad(a, 1, mode="edge")
        ng = 1
    c = a[ng:-ng, ng:-ng, ng:-ng]
    out = np.zeros_like(c)
    sl = lambda o, n: slice(ng + o, n + ng + o)
    nz, ny, nx = c.shape
    any_fluid = np.zeros(c.shape, dtype=bool)
    for dz, dy, dxx in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)):
        any_fluid |= a[sl(dz, nz), sl(dy, ny), sl(dxx, nx)] == 1
    out[(c == 0) & any_fluid] = 1
    return out


# ---------------------------------------------------------------------------------------------------------
# Checkpoints (LBM::write_checkpoint_file / read_checkpoint_file, Source/LBM.cpp:1692-1915): a text Header and
# the two lattices as VisMF files WITH their ghost cells, `chkNNNNN/Level_0/f_00_{H,D_00000}` and `g_00_...`.
# A checkpoint written here restarts the unmodified reference (amr.restart=...), and the reference's checkpoints
# load into `LBM` (tests/test_checkpoint.py runs both directions against the reference executable).
# ---------------------------------------------------------------------------------------------------------
def _grown(lo, hi, ng):
    return tuple(v - ng for v in lo), tuple(v + ng for v in hi)


def _write_vismf_data(lev_dir: str, dname: str, data: np.ndarray, ng: int, boxes, zoff: int = 0) -> dict:
    """the FABs of `boxes` (global index space; data starts at global plane zoff, grown by ng) into one data file;
    returns what the VisMF header needs"""
    ncomp = data.shape[0]
    os.makedirs(lev_dir, exist_ok=True)
    offsets, mins, maxs = [], [], []
    with open(os.path.join(lev_dir, dname), "wb") as fh:
        for lo, hi in boxes:
            offsets.append(fh.tell())
            glo, ghi = _grown(lo, hi, ng)
            z0, z1 = lo[2] - zoff, hi[2] - zoff
            sub = np.ascontiguousarray(data[:, z0:z1 + 1 + 2 * ng, lo[1]:hi[1] + 1 + 2 * ng, lo[0]:hi[0] + 1 + 2 * ng])
            fh.write(f"{FAB_HEADER}{_box(glo, ghi)} {ncomp}\n".encode())
            fh.write(sub.astype("<f8").tobytes())
            val = sub[:, ng:sub.shape[1] - ng, ng:sub.shape[2] - ng, ng:sub.shape[3] - ng] if ng else sub
            flat = val.reshape(ncomp, -1)
            mins.append(flat.min(axis=1).tolist())
            maxs.append(flat.max(axis=1).tolist())
    return {"file": dname, "boxes": [(tuple(lo), tuple(hi)) for lo, hi in boxes], "offsets": offsets, "mins": mins,
            "maxs": maxs}


def _write_vismf_header(lev_dir: str, prefix: str, ncomp: int, ng: int, parts) -> None:
    boxes = [b for p in parts for b in p["boxes"]]
    with open(os.path.join(lev_dir, f"{prefix}_H"), "w") as fh:
        fh.write(f"1\n1\n{ncomp}\n{ng}\n")
        fh.write(f"({len(boxes)} 0\n")
        for lo, hi in boxes:
            fh.write(_box(lo, hi) + "\n")
        fh.write(")\n")
        fh.write(f"{len(boxes)}\n")
        for p in parts:
            for off in p["offsets"]:
                fh.write(f"FabOnDisk: {p['file']} {off}\n")
        fh.write("\n")
        for key in ("mins", "maxs"):
            fh.write(f"{len(boxes)},{ncomp}\n")
            for p in parts:
                for row in p[key]:
                    fh.write("".join("%.17e," % v for v in row) + "\n")
            fh.write("\n")


def write_vismf(lev_dir: str, prefix: str, data: np.ndarray, ng: int, boxes) -> None:
    """VisMF::Write of one MultiFab: data is [ncomp, nz+2ng, ny+2ng, nx+2ng] over the domain grown by ng; every
    FAB goes out with its own ghost cells (taken from the grown array), minima / maxima over the valid box."""
    part = _write_vismf_data(lev_dir, f"{prefix}_D_00000", data, ng, boxes)
    _write_vismf_header(lev_dir, prefix, data.shape[0], ng, [part])


def read_vismf(lev_dir: str, prefix: str, zrange=None):
    """-> (data [ncomp, nz+2ng, ny+2ng, nx+2ng] over the bounding box of the FABs read, grown by ng, ng, boxes).
    Cells covered by several FABs (ghost overlap) take the VALID cell's value.  zrange = (zlo, zhi): only the FABs
    whose valid box lies inside those planes (a rank reading its own slab of a checkpoint)."""
    import re
    with open(os.path.join(lev_dir, f"{prefix}_H")) as fh:
        ch = fh.read().split("\n")
    ncomp, ng = int(ch[2]), int(ch[3].split()[0].strip("(),"))
    i = next(k for k, l in enumerate(ch) if l.startswith("(") and k >= 4)
    nbox = int(ch[i].strip("(").split()[0])
    boxes = []
    for b in range(nbox):
        m = [int(v) for v in re.findall(r"-?\d+", ch[i + 1 + b])]
        boxes.append((tuple(m[0:3]), tuple(m[3:6])))
    j = next(k for k, l in enumerate(ch) if l.startswith("FabOnDisk"))
    fod = [ch[j + b] for b in range(nbox)]
    if zrange is not None:
        keep = [b for b in range(nbox) if boxes[b][1][2] >= zrange[0] and boxes[b][0][2] <= zrange[1]]
        boxes, fod = [boxes[b] for b in keep], [fod[b] for b in keep]
    dlo = [min(b[0][d] for b in boxes) for d in range(3)]
    dhi = [max(b[1][d] for b in boxes) for d in range(3)]
    n = [dhi[d] - dlo[d] + 1 for d in range(3)]
    data = np.zeros((ncomp, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng))
    fabs = []
    for b, (lo, hi) in enumerate(boxes):
        _, fname, off = fod[b].split()
        bn = [hi[d] - lo[d] + 1 + 2 * ng for d in range(3)]
        with open(os.path.join(lev_dir, fname), "rb") as fh:
            fh.seek(int(off))
            fh.readline()
            arr = np.frombuffer(fh.read(8 * ncomp * bn[0] * bn[1] * bn[2]), dtype="<f8").reshape(ncomp, bn[2], bn[1], bn[0])
        fabs.append(arr)
        o = [lo[d] - dlo[d] for d in range(3)]
        data[:, o[2]:o[2] + bn[2], o[1]:o[1] + bn[1], o[0]:o[0] + bn[0]] = arr
    for (lo, hi), arr in zip(boxes, fabs):  # valid cells win over other FABs' ghost copies
        o = [lo[d] - dlo[d] + ng for d in range(3)]
        bn = [hi[d] - lo[d] + 1 for d in range(3)]
        data[:, o[2]:o[2] + bn[2], o[1]:o[1] + bn[1], o[0]:o[0] + bn[0]] = \
            arr[:, ng:ng + bn[2], ng:ng + bn[1], ng:ng + bn[0]]
    return data, ng, boxes


def chk_file_name(prefix: str, step: int, digits: int = 5) -> str:
    return f"{prefix}{step:0{digits}d}"


def write_checkpoint(path: str, f: np.ndarray, g: np.ndarray, *, step: int, dt: float, time: float, periodic,
                     max_grid_size: int = 32, ng: int = 3) -> None:
    """f, g: [27, nz, ny, nx] valid cells.  Ghost cells of the files hold the periodic images in periodic
    directions and the nearest valid cell elsewhere; the reference refills every ghost cell after a restart
    (FillBoundary in read_checkpoint_file, fillpatch at the start of each step)."""
    n = (f.shape[3], f.shape[2], f.shape[1])
    boxes = chop_boxes(n, max_grid_size)

    def grow(a):
        for axis, d in ((3, 0), (2, 1), (1, 2)):
            pad = [(0, 0)] * 4
            pad[axis] = (ng, ng)
            a = np.pad(a, pad, mode="wrap" if periodic[d] else "edge")
        return a

    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "Header"), "w") as fh:
        fh.write("Checkpoint file for LBM\n0\n")
        fh.write(f"{step} \n{_g17(dt)} \n{_g17(time)} \n")
        fh.write(f"({len(boxes)} 0\n")
        for lo, hi in boxes:
            fh.write(_box(lo, hi) + "\n")
        fh.write(")\n")
    lev_dir = os.path.join(path, "Level_0")
    write_vismf(lev_dir, "f_00", grow(np.asarray(f, dtype=np.float64)), ng, boxes)
    write_vismf(lev_dir, "g_00", grow(np.asarray(g, dtype=np.float64)), ng, boxes)


def read_checkpoint(path: str) -> dict:
    """-> {"step", "dt", "time", "f", "g"} with f, g the valid cells [27, nz, ny, nx] of level 0"""
    with open(os.path.join(path, "Header")) as fh:
        lines = fh.read().split("\n")
    if int(lines[1]) != 0:
        raise ValueError("only single-level checkpoints are supported")
    out = {"step": int(lines[2].split()[0]), "dt": float(lines[3].split()[0]), "time": float(lines[4].split()[0])}
    for name in ("f", "g"):
        data, ng, _ = read_vismf(os.path.join(path, "Level_0"), f"{name}_00")
        out[name] = np.ascontiguousarray(data[:, ng:data.shape[1] - ng, ng:data.shape[2] - ng, ng:data.shape[3] - ng]
                                         if ng else data)
    return out


def _write_vismf_fabs(lev_dir: str, prefix: str, fabs, ng: int, boxes, rank: int = 0, own=None, gather=None) -> None:
    """VisMF::Write of a MultiFab given FAB by FAB: fabs[i] = [ncomp, nz+2ng, ny+2ng, nx+2ng] of box i with its ghost
    cells; minima / maxima over the valid box.  Distributed (own[i] = rank of box i, fabs[i] only for this rank's boxes):
    every rank writes `<prefix>_D_<rank>`, rank 0 the header from the gathered offsets and extrema."""
    os.makedirs(lev_dir, exist_ok=True)
    own = own if own is not None else [rank] * len(boxes)
    mine, ncomp = {}, None
    if any(o == rank for o in own):
        with open(os.path.join(lev_dir, f"{prefix}_D_{rank:05d}"), "wb") as fh:
            for ib, ((lo, hi), fab) in enumerate(zip(boxes, fabs)):
                if own[ib] != rank:
                    continue
                fab = np.ascontiguousarray(fab, dtype="<f8")
                ncomp = fab.shape[0]
                assert fab.shape == (ncomp,) + tuple(hi[d] - lo[d] + 1 + 2 * ng for d in (2, 1, 0))
                off = fh.tell()
                glo, ghi = _grown(lo, hi, ng)
                fh.write(f"{FAB_HEADER}{_box(glo, ghi)} {ncomp}\n".encode())
                fh.write(fab.tobytes())
                val = fab[:, ng:fab.shape[1] - ng, ng:fab.shape[2] - ng, ng:fab.shape[3] - ng] if ng else fab
                flat = val.reshape(ncomp, -1)
                mine[ib] = (off, flat.min(axis=1).tolist(), flat.max(axis=1).tolist(), ncomp)
    parts = gather(mine) if gather is not None else [mine]
    if rank != 0:
        return
    rec = [parts[own[ib]][ib] for ib in range(len(boxes))]
    ncomp = rec[0][3]
    with open(os.path.join(lev_dir, f"{prefix}_H"), "w") as fh:
        fh.write(f"1\n1\n{ncomp}\n{ng}\n")
        fh.write(f"({len(boxes)} 0\n")
        for lo, hi in boxes:
            fh.write(_box(lo, hi) + "\n")
        fh.write(")\n")
        fh.write(f"{len(boxes)}\n")
        for ib, r in enumerate(rec):
            fh.write(f"FabOnDisk: {prefix}_D_{own[ib]:05d} {r[0]}\n")
        fh.write("\n")
        for col in (1, 2):
            fh.write(f"{len(boxes)},{ncomp}\n")
            for r in rec:
                fh.write("".join("%.17e," % v for v in r[col]) + "\n")
            fh.write("\n")


def _read_vismf_fabs(lev_dir: str, prefix: str):
    """-> (fabs [ncomp, nz+2ng, ny+2ng, nx+2ng] per box, ng, boxes)"""
    import re
    with open(os.path.join(lev_dir, f"{prefix}_H")) as fh:
        ch = fh.read().split("\n")
    ncomp, ng = int(ch[2]), int(ch[3].split()[0].strip("(),"))
    i = next(k for k, l in enumerate(ch) if l.startswith("(") and k >= 4)
    nbox = int(ch[i].strip("(").split()[0])
    boxes = []
    for b in range(nbox):
        m = [int(v) for v in re.findall(r"-?\d+", ch[i + 1 + b])]
        boxes.append((tuple(m[0:3]), tuple(m[3:6])))
    j = next(k for k, l in enumerate(ch) if l.startswith("FabOnDisk"))
    fabs = []
    for b, (lo, hi) in enumerate(boxes):
        _, fname, off = ch[j + b].split()
        bn = [hi[d] - lo[d] + 1 + 2 * ng for d in range(3)]
        with open(os.path.join(lev_dir, fname), "rb") as fh:
            fh.seek(int(off))
            fh.readline()
            fabs.append(np.frombuffer(fh.read(8 * ncomp * bn[0] * bn[1] * bn[2]), dtype="<f8")
                        .reshape(ncomp, bn[2], bn[1], bn[0]).copy())
    return fabs, ng, boxes


def write_checkpoint_levels(path: str, levels, *, isteps, dts, times, ng: int = 3, rank: int = 0, owners=None,
                            gather=None) -> None:
    """LBM::write_checkpoint_file (Source/LBM.cpp:1692-1783) for a hierarchy: levels[lev] = (boxes, f_fabs, g_fabs) with
    the FABs of every box INCLUDING their ng ghost cells (VisMF::Write stores them); isteps / dts / times are the
    reference's m_isteps / m_dts / m_ts_new, one entry per level up to amr.max_level (levels that do not exist keep
    their initial values there).  Distributed levels: owners[lev][i] = rank of box i, FABs only for this rank's boxes,
    `gather` as in write_plotfile_levels; every rank writes its data files, rank 0 the headers."""
    os.makedirs(path, exist_ok=True)
    if rank == 0:
      with open(os.path.join(path, "Header"), "w") as fh:
        fh.write(f"Checkpoint file for LBM\n{len(levels) - 1}\n")
        fh.write("".join(f"{int(v)} " for v in isteps) + "\n")
        fh.write("".join(f"{_g17(v)} " for v in dts) + "\n")
        fh.write("".join(f"{_g17(v)} " for v in times) + "\n")
        for boxes, _, _ in levels:
            fh.write(f"({len(boxes)} 0\n")
            for lo, hi in boxes:
                fh.write(_box(lo, hi) + "\n")
            fh.write(")\n")
    for lev, (boxes, ff, gg) in enumerate(levels):
        lev_dir = os.path.join(path, f"Level_{lev}")
        own = owners[lev] if owners is not None else None
        _write_vismf_fabs(lev_dir, "f_00", ff, ng, boxes, rank, own, gather)
        _write_vismf_fabs(lev_dir, "g_00", gg, ng, boxes, rank, own, gather)


def read_checkpoint_levels(path: str) -> dict:
    """-> {"isteps", "dts", "times", "levels": [(boxes, f_fabs, g_fabs), ...], "ng"} (FABs with their ghost cells)"""
    with open(os.path.join(path, "Header")) as fh:
        lines = fh.read().split("\n")
    finest = int(lines[1])
    out = {"isteps": [int(v) for v in lines[2].split()], "dts": [float(v) for v in lines[3].split()],
           "times": [float(v) for v in lines[4].split()], "levels": []}
    for lev in range(finest + 1):
        ff, ng, boxes = _read_vismf_fabs(os.path.join(path, f"Level_{lev}"), "f_00")
        gg, _, _ = _read_vismf_fabs(os.path.join(path, f"Level_{lev}"), "g_00")
        out["levels"].append((boxes, ff, gg))
        out["ng"] = ng
    return out


def write_amr_checkpoint(amr, directory: str = ".", prefix: str = "chk", digits: int = 5, max_level: int | None = None) -> str:
    """LBM::write_checkpoint_file for a hierarchy on the device: f and g of every box with their 3 ghost cells.  The
    unmodified reference restarts from it (amr.restart).  With distributed levels a collective call."""
    nl = amr.finest + 1
    nmax = (max_level if max_level is not None else amr.finest) + 1
    fab = lambda lev, ib, which: amr.get_box(lev, ib, which, ng=3) if amr.is_local(lev, ib) else None
    levels = [(amr.boxes[lev], [fab(lev, ib, 0) for ib in range(len(amr.boxes[lev]))],
               [fab(lev, ib, 1) for ib in range(len(amr.boxes[lev]))]) for lev in range(nl)]
    path = os.path.join(directory, chk_file_name(prefix, amr.isteps, digits))
    gather = None
    if amr.world > 1:  # collective call: every rank writes the FABs it holds, rank 0 the headers
        import torch.distributed as dist

        def gather(obj):
            out = [None] * amr.world
            dist.all_gather_object(out, obj)
            return out

    write_checkpoint_levels(path, levels, isteps=[amr.isteps * 2 ** l if l < nl else 0 for l in range(nmax)],
                            dts=[1.0 / 2 ** l for l in range(nmax)], times=[amr.time if l < nl else 0.0 for l in range(nmax)],
                            rank=amr.rank, owners=amr.owner if amr.world > 1 else None, gather=gather)
    return path


# ---------------------------------------------------------------------------------------------------------
# Multi-rank plotfiles: every rank writes the FABs of its own z-slab into Level_0/Cell_D_<rank>, rank 0 writes
# Header and Cell_H from the gathered box lists, offsets and extrema -- the layout VisMF uses when every rank
# owns a file (AMReX_VisMF.cpp).  `gather(obj) -> [obj of rank 0, obj of rank 1, ...]` is the only communication
# (torch.distributed.all_gather_object in marbles_b200.lbm; any callable in tests).
# ---------------------------------------------------------------------------------------------------------
def write_plotfile_slabs(path: str, names: list[str], data_local: np.ndarray, *, zlo: int, nz_total: int, rank: int,
                         gather, time: float, step: int, prob_lo, prob_hi, max_grid_size: int = 32) -> None:
    """data_local: [ncomp, nz_local, ny, nx] of this rank's slab, whose first plane is global plane zlo."""
    data_local = np.asarray(data_local, dtype=np.float64)
    ncomp, nzl, ny, nx = data_local.shape
    lev_dir = os.path.join(path, "Level_0")
    os.makedirs(lev_dir, exist_ok=True)
    dname = f"Cell_D_{rank:05d}"
    boxes, offsets, mins, maxs = [], [], [], []
    with open(os.path.join(lev_dir, dname), "wb") as fh:
        for lo, hi in chop_boxes((nx, ny, nzl), max_grid_size):
            glo, ghi = (lo[0], lo[1], lo[2] + zlo), (hi[0], hi[1], hi[2] + zlo)
            offsets.append(fh.tell())
            sub = np.ascontiguousarray(data_local[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1])
            fh.write(f"{FAB_HEADER}{_box(glo, ghi)} {ncomp}\n".encode())
            fh.write(sub.astype("<f8").tobytes())
            flat = sub.reshape(ncomp, -1)
            boxes.append((glo, ghi))
            mins.append(flat.min(axis=1).tolist())
            maxs.append(flat.max(axis=1).tolist())
    parts = gather({"file": dname, "boxes": boxes, "offsets": offsets, "mins": mins, "maxs": maxs})
    if rank != 0:
        return
    all_boxes = [b for p in parts for b in p["boxes"]]
    n = (nx, ny, nz_total)
    dx = [(float(prob_hi[d]) - float(prob_lo[d])) / n[d] for d in range(3)]
    with open(os.path.join(lev_dir, "Cell_H"), "w") as fh:
        fh.write(f"1\n1\n{ncomp}\n0\n")
        fh.write(f"({len(all_boxes)} 0\n")
        for lo, hi in all_boxes:
            fh.write(_box(lo, hi) + "\n")
        fh.write(")\n")
        fh.write(f"{len(all_boxes)}\n")
        for p in parts:
            for off in p["offsets"]:
                fh.write(f"FabOnDisk: {p['file']} {off}\n")
        fh.write("\n")
        for key in ("mins", "maxs"):
            fh.write(f"{len(all_boxes)},{ncomp}\n")
            for p in parts:
                for row in p[key]:
                    fh.write("".join("%.17e," % v for v in row) + "\n")
            fh.write("\n")
    with open(os.path.join(path, "Header"), "w") as fh:
        fh.write("HyperCLaw-V1.1\n")
        fh.write(f"{ncomp}\n")
        for nm in names:
            fh.write(nm + "\n")
        fh.write("3\n")
        fh.write(_g17(time) + "\n")
        fh.write("0\n")
        fh.write(" ".join(_g17(v) for v in prob_lo) + " \n")
        fh.write(" ".join(_g17(v) for v in prob_hi) + " \n")
        fh.write("\n")
        fh.write(_box((0, 0, 0), (nx - 1, ny - 1, nz_total - 1)) + " \n")
        fh.write(f"{step} \n")
        fh.write(" ".join(_g17(v) for v in dx) + " \n")
        fh.write("0\n0\n")
        fh.write(f"0 {len(all_boxes)} {_g17(time)}\n")
        fh.write(f"{step}\n")
        for lo, hi in all_boxes:
            for d in range(3):
                fh.write(f"{_g17(prob_lo[d] + lo[d] * dx[d])} {_g17(prob_lo[d] + (hi[d] + 1) * dx[d])}\n")
        fh.write("Level_0/Cell\n")


def write_checkpoint_slabs(path: str, f_local: np.ndarray, g_local: np.ndarray, *, zlo: int, nz_total: int, rank: int,
                           gather, step: int, dt: float, time: float, periodic, max_grid_size: int = 32, ng: int = 3):
    """Checkpoint of a z-slab run: every rank writes the FABs of its slab (f_00_D_<rank>, g_00_D_<rank>), rank 0 the
    Header and the two VisMF headers.  Ghost cells hold periodic images in x and y and the slab's own nearest
    planes in z (the reference refills every ghost cell after a restart)."""
    nzl, ny, nx = f_local.shape[1:]
    boxes = [((lo[0], lo[1], lo[2] + zlo), (hi[0], hi[1], hi[2] + zlo)) for lo, hi in chop_boxes((nx, ny, nzl), max_grid_size)]

    def grow(a):
        for axis, d in ((3, 0), (2, 1), (1, 2)):
            pad = [(0, 0)] * 4
            pad[axis] = (ng, ng)
            a = np.pad(a, pad, mode="wrap" if (periodic[d] and d != 2) else "edge")
        return a

    lev_dir = os.path.join(path, "Level_0")
    os.makedirs(lev_dir, exist_ok=True)
    parts = {}
    for name, a in (("f_00", f_local), ("g_00", g_local)):
        mine = _write_vismf_data(lev_dir, f"{name}_D_{rank:05d}", grow(np.asarray(a, dtype=np.float64)), ng, boxes, zoff=zlo)
        parts[name] = gather(mine)
    if rank != 0:
        return
    all_boxes = [b for p in parts["f_00"] for b in p["boxes"]]
    with open(os.path.join(path, "Header"), "w") as fh:
        fh.write("Checkpoint file for LBM\n0\n")
        fh.write(f"{step} \n{_g17(dt)} \n{_g17(time)} \n")
        fh.write(f"({len(all_boxes)} 0\n")
        for lo, hi in all_boxes:
            fh.write(_box(lo, hi) + "\n")
        fh.write(")\n")
    for name in ("f_00", "g_00"):
        _write_vismf_header(lev_dir, name, f_local.shape[0], ng, parts[name])
    _ = nz_total


def read_checkpoint_slab(path: str, zlo: int, zhi: int) -> dict:
    """the planes [zlo, zhi] of a single-level checkpoint (whoever wrote it, with whatever boxes)"""
    with open(os.path.join(path, "Header")) as fh:
        lines = fh.read().split("\n")
    if int(lines[1]) != 0:
        raise ValueError("only single-level checkpoints are supported")
    out = {"step": int(lines[2].split()[0]), "dt": float(lines[3].split()[0]), "time": float(lines[4].split()[0])}
    for name in ("f", "g"):
        data, ng, boxes = read_vismf(os.path.join(path, "Level_0"), f"{name}_00", zrange=(zlo, zhi))
        z0 = min(b[0][2] for b in boxes)
        a = data[:, ng:data.shape[1] - ng, ng:data.shape[2] - ng, ng:data.shape[3] - ng] if ng else data
        out[name] = np.ascontiguousarray(a[:, zlo - z0:zhi - z0 + 1])
    return out


def read_plotfile(path: str) -> dict:
    """A single-level plotfile as {variable name: [nz, ny, nx] array} plus "__names__", "__time__", "__step__";
    any number of data files (one per rank) and boxes."""
    import re
    with open(os.path.join(path, "Header")) as fh:
        lines = fh.read().split("\n")
    ncomp = int(lines[1])
    names = lines[2:2 + ncomp]
    pos = 2 + ncomp
    time = float(lines[pos + 1])
    if int(lines[pos + 2]) != 0:
        raise ValueError("only single-level plotfiles are supported")
    m = [int(v) for v in re.findall(r"-?\d+", lines[pos + 6])]
    lo, hi = m[0:3], m[3:6]
    step = int(lines[pos + 7].split()[0])
    data, ng, _ = read_vismf(os.path.join(path, "Level_0"), "Cell")
    assert ng == 0 and list(data.shape[1:]) == [hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1]
    out = {name: data[c] for c, name in enumerate(names)}
    out.update({"__names__": names, "__time__": time, "__step__": step})
    return out
