"""Host-side mirror of the reference's multi-level time loop over the multi-box C ABI.

The reference keeps the AMR hierarchy in AMReX (`lbm::LBM : amrex::AmrCore`); what the accelerated path
needs from it is the box list of every level.  `AmrLBM` takes those lists (as the reference-side shim
would hand over `boxArray(lev)` after `MakeNewLevelFromScratch` / `RemakeLevel`) and drives the
reference-granular entry points in the reference's order:

    LBM::init_data            Source/LBM.cpp:155-194   -> mbl_level_define_boxes, mbl_box_set_is_fluid,
                                                          mbl_initialize per level, mbl_average_down(ng 0)
    LBM::evolve (one step)    Source/LBM.cpp:416-422   -> fillpatch(0); time_step(0); post_time_step
    LBM::time_step            Source/LBM.cpp:452-521   -> fillpatch(lev+1); 2 x (physbc(lev+1); time_step(lev+1)); advance
    LBM::advance              Source/LBM.cpp:523-544   -> stream; average_down_to(lev, 1 ghost ring); collide
    LBM::MakeNewLevelFromCoarse / ClearLevel  Source/LBM.cpp:1088-1144, 1367-1380 -> `make_level_from_coarse`, `clear_level`
    LBM::RemakeLevel          Source/LBM.cpp:1302-1364 -> `regrid_level`: mbl_level_regrid with the box list AmrCore::regrid
                                                          produced (old level + coarse interpolation into new FABs),
                                                          new is_fluid, mbl_fill_f_inside_eb

No CPU fallback: every operator is a kernel of marbles_b200/csrc/patch.cu.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import LevelGeom, MarblesError, Params, check
from .geometry import is_fluid_from_deck
from .inputs import LbmInputs, lbm_inputs, parse_deck
from .lbm import DERIVED_NAMES, F_NGHOST, MACRO_NAMES, NDERIVED, NMACRO, NQ, _dptr

REF_RATIO = 2


class AmrLBM:
    """Multi-level lattice-Boltzmann state (all boxes of every level on this rank's B200)."""

    def __init__(self, deck, level_boxes, is_fluid=None, *, overrides=None, inputs: LbmInputs | None = None,
                 device: int = 0, cuda_stream: int | None = None, rank: int = 0, world: int = 1, owners=None,
                 exchange=None):
        """level_boxes[lev] = [(lo, hi), ...] valid boxes in the index space of level lev (AMReX BoxArray);
        is_fluid[lev] = dense int array over the level domain (component 0 of m_is_fluid on valid cells; ghost
        cells take the value of the cell they lie on, periodic images included, and 1 beyond non-periodic faces)
        or None -> the analytic body of the deck evaluated at each level's resolution.

        Distributed levels (one process per GPU, AMReX's DistributionMapping): every rank passes the whole box lists,
        `owners(lev, boxes) -> [rank per box]` (default: contiguous runs of the list, `default_owners`) says where each
        box lives, `exchange(rank, parts, stream)` moves the device messages (marbles_b200/amr_comm.py)."""
        self.rank, self.world = rank, world
        self._owners_of = owners or (lambda lev, boxes: default_owners(boxes, world))
        self._exchange = exchange
        self.owner: list[list[int]] = []
        self._isfl: dict = {}  # (lev, ib) -> is_fluid of the grown box as handed to the library (plotfiles)
        if inputs is None:
            d = deck if isinstance(deck, dict) else parse_deck(deck, overrides)
            inputs = lbm_inputs(d)
        self.inp = inputs
        self.lib = _lib.load()
        self.ctx = C.c_void_p()
        p = Params()
        p.nu, p.alpha, p.R, p.gamma, p.mesh_speed = inputs.nu, inputs.alpha, inputs.R, inputs.gamma, inputs.mesh_speed
        for d in range(3):
            p.bc_type[d], p.bc_type[d + 3], p.periodic[d] = inputs.bc_lo[d], inputs.bc_hi[d], inputs.periodic[d]
        p.vbc_kind, p.vbc_dir = inputs.vbc_kind, inputs.vbc_dir
        p.vbc_normal_dir, p.vbc_tangential_dir = inputs.vbc_normal_dir, inputs.vbc_tangential_dir
        p.vbc_u, p.vbc_rho, p.vbc_T = inputs.vbc_u, inputs.vbc_rho, inputs.vbc_T
        p.vbc_gamma, p.vbc_R = inputs.vbc_gamma, inputs.vbc_R
        self.params = p
        check(self.lib.mbl_create(C.byref(p), device, C.byref(self.ctx)))
        if cuda_stream is not None:
            check(self.lib.mbl_set_stream(self.ctx, C.c_void_p(cuda_stream)))
        if world > 1:
            if exchange is None:
                raise ValueError("distributed levels need an exchange (marbles_b200.amr_comm)")
            self._exchange_cb = _lib.EXCHANGE_FN(self._exchange_trampoline)  # kept alive with the object
            check(self.lib.mbl_set_exchange(self.ctx, rank, world, self._exchange_cb, None))
        self.boxes: list[list[tuple]] = []
        self.n: list[list[int]] = []
        self.dt: list[float] = []
        self.time = 0.0
        self.isteps = 0
        self._is_fluid_dense = list(is_fluid) if is_fluid is not None else None
        for lev, bxs in enumerate(level_boxes):
            dense = None if is_fluid is None else is_fluid[lev]
            self.define_level(lev, bxs, dense)

    def _exchange_trampoline(self, user, npeers, peers, send, nsend, recv, nrecv, stream):
        """mbl_exchange_fn: hands the library's per-peer device messages to the Python exchange object"""
        try:
            parts = [(int(peers[i]), int(send[i] or 0), int(nsend[i]), int(recv[i] or 0), int(nrecv[i])) for i in range(npeers)]
            self._exchange(self.rank, parts, int(stream or 0))
            return 0
        except Exception:  # an exception must not unwind through the C frames
            import traceback
            traceback.print_exc()
            return 1

    def _owner_array(self, lev: int, boxes):
        own = [int(o) for o in self._owners_of(lev, boxes)]
        assert len(own) == len(boxes) and all(0 <= o < self.world for o in own)
        return own, (C.c_int * len(own))(*own)

    def is_local(self, lev: int, ib: int) -> bool:
        return self.owner[lev][ib] == self.rank

    # ------------------------------------------------------------------ setup
    def close(self):
        if self.ctx:
            self.lib.mbl_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def finest(self) -> int:
        return len(self.boxes) - 1

    def level_geom(self, lev: int) -> LevelGeom:
        r = REF_RATIO ** lev
        g = LevelGeom()
        for d in range(3):
            g.dom_lo[d], g.dom_hi[d] = 0, self.inp.n_cell[d] * r - 1
            g.lo[d], g.hi[d] = g.dom_lo[d], g.dom_hi[d]
            g.dx[d] = self.inp.dx[d] / r
            g.inv_dx[d] = 1.0 / g.dx[d]
            g.prob_lo[d], g.prob_hi[d] = self.inp.prob_lo[d], self.inp.prob_hi[d]
        g.dt = 1.0 / r  # m_dts[lev] = m_dts[lev-1] / MaxRefRatio (Source/LBM.cpp:1073-1076), est_time_step == 1
        return g

    def define_level(self, lev: int, boxes, is_fluid_dense=None):
        """MakeNewLevelFromScratch's allocations + initialize_is_fluid (Source/LBM.cpp:1148-1262) for one level"""
        boxes = [(tuple(int(v) for v in lo), tuple(int(v) for v in hi)) for lo, hi in boxes]
        g = self.level_geom(lev)
        nb = len(boxes)
        lo = (C.c_int * (3 * nb))(*[v for b in boxes for v in b[0]])
        hi = (C.c_int * (3 * nb))(*[v for b in boxes for v in b[1]])
        own, cown = self._owner_array(lev, boxes)
        check(self.lib.mbl_level_define_boxes_on(self.ctx, lev, C.byref(g), nb, lo, hi, cown))
        n = [self.inp.n_cell[d] * REF_RATIO ** lev for d in range(3)]
        if lev < len(self.boxes):
            self.boxes[lev], self.n[lev], self.dt[lev], self.owner[lev] = boxes, n, g.dt, own
        else:
            assert lev == len(self.boxes), "levels are defined in order"
            self.boxes.append(boxes)
            self.n.append(n)
            self.dt.append(g.dt)
            self.owner.append(own)
        self._set_level_is_fluid(lev, is_fluid_dense)

    def _set_level_is_fluid(self, lev: int, is_fluid_dense=None):
        """initialize_is_fluid (Source/LBM.cpp:1213-1262) for every box of a level"""
        boxes, n = self.boxes[lev], self.n[lev]
        ng = F_NGHOST
        dx = [self.inp.dx[d] / REF_RATIO ** lev for d in range(3)]
        # the level's valid-cell field, once: ghost cells on cells of the level domain (periodic images included) take
        # its value -- m_is_fluid.FillBoundary(periodicity), Source/LBM.cpp:1234
        dense = is_fluid_dense
        if dense is None:
            dense = is_fluid_from_deck(self.inp.deck, n, self.inp.prob_lo, dx, (0, 0, 0), n, 0)
        dense = np.asarray(dense)
        all_fluid = bool(dense.min() == 1)
        for ib, (blo, bhi) in enumerate(boxes):
            if not self.is_local(lev, ib):
                continue
            nl = [bhi[d] - blo[d] + 1 for d in range(3)]
            if is_fluid_dense is None and not all_fluid:
                # beyond non-periodic faces: the body evaluated there (EB2 covers the grown domain)
                a = is_fluid_from_deck(self.inp.deck, n, self.inp.prob_lo, dx, blo, nl, ng)
            else:
                a = np.ones(tuple(nl[d] + 2 * ng for d in (2, 1, 0)), dtype=np.int32)
            idx, ok = [], []
            for d in (2, 1, 0):
                x = np.arange(blo[d] - ng, bhi[d] + ng + 1)
                if self.inp.periodic[d]:
                    idx.append(x % n[d])
                    ok.append(np.ones(x.shape, bool))
                else:
                    idx.append(np.clip(x, 0, n[d] - 1))
                    ok.append((x >= 0) & (x < n[d]))
            m = ok[0][:, None, None] & ok[1][None, :, None] & ok[2][None, None, :]
            v = dense[idx[0][:, None, None], idx[1][None, :, None], idx[2][None, None, :]]
            a = np.ascontiguousarray(np.where(m, v, a).astype(np.int32))
            self._isfl[(lev, ib)] = a
            if a.min() == 1:
                continue  # freshly defined boxes (mbl_level_define_boxes, _regrid, _make_from_coarse) start all fluid
            check(self.lib.mbl_box_set_is_fluid(self.ctx, lev, ib, a.ctypes.data_as(C.POINTER(C.c_int32)), ng))

    def init_data(self):
        """initialize_f on every level (IC on the grown boxes) and average_down(0 ghost cells), Source/LBM.cpp:163-167"""
        v = (C.c_double * 16)(*self.inp.ic_params)
        for lev in range(self.finest + 1):
            check(self.lib.mbl_initialize(self.ctx, lev, self.inp.ic_kind, v, 16))
        for lev in range(self.finest - 1, -1, -1):
            self.average_down_to(lev, 0)
        self.time, self.isteps = 0.0, 0

    # ---------------------------------------------------------------- operators
    def fillpatch(self, lev: int):
        check(self.lib.mbl_fillpatch(self.ctx, lev, self.time))

    def physbc(self, lev: int):
        check(self.lib.mbl_physbc(self.ctx, lev, self.time))

    def stream(self, lev: int):
        check(self.lib.mbl_stream(self.ctx, lev))

    def collide(self, lev: int, want_macrodata: bool = False):
        check(self.lib.mbl_collide(self.ctx, lev, int(want_macrodata)))

    def average_down_to(self, crse_lev: int, ng: int = 1):
        check(self.lib.mbl_average_down(self.ctx, crse_lev, ng))

    def advance(self, lev: int, want_macrodata: bool = False):
        """LBM::advance (Source/LBM.cpp:523-544): mbl_advance = stream; average_down_to(lev, 1) below a finer level;
        collide -- one fused pass on the finest level"""
        check(self.lib.mbl_advance(self.ctx, lev, int(want_macrodata)))

    def advance_unfused(self, lev: int, want_macrodata: bool = False):
        """the same through the three reference-granular entry points"""
        self.stream(lev)
        if lev < self.finest:
            self.average_down_to(lev, 1)
        self.collide(lev, want_macrodata)

    def time_step(self, lev: int, want_macrodata: bool = False):
        """LBM::time_step without the regrid check (Source/LBM.cpp:452-521): finer levels first, two substeps"""
        if lev < self.finest:
            self.fillpatch(lev + 1)
            for _ in range(REF_RATIO):
                self.physbc(lev + 1)
                self.time_step(lev + 1, want_macrodata)
        self.advance(lev, want_macrodata)

    def step(self, nsteps: int = 1, want_macrodata: bool = False):
        """nsteps coarse steps of LBM::evolve (Source/LBM.cpp:416-422); macrodata of the last collide of every
        level is stored for the last step when asked for"""
        for s in range(nsteps):
            self.fillpatch(0)
            self.time_step(0, want_macrodata and s == nsteps - 1)
            self.time += 1.0
            self.isteps += 1

    def compute_derived(self):
        """post_time_step: compute_derived on every level (Source/LBM.cpp:546-555)"""
        for lev in range(self.finest + 1):
            check(self.lib.mbl_compute_derived(self.ctx, lev))

    def write_checkpoint_file(self, directory: str = ".", prefix: str = "chk", digits: int = 5) -> str:
        """LBM::write_checkpoint_file (Source/LBM.cpp:1692-1783): Header + f and g of every box with their ghost cells;
        the unmodified reference restarts from it (amr.restart)"""
        from .plotfile import write_amr_checkpoint
        max_level = int(self.inp.deck.get("amr.max_level", [self.finest])[0])
        return write_amr_checkpoint(self, directory, prefix, digits, max_level=max(max_level, self.finest))

    @classmethod
    def from_checkpoint(cls, deck, path: str, is_fluid=None, **kw):
        """LBM::read_checkpoint_file (Source/LBM.cpp:1785-1915): the box lists, step counters and times of the Header, f
        and g of every FAB with their ghost cells, then is_fluid, fill_f_inside_eb and FillBoundary as the reference
        populates "the other data".  Written by the reference or by write_checkpoint_file."""
        from .plotfile import read_checkpoint_levels
        c = read_checkpoint_levels(path)
        amr = cls(deck, [lv[0] for lv in c["levels"]], is_fluid, **kw)
        for lev, (boxes, ff, gg) in enumerate(c["levels"]):
            for ib in range(len(boxes)):
                if amr.is_local(lev, ib):
                    amr.set_box(lev, ib, 0, ff[ib], ng=c["ng"])
                    amr.set_box(lev, ib, 1, gg[ib], ng=c["ng"])
            check(amr.lib.mbl_fill_f_inside_eb(amr.ctx, lev))
        amr.isteps, amr.time = int(c["isteps"][0]), float(c["times"][0])
        return amr

    def compute_eb_forces(self) -> np.ndarray:
        """LBM::compute_eb_forces (Source/LBM.cpp:994-1044): the levels' sums added as the reference adds them (m_mask is
        empty in every run the reference completes, so no cell is left out); this rank's boxes only -- the ranks of a
        distributed hierarchy add theirs (ParallelDescriptor::ReduceRealSum)"""
        total = np.zeros(3)
        out = (C.c_double * 3)()
        for lev in range(self.finest + 1):
            check(self.lib.mbl_eb_forces(self.ctx, lev, out))
            total += np.array(out[:])
        return total

    def sync(self):
        check(self.lib.mbl_sync(self.ctx))

    def regrid_level(self, lev: int, boxes, is_fluid_dense=None):
        """LBM::RemakeLevel (Source/LBM.cpp:1302-1364) with the box list AmrCore::regrid produced for level lev >= 1:
        mbl_level_regrid fills the new boxes from the old level and, where that does not cover them, from level lev-1
        (FillPatchOps::fillpatch into new MultiFabs); then the new is_fluid and fill_f_inside_eb + FillBoundary."""
        boxes = [(tuple(int(v) for v in lo), tuple(int(v) for v in hi)) for lo, hi in boxes]
        nb = len(boxes)
        lo = (C.c_int * (3 * nb))(*[v for b in boxes for v in b[0]])
        hi = (C.c_int * (3 * nb))(*[v for b in boxes for v in b[1]])
        own, cown = self._owner_array(lev, boxes)
        check(self.lib.mbl_level_regrid_on(self.ctx, lev, nb, lo, hi, cown))
        self.boxes[lev], self.owner[lev] = boxes, own
        self._set_level_is_fluid(lev, is_fluid_dense)
        check(self.lib.mbl_fill_f_inside_eb(self.ctx, lev))

    def make_level_from_coarse(self, lev: int, boxes, is_fluid_dense=None):
        """LBM::MakeNewLevelFromCoarse (Source/LBM.cpp:1088-1144): level lev = finest + 1 appears in a regrid; every cell
        of its grown boxes is interpolated from level lev-1 (mbl_level_make_from_coarse), then is_fluid is set."""
        assert lev == self.finest + 1, "a new level appears above the finest one"
        boxes = [(tuple(int(v) for v in lo), tuple(int(v) for v in hi)) for lo, hi in boxes]
        g = self.level_geom(lev)
        nb = len(boxes)
        lo = (C.c_int * (3 * nb))(*[v for b in boxes for v in b[0]])
        hi = (C.c_int * (3 * nb))(*[v for b in boxes for v in b[1]])
        own, cown = self._owner_array(lev, boxes)
        check(self.lib.mbl_level_make_from_coarse_on(self.ctx, lev, C.byref(g), nb, lo, hi, cown))
        self.boxes.append(boxes)
        self.owner.append(own)
        self.n.append([self.inp.n_cell[d] * REF_RATIO ** lev for d in range(3)])
        self.dt.append(g.dt)
        self._set_level_is_fluid(lev, is_fluid_dense)

    def clear_level(self, lev: int):
        """LBM::ClearLevel (Source/LBM.cpp:1367-1380): the finest level vanishes in a regrid"""
        assert lev == self.finest and lev >= 1
        check(self.lib.mbl_level_clear(self.ctx, lev))
        self.boxes.pop()
        self.n.pop()
        self.dt.pop()
        self.owner.pop()

    # ------------------------------------------------------------------ access
    def box_shape(self, lev: int, ib: int, ncomp: int, ng: int):
        lo, hi = self.boxes[lev][ib]
        return (ncomp,) + tuple(hi[d] - lo[d] + 1 + 2 * ng for d in (2, 1, 0))

    def get_box(self, lev: int, ib: int, which: int, ng: int = 0) -> np.ndarray:
        a = np.zeros(self.box_shape(lev, ib, NQ, ng))
        check(self.lib.mbl_box_download(self.ctx, lev, ib, which, _dptr(a), ng))
        return a

    def get_box_macrodata(self, lev: int, ib: int, derived: bool = False, ng: int = 0) -> np.ndarray:
        a = np.zeros(self.box_shape(lev, ib, NDERIVED if derived else NMACRO, ng))
        check(self.lib.mbl_box_download_macrodata(self.ctx, lev, ib, _dptr(a), ng, int(derived)))
        return a

    def box_is_fluid(self, lev: int, ib: int) -> np.ndarray:
        """component 0 of m_is_fluid on the box grown by 3 ghost cells, as the library was given it"""
        return self._isfl[(lev, ib)]

    def set_box(self, lev: int, ib: int, which: int, a: np.ndarray, ng: int = 0):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self.box_shape(lev, ib, NQ, ng)
        check(self.lib.mbl_box_upload(self.ctx, lev, ib, which, _dptr(a), ng))

    def dense(self, lev: int, which: str) -> np.ndarray:
        """valid cells of every box of a level THIS RANK holds, gathered over the level domain (NaN where the level has
        no box here; `merge_dense` puts the ranks' arrays together); which = 'f' | 'g' | 'macro' | 'derived'"""
        ncomp = {"f": NQ, "g": NQ, "macro": NMACRO, "derived": NDERIVED}[which]
        n = self.n[lev]
        out = np.full((ncomp, n[2], n[1], n[0]), np.nan)
        for ib, (lo, hi) in enumerate(self.boxes[lev]):
            if not self.is_local(lev, ib):
                continue
            a = np.zeros(self.box_shape(lev, ib, ncomp, 0))
            if which in ("f", "g"):
                check(self.lib.mbl_box_download(self.ctx, lev, ib, 0 if which == "f" else 1, _dptr(a), 0))
            else:
                check(self.lib.mbl_box_download_macrodata(self.ctx, lev, ib, _dptr(a), 0, int(which == "derived")))
            out[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] = a
        return out

    def fields(self, lev: int, macro: bool = True) -> dict:
        """dense valid-cell fields of one level under the reference's plotfile names"""
        out = {}
        f, g = self.dense(lev, "f"), self.dense(lev, "g")
        for q in range(NQ):
            out[f"f_{q:02d}"] = f[q]
            out[f"g_{q:02d}"] = g[q]
        if macro:
            m = self.dense(lev, "macro")
            for n, name in enumerate(MACRO_NAMES):
                out[name] = m[n]
            d = self.dense(lev, "derived")
            for n, name in enumerate(DERIVED_NAMES):
                out[name] = d[n]
        return out

    @property
    def launches(self) -> int:
        return int(self.lib.mbl_launch_count(self.ctx))

    def ncells(self, lev: int | None = None) -> int:
        levs = range(self.finest + 1) if lev is None else [lev]
        return sum(int(np.prod([hi[d] - lo[d] + 1 for d in range(3)])) for l in levs for lo, hi in self.boxes[l])


def default_owners(boxes, world: int):
    """box -> rank as AMReX's space-filling-curve DistributionMapping does it (AMReX_DistributionMapping.cpp, SFC strategy):
    the boxes sorted along a Morton curve through their positions, cut into `world` runs of about the same number of
    cells -- every rank gets a compact region, so the cells that cross ranks are few"""
    if world <= 1:
        return [0] * len(boxes)
    ext = [max(hi[d] - lo[d] + 1 for lo, hi in boxes) for d in range(3)]

    def morton(lo):
        c = [lo[d] // ext[d] for d in range(3)]
        key = 0
        for bit in range(21):
            for d in range(3):
                key |= ((c[d] >> bit) & 1) << (3 * bit + d)
        return key

    order = sorted(range(len(boxes)), key=lambda i: (morton(boxes[i][0]), i))
    cells = [int(np.prod([boxes[i][1][d] - boxes[i][0][d] + 1 for d in range(3)])) for i in order]
    total, acc, own = float(sum(cells)), 0.0, [0] * len(boxes)
    for i, c in zip(order, cells):
        own[i] = min(world - 1, int((acc + 0.5 * c) * world / total))
        acc += c
    return own


def merge_dense(parts):
    """dense arrays of the ranks (NaN where a rank holds no box) -> one array"""
    out = np.array(parts[0], copy=True)
    for p in parts[1:]:
        m = np.isnan(out)
        out[m] = np.asarray(p)[m]
    return out
