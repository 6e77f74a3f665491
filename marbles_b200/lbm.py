"""Host-side mirror of the reference's `lbm::LBM` surface for the accelerated path.

The reference drives the lattice update through member functions of
`lbm::LBM : amrex::AmrCore` (Source/LBM.H:28-119).  This class keeps their names,
argument meaning and call order for the single-level hot path and forwards each of
them to the C ABI (include/marbles_b200.h):

    LBM::init_data            Source/LBM.cpp:155-194   -> mbl_level_define, mbl_set_is_fluid, mbl_initialize
    LBM::evolve               Source/LBM.cpp:398-448   -> fillpatch, time_step, post_time_step per step
    FillPatchOps::fillpatch   Source/FillPatchOps.H:75-132 -> mbl_fillpatch (+ z-halo exchange between ranks)
    LBM::advance              Source/LBM.cpp:523-544   -> stream(f), stream(g), collide
    LBM::stream               Source/LBM.cpp:558-604   -> mbl_stream
    LBM::collide              Source/LBM.cpp:607-618   -> mbl_collide
    LBM::f_to_macrodata       Source/LBM.cpp:810-906   -> mbl_f_to_macrodata
    LBM::compute_derived      Source/LBM.cpp:909-955   -> mbl_compute_derived
    LBM::compute_eb_forces    Source/LBM.cpp:994-1044  -> mbl_eb_forces (+ all-reduce)

`step()` is the fused fast path (mbl_step): same result as fillpatch + advance.
Errors raise (the reference calls amrex::Abort).  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import LevelGeom, Layout, MarblesError, Params, check
from .geometry import body_from_deck, is_fluid_from_deck
from .inputs import LbmInputs, lbm_inputs, parse_deck

NQ, NMACRO, NDERIVED = 27, 19, 7
MACRO_NAMES = ["rho", "vel_x", "vel_y", "vel_z", "vel_mag", "two_rho_e", "QCorrX", "QCorrY", "QCorrZ",
               "pxx", "pyy", "pzz", "pxy", "pxz", "pyz", "qx", "qy", "qz", "temperature"]
DERIVED_NAMES = ["vort_x", "vort_y", "vort_z", "vort_mag", "dQCorrX", "dQCorrY", "dQCorrZ"]
F_NGHOST = 3  # m_f_nghost, Source/LBM.H:230


def _dptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def slab_bounds(nz: int, rank: int, world: int) -> tuple[int, int]:
    """z-range [lo, hi] (inclusive) of `rank`'s slab: contiguous, sizes differ by at most one plane."""
    base, rem = divmod(nz, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0) - 1


def wrap_periodic(a: np.ndarray, ng: int, n, periodic) -> np.ndarray:
    """m_is_fluid.FillBoundary(periodicity) on one box that spans the domain (n cells per direction): ghost cells in
    periodic directions mirror the valid cells; a is [nz + 2 ng, ny + 2 ng, nx + 2 ng]."""
    for axis, d in ((2, 0), (1, 1), (0, 2)):
        if periodic[d]:
            idx = (np.arange(-ng, n[d] + ng) % n[d]) + ng
            a = np.take(a, idx, axis=axis)
    return np.ascontiguousarray(a)


def slab_is_fluid_from_deck(inp: LbmInputs, zlo: int, zhi: int, ng: int = F_NGHOST) -> np.ndarray:
    """is_fluid (component 0) of the z-slab [zlo, zhi] grown by ng for the analytic body of the deck.  Ghost planes
    that belong to a neighbouring rank, or to the periodic image at the domain ends, hold THOSE cells
    (m_is_fluid.FillBoundary(periodicity), Source/LBM.cpp:1234); planes beyond a non-periodic domain end keep the
    geometry evaluated beyond the domain (Source/LBM.cpp:1222-1232).  The body is analytic, so the whole domain is
    evaluated once, wrapped, and the slab cut out."""
    glob = is_fluid_from_deck(inp.deck, inp.n_cell, inp.prob_lo, inp.dx, (0, 0, 0), inp.n_cell, ng)
    glob = wrap_periodic(glob, ng, inp.n_cell, inp.periodic)
    return np.ascontiguousarray(glob[zlo:zhi + 1 + 2 * ng])


class LBM:
    """Single-level lattice-Boltzmann state of one rank (one z-slab of the domain) on one B200."""

    def __init__(self, deck=None, overrides=None, *, inputs: LbmInputs | None = None, device: int = 0,
                 rank: int = 0, world: int = 1, comm=None, is_fluid: np.ndarray | None = None,
                 cuda_stream: int | None = None, variant: int | None = None):
        if inputs is None:
            d = deck if isinstance(deck, dict) else parse_deck(deck, overrides)
            if isinstance(deck, dict) and overrides:
                d = dict(d)
                d.update(parse_deck(None, overrides))
            inputs = lbm_inputs(d)
        self.inp = inputs
        self.rank, self.world, self.comm = rank, world, comm
        self.cuda_stream = cuda_stream
        self.overlap = os.environ.get("MBL_OVERLAP", "1") != "0"
        self._ghosts_fresh = False  # the ghost planes of the current buffers hold the neighbours' planes
        self.lib = _lib.load()
        self.ctx = C.c_void_p()
        self.lev = 0
        self.time = 0.0
        self.isteps = 0
        self.dt = 1.0  # est_time_step == 1 (Source/LBM.cpp:1080-1084)
        n = inputs.n_cell
        zlo, zhi = slab_bounds(n[2], rank, world)
        self.lo = (0, 0, zlo)
        self.hi = (n[0] - 1, n[1] - 1, zhi)
        self.n_local = (n[0], n[1], zhi - zlo + 1)

        p = Params()
        p.nu, p.alpha, p.R, p.gamma, p.mesh_speed = inputs.nu, inputs.alpha, inputs.R, inputs.gamma, inputs.mesh_speed
        for d in range(3):
            p.bc_type[d] = inputs.bc_lo[d]
            p.bc_type[d + 3] = inputs.bc_hi[d]
            p.periodic[d] = inputs.periodic[d]
        p.vbc_kind, p.vbc_dir = inputs.vbc_kind, inputs.vbc_dir
        p.vbc_normal_dir, p.vbc_tangential_dir = inputs.vbc_normal_dir, inputs.vbc_tangential_dir
        p.vbc_u, p.vbc_rho, p.vbc_T = inputs.vbc_u, inputs.vbc_rho, inputs.vbc_T
        p.vbc_gamma, p.vbc_R = inputs.vbc_gamma, inputs.vbc_R
        self.params = p
        check(self.lib.mbl_create(C.byref(p), device, C.byref(self.ctx)))
        if cuda_stream is not None:
            check(self.lib.mbl_set_stream(self.ctx, C.c_void_p(cuda_stream)))
        if variant is not None:
            check(self.lib.mbl_set_variant(self.ctx, variant))
        self.variant = int(self.lib.mbl_get_variant(self.ctx))  # library default (MBL_VARIANT) unless given

        g = LevelGeom()
        dx = inputs.dx
        for d in range(3):
            g.dom_lo[d], g.dom_hi[d] = 0, n[d] - 1
            g.lo[d], g.hi[d] = self.lo[d], self.hi[d]
            g.inv_dx[d] = 1.0 / dx[d]
            g.prob_lo[d], g.prob_hi[d], g.dx[d] = inputs.prob_lo[d], inputs.prob_hi[d], dx[d]
        g.dt = self.dt
        self.geom = g
        self.layout = Layout()
        check(self.lib.mbl_level_layout(C.byref(g), C.byref(self.layout)))
        check(self.lib.mbl_level_define(self.ctx, self.lev, C.byref(g), None))
        self._is_fluid_cache = None
        self._all_fluid = True
        self._halo = None
        self.set_is_fluid(is_fluid)
        # lean z-halo (half the bytes per exchange): legal where no bounce-back and no boundary ghost value sits next
        # to a slab cut -- all-periodic decks without solid cells
        self.halo_lean = (world > 1 and all(self.inp.periodic) and self._all_fluid
                          and os.environ.get("MBL_HALO_LEAN", "1") != "0")
        check(self.lib.mbl_set_halo_lean(self.ctx, int(self.halo_lean)))

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

    def set_is_fluid(self, is_fluid: np.ndarray | None):
        """LBM::initialize_is_fluid (Source/LBM.cpp:1213-1262).  `is_fluid`: component 0 on the local
        valid box [nz,ny,nx] or on the box grown by F_NGHOST; None -> eb2.geom_type of the deck."""
        ng = F_NGHOST
        nx, ny, nz = self.n_local
        full = (nz + 2 * ng, ny + 2 * ng, nx + 2 * ng)
        if is_fluid is None:
            body = body_from_deck(self.inp.deck)
            if body is not None and os.environ.get("MBL_HOST_GEOMETRY", "0") != "1":
                # the analytic bodies of the shipped decks are evaluated on the device (mbl_set_body): no host field,
                # no upload; the host copy (plotfiles) is made on demand
                kind, par = body
                self._is_fluid_cache, self._all_fluid = None, kind == 0
                v = (C.c_double * max(len(par), 1))(*par)
                check(self.lib.mbl_set_body(self.ctx, self.lev, kind, v, len(par)))
                return
            a = slab_is_fluid_from_deck(self.inp, self.lo[2], self.hi[2], ng)
            if a.min() == 1:
                self._is_fluid = a
                check(self.lib.mbl_set_all_fluid(self.ctx, self.lev))
                return
        else:
            is_fluid = np.asarray(is_fluid, dtype=np.int32)
            if is_fluid.shape == full:
                a = np.ascontiguousarray(is_fluid)
            elif is_fluid.shape == (nz, ny, nx):
                if self.world > 1:
                    raise MarblesError(
                        "on a z-slab (world > 1) is_fluid must cover the slab grown by 3 cells: the ghost planes are "
                        "the neighbouring ranks' cells (m_is_fluid.FillBoundary), which a valid-box array cannot supply")
                a = np.ones(full, dtype=np.int32)
                a[ng:-ng, ng:-ng, ng:-ng] = is_fluid
                a = self._wrap_periodic(a, ng, z_local=True)
            else:
                raise MarblesError(f"is_fluid has shape {is_fluid.shape}, expected {(nz, ny, nx)} or {full}")
        self._is_fluid = a
        check(self.lib.mbl_set_is_fluid(self.ctx, self.lev, a.ctypes.data_as(C.POINTER(C.c_int32)), ng))

    @property
    def _is_fluid(self) -> np.ndarray:
        """host copy of is_fluid (component 0, slab grown by 3); evaluated on demand when the device built the flags"""
        if self._is_fluid_cache is None:
            self._is_fluid_cache = slab_is_fluid_from_deck(self.inp, self.lo[2], self.hi[2], F_NGHOST)
        return self._is_fluid_cache

    @_is_fluid.setter
    def _is_fluid(self, a: np.ndarray):
        self._is_fluid_cache = a
        self._all_fluid = bool(int(a.min()) == 1)

    def _wrap_periodic(self, a: np.ndarray, ng: int, z_local: bool, n_local=None) -> np.ndarray:
        """m_is_fluid.FillBoundary(periodicity): ghost cells in periodic directions mirror the valid cells."""
        a = a.copy()
        nloc = self.n_local if n_local is None else n_local
        for axis, d in ((2, 0), (1, 1), (0, 2)):
            if not self.inp.periodic[d] or (d == 2 and not z_local):
                continue
            n = nloc[d]
            idx = (np.arange(-ng, n + ng) % n) + ng
            a = np.take(a, idx, axis=axis)
        return np.ascontiguousarray(a)

    def init_data(self):
        """initialize_f of MakeNewLevelFromScratch (Source/LBM.cpp:1186-1198) on the device."""
        v = (C.c_double * 16)(*self.inp.ic_params)
        check(self.lib.mbl_initialize(self.ctx, self.lev, self.inp.ic_kind, v, 16))
        self.time, self.isteps = 0.0, 0
        self._ghosts_fresh = False

    # ---------------------------------------------------------------- operators
    def exchange_halo(self):
        """ghost planes owned by neighbouring ranks (the FillBoundary part that crosses ranks)."""
        if self.world > 1:
            if self.comm is None:
                raise MarblesError("world > 1 needs a halo communicator (marbles_b200.parallel.HaloComm)")
            self.comm.exchange(self)

    def fillpatch(self, lev: int = 0, time: float | None = None):
        """m_fillpatch_op->fillpatch(lev, time, m_f[lev]) and the same for m_g (Source/LBM.cpp:416-418)."""
        self.exchange_halo()
        check(self.lib.mbl_fillpatch(self.ctx, lev, self.time if time is None else time))

    def physbc(self, lev: int = 0, time: float | None = None):
        check(self.lib.mbl_physbc(self.ctx, lev, self.time if time is None else time))

    def stream(self, lev: int = 0):
        """stream(lev, m_f); stream(lev, m_g) (Source/LBM.cpp:535-537)."""
        check(self.lib.mbl_stream(self.ctx, lev))
        self._ghosts_fresh = False

    def collide(self, lev: int = 0, want_macrodata: bool = True):
        if self.world > 1:
            self.exchange_halo()  # post-stream state of the neighbours' edge planes (q-correction stencil)
        check(self.lib.mbl_collide(self.ctx, lev, int(want_macrodata)))
        self._ghosts_fresh = False

    def advance(self, lev: int = 0):
        """LBM::advance (Source/LBM.cpp:523-544), un-fused."""
        self.stream(lev)
        self.collide(lev)

    def f_to_macrodata(self, lev: int = 0):
        check(self.lib.mbl_f_to_macrodata(self.ctx, lev))

    def compute_derived(self, lev: int = 0):
        """LBM::compute_derived.  On a slab the planes next to another rank difference across the rank boundary:
        the neighbours' adjacent macrodata planes are exchanged first (collective call)."""
        if self.world > 1 and self.comm is not None and hasattr(self.comm, "exchange_macro"):
            self.comm.exchange_macro(self)
            check(self.lib.mbl_compute_derived_slab(self.ctx, lev, int(self.comm.lower is not None),
                                                    int(self.comm.upper is not None)))
        else:
            check(self.lib.mbl_compute_derived(self.ctx, lev))

    def compute_eb_forces(self) -> np.ndarray:
        out = (C.c_double * 3)()
        if self.world > 1:
            self.exchange_halo()
        check(self.lib.mbl_eb_forces(self.ctx, self.lev, out))
        f = np.array(out[:])
        if self.world > 1:
            f = self.comm.allreduce_sum(f)
        return f

    def can_overlap(self) -> bool:
        """The overlapped slab step (mbl_step_split: boundary planes first, their exchange behind the interior planes)
        applies to the two-kernel and the tile-carry variants and to slabs of at least 8 planes; levels with walls,
        inlets or outlets fill their ghost cells at the start of part 0."""
        return (self.world > 1 and self.variant in (0, 5, 7, 8, 9, 10) and self.n_local[2] >= 8
                and self.comm is not None and hasattr(self.comm, "exchange_next") and self.overlap)

    def step(self, nsteps: int = 1, want_macrodata: bool = False):
        """Fused fast path: nsteps x (fillpatch f, g; stream f, g; collide)."""
        if self.world == 1:
            check(self.lib.mbl_step(self.ctx, self.lev, nsteps, self.time, int(want_macrodata)))
        elif self.can_overlap() and nsteps > (1 if want_macrodata else 0):
            self._step_overlapped(nsteps - (1 if want_macrodata else 0))
            if want_macrodata:  # the last step stores all 19 fields: the plain slab step
                self._ghosts_fresh = False
                check(self.lib.mbl_step_local(self.ctx, self.lev, self.time + (nsteps - 1) * self.dt, 1))
        else:
            for s in range(nsteps):
                self.exchange_halo()
                check(self.lib.mbl_step_local(self.ctx, self.lev, self.time + s * self.dt,
                                              int(want_macrodata and s == nsteps - 1)))
            self._ghosts_fresh = False
        self.time += nsteps * self.dt
        self.isteps += nsteps

    def _step_overlapped(self, nsteps: int):
        """Slab steps with the ghost-plane exchange hidden behind the interior planes: the two boundary planes
        at each z-end are collided first (part 0), travel to the neighbours on the communicator's stream while
        the interior planes are collided (part 1), and are awaited only by the next step's part 0."""
        if not self._ghosts_fresh:
            self.exchange_halo()  # blocking exchange of the current buffers' ghost planes
            self._ghosts_fresh = True
        for _ in range(nsteps):
            self.comm.wait_exchange(self)                        # ghost planes of the current buffers are in
            check(self.lib.mbl_step_split(self.ctx, self.lev, 0))
            self.comm.exchange_next(self)                        # async: boundary planes of the written buffers
            check(self.lib.mbl_step_split(self.ctx, self.lev, 1))
        self.comm.wait_exchange(self)

    def evolve(self, max_step: int | None = None, fused: bool = True, want_macrodata: bool = False):
        """LBM::evolve (Source/LBM.cpp:398-448) without I/O."""
        nsteps = self.inp.max_step if max_step is None else max_step
        if fused:
            self.step(nsteps, want_macrodata)
            return
        for _ in range(nsteps):
            self.fillpatch(0)
            self.advance(0)
            self.time += self.dt
            self.isteps += 1

    def sync(self):
        check(self.lib.mbl_sync(self.ctx))

    def step_host(self, f_fab: np.ndarray, g_fab: np.ndarray, nsteps: int = 1, ng: int = F_NGHOST):
        """The reference-facing call with HOST buffers (FAB layout): upload, nsteps, download in place.  On a
        z-slab of an all-periodic multi-rank run the boundary planes go up first, are exchanged with the
        neighbours, and the rest of the slab is pipelined (mbl_step_host_begin / _finish)."""
        if self.world > 1:
            if nsteps != 1 or not all(self.inp.periodic) or self.n_local[2] < 8:
                self.set_state(f_fab, g_fab, ng)
                self.step(nsteps)
                check(self.lib.mbl_download(self.ctx, self.lev, 0, _dptr(f_fab), ng))
                check(self.lib.mbl_download(self.ctx, self.lev, 1, _dptr(g_fab), ng))
                return
            check(self.lib.mbl_step_host_begin(self.ctx, self.lev, _dptr(f_fab), _dptr(g_fab), ng))
            self.exchange_halo()
            check(self.lib.mbl_step_host_finish(self.ctx, self.lev, _dptr(f_fab), _dptr(g_fab), ng))
            self._ghosts_fresh = False
            self.time += self.dt
            self.isteps += 1
            return
        check(self.lib.mbl_step_host(self.ctx, self.lev, nsteps, self.time, _dptr(f_fab), _dptr(g_fab), ng))
        self._ghosts_fresh = False
        self.time += nsteps * self.dt
        self.isteps += nsteps

    # ------------------------------------------------------------------ access
    def fab_shape(self, ncomp: int, ng: int):
        nx, ny, nz = self.n_local
        return (ncomp, nz + 2 * ng, ny + 2 * ng, nx + 2 * ng)

    def set_state(self, f: np.ndarray, g: np.ndarray, ng: int = F_NGHOST):
        f = np.ascontiguousarray(f, dtype=np.float64)
        g = np.ascontiguousarray(g, dtype=np.float64)
        assert f.shape == self.fab_shape(NQ, ng) and g.shape == f.shape, (f.shape, self.fab_shape(NQ, ng))
        check(self.lib.mbl_upload(self.ctx, self.lev, 0, _dptr(f), ng))
        check(self.lib.mbl_upload(self.ctx, self.lev, 1, _dptr(g), ng))
        self._ghosts_fresh = False

    def get_f(self, ng: int = 0) -> np.ndarray:
        a = np.zeros(self.fab_shape(NQ, ng))
        check(self.lib.mbl_download(self.ctx, self.lev, 0, _dptr(a), ng))
        return a

    def get_g(self, ng: int = 0) -> np.ndarray:
        a = np.zeros(self.fab_shape(NQ, ng))
        check(self.lib.mbl_download(self.ctx, self.lev, 1, _dptr(a), ng))
        return a

    def get_macrodata(self, ng: int = 0) -> np.ndarray:
        a = np.zeros(self.fab_shape(NMACRO, ng))
        check(self.lib.mbl_download_macrodata(self.ctx, self.lev, _dptr(a), ng))
        return a

    def get_derived(self) -> np.ndarray:
        a = np.zeros(self.fab_shape(NDERIVED, 0))
        check(self.lib.mbl_download_derived(self.ctx, self.lev, _dptr(a)))
        return a

    def fields(self) -> dict:
        """Valid-cell fields under the reference's plotfile names (Source/LBM.cpp:302-340)."""
        out = {}
        m = self.get_macrodata()
        for n, name in enumerate(MACRO_NAMES):
            out[name] = m[n]
        f, g = self.get_f(), self.get_g()
        for q in range(NQ):
            out[f"f_{q:02d}"] = f[q]
            out[f"g_{q:02d}"] = g[q]
        return out

    def write_plot_file(self, directory: str = ".", prefix: str = "plt") -> str:
        """LBM::write_plot_file (Source/LBM.cpp:1677-1690): the current state as an AMReX plotfile
        `<prefix><step:05d>`; the last step must have been taken with want_macrodata=True."""
        from .plotfile import write_lbm_plotfile
        return write_lbm_plotfile(self, directory, prefix)

    def write_checkpoint_file(self, directory: str = ".", prefix: str = "chk", digits: int = 5) -> str:
        """LBM::write_checkpoint_file (Source/LBM.cpp:1692-1783): `<prefix><step:05d>` with Header, f_00 and g_00
        VisMF files (3 ghost cells); the unmodified reference restarts from it (amr.restart)."""
        from . import plotfile as P
        path = os.path.join(directory, P.chk_file_name(prefix, self.isteps, digits))
        deck = self.inp.deck
        mgs = deck.get("amr.max_grid_size", 32)
        mgs = int(str(mgs[0] if isinstance(mgs, (list, tuple)) else mgs).split()[0])
        if self.world == 1:
            P.write_checkpoint(path, self.get_f(), self.get_g(), step=self.isteps, dt=self.dt, time=self.time,
                               periodic=self.inp.periodic, max_grid_size=mgs, ng=F_NGHOST)
        else:  # collective: every rank writes its slab's FABs, rank 0 the headers
            import torch.distributed as dist

            def gather(obj):
                out = [None] * self.world
                dist.all_gather_object(out, obj)
                return out

            P.write_checkpoint_slabs(path, self.get_f(), self.get_g(), zlo=self.lo[2], nz_total=self.inp.n_cell[2],
                                     rank=self.rank, gather=gather, step=self.isteps, dt=self.dt, time=self.time,
                                     periodic=self.inp.periodic, max_grid_size=mgs, ng=F_NGHOST)
        return path

    def read_checkpoint_file(self, path: str):
        """LBM::read_checkpoint_file (Source/LBM.cpp:1785-1915): load f, g, step and time of a single-level
        checkpoint written by the reference or by write_checkpoint_file."""
        from . import plotfile as P
        c = P.read_checkpoint(path) if self.world == 1 else P.read_checkpoint_slab(path, self.lo[2], self.hi[2])
        if c["f"].shape != self.fab_shape(NQ, 0):
            raise MarblesError(f"checkpoint holds {c['f'].shape[1:]} cells, this level {self.fab_shape(NQ, 0)[1:]}")
        self.set_state(c["f"], c["g"], ng=0)
        self.isteps, self.time, self.dt = c["step"], c["time"], c["dt"]

    @property
    def launches(self) -> int:
        return int(self.lib.mbl_launch_count(self.ctx))

    @property
    def ncells(self) -> int:
        return int(np.prod(self.n_local))
