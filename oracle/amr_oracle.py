"""CPU ORACLE, multi-level part -- TEST INFRASTRUCTURE ONLY.

Restates what the reference does AROUND the per-box lattice update when a run has several boxes per level and
several levels (BASELINE configs 4-5): the sub-cycled time step, the same-level ghost exchange, the
fine -> coarse average and the coarse -> fine ghost interpolation.  The per-box arithmetic is the plain-C oracle
(``marbles_oracle.c``: stream, macrodata, q-corrections, equilibria, relax, K6 pre-pass, BCFill); this module adds

* ``Level.fill_boundary``      FabArray::FillBoundary(periodicity): ghost cells that lie on valid cells of the
                               same level (other boxes, periodic images) take those values
                               (Submodules/AMReX/Src/Base/AMReX_FabArrayCommI.H:8-253)
* ``AmrOracle.average_down_to`` lbm::average_down_with_ghosts / masked_avgdown (Source/Utilities.cpp:5-28,
                               Source/Utilities.H:315-350) including one coarse ghost ring, the -1 sentinel mask
                               and AMReX's copy order where the rings of neighbouring fine boxes overlap
                               (FabArrayBase::CPC::define, AMReX_FabArrayBase.cpp:328-472; BoxArray::intersections,
                               AMReX_BoxArray.cpp:1219-1310; Periodicity::shiftIntVect, AMReX_Periodicity.cpp:8-33)
* ``AmrOracle.fillpatch``      FillPatchOps::fillpatch for lev > 0 (Source/FillPatchOps.H:75-132):
                               K6 pre-pass, FillPatchTwoLevels (AMReX_FillPatchUtil_I.H:450-621) with
                               cell_cons_interp = CellConservativeLinear(do_linear_limiting = false)
                               (AMReX_Interpolater.cpp:41, 833-982; slopes AMReX_MFInterp_3D_C.H:176-249,
                               AMReX_MFInterp_C.H:11-34), FillBoundary, BCFill
* ``AmrOracle.time_step``      LBM::time_step / advance (Source/LBM.cpp:452-544): finer levels first, two substeps,
                               stream -> average_down_to(lev, 1 ghost ring) -> collide

Parity is PINNED: tests/test_oracle_amr.py compares this module with golden vectors written by the unmodified
reference on 2- and 3-level decks (tests/golden/make_golden.py, cases ``amr*``).

Not restated (asserted against): fine boxes that touch a NON-periodic domain face (the coarse values AMReX
interpolates from there depend on the internal box list of its FPinfo cache), and regridding itself (the box
lists of every level are inputs, as they are for the C ABI).
"""
from __future__ import annotations

import ctypes as C
import itertools

import numpy as np

from . import oracle as O

NQ = O.NQ
SMALL_NUM = np.finfo(np.float64).eps * 1e10  # constants::SMALL_NUM, Source/Constants.H:57-58


def _ptr(a, ctype=C.c_double):
    return a.ctypes.data_as(C.POINTER(ctype))


def box_intersect(alo, ahi, blo, bhi):
    lo = [max(a, b) for a, b in zip(alo, blo)]
    hi = [min(a, b) for a, b in zip(ahi, bhi)]
    return (lo, hi) if all(l <= h for l, h in zip(lo, hi)) else None


def hash_order(boxes):
    """Order in which BoxArray::intersections visits the boxes of a BoxArray: hash bins keyed by
    coarsen(smallEnd, max box extent), visited x fastest, then y, then z; inside a bin by index."""
    maxext = [max(hi[d] - lo[d] + 1 for lo, hi in boxes) for d in range(3)]
    key = lambda n: tuple((boxes[n][0][d] // maxext[d]) for d in (2, 1, 0)) + (n,)
    return sorted(range(len(boxes)), key=key)


def periodic_shifts(periodic, n, nghost):
    """Periodicity::shiftIntVect: x outermost, z innermost, -per..per in steps of the period."""
    rng = []
    for d in range(3):
        if periodic[d]:
            per = n[d]
            while per < nghost:
                per += n[d]
            rng.append(range(-per, per + 1, n[d]))
        else:
            rng.append([0])
    return [s for s in itertools.product(*rng)]


class Box:
    """One FAB of a level: reference-shaped arrays [comp, k, j, i] (f, g, is_fluid with 3 ghost cells,
    macrodata with 1)."""

    def __init__(self, lev, lo, hi):
        self.lo, self.hi = list(lo), list(hi)
        self.n = [hi[d] - lo[d] + 1 for d in range(3)]
        p = O._Params.from_buffer_copy(lev.params)
        for d in range(3):
            p.lo[d], p.hi[d] = lo[d], hi[d]
        self.p = p
        ng = p.ng
        gs = lambda g: tuple(self.n[d] + 2 * g for d in (2, 1, 0))
        self.f = np.zeros((NQ,) + gs(ng))
        self.g = np.zeros((NQ,) + gs(ng))
        self.is_fluid = np.ones((2,) + gs(ng), dtype=np.int32)
        self.is_fluid[1] = 0
        self.macro = np.zeros((O.NMACRO,) + gs(1))
        self.derived = np.zeros((O.NDERIVED,) + gs(0))
        self.eq = np.zeros((NQ,) + gs(0))
        self.eq_g = np.zeros((NQ,) + gs(0))

    def grown_index(self, ng):
        """global (k, j, i) index arrays of the box grown by ng"""
        return [np.arange(self.lo[d] - ng, self.hi[d] + ng + 1) for d in (2, 1, 0)]


class Level:
    def __init__(self, lev, setup: O.Setup, boxes, ref_ratio=2):
        self.lev = lev
        p0 = setup.params
        r = ref_ratio ** lev
        p = O._Params.from_buffer_copy(p0)
        self.n = [setup.n[d] * r for d in range(3)]
        for d in range(3):
            p.dom_lo[d], p.dom_hi[d] = 0, self.n[d] - 1
            p.dx[d] = p0.dx[d] / r
            p.inv_dx[d] = 1.0 / p.dx[d]
        p.dt = p0.dt / r  # m_dts[lev] = m_dts[lev-1] / MaxRefRatio, Source/LBM.cpp:1073-1076
        self.params = p
        self.periodic = [int(p.periodic[d]) for d in range(3)]
        self.boxes = [Box(self, lo, hi) for lo, hi in boxes]
        self.time = 0.0
        self.cover = np.full(tuple(self.n[d] for d in (2, 1, 0)), -1, dtype=np.int32)
        for n, b in enumerate(self.boxes):
            self.cover[b.lo[2]:b.hi[2] + 1, b.lo[1]:b.hi[1] + 1, b.lo[0]:b.hi[0] + 1] = n

    # dense view of the valid data of every box (NaN where the level has no box)
    def gather(self, name, ncomp, dtype=np.float64):
        G = np.full((ncomp,) + self.cover.shape, np.nan if dtype == np.float64 else -9, dtype=dtype)
        for b in self.boxes:
            a = getattr(b, name)
            ng = (a.shape[1] - b.n[2]) // 2
            v = a if ng == 0 else a[:, ng:-ng, ng:-ng, ng:-ng]
            G[:, b.lo[2]:b.hi[2] + 1, b.lo[1]:b.hi[1] + 1, b.lo[0]:b.hi[0] + 1] = v
        return G

    def wrapped(self, idx):
        """global index arrays -> (wrapped index arrays, inside-the-periodic-grown-domain masks)"""
        out, ok = [], []
        for a, d in zip(idx, (2, 1, 0)):
            if self.periodic[d]:
                out.append(a % self.n[d])
                ok.append(np.ones(a.shape, bool))
            else:
                out.append(np.clip(a, 0, self.n[d] - 1))
                ok.append((a >= 0) & (a < self.n[d]))
        return out, ok

    def fill_boundary(self, name, ng):
        """FillBoundary(periodicity) of `name` over ng ghost cells"""
        a0 = getattr(self.boxes[0], name)
        G = self.gather(name, a0.shape[0], a0.dtype)
        for b in self.boxes:
            a = getattr(b, name)
            nga = (a.shape[1] - b.n[2]) // 2
            idx = b.grown_index(ng)
            w, ok = self.wrapped(idx)
            K, J, I = np.meshgrid(*w, indexing="ij")
            m = ok[0][:, None, None] & ok[1][None, :, None] & ok[2][None, None, :]
            m &= self.cover[K, J, I] >= 0
            # valid cells of the box itself keep their values
            m[ng:ng + b.n[2], ng:ng + b.n[1], ng:ng + b.n[0]] = False
            s = slice(nga - ng, a.shape[1] - (nga - ng)), slice(nga - ng, a.shape[2] - (nga - ng)), \
                slice(nga - ng, a.shape[3] - (nga - ng))
            view = a[(slice(None),) + s]
            view[:, m] = G[:, K[m], J[m], I[m]]


class AmrOracle:
    """Multi-level state driven in the reference's order (LBM::evolve / time_step / advance)."""

    def __init__(self, setup: O.Setup, level_boxes, is_fluid=None):
        """level_boxes[lev] = [(lo, hi), ...] valid boxes in the index space of level lev;
        is_fluid[lev] = dense int array over the level domain (1 where no box: unused) or None (all fluid)."""
        self.setup = setup
        self.levels = [Level(l, setup, bxs) for l, bxs in enumerate(level_boxes)]
        self.finest = len(self.levels) - 1
        self.lib = O.lib()
        for l, L in enumerate(self.levels):
            fl = None if is_fluid is None else is_fluid[l]
            self._set_is_fluid(L, fl)

    # ------------------------------------------------------------------ setup
    def _set_is_fluid(self, L: Level, dense):
        """comp 0 from the dense field (ghost cells: the value of the cell they lie on, periodic images
        included; fluid beyond non-periodic faces and where the level has no box), comp 1 = eb_boundary"""
        ng = L.params.ng
        for b in L.boxes:
            if dense is not None:
                idx = b.grown_index(ng)
                w, ok = L.wrapped(idx)
                K, J, I = np.meshgrid(*w, indexing="ij")
                m = ok[0][:, None, None] & ok[1][None, :, None] & ok[2][None, None, :]
                v = np.where(m, dense[K, J, I], 1).astype(np.int32)
                b.is_fluid[0] = v
            self.lib.orc_eb_boundary(C.byref(b.p), _ptr(b.is_fluid, C.c_int))
        L.fill_boundary("is_fluid", ng)

    def initialize(self):
        """MakeNewLevelFromScratch on every level (Source/LBM.cpp:1148-1199), then average_down (:183)"""
        for L in self.levels:
            for b in L.boxes:
                self.lib.orc_initialize(C.byref(b.p), C.byref(self.setup.ic), _ptr(b.is_fluid, C.c_int), _ptr(b.f),
                                        _ptr(b.g))
            L.fill_boundary("f", L.params.ng)
            L.fill_boundary("g", L.params.ng)
            self._macrodata(L)
            for b in L.boxes:
                self.lib.orc_macrodata_to_equilibrium(C.byref(b.p), _ptr(b.is_fluid, C.c_int), _ptr(b.macro),
                                                      _ptr(b.derived), _ptr(b.eq), _ptr(b.eq_g))
                self.lib.orc_compute_derived(C.byref(b.p), _ptr(b.is_fluid, C.c_int), _ptr(b.macro), _ptr(b.derived))
                self.lib.orc_compute_q_corrections(C.byref(b.p), _ptr(b.is_fluid, C.c_int), _ptr(b.macro),
                                                   _ptr(b.derived))
        for lev in range(self.finest - 1, -1, -1):
            self.average_down_to(lev, ng=0)

    # ---------------------------------------------------------------- operators
    def _macrodata(self, L):
        for b in L.boxes:
            self.lib.orc_f_to_macrodata(C.byref(b.p), _ptr(b.is_fluid, C.c_int), _ptr(b.f), _ptr(b.g), _ptr(b.macro), 0)
        L.fill_boundary("macro", 1)  # Source/LBM.cpp:905

    def stream(self, lev):
        L = self.levels[lev]
        for name in ("f", "g"):
            for b in L.boxes:
                self.lib.orc_stream(C.byref(b.p), _ptr(b.is_fluid, C.c_int), _ptr(getattr(b, name)), 0)
            L.fill_boundary(name, L.params.ng)  # Source/LBM.cpp:603

    def collide(self, lev):
        L = self.levels[lev]
        self._macrodata(L)
        for b in L.boxes:
            p, fl = C.byref(b.p), _ptr(b.is_fluid, C.c_int)
            self.lib.orc_compute_q_corrections(p, fl, _ptr(b.macro), _ptr(b.derived))
            self.lib.orc_macrodata_to_equilibrium(p, fl, _ptr(b.macro), _ptr(b.derived), _ptr(b.eq), _ptr(b.eq_g))
            self.lib.orc_relax(p, fl, _ptr(b.macro), _ptr(b.eq), _ptr(b.eq_g), _ptr(b.f), _ptr(b.g), 0)
        L.fill_boundary("f", L.params.ng)  # Source/LBM.cpp:805-806
        L.fill_boundary("g", L.params.ng)

    def physbc(self, lev):
        L = self.levels[lev]
        for b in L.boxes:
            self.lib.orc_physbc(C.byref(b.p), _ptr(b.f), 0, C.c_double(L.time))
            self.lib.orc_physbc(C.byref(b.p), _ptr(b.g), 1, C.c_double(L.time))

    def fillpatch(self, lev):
        """FillPatchOps::fillpatch(lev, time, m_f[lev]) and the same for g"""
        L = self.levels[lev]
        for b in L.boxes:
            self.lib.orc_prepass(C.byref(b.p), _ptr(b.f))  # K6, FillPatchOps.H:92-108
            self.lib.orc_prepass(C.byref(b.p), _ptr(b.g))
        if lev > 0:
            self._interp_from_coarse(lev, "f")
            self._interp_from_coarse(lev, "g")
        L.fill_boundary("f", L.params.ng)
        L.fill_boundary("g", L.params.ng)
        self.physbc(lev)

    def _interp_from_coarse(self, lev, name, ratio=2, target=None, cover=None):
        """cells of the target boxes (default: level lev itself) grown by the ghost width, inside the (periodically
        grown) domain, that no valid cell of `cover` (default: level lev's own boxes) lies on -- FPinfo: complementIn
        WITHOUT periodic shifts: CellConservativeLinear from level lev-1.  With another `target` / `cover` this is
        the fill of a re-made level from the old one (RemakeLevel)."""
        Lf = self.levels[lev] if target is None else target
        Lc = self.levels[lev - 1]
        cover_mask = Lf.cover if cover is None else cover
        ng = Lf.params.ng
        Gc = Lc.gather(name, NQ)
        for b in Lf.boxes:
            a = getattr(b, name)
            idx = b.grown_index(ng)  # fine global indices k, j, i
            # leftover mask: inside dstdomain (domain grown by ng in periodic directions), not on a valid cell
            m = np.ones(a.shape[1:], bool)
            for ax, d in enumerate((2, 1, 0)):
                if not Lf.periodic[d]:
                    ok = (idx[ax] >= 0) & (idx[ax] < Lf.n[d])
                    sh = [1, 1, 1]
                    sh[ax] = -1
                    m &= ok.reshape(sh)
            cl = [np.clip(x, 0, Lf.n[d] - 1) for x, d in zip(idx, (2, 1, 0))]
            K, J, I = np.meshgrid(*cl, indexing="ij")
            ins = np.ones(a.shape[1:], bool)
            for ax, d in enumerate((2, 1, 0)):
                ok = (idx[ax] >= 0) & (idx[ax] < Lf.n[d])
                sh = [1, 1, 1]
                sh[ax] = -1
                ins &= ok.reshape(sh)
            m &= ~(ins & (cover_mask[K, J, I] >= 0))
            if not m.any():
                continue
            # coarse patch: coarsen(grown box) grown by 1, filled from the coarse level's valid cells (periodic)
            clo = [(b.lo[d] - ng) // ratio - 1 for d in range(3)]
            chi = [(b.hi[d] + ng) // ratio + 1 for d in range(3)]
            cidx = [np.arange(clo[d], chi[d] + 1) for d in (2, 1, 0)]
            w, ok = Lc.wrapped(cidx)
            cp = Gc[:, w[0][:, None, None], w[1][None, :, None], w[2][None, None, :]]
            okc = ok[0][:, None, None] & ok[1][None, :, None] & ok[2][None, None, :]
            cp = np.where(okc[None], cp, np.nan)
            u0 = cp[:, 1:-1, 1:-1, 1:-1]

            def slope(ax):
                sl = lambda o: tuple(slice(1 + o, cp.shape[x + 1] - 1 + o) if x == ax else slice(1, -1) for x in range(3))
                up, um = cp[(slice(None),) + sl(1)], cp[(slice(None),) + sl(-1)]
                dc = 0.5 * (up - um)
                df = 2.0 * (up - u0)
                db = 2.0 * (u0 - um)
                with np.errstate(invalid="ignore"):
                    s = np.where(df * db >= 0.0, np.minimum(np.abs(df), np.abs(db)), 0.0)
                    return np.copysign(1.0, dc) * np.minimum(s, np.abs(dc))

            sz, sy, sx = slope(0), slope(1), slope(2)
            with np.errstate(invalid="ignore", divide="ignore"):
                dumax = np.abs(sx) * 0.25 + np.abs(sy) * 0.25 + np.abs(sz) * 0.25
                umax, umin = u0.copy(), u0.copy()
                for dk, dj, di in itertools.product((-1, 0, 1), repeat=3):
                    nb = cp[:, 1 + dk:cp.shape[1] - 1 + dk, 1 + dj:cp.shape[2] - 1 + dj, 1 + di:cp.shape[3] - 1 + di]
                    umin = np.minimum(umin, nb)
                    umax = np.maximum(umax, nb)
                alpha = np.ones_like(u0)
                any_s = (sx != 0.0) | (sy != 0.0) | (sz != 0.0)
                c1 = any_s & (dumax * alpha > (umax - u0))
                alpha = np.where(c1, (umax - u0) / dumax, alpha)
                c2 = any_s & (dumax * alpha > (u0 - umin))
                alpha = np.where(c2, (u0 - umin) / dumax, alpha)
            sx, sy, sz = sx * alpha, sy * alpha, sz * alpha
            # fine cells: parent index relative to the slope array, offsets -+ 0.25
            par = [np.floor_divide(x, ratio) - (clo[d] + 1) for x, d in zip(idx, (2, 1, 0))]
            off = [((x - np.floor_divide(x, ratio) * ratio) + 0.5) / ratio - 0.5 for x in idx]
            PK, PJ, PI = np.meshgrid(*par, indexing="ij")
            OK_, OJ, OI = np.meshgrid(*off, indexing="ij")
            val = u0[:, PK, PJ, PI] + OI * sx[:, PK, PJ, PI] + OJ * sy[:, PK, PJ, PI] + OK_ * sz[:, PK, PJ, PI]
            # guard: slopes of coarse cells on a non-periodic domain face use one-sided formulas that depend on the
            # extents of AMReX's internal coarse patch (AMReX_MFInterp_C.H:15-33) and on ghost values its BCFill put
            # there: not restated.  Only cells that SURVIVE matter: the copy from the fine level that follows
            # (FillPatchSingleLevel, periodic images included) overwrites every cell a box of `cover` lies on, so a fine
            # box may touch a non-periodic face as long as no coarse-fine interface cell has its parent there.
            wi, wok = Lf.wrapped(idx)
            WK, WJ, WI = np.meshgrid(*wi, indexing="ij")
            wm = wok[0][:, None, None] & wok[1][None, :, None] & wok[2][None, None, :]
            surv = m & ~(wm & (cover_mask[WK, WJ, WI] >= 0))
            for ax, d in enumerate((2, 1, 0)):
                if Lc.periodic[d]:
                    continue
                pc = np.floor_divide(idx[ax], ratio)
                sel = surv.any(axis=tuple(x for x in range(3) if x != ax))
                if ((pc[sel] <= 0) | (pc[sel] >= Lc.n[d] - 1)).any():
                    raise NotImplementedError("coarse-fine interpolation next to a non-periodic domain face is not restated")
            a[:, m] = val[:, m]  # NaN (see average_down_to) can only reach cells the FillBoundary below overwrites

    def average_down_to(self, crse_lev, ng=1, ratio=2):
        """average_down_with_ghosts(m_f[crse_lev+1], m_f[crse_lev], geom, ng, ratio) and the same for g;
        ng = 1 inside advance (Source/LBM.cpp:539-541), ng = 0 after initialisation (Source/LBM.cpp:167)"""
        Lf, Lc = self.levels[crse_lev + 1], self.levels[crse_lev]
        fboxes = [(b.lo, b.hi) for b in Lf.boxes]
        order = hash_order(fboxes)
        shifts = periodic_shifts(Lc.periodic, Lc.n, 0)
        for name in ("f", "g"):
            Gc = Lc.gather(name, NQ)
            cfine = []
            for b in Lf.boxes:
                a = getattr(b, name)  # 3 ghost cells
                clo = [b.lo[d] // ratio - ng for d in range(3)]
                chi = [b.hi[d] // ratio + ng for d in range(3)]
                cidx = [np.arange(clo[d], chi[d] + 1) for d in (2, 1, 0)]
                # cfine.ParallelCopy(crse, ..., src ng 0, dst ng 1): NOT periodic; cells on no coarse valid cell
                # stay uninitialised (NaN here)
                ok = [(x >= 0) & (x < Lc.n[d]) for x, d in zip(cidx, (2, 1, 0))]
                cl = [np.clip(x, 0, Lc.n[d] - 1) for x, d in zip(cidx, (2, 1, 0))]
                cf = Gc[:, cl[0][:, None, None], cl[1][None, :, None], cl[2][None, None, :]].copy()
                okc = ok[0][:, None, None] & ok[1][None, :, None] & ok[2][None, None, :]
                cf[:, ~okc] = np.nan
                # masked_avgdown over the coarsened box grown by 1 = fine valid cells + 2 ghost layers
                t = 3 - ratio * ng  # fine ghost layers not under the coarsened box grown by ng
                fa = a[:, t:a.shape[1] - t, t:a.shape[2] - t, t:a.shape[3] - t]
                nzc, nyc, nxc = cf.shape[1:]
                f8 = fa.reshape(NQ, nzc, 2, nyc, 2, nxc, 2)
                c = np.zeros_like(cf)
                vol = np.zeros_like(cf)
                for kr, jr, ir in itertools.product((0, 1), repeat=3):  # kref outermost, iref innermost
                    fv = f8[:, :, kr, :, jr, :, ir]
                    use = np.abs(fv - (-1.0)) > SMALL_NUM
                    c = np.where(use, c + fv, c)
                    vol = np.where(use, vol + 1.0, vol)
                with np.errstate(invalid="ignore", divide="ignore"):
                    cf = np.where(vol > 0.0, c / vol, cf)
                cfine.append((clo, chi, cf))
            # crse.ParallelCopy(cfine, src ng 1, dst ng 0, periodicity): tags in CPC order, later tags overwrite
            for cb in Lc.boxes:
                dst = getattr(cb, name)
                ngc = Lc.params.ng
                for sh in shifts:
                    for n in order:
                        clo, chi, cf = cfine[n]
                        slo = [clo[d] + sh[d] for d in range(3)]
                        shi = [chi[d] + sh[d] for d in range(3)]
                        it = box_intersect(cb.lo, cb.hi, slo, shi)
                        if it is None:
                            continue
                        lo, hi = it
                        d_s = tuple(slice(lo[d] - cb.lo[d] + ngc, hi[d] - cb.lo[d] + ngc + 1) for d in (2, 1, 0))
                        s_s = tuple(slice(lo[d] - slo[d], hi[d] - slo[d] + 1) for d in (2, 1, 0))
                        src = cf[(slice(None),) + s_s]
                        # NaN = a cfine cell the reference leaves uninitialised (outside the coarse domain in a periodic
                        # direction, all eight fine cells masked: solid cells of a body that crosses the periodic face);
                        # the reference copies whatever its arena held there.  Only solid cells can receive it.
                        dst[(slice(None),) + d_s] = src

    def regrid_level(self, lev, new_boxes, is_fluid_dense=None):
        """LBM::RemakeLevel (Source/LBM.cpp:1302-1364) for a box list that AmrCore::regrid changed: the new level's f, g
        by FillPatchOps::fillpatch into NEW MultiFabs -- K6 pre-pass on the old level, cells the OLD level's valid
        boxes do not cover by interpolation from level lev-1, the rest copied from the old level (periodic images
        included), BCFill -- then is_fluid, fill_f_inside_eb (zero in solid cells, LBM.cpp:1278-1298) and FillBoundary."""
        assert lev >= 1
        old = self.levels[lev]
        for b in old.boxes:
            self.lib.orc_prepass(C.byref(b.p), _ptr(b.f))
            self.lib.orc_prepass(C.byref(b.p), _ptr(b.g))
        new = Level(lev, self.setup, new_boxes)
        new.time = old.time
        ng = new.params.ng
        for name in ("f", "g"):
            self._interp_from_coarse(lev, name, target=new, cover=old.cover)
            G = old.gather(name, NQ)  # FillPatchSingleLevel(mf, nghost, {old fine}): valid + ghost cells, periodic
            for b in new.boxes:
                a = getattr(b, name)
                w, ok = new.wrapped(b.grown_index(ng))
                K, J, I = np.meshgrid(*w, indexing="ij")
                m = ok[0][:, None, None] & ok[1][None, :, None] & ok[2][None, None, :]
                m &= old.cover[K, J, I] >= 0
                a[:, m] = G[:, K[m], J[m], I[m]]
        self.levels[lev] = new
        self.physbc(lev)
        self._set_is_fluid(new, is_fluid_dense)
        for b in new.boxes:  # fill_f_inside_eb
            solid = b.is_fluid[0] == 0
            b.f[:, solid] = 0.0
            b.g[:, solid] = 0.0
        new.fill_boundary("f", ng)
        new.fill_boundary("g", ng)
        self._macrodata(new)

    def make_level_from_coarse(self, lev, boxes, is_fluid_dense=None):
        """LBM::MakeNewLevelFromCoarse (Source/LBM.cpp:1088-1144), a level AmrCore::regrid adds above the finest one:
        is_fluid, then f and g of EVERY cell (valid and ghost, inside the periodically grown domain) by
        FillPatchOps::fillpatch_from_coarse = InterpFromCoarseLevel (CellConservativeLinear from level lev-1) and the
        fine BCFill -- no K6 pre-pass, no fill_f_inside_eb, no FillBoundary -- then the macrodata."""
        assert lev == len(self.levels) and lev >= 1
        new = Level(lev, self.setup, boxes)
        new.time = self.levels[lev - 1].time
        self._set_is_fluid(new, is_fluid_dense)
        nothing = np.full_like(new.cover, -1)
        for name in ("f", "g"):
            self._interp_from_coarse(lev, name, target=new, cover=nothing)
        self.levels.append(new)
        self.finest = lev
        self.physbc(lev)
        self._macrodata(new)

    def clear_level(self, lev):
        """LBM::ClearLevel (Source/LBM.cpp:1367-1380): the finest level disappears in a regrid"""
        assert lev == self.finest and lev >= 1
        self.levels.pop()
        self.finest = lev - 1

    def eb_forces(self):
        """LBM::compute_eb_forces (Source/LBM.cpp:994-1044): momentum exchange over the eb_boundary cells of EVERY level,
        summed without any weighting by the cell size.  m_mask[lev] is empty in every run the reference completes
        (tests/golden/make_golden.py, amr3_chcyl_forces), so the coarse cells under a fine level count as well."""
        total = np.zeros(3)
        for L in self.levels:
            for b in L.boxes:
                out = (C.c_double * 3)()
                self.lib.orc_eb_forces(C.byref(b.p), _ptr(b.is_fluid, C.c_int), _ptr(b.f), out)
                total += np.array(out[:])
        return total

    # ------------------------------------------------------------ time stepping
    def advance(self, lev):
        """LBM::advance (Source/LBM.cpp:523-544)"""
        self.stream(lev)
        if lev < self.finest:
            self.average_down_to(lev)
        self.collide(lev)
        self.levels[lev].time += self.levels[lev].params.dt

    def time_step(self, lev):
        """LBM::time_step without regridding (Source/LBM.cpp:452-521)"""
        if lev < self.finest:
            self.fillpatch(lev + 1)
            for _ in range(2):  # m_nsubsteps[lev + 1] = MaxRefRatio(lev)
                self.physbc(lev + 1)
                self.time_step(lev + 1)
        self.advance(lev)

    def post_time_step(self):
        for L in self.levels:
            for b in L.boxes:
                self.lib.orc_compute_derived(C.byref(b.p), _ptr(b.is_fluid, C.c_int), _ptr(b.macro), _ptr(b.derived))

    def step(self, nsteps=1):
        """LBM::evolve body (Source/LBM.cpp:416-422)"""
        for _ in range(nsteps):
            self.fillpatch(0)
            self.time_step(0)
            self.post_time_step()

    # ------------------------------------------------------------------ views
    def fields(self, lev):
        """dense valid-cell fields of one level under the plotfile names (NaN where the level has no box)"""
        L = self.levels[lev]
        out = {}
        m = L.gather("macro", O.NMACRO)
        for n, name in enumerate(O.MACRO_NAMES):
            out[name] = m[n]
        f, g = L.gather("f", NQ), L.gather("g", NQ)
        for q in range(NQ):
            out[f"f_{q:02d}"] = f[q]
            out[f"g_{q:02d}"] = g[q]
        d = L.gather("derived", O.NDERIVED)
        for n, name in enumerate(O.DERIVED_NAMES):
            out[name] = d[n]
        fl = L.gather("is_fluid", 2, np.int32)
        out["is_fluid"] = np.where(fl[0] < 0, np.nan, fl[0].astype(float))
        out["eb_boundary"] = np.where(fl[1] < 0, np.nan, fl[1].astype(float))
        return out
