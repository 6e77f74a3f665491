"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes/numpy front end of ``oracle/liboracle.so`` (the plain-C restatement of
the reference's lattice update, ``oracle/marbles_oracle.c``) plus the two helpers
the parity tests need around it:

* ``parse_deck`` / ``lbm_setup``: read a MARBLES ``.inp`` deck (+ ``key=value``
  overrides) and derive the scalars the reference derives in
  ``Source/LBM.cpp:196-300``, ``Source/VelocityBC.cpp:6-54``, ``Source/IC.cpp:6-141``.
* ``read_plotfile``: read an AMReX plotfile written by the reference
  (``Source/LBM.cpp:1629-1690``; format: SURVEY.md section 3.5).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
``marbles_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import re
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_SERIAL = os.path.join(REF_DIR, "marbles3d.ex")
REF_OMP = os.path.join(REF_DIR, "marbles3d.omp.ex")

NQ, NMACRO, NDERIVED = 27, 19, 7
MACRO_NAMES = ["rho", "vel_x", "vel_y", "vel_z", "vel_mag", "two_rho_e", "QCorrX", "QCorrY", "QCorrZ",
               "pxx", "pyy", "pzz", "pxy", "pxz", "pyz", "qx", "qy", "qz", "temperature"]
DERIVED_NAMES = ["vort_x", "vort_y", "vort_z", "vort_mag", "dQCorrX", "dQCorrY", "dQCorrZ"]
R_U = 28.96  # Source/Constants.H:66


def build(force: bool = False) -> str:
    """Compile liboracle.so (gcc, seconds)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
            os.path.join(HERE, "marbles_oracle.c")):
        subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)
    return LIB_PATH


class _Params(C.Structure):
    _fields_ = [
        ("dom_lo", C.c_int * 3), ("dom_hi", C.c_int * 3), ("lo", C.c_int * 3), ("hi", C.c_int * 3),
        ("ng", C.c_int), ("periodic", C.c_int * 3), ("bc_type", C.c_int * 6),
        ("nu", C.c_double), ("alpha", C.c_double), ("R", C.c_double), ("gamma", C.c_double),
        ("mesh_speed", C.c_double), ("dt", C.c_double), ("inv_dx", C.c_double * 3),
        ("prob_lo", C.c_double * 3), ("prob_hi", C.c_double * 3), ("dx", C.c_double * 3),
        ("vbc_kind", C.c_int), ("vbc_dir", C.c_int), ("vbc_normal_dir", C.c_int), ("vbc_tangential_dir", C.c_int),
        ("vbc_u", C.c_double), ("vbc_rho", C.c_double), ("vbc_T", C.c_double), ("vbc_gamma", C.c_double),
        ("vbc_R", C.c_double),
    ]


class _IC(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("density", C.c_double), ("velocity", C.c_double * 3), ("v0", C.c_double),
        ("omega", C.c_double * 3), ("wave_length", C.c_double), ("T0", C.c_double), ("gamma", C.c_double),
        ("R", C.c_double), ("c_s", C.c_double), ("density_ratio", C.c_double), ("temperature_ratio", C.c_double),
        ("x_discontinuity", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_ncell_grown.restype = C.c_long
        _lib.orc_check_stencil.restype = C.c_int
    return _lib


# --------------------------------------------------------------------------
# deck parsing (AMReX ParmParse subset: `key = v1 v2 ...`, '#' comments, quotes)
# --------------------------------------------------------------------------
def parse_deck(path: str | None, overrides: dict | list | None = None) -> dict:
    deck: dict[str, list[str]] = {}
    lines: list[str] = []
    if path is not None:
        with open(path) as fh:
            lines = fh.read().splitlines()
    if isinstance(overrides, dict):
        lines += [f"{k} = {v}" for k, v in overrides.items()]
    elif overrides:
        lines += list(overrides)
    for line in lines:
        line = line.split("#", 1)[0].strip()
        if "=" not in line:
            continue
        key, val = line.split("=", 1)
        toks = re.findall(r'"[^"]*"|\S+', val.strip())
        deck[key.strip()] = [t.strip('"') for t in toks]
    return deck


def _get(deck, key, default, conv=float):
    if key not in deck:
        return default
    vals = [conv(v) for v in deck[key]]
    if isinstance(default, (list, tuple)):
        out = list(default)
        out[:len(vals)] = vals[:len(out)]
        return out
    return vals[0]


VBC_KINDS = {"noop": 0, "constant": 1, "channel": 2, "parabolic": 3}
IC_KINDS = {"constant": 0, "taylorgreen": 1, "viscosity_test": 2, "thermaldiffusivity_test": 3, "sod": 4}


@dataclass
class Setup:
    """Everything the oracle needs for a single-level, single-box run."""
    n: tuple
    params: _Params
    ic: _IC
    deck: dict = field(default_factory=dict)


def lbm_setup(deck: dict, lo=None, hi=None, ng: int = 3) -> Setup:
    """Derive the reference's run-time scalars from a parsed deck."""
    n = [int(v) for v in deck["amr.n_cell"]]
    p = _Params()
    for d in range(3):
        p.dom_lo[d] = 0
        p.dom_hi[d] = n[d] - 1
        p.lo[d] = 0 if lo is None else lo[d]
        p.hi[d] = n[d] - 1 if hi is None else hi[d]
    p.ng = ng
    periodic = _get(deck, "geometry.is_periodic", [0, 0, 0], int)
    bc_lo = _get(deck, "lbm.bc_lo", [0, 0, 0], int)
    bc_hi = _get(deck, "lbm.bc_hi", [0, 0, 0], int)
    prob_lo = _get(deck, "geometry.prob_lo", [0.0, 0.0, 0.0])
    prob_hi = _get(deck, "geometry.prob_hi", [1.0, 1.0, 1.0])
    for d in range(3):
        p.periodic[d] = periodic[d]
        p.bc_type[d] = bc_lo[d]
        p.bc_type[d + 3] = bc_hi[d]
        p.prob_lo[d] = prob_lo[d]
        p.prob_hi[d] = prob_hi[d]
        # AMReX CoordSys: dx = (hi - lo)/n ; inv_dx = 1/dx
        p.dx[d] = (prob_hi[d] - prob_lo[d]) / n[d]
        p.inv_dx[d] = 1.0 / p.dx[d]
    p.nu = _get(deck, "lbm.nu", 1.0)
    p.alpha = _get(deck, "lbm.alpha", p.nu)  # Source/LBM.cpp:277-279
    gamma = _get(deck, "lbm.adiabatic_exponent", 5.0 / 3.0)
    m_bar = _get(deck, "lbm.mean_molecular_mass", R_U)
    p.R = R_U / m_bar
    p.gamma = gamma
    p.mesh_speed = _get(deck, "lbm.dx_outer", 1.0) / _get(deck, "lbm.dt_outer", 1.0)
    p.dt = 1.0  # est_time_step == 1, Source/LBM.cpp:1080-1084 (level 0)

    # inlet functor: Source/VelocityBC.cpp:6-54
    kind = deck.get("lbm.velocity_bc_type", ["noop"])[0]
    p.vbc_kind = VBC_KINDS[kind]
    pre = f"velocity_bc_{kind}."
    mach_default = {"constant": 0.005, "channel": 0.005, "parabolic": 0.05}.get(kind, 0.0)
    mach = _get(deck, pre + "Mach_ref", mach_default)
    p.vbc_rho = _get(deck, pre + "initial_density", 1.0)
    p.vbc_T = _get(deck, pre + "initial_temperature", 1.0 / 3.0)
    p.vbc_gamma = _get(deck, pre + "adiabatic_exponent", 5.0 / 3.0)
    p.vbc_R = R_U / _get(deck, pre + "mean_molecular_mass", R_U)
    p.vbc_u = mach * math.sqrt(p.vbc_gamma * p.vbc_R * p.vbc_T)
    p.vbc_dir = _get(deck, pre + "dir", 1, int)
    p.vbc_normal_dir = _get(deck, pre + "normal_dir", 0, int)
    p.vbc_tangential_dir = _get(deck, pre + "tangential_dir", 1, int)

    # initial condition: Source/IC.cpp:6-141
    ic = _IC()
    ictype = deck["lbm.ic_type"][0]
    ic.kind = IC_KINDS[ictype]
    pre = f"ic_{ictype}."
    ic.T0 = _get(deck, pre + "initial_temperature", 1.0 / 3.0)
    ic.gamma = _get(deck, pre + "adiabatic_exponent", 5.0 / 3.0)
    ic.R = R_U / _get(deck, pre + "mean_molecular_mass", R_U)
    ic.c_s = math.sqrt(ic.gamma * ic.R * ic.T0)
    vel = _get(deck, pre + "velocity", [0.0, 0.0, 0.0])
    machs = _get(deck, pre + "mach_components", [0.0, 0.0, 0.0])
    if ictype == "taylorgreen":
        ic.density = _get(deck, pre + "rho0", 1.0)
        ic.v0 = _get(deck, pre + "v0", 1.0)
        om = _get(deck, pre + "omega", [1.0, 1.0, 1.0])
        for d in range(3):
            ic.omega[d] = om[d]
        ic.T0 = 1.0 / 3.0  # TaylorGreen never reads initial_temperature (IC.cpp:38-49)
        ic.gamma = 5.0 / 3.0
        ic.R = 1.0
    else:
        ic.density = _get(deck, pre + "density", 1.0)
        for d in range(3):
            # the constructors overwrite velocity with mach_components * c_s (IC.cpp:30-32 ...)
            ic.velocity[d] = machs[d] * ic.c_s
        _ = vel
    ic.wave_length = _get(deck, pre + "wave_length", 1.0)
    ic.density_ratio = _get(deck, pre + "density_ratio", 1.0)
    ic.temperature_ratio = _get(deck, pre + "temperature_ratio", 1.0)
    ic.x_discontinuity = _get(deck, pre + "x_discontinuity", 10.0)
    return Setup(tuple(n), p, ic, deck)


# --------------------------------------------------------------------------
# state + stepping
# --------------------------------------------------------------------------
def _ptr(a, ctype=C.c_double):
    return a.ctypes.data_as(C.POINTER(ctype))


class Oracle:
    """Single-box reference-shaped state: f, g (ng ghosts), is_fluid (2 comps, ng ghosts),
    macrodata (1 ghost), derived/eq/eq_g (0 ghosts); arrays indexed [comp, k, j, i]."""

    def __init__(self, setup: Setup, is_fluid: np.ndarray | None = None):
        self.s = setup
        self.p = setup.params
        p = self.p
        self.ng = p.ng
        self.nv = tuple(p.hi[d] - p.lo[d] + 1 for d in range(3))
        gs = lambda g: tuple(self.nv[d] + 2 * g for d in (2, 1, 0))
        self.f = np.zeros((NQ,) + gs(self.ng))
        self.g = np.zeros((NQ,) + gs(self.ng))
        self.is_fluid = np.zeros((2,) + gs(self.ng), dtype=np.int32)
        self.macro = np.zeros((NMACRO,) + gs(1))
        self.derived = np.zeros((NDERIVED,) + gs(0))
        self.eq = np.zeros((NQ,) + gs(0))
        self.eq_g = np.zeros((NQ,) + gs(0))
        self.time = 0.0
        if is_fluid is None:
            self.is_fluid[0] = 1
        else:
            self.set_is_fluid(is_fluid)

    # -- geometry ----------------------------------------------------------
    def set_is_fluid(self, a: np.ndarray):
        """`a`: comp-0 flags either on the valid box [nz,ny,nx] (ghosts: periodic wrap, fluid
        beyond non-periodic faces) or on the grown box."""
        ng = self.ng
        if a.shape == self.is_fluid.shape[1:]:
            self.is_fluid[0] = a
        else:
            assert a.shape == tuple(self.nv[d] for d in (2, 1, 0)), a.shape
            self.is_fluid[0] = 1
            self.is_fluid[0][ng:-ng, ng:-ng, ng:-ng] = a
            lib().orc_fill_periodic_int(C.byref(self.p), _ptr(self.is_fluid, C.c_int), 1, ng)
        lib().orc_eb_boundary(C.byref(self.p), _ptr(self.is_fluid, C.c_int))
        lib().orc_fill_periodic_int(C.byref(self.p), _ptr(self.is_fluid, C.c_int), 2, ng)

    # -- reference call sequence --------------------------------------------
    def initialize(self):
        """LBM::initialize_f + f_to_macrodata.. (Source/LBM.cpp:1186-1198)"""
        L, p = lib(), C.byref(self.p)
        L.orc_initialize(p, C.byref(self.s.ic), _ptr(self.is_fluid, C.c_int), _ptr(self.f), _ptr(self.g))
        L.orc_fill_periodic(p, _ptr(self.f), NQ, self.ng)
        L.orc_fill_periodic(p, _ptr(self.g), NQ, self.ng)
        L.orc_f_to_macrodata(p, _ptr(self.is_fluid, C.c_int), _ptr(self.f), _ptr(self.g), _ptr(self.macro), 1)
        L.orc_macrodata_to_equilibrium(p, _ptr(self.is_fluid, C.c_int), _ptr(self.macro), _ptr(self.derived),
                                       _ptr(self.eq), _ptr(self.eq_g))
        L.orc_compute_derived(p, _ptr(self.is_fluid, C.c_int), _ptr(self.macro), _ptr(self.derived))
        L.orc_compute_q_corrections(p, _ptr(self.is_fluid, C.c_int), _ptr(self.macro), _ptr(self.derived))

    def fillpatch(self):
        L, p = lib(), C.byref(self.p)
        L.orc_fillpatch(p, _ptr(self.f), 0, C.c_double(self.time))
        L.orc_fillpatch(p, _ptr(self.g), 1, C.c_double(self.time))

    def stream(self):
        L, p = lib(), C.byref(self.p)
        L.orc_stream(p, _ptr(self.is_fluid, C.c_int), _ptr(self.f), 1)
        L.orc_stream(p, _ptr(self.is_fluid, C.c_int), _ptr(self.g), 1)

    def collide(self):
        lib().orc_collide(C.byref(self.p), _ptr(self.is_fluid, C.c_int), _ptr(self.f), _ptr(self.g),
                          _ptr(self.macro), _ptr(self.derived), _ptr(self.eq), _ptr(self.eq_g), 1)

    def post_time_step(self):
        lib().orc_compute_derived(C.byref(self.p), _ptr(self.is_fluid, C.c_int), _ptr(self.macro),
                                  _ptr(self.derived))

    def step(self, nsteps: int = 1):
        for _ in range(nsteps):
            self.fillpatch()
            self.stream()
            self.collide()
            self.post_time_step()
            self.time += self.p.dt

    def eb_forces(self):
        out = (C.c_double * 3)()
        lib().orc_eb_forces(C.byref(self.p), _ptr(self.is_fluid, C.c_int), _ptr(self.f), out)
        return np.array(out[:])

    # -- views ---------------------------------------------------------------
    def valid(self, a: np.ndarray, ng: int) -> np.ndarray:
        return a if ng == 0 else a[:, ng:-ng, ng:-ng, ng:-ng]

    @property
    def f_valid(self):
        return self.valid(self.f, self.ng)

    @property
    def g_valid(self):
        return self.valid(self.g, self.ng)

    @property
    def macro_valid(self):
        return self.valid(self.macro, 1)

    def fields(self) -> dict:
        """Same names as the reference plotfile (Source/LBM.cpp:302-340)."""
        out = {}
        for n, name in enumerate(MACRO_NAMES):
            out[name] = self.macro_valid[n]
        for q in range(NQ):
            out[f"f_{q:02d}"] = self.f_valid[q]
            out[f"g_{q:02d}"] = self.g_valid[q]
        for n, name in enumerate(DERIVED_NAMES):
            out[name] = self.derived[n]
        ng = self.ng
        out["is_fluid"] = self.is_fluid[0][ng:-ng, ng:-ng, ng:-ng].astype(float)
        out["eb_boundary"] = self.is_fluid[1][ng:-ng, ng:-ng, ng:-ng].astype(float)
        return out


def stencil():
    ev = (C.c_int * 81)()
    w = (C.c_double * 27)()
    b = [(C.c_int * 27)() for _ in range(4)]
    lib().orc_stencil(ev, w, *b)
    return (np.array(ev[:]).reshape(27, 3), np.array(w[:]), *[np.array(x[:]) for x in b])


# --------------------------------------------------------------------------
# AMReX plotfile reader (single level, level 0)
# --------------------------------------------------------------------------
def read_plotfile(path: str, level: int = 0) -> dict:
    with open(os.path.join(path, "Header")) as fh:
        lines = fh.read().split("\n")
    ncomp = int(lines[1])
    names = lines[2:2 + ncomp]
    pos = 2 + ncomp
    dim = int(lines[pos]); pos += 1
    time = float(lines[pos]); pos += 1
    finest = int(lines[pos]); pos += 1
    pos += 2  # prob_lo, prob_hi
    pos += 1  # ref ratios
    dom_line = lines[pos]
    doms = re.findall(r"\(\((-?\d+),(-?\d+),(-?\d+)\) \((-?\d+),(-?\d+),(-?\d+)\)", dom_line)
    lo = [int(v) for v in doms[level][0:3]]
    hi = [int(v) for v in doms[level][3:6]]
    n = [hi[d] - lo[d] + 1 for d in range(3)]
    data = np.full((ncomp, n[2], n[1], n[0]), np.nan)

    lev_dir = os.path.join(path, f"Level_{level}")
    with open(os.path.join(lev_dir, "Cell_H")) as fh:
        ch = fh.read().split("\n")
    i = 0
    while not ch[i].startswith("("):
        i += 1
    nbox = int(ch[i].strip("(").split()[0])
    boxes = []
    for b in range(nbox):
        m = re.findall(r"-?\d+", ch[i + 1 + b])
        boxes.append(([int(v) for v in m[0:3]], [int(v) for v in m[3:6]]))
    j = i + 1 + nbox
    while not ch[j].startswith("FabOnDisk"):
        j += 1
    for b in range(nbox):
        _, fname, off = ch[j + b].split()
        blo, bhi = boxes[b]
        bn = [bhi[d] - blo[d] + 1 for d in range(3)]
        with open(os.path.join(lev_dir, fname), "rb") as fh:
            fh.seek(int(off))
            fh.readline()  # FAB header line
            arr = np.frombuffer(fh.read(8 * ncomp * bn[0] * bn[1] * bn[2]), dtype="<f8")
        arr = arr.reshape(ncomp, bn[2], bn[1], bn[0])
        data[:, blo[2] - lo[2]:bhi[2] - lo[2] + 1, blo[1] - lo[1]:bhi[1] - lo[1] + 1,
             blo[0] - lo[0]:bhi[0] - lo[0] + 1] = arr
    out = {name: data[c] for c, name in enumerate(names)}
    out["__time__"] = time
    out["__names__"] = names
    _ = (dim, finest)
    return out


def read_plotfile_boxes(path: str, level: int = 0) -> list:
    """the valid boxes of one level of a plotfile: [[lo, hi], ...] (Level_k/Cell_H)"""
    with open(os.path.join(path, f"Level_{level}", "Cell_H")) as fh:
        ch = fh.read().split("\n")
    i = 0
    while not ch[i].startswith("("):
        i += 1
    nbox = int(ch[i].strip("(").split()[0])
    boxes = []
    for b in range(nbox):
        m = re.findall(r"-?\d+", ch[i + 1 + b])
        boxes.append([[int(v) for v in m[0:3]], [int(v) for v in m[3:6]]])
    return boxes


def run_reference(deck_path: str, workdir: str, overrides: list[str], omp: bool = False, threads: int | None = None,
                  timeout: float = 3600.0) -> str:
    """Run the unmodified reference executable (oracle/_ref) on a deck. Returns its stdout."""
    exe = REF_OMP if omp else REF_SERIAL
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    os.makedirs(workdir, exist_ok=True)
    res = subprocess.run([exe, deck_path] + list(overrides), cwd=workdir, env=env, capture_output=True, text=True,
                         timeout=timeout)
    if res.returncode != 0:
        raise RuntimeError(f"reference failed rc={res.returncode}\n{res.stdout[-2000:]}\n{res.stderr[-2000:]}")
    return res.stdout
