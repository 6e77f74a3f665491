"""Input decks: the reference's `key = value` deck format (AMReX ParmParse) and the
scalars LBM derives from it.

Mirrors LBM::read_parameters (Source/LBM.cpp:196-300), the inlet functor
constructors (Source/VelocityBC.cpp:6-54) and the initial-condition constructors
(Source/IC.cpp:6-141), so the shipped decks (Tests/test_files/*/*.inp) run
unchanged.  Keys this path does not use (amr.*, amrex.*, tagging.*) are kept in
the dict and ignored.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

R_U = 28.96  # universal gas constant in LB units, Source/LBM.H:311, Source/Constants.H:66

BC_PERIODIC, BC_NOSLIP, BC_VELOCITY, BC_PRESSURE, BC_OUTFLOW = 0, 1, 2, 3, 5
VBC_KINDS = {"noop": 0, "constant": 1, "channel": 2, "parabolic": 3}
IC_KINDS = {"constant": 0, "taylorgreen": 1, "viscosity_test": 2, "thermaldiffusivity_test": 3, "sod": 4}


class DeckError(ValueError):
    """Invalid input deck (the reference calls amrex::Abort)."""


def parse_deck(path: str | None = None, overrides=None, text: str | None = None) -> dict:
    """Parse a deck file plus command-line style overrides (`key=value` strings or a dict)."""
    lines: list[str] = []
    if path is not None:
        with open(path) as fh:
            lines += fh.read().splitlines()
    if text is not None:
        lines += text.splitlines()
    if isinstance(overrides, dict):
        lines += [f"{k} = {v}" for k, v in overrides.items()]
    elif overrides:
        lines += list(overrides)
    deck: dict[str, list[str]] = {}
    for line in lines:
        line = line.split("#", 1)[0].strip()
        if "=" not in line:
            continue
        key, val = line.split("=", 1)
        toks = re.findall(r'"[^"]*"|\S+', val.strip())
        deck[key.strip()] = [t.strip('"') for t in toks]
    return deck


def _scalar(deck, key, default, conv=float):
    return conv(deck[key][0]) if key in deck else default


def _vector(deck, key, default, conv=float):
    out = list(default)
    if key in deck:
        vals = [conv(v) for v in deck[key]]
        out[:len(vals)] = vals[:len(out)]
    return out


@dataclass
class LbmInputs:
    n_cell: tuple
    prob_lo: tuple
    prob_hi: tuple
    periodic: tuple
    bc_lo: tuple
    bc_hi: tuple
    nu: float
    alpha: float
    R: float
    gamma: float
    mesh_speed: float
    max_step: int
    # inlet functor
    vbc_kind: int = 0
    vbc_dir: int = 1
    vbc_normal_dir: int = 0
    vbc_tangential_dir: int = 1
    vbc_u: float = 0.0
    vbc_rho: float = 1.0
    vbc_T: float = 1.0 / 3.0
    vbc_gamma: float = 5.0 / 3.0
    vbc_R: float = 1.0
    # initial condition: kind + the 16 scalars mbl_initialize takes
    ic_kind: int = 0
    ic_params: list = field(default_factory=lambda: [0.0] * 16)
    deck: dict = field(default_factory=dict)

    @property
    def dx(self):
        return tuple((self.prob_hi[d] - self.prob_lo[d]) / self.n_cell[d] for d in range(3))


def lbm_inputs(deck: dict) -> LbmInputs:
    if "amr.n_cell" not in deck:
        raise DeckError("amr.n_cell is required")
    n = tuple(int(v) for v in deck["amr.n_cell"])
    if len(n) != 3:
        raise DeckError("only AMREX_SPACEDIM == 3 (D3Q27) is supported")
    periodic = tuple(_vector(deck, "geometry.is_periodic", [0, 0, 0], int))
    bc_lo = tuple(_vector(deck, "lbm.bc_lo", [0, 0, 0], int))
    bc_hi = tuple(_vector(deck, "lbm.bc_hi", [0, 0, 0], int))
    for b in bc_lo + bc_hi:  # Source/LBM.cpp:109-150: anything else aborts with "Invalid bc_lo" / "Invalid bc_hi"
        if b not in (0, 1, 2, 3, 5, 6, 7, 8):
            raise DeckError(f"Invalid bc_lo / bc_hi code {b}")
    for d in range(3):  # Source/LBM.cpp:227-252
        if periodic[d]:
            if bc_lo[d] != BC_PERIODIC:
                raise DeckError(f"BC is periodic in direction {d} but low BC is not 0")
            if bc_hi[d] != BC_PERIODIC:
                raise DeckError(f"BC is periodic in direction {d} but high BC is not 0")
        else:
            if bc_lo[d] == BC_PERIODIC or bc_hi[d] == BC_PERIODIC:
                raise DeckError(f"BC is interior in direction {d} but not periodic")
    has_vel = any(b == BC_VELOCITY for b in bc_lo + bc_hi)
    if has_vel and "lbm.velocity_bc_type" not in deck:  # Source/LBM.cpp:264-269
        raise DeckError("LBM::read_paramaters: velocity BC is used without specifying the type to be used")
    if "lbm.ic_type" not in deck:
        raise DeckError("lbm.ic_type is required")
    nu = _scalar(deck, "lbm.nu", 1.0)
    gamma = _scalar(deck, "lbm.adiabatic_exponent", 5.0 / 3.0)
    m_bar = _scalar(deck, "lbm.mean_molecular_mass", R_U)
    inp = LbmInputs(
        n_cell=n,
        prob_lo=tuple(_vector(deck, "geometry.prob_lo", [0.0, 0.0, 0.0])),
        prob_hi=tuple(_vector(deck, "geometry.prob_hi", [1.0, 1.0, 1.0])),
        periodic=periodic, bc_lo=bc_lo, bc_hi=bc_hi,
        nu=nu, alpha=_scalar(deck, "lbm.alpha", nu),  # Source/LBM.cpp:277-279
        R=R_U / m_bar, gamma=gamma,
        mesh_speed=_scalar(deck, "lbm.dx_outer", 1.0) / _scalar(deck, "lbm.dt_outer", 1.0),
        max_step=_scalar(deck, "max_step", 2 ** 31 - 1, int),
        deck=deck,
    )
    # inlet functor
    kind = deck.get("lbm.velocity_bc_type", ["noop"])[0]
    if kind not in VBC_KINDS:
        raise DeckError("LBM::set_bcs(): Unknown velocity BC")  # Source/LBM.cpp:1449
    inp.vbc_kind = VBC_KINDS[kind]
    pre = f"velocity_bc_{kind}."
    mach = _scalar(deck, pre + "Mach_ref", {"parabolic": 0.05}.get(kind, 0.005))
    inp.vbc_rho = _scalar(deck, pre + "initial_density", 1.0)
    inp.vbc_T = _scalar(deck, pre + "initial_temperature", 1.0 / 3.0)
    inp.vbc_gamma = _scalar(deck, pre + "adiabatic_exponent", 5.0 / 3.0)
    inp.vbc_R = R_U / _scalar(deck, pre + "mean_molecular_mass", R_U)
    inp.vbc_u = mach * math.sqrt(inp.vbc_gamma * inp.vbc_R * inp.vbc_T)
    inp.vbc_dir = _scalar(deck, pre + "dir", 1, int)
    inp.vbc_normal_dir = _scalar(deck, pre + "normal_dir", 0, int)
    inp.vbc_tangential_dir = _scalar(deck, pre + "tangential_dir", 1, int)
    # initial condition
    ictype = deck["lbm.ic_type"][0]
    if ictype not in IC_KINDS:
        raise DeckError("LBM::set_ics(): User must specify a valid initial condition")  # Source/LBM.cpp:1472
    inp.ic_kind = IC_KINDS[ictype]
    pre = f"ic_{ictype}."
    T0 = _scalar(deck, pre + "initial_temperature", 1.0 / 3.0)
    g_ic = _scalar(deck, pre + "adiabatic_exponent", 5.0 / 3.0)
    R_ic = R_U / _scalar(deck, pre + "mean_molecular_mass", R_U)
    density = _scalar(deck, pre + "density", 1.0)
    v0, omega = 1.0, [1.0, 1.0, 1.0]
    if ictype == "taylorgreen":  # reads rho0, v0, omega only (Source/IC.cpp:36-49)
        density = _scalar(deck, pre + "rho0", 1.0)
        v0 = _scalar(deck, pre + "v0", 1.0)
        omega = _vector(deck, pre + "omega", [1.0, 1.0, 1.0])
        T0, g_ic, R_ic = 1.0 / 3.0, 5.0 / 3.0, 1.0
    c_s = math.sqrt(g_ic * R_ic * T0)
    machs = _vector(deck, pre + "mach_components", [0.0, 0.0, 0.0])
    vel = [m * c_s for m in machs]  # velocity is overwritten by mach_components * c_s (IC.cpp:30-32)
    inp.ic_params = [density, vel[0], vel[1], vel[2], v0, omega[0], omega[1], omega[2],
                     _scalar(deck, pre + "wave_length", 1.0), T0, g_ic, R_ic, c_s,
                     _scalar(deck, pre + "density_ratio", 1.0), _scalar(deck, pre + "temperature_ratio", 1.0),
                     _scalar(deck, pre + "x_discontinuity", 10.0)]
    return inp
