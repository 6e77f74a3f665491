"""is_fluid for the simple embedded-boundary bodies of the shipped decks.

In the reference the flag field comes from AMReX's EB2 cut-cell generator
(Source/EB.cpp:5-38 -> EB2::Build) and LBM::initialize_is_fluid marks a cell
fluid when it is regular or cut (Source/LBM.cpp:1222-1230).  EB generation stays
outside the accelerated path (the C ABI takes `is_fluid`
from the caller); this module only lets the stand-alone host mirror run decks
with `eb2.geom_type = all_regular | sphere | cylinder | box` without AMReX: a
cell is solid when all 8 of its corners lie inside the body, which is EB2's
covered-cell rule for these implicit functions.  tests/ check it against the
reference's own is_fluid output.
"""
from __future__ import annotations

import numpy as np


def _vec(deck, key, default):
    return [float(v) for v in deck[key]] if key in deck else list(default)


def body_from_deck(deck: dict):
    """(kind, parameters) of the deck's analytic body for mbl_set_body, which evaluates the flag field on the device
    (include/marbles_b200.h); None for a body the library does not know (the caller then supplies is_fluid)."""
    gtype = deck.get("eb2.geom_type", ["all_regular"])[0]
    if gtype == "all_regular":
        return 0, []
    if gtype == "sphere":
        c = _vec(deck, "eb2.sphere_center", [0, 0, 0])
        return 1, c + [float(deck["eb2.sphere_radius"][0]), float(int(deck.get("eb2.sphere_has_fluid_inside", ["0"])[0]))]
    if gtype == "cylinder":
        c = _vec(deck, "eb2.cylinder_center", [0, 0, 0])
        return 2, c + [float(deck["eb2.cylinder_radius"][0]), float(deck.get("eb2.cylinder_height", ["-1"])[0]),
                       float(int(deck.get("eb2.cylinder_direction", ["0"])[0])),
                       float(int(deck.get("eb2.cylinder_has_fluid_inside", ["0"])[0]))]
    if gtype == "box":
        return 3, _vec(deck, "eb2.box_lo", [0, 0, 0]) + _vec(deck, "eb2.box_hi", [0, 0, 0]) + \
            [float(int(deck.get("eb2.box_has_fluid_inside", ["0"])[0]))]
    return None


def _corner_coords(n, lo, prob_lo, dx, ng):
    """node coordinates of the grown box, per dimension (length n+2ng+1)"""
    return [prob_lo[d] + (np.arange(lo[d] - ng, lo[d] + n[d] + ng + 1)) * dx[d] for d in range(3)]


def is_fluid_from_deck(deck: dict, n_cell, prob_lo, dx, lo=(0, 0, 0), n_local=None, ng: int = 3) -> np.ndarray:
    """int32 array [nz+2ng, ny+2ng, nx+2ng] (component 0 of m_is_fluid) for the local box."""
    n = tuple(n_local) if n_local is not None else tuple(n_cell)
    gtype = deck.get("eb2.geom_type", ["all_regular"])[0]
    shape = tuple(n[d] + 2 * ng for d in (2, 1, 0))
    if gtype == "all_regular":
        return np.ones(shape, dtype=np.int32)
    xs, ys, zs = _corner_coords(n, lo, prob_lo, dx, ng)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    if gtype == "sphere":
        c = _vec(deck, "eb2.sphere_center", [0, 0, 0])
        r = float(deck["eb2.sphere_radius"][0])
        inside_body = (X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2 < r * r
        fluid_inside = int(deck.get("eb2.sphere_has_fluid_inside", ["0"])[0])
    elif gtype == "cylinder":
        c = _vec(deck, "eb2.cylinder_center", [0, 0, 0])
        r = float(deck["eb2.cylinder_radius"][0])
        h = float(deck.get("eb2.cylinder_height", ["-1"])[0])
        ax = int(deck.get("eb2.cylinder_direction", ["0"])[0])
        P = [X - c[0], Y - c[1], Z - c[2]]
        rad2 = sum(P[d] ** 2 for d in range(3) if d != ax)
        inside_body = rad2 < r * r
        if h > 0:
            inside_body &= np.abs(P[ax]) < 0.5 * h
        fluid_inside = int(deck.get("eb2.cylinder_has_fluid_inside", ["0"])[0])
    elif gtype == "box":
        blo = _vec(deck, "eb2.box_lo", [0, 0, 0])
        bhi = _vec(deck, "eb2.box_hi", [0, 0, 0])
        inside_body = ((X > blo[0]) & (X < bhi[0]) & (Y > blo[1]) & (Y < bhi[1]) & (Z > blo[2]) & (Z < bhi[2]))
        fluid_inside = int(deck.get("eb2.box_has_fluid_inside", ["0"])[0])
    else:
        raise ValueError(f"eb2.geom_type = {gtype} needs the caller to supply is_fluid (AMReX EB2 / STL)")
    solid_node = ~inside_body if fluid_inside else inside_body
    s = solid_node
    all_corners = (s[:-1, :-1, :-1] & s[1:, :-1, :-1] & s[:-1, 1:, :-1] & s[1:, 1:, :-1] &
                   s[:-1, :-1, 1:] & s[1:, :-1, 1:] & s[:-1, 1:, 1:] & s[1:, 1:, 1:])
    return np.where(all_corners, 0, 1).astype(np.int32)
