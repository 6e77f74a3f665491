"""ctypes binding of the C ABI in include/marbles_b200.h.

The library is built in-tree (marbles_b200/libmarbles_b200.so).  There is no
CPU path: if the library is missing or no CUDA device is present the calls fail
loudly (MarblesError), they never fall back to another implementation.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmarbles_b200.so")


class MarblesError(RuntimeError):
    """Raised for every nonzero return code of the C ABI (the reference aborts)."""


class Params(C.Structure):
    _fields_ = [
        ("nu", C.c_double), ("alpha", C.c_double), ("R", C.c_double), ("gamma", C.c_double),
        ("mesh_speed", C.c_double), ("bc_type", C.c_int * 6), ("periodic", C.c_int * 3),
        ("vbc_kind", C.c_int), ("vbc_dir", C.c_int), ("vbc_normal_dir", C.c_int), ("vbc_tangential_dir", C.c_int),
        ("vbc_u", C.c_double), ("vbc_rho", C.c_double), ("vbc_T", C.c_double), ("vbc_gamma", C.c_double),
        ("vbc_R", C.c_double),
    ]


class LevelGeom(C.Structure):
    _fields_ = [
        ("dom_lo", C.c_int * 3), ("dom_hi", C.c_int * 3), ("lo", C.c_int * 3), ("hi", C.c_int * 3),
        ("dt", C.c_double), ("inv_dx", C.c_double * 3),
        ("prob_lo", C.c_double * 3), ("prob_hi", C.c_double * 3), ("dx", C.c_double * 3),
    ]


class Layout(C.Structure):
    _fields_ = [
        ("pitch", C.c_int64), ("plane_stride", C.c_int64), ("comp_stride", C.c_int64),
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("ox", C.c_int32), ("gy", C.c_int32), ("gz", C.c_int32),
        ("lattice_doubles", C.c_int64), ("state_bytes", C.c_int64),
    ]


# every symbol include/marbles_b200.h declares, with its argument types
_P = C.c_void_p
_D = C.POINTER(C.c_double)
# mbl_exchange_fn: (user, npeers, peers[], send[], nsend[], recv[], nrecv[], stream) -> int
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                          C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_void_p)

SYMBOLS = {
    "mbl_last_error": (C.c_char_p, []),
    "mbl_version": (C.c_int, []),
    "mbl_create": (C.c_int, [C.POINTER(Params), C.c_int, C.POINTER(_P)]),
    "mbl_destroy": (C.c_int, [_P]),
    "mbl_set_stream": (C.c_int, [_P, _P]),
    "mbl_sync": (C.c_int, [_P]),
    "mbl_level_layout": (C.c_int, [C.POINTER(LevelGeom), C.POINTER(Layout)]),
    "mbl_level_define": (C.c_int, [_P, C.c_int, C.POINTER(LevelGeom), _P]),
    "mbl_level_clear": (C.c_int, [_P, C.c_int]),
    "mbl_level_lattice_ptr": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P)]),
    "mbl_set_is_fluid": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.c_int]),
    "mbl_set_all_fluid": (C.c_int, [_P, C.c_int]),
    "mbl_set_body": (C.c_int, [_P, C.c_int, C.c_int, _D, C.c_int]),
    "mbl_upload": (C.c_int, [_P, C.c_int, C.c_int, _D, C.c_int]),
    "mbl_download": (C.c_int, [_P, C.c_int, C.c_int, _D, C.c_int]),
    "mbl_download_macrodata": (C.c_int, [_P, C.c_int, _D, C.c_int]),
    "mbl_download_derived": (C.c_int, [_P, C.c_int, _D]),
    "mbl_initialize": (C.c_int, [_P, C.c_int, C.c_int, _D, C.c_int]),
    "mbl_fillpatch": (C.c_int, [_P, C.c_int, C.c_double]),
    "mbl_physbc": (C.c_int, [_P, C.c_int, C.c_double]),
    "mbl_stream": (C.c_int, [_P, C.c_int]),
    "mbl_collide": (C.c_int, [_P, C.c_int, C.c_int]),
    "mbl_advance": (C.c_int, [_P, C.c_int, C.c_int]),
    "mbl_f_to_macrodata": (C.c_int, [_P, C.c_int]),
    "mbl_compute_derived": (C.c_int, [_P, C.c_int]),
    "mbl_eb_forces": (C.c_int, [_P, C.c_int, _D]),
    "mbl_macro_halo_doubles": (C.c_int64, [_P, C.c_int]),
    "mbl_macro_halo": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int]),
    "mbl_compute_derived_slab": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "mbl_step": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_int]),
    "mbl_step_local": (C.c_int, [_P, C.c_int, C.c_double, C.c_int]),
    "mbl_halo_doubles": (C.c_int64, [_P, C.c_int]),
    "mbl_set_halo_lean": (C.c_int, [_P, C.c_int]),
    "mbl_halo_pack": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "mbl_halo_unpack": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "mbl_step_split": (C.c_int, [_P, C.c_int, C.c_int]),
    "mbl_halo_pack_next": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "mbl_halo_unpack_next": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "mbl_step_host": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, _D, _D, C.c_int]),
    "mbl_step_host_begin": (C.c_int, [_P, C.c_int, _D, _D, C.c_int]),
    "mbl_step_host_finish": (C.c_int, [_P, C.c_int, _D, _D, C.c_int]),
    "mbl_level_define_boxes": (C.c_int, [_P, C.c_int, C.POINTER(LevelGeom), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mbl_level_num_boxes": (C.c_int, [_P, C.c_int]),
    "mbl_level_bind": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "mbl_box_set_is_fluid": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int]),
    "mbl_box_upload": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _D, C.c_int]),
    "mbl_box_download": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _D, C.c_int]),
    "mbl_box_download_macrodata": (C.c_int, [_P, C.c_int, C.c_int, _D, C.c_int, C.c_int]),
    "mbl_average_down": (C.c_int, [_P, C.c_int, C.c_int]),
    "mbl_level_regrid": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mbl_level_make_from_coarse": (C.c_int, [_P, C.c_int, C.POINTER(LevelGeom), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mbl_level_make_from_coarse_on": (C.c_int, [_P, C.c_int, C.POINTER(LevelGeom), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                                C.POINTER(C.c_int)]),
    "mbl_level_regrid_on": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mbl_level_define_boxes_on": (C.c_int, [_P, C.c_int, C.POINTER(LevelGeom), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                            C.POINTER(C.c_int)]),
    "mbl_level_box_owner": (C.c_int, [_P, C.c_int, C.c_int]),
    "mbl_fill_boundary_plan": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mbl_set_exchange": (C.c_int, [_P, C.c_int, C.c_int, EXCHANGE_FN, C.c_void_p]),
    "mbl_fill_f_inside_eb": (C.c_int, [_P, C.c_int]),
    "mbl_launch_count": (C.c_int64, [_P]),
    "mbl_set_variant": (C.c_int, [_P, C.c_int]),
    "mbl_get_variant": (C.c_int, [_P]),
    "mbl_set_timing": (C.c_int, [_P, C.c_int]),
    "mbl_get_timing": (C.c_int, [_P, _D, C.POINTER(C.c_int)]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MarblesError(
                f"{LIB_PATH} is missing: build it with `python -m marbles_b200.build` "
                "(marbles_b200 has no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise MarblesError(load().mbl_last_error().decode())
