"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

    python tests/golden/make_golden.py          # needs oracle/_ref/marbles3d.ex (make -C oracle ref)

Each case is a small deck (written here, in the reference's deck format) run by
the reference executable built from /root/reference by oracle/refbuild/Makefile.
The plotfiles it writes (lbm.save_streaming=1, lbm.save_derived=1) are read back
with oracle.read_plotfile and stored as <case>.npz:

    deck        the deck text (tests re-parse it; nothing under /root/reference is read at test time)
    steps       plotted step numbers stored
    is_fluid    component 0 on valid cells (from plt00000)
    s<N>_<name> every plotfile component of step N  (all 82 for the last step; f,g + macro for step 0)

The reference ships no goldens of its own: these files are
the pin for the oracle, and through it for the CUDA path.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

COMMON = """
lbm.dx_outer = 1.0
lbm.dt_outer = 1.0
lbm.save_streaming = 1
lbm.save_derived = 1
amr.max_level = 0
amr.plot_int = 1
amr.chk_int = -1
amr.max_grid_size = 64
amr.blocking_factor = 2
amrex.fpe_trap_invalid = 1
amrex.fpe_trap_zero = 1
amrex.fpe_trap_overflow = 1
amrex.the_arena_is_managed = 0
"""

CASES = {
    # config 1 of BASELINE.json at reduced size: periodic Taylor-Green vortex, uniform T0 = 1/3
    "tg12": ("""
max_step = 10
geometry.prob_lo = -1.0 -1.0 -1.0
geometry.prob_hi =  1.0  1.0  1.0
geometry.is_periodic = 1 1 1
amr.n_cell = 12 12 12
lbm.bc_lo = 0 0 0
lbm.bc_hi = 0 0 0
lbm.nu = 0.1733333333333333
lbm.ic_type = "taylorgreen"
ic_taylorgreen.rho0 = 1.0
ic_taylorgreen.v0 = 0.1
eb2.geom_type = "all_regular"
""", [0, 1, 10]),
    # config 2: thermal Sod tube (gamma = 2, alpha set, zeroth-order outflow in x)
    "sod48": ("""
max_step = 30
geometry.prob_lo = 0.0 -1.0 -1.0
geometry.prob_hi = 48.0 1.0 1.0
geometry.is_periodic = 0 1 1
amr.n_cell = 48 2 2
lbm.bc_lo = 5 0 0
lbm.bc_hi = 5 0 0
lbm.nu = 0.010
lbm.alpha = 0.010
lbm.ic_type = "sod"
ic_sod.density = 0.50
ic_sod.mach_components = 0.0 0.0 0.0
ic_sod.x_discontinuity = 24.0
ic_sod.initial_temperature = 0.20
ic_sod.adiabatic_exponent = 2.0
ic_sod.mean_molecular_mass = 28.96
ic_sod.density_ratio = 4.00
ic_sod.temperature_ratio = 0.1250
lbm.initial_temperature = 0.20
lbm.adiabatic_exponent = 2.0
lbm.mean_molecular_mass = 28.96
eb2.geom_type = "all_regular"
""", [0, 1, 30]),
    # config 4 at reduced size: channel inlet, outflow, no-slip walls, periodic z, EB cylinder
    "chcyl": ("""
max_step = 8
geometry.prob_lo = 0.0 0.0 0.0
geometry.prob_hi = 32.0 12.0 4.0
geometry.is_periodic = 0 0 1
amr.n_cell = 32 12 4
lbm.bc_lo = 2 1 0
lbm.bc_hi = 5 1 0
lbm.nu = 0.0050
lbm.velocity_bc_type = "channel"
velocity_bc_channel.initial_density = 1.0
velocity_bc_channel.Mach_ref = 0.01
velocity_bc_channel.initial_temperature = 0.03
lbm.ic_type = "constant"
ic_constant.density = 1.0
ic_constant.initial_temperature = 0.03
ic_constant.mach_components = 0.0 0.0 0.0
eb2.geom_type = "cylinder"
eb2.cylinder_radius = 2.6
eb2.cylinder_center = 9.0 6.0 8.0
eb2.cylinder_has_fluid_inside = 0
eb2.cylinder_height = 256.0
eb2.cylinder_direction = 2
""", [0, 1, 8]),
    # Pr != 1, R != 1: exercises omega_one/omega != 1 (MRT heat flux) near the over-relaxation limit
    "thermal": ("""
max_step = 20
geometry.prob_lo = 0.0 0.0 -1.0
geometry.prob_hi = 4.0 24.0 1.0
geometry.is_periodic = 1 1 1
amr.n_cell = 4 24 2
lbm.bc_lo = 0 0 0
lbm.bc_hi = 0 0 0
lbm.nu = 0.000008076
lbm.alpha = 0.000010
lbm.ic_type = "thermaldiffusivity_test"
ic_thermaldiffusivity_test.density = 1.0
ic_thermaldiffusivity_test.mach_components = 0.0 0.0 0.0
ic_thermaldiffusivity_test.wave_length = 24.0
ic_thermaldiffusivity_test.initial_temperature = 0.003333
ic_thermaldiffusivity_test.adiabatic_exponent = 1.32
ic_thermaldiffusivity_test.mean_molecular_mass = 17.031
lbm.initial_temperature = 0.003333
lbm.adiabatic_exponent = 1.32
lbm.mean_molecular_mass = 17.031
eb2.geom_type = "all_regular"
""", [0, 1, 20]),
    # pine_box boundary set on a clean geometry: constant-velocity inlet (z-lo), pressure outlet
    # (z-hi), four no-slip walls, EB sphere; needs the K6 pre-pass at the outlet corners
    "pressure": ("""
max_step = 6
geometry.prob_lo = 0.0 0.0 0.0
geometry.prob_hi = 10.0 10.0 14.0
geometry.is_periodic = 0 0 0
amr.n_cell = 10 10 14
lbm.bc_lo = 1 1 2
lbm.bc_hi = 1 1 3
lbm.nu = 0.01733333333333333
lbm.velocity_bc_type = "constant"
velocity_bc_constant.dir = 2
velocity_bc_constant.Mach_ref = 0.01
lbm.ic_type = "constant"
ic_constant.density = 1.0
ic_constant.mach_components = 0.0 0.0 0.002
eb2.geom_type = "sphere"
eb2.sphere_radius = 2.2
eb2.sphere_center = 5.0 5.0 6.0
eb2.sphere_has_fluid_inside = 0
""", [0, 1, 6]),
    # single_rotated_box boundary set: slip walls (x), parabolic inlet (y-lo) with prob_lo != 0,
    # outflow (y-hi), periodic z, EB box
    "slip": ("""
max_step = 6
geometry.prob_lo = -8.0 0.0 -2.0
geometry.prob_hi = 8.0 24.0 2.0
geometry.is_periodic = 0 0 1
amr.n_cell = 16 24 4
lbm.bc_lo = 6 2 0
lbm.bc_hi = 6 5 0
lbm.nu = 0.20
lbm.velocity_bc_type = "parabolic"
velocity_bc_parabolic.Mach_ref = 0.0025
velocity_bc_parabolic.normal_dir = 0
velocity_bc_parabolic.tangential_dir = 1
lbm.ic_type = "constant"
ic_constant.density = 1.0
ic_constant.mach_components = 0.0 0.0 0.0
eb2.geom_type = "box"
eb2.box_lo = -2.5 8.5 -1000.0
eb2.box_hi = 2.5 13.5 1000.0
eb2.box_has_fluid_inside = 0
""", [0, 1, 6]),
    # viscosity-test shear wave (all periodic, non-default IC kind 2)
    "shear": ("""
max_step = 12
geometry.prob_lo = 0.0 0.0 -1.0
geometry.prob_hi = 4.0 20.0 1.0
geometry.is_periodic = 1 1 1
amr.n_cell = 4 20 2
lbm.bc_lo = 0 0 0
lbm.bc_hi = 0 0 0
lbm.nu = 0.0001
lbm.ic_type = "viscosity_test"
ic_viscosity_test.density = 1.0
ic_viscosity_test.mach_components = 0.0 0.0 0.0
ic_viscosity_test.wave_length = 20.0
ic_viscosity_test.initial_temperature = 0.03333
lbm.initial_temperature = 0.03333
eb2.geom_type = "all_regular"
""", [0, 1, 12]),
    # slip walls with y and z normals (codes 7 and 8, BC.H:119, 138), constant-velocity inlet, outflow, EB sphere
    "slipyz": ("""
max_step = 6
geometry.prob_lo = 0.0 0.0 0.0
geometry.prob_hi = 20.0 10.0 8.0
geometry.is_periodic = 0 0 0
amr.n_cell = 20 10 8
lbm.bc_lo = 2 7 8
lbm.bc_hi = 5 7 8
lbm.nu = 0.02
lbm.velocity_bc_type = "constant"
velocity_bc_constant.dir = 0
velocity_bc_constant.Mach_ref = 0.02
lbm.ic_type = "constant"
ic_constant.density = 1.0
ic_constant.mach_components = 0.005 0.0 0.0
eb2.geom_type = "sphere"
eb2.sphere_radius = 1.9
eb2.sphere_center = 7.0 5.0 4.0
eb2.sphere_has_fluid_inside = 0
""", [0, 1, 6]),
    # a body that crosses the outlet and a no-slip wall: is_fluid of the out-of-domain ghost cells comes from the
    # geometry evaluated beyond the domain (SURVEY A.4 last paragraph, the pine_box situation).  The outlet is the
    # zeroth-order outflow: with the pressure outlet of `pressure` the reference itself blows up within 8 steps when a
    # body sits on the outlet face (FPE trap at step 8), which is no basis for a round-off comparison
    "touch": ("""
max_step = 8
geometry.prob_lo = 0.0 0.0 0.0
geometry.prob_hi = 10.0 10.0 14.0
geometry.is_periodic = 0 0 0
amr.n_cell = 10 10 14
lbm.bc_lo = 1 1 2
lbm.bc_hi = 1 1 5
lbm.nu = 0.01733333333333333
lbm.velocity_bc_type = "constant"
velocity_bc_constant.dir = 2
velocity_bc_constant.Mach_ref = 0.01
lbm.ic_type = "constant"
ic_constant.density = 1.0
ic_constant.mach_components = 0.0 0.0 0.002
eb2.geom_type = "box"
eb2.box_lo = 5.5 3.5 10.5
eb2.box_hi = 12.5 6.5 17.5
eb2.box_has_fluid_inside = 0
""", [0, 1, 8]),
}

KEEP_STEP0 = ([f"f_{q:02d}" for q in range(27)] + [f"g_{q:02d}" for q in range(27)] + O.MACRO_NAMES)



# ---------------------------------------------------------------------------------------------------------
# multi-level cases (BASELINE configs 4-5): several boxes per level, 2 and 3 levels, static tagging boxes.
# Stored per level: the box list the reference made (Level_k/Cell_H), is_fluid, and dense fields (NaN where
# the level has no box).  Fine levels stay away from non-periodic faces, and no body crosses a periodic face
# inside a refined region (there the reference averages uninitialised memory into the coarse level).
# ---------------------------------------------------------------------------------------------------------
AMR_COMMON = """
lbm.dx_outer = 1.0
lbm.dt_outer = 1.0
lbm.save_streaming = 1
lbm.save_derived = 1
amr.plot_int = 1
amr.chk_int = -1
amr.blocking_factor = 4
amr.regrid_int = 1000000
amrex.fpe_trap_invalid = 0
amrex.fpe_trap_zero = 0
amrex.fpe_trap_overflow = 0
amrex.the_arena_is_managed = 0
"""

CHANNEL_BODY = """
geometry.is_periodic = 0 0 1
lbm.bc_lo = 2 1 0
lbm.bc_hi = 5 1 0
lbm.nu = 0.0050
lbm.velocity_bc_type = "channel"
velocity_bc_channel.initial_density = 1.0
velocity_bc_channel.Mach_ref = 0.01
velocity_bc_channel.initial_temperature = 0.03
lbm.ic_type = "constant"
ic_constant.density = 1.0
ic_constant.initial_temperature = 0.03
ic_constant.mach_components = 0.0 0.0 0.0
eb2.geom_type = "cylinder"
eb2.cylinder_radius = 2.3
eb2.cylinder_has_fluid_inside = 0
eb2.cylinder_height = 2.0
eb2.cylinder_direction = 2
"""

AMR_CASES = {
    # periodic Taylor-Green, 2 levels, 8 coarse + 8 fine boxes of different sizes
    "amr2_tg": ("""
max_step = 3
geometry.prob_lo = -1.0 -1.0 -1.0
geometry.prob_hi =  1.0  1.0  1.0
geometry.is_periodic = 1 1 1
amr.n_cell = 16 16 16
lbm.bc_lo = 0 0 0
lbm.bc_hi = 0 0 0
lbm.nu = 0.1733333333333333
lbm.ic_type = "taylorgreen"
ic_taylorgreen.rho0 = 1.0
ic_taylorgreen.v0 = 0.1
eb2.geom_type = "all_regular"
amr.max_level = 1
amr.max_grid_size = 8
amr.n_error_buf = 0
tagging.refinement_indicators = box
tagging.box.in_box_lo = -0.45 -0.45 -0.45
tagging.box.in_box_hi = 0.45 0.3 0.2
""", 2, [0, 1, 3]),
    # BASELINE config 4 at reduced size: channel inlet / outflow / no-slip walls / periodic z, EB cylinder inside
    # the refined region, the fine level spans the periodic direction
    "amr2_chcyl": ("""
max_step = 4
geometry.prob_lo = 0.0 0.0 0.0
geometry.prob_hi = 48.0 16.0 4.0
amr.n_cell = 48 16 4
eb2.cylinder_center = 14.0 8.0 2.0
amr.max_level = 1
amr.max_grid_size = 16
amr.n_error_buf = 0
tagging.refinement_indicators = box
tagging.box.in_box_lo = 8.0 4.0 -1.0
tagging.box.in_box_hi = 22.0 12.0 5.0
""" + CHANNEL_BODY, 2, [0, 1, 4]),
    # BASELINE config 5 at reduced size: 3 levels (no 3-level deck is shipped; built from the nesting of
    # channel_cylinder_amr.inp:39-41), 6 boxes per level
    "amr3_chcyl": ("""
max_step = 3
geometry.prob_lo = 0.0 0.0 0.0
geometry.prob_hi = 48.0 24.0 4.0
amr.n_cell = 48 24 4
eb2.cylinder_center = 14.0 12.0 2.0
amr.max_level = 2
amr.max_grid_size = 16
amr.n_error_buf = 1
tagging.refinement_indicators = a b
tagging.a.in_box_lo = 8.0 7.0 -1.0
tagging.a.in_box_hi = 24.0 17.0 5.0
tagging.a.max_level = 1
tagging.b.in_box_lo = 10.5 9.0 -1.0
tagging.b.in_box_hi = 18.0 15.0 5.0
""" + CHANNEL_BODY, 3, [0, 1, 3]),
    # sod_amr.inp with a static refined band around the discontinuity (thermal path, gamma = 2, outflow in x)
    "amr2_sod": ("""
max_step = 4
geometry.prob_lo = 0.0 -2.0 -2.0
geometry.prob_hi = 64.0 2.0 2.0
geometry.is_periodic = 0 1 1
amr.n_cell = 64 4 4
lbm.bc_lo = 5 0 0
lbm.bc_hi = 5 0 0
lbm.nu = 0.010
lbm.alpha = 0.010
lbm.ic_type = "sod"
ic_sod.density = 0.50
ic_sod.mach_components = 0.0 0.0 0.0
ic_sod.x_discontinuity = 32.0
ic_sod.initial_temperature = 0.20
ic_sod.adiabatic_exponent = 2.0
ic_sod.mean_molecular_mass = 28.96
ic_sod.density_ratio = 4.00
ic_sod.temperature_ratio = 0.1250
lbm.initial_temperature = 0.20
lbm.adiabatic_exponent = 2.0
lbm.mean_molecular_mass = 28.96
eb2.geom_type = "all_regular"
amr.max_level = 1
amr.max_grid_size = 16
amr.n_error_buf = 0
tagging.refinement_indicators = box
tagging.box.in_box_lo = 24.0 -3.0 -3.0
tagging.box.in_box_hi = 40.0 3.0 3.0
""", 2, [0, 1, 4]),
}

# sod_amr.inp's dynamic refinement: regrid every coarse step (amr.regrid_int = 1), cells tagged where the density
# differs from a neighbour's; the fine boxes grow, split and shrink as the waves travel.  The box list of level 1 is
# stored for EVERY step (`boxes_s<N>_l1`): it is the input a regrid hands to RemakeLevel.
AMR_CASES["amr2_sod_regrid"] = (AMR_CASES["amr2_sod"][0].replace("max_step = 4", "max_step = 16")
                                .replace("amr.n_error_buf = 0", "amr.n_error_buf = 1")
                                .split("tagging.refinement_indicators")[0] + """tagging.refinement_indicators = rho
tagging.rho.adjacent_difference_greater = 0.02
tagging.rho.field_name = rho
""", 2, [0, 1, 8, 16])
# A level that appears and vanishes while the run goes on (MakeNewLevelFromCoarse, ClearLevel): tagging boxes with a
# time window (tagging.*.start_time / end_time, Source/LBM.cpp:332-343), regrid every coarse step.  Level 1 does not
# exist in plt00000; it is interpolated from level 0 when the window opens and dropped when it closes.
AMR_CASES["amr2_tg_appear"] = (AMR_CASES["amr2_tg"][0].replace("max_step = 3", "max_step = 9") + """tagging.box.start_time = 2.5
tagging.box.end_time = 6.5
""", 2, [0, 2, 3, 4, 6, 7, 9])
AMR_CASES["amr2_chcyl_appear"] = (AMR_CASES["amr2_chcyl"][0].replace("max_step = 4", "max_step = 8") + """tagging.box.start_time = 1.5
tagging.box.end_time = 5.5
""", 2, [0, 1, 2, 3, 5, 6, 8])
# sod_amr.inp itself at reduced size: besides the density tagging, its `bc` box keeps the outflow face at x-lo refined
# (Tests/test_files/sod_amr/sod_amr.inp:51-57) -- fine boxes TOUCH a non-periodic face; a second box does the same at
# x-hi.  The coarse-fine interfaces stay x-normal planes away from the faces.
AMR_CASES["amr2_sod_bc"] = (AMR_CASES["amr2_sod_regrid"][0].replace("max_step = 16", "max_step = 12")
                            .replace("tagging.refinement_indicators = rho", "tagging.refinement_indicators = rho bc bchi") + """tagging.bc.in_box_lo = -10 -10 -10
tagging.bc.in_box_hi = 6 10 10
tagging.bchi.in_box_lo = 59 -10 -10
tagging.bchi.in_box_hi = 80 10 10
""", 2, [0, 1, 6, 12])
# compute_eb_forces on a hierarchy (lbm.compute_forces = 1; the forces file is stored as `forces`): a uniform flow hits the
# cylinder from the first step.  m_mask of every level is EMPTY in this and in every other run the reference completes:
# LBM::initialize_mask only fills it when the level is (re)made while a finer one exists (Source/LBM.cpp:1264-1276), which
# from scratch never happens (each level is made while it is the finest), and a RemakeLevel of level 1 under an existing
# level 2 -- the one way to get there -- segfaults in the reference (tried: tagging box of level 1 widened at step 3).
# So the coarse cells under a fine level are counted as well, and the levels' sums are added without any weighting.
AMR_CASES["amr3_chcyl_forces"] = (AMR_CASES["amr3_chcyl"][0].replace("max_step = 3", "max_step = 6")
                                  .replace("ic_constant.mach_components = 0.0 0.0 0.0", "ic_constant.mach_components = 0.05 0.0 0.0") + """lbm.compute_forces = 1
lbm.forces_file = forces.txt
""", 3, [0, 3, 6])
AMR_REGRID_INT = {"amr2_sod_regrid": 1, "amr2_tg_appear": 1, "amr2_chcyl_appear": 1, "amr2_sod_bc": 1}
# steps stored with f and g as well (besides the first and the last): the first step of the new level
AMR_FULL_STEPS = {"amr2_tg_appear": [4], "amr2_chcyl_appear": [3]}

AMR_KEEP_LAST = KEEP_STEP0 + ["dQCorrX", "dQCorrY", "dQCorrZ"]
AMR_KEEP_MID = O.MACRO_NAMES


def make_amr(out_dir, only):
    for name, (deck, nlev, steps) in AMR_CASES.items():
        if only and name not in only:
            continue
        work = tempfile.mkdtemp(prefix=f"golden_{name}_")
        deck_text = deck.strip() + "\n" + AMR_COMMON
        if name in AMR_REGRID_INT:
            deck_text = deck_text.replace("amr.regrid_int = 1000000", f"amr.regrid_int = {AMR_REGRID_INT[name]}")
        deck_path = os.path.join(work, "case.inp")
        with open(deck_path, "w") as fh:
            fh.write(deck_text)
        O.run_reference(deck_path, work, [], omp=False)
        data = {"deck": np.array(deck_text), "steps": np.array(steps), "nlev": np.array(nlev)}
        if os.path.exists(os.path.join(work, "forces.txt")):
            data["forces"] = np.loadtxt(os.path.join(work, "forces.txt"), skiprows=1)
        def boxes_of(st, lev):  # empty when the level does not exist in that plotfile
            pdir = os.path.join(work, f"plt{st:05d}")
            if not os.path.isdir(os.path.join(pdir, f"Level_{lev}")):
                return np.zeros((0, 2, 3), dtype=np.int64)
            return np.array(O.read_plotfile_boxes(pdir, lev))
        if name in AMR_REGRID_INT:
            for st in range(1, steps[-1] + 1):
                for lev in range(1, nlev):
                    data[f"boxes_s{st}_l{lev}"] = boxes_of(st, lev)
        for lev in range(nlev):
            data[f"boxes_l{lev}"] = boxes_of(0, lev)
            for s in steps:
                if len(boxes_of(s, lev)) == 0:
                    continue
                pf = O.read_plotfile(os.path.join(work, f"plt{s:05d}"), lev)
                if f"is_fluid_l{lev}" not in data:
                    fl = pf["is_fluid"]
                    data[f"is_fluid_l{lev}"] = np.where(np.isnan(fl), 1, fl).astype(np.int8)
                keep = AMR_KEEP_LAST if s == steps[-1] else (
                    KEEP_STEP0 if s == steps[0] or s in AMR_FULL_STEPS.get(name, []) else AMR_KEEP_MID)
                for n in keep:
                    data[f"s{s}_l{lev}_{n}"] = pf[n]
        if name.endswith("_forces"):  # the forces file is the golden; the fields of this deck's hierarchy are pinned elsewhere
            data = {k: v for k, v in data.items() if not (k[0] == "s" and k[1].isdigit())}
        path = os.path.join(out_dir, f"{name}.npz")
        np.savez_compressed(path, **data)
        print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, boxes per level "
              f"{[len(data[f'boxes_l{l}']) for l in range(nlev)]}")
        shutil.rmtree(work)


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]
    for name, (deck, steps) in CASES.items():
        if only and name not in only:
            continue
        work = tempfile.mkdtemp(prefix=f"golden_{name}_")
        deck_text = deck.strip() + "\n" + COMMON
        deck_path = os.path.join(work, "case.inp")
        with open(deck_path, "w") as fh:
            fh.write(deck_text)
        O.run_reference(deck_path, work, [], omp=False)
        data = {"deck": np.array(deck_text), "steps": np.array(steps)}
        for s in steps:
            pf = O.read_plotfile(os.path.join(work, f"plt{s:05d}"))
            names = pf["__names__"] if s == steps[-1] else KEEP_STEP0
            for n in names:
                data[f"s{s}_{n}"] = pf[n]
            if s == steps[0]:
                data["is_fluid"] = pf["is_fluid"].astype(np.int8)
        path = os.path.join(out_dir, f"{name}.npz")
        np.savez_compressed(path, **data)
        print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, solid cells {int((data['is_fluid'] == 0).sum())}")
        shutil.rmtree(work)


if __name__ == "__main__":
    main()
    make_amr(os.path.dirname(os.path.abspath(__file__)), sys.argv[1:])
