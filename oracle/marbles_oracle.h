/*
 * marbles_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * Plain-C restatement of the reference's per-timestep lattice update
 * (NREL/marbles @ b1b50272) on ONE box with ghost cells, written so that the
 * order of floating-point operations follows the reference and results can be
 * compared bit for bit with the reference executable built by
 * oracle/refbuild/Makefile (parity is PINNED: see tests/test_oracle_golden.py
 * and tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (marbles_b200/) never links, imports or calls it.
 *
 * Array layout is the reference's FAB layout (AMReX_Array4.H:60-94): x fastest,
 * then y, then z, component slowest, over the box grown by `ng` cells.
 */
#ifndef MARBLES_ORACLE_H
#define MARBLES_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NQ 27
#define ORC_NMACRO 19   /* Source/Constants.H:8-31 */
#define ORC_NDERIVED 7  /* Source/Constants.H:39-47 */

/* boundary-condition codes, Source/BC.H:13-20 */
enum {
    ORC_BC_PERIODIC = 0,
    ORC_BC_NOSLIP = 1,
    ORC_BC_VELOCITY = 2,
    ORC_BC_PRESSURE = 3,
    ORC_BC_OUTFLOW = 5,
    ORC_BC_SLIP_X = 6,
    ORC_BC_SLIP_Y = 7,
    ORC_BC_SLIP_Z = 8
};

/* inlet functors, Source/VelocityBC.H */
enum { ORC_VBC_NOOP = 0, ORC_VBC_CONSTANT = 1, ORC_VBC_CHANNEL = 2, ORC_VBC_PARABOLIC = 3 };

/* initial conditions, Source/IC.H */
enum {
    ORC_IC_CONSTANT = 0,
    ORC_IC_TAYLORGREEN = 1,
    ORC_IC_VISCOSITY = 2,
    ORC_IC_THERMALDIFF = 3,
    ORC_IC_SOD = 4
};

typedef struct orc_params {
    int dom_lo[3], dom_hi[3]; /* level domain box, inclusive cell indices   */
    int lo[3], hi[3];         /* valid box of this FAB, inclusive           */
    int ng;                   /* ghost cells of f, g, is_fluid (reference 3)*/
    int periodic[3];          /* geometry.is_periodic                       */
    int bc_type[6];           /* idir + 3*lohi, Source/LBM.cpp:222-225      */
    double nu, alpha;         /* lbm.nu, lbm.alpha                          */
    double R;                 /* R_u / m_bar, Source/LBM.cpp:639            */
    double gamma;             /* lbm.adiabatic_exponent                     */
    double mesh_speed;        /* dx_outer / dt_outer                        */
    double dt;                /* m_dts[lev]                                 */
    double inv_dx[3];         /* geom[lev].InvCellSizeArray()               */
    double prob_lo[3], prob_hi[3], dx[3];
    /* inlet functor (lbm.velocity_bc_type + velocity_bc_<type>.*) */
    int vbc_kind;
    int vbc_dir;            /* constant: dir                                */
    int vbc_normal_dir;     /* parabolic                                    */
    int vbc_tangential_dir; /* parabolic                                    */
    double vbc_u;           /* u0 / u_ref / um = Mach_ref * c_s             */
    double vbc_rho, vbc_T, vbc_gamma, vbc_R;
} orc_params;

typedef struct orc_ic {
    int kind;
    double density;          /* constant/visc/thermal/sod: density; TG: rho0 */
    double velocity[3];      /* already includes mach_components * c_s       */
    double v0;               /* TG                                           */
    double omega[3];         /* TG                                           */
    double wave_length;      /* visc / thermal                               */
    double T0, gamma, R;     /* initial_temperature, adiabatic_exponent, R_u/m_bar */
    double c_s;              /* speed_of_sound_ref                           */
    double density_ratio, temperature_ratio, x_discontinuity; /* sod       */
} orc_ic;

/* sizes */
long orc_ncell_grown(const orc_params* p, int ng);

/* D3Q27 tables (Source/Stencil.H:49-167) */
void orc_stencil(int evs[27][3], double w[27], int bounce[27], int bx[27], int by[27], int bz[27]);
int orc_check_stencil(void); /* Source/Stencil.cpp:5-62; 0 = ok */

/* Source/IC.H:474-519 + LBM.cpp:1287-1295 (zero in solid); no FillBoundary */
void orc_initialize(const orc_params* p, const orc_ic* ic, const int* is_fluid, double* f, double* g);

/* ghost fill: FillPatchOps.H:75-132 for lev 0: K6 pre-pass, periodic FillBoundary
 * (only images that lie in this box), then the BCFill pass (BC.H:345-471) in the
 * region order of AMReX_PhysBCFunct.H:406-682 (CPU branch). */
void orc_prepass(const orc_params* p, double* f);
void orc_fill_periodic(const orc_params* p, double* a, int ncomp, int ng);
void orc_fill_periodic_int(const orc_params* p, int* a, int ncomp, int ng);
void orc_physbc(const orc_params* p, double* f, int is_energy_lattice, double time);
void orc_fillpatch(const orc_params* p, double* f, int is_energy_lattice, double time);

/* LBM.cpp:558-604; `fill_boundary` = run the periodic FillBoundary at the end */
void orc_stream(const orc_params* p, const int* is_fluid, double* f, int fill_boundary);

/* LBM.cpp:810-906 (macro has 1 ghost cell), 959-991, 621-762, 765-807 */
void orc_f_to_macrodata(const orc_params* p, const int* is_fluid, const double* f, const double* g,
                        double* macro, int fill_boundary);
void orc_compute_q_corrections(const orc_params* p, const int* is_fluid, const double* macro, double* derived);
void orc_compute_derived(const orc_params* p, const int* is_fluid, const double* macro, double* derived);
void orc_macrodata_to_equilibrium(const orc_params* p, const int* is_fluid, const double* macro,
                                  const double* derived, double* eq, double* eq_g);
void orc_relax(const orc_params* p, const int* is_fluid, const double* macro, const double* eq,
               const double* eq_g, double* f, double* g, int fill_boundary);
void orc_collide(const orc_params* p, const int* is_fluid, double* f, double* g, double* macro, double* derived,
                 double* eq, double* eq_g, int fill_boundary);

/* is_fluid comp 1 ("eb_boundary"), LBM.cpp:1236-1261; comp 0 must be filled incl. ghosts */
void orc_eb_boundary(const orc_params* p, int* is_fluid);

/* LBM.cpp:994-1044 (single level, mask==0) */
void orc_eb_forces(const orc_params* p, const int* is_fluid, const double* f, double forces[3]);

/* one coarse step of a single-level run, LBM.cpp:416-422 + 523-544 */
void orc_step(const orc_params* p, const int* is_fluid, double* f, double* g, double* macro, double* derived,
              double* eq, double* eq_g, double time);

#ifdef __cplusplus
}
#endif
#endif
