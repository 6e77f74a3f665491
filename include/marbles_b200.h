/*
 * marbles_b200.h -- C ABI of the B200-native D3Q27 f+g lattice update.
 *
 * This is the drop-in boundary for ONE path of NREL/marbles: the per-timestep
 * lattice update behind lbm::LBM (stream / collide / f_to_macrodata /
 * FillPatchOps::fillpatch / physbc, reference Source/LBM.H:78-112,
 * Source/FillPatchOps.H:15-42).  The reference has no FFI for this path (it calls
 * amrex::ParallelFor lambdas in place); each entry point below names the
 * reference member function it replaces, and INTEGRATION.md shows the patch to
 * Source/LBM.cpp that calls them.
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns 0 on success, nonzero
 *    on error (then mbl_last_error() describes it; the reference-side shim turns
 *    that into amrex::Abort, which is how the reference reports errors);
 *  - work is enqueued on the context's CUDA stream (mbl_set_stream; pass
 *    amrex::Gpu::gpuStream()) and is asynchronous unless the call moves data to
 *    or from HOST memory; mbl_sync() waits;
 *  - "FAB layout" means the reference's array layout for one box grown by `ng`
 *    ghost cells: x fastest, then y, z, component slowest
 *    (Submodules/AMReX/Src/Base/AMReX_Array4.H:60-94);
 *  - there is NO CPU fallback: without a CUDA device every call fails.
 */
#ifndef MARBLES_B200_H
#define MARBLES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBL_NQ 27        /* constants::N_MICRO_STATES, Source/Constants.H:36 */
#define MBL_NMACRO 19    /* constants::N_MACRO_STATES, Source/Constants.H:8  */
#define MBL_NDERIVED 7   /* constants::N_DERIVED,      Source/Constants.H:39 */

/* lbm.bc_lo / lbm.bc_hi codes, Source/BC.H:13-20 */
enum {
    MBL_BC_PERIODIC = 0,
    MBL_BC_NOSLIP = 1,
    MBL_BC_VELOCITY = 2,
    MBL_BC_PRESSURE = 3,
    MBL_BC_OUTFLOW = 5,
    MBL_BC_SLIP_X = 6,
    MBL_BC_SLIP_Y = 7,
    MBL_BC_SLIP_Z = 8
};
/* lbm.velocity_bc_type, Source/LBM.cpp:1382-1451 / Source/VelocityBC.H */
enum { MBL_VBC_NOOP = 0, MBL_VBC_CONSTANT = 1, MBL_VBC_CHANNEL = 2, MBL_VBC_PARABOLIC = 3 };
/* which lattice */
enum { MBL_F = 0, MBL_G = 1 };

/* The parsed scalars of LBM::read_parameters (Source/LBM.cpp:196-300) and of the
 * inlet functor constructors (Source/VelocityBC.cpp:6-54) that the path needs. */
typedef struct mbl_params {
    double nu;          /* lbm.nu                                   */
    double alpha;       /* lbm.alpha (defaults to nu)               */
    double R;           /* m_R_u / m_m_bar                          */
    double gamma;       /* lbm.adiabatic_exponent                   */
    double mesh_speed;  /* lbm.dx_outer / lbm.dt_outer              */
    int bc_type[6];     /* m_bc_type: idir + 3*lohi                 */
    int periodic[3];    /* geometry.is_periodic                     */
    int vbc_kind;       /* MBL_VBC_*                                */
    int vbc_dir;        /* velocity_bc_constant.dir                 */
    int vbc_normal_dir;     /* velocity_bc_parabolic.normal_dir     */
    int vbc_tangential_dir; /* velocity_bc_parabolic.tangential_dir */
    double vbc_u;       /* u0 / u_ref / um = Mach_ref * c_s         */
    double vbc_rho, vbc_T, vbc_gamma, vbc_R; /* functor state       */
} mbl_params;

/* One level of the hierarchy as this rank sees it: the level domain and the ONE
 * box (z-slab of the domain) this rank owns.  Replaces the BoxArray /
 * DistributionMapping / Geometry triple handed to MakeNewLevelFromScratch
 * (Source/LBM.cpp:1148-1199). */
typedef struct mbl_level_geom {
    int dom_lo[3], dom_hi[3]; /* geom[lev].Domain()                 */
    int lo[3], hi[3];         /* valid box owned by this rank       */
    double dt;                /* m_dts[lev]                          */
    double inv_dx[3];         /* geom[lev].InvCellSizeArray()        */
    double prob_lo[3], prob_hi[3], dx[3];
} mbl_level_geom;

/* Device layout of one level's state (structure of arrays, ghost cells
 * included; DESIGN.md "Data layout").  element (q,k,j,i) of a lattice buffer is
 * at  q*comp_stride + (k+gz)*plane_stride + (j+gy)*pitch + (i+ox). */
typedef struct mbl_layout {
    int64_t pitch, plane_stride, comp_stride; /* in doubles          */
    int32_t nx, ny, nz;                        /* valid cells         */
    int32_t ox, gy, gz;                        /* offsets of cell 0   */
    int64_t lattice_doubles;                   /* 27 * comp_stride    */
    int64_t state_bytes;                       /* all device state    */
} mbl_layout;

typedef struct mbl_ctx mbl_ctx;

const char* mbl_last_error(void);
int mbl_version(void);

/* lifetime ------------------------------------------------------------- */
int mbl_create(const mbl_params* params, int device, mbl_ctx** out);
int mbl_destroy(mbl_ctx* ctx);
int mbl_set_stream(mbl_ctx* ctx, void* cuda_stream);
int mbl_sync(mbl_ctx* ctx);

/* level definition: call again after every regrid (RemakeLevel /
 * MakeNewLevelFromCoarse / ClearLevel, Source/LBM.cpp:1088-1379).
 * `device_state` may be NULL (the library allocates) or caller-owned device
 * memory of mbl_level_layout().state_bytes bytes (no ownership transfer). */
int mbl_level_layout(const mbl_level_geom* geom, mbl_layout* out);
int mbl_level_define(mbl_ctx* ctx, int lev, const mbl_level_geom* geom, void* device_state);
int mbl_level_clear(mbl_ctx* ctx, int lev);
/* device pointer of the CURRENT f (which=0) / g (which=1) buffer */
int mbl_level_lattice_ptr(mbl_ctx* ctx, int lev, int which, void** out);

/* EB flag field: replaces LBM::initialize_is_fluid's product m_is_fluid
 * (Source/LBM.cpp:1213-1262).  `is_fluid` is HOST memory, FAB layout, int32,
 * component 0 over the valid box grown by ng (ng >= 2; values beyond the
 * domain come from the geometry).  Builds the 1-byte flag field
 * (bit0 fluid, bit1 eb_boundary, bits2-7 gradient-neighbour usable) and the
 * 27-bit pull mask on the device. */
int mbl_set_is_fluid(mbl_ctx* ctx, int lev, const int32_t* is_fluid, int ng);
int mbl_set_all_fluid(mbl_ctx* ctx, int lev);
/* the same flag field for the analytic bodies of the shipped decks (eb2.geom_type, Source/EB.cpp:5-38), evaluated on
 * the device with EB2's covered-cell rule (all 8 corners inside the body), ghost layers included (periodic images in
 * periodic directions, the geometry beyond the domain elsewhere): no host field, no upload.
 * kind 0 all_regular; 1 sphere {cx, cy, cz, r, fluid_inside}; 2 cylinder {cx, cy, cz, r, height (<= 0: unbounded),
 * direction, fluid_inside}; 3 box {lox, loy, loz, hix, hiy, hiz, fluid_inside}.  STL / general EB2 bodies: pass
 * m_is_fluid through mbl_set_is_fluid. */
int mbl_set_body(mbl_ctx* ctx, int lev, int kind, const double* params, int nparams);

/* state transfer between a FAB (27 comps, ghost width ng, x fastest, component slowest:
 * AMReX_Array4.H:60-94) and the library's padded SoA buffers: one pitched DMA per component.
 * `fab` may point to HOST memory or to DEVICE memory (an AMReX device-arena MultiFab: pass
 * mf[mfi].dataPtr()); upload also takes the FAB's ghost cells the device layout has room for,
 * download writes valid cells only.  These are the calls that carry m_f[lev] / m_g[lev] across
 * the boundary (checkpoint restart, plotfiles, regrid: Source/LBM.cpp:1629-1675, 1772-1782,
 * 1897-1915). */
int mbl_upload(mbl_ctx* ctx, int lev, int which, const double* fab, int ng);
int mbl_download(mbl_ctx* ctx, int lev, int which, double* fab, int ng);
/* macrodata of the last collide/step with want_macrodata (19 comps, ghost ng<=1,
 * valid cells written) and derived data (7 comps, ng 0) */
int mbl_download_macrodata(mbl_ctx* ctx, int lev, double* fab, int ng);
int mbl_download_derived(mbl_ctx* ctx, int lev, double* fab);

/* initial state on the device: ic::Initializer<ICOp>::initialize + fill_f_inside_eb
 * (Source/IC.H:474-519, Source/LBM.cpp:1287-1295).  `ic` = {kind, 16 doubles}, see
 * marbles_b200/lbm.py. */
int mbl_initialize(mbl_ctx* ctx, int lev, int ic_kind, const double* ic_params, int n_ic_params);

/* reference-granular operators (AMR-compatible sequencing stays in the caller) */
/* FillPatchOps::fillpatch(lev,time,mf) for lev 0 / same-level part
 * (Source/FillPatchOps.H:75-132): K6 pre-pass, periodic fill, BCFill pass.
 * Ghost planes owned by other ranks are filled by mbl_halo_* first. */
int mbl_fillpatch(mbl_ctx* ctx, int lev, double time);
/* FillPatchOps::physbc (Source/FillPatchOps.H:142-159) */
int mbl_physbc(mbl_ctx* ctx, int lev, double time);
/* LBM::stream(lev, m_f) and LBM::stream(lev, m_g) (Source/LBM.cpp:558-604), pull form */
int mbl_stream(mbl_ctx* ctx, int lev);
/* LBM::collide(lev) = f_to_macrodata + compute_q_corrections +
 * macrodata_to_equilibrium + relax_f_to_equilibrium (Source/LBM.cpp:607-618) on the
 * already streamed state */
int mbl_collide(mbl_ctx* ctx, int lev, int want_macrodata);
/* LBM::advance(lev) (Source/LBM.cpp:523-544): stream; average_down_to(lev, 1 ghost ring) if level lev+1 is defined;
 * collide.  On the finest multi-box level the stream and the collide are ONE pass over the boxes (results bit-identical
 * to the three calls; MBL_AMR_FUSED=0 keeps them apart). */
int mbl_advance(mbl_ctx* ctx, int lev, int want_macrodata);
/* LBM::f_to_macrodata(lev) (Source/LBM.cpp:810-906) on the current state */
int mbl_f_to_macrodata(mbl_ctx* ctx, int lev);
/* LBM::compute_derived(lev) (Source/LBM.cpp:909-955), needs macrodata */
int mbl_compute_derived(mbl_ctx* ctx, int lev);
/* compute_derived on a z-slab: the vorticity (and, for an initial state, the differenced q-corrections) of the
 * planes next to another rank needs that rank's adjacent macrodata plane.  mbl_macro_halo(pack=1) copies the
 * outermost valid plane of velocity and QCorr (6 comps, mbl_macro_halo_doubles() doubles) on `side` into
 * device_buf, (pack=0) writes a neighbour's buffer into the ghost plane; mbl_compute_derived_slab then
 * differences across the sides that have one. */
int64_t mbl_macro_halo_doubles(mbl_ctx* ctx, int lev);
int mbl_macro_halo(mbl_ctx* ctx, int lev, int side, double* device_buf, int pack);
int mbl_compute_derived_slab(mbl_ctx* ctx, int lev, int has_lo, int has_hi);
/* LBM::compute_eb_forces() for one level (Source/LBM.cpp:994-1044): local sum; on a multi-box level over the boxes that
 * live here.  The caller adds the levels (the reference adds them without weighting) and the ranks. */
int mbl_eb_forces(mbl_ctx* ctx, int lev, double out[3]);

/* ---------------------------------------------------------------------------------------------
 * Multi-box / multi-level levels (AMR-exact mode; BASELINE configs 4-5).
 *
 * mbl_level_define_boxes replaces the BoxArray handed to MakeNewLevelFromScratch / RemakeLevel /
 * MakeNewLevelFromCoarse (Source/LBM.cpp:1088-1199, 1302-1364): `nboxes` valid boxes of level `lev`
 * owned by this rank, lo/hi = 3 ints per box in the level's index space (geom->lo/hi are ignored).
 * Each box is stored as an AMReX-shaped FAB with 3 ghost cells (m_f_nghost, Source/LBM.H:230).  Call it
 * again after every regrid.  On such a level the reference-granular entry points below keep their
 * meaning and follow the reference's un-fused order exactly:
 *   mbl_fillpatch   K6 pre-pass; for lev > 0 FillPatchTwoLevels from level lev-1 (cell_cons_interp into the
 *                   ghost cells no fine valid cell covers, Source/FillPatchOps.H:118-129); FillBoundary; BCFill
 *   mbl_physbc      BCFill only (between the fine substeps, Source/LBM.cpp:503-506)
 *   mbl_stream      stream on the GROWN boxes with the -1 sentinel + FillBoundary (Source/LBM.cpp:558-604)
 *   mbl_collide     f_to_macrodata on valid + 1, FillBoundary, q-corrections, equilibria, relax, FillBoundary
 *   mbl_average_down  average_down_with_ghosts / masked_avgdown of level crse_lev + 1 onto crse_lev
 *                   (Source/Utilities.cpp:5-28, Source/Utilities.H:315-350; ng = 1 inside advance, 0 after init)
 * The sub-cycling order (LBM::time_step, Source/LBM.cpp:452-521) stays in the caller.
 * Restrictions: refinement ratio 2; fine boxes may touch a NON-periodic domain
 * face (sod_amr.inp refines its outflow face), but no coarse-fine INTERFACE cell may have its coarse parent on such a
 * face (mbl_fillpatch reports it).
 * ------------------------------------------------------------------------------------------- */
int mbl_level_define_boxes(mbl_ctx* ctx, int lev, const mbl_level_geom* geom, int nboxes, const int* lo, const int* hi);
int mbl_level_num_boxes(mbl_ctx* ctx, int lev);
/* Distributed levels (AMReX: DistributionMapping, one process per GPU).  Every rank passes the WHOLE box list of the
 * level plus owner[ibox] = the rank that holds the box; only owned boxes get memory here.  The copy-tag lists are
 * built from the global lists on every rank in the same order; a tag whose two boxes live on different ranks becomes a
 * piece of one device message per peer and operator (FabArray::FillBoundary / ParallelCopy,
 * AMReX_FabArrayCommI.H:8-253, AMReX_FabArrayBase.cpp:328-472).  The library packs and unpacks; the caller moves the
 * messages: mbl_set_exchange registers a function that, for npeers peers, sends send[p] (nsend[p] doubles, device
 * memory) to peers[p] and receives nrecv[p] doubles from it into recv[p], ordered after the work already queued on
 * `stream` and complete (or stream-ordered) before it returns -- MPI_Isend/Irecv + Waitall on a CUDA-aware MPI, or
 * ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd.  Every rank calls the same operators in the same order.
 * The _on variants of define / regrid / make_from_coarse take the owner list; the plain ones put every box here. */
typedef int (*mbl_exchange_fn)(void* user, int npeers, const int* peers, double* const* send, const int64_t* nsend,
                               double* const* recv, const int64_t* nrecv, void* stream);
int mbl_set_exchange(mbl_ctx* ctx, int rank, int world, mbl_exchange_fn fn, void* user);
int mbl_level_define_boxes_on(mbl_ctx* ctx, int lev, const mbl_level_geom* geom, int nboxes, const int* lo, const int* hi,
                              const int* owner);
int mbl_level_box_owner(mbl_ctx* ctx, int lev, int ibox);
/* Host only (no device, no context): the cells a FillBoundary over ng ghost cells of a distributed level moves between
 * `rank` and every other rank (send_cells[world], recv_cells[world]) and inside the rank -- the split of the copy-tag
 * list mbl_level_define_boxes_on makes.  What rank a sends to b is what b receives from a. */
int mbl_fill_boundary_plan(int nboxes, const int* lo, const int* hi, const int* owner, const int dom_lo[3], const int dom_hi[3],
                           const int periodic[3], int ng, int rank, int world, int64_t* send_cells, int64_t* recv_cells,
                           int64_t* local_cells);
/* zero-copy: use the DEVICE memory of an AMReX FAB (27 comps, 3 ghost cells: m_f[lev][mfi].dataPtr()) as box
 * `ibox`'s f (which = 0) or g (which = 1); no ownership transfer.  The library keeps the result of every operator
 * in that memory (a scratch copy is used inside mbl_stream, as the reference's f_star). */
int mbl_level_bind(mbl_ctx* ctx, int lev, int ibox, int which, double* device_fab);
/* m_is_fluid[lev][mfi] comp 0 (int32 FAB, ng >= 3 ghost cells already FillBoundary'd, host or device memory) */
int mbl_box_set_is_fluid(mbl_ctx* ctx, int lev, int ibox, const int32_t* fab, int ng);
/* FAB <-> box transfers (27 comps, ghost width ng of the caller's FAB; ghost cells up to 3 are transferred) */
int mbl_box_upload(mbl_ctx* ctx, int lev, int ibox, int which, const double* fab, int ng);
int mbl_box_download(mbl_ctx* ctx, int lev, int ibox, int which, double* fab, int ng);
/* derived = 0: the 19 macrodata comps, 1: the 7 derived comps */
int mbl_box_download_macrodata(mbl_ctx* ctx, int lev, int ibox, double* fab, int ng, int derived);
int mbl_average_down(mbl_ctx* ctx, int crse_lev, int ng);
/* LBM::RemakeLevel (Source/LBM.cpp:1302-1364) after AmrCore::regrid changed the box list of level lev >= 1: the new
 * boxes take the old level's f, g where it covers them (periodic images included) and CellConservativeLinear values
 * from level lev-1 elsewhere, then BCFill -- FillPatchOps::fillpatch into new MultiFabs.  The caller then passes the
 * new is_fluid (mbl_box_set_is_fluid) and calls mbl_fill_f_inside_eb (zero in solid cells + FillBoundary,
 * Source/LBM.cpp:1278-1298, 1347-1348). */
int mbl_level_regrid(mbl_ctx* ctx, int lev, int nboxes, const int* lo, const int* hi);
int mbl_level_regrid_on(mbl_ctx* ctx, int lev, int nboxes, const int* lo, const int* hi, const int* owner);
/* LBM::MakeNewLevelFromCoarse (Source/LBM.cpp:1088-1144): a level lev >= 1 that did not exist appears in a regrid.
 * Every cell of the new boxes (valid and ghost, inside the periodically grown domain) gets CellConservativeLinear
 * values from level lev-1 (FillPatchOps::fillpatch_from_coarse, Source/FillPatchOps.H:164-181), then BCFill.  The
 * caller passes is_fluid afterwards (no fill_f_inside_eb here: the reference does not call it for a new level).
 * LBM::ClearLevel (Source/LBM.cpp:1367-1380), a level that vanishes, is mbl_level_clear. */
int mbl_level_make_from_coarse(mbl_ctx* ctx, int lev, const mbl_level_geom* geom, int nboxes, const int* lo, const int* hi);
int mbl_level_make_from_coarse_on(mbl_ctx* ctx, int lev, const mbl_level_geom* geom, int nboxes, const int* lo, const int* hi,
                                  const int* owner);
int mbl_fill_f_inside_eb(mbl_ctx* ctx, int lev);

/* fused fast path: one coarse step of a single-level run =
 * fillpatch(f), fillpatch(g), stream(f), stream(g), collide
 * (Source/LBM.cpp:416-422, 523-544).  nsteps > 1 only on a single rank; macrodata (all 19
 * fields + differenced q-corrections) is stored for the last step if want_macrodata.  On boxes
 * below ~16 M cells, nsteps >= 5 replays pairs of steps as one CUDA graph (MBL_GRAPH=0 disables). */
int mbl_step(mbl_ctx* ctx, int lev, int nsteps, double time, int want_macrodata);
/* one step of a z-slab whose ghost planes are current: the caller exchanges the halo planes
 * (mbl_halo_pack / transport / mbl_halo_unpack) before every call */
int mbl_step_local(mbl_ctx* ctx, int lev, double time, int want_macrodata);

/* z-halo planes (multi-rank slabs): side 0 = low-z, 1 = high-z.  pack copies
 * the gz outermost VALID planes of f and g (all 27 comps) into `device_buf`
 * (mbl_halo_doubles() doubles); unpack writes a neighbour's packed planes into
 * the ghost planes on that side.  Transport (NCCL over NVLink) is the caller's. */
int64_t mbl_halo_doubles(mbl_ctx* ctx, int lev);
/* lean halo: pack / unpack only the 27 of 54 plane-components per lattice a neighbour's pull can reach (the 18 with
 * e_z towards it or 0 of the adjacent plane, the 9 with e_z towards it of the second plane): half the bytes.  Only
 * for levels without bounce-back or boundary ghost values next to the slab cut (all-fluid, all-periodic). */
int mbl_set_halo_lean(mbl_ctx* ctx, int on);
int mbl_halo_pack(mbl_ctx* ctx, int lev, int side, double* device_buf);
int mbl_halo_unpack(mbl_ctx* ctx, int lev, int side, const double* device_buf);

/* The slab step with the halo exchange overlapped with the interior planes (the cross-rank part of
 * FillBoundary, AMReX_FabArrayCommI.H:8-253, hidden behind the kernels; any boundary conditions).  Output plane k
 * depends on input planes k-2..k+2, so:
 *   mbl_step_split(part 0): ghost fill of the current buffers (levels with non-periodic faces), then the 2 outermost
 *     planes at each z-end (needs the current z ghost planes)
 *   mbl_halo_pack_next / exchange / mbl_halo_unpack_next on ANOTHER stream (mbl_set_stream): the boundary
 *     planes just written travel to the neighbours' ghost planes of the buffers being written
 *   mbl_step_split(part 1): the interior planes; the written buffers become current.
 * The caller orders the streams with events (marbles_b200/lbm.py: LBM._step_overlapped). */
int mbl_step_split(mbl_ctx* ctx, int lev, int part);
int mbl_halo_pack_next(mbl_ctx* ctx, int lev, int side, double* device_buf);
int mbl_halo_unpack_next(mbl_ctx* ctx, int lev, int side, const double* device_buf);

/* host-buffer convenience used for the end-to-end measurement: upload f,g
 * (FAB layout, ghost ng), run nsteps, download f,g into the same buffers */
int mbl_step_host(mbl_ctx* ctx, int lev, int nsteps, double time, double* f_fab, double* g_fab, int ng);

/* the same for a z-slab of a multi-rank run (all-periodic level): begin uploads the two outermost planes at each
 * z-end, the caller exchanges them (mbl_halo_pack / transport / mbl_halo_unpack on the context's stream), finish
 * uploads the interior in chunks with the kernels following the upload frontier and the result going back down */
int mbl_step_host_begin(mbl_ctx* ctx, int lev, double* f_fab, double* g_fab, int ng);
int mbl_step_host_finish(mbl_ctx* ctx, int lev, double* f_fab, double* g_fab, int ng);

/* number of kernels launched by this context so far (bench.py gpu_launches) */
int64_t mbl_launch_count(mbl_ctx* ctx);
/* per-kernel device timing of mbl_step / mbl_step_local with CUDA events on the context's
 * stream: ms[0] ghost fill, ms[1] q-correction pass, ms[2] collide pass, summed over the
 * *nsteps steps recorded since the last call (bench.py roofline) */
int mbl_set_timing(mbl_ctx* ctx, int on);
int mbl_get_timing(mbl_ctx* ctx, double ms[3], int* nsteps);
/* select the implementation of mbl_step: 0 = two kernels, k_qcorr (q-corrections of the
 * post-stream state) then k_collide_lean (pull + collide); 1 = ONE persistent kernel per level with bulk-TMA
 * staged pulls, q-correction jobs and collide jobs interleaved; 2 = the same kernel launched once per job
 * type; 3 = one persistent warp-autonomous kernel with plain loads; 4, 5 = "carry" steps whose collide kernel
 * also emits partial sums of the next step's conserved moments, so the q-correction pass does not read the
 * populations again (4: threads march through rows, 5: one cell per thread, rows exchanged inside the CTA).
 * 6 = 0 with a chosen number of CTAs per SM; 7 = 5 with the z sum of plane pairs completed on chip (9 carried words
 * instead of 12; even nz, else 5); 8 = one z-marching kernel per step; 9 = 5 marching through z-chunks of 8 planes
 * (MBL_ZMARCH): the z sums are completed on chip and the collide kernel itself stores the next step's QCorr for the
 * cells that are complete (3 words instead of 12; needs a second QCorr array).  9 is the default (boxes whose
 * components exceed 4 GB fall back to 0, boxes with fewer than 4 planes to 5).  Variants 1-4 and 8 are measured
 * negative results and are compiled only with MBL_EXPERIMENTS=1.  All variants agree to round-off (the carried moments are summed in another order);
 * DESIGN.md has the measurements. */
int mbl_set_variant(mbl_ctx* ctx, int variant);
int mbl_get_variant(mbl_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MARBLES_B200_H */
