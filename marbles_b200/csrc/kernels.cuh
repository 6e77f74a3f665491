// kernels.cuh -- launch interface between the C ABI (api.cu) and the CUDA kernels
// (kernels.cu).  Every launcher enqueues on `st` and returns the number of kernels
// it launched.
#pragma once
#include "lattice.cuh"

namespace mbl {

struct BcInfo {
    int periodic[3];
    int bc[6];  // idir + 3*lohi
    int vbc_kind, vbc_dir, vbc_normal_dir, vbc_tangential_dir;
    double vbc_u, vbc_rho, vbc_T, vbc_gamma, vbc_R;
    double prob_lo[3], prob_hi[3], dx[3];
};

struct IcInfo {
    int kind;
    double density, vel[3], v0, omega[3], wave_length, T0, gamma, R, c_s, density_ratio, temperature_ratio, x_disc;
};

// analytic embedded-boundary body of a deck (eb2.geom_type = sphere | cylinder | box, Source/EB.cpp:5-38)
struct BodyInfo {
    int kind;          // 1 sphere, 2 cylinder, 3 box
    int axis;          // cylinder direction
    int fluid_inside;  // eb2.*_has_fluid_inside
    double a[3], b[3]; // centre (sphere, cylinder) or box lo; box hi
    double r, h;       // radius; cylinder height (<= 0: unbounded)
};

struct LevelPtrs {
    double* f[2];  // ping-pong lattice buffers
    double* g[2];
    double* qc;     // 3 comps: QCorr_x, QCorr_y, QCorr_z of the post-stream state
    double* macro;  // 19 comps or nullptr
    double* derived;  // 7 comps or nullptr
    uint32_t* nbr;  // 27-bit pull mask
    uint8_t* flag;  // flag byte field
};

void init_tables();

int launch_flags(const Layout& L, const BcInfo& B, const int32_t* d_isfluid_fab, int ng, uint32_t* nbr, uint8_t* flag,
                 cudaStream_t st);
int launch_flags_all_fluid(const Layout& L, const BcInfo& B, uint32_t* nbr, uint8_t* flag, cudaStream_t st);

// is_fluid (FAB layout, box grown by ng) of an analytic body, on the device
int launch_body_is_fluid(const Layout& L, const BcInfo& B, const BodyInfo& G, int32_t* fab, int ng, cudaStream_t st);

int launch_fill(double* p, long long n, double v, cudaStream_t st);

int launch_initialize(const Layout& L, const BcInfo& B, const IcInfo& I, const uint8_t* flag, double* f, double* g,
                      cudaStream_t st);

// ghost fill of the CURRENT buffers; `local_z` = this box spans the whole domain in z
// (periodic z is then wrapped locally, otherwise z ghost planes come from mbl_halo_*)
int launch_ghost_fill(const Layout& L, const BcInfo& B, double* f, double* g, bool local_z, bool do_prepass,
                      bool do_periodic, cudaStream_t st);

// two-pass collide.  pull=true: stream+collide fused, reads (fin,gin) writes (fout,gout);
// pull=false: collide in place on the streamed state.  [ka, kb) restricts the launch to those planes
// (default: every plane the step needs).
int launch_qcorr(const Layout& L, const Phys& P, const double* fin, const double* gin, const uint32_t* nbr, double* qc,
                 bool pull, cudaStream_t st, int ka = 0, int kb = 0);
int launch_collide(const Layout& L, const Phys& P, const double* fin, const double* gin, double* fout, double* gout,
                   const uint32_t* nbr, const uint8_t* flag, const double* qc, double* macro, bool pull,
                   cudaStream_t st, int ka = 0, int kb = 0);
// "carry" step (kernels.cu: k_collide_carry / k_qcorr_combine): the collide kernel also emits partial sums of
// the next post-stream state's conserved moments (12 words per cell, `part`), from which the next step's
// q-corrections are assembled without touching the populations a second time.
struct CarryPlan {
    int own, halo;  // cells a warp owns (own + 2 * halo = 32 lanes) and redundant cells on each side
    int ky;         // rows a thread marches through
    int nxc;        // warps per row
    int prefetch;   // rows of L2 bulk-prefetch distance (0: off)
    unsigned esz8;  // k_collide_tile: bytes per plane of an edge array
};
constexpr int CARRY_WORDS = 12;
constexpr int CARRY_EDGE_WORDS = 18;  // k_collide_tile: [first row down | last row up][c][rho, jx, e2]
// per-component base pointers, passed by value in kernel parameter space
struct CarryPtrs {
    const double* fin[NQ];
    const double* gin[NQ];
    double* fout[NQ];
    double* gout[NQ];
    const double* qc[3];
    double* part[CARRY_WORDS];
    double* edge[CARRY_EDGE_WORDS];
};
// q-correction array of the NEXT step (k_collide_tile_march finishes most cells itself)
struct MarchOut {
    double* qcn[3];
};
CarryPlan make_carry_plan(const Layout& L, int own, int ky);
int launch_collide_carry(const Layout& L, const Phys& P, const CarryPlan& C, int min_blocks, const double* fin,
                         const double* gin, double* fout, double* gout, const uint32_t* nbr, const uint8_t* flag,
                         const double* qc, double* part, cudaStream_t st);
// k_collide<pull> with g staged through shared memory and 32-bit addressing (bit-identical results)
int launch_collide_lean(const Layout& L, const Phys& P, int min_blocks, const double* fin, const double* gin,
                        double* fout, double* gout, const uint32_t* nbr, const uint8_t* flag, const double* qc,
                        cudaStream_t st, int ka = 0, int kb = 0);
// the same without marching: CTA = `rows` warps = rows - 2 owned rows + 2 halo rows of one column strip
int launch_collide_tile(const Layout& L, const Phys& P, const CarryPlan& C, int rows, const double* fin,
                        const double* gin, double* fout, double* gout, const uint32_t* nbr, const uint8_t* flag,
                        const double* qc, double* part, double* edge, cudaStream_t st, int ka = 0, int kb = 0);
// doubles per plane of one edge array for `rows` rows per CTA (the array spans nz + 2 GZ planes)
// the tile carry step with the z sum of plane pairs completed on chip (9 part words instead of 12; even nz)
int launch_collide_tile_pair(const Layout& L, const Phys& P, const CarryPlan& C, const double* fin, const double* gin,
                             double* fout, double* gout, const uint32_t* nbr, const uint8_t* flag, const double* qc,
                             double* part, double* edge, cudaStream_t st, int ka = 0, int kb = 0);
int launch_qcorr_combine_pair(const Layout& L, const Phys& P, const double* fin, const double* gin, const uint32_t* nbr,
                              const double* part, const double* edge, double* qc, cudaStream_t st, int ka = 0, int kb = 0);
// variant 9: the tile carry step marching through z-chunks of zm planes, z sums completed on chip, QCorr of the next step
// written by the collide kernel for the cells that are complete (kernels.cu); zpos[k + GZ]: 0 interior, 1 first, 2 last
// plane of its chunk
// W rows per CTA; pipe: plane k+1 is pulled into shared memory while plane k is collided (variant 10; W = 4 or 8)
bool march_rows_supported(int W, bool pipe);
int launch_collide_tile_march(const Layout& L, const Phys& P, const CarryPlan& C, int W, bool pipe, int zm, const double* fin,
                              const double* gin, double* fout, double* gout, const uint32_t* nbr, const uint8_t* flag,
                              const double* qc, double* qc_next, double* part, double* edge, cudaStream_t st, int ka = 0,
                              int kb = 0);
int launch_qcorr_combine_march(const Layout& L, const Phys& P, const double* fin, const double* gin, const uint32_t* nbr,
                               const double* part, const double* edge, int W, const signed char* zpos, double* qc,
                               cudaStream_t st, int ka = 0, int kb = 0);
long long carry_edge_plane(const Layout& L, int rows);
int carry_tile_rows(int rows);  // supported rows per CTA: 4, 6 (default), 8, 12
// edge / edge_rows: the edge arrays k_collide_tile wrote (nullptr after k_collide_carry)
int launch_qcorr_combine(const Layout& L, const Phys& P, const double* fin, const double* gin, const uint32_t* nbr,
                         const double* part, const double* edge, int edge_rows, double* qc, cudaStream_t st, int ka = 0,
                         int kb = 0);

// march.cu: the whole step as ONE kernel.  A CTA marches through the planes of a z-chunk, every population is
// pulled once (cp.async into shared memory, where it stays for one iteration), the q-corrections of the face
// neighbours are recomputed on a one-cell halo (two extra rows, two extra lanes) instead of being stored.
struct MarchPtrs {
    const double* fin[NQ];
    const double* gin[NQ];
    double* fout[NQ];
    double* gout[NQ];
};
struct MarchPlan {
    int own, halo;  // cells a warp owns and halo lanes on each side (own + 2 halo = 32)
    int nxc;        // column strips per row
    int zm;         // planes per march
    int ka, kb;     // plane range [ka, kb) of the launch
};
MarchPlan make_march_plan(const Layout& L, int zm, int ka, int kb);
size_t march_smem_bytes(int rows);
// rows: productive rows per CTA (4, 6 or 7); zm: planes per march; pipe: pull plane k+2 behind the arithmetic of plane k
int launch_march(const Layout& L, const Phys& P, int rows, int zm, int pipe, const double* fin, const double* gin,
                 double* fout, double* gout, const uint32_t* nbr, const uint8_t* flag, cudaStream_t st, int ka = 0,
                 int kb = 0);

int launch_stream(const Layout& L, const double* fin, const double* gin, double* fout, double* gout,
                  const uint32_t* nbr, cudaStream_t st);
int launch_macrodata(const Layout& L, const Phys& P, const double* f, const double* g, const uint8_t* flag,
                     double* macro, cudaStream_t st);
int launch_derived(const Layout& L, const Phys& P, const uint8_t* flag, const double* macro, double* derived,
                   cudaStream_t st, int with_dq = 0, int zghost = 0);
int launch_eb_forces(const Layout& L, const double* f, const uint8_t* flag, double* d_out3, cudaStream_t st);

// fused.cu: persistent TMA-pipelined kernel.  mode 0: q-correction jobs only (pass 1), 1: collide jobs only
// (pass 2), 2: both interleaved in one launch (the second touch of every population is an L2 hit).
struct FusedPlan {
    int B, NB;       // band height (rows), number of bands
    int kq0, NK;     // first plane with q-correction jobs, number of planes
    int UPR, JPS;    // units per row, job indices per slab = (B + 1) * UPR
    long long NJ;    // job indices per type
    long long LAG;   // collide job n runs after q-correction job n + LAG was handed out
    long long total_tickets;
    int mode;
};
FusedPlan make_fused_plan(const Layout& L, int uw, int band_rows, int mode, int grid, int lag_per_cta);
size_t fused_counter_ints(const Layout& L);
int launch_fused(const Layout& L, const Phys& P, int uw, int band_rows, int lag_per_cta, int mode, int sm_count,
                 const double* fin, const double* gin, double* fout, double* gout, const uint32_t* nbr,
                 const uint8_t* flag, double* qc, double* macro, int* counters, cudaStream_t st);

int launch_fused_plain(const Layout& L, const Phys& P, int band_rows, int lag_per_cta, int mode, int sm_count,
                       const double* fin, const double* gin, double* fout, double* gout, const uint32_t* nbr,
                       const uint8_t* flag, double* qc, double* macro, int* counters, cudaStream_t st);

// lean: only the 27 (of 54) plane-components per lattice a neighbour's pull can reach (kernels.cu: LeanList)
int launch_halo_pack(const Layout& L, const double* f, const double* g, int side, double* buf, cudaStream_t st,
                     bool lean = false);
int launch_halo_unpack(const Layout& L, double* f, double* g, int side, const double* buf, cudaStream_t st,
                       bool lean = false);

}  // namespace mbl
