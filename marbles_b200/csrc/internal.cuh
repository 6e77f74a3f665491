// internal.cuh -- state shared by the two translation units behind the C ABI: api.cu (context, single-box levels,
// the fused step) and amr.cu (multi-box / multi-level patch levels).
#pragma once
#include "../../include/marbles_b200.h"

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace mbl {

int fail(const char* fmt, ...);  // records mbl_last_error(), returns 1

#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) return mbl::fail("%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

constexpr int MAX_LEVELS = 16;
constexpr int NMACRO_ALL = MBL_NMACRO + MBL_NDERIVED;  // macro comps followed by derived comps

struct PatchLevel;  // amr.cu

// the per-level physics and boundary-condition records the kernels take, from the parsed scalars of
// LBM::read_parameters (mbl_params) and the level's geometry
inline void level_phys_bc(const mbl_params& pr, const mbl_level_geom& g, Phys& P, BcInfo& B)
{
    P.nu = pr.nu;
    P.alpha = pr.alpha;
    P.R = pr.R;
    P.gamma = pr.gamma;
    P.cv = pr.R / (pr.gamma - 1.0);  // LBM.cpp:640
    P.dt = g.dt;
    P.mesh_speed = pr.mesh_speed;
    for (int d = 0; d < 3; ++d) {
        P.idx[d] = g.inv_dx[d];
        B.periodic[d] = pr.periodic[d];
        B.prob_lo[d] = g.prob_lo[d];
        B.prob_hi[d] = g.prob_hi[d];
        B.dx[d] = g.dx[d];
    }
    for (int n = 0; n < 6; ++n) B.bc[n] = pr.bc_type[n];
    B.vbc_kind = pr.vbc_kind;
    B.vbc_dir = pr.vbc_dir;
    B.vbc_normal_dir = pr.vbc_normal_dir;
    B.vbc_tangential_dir = pr.vbc_tangential_dir;
    B.vbc_u = pr.vbc_u;
    B.vbc_rho = pr.vbc_rho;
    B.vbc_T = pr.vbc_T;
    B.vbc_gamma = pr.vbc_gamma;
    B.vbc_R = pr.vbc_R;
}

struct Level {
    bool defined = false;
    Layout L;
    Phys P;
    BcInfo B;
    mbl_level_geom geom;
    char* base = nullptr;  // device state block
    bool owned = false;
    LevelPtrs p;
    int cur = 0;  // index of the current lattice buffers
    bool local_z = true;
    int32_t* flag_stage = nullptr;
    double* d_red = nullptr;  // 3 doubles for reductions
    double* macro = nullptr;  // lazily allocated (26 comps)
    int* counters = nullptr;  // fused kernel: ticket + per-slab completion counters
    double* part = nullptr;   // carry step: 12 partial-sum words per cell (lazily allocated)
    double* edge = nullptr;   // tile carry step: 18 words per CTA row
    bool dq_from_macro = false;  // macrodata came from mbl_f_to_macrodata: compute_derived also differences QCorr
    int part_pair = 0;        // layout of `part`: 0 twelve words per cell, 1 plane pairs (variant 7), 2 z-march chunks (variant 9)
    double* qc2 = nullptr;    // variant 9: the q-correction array the collide kernel writes for the next step
    signed char* zpos = nullptr;  // variant 9: position of each plane in its z-chunk (0 interior, 1 first, 2 last)
    int zpos_zm = 0;          // chunk length zpos was written for
    int edge_rows = 0;        // rows per CTA the edge arrays were written with (0: written by the marching kernel)
    // two consecutive steps (buffers a -> b -> a) captured as one CUDA graph: small boxes are launch-bound
    // (a non-periodic level issues ~30 ghost-fill launches per step)
    cudaGraphExec_t graph = nullptr;
    int graph_cur = -1, graph_variant = -1;  // buffer parity and step variant the graph was captured for
    const double* graph_qc = nullptr;        // ... and which of the two QCorr arrays was current (variant 9 swaps them)
    int64_t graph_launches = 0;              // kernels per replay
    bool carry_valid = false; // `part` holds the partial sums of the current lattice buffers' next post-stream state
};

}  // namespace mbl

struct mbl_ctx {
    mbl_params prm;
    int device = 0;
    cudaStream_t stream = nullptr;
    mbl::Level lev[mbl::MAX_LEVELS];
    mbl::PatchLevel* plev[mbl::MAX_LEVELS] = {};  // multi-box levels (mbl_level_define_boxes), amr.cu
    // distributed multi-box levels (mbl_set_exchange): this process' rank, the caller's pairwise exchange of device
    // messages, one send and one receive message per peer (grown on demand)
    int rank = 0, world = 1;
    mbl_exchange_fn exchange = nullptr;
    void* exchange_user = nullptr;
    struct PeerBuf {
        double* send = nullptr;
        double* recv = nullptr;
        long long cap_send = 0, cap_recv = 0;
    };
    std::vector<PeerBuf> peer_buf;
    int64_t launches = 0;
    // implementation of mbl_step: 0 (default, fastest measured): two kernels, k_qcorr + k_collide;
    // 1: persistent TMA-pipelined kernel with both job types; 2: the same kernel, one launch per job type;
    // 3: persistent warp-autonomous kernel (plain loads) with both job types.  DESIGN.md has the numbers.
    int variant = 9;
    int uw = 128, band_rows = 16, lag_per_cta = 4;  // variants 1-3 tuning (MBL_UW / MBL_BAND / MBL_LAG)
    int carry_own = 30, carry_ky = 32, carry_minb = 2, carry_rows = 6;  // variant 4 tuning (MBL_OWN / MBL_KY / MBL_MINB)
    int march_rows = 6, march_zm = 64, march_pipe = 1;  // variant 8 tuning (MBL_MROWS / MBL_ZM / MBL_PIPE)
    int zmarch = 8;  // variant 9: planes per z-chunk (MBL_ZMARCH)
    int sm_count = 148;
    cudaStream_t s_up = nullptr, s_down = nullptr;  // copy streams of the pipelined mbl_step_host
    cudaStream_t s_capture = nullptr;               // CUDA graph capture of step pairs
    int host_chunk = 16;  // planes per upload chunk (MBL_HOST_CHUNK; negative: no pipelining)
    bool use_graphs = true;  // MBL_GRAPH=0 disables
    bool halo_lean = false;  // mbl_set_halo_lean: ghost planes carry only the populations a pull can reach
    bool timing = false;
    std::vector<cudaEvent_t> events;  // 4 per timed record: before ghost fill, q-corr, collide, after
    int timed_steps = 0;              // steps covered by the records (a split step makes two records)
};

namespace mbl {
// amr.cu: patch-level counterparts of the per-level entry points (dispatched from api.cu)
int patch_clear(mbl_ctx* ctx, int lev);
int patch_initialize(mbl_ctx* ctx, int lev, const IcInfo& I);
int patch_fillpatch(mbl_ctx* ctx, int lev, double time);
int patch_physbc(mbl_ctx* ctx, int lev, double time);
int patch_stream(mbl_ctx* ctx, int lev);
int patch_collide(mbl_ctx* ctx, int lev, int want_macro);
int patch_advance(mbl_ctx* ctx, int lev, int want_macro);
int patch_eb_forces(mbl_ctx* ctx, int lev, double out[3]);
int patch_f_to_macrodata(mbl_ctx* ctx, int lev);
int patch_compute_derived(mbl_ctx* ctx, int lev);
}  // namespace mbl
