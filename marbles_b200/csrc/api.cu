// api.cu -- C ABI (include/marbles_b200.h) over the CUDA kernels: context, levels,
// device state ownership, host<->device transfer and the per-step sequencing that
// replaces LBM::advance / FillPatchOps::fillpatch (Source/LBM.cpp:523-544,
// Source/FillPatchOps.H:75-132).  No CPU fallback: every entry point needs a device.
#include <cstdlib>
#include <algorithm>
#include <cstring>

#include "internal.cuh"

using namespace mbl;

namespace mbl {

thread_local std::string g_err;

int fail(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

}  // namespace mbl

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct StateMap {
    size_t f0, g0, f1, g1, qc, nbr, flag, total;
};

StateMap state_map(const Layout& L)
{
    StateMap m;
    const size_t lat = (size_t)NQ * L.sq * sizeof(double);
    size_t o = 0;
    m.f0 = o, o += align_up(lat, 256);
    m.g0 = o, o += align_up(lat, 256);
    m.f1 = o, o += align_up(lat, 256);
    m.g1 = o, o += align_up(lat, 256);
    m.qc = o, o += align_up((size_t)3 * L.sq * sizeof(double), 256);
    m.nbr = o, o += align_up((size_t)L.sq * sizeof(uint32_t), 256);
    m.flag = o, o += align_up((size_t)L.sq, 256);
    m.total = o;
    return m;
}

Layout layout_of(const mbl_level_geom* g) { return make_layout(g->lo, g->hi, g->dom_lo, g->dom_hi); }

int check_level(mbl_ctx* ctx, int lev)
{
    if (!ctx) return fail("null context");
    if (lev < 0 || lev >= MAX_LEVELS || !ctx->lev[lev].defined) return fail("level %d is not defined", lev);
    return 0;
}

bool is_patch(mbl_ctx* ctx, int lev) { return ctx && lev >= 0 && lev < MAX_LEVELS && ctx->plev[lev] != nullptr; }

int ensure_macro(Level& lv, cudaStream_t st, int64_t& launches)
{
    if (lv.macro) return 0;
    const size_t n = (size_t)NMACRO_ALL * lv.L.sq;
    CU(cudaMalloc(&lv.macro, n * sizeof(double)));
    launches += launch_fill(lv.macro, (long long)n, 0.0, st);  // m_macrodata.setVal(0), LBM.cpp:1187-1190
    return 0;
}

double* curf(Level& lv) { return lv.p.f[lv.cur]; }
double* curg(Level& lv) { return lv.p.g[lv.cur]; }

// one fused step on the local box; z ghost planes owned by other ranks must be current
int step_local(mbl_ctx* ctx, Level& lv, double /*time*/, int want_macro)
{
    cudaStream_t st = ctx->stream;
    if (want_macro && ensure_macro(lv, st, ctx->launches)) return 1;
    if (want_macro) lv.dq_from_macro = false;  // the collide kernel stores the differenced q-corrections itself
    const int a = lv.cur, b = 1 - lv.cur;
    auto mark = [&]() {
        if (!ctx->timing) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ctx->events.push_back(e);
    };
    mark();
    // all-periodic level: the plain kernels wrap source indices themselves and no ghost cell is read;
    // otherwise (and for the TMA kernels, which stage un-wrapped rows) fill every ghost cell first
    const bool tma = ctx->variant == 1 || ctx->variant == 2;  // experiments only
    Layout Lk = lv.L;
    if (tma) Lk.wrap[0] = Lk.wrap[1] = Lk.wrap[2] = 0;
    const bool need_ghosts = !(Lk.wrap[0] && Lk.wrap[1]);
    if (need_ghosts) ctx->launches += launch_ghost_fill(lv.L, lv.B, lv.p.f[a], lv.p.g[a], lv.local_z, true, true, st);
    mark();
    double* macro = want_macro ? lv.macro : nullptr;
    // the carry kernels address a component with 32-bit byte offsets: larger boxes take the two-kernel step
    int variant = ctx->variant;
    if (variant == 7 && ((lv.L.nz & 1) || carry_tile_rows(ctx->carry_rows) != 6)) variant = 5;  // pairs need an even nz
    if (variant == 9 && (lv.L.nz < 4 || carry_tile_rows(ctx->carry_rows) != 6)) variant = 5;    // z-march needs chunks of >= 2 planes
    if (variant == 10 && lv.L.nz < 4) variant = 5;
    const bool zmarch = variant == 9 || variant == 10;
    if ((variant == 4 || variant == 5 || variant == 7 || variant == 8 || zmarch) && lv.L.sq * 8 >= (1LL << 32)) variant = 0;
#ifdef MBL_EXPERIMENTS
    if (variant == 8 && !macro) {
        // march step: ONE kernel, no q-correction pass and no carried sums (experiments/march.cu)
        mark();
        const int nl = launch_march(Lk, lv.P, ctx->march_rows, ctx->march_zm, ctx->march_pipe, lv.p.f[a], lv.p.g[a],
                                    lv.p.f[b], lv.p.g[b], lv.p.nbr, lv.p.flag, st);
        if (nl < 0) return fail("march step: launch failed (%d)", nl);
        ctx->launches += nl;
        lv.carry_valid = false;
    } else
#endif
    if (variant == 5 || variant == 7 || variant == 4 || variant == 9 || variant == 10) {
        // carry step: q-corrections from the partial sums the previous collide left behind (first step, or after
        // anything else wrote the lattice: the full q-correction pass)
        if (!lv.part) CU(cudaMalloc(&lv.part, (size_t)CARRY_WORDS * lv.L.sq * sizeof(double)));
        const int W = variant == 10 ? (ctx->carry_rows == 4 ? 4 : 8) : carry_tile_rows(ctx->carry_rows);
        if (variant != 4 && !lv.edge)
            CU(cudaMalloc(&lv.edge, (size_t)CARRY_EDGE_WORDS * carry_edge_plane(lv.L, 4) * (lv.L.nz + 2 * GZ) * sizeof(double)));
        if (variant == 9 || variant == 10) {
            if (!lv.qc2) {
                CU(cudaMalloc(&lv.qc2, (size_t)3 * lv.L.sq * sizeof(double)));
                CU(cudaMemsetAsync(lv.qc2, 0, (size_t)3 * lv.L.sq * sizeof(double), st));
            }
            if (!lv.zpos) CU(cudaMalloc(&lv.zpos, (size_t)(lv.L.nz + 2 * GZ)));
        }
        if (lv.carry_valid && lv.part_pair == 2)
            ctx->launches += launch_qcorr_combine_march(Lk, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.part, lv.edge, lv.edge_rows,
                                                        lv.zpos, lv.p.qc, st);
        else if (lv.carry_valid && lv.part_pair)
            ctx->launches += launch_qcorr_combine_pair(Lk, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.part, lv.edge, lv.p.qc, st);
        else if (lv.carry_valid)
            ctx->launches += launch_qcorr_combine(Lk, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.part,
                                                  lv.edge_rows ? lv.edge : nullptr, lv.edge_rows, lv.p.qc, st);
        else
            ctx->launches += launch_qcorr(Lk, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.p.qc, true, st);
        mark();
        if (macro) {
            ctx->launches += launch_collide(Lk, lv.P, lv.p.f[a], lv.p.g[a], lv.p.f[b], lv.p.g[b], lv.p.nbr, lv.p.flag,
                                            lv.p.qc, macro, true, st);
            lv.carry_valid = false;
        } else {
            const CarryPlan C = make_carry_plan(Lk, ctx->carry_own, ctx->carry_ky);
            int nl;
            if (variant == 9 || variant == 10) {
                // chunks of zm planes, none of a single plane; zpos says where each plane sits in its chunk
                int zm = ctx->zmarch > 1 ? ctx->zmarch : 8;
                while (lv.L.nz % zm == 1) ++zm;
                if (lv.zpos_zm != zm) {
                    std::vector<signed char> zp((size_t)lv.L.nz + 2 * GZ, 0);
                    for (int k = 0; k < lv.L.nz; ++k) {
                        const int kk = k % zm, nk = std::min(zm, lv.L.nz - (k - kk));
                        zp[k + GZ] = kk == 0 ? 1 : kk == nk - 1 ? 2 : 0;
                    }
                    CU(cudaMemcpyAsync(lv.zpos, zp.data(), zp.size(), cudaMemcpyHostToDevice, st));
                    CU(cudaStreamSynchronize(st));
                    lv.zpos_zm = zm;
                }
                nl = launch_collide_tile_march(Lk, lv.P, C, W, variant == 10, zm, lv.p.f[a], lv.p.g[a], lv.p.f[b], lv.p.g[b],
                                               lv.p.nbr, lv.p.flag, lv.p.qc, lv.qc2, lv.part, lv.edge, st);
                if (nl > 0) std::swap(lv.p.qc, lv.qc2);  // the cells the kernel finished are in what is now lv.p.qc
            } else if (variant == 7) {
                nl = launch_collide_tile_pair(Lk, lv.P, C, lv.p.f[a], lv.p.g[a], lv.p.f[b], lv.p.g[b], lv.p.nbr, lv.p.flag,
                                              lv.p.qc, lv.part, lv.edge, st);
            } else if (variant == 4) {
#ifdef MBL_EXPERIMENTS
                nl = launch_collide_carry(Lk, lv.P, C, ctx->carry_minb, lv.p.f[a], lv.p.g[a], lv.p.f[b], lv.p.g[b],
                                          lv.p.nbr, lv.p.flag, lv.p.qc, lv.part, st);
#else
                nl = -1;
#endif
            } else {
                nl = launch_collide_tile(Lk, lv.P, C, W, lv.p.f[a], lv.p.g[a], lv.p.f[b], lv.p.g[b], lv.p.nbr, lv.p.flag,
                                         lv.p.qc, lv.part, lv.edge, st);
            }
            lv.edge_rows = variant == 4 ? 0 : W;
            lv.part_pair = variant == 7 ? 1 : (variant == 9 || variant == 10) ? 2 : 0;
            if (nl < 0) return fail("carry step: launch failed (%d)", nl);
            ctx->launches += nl;
            lv.carry_valid = true;
        }
    } else if (variant == 6 && !macro) {
        // the lean collide with a chosen number of CTAs per SM (MBL_MINB); variant 0 uses it with 3
        ctx->launches += launch_qcorr(Lk, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.p.qc, true, st);
        mark();
        const int nl = launch_collide_lean(Lk, lv.P, ctx->carry_minb, lv.p.f[a], lv.p.g[a], lv.p.f[b], lv.p.g[b],
                                           lv.p.nbr, lv.p.flag, lv.p.qc, st);
        if (nl < 0) return fail("lean collide: a lattice component exceeds 4 GB (32-bit byte offsets)");
        ctx->launches += nl;
    } else if (variant == 0 || variant == 6 || variant == 8) {
        ctx->launches += launch_qcorr(Lk, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.p.qc, true, st);
        mark();
        ctx->launches += launch_collide(Lk, lv.P, lv.p.f[a], lv.p.g[a], lv.p.f[b], lv.p.g[b], lv.p.nbr, lv.p.flag,
                                        lv.p.qc, macro, true, st);
        lv.carry_valid = false;
    }
#ifdef MBL_EXPERIMENTS
    else if (variant == 2) {
        ctx->launches += launch_fused(Lk, lv.P, ctx->uw, ctx->band_rows, ctx->lag_per_cta, 0, ctx->sm_count, lv.p.f[a],
                                      lv.p.g[a], lv.p.f[b], lv.p.g[b], lv.p.nbr, lv.p.flag, lv.p.qc, macro, lv.counters, st);
        mark();
        ctx->launches += launch_fused(Lk, lv.P, ctx->uw, ctx->band_rows, ctx->lag_per_cta, 1, ctx->sm_count, lv.p.f[a],
                                      lv.p.g[a], lv.p.f[b], lv.p.g[b], lv.p.nbr, lv.p.flag, lv.p.qc, macro, lv.counters, st);
    } else if (variant == 3) {
        mark();  // no separate q-correction pass
        ctx->launches += launch_fused_plain(Lk, lv.P, ctx->band_rows, ctx->lag_per_cta, 2, ctx->sm_count, lv.p.f[a],
                                            lv.p.g[a], lv.p.f[b], lv.p.g[b], lv.p.nbr, lv.p.flag, lv.p.qc, macro,
                                            lv.counters, st);
    } else {
        mark();  // no separate q-correction pass
        ctx->launches += launch_fused(Lk, lv.P, ctx->uw, ctx->band_rows, ctx->lag_per_cta, 2, ctx->sm_count, lv.p.f[a],
                                      lv.p.g[a], lv.p.f[b], lv.p.g[b], lv.p.nbr, lv.p.flag, lv.p.qc, macro, lv.counters, st);
    }
#else
    else {
        return fail("step variant %d is an experiment: rebuild the library with MBL_EXPERIMENTS=1", variant);
    }
#endif
    mark();
    if (ctx->timing) ctx->timed_steps++;
    lv.cur = b;
    CU(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" {

const char* mbl_last_error(void) { return g_err.c_str(); }
int mbl_version(void) { return 100; }

int mbl_create(const mbl_params* params, int device, mbl_ctx** out)
{
    if (!params || !out) return fail("mbl_create: null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail("mbl_create: no CUDA device (%s); marbles_b200 has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail("mbl_create: device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    for (int d = 0; d < 3; ++d) {
        // LBM::read_parameters consistency checks (LBM.cpp:227-252)
        if (params->periodic[d] && (params->bc_type[d] != 0 || params->bc_type[d + 3] != 0))
            return fail("BC is periodic in direction %d but bc_lo/bc_hi is not 0", d);
        if (!params->periodic[d] && (params->bc_type[d] == 0 || params->bc_type[d + 3] == 0))
            return fail("BC is interior in direction %d but not periodic", d);
    }
    int sm_count = 148;
    CU(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    init_tables();
    CU(cudaGetLastError());
    mbl_ctx* c = new mbl_ctx();  // nothing below can fail: the context is never leaked
    c->prm = *params;
    c->device = device;
    c->sm_count = sm_count;
    if (const char* e = getenv("MBL_VARIANT")) c->variant = atoi(e);
    if (const char* e = getenv("MBL_UW")) c->uw = atoi(e) == 256 ? 256 : 128;
    if (const char* e = getenv("MBL_BAND")) c->band_rows = atoi(e) > 0 ? atoi(e) : 16;
    if (const char* e = getenv("MBL_LAG")) c->lag_per_cta = atoi(e) > 0 ? atoi(e) : 4;
    if (const char* e = getenv("MBL_ROWS")) c->carry_rows = atoi(e);
    if (const char* e = getenv("MBL_HOST_CHUNK")) c->host_chunk = atoi(e);
    if (const char* e = getenv("MBL_MROWS")) c->march_rows = atoi(e) == 4 ? 4 : 6;
    if (const char* e = getenv("MBL_ZM")) c->march_zm = atoi(e) > 0 ? atoi(e) : 64;
    if (const char* e = getenv("MBL_PIPE")) c->march_pipe = atoi(e) != 0;
    if (const char* e = getenv("MBL_GRAPH")) c->use_graphs = atoi(e) != 0;
    if (const char* e = getenv("MBL_OWN")) c->carry_own = atoi(e) == 28 ? 28 : 30;
    if (const char* e = getenv("MBL_KY")) c->carry_ky = atoi(e) > 0 ? atoi(e) : 32;
    if (const char* e = getenv("MBL_ZMARCH")) c->zmarch = atoi(e) > 1 ? atoi(e) : 8;
    if (const char* e = getenv("MBL_MINB")) c->carry_minb = atoi(e) >= 2 && atoi(e) <= 5 ? atoi(e) : 2;
    *out = c;
    return 0;
}

int mbl_destroy(mbl_ctx* ctx)
{
    if (ctx)
        for (auto& b : ctx->peer_buf) {
            if (b.send) cudaFree(b.send);
            if (b.recv) cudaFree(b.recv);
        }
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    for (int l = 0; l < MAX_LEVELS; ++l) mbl_level_clear(ctx, l);
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    if (ctx->s_up) cudaStreamDestroy(ctx->s_up);
    if (ctx->s_down) cudaStreamDestroy(ctx->s_down);
    if (ctx->s_capture) cudaStreamDestroy(ctx->s_capture);
    delete ctx;
    return 0;
}

int mbl_set_stream(mbl_ctx* ctx, void* s)
{
    if (!ctx) return fail("null context");
    ctx->stream = (cudaStream_t)s;
    return 0;
}

int mbl_sync(mbl_ctx* ctx)
{
    if (!ctx) return fail("null context");
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mbl_level_layout(const mbl_level_geom* geom, mbl_layout* out)
{
    if (!geom || !out) return fail("null argument");
    const Layout L = layout_of(geom);
    out->pitch = L.px;
    out->plane_stride = L.sz;
    out->comp_stride = L.sq;
    out->nx = L.nx, out->ny = L.ny, out->nz = L.nz;
    out->ox = OX, out->gy = GY, out->gz = GZ;
    out->lattice_doubles = (int64_t)NQ * L.sq;
    out->state_bytes = (int64_t)state_map(L).total;
    return 0;
}

int mbl_level_clear(mbl_ctx* ctx, int lev)
{
    if (!ctx || lev < 0 || lev >= MAX_LEVELS) return fail("bad level");
    Level& lv = ctx->lev[lev];
    cudaSetDevice(ctx->device);
    if (ctx->plev[lev]) patch_clear(ctx, lev);
    if (!lv.defined) return 0;
    cudaStreamSynchronize(ctx->stream);
    if (lv.owned && lv.base) cudaFree(lv.base);
    if (lv.flag_stage) cudaFree(lv.flag_stage);
    if (lv.d_red) cudaFree(lv.d_red);
    if (lv.macro) cudaFree(lv.macro);
    if (lv.counters) cudaFree(lv.counters);
    if (lv.graph) cudaGraphExecDestroy(lv.graph);
    if (lv.part) cudaFree(lv.part);
    if (lv.edge) cudaFree(lv.edge);
    if (lv.zpos) cudaFree(lv.zpos);
    if (lv.qc2) {
        // the two q-correction arrays may have changed places: free the one that is not part of the state block
        char* q = (char*)lv.p.qc;
        const bool qc_in_block = q >= lv.base && q < lv.base + state_map(lv.L).total;
        cudaFree(qc_in_block ? lv.qc2 : lv.p.qc);
    }
    lv = Level();
    return 0;
}

int mbl_level_define(mbl_ctx* ctx, int lev, const mbl_level_geom* g, void* device_state)
{
    if (!ctx || !g) return fail("null argument");
    if (lev < 0 || lev >= MAX_LEVELS) return fail("level %d out of range", lev);
    for (int d = 0; d < 3; ++d) {
        if (g->hi[d] < g->lo[d]) return fail("empty box in direction %d", d);
        if (g->lo[d] < g->dom_lo[d] || g->hi[d] > g->dom_hi[d]) return fail("box leaves the domain in direction %d", d);
    }
    if (g->lo[0] != g->dom_lo[0] || g->hi[0] != g->dom_hi[0] || g->lo[1] != g->dom_lo[1] || g->hi[1] != g->dom_hi[1])
        return fail("a rank's box must span the level domain in x and y (z-slab decomposition)");
    CU(cudaSetDevice(ctx->device));
    mbl_level_clear(ctx, lev);
    Level& lv = ctx->lev[lev];
    lv.geom = *g;
    lv.L = layout_of(g);
    lv.local_z = (g->lo[2] == g->dom_lo[2] && g->hi[2] == g->dom_hi[2]);
    const bool all_periodic = ctx->prm.periodic[0] && ctx->prm.periodic[1] && ctx->prm.periodic[2];
    lv.L.wrap[0] = lv.L.wrap[1] = all_periodic;
    lv.L.wrap[2] = all_periodic && lv.local_z;
    if (!lv.local_z && lv.L.nz < GZ) return fail("a z-slab needs at least %d planes", GZ);
    level_phys_bc(ctx->prm, *g, lv.P, lv.B);

    const StateMap m = state_map(lv.L);
    if (device_state) {
        lv.base = (char*)device_state;
        lv.owned = false;
    } else {
        CU(cudaMalloc(&lv.base, m.total));
        lv.owned = true;
    }
    lv.p.f[0] = (double*)(lv.base + m.f0);
    lv.p.g[0] = (double*)(lv.base + m.g0);
    lv.p.f[1] = (double*)(lv.base + m.f1);
    lv.p.g[1] = (double*)(lv.base + m.g1);
    lv.p.qc = (double*)(lv.base + m.qc);
    lv.p.nbr = (uint32_t*)(lv.base + m.nbr);
    lv.p.flag = (uint8_t*)(lv.base + m.flag);
    lv.cur = 0;
    CU(cudaMalloc(&lv.d_red, 8 * sizeof(double)));
#ifdef MBL_EXPERIMENTS
    CU(cudaMalloc(&lv.counters, fused_counter_ints(lv.L) * sizeof(int)));
#endif
    // zero everything once: pad cells are never written by the kernels
    CU(cudaMemsetAsync(lv.base, 0, m.total, ctx->stream));
    lv.defined = true;
    ctx->launches += launch_flags_all_fluid(lv.L, lv.B, lv.p.nbr, lv.p.flag, ctx->stream);
    CU(cudaGetLastError());
    return 0;
}

int mbl_level_lattice_ptr(mbl_ctx* ctx, int lev, int which, void** out)
{
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    *out = which == MBL_G ? (void*)curg(lv) : (void*)curf(lv);
    return 0;
}

int mbl_set_all_fluid(mbl_ctx* ctx, int lev)
{
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    lv.carry_valid = false;
    CU(cudaSetDevice(ctx->device));
    ctx->launches += launch_flags_all_fluid(lv.L, lv.B, lv.p.nbr, lv.p.flag, ctx->stream);
    CU(cudaGetLastError());
    return 0;
}

int mbl_set_is_fluid(mbl_ctx* ctx, int lev, const int32_t* is_fluid, int ng)
{
    if (check_level(ctx, lev)) return 1;
    if (!is_fluid) return fail("null is_fluid");
    if (ng < 2) return fail("is_fluid needs at least 2 ghost cells (got %d)", ng);
    Level& lv = ctx->lev[lev];
    lv.carry_valid = false;
    CU(cudaSetDevice(ctx->device));
    const size_t n = (size_t)(lv.L.nx + 2 * ng) * (lv.L.ny + 2 * ng) * (lv.L.nz + 2 * ng);
    if (lv.flag_stage) cudaFree(lv.flag_stage);
    CU(cudaMalloc(&lv.flag_stage, n * sizeof(int32_t)));
    CU(cudaMemcpyAsync(lv.flag_stage, is_fluid, n * sizeof(int32_t), cudaMemcpyDefault, ctx->stream));
    ctx->launches += launch_flags(lv.L, lv.B, lv.flag_stage, ng, lv.p.nbr, lv.p.flag, ctx->stream);
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(lv.flag_stage);
    lv.flag_stage = nullptr;
    CU(cudaGetLastError());
    return 0;
}

// LBM::initialize_is_fluid for the analytic bodies of the shipped decks, evaluated on the device (no host field, no
// upload): kind 1 sphere {cx, cy, cz, r, fluid_inside}, 2 cylinder {cx, cy, cz, r, height, direction, fluid_inside},
// 3 box {lox, loy, loz, hix, hiy, hiz, fluid_inside}; 0 = all_regular
int mbl_set_body(mbl_ctx* ctx, int lev, int kind, const double* v, int nv)
{
    if (check_level(ctx, lev)) return 1;
    if (kind == 0) return mbl_set_all_fluid(ctx, lev);
    if (kind < 1 || kind > 3 || !v) return fail("mbl_set_body: unknown body kind %d", kind);
    if (nv < (kind == 1 ? 5 : 7)) return fail("mbl_set_body: too few parameters for body kind %d", kind);
    Level& lv = ctx->lev[lev];
    lv.carry_valid = false;
    CU(cudaSetDevice(ctx->device));
    BodyInfo G;
    memset(&G, 0, sizeof(G));
    G.kind = kind;
    if (kind == 1) {
        G.a[0] = v[0], G.a[1] = v[1], G.a[2] = v[2], G.r = v[3], G.fluid_inside = v[4] != 0.0;
    } else if (kind == 2) {
        G.a[0] = v[0], G.a[1] = v[1], G.a[2] = v[2], G.r = v[3], G.h = v[4], G.axis = (int)v[5], G.fluid_inside = v[6] != 0.0;
        if (G.axis < 0 || G.axis > 2) return fail("mbl_set_body: cylinder direction %d", G.axis);
    } else {
        for (int d = 0; d < 3; ++d) G.a[d] = v[d], G.b[d] = v[3 + d];
        G.fluid_inside = v[6] != 0.0;
    }
    const int ng = 3;  // m_is_fluid's ghost width (Source/LBM.cpp:1166)
    const size_t n = (size_t)(lv.L.nx + 2 * ng) * (lv.L.ny + 2 * ng) * (lv.L.nz + 2 * ng);
    int32_t* stage = nullptr;
    CU(cudaMalloc(&stage, n * sizeof(int32_t)));
    ctx->launches += launch_body_is_fluid(lv.L, lv.B, G, stage, ng, ctx->stream);
    ctx->launches += launch_flags(lv.L, lv.B, stage, ng, lv.p.nbr, lv.p.flag, ctx->stream);
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(stage);
    CU(cudaGetLastError());
    return 0;
}

// planes [ka, kb) of the components of one lattice (or of the macrodata) between a host FAB (ghost width ng, interior only)
// and the padded SoA buffer: one pitched DMA per component, no staging kernel
// (with_ghosts: also the FAB's ghost cells the device layout has room for, as mbl_upload does)
static int copy_planes(Level& lv, double* soa, double* fab, int ng, int ka, int kb, bool to_device, cudaStream_t st,
                       bool with_ghosts = false, int ncomp = NQ)
{
    const Layout& L = lv.L;
    const size_t sx = L.nx + 2 * ng, sy = L.ny + 2 * ng, n = sx * sy * (L.nz + 2 * ng);
    const int gx = with_ghosts ? std::min(ng, GX) : 0, gy = with_ghosts ? std::min(ng, GY) : 0,
              gz = with_ghosts ? std::min(ng, GZ) : 0;
    ka -= gz, kb += gz;
    for (int q = 0; q < ncomp; ++q) {
        cudaMemcpy3DParms p;
        memset(&p, 0, sizeof(p));
        cudaPitchedPtr host = make_cudaPitchedPtr(fab + (size_t)q * n, sx * sizeof(double), sx, sy);
        cudaPitchedPtr dev = make_cudaPitchedPtr(soa + (size_t)q * L.sq, L.px * sizeof(double), L.px, L.ny + 2 * GY);
        const cudaPos hpos = make_cudaPos((size_t)(ng - gx) * sizeof(double), ng - gy, ng + ka);
        const cudaPos dpos = make_cudaPos((size_t)(OX - gx) * sizeof(double), GY - gy, GZ + ka);
        p.srcPtr = to_device ? host : dev;
        p.srcPos = to_device ? hpos : dpos;
        p.dstPtr = to_device ? dev : host;
        p.dstPos = to_device ? dpos : hpos;
        p.extent = make_cudaExtent((L.nx + 2 * gx) * sizeof(double), L.ny + 2 * gy, kb - ka);
        p.kind = cudaMemcpyDefault;  // the FAB may live in host memory or in device memory (AMReX device arena)
        CU(cudaMemcpy3DAsync(&p, st));
    }
    return 0;
}

int mbl_upload(mbl_ctx* ctx, int lev, int which, const double* fab, int ng)
{
    if (check_level(ctx, lev)) return 1;
    if (!fab || ng < 0) return fail("mbl_upload: bad argument");
    Level& lv = ctx->lev[lev];
    lv.carry_valid = false;
    CU(cudaSetDevice(ctx->device));
    // valid cells plus the FAB ghost cells the device layout has room for, one pitched DMA per component
    if (copy_planes(lv, which == MBL_G ? curg(lv) : curf(lv), const_cast<double*>(fab), ng, 0, lv.L.nz, true, ctx->stream,
                    true))
        return 1;
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int download_comps(mbl_ctx* ctx, Level& lv, const double* src, int ncomp, double* fab, int ng)
{
    // valid cells only, straight into the interior of the host FAB (its ghost cells are left untouched)
    if (copy_planes(lv, const_cast<double*>(src), fab, ng, 0, lv.L.nz, false, ctx->stream, false, ncomp)) return 1;
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mbl_download(mbl_ctx* ctx, int lev, int which, double* fab, int ng)
{
    if (check_level(ctx, lev)) return 1;
    if (!fab || ng < 0) return fail("mbl_download: bad argument");
    Level& lv = ctx->lev[lev];
    CU(cudaSetDevice(ctx->device));
    return download_comps(ctx, lv, which == MBL_G ? curg(lv) : curf(lv), NQ, fab, ng);
}

int mbl_download_macrodata(mbl_ctx* ctx, int lev, double* fab, int ng)
{
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    if (!lv.macro) return fail("no macrodata yet: call mbl_collide/mbl_step with want_macrodata or mbl_f_to_macrodata");
    CU(cudaSetDevice(ctx->device));
    return download_comps(ctx, lv, lv.macro, MBL_NMACRO, fab, ng);
}

int mbl_download_derived(mbl_ctx* ctx, int lev, double* fab)
{
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    if (!lv.macro) return fail("no derived data yet");
    CU(cudaSetDevice(ctx->device));
    return download_comps(ctx, lv, lv.macro + (size_t)MBL_NMACRO * lv.L.sq, MBL_NDERIVED, fab, 0);
}

int mbl_initialize(mbl_ctx* ctx, int lev, int ic_kind, const double* v, int nv)
{
    const bool patch = is_patch(ctx, lev);
    if (!patch && check_level(ctx, lev)) return 1;
    if (nv < 16 || !v) return fail("mbl_initialize: need 16 parameters");
    if (ic_kind < 0 || ic_kind > 4) return fail("mbl_initialize: unknown initial condition %d", ic_kind);
    CU(cudaSetDevice(ctx->device));
    IcInfo I;
    I.kind = ic_kind;
    I.density = v[0];
    I.vel[0] = v[1], I.vel[1] = v[2], I.vel[2] = v[3];
    I.v0 = v[4];
    I.omega[0] = v[5], I.omega[1] = v[6], I.omega[2] = v[7];
    I.wave_length = v[8];
    I.T0 = v[9], I.gamma = v[10], I.R = v[11], I.c_s = v[12];
    I.density_ratio = v[13], I.temperature_ratio = v[14], I.x_disc = v[15];
    if (patch) return patch_initialize(ctx, lev, I);
    Level& lv = ctx->lev[lev];
    lv.carry_valid = false;
    ctx->launches += launch_initialize(lv.L, lv.B, I, lv.p.flag, curf(lv), curg(lv), ctx->stream);
    CU(cudaGetLastError());
    return 0;
}

int mbl_fillpatch(mbl_ctx* ctx, int lev, double /*time*/)
{
    if (is_patch(ctx, lev)) {
        CU(cudaSetDevice(ctx->device));
        return patch_fillpatch(ctx, lev, 0.0);
    }
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    CU(cudaSetDevice(ctx->device));
    ctx->launches += launch_ghost_fill(lv.L, lv.B, curf(lv), curg(lv), lv.local_z, true, true, ctx->stream);
    CU(cudaGetLastError());
    return 0;
}

int mbl_physbc(mbl_ctx* ctx, int lev, double /*time*/)
{
    if (is_patch(ctx, lev)) {
        CU(cudaSetDevice(ctx->device));
        return patch_physbc(ctx, lev, 0.0);
    }
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    CU(cudaSetDevice(ctx->device));
    ctx->launches += launch_ghost_fill(lv.L, lv.B, curf(lv), curg(lv), lv.local_z, false, false, ctx->stream);
    CU(cudaGetLastError());
    return 0;
}

int mbl_stream(mbl_ctx* ctx, int lev)
{
    if (is_patch(ctx, lev)) {
        CU(cudaSetDevice(ctx->device));
        return patch_stream(ctx, lev);
    }
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    lv.carry_valid = false;
    CU(cudaSetDevice(ctx->device));
    const int a = lv.cur, b = 1 - lv.cur;
    ctx->launches += launch_stream(lv.L, lv.p.f[a], lv.p.g[a], lv.p.f[b], lv.p.g[b], lv.p.nbr, ctx->stream);
    lv.cur = b;
    CU(cudaGetLastError());
    return 0;
}

// LBM::advance(lev) (LBM.cpp:523-544)
int mbl_advance(mbl_ctx* ctx, int lev, int want_macro)
{
    if (is_patch(ctx, lev)) {
        CU(cudaSetDevice(ctx->device));
        return patch_advance(ctx, lev, want_macro);
    }
    return mbl_stream(ctx, lev) || mbl_collide(ctx, lev, want_macro);
}

int mbl_collide(mbl_ctx* ctx, int lev, int want_macro)
{
    if (is_patch(ctx, lev)) {
        CU(cudaSetDevice(ctx->device));
        return patch_collide(ctx, lev, want_macro);
    }
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    lv.carry_valid = false;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (want_macro && ensure_macro(lv, st, ctx->launches)) return 1;
    if (want_macro) lv.dq_from_macro = false;
    double *f = curf(lv), *g = curg(lv);
    ctx->launches += launch_qcorr(lv.L, lv.P, f, g, lv.p.nbr, lv.p.qc, false, st);
    ctx->launches += launch_collide(lv.L, lv.P, f, g, f, g, lv.p.nbr, lv.p.flag, lv.p.qc,
                                    want_macro ? lv.macro : nullptr, false, st);
    CU(cudaGetLastError());
    return 0;
}

int mbl_f_to_macrodata(mbl_ctx* ctx, int lev)
{
    if (is_patch(ctx, lev)) {
        CU(cudaSetDevice(ctx->device));
        return patch_f_to_macrodata(ctx, lev);
    }
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    CU(cudaSetDevice(ctx->device));
    if (ensure_macro(lv, ctx->stream, ctx->launches)) return 1;
    ctx->launches += launch_macrodata(lv.L, lv.P, curf(lv), curg(lv), lv.p.flag, lv.macro, ctx->stream);
    lv.dq_from_macro = true;
    CU(cudaGetLastError());
    return 0;
}

int mbl_compute_derived(mbl_ctx* ctx, int lev)
{
    if (is_patch(ctx, lev)) {
        CU(cudaSetDevice(ctx->device));
        return patch_compute_derived(ctx, lev);
    }
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    if (!lv.macro) return fail("mbl_compute_derived needs macrodata");
    CU(cudaSetDevice(ctx->device));
    ctx->launches += launch_derived(lv.L, lv.P, lv.p.flag, lv.macro, lv.macro + (size_t)MBL_NMACRO * lv.L.sq, ctx->stream,
                                    lv.dq_from_macro ? 1 : 0);
    CU(cudaGetLastError());
    return 0;
}

// macrodata planes a z-neighbour needs for compute_derived on a slab: velocity (vorticity) and QCorr (their
// differences), one valid plane per side.  side 0 = low z, 1 = high z; pack copies the outermost valid plane into
// buf, unpack copies a neighbour's plane into the ghost plane on that side.  Whole padded planes, plain copies.
static const int MACRO_HALO_COMPS[6] = {1, 2, 3, 6, 7, 8};

int64_t mbl_macro_halo_doubles(mbl_ctx* ctx, int lev)
{
    if (check_level(ctx, lev)) return -1;
    return 6LL * ctx->lev[lev].L.sz;
}

int mbl_macro_halo(mbl_ctx* ctx, int lev, int side, double* buf, int pack)
{
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    if (!lv.macro) return fail("mbl_macro_halo needs macrodata");
    if (!buf || side < 0 || side > 1) return fail("mbl_macro_halo: bad argument");
    CU(cudaSetDevice(ctx->device));
    const Layout& L = lv.L;
    const int plane = pack ? (side == 0 ? 0 : L.nz - 1) : (side == 0 ? -1 : L.nz);
    for (int n = 0; n < 6; ++n) {
        double* p = lv.macro + (size_t)MACRO_HALO_COMPS[n] * L.sq + (size_t)(plane + GZ) * L.sz;
        double* b = buf + (size_t)n * L.sz;
        CU(cudaMemcpyAsync(pack ? b : p, pack ? p : b, (size_t)L.sz * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return 0;
}

int mbl_compute_derived_slab(mbl_ctx* ctx, int lev, int has_lo, int has_hi)
{
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    if (!lv.macro) return fail("mbl_compute_derived needs macrodata");
    CU(cudaSetDevice(ctx->device));
    ctx->launches += launch_derived(lv.L, lv.P, lv.p.flag, lv.macro, lv.macro + (size_t)MBL_NMACRO * lv.L.sq, ctx->stream,
                                    lv.dq_from_macro ? 1 : 0, (has_lo ? 1 : 0) | (has_hi ? 2 : 0));
    CU(cudaGetLastError());
    return 0;
}

int mbl_eb_forces(mbl_ctx* ctx, int lev, double out[3])
{
    if (is_patch(ctx, lev)) {
        CU(cudaSetDevice(ctx->device));
        return patch_eb_forces(ctx, lev, out);
    }
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    CU(cudaSetDevice(ctx->device));
    // the reference reads f in ghost cells that FillBoundary refreshed after the collision
    // (LBM.cpp:805); refresh the periodic images here
    ctx->launches += launch_ghost_fill(lv.L, lv.B, curf(lv), curg(lv), lv.local_z, false, true, ctx->stream);
    ctx->launches += launch_eb_forces(lv.L, curf(lv), lv.p.flag, lv.d_red, ctx->stream);
    CU(cudaMemcpyAsync(out, lv.d_red, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mbl_step_local(mbl_ctx* ctx, int lev, double time, int want_macro)
{
    if (check_level(ctx, lev)) return 1;
    CU(cudaSetDevice(ctx->device));
    return step_local(ctx, ctx->lev[lev], time, want_macro);
}

int mbl_step(mbl_ctx* ctx, int lev, int nsteps, double time, int want_macro)
{
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    if (!lv.local_z && nsteps > 1) return fail("mbl_step: nsteps > 1 needs a box that spans the domain in z");
    CU(cudaSetDevice(ctx->device));
    int s = 0;
    // Launch-bound boxes: replay pairs of steps as one CUDA graph.  The first step runs eagerly (it allocates
    // the carry arrays and brings the level into its steady kernel sequence), the last one too if macrodata
    // is wanted; variants 1-3 (persistent kernels with host-side plans) and timed runs are not captured.
    const bool graphable = ctx->use_graphs && !ctx->timing && (ctx->variant == 0 || ctx->variant >= 4) &&
                           lv.L.sq < (1LL << 24);  // ~16 M cells: beyond that launches are noise
    const int tail = want_macro ? 1 : 0;
    if (graphable && nsteps - tail >= 5) {
        if (step_local(ctx, lv, time, 0)) return 1;
        s = 1;
        if (lv.graph && (lv.graph_cur != lv.cur || lv.graph_variant != ctx->variant || lv.graph_qc != lv.p.qc)) {
            cudaGraphExecDestroy(lv.graph);
            lv.graph = nullptr;
        }
        if (!lv.graph) {
            // capture on a stream of our own (the caller's may be the legacy default stream, which cannot be
            // captured); the instantiated graph is launched on the caller's stream
            if (!ctx->s_capture) CU(cudaStreamCreateWithFlags(&ctx->s_capture, cudaStreamNonBlocking));
            cudaGraph_t g = nullptr;
            const int64_t l0 = ctx->launches;
            cudaStream_t user = ctx->stream;
            ctx->stream = ctx->s_capture;
            cudaError_t ce = cudaStreamBeginCapture(ctx->s_capture, cudaStreamCaptureModeThreadLocal);
            int rc = 0;
            if (ce == cudaSuccess) {
                rc = step_local(ctx, lv, time, 0) || step_local(ctx, lv, time, 0);
                ce = cudaStreamEndCapture(ctx->s_capture, &g);
            }
            ctx->stream = user;
            const int64_t captured = ctx->launches - l0;
            ctx->launches = l0;  // nothing ran
            if (!rc && ce == cudaSuccess && cudaGraphInstantiate(&lv.graph, g, 0) == cudaSuccess)
                lv.graph_launches = captured;
            if (g) cudaGraphDestroy(g);
            if (rc || ce != cudaSuccess || !lv.graph) {
                // not capturable here (driver, lazy loading, ...): run eagerly from now on
                cudaGetLastError();
                lv.graph = nullptr;
                ctx->use_graphs = false;
                if (rc) return 1;
            } else {
                lv.graph_cur = lv.cur;  // two steps: the parity is back where the capture started
                lv.graph_variant = ctx->variant;
                lv.graph_qc = lv.p.qc;
            }
        }
        for (; lv.graph && s + 2 <= nsteps - tail; s += 2) {
            CU(cudaGraphLaunch(lv.graph, ctx->stream));
            ctx->launches += lv.graph_launches;
        }
    }
    for (; s < nsteps; ++s)
        if (step_local(ctx, lv, time + s * lv.P.dt, want_macro && s == nsteps - 1)) return 1;
    return 0;
}

// Overlapped slab step.  Output plane k needs input planes k-2 .. k+2 (the q-correction of k+-1 pulls from
// k+-2), so only the GZ = 2 outermost planes at each z-end depend on the neighbours' ghost planes.
//   part 0: q-corrections of planes [-1, 3) and [nz-3, nz+1), collide of planes [0, 2) and [nz-2, nz)
//   part 1: q-corrections of planes [3, nz-3), collide of planes [2, nz-2); the written buffers become current
// Between the two the caller packs the freshly written boundary planes (mbl_halo_pack_next), exchanges them
// and unpacks them into the ghost planes of the written buffers (mbl_halo_unpack_next) on another stream.
int mbl_step_split(mbl_ctx* ctx, int lev, int part)
{
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    const Layout& L = lv.L;
    if (L.nz < 8) return fail("mbl_step_split: a slab needs at least 8 planes");
    if (part != 0 && part != 1) return fail("mbl_step_split: part must be 0 or 1");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int a = lv.cur, b = 1 - lv.cur, nz = L.nz;
    auto mark = [&]() {
        if (!ctx->timing) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ctx->events.push_back(e);
    };
    // the carry variants keep carrying through the split: q-corrections from the previous step's partial sums
    // (k_qcorr_combine*; planes next to a ghost plane are pulled as before), collide by k_collide_tile*
    const bool tile = (ctx->variant == 5 || ctx->variant == 7 || ctx->variant == 9 || ctx->variant == 10) && L.sq * 8 < (1LL << 32);
    const int W = carry_tile_rows(ctx->carry_rows);
    // variant 9: the z-march kernels; the three plane ranges of the split are chunked on their own -- [0, 2) and
    // [nz-2, nz) are one chunk each, [2, nz-2) is cut into chunks of zm planes -- and zpos describes that layout
    const bool zmarch = tile && ctx->variant == 9 && W == 6;
    int zm = ctx->zmarch > 1 ? ctx->zmarch : 8;
    while ((nz - 4) % zm == 1) ++zm;
    const int zkey = -zm;  // layout key of the split (step_local's is +zm)
    if (tile) {
        if (!lv.part) CU(cudaMalloc(&lv.part, (size_t)CARRY_WORDS * L.sq * sizeof(double)));
        if (!lv.edge)
            CU(cudaMalloc(&lv.edge, (size_t)CARRY_EDGE_WORDS * carry_edge_plane(L, 4) * (L.nz + 2 * GZ) * sizeof(double)));
    }
    if (zmarch) {
        if (!lv.qc2) {
            CU(cudaMalloc(&lv.qc2, (size_t)3 * L.sq * sizeof(double)));
            CU(cudaMemsetAsync(lv.qc2, 0, (size_t)3 * L.sq * sizeof(double), st));
        }
        if (!lv.zpos) CU(cudaMalloc(&lv.zpos, (size_t)(nz + 2 * GZ)));
    }
    // sums of a z-march step can only be combined with the table of the layout that wrote them (lv.zpos_zm)
    // (nothing part 0 does changes these, so part 1 decides the same)
    const bool from_march = lv.carry_valid && lv.part_pair == 2 && lv.edge_rows == 6 && lv.zpos && lv.zpos_zm != 0;
    const bool from_sums = tile && lv.carry_valid && lv.edge_rows == W && !lv.part_pair;
    const CarryPlan C = make_carry_plan(L, ctx->carry_own, ctx->carry_ky);
#ifdef MBL_EXPERIMENTS
    const bool march = ctx->variant == 8 && L.sq * 8 < (1LL << 32);  // one kernel: the q-correction ranges are empty
#else
    const bool march = false;
#endif
    auto q = [&](int ka, int kb) {
        if (march) return;
        if (from_march)
            ctx->launches += launch_qcorr_combine_march(L, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.part, lv.edge, 6, lv.zpos,
                                                        lv.p.qc, st, ka, kb);
        else if (from_sums)
            ctx->launches += launch_qcorr_combine(L, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.part, lv.edge, W, lv.p.qc, st,
                                                  ka, kb);
        else
            ctx->launches += launch_qcorr(L, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.p.qc, true, st, ka, kb);
    };
    auto c = [&](int ka, int kb) {
#ifdef MBL_EXPERIMENTS
        if (march) {
            ctx->launches += launch_march(L, lv.P, ctx->march_rows, ctx->march_zm, ctx->march_pipe, lv.p.f[a], lv.p.g[a],
                                          lv.p.f[b], lv.p.g[b], lv.p.nbr, lv.p.flag, st, ka, kb);
            return;
        }
#endif
        if (zmarch)
            ctx->launches += launch_collide_tile_march(L, lv.P, C, 6, false, zm, lv.p.f[a], lv.p.g[a], lv.p.f[b], lv.p.g[b],
                                                       lv.p.nbr, lv.p.flag, lv.p.qc, lv.qc2, lv.part, lv.edge, st, ka, kb);
        else if (tile)
            ctx->launches += launch_collide_tile(L, lv.P, C, W, lv.p.f[a], lv.p.g[a], lv.p.f[b], lv.p.g[b], lv.p.nbr,
                                                 lv.p.flag, lv.p.qc, lv.part, lv.edge, st, ka, kb);
        else
            ctx->launches += launch_collide(L, lv.P, lv.p.f[a], lv.p.g[a], lv.p.f[b], lv.p.g[b], lv.p.nbr, lv.p.flag,
                                            lv.p.qc, nullptr, true, st, ka, kb);
    };
    if (part == 0) {
        // a z-end that is a periodic image of the box itself (single rank in z) has no ghost planes to wait for,
        // but the split is still valid: the kernels wrap
        const int klo = L.wrap[2] ? 0 : -1, khi = L.wrap[2] ? nz : nz + 1;
        mark();
        // levels with walls / inlets / outlets: the ghost cells of the CURRENT buffers (whose z ghost planes the
        // previous exchange delivered) are filled once, before either part reads them -- K6, periodic images in x, y,
        // BCFill regions, exactly as mbl_step_local does
        if (!(L.wrap[0] && L.wrap[1]))
            ctx->launches += launch_ghost_fill(L, lv.B, lv.p.f[a], lv.p.g[a], lv.local_z, true, true, st);
        mark();
        q(klo, 3);
        q(nz - 3, khi);
        mark();
        c(0, 2);
        c(nz - 2, nz);
        mark();
    } else {
        mark(), mark();
        if (nz > 6) q(3, nz - 3);
        mark();
        c(2, nz - 2);
        mark();
        if (ctx->timing) ctx->timed_steps++;
        lv.cur = b;
        lv.carry_valid = tile;
        lv.edge_rows = tile ? W : 0;
        lv.part_pair = zmarch ? 2 : 0;
        if (zmarch) {
            std::swap(lv.p.qc, lv.qc2);  // the cells the collide kernels finished are in what is now lv.p.qc
            if (lv.zpos_zm != zkey) {
                // the table of the layout the sums now have (the combine launches above used the previous one)
                std::vector<signed char> zp((size_t)nz + 2 * GZ, 0);
                zp[0 + GZ] = 1, zp[1 + GZ] = 2, zp[nz - 2 + GZ] = 1, zp[nz - 1 + GZ] = 2;
                for (int k = 2; k < nz - 2; ++k) {
                    const int kk = (k - 2) % zm, nk = std::min(zm, nz - 2 - (k - kk));
                    zp[k + GZ] = kk == 0 ? 1 : kk == nk - 1 ? 2 : 0;
                }
                CU(cudaMemcpyAsync(lv.zpos, zp.data(), zp.size(), cudaMemcpyHostToDevice, st));
                CU(cudaStreamSynchronize(st));
                lv.zpos_zm = zkey;
            }
        }
    }
    CU(cudaGetLastError());
    return 0;
}

int mbl_halo_pack_next(mbl_ctx* ctx, int lev, int side, double* buf)
{
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    CU(cudaSetDevice(ctx->device));
    ctx->launches += launch_halo_pack(lv.L, lv.p.f[1 - lv.cur], lv.p.g[1 - lv.cur], side, buf, ctx->stream, ctx->halo_lean);
    CU(cudaGetLastError());
    return 0;
}

int mbl_halo_unpack_next(mbl_ctx* ctx, int lev, int side, const double* buf)
{
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    CU(cudaSetDevice(ctx->device));
    ctx->launches += launch_halo_unpack(lv.L, lv.p.f[1 - lv.cur], lv.p.g[1 - lv.cur], side, buf, ctx->stream, ctx->halo_lean);
    CU(cudaGetLastError());
    return 0;
}

int64_t mbl_halo_doubles(mbl_ctx* ctx, int lev)
{
    if (check_level(ctx, lev)) return -1;
    return (ctx->halo_lean ? 1LL : (long long)GZ) * 2LL * NQ * ctx->lev[lev].L.sz;
}

int mbl_set_halo_lean(mbl_ctx* ctx, int on)
{
    if (!ctx) return fail("null context");
    ctx->halo_lean = on != 0;
    return 0;
}

int mbl_halo_pack(mbl_ctx* ctx, int lev, int side, double* buf)
{
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    CU(cudaSetDevice(ctx->device));
    ctx->launches += launch_halo_pack(lv.L, curf(lv), curg(lv), side, buf, ctx->stream, ctx->halo_lean);
    CU(cudaGetLastError());
    return 0;
}

int mbl_halo_unpack(mbl_ctx* ctx, int lev, int side, const double* buf)
{
    if (check_level(ctx, lev)) return 1;
    Level& lv = ctx->lev[lev];
    CU(cudaSetDevice(ctx->device));
    ctx->launches += launch_halo_unpack(lv.L, curf(lv), curg(lv), side, buf, ctx->stream, ctx->halo_lean);
    CU(cudaGetLastError());
    return 0;
}

// One step of an all-periodic box that spans the domain, with the host<->device copies of the step's input and
// result overlapped with the kernels: the box is cut into z-chunks; chunk c+1 is uploaded (copy stream) while
// the q-correction and collide kernels run on the planes whose three-plane neighbourhoods are already on the
// device (compute stream) and finished planes of the result go back to the host (second copy stream).  The
// two planes at the top are uploaded first, because the periodic wrap makes plane 0 depend on them.
static int step_host_pipelined(mbl_ctx* ctx, Level& lv, double* f_fab, double* g_fab, int ng)
{
    const Layout& L = lv.L;
    const int nz = L.nz;
    if (!ctx->s_up) {
        CU(cudaStreamCreateWithFlags(&ctx->s_up, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&ctx->s_down, cudaStreamNonBlocking));
    }
    cudaStream_t sc = ctx->stream, su = ctx->s_up, sd = ctx->s_down;
    const int a = lv.cur, b = 1 - lv.cur;
    const int top = nz - 2;  // planes [top, nz) go first
    const int cz = ctx->host_chunk > 0 ? ctx->host_chunk : 16;
    std::vector<cudaEvent_t> evs;
    auto event = [&](cudaStream_t st) {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        cudaEventRecord(e, st);
        evs.push_back(e);
        return e;
    };
    // the copy streams start after whatever the caller queued on the compute stream
    cudaEvent_t e0 = event(sc);
    CU(cudaStreamWaitEvent(su, e0, 0));
    CU(cudaStreamWaitEvent(sd, e0, 0));
    if (copy_planes(lv, lv.p.f[a], f_fab, ng, top, nz, true, su)) return 1;
    if (copy_planes(lv, lv.p.g[a], g_fab, ng, top, nz, true, su)) return 1;
    int qdone = 0, cdone = 0;
    bool first = true;
    for (int u0 = 0; u0 < top; u0 += cz) {
        const int u = std::min(top, u0 + cz);  // planes [0, u) and [top, nz) are on the device after this upload
        if (copy_planes(lv, lv.p.f[a], f_fab, ng, u0, u, true, su)) return 1;
        if (copy_planes(lv, lv.p.g[a], g_fab, ng, u0, u, true, su)) return 1;
        CU(cudaStreamWaitEvent(sc, event(su), 0));
        if (first) {  // plane nz-1 pulls from nz-2, nz-1 and 0
            ctx->launches += launch_qcorr(L, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.p.qc, true, sc, nz - 1, nz);
            first = false;
        }
        const int qhi = (u == top) ? nz - 1 : u - 1;  // q-correction of plane k pulls from plane k+1
        if (qhi > qdone) {
            ctx->launches += launch_qcorr(L, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.p.qc, true, sc, qdone, qhi);
            qdone = qhi;
        }
        const int chi = (qdone == nz - 1) ? nz : qdone - 1;  // collide of plane k differences the q-correction of k+1
        if (chi > cdone) {
            ctx->launches += launch_collide(L, lv.P, lv.p.f[a], lv.p.g[a], lv.p.f[b], lv.p.g[b], lv.p.nbr, lv.p.flag,
                                            lv.p.qc, nullptr, true, sc, cdone, chi);
            CU(cudaStreamWaitEvent(sd, event(sc), 0));
            if (copy_planes(lv, lv.p.f[b], f_fab, ng, cdone, chi, false, sd)) return 1;
            if (copy_planes(lv, lv.p.g[b], g_fab, ng, cdone, chi, false, sd)) return 1;
            cdone = chi;
        }
    }
    lv.cur = b;
    lv.carry_valid = false;
    CU(cudaStreamWaitEvent(sc, event(sd), 0));  // later work on the caller's stream sees the finished step
    CU(cudaStreamSynchronize(sd));              // the host buffers are valid on return
    CU(cudaStreamSynchronize(sc));
    for (cudaEvent_t e : evs) cudaEventDestroy(e);
    CU(cudaGetLastError());
    return 0;
}

// The same for a z-slab of a multi-rank run (all-periodic level, not wrapped in z): the two outermost planes at
// each z-end go up first (begin), the caller exchanges them with the neighbours (mbl_halo_pack / transport /
// mbl_halo_unpack on the context's stream), then finish() uploads the interior planes in chunks while the kernels
// follow the upload frontier and finished planes go back down.
int mbl_step_host_begin(mbl_ctx* ctx, int lev, double* f_fab, double* g_fab, int ng)
{
    if (check_level(ctx, lev)) return 1;
    if (!f_fab || !g_fab || ng < 0) return fail("mbl_step_host_begin: bad argument");
    Level& lv = ctx->lev[lev];
    const Layout& L = lv.L;
    if (!(L.wrap[0] && L.wrap[1]) || L.wrap[2] || L.nz < 8)
        return fail("mbl_step_host_begin: for z-slabs (at least 8 planes) of all-periodic levels");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->s_up) {
        CU(cudaStreamCreateWithFlags(&ctx->s_up, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&ctx->s_down, cudaStreamNonBlocking));
    }
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU(cudaEventRecord(e, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->s_up, e, 0));
    const int a = lv.cur, nz = L.nz;
    for (int pass = 0; pass < 2; ++pass) {
        const int ka = pass ? nz - GZ : 0, kb = pass ? nz : GZ;
        if (copy_planes(lv, lv.p.f[a], f_fab, ng, ka, kb, true, ctx->s_up)) return 1;
        if (copy_planes(lv, lv.p.g[a], g_fab, ng, ka, kb, true, ctx->s_up)) return 1;
    }
    CU(cudaEventRecord(e, ctx->s_up));
    CU(cudaStreamWaitEvent(ctx->stream, e, 0));  // the halo pack that follows reads these planes
    CU(cudaEventDestroy(e));
    lv.carry_valid = false;
    return 0;
}

int mbl_step_host_finish(mbl_ctx* ctx, int lev, double* f_fab, double* g_fab, int ng)
{
    if (check_level(ctx, lev)) return 1;
    if (!f_fab || !g_fab || ng < 0) return fail("mbl_step_host_finish: bad argument");
    Level& lv = ctx->lev[lev];
    const Layout& L = lv.L;
    if (!(L.wrap[0] && L.wrap[1]) || L.wrap[2] || L.nz < 8 || !ctx->s_up)
        return fail("mbl_step_host_finish without mbl_step_host_begin");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t sc = ctx->stream, su = ctx->s_up, sd = ctx->s_down;
    const int a = lv.cur, b = 1 - lv.cur, nz = L.nz;
    const int cz = ctx->host_chunk > 0 ? ctx->host_chunk : 16;
    std::vector<cudaEvent_t> evs;
    auto event = [&](cudaStream_t st) {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        cudaEventRecord(e, st);
        evs.push_back(e);
        return e;
    };
    cudaEvent_t e0 = event(sc);  // the ghost planes are in (halo unpack on the compute stream)
    CU(cudaStreamWaitEvent(sd, e0, 0));
    // planes [-2, 2) and [nz-2, nz+2) are on the device: q-corrections that need nothing else
    ctx->launches += launch_qcorr(L, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.p.qc, true, sc, -1, 1);
    ctx->launches += launch_qcorr(L, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.p.qc, true, sc, nz - 1, nz + 1);
    int qdone = 1, cdone = 0;
    const int top = nz - GZ;
    for (int u0 = GZ; u0 < top; u0 += cz) {
        const int u = std::min(top, u0 + cz);  // planes [-2, u) and [top, nz+2) are on the device after this upload
        if (copy_planes(lv, lv.p.f[a], f_fab, ng, u0, u, true, su)) return 1;
        if (copy_planes(lv, lv.p.g[a], g_fab, ng, u0, u, true, su)) return 1;
        CU(cudaStreamWaitEvent(sc, event(su), 0));
        const int qhi = (u == top) ? nz - 1 : u - 1;
        if (qhi > qdone) {
            ctx->launches += launch_qcorr(L, lv.P, lv.p.f[a], lv.p.g[a], lv.p.nbr, lv.p.qc, true, sc, qdone, qhi);
            qdone = qhi;
        }
        const int chi = (qdone == nz - 1) ? nz : qdone - 1;
        if (chi > cdone) {
            ctx->launches += launch_collide(L, lv.P, lv.p.f[a], lv.p.g[a], lv.p.f[b], lv.p.g[b], lv.p.nbr, lv.p.flag,
                                            lv.p.qc, nullptr, true, sc, cdone, chi);
            CU(cudaStreamWaitEvent(sd, event(sc), 0));
            if (copy_planes(lv, lv.p.f[b], f_fab, ng, cdone, chi, false, sd)) return 1;
            if (copy_planes(lv, lv.p.g[b], g_fab, ng, cdone, chi, false, sd)) return 1;
            cdone = chi;
        }
    }
    lv.cur = b;
    lv.carry_valid = false;
    CU(cudaStreamWaitEvent(sc, event(sd), 0));
    CU(cudaStreamSynchronize(sd));
    CU(cudaStreamSynchronize(sc));
    for (cudaEvent_t e : evs) cudaEventDestroy(e);
    CU(cudaGetLastError());
    return 0;
}

int mbl_step_host(mbl_ctx* ctx, int lev, int nsteps, double time, double* f_fab, double* g_fab, int ng)
{
    if (check_level(ctx, lev)) return 1;
    if (!f_fab || !g_fab || ng < 0 || nsteps < 1) return fail("mbl_step_host: bad argument");
    Level& lv = ctx->lev[lev];
    CU(cudaSetDevice(ctx->device));
    const Layout& L = lv.L;
    if (nsteps == 1 && ctx->host_chunk >= 0 && L.wrap[0] && L.wrap[1] && L.wrap[2] && L.nz >= 8)
        return step_host_pipelined(ctx, lv, f_fab, g_fab, ng);
    // general boxes: upload everything (the step refills every ghost cell), step, download
    lv.carry_valid = false;
    if (copy_planes(lv, curf(lv), f_fab, ng, 0, L.nz, true, ctx->stream, true)) return 1;
    if (copy_planes(lv, curg(lv), g_fab, ng, 0, L.nz, true, ctx->stream, true)) return 1;
    if (mbl_step(ctx, lev, nsteps, time, 0)) return 1;
    if (copy_planes(lv, curf(lv), f_fab, ng, 0, L.nz, false, ctx->stream)) return 1;
    if (copy_planes(lv, curg(lv), g_fab, ng, 0, L.nz, false, ctx->stream)) return 1;
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int64_t mbl_launch_count(mbl_ctx* ctx) { return ctx ? ctx->launches : -1; }

int mbl_set_timing(mbl_ctx* ctx, int on)
{
    if (!ctx) return fail("null context");
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    ctx->events.clear();
    ctx->timed_steps = 0;
    ctx->timing = on != 0;
    return 0;
}

int mbl_get_timing(mbl_ctx* ctx, double ms[3], int* nsteps)
{
    if (!ctx || !ms || !nsteps) return fail("null argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ms[0] = ms[1] = ms[2] = 0.0;
    *nsteps = ctx->timed_steps;
    const int nrec = (int)(ctx->events.size() / 4);
    for (int s = 0; s < nrec; ++s)
        for (int k = 0; k < 3; ++k) {
            float t = 0.f;
            CU(cudaEventElapsedTime(&t, ctx->events[4 * s + k], ctx->events[4 * s + k + 1]));
            ms[k] += t;
        }
    return mbl_set_timing(ctx, ctx->timing ? 1 : 0);
}

int mbl_get_variant(mbl_ctx* ctx) { return ctx ? ctx->variant : -1; }

int mbl_set_variant(mbl_ctx* ctx, int variant)
{
    if (!ctx) return fail("null context");
    if (variant < 0 || variant > 10) return fail("variant %d is not available", variant);
#ifndef MBL_EXPERIMENTS
    if ((variant >= 1 && variant <= 4) || variant == 8 || variant == 10)
        return fail("step variant %d is an experiment: rebuild the library with MBL_EXPERIMENTS=1", variant);
#endif
    ctx->variant = variant;
    for (int l = 0; l < MAX_LEVELS; ++l) ctx->lev[l].carry_valid = false;
    return 0;
}

}  // extern "C"
