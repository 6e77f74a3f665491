// kernels.cu -- CUDA kernels (sm_100a) of the D3Q27 f+g lattice update.
//
//   k_qcorr      pass 1: pull f,g -> rho,u,T -> QCorr_x,y,z      (LBM.cpp:810-906, pulled LBM.cpp:558-604)
//   k_collide    pass 2: pull f,g -> moments -> grad QCorr -> feq/geq -> BGK relax -> store
//                (LBM.cpp:558-604 + 607-618 fused; pull=false: collide only, in place)
//   k_stream     pull-scheme stream with halfway bounce-back (LBM.cpp:558-604)
//   ghost fill   K6 pre-pass (FillPatchOps.H:92-108), periodic wrap (FillBoundary), BCFill regions
//                (BC.H:345-471 in the region order of AMReX_PhysBCFunct.H:593-678)
//   flags        is_fluid (LBM.cpp:1213-1262) -> flag byte field + 27-bit pull mask
//
// All kernels are bandwidth-bound fp64 stencil/byte work: no tensor cores.  Thread x
// is the fastest index so every load/store of a component plane is coalesced; stores
// and e_x = 0 loads are 128-byte aligned, e_x = +-1 loads are shifted by one element.
#include "kernels.cuh"
#include "bcic.cuh"

#include <cstdio>
#include <cstdlib>

namespace mbl {


void init_tables()
{
    DirTables t;
    for (int q = 0; q < NQ; ++q) {
        t.ex[q] = ex(q);
        t.ey[q] = ey(q);
        t.ez[q] = ez(q);
        t.opp[q] = opp(q);
        t.mx[q] = mirror_x(q);
        t.my[q] = mirror_y(q);
        t.mz[q] = mirror_z(q);
        t.w[q] = weight(q);
    }
    cudaMemcpyToSymbol(c_dir, &t, sizeof(t));
}

// ---------------------------------------------------------------------------
// pull of one lattice (all 27 directions) for cell c
// ---------------------------------------------------------------------------
// L2 eviction-priority hints (createpolicy + ld/st .L2::cache_hint).  The fused kernel touches every
// population twice: the q-correction job's pull should stay in L2 (HINT_KEEP) until the collide job of the
// same row re-reads it (HINT_LAST: last use) and the collide job's stores must not push it out.
constexpr int HINT_NONE = 0, HINT_KEEP = 1, HINT_LAST = 2;
__device__ __forceinline__ uint64_t make_policy(int hint)
{
    uint64_t p = 0;
    if (hint == HINT_KEEP) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    if (hint == HINT_LAST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
template <int HINT>
__device__ __forceinline__ double ld_hint(const double* p, uint64_t pol)
{
    if constexpr (HINT == HINT_NONE) {
        return *p;
    } else {
        double v;
        asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
        return v;
    }
}
template <int HINT>
__device__ __forceinline__ void st_hint(double* p, double v, uint64_t pol)
{
    if constexpr (HINT == HINT_NONE) {
        *p = v;
    } else {
        asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
    }
}

template <bool PULL, bool FAST, int HINT = HINT_NONE, typename F>
__device__ __forceinline__ void gather27(const double* __restrict__ in, long long c, uint32_t m, const Layout& L,
                                         const PullOffsets& o, F&& sink, uint64_t pol = 0)
{
    static_for<0, NQ>([&](auto qc) {
        constexpr int Q = decltype(qc)::value;
        double v;
        if constexpr (!PULL) {
            v = ld_hint<HINT>(in + (long long)Q * L.sq + c, pol);
        } else if constexpr (FAST) {
            v = ld_hint<HINT>(in + (long long)Q * L.sq + c + (o.xo[ex(Q) + 1] + o.yo[ey(Q) + 1] + o.zo[ez(Q) + 1]), pol);
        } else {
            // fluid source: take its population; solid source: halfway bounce-back of the
            // cell's own opposite population (LBM.cpp:590-595 in pull form)
            const bool fl = (m >> Q) & 1u;
            const long long a = (long long)Q * L.sq + c + (o.xo[ex(Q) + 1] + o.yo[ey(Q) + 1] + o.zo[ez(Q) + 1]);
            const long long b = (long long)opp(Q) * L.sq + c;
            v = ld_hint<HINT>(in + (fl ? a : b), pol);
        }
        sink(qc, v);
    });
}

// ---------------------------------------------------------------------------
// pass 1: q-corrections of the post-stream state of one cell
// ---------------------------------------------------------------------------
template <bool PULL, int HINT = HINT_NONE>
__device__ __forceinline__ void qcorr_cell(const double* __restrict__ fin, const double* __restrict__ gin,
                                           const uint32_t* __restrict__ nbr, double* __restrict__ qc, const Layout& L,
                                           const Phys& P, int i, int j, int k, uint64_t pol = 0)
{
    const long long c = L.cell(i, j, k);
    const long long n = L.sq;
    // 40 registers -> 12 CTAs per SM: here occupancy hides the mask -> pull dependency (ncu: 6.5 TB/s)
    const uint32_t m = nbr[c];
    if (!(m & 1u)) return;
    const PullOffsets o = pull_offsets(L, i, j, k);
    MomL ml = {0.0, 0.0, 0.0, 0.0};
    double e2 = 0.0;
    if (__all_sync(__activemask(), m == ALL_FLUID)) {
        gather27<PULL, true, HINT>(fin, c, m, L, o, [&](auto qc_, double v) { acc_l<decltype(qc_)::value>(ml, v); }, pol);
        gather27<PULL, true, HINT>(gin, c, m, L, o, [&](auto, double v) { e2 += v; }, pol);
    } else {
        gather27<PULL, false, HINT>(fin, c, m, L, o, [&](auto qc_, double v) { acc_l<decltype(qc_)::value>(ml, v); }, pol);
        gather27<PULL, false, HINT>(gin, c, m, L, o, [&](auto, double v) { e2 += v; }, pol);
    }
    const Prim s = primitives(ml.rho, ml.jx, ml.jy, ml.jz, e2, P);
    qc[c] = s.qcx;
    qc[n + c] = s.qcy;
    qc[2 * n + c] = s.qcz;
}

template <bool PULL>
__global__ void __launch_bounds__(128) k_qcorr(const double* __restrict__ fin, const double* __restrict__ gin,
                                               const uint32_t* __restrict__ nbr, double* __restrict__ qc,
                                               const __grid_constant__ Layout L, const __grid_constant__ Phys P, int k0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.nx) return;
    qcorr_cell<PULL>(fin, gin, nbr, qc, L, P, i, blockIdx.y, blockIdx.z + k0);
}

// ---------------------------------------------------------------------------
// pass 2: (pull +) collide of one cell
// ---------------------------------------------------------------------------
template <bool PULL, bool MACRO, int HINT = HINT_NONE>
__device__ __forceinline__ void collide_cell(const double* __restrict__ fin, const double* __restrict__ gin,
                                             double* __restrict__ fout, double* __restrict__ gout,
                                             const uint32_t* __restrict__ nbr, const uint8_t* __restrict__ flag,
                                             const double* __restrict__ qc, double* __restrict__ macro, const Layout& L,
                                             const Phys& P, int i, int j, int k, uint64_t pol = 0)
{
    const long long c = L.cell(i, j, k);
    const long long n = L.sq;
    // Every load of the cell is issued before anything is waited for (ncu showed four dependent global
    // latencies per cell -- mask, populations, flag byte, neighbour q-corrections -- costing 64 % of the
    // kernel): the mask, the flag byte, the six neighbour q-corrections (always inside the padded box;
    // unusable ones are discarded by the flag bits afterwards) and, speculatively, the 54 all-fluid pulls;
    // directions whose source turns out to be solid are patched.
    const uint32_t m = nbr[c];
    const unsigned fb = flag[c];
    const double qxp = qc[c + 1], qxm = qc[c - 1];
    const double qyp = qc[n + c + L.px], qym = qc[n + c - L.px];
    const double qzp = qc[2 * n + c + L.sz], qzm = qc[2 * n + c - L.sz];
    const PullOffsets o = pull_offsets(L, i, j, k);
    double f[NQ], g[NQ];
    gather27<PULL, true, HINT>(fin, c, ALL_FLUID, L, o, [&](auto qc_, double v) { f[decltype(qc_)::value] = v; }, pol);
    gather27<PULL, true, HINT>(gin, c, ALL_FLUID, L, o, [&](auto qc_, double v) { g[decltype(qc_)::value] = v; }, pol);
    if (!(m & 1u)) {
        // solid cell: the streamed value is the -1 sentinel (LBM.cpp:565, 582); collide skips it
        if constexpr (PULL) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                fout[q * n + c] = -1.0;
                gout[q * n + c] = -1.0;
            }
        }
        return;
    }
    if constexpr (PULL) {
        if (m != ALL_FLUID) {
            // halfway bounce-back: the cell's own opposite population (LBM.cpp:590-595 in pull form)
            static_for<1, NQ>([&](auto qc_) {
                constexpr int Q = decltype(qc_)::value;
                if (!((m >> Q) & 1u)) {
                    f[Q] = fin[(long long)opp(Q) * n + c];
                    g[Q] = gin[(long long)opp(Q) * n + c];
                }
            });
        }
    }
    const MomF mf = moments_f([&](int q) { return f[q]; });
    const MomG mg = moments_g([&](int q) { return g[q]; });
    const Prim s = primitives(mf.rho, mf.jx, mf.jy, mf.jz, mg.e2, P);

    // grad of the q-correction (LBM.cpp:959-991, Utilities.H:279-312)
    const double dqx = one_sided_gradient(fb & GRAD_PX, fb & GRAD_MX, (fb & GRAD_PX) ? qxp : 0.0, s.qcx,
                                          (fb & GRAD_MX) ? qxm : 0.0, P.idx[0]);
    const double dqy = one_sided_gradient(fb & GRAD_PY, fb & GRAD_MY, (fb & GRAD_PY) ? qyp : 0.0, s.qcy,
                                          (fb & GRAD_MY) ? qym : 0.0, P.idx[1]);
    const double dqz = one_sided_gradient(fb & GRAD_PZ, fb & GRAD_MZ, (fb & GRAD_PZ) ? qzp : 0.0, s.qcz,
                                          (fb & GRAD_MZ) ? qzm : 0.0, P.idx[2]);

    if constexpr (MACRO) {
        // m_macrodata of the post-stream state (Constants.H:8-31, LBM.cpp:867-901)
        macro[0 * n + c] = s.rho;
        macro[1 * n + c] = s.u;
        macro[2 * n + c] = s.v;
        macro[3 * n + c] = s.w;
        macro[4 * n + c] = sqrt(s.u * s.u + s.v * s.v + s.w * s.w);
        macro[5 * n + c] = mg.e2;
        macro[6 * n + c] = s.qcx;
        macro[7 * n + c] = s.qcy;
        macro[8 * n + c] = s.qcz;
        macro[9 * n + c] = mf.pxx;
        macro[10 * n + c] = mf.pyy;
        macro[11 * n + c] = mf.pzz;
        macro[12 * n + c] = mf.pxy;
        macro[13 * n + c] = mf.pxz;
        macro[14 * n + c] = mf.pyz;
        macro[15 * n + c] = mg.qx;
        macro[16 * n + c] = mg.qy;
        macro[17 * n + c] = mg.qz;
        macro[18 * n + c] = s.T;
        // m_derived (Constants.H:39-47) lives behind the 19 macro comps: 19..22 vorticity, 23..25 dQCorr
        macro[23 * n + c] = dqx;
        macro[24 * n + c] = dqy;
        macro[25 * n + c] = dqz;
    }

    const Coll cc = collision_coefficients(s, mf, mg, dqx, dqy, dqz, P);
    // relax_f_to_equilibrium (LBM.cpp:799-801)
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        st_hint<HINT>(fout + (long long)Q * n + c, f[Q] + cc.omega * (feq_q<Q>(cc) - f[Q]), pol);
    });
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        st_hint<HINT>(gout + (long long)Q * n + c, g[Q] + cc.omega * (geq_q<Q>(cc) - g[Q]), pol);
    });
}

template <bool PULL, bool MACRO>
__global__ void __launch_bounds__(128, 3) k_collide(const double* __restrict__ fin, const double* __restrict__ gin,
                                                 double* __restrict__ fout, double* __restrict__ gout,
                                                 const uint32_t* __restrict__ nbr, const uint8_t* __restrict__ flag,
                                                 const double* __restrict__ qc, double* __restrict__ macro,
                                                 const __grid_constant__ Layout L, const __grid_constant__ Phys P, int k0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.nx) return;
    collide_cell<PULL, MACRO>(fin, gin, fout, gout, nbr, flag, qc, macro, L, P, i, blockIdx.y, blockIdx.z + k0);
}

// ---------------------------------------------------------------------------
// "carry" step: the collide kernel of step n also emits the conserved moments of the POST-STREAM state of
// step n+1, so that the q-correction pass (the second touch of all 54 populations) disappears.
//
// rho, rho*u and 2rhoE of the next post-stream state at cell d are sums over the 27 cells d - e_q of this
// step's post-collision populations: M(d) = sum_q phi_q f*_q(d - e_q).  The sum is separable on the
// tensor-product lattice, so it is reduced in three stages without ever re-reading a population:
//   x  inside the warp: lane l holds cell i0 + l of a row, the e_x = +-1 populations are shifted one lane
//      by shuffles.  Warps overlap by `halo` cells on each side (those lanes collide redundantly and
//      store nothing), so no partial sums cross a warp edge.
//   y  by marching: a thread walks its column through KY rows of one plane (plus one redundant row at
//      either end) and keeps the sums destined for rows j-1 and j in a private shared-memory ring; row j-1
//      is complete once row j has been collided.  Rows, not planes, are marched so that the CTAs resident at
//      any time cover a few whole planes: marching through planes spreads them over every plane of a chunk
//      (thousands of 2 MB pages and DRAM rows in flight; measured 40 % slower).
//   z  through memory: the thread of source plane k writes, per destination plane k+c (c = -1,0,1), the four
//      words (rho, jx, jy, 2rhoE) -- jz is c * rho.  k_qcorr_combine adds the three planes of a cell.
// Cells whose 27 pull sources are not all collided cells of this box (EB neighbours, non-wrapped box
// faces, ghost planes owned by another rank) are not served by the carried sums: k_qcorr_combine pulls
// their populations itself, exactly as k_qcorr does.
// Real traffic per cell: collide 54 + 54 + 12 words, combine 12 + 3 words (+ masks) = 1085 B against
// 1360 B of the two-pass step.
// ---------------------------------------------------------------------------
static_assert(CARRY_WORDS == 12, "part layout: [c = -1,0,1][rho, jx, jy, e2]");

__device__ __forceinline__ void cp_async8(unsigned smem_addr, const void* gptr)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all_after(double& dep) { asm volatile("cp.async.wait_all;" : "+d"(dep)::"memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---------------------------------------------------------------------------
// k_collide with the two register savers of the carry kernels: g travels global -> shared by cp.async (a private
// 27-slot column per thread) instead of occupying 54 registers, and every access is a parameter-space base pointer
// plus an unsigned 32-bit byte offset.  Same arithmetic in the same order as collide_cell<true, false>: results
// are bit-identical to k_collide.
// ---------------------------------------------------------------------------
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
    k_collide_lean(const __grid_constant__ CarryPtrs A, const uint32_t* __restrict__ nbr,
                   const uint8_t* __restrict__ flag, const __grid_constant__ Layout L, const __grid_constant__ Phys P,
                   int k0)
{
    constexpr int T = 128;
    extern __shared__ double smem[];
    double* const sg = smem + threadIdx.x;
    const unsigned sg_addr = (unsigned)__cvta_generic_to_shared(sg);
    const int i = blockIdx.x * T + threadIdx.x;
    if (i >= L.nx) return;
    const int j = blockIdx.y, k = blockIdx.z + k0;
    const unsigned px8 = (unsigned)L.px * 8u, sz8 = (unsigned)L.sz * 8u;
    unsigned xo[3], yo[3], zo[3];
    xo[1] = yo[1] = zo[1] = 0u;
    xo[2] = (L.wrap[0] && i == 0) ? (unsigned)(L.nx - 1) * 8u : 0u - 8u;
    xo[0] = (L.wrap[0] && i == L.nx - 1) ? 0u - (unsigned)(L.nx - 1) * 8u : 8u;
    yo[2] = (L.wrap[1] && j == 0) ? (unsigned)(L.ny - 1) * px8 : 0u - px8;
    yo[0] = (L.wrap[1] && j == L.ny - 1) ? 0u - (unsigned)(L.ny - 1) * px8 : px8;
    zo[2] = (L.wrap[2] && k == 0) ? (unsigned)(L.nz - 1) * sz8 : 0u - sz8;
    zo[0] = (L.wrap[2] && k == L.nz - 1) ? 0u - (unsigned)(L.nz - 1) * sz8 : sz8;
    const unsigned c = (unsigned)(i + OX) * 8u + (unsigned)(j + GY) * px8 + (unsigned)(k + GZ) * sz8;
    auto ldb = [](const double* base, unsigned off) { return *(const double*)((const char*)base + off); };
    auto stb = [](double* base, unsigned off, double v) { *(double*)((char*)base + off) = v; };
    unsigned cyz[3][3];
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int d = 0; d < 3; ++d) cyz[b][d] = c + yo[b] + zo[d];
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        cp_async8(sg_addr + Q * T * 8, (const char*)A.gin[Q] + (cyz[ey(Q) + 1][ez(Q) + 1] + xo[ex(Q) + 1]));
    });
    const uint32_t m = *(const uint32_t*)((const char*)nbr + (c >> 1));
    const unsigned fb = flag[c >> 3];
    const double qxp = ldb(A.qc[0], c + 8u), qxm = ldb(A.qc[0], c - 8u);
    const double qyp = ldb(A.qc[1], c + px8), qym = ldb(A.qc[1], c - px8);
    const double qzp = ldb(A.qc[2], c + sz8), qzm = ldb(A.qc[2], c - sz8);
    double f[NQ];
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        f[Q] = ldb(A.fin[Q], cyz[ey(Q) + 1][ez(Q) + 1] + xo[ex(Q) + 1]);
    });
    const bool fluid = m & 1u;
    cp_async_wait_all();
    if (m != ALL_FLUID) {
        if (fluid) {
            // halfway bounce-back: the cell's own opposite population (LBM.cpp:590-595 in pull form)
            static_for<1, NQ>([&](auto qc_) {
                constexpr int Q = decltype(qc_)::value;
                if (!((m >> Q) & 1u)) {
                    f[Q] = ldb(A.fin[opp(Q)], c);
                    sg[Q * T] = ldb(A.gin[opp(Q)], c);
                }
            });
        } else {
            // solid cell: the streamed value is the -1 sentinel (LBM.cpp:565, 582) and collide skips it;
            // with omega = 0 below the "relaxed" value is exactly -1 again
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                f[q] = -1.0;
                sg[q * T] = -1.0;
            }
        }
    }
    const MomF mf = moments_f([&](int q) { return f[q]; });
    const MomG mg = moments_g([&](int q) { return sg[q * T]; });
    const Prim s = primitives(mf.rho, mf.jx, mf.jy, mf.jz, mg.e2, P);
    const double dqx = one_sided_gradient(fb & GRAD_PX, fb & GRAD_MX, (fb & GRAD_PX) ? qxp : 0.0, s.qcx,
                                          (fb & GRAD_MX) ? qxm : 0.0, P.idx[0]);
    const double dqy = one_sided_gradient(fb & GRAD_PY, fb & GRAD_MY, (fb & GRAD_PY) ? qyp : 0.0, s.qcy,
                                          (fb & GRAD_MY) ? qym : 0.0, P.idx[1]);
    const double dqz = one_sided_gradient(fb & GRAD_PZ, fb & GRAD_MZ, (fb & GRAD_PZ) ? qzp : 0.0, s.qcz,
                                          (fb & GRAD_MZ) ? qzm : 0.0, P.idx[2]);
    const Coll cc = collision_coefficients(s, mf, mg, dqx, dqy, dqz, P);
    const double omega = fluid ? cc.omega : 0.0;
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        stb(A.fout[Q], c, f[Q] + omega * (feq_q<Q>(cc) - f[Q]));
        asm volatile("" ::: "memory");  // keep the relax / store pairs in order: fewer values live at once
    });
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        const double gq = sg[Q * T];
        stb(A.gout[Q], c, gq + omega * (geq_q<Q>(cc) - gq));
        asm volatile("" ::: "memory");
    });
}

// ---------------------------------------------------------------------------
// The same carried sums without marching: a CTA is W warps = W consecutive rows of one warp-wide column strip,
// one cell per thread, so CTAs sweep the box in launch order exactly like k_collide (the DRAM streams stay
// compact; marching spreads the resident CTAs over rows that are KY rows apart).  x is reduced by shuffles as
// above, y by one exchange through shared memory between the warps of the CTA.  What the first row of a CTA
// sends down and its last row sends up leaves the CTA: those 2 x 9 words go to the compact `edge` arrays
// (one entry per CTA row, 18 / W words per cell) and k_qcorr_combine adds them to the two rows concerned, so no
// row is collided twice.  g again travels global -> shared by cp.async, and the exchange reuses those slots.
// ---------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(32 * W, (W <= 4 ? 3 : W <= 8 ? 2 : 1))
    k_collide_tile(const __grid_constant__ CarryPtrs A, const uint32_t* __restrict__ nbr,
                   const uint8_t* __restrict__ flag, const __grid_constant__ Layout L, const __grid_constant__ Phys P,
                   const __grid_constant__ CarryPlan C, int k0)
{
    constexpr int T = 32 * W;
    extern __shared__ double smem[];
    double* const sg = smem + threadIdx.x;  // sg[slot * T]
    const unsigned sg_addr = (unsigned)__cvta_generic_to_shared(sg);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int xc = blockIdx.x;
    const int y0 = blockIdx.y * W;
    const int rows = min(W, L.ny - y0);  // rows of this CTA inside the box
    const int k = blockIdx.z + k0;
    const int i = xc * C.own - C.halo + lane;
    const bool own = lane >= C.halo && lane < C.halo + C.own && i < L.nx && w < rows;
    // cell this thread collides: its own, the periodic image for halo lanes over a wrapped edge, otherwise
    // clamped (what such a thread contributes lands in cells k_qcorr_combine does not take from the carried sums)
    int is = i, j = min(y0 + w, L.ny - 1);
    if (is < 0) is = L.wrap[0] ? is + L.nx : 0;
    if (is >= L.nx) is = (L.wrap[0] && is - L.nx < L.nx) ? is - L.nx : L.nx - 1;
    const unsigned FULL = 0xffffffffu;
    const unsigned px8 = (unsigned)L.px * 8u, sz8 = (unsigned)L.sz * 8u;
    unsigned xo[3], yo[3], zo[3];
    xo[1] = yo[1] = zo[1] = 0u;
    xo[2] = (L.wrap[0] && is == 0) ? (unsigned)(L.nx - 1) * 8u : 0u - 8u;
    xo[0] = (L.wrap[0] && is == L.nx - 1) ? 0u - (unsigned)(L.nx - 1) * 8u : 8u;
    yo[2] = (L.wrap[1] && j == 0) ? (unsigned)(L.ny - 1) * px8 : 0u - px8;
    yo[0] = (L.wrap[1] && j == L.ny - 1) ? 0u - (unsigned)(L.ny - 1) * px8 : px8;
    zo[2] = (L.wrap[2] && k == 0) ? (unsigned)(L.nz - 1) * sz8 : 0u - sz8;
    zo[0] = (L.wrap[2] && k == L.nz - 1) ? 0u - (unsigned)(L.nz - 1) * sz8 : sz8;
    const unsigned c = (unsigned)(is + OX) * 8u + (unsigned)(j + GY) * px8 + (unsigned)(k + GZ) * sz8;
    auto ldb = [](const double* base, unsigned off) { return *(const double*)((const char*)base + off); };
    auto stb = [](double* base, unsigned off, double v) { *(double*)((char*)base + off) = v; };
    unsigned cyz[3][3];
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int d = 0; d < 3; ++d) cyz[b][d] = c + yo[b] + zo[d];
    // g: global -> shared, asynchronously, no registers
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        cp_async8(sg_addr + Q * T * 8, (const char*)A.gin[Q] + (cyz[ey(Q) + 1][ez(Q) + 1] + xo[ex(Q) + 1]));
    });
    const uint32_t m = *(const uint32_t*)((const char*)nbr + (c >> 1));
    const unsigned fb = flag[c >> 3];
    const double qxp = ldb(A.qc[0], c + 8u), qxm = ldb(A.qc[0], c - 8u);
    const double qyp = ldb(A.qc[1], c + px8), qym = ldb(A.qc[1], c - px8);
    const double qzp = ldb(A.qc[2], c + sz8), qzm = ldb(A.qc[2], c - sz8);
    double f[NQ];
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        f[Q] = ldb(A.fin[Q], cyz[ey(Q) + 1][ez(Q) + 1] + xo[ex(Q) + 1]);
    });
    const bool fluid = m & 1u;
    if (m != ALL_FLUID) {
        if (fluid) {
            // halfway bounce-back: the cell's own opposite population (LBM.cpp:590-595 in pull form)
            static_for<1, NQ>([&](auto qc_) {
                constexpr int Q = decltype(qc_)::value;
                if (!((m >> Q) & 1u)) f[Q] = ldb(A.fin[opp(Q)], c);
            });
        } else {
            // solid cell: the streamed value is the -1 sentinel (LBM.cpp:565, 582) and collide skips it;
            // with omega = 0 below the "relaxed" value is exactly -1 again
#pragma unroll
            for (int q = 0; q < NQ; ++q) f[q] = -1.0;
        }
    }
    MomF mf = moments_f([&](int q) { return f[q]; });
    // wait for g only now, and tied to a value that needs every f: placed right after the cp.async issue, ptxas
    // schedules the f loads behind the wait and the cell pays two DRAM latencies in series
    cp_async_wait_all_after(mf.rho);
    if (m != ALL_FLUID) {
        if (fluid) {
            static_for<1, NQ>([&](auto qc_) {
                constexpr int Q = decltype(qc_)::value;
                if (!((m >> Q) & 1u)) sg[Q * T] = ldb(A.gin[opp(Q)], c);
            });
        } else {
#pragma unroll
            for (int q = 0; q < NQ; ++q) sg[q * T] = -1.0;
        }
    }
    const MomG mg = moments_g([&](int q) { return sg[q * T]; });
    const Prim s = primitives(mf.rho, mf.jx, mf.jy, mf.jz, mg.e2, P);
    const double dqx = one_sided_gradient(fb & GRAD_PX, fb & GRAD_MX, (fb & GRAD_PX) ? qxp : 0.0, s.qcx,
                                          (fb & GRAD_MX) ? qxm : 0.0, P.idx[0]);
    const double dqy = one_sided_gradient(fb & GRAD_PY, fb & GRAD_MY, (fb & GRAD_PY) ? qyp : 0.0, s.qcy,
                                          (fb & GRAD_MY) ? qym : 0.0, P.idx[1]);
    const double dqz = one_sided_gradient(fb & GRAD_PZ, fb & GRAD_MZ, (fb & GRAD_PZ) ? qzp : 0.0, s.qcz,
                                          (fb & GRAD_MZ) ? qzm : 0.0, P.idx[2]);
    const Coll cc = collision_coefficients(s, mf, mg, dqx, dqy, dqz, P);
    const double omega = fluid ? cc.omega : 0.0;
    // this cell's contributions to rows j-1 (TA), j (TB), j+1 (TC): [plane c][rho, jx], then e2
    double TA[3][2] = {}, TB[3][2] = {}, TC[3][2] = {};
    double EA[3] = {}, EB[3] = {}, EC[3] = {};
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        constexpr int d = ez(Q) + 1;
        const double v = f[Q] + omega * (feq_q<Q>(cc) - f[Q]);
        if (own) stb(A.fout[Q], c, v);
        double t = v;
        if constexpr (ex(Q) == 1) t = __shfl_up_sync(FULL, v, 1);
        if constexpr (ex(Q) == -1) t = __shfl_down_sync(FULL, v, 1);
        double(&Tt)[3][2] = ey(Q) == -1 ? TA : ey(Q) == 0 ? TB : TC;
        Tt[d][0] += t;
        if constexpr (ex(Q) == 1) Tt[d][1] += t;
        if constexpr (ex(Q) == -1) Tt[d][1] -= t;
    });
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        constexpr int d = ez(Q) + 1;
        const double gq = sg[Q * T];
        const double v = gq + omega * (geq_q<Q>(cc) - gq);
        if (own) stb(A.gout[Q], c, v);
        double t = v;
        if constexpr (ex(Q) == 1) t = __shfl_up_sync(FULL, v, 1);
        if constexpr (ex(Q) == -1) t = __shfl_down_sync(FULL, v, 1);
        double(&E)[3] = ey(Q) == -1 ? EA : ey(Q) == 0 ? EB : EC;
        E[d] += t;
    });
    // y exchange: what this row sends down (TA, EA: slots 0..8) and up (TC, EC: slots 9..17), in the g slots;
    // the first row's "down" and the last row's "up" leave the CTA through the edge arrays
    const unsigned ce = (unsigned)(i + OX) * 8u + (unsigned)blockIdx.y * px8 + (unsigned)(k + GZ) * (unsigned)C.esz8;
    const bool first = w == 0, last = w == rows - 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (first) {
            if (own) {
                stb(A.edge[3 * d + 0], ce, TA[d][0]);
                stb(A.edge[3 * d + 1], ce, TA[d][1]);
                stb(A.edge[3 * d + 2], ce, EA[d]);
            }
        } else {
            sg[(3 * d + 0) * T] = TA[d][0];
            sg[(3 * d + 1) * T] = TA[d][1];
            sg[(3 * d + 2) * T] = EA[d];
        }
        if (last) {
            if (own) {
                stb(A.edge[9 + 3 * d + 0], ce, TC[d][0]);
                stb(A.edge[9 + 3 * d + 1], ce, TC[d][1]);
                stb(A.edge[9 + 3 * d + 2], ce, EC[d]);
            }
        } else {
            sg[(9 + 3 * d + 0) * T] = TC[d][0];
            sg[(9 + 3 * d + 1) * T] = TC[d][1];
            sg[(9 + 3 * d + 2) * T] = EC[d];
        }
    }
    __syncthreads();
    if (own) {
        const double* up = sg + 32;   // row j+1: its e_y = -1 terms arrive here
        const double* dn = sg - 32;   // row j-1: its e_y = +1 terms
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double ua = last ? 0.0 : up[(3 * d + 0) * T], dc = first ? 0.0 : dn[(9 + 3 * d + 0) * T];
            const double ux = last ? 0.0 : up[(3 * d + 1) * T], dx = first ? 0.0 : dn[(9 + 3 * d + 1) * T];
            const double ue = last ? 0.0 : up[(3 * d + 2) * T], de = first ? 0.0 : dn[(9 + 3 * d + 2) * T];
            stb(A.part[4 * d + 0], c, TB[d][0] + ua + dc);
            stb(A.part[4 * d + 1], c, TB[d][1] + ux + dx);
            stb(A.part[4 * d + 2], c, dc - ua);
            stb(A.part[4 * d + 3], c, EB[d] + ue + de);
        }
    }
}

// ---------------------------------------------------------------------------
// k_collide_tile with the z sum of a PAIR of planes (2b, 2b+1) completed on chip: a thread collides its cell in
// plane 2b, then in plane 2b+1, and keeps what the first sends up / receives from the second in shared memory.
// Per cell 9 words leave instead of 12: own[rho, jx, jy, jz, e2] (its own and its partner's contributions) and
// out[rho, jx, jy, e2] (even plane: what it sends down; odd plane: what it sends up).  Needs an even number of
// planes; k_qcorr_combine_pair adds own(k) and out(k -+ 1).
// ---------------------------------------------------------------------------
constexpr int PAIR_KEEP_SLOTS = 8;
template <int W>
__global__ void __launch_bounds__(32 * W, (W <= 4 ? 3 : W <= 8 ? 2 : 1))
    k_collide_tile_pair(const __grid_constant__ CarryPtrs A, const uint32_t* __restrict__ nbr,
                   const uint8_t* __restrict__ flag, const __grid_constant__ Layout L, const __grid_constant__ Phys P,
                   const __grid_constant__ CarryPlan C, int k0)
{
    constexpr int T = 32 * W;
    extern __shared__ double smem[];
    double* const sg = smem + threadIdx.x;  // sg[slot * T]
    const unsigned sg_addr = (unsigned)__cvta_generic_to_shared(sg);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int xc = blockIdx.x;
    const int y0 = blockIdx.y * W;
    const int rows = min(W, L.ny - y0);  // rows of this CTA inside the box
    double* const sk = smem + NQ * T + threadIdx.x;  // sk[slot * T]: own and up sums of the first plane
    const int i = xc * C.own - C.halo + lane;
    const bool own = lane >= C.halo && lane < C.halo + C.own && i < L.nx && w < rows;
    // cell this thread collides: its own, the periodic image for halo lanes over a wrapped edge, otherwise
    // clamped (what such a thread contributes lands in cells k_qcorr_combine does not take from the carried sums)
    int is = i, j = min(y0 + w, L.ny - 1);
    if (is < 0) is = L.wrap[0] ? is + L.nx : 0;
    if (is >= L.nx) is = (L.wrap[0] && is - L.nx < L.nx) ? is - L.nx : L.nx - 1;
    const unsigned FULL = 0xffffffffu;
    const unsigned px8 = (unsigned)L.px * 8u, sz8 = (unsigned)L.sz * 8u;
    unsigned xo[3], yo[3], zo[3];
    xo[1] = yo[1] = zo[1] = 0u;
    xo[2] = (L.wrap[0] && is == 0) ? (unsigned)(L.nx - 1) * 8u : 0u - 8u;
    xo[0] = (L.wrap[0] && is == L.nx - 1) ? 0u - (unsigned)(L.nx - 1) * 8u : 8u;
    yo[2] = (L.wrap[1] && j == 0) ? (unsigned)(L.ny - 1) * px8 : 0u - px8;
    yo[0] = (L.wrap[1] && j == L.ny - 1) ? 0u - (unsigned)(L.ny - 1) * px8 : px8;
    unsigned c_first = 0u;
#pragma unroll 1
    for (int kk = 0; kk < 2; ++kk) {
    const int k = 2 * (int)blockIdx.z + kk + k0;
    zo[2] = (L.wrap[2] && k == 0) ? (unsigned)(L.nz - 1) * sz8 : 0u - sz8;
    zo[0] = (L.wrap[2] && k == L.nz - 1) ? 0u - (unsigned)(L.nz - 1) * sz8 : sz8;
    const unsigned c = (unsigned)(is + OX) * 8u + (unsigned)(j + GY) * px8 + (unsigned)(k + GZ) * sz8;
    auto ldb = [](const double* base, unsigned off) { return *(const double*)((const char*)base + off); };
    auto stb = [](double* base, unsigned off, double v) { *(double*)((char*)base + off) = v; };
    unsigned cyz[3][3];
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int d = 0; d < 3; ++d) cyz[b][d] = c + yo[b] + zo[d];
    // g: global -> shared, asynchronously, no registers
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        cp_async8(sg_addr + Q * T * 8, (const char*)A.gin[Q] + (cyz[ey(Q) + 1][ez(Q) + 1] + xo[ex(Q) + 1]));
    });
    const uint32_t m = *(const uint32_t*)((const char*)nbr + (c >> 1));
    const unsigned fb = flag[c >> 3];
    const double qxp = ldb(A.qc[0], c + 8u), qxm = ldb(A.qc[0], c - 8u);
    const double qyp = ldb(A.qc[1], c + px8), qym = ldb(A.qc[1], c - px8);
    const double qzp = ldb(A.qc[2], c + sz8), qzm = ldb(A.qc[2], c - sz8);
    double f[NQ];
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        f[Q] = ldb(A.fin[Q], cyz[ey(Q) + 1][ez(Q) + 1] + xo[ex(Q) + 1]);
    });
    const bool fluid = m & 1u;
    if (m != ALL_FLUID) {
        if (fluid) {
            // halfway bounce-back: the cell's own opposite population (LBM.cpp:590-595 in pull form)
            static_for<1, NQ>([&](auto qc_) {
                constexpr int Q = decltype(qc_)::value;
                if (!((m >> Q) & 1u)) f[Q] = ldb(A.fin[opp(Q)], c);
            });
        } else {
            // solid cell: the streamed value is the -1 sentinel (LBM.cpp:565, 582) and collide skips it;
            // with omega = 0 below the "relaxed" value is exactly -1 again
#pragma unroll
            for (int q = 0; q < NQ; ++q) f[q] = -1.0;
        }
    }
    MomF mf = moments_f([&](int q) { return f[q]; });
    // wait for g only now, and tied to a value that needs every f: placed right after the cp.async issue, ptxas
    // schedules the f loads behind the wait and the cell pays two DRAM latencies in series
    cp_async_wait_all_after(mf.rho);
    if (m != ALL_FLUID) {
        if (fluid) {
            static_for<1, NQ>([&](auto qc_) {
                constexpr int Q = decltype(qc_)::value;
                if (!((m >> Q) & 1u)) sg[Q * T] = ldb(A.gin[opp(Q)], c);
            });
        } else {
#pragma unroll
            for (int q = 0; q < NQ; ++q) sg[q * T] = -1.0;
        }
    }
    const MomG mg = moments_g([&](int q) { return sg[q * T]; });
    const Prim s = primitives(mf.rho, mf.jx, mf.jy, mf.jz, mg.e2, P);
    const double dqx = one_sided_gradient(fb & GRAD_PX, fb & GRAD_MX, (fb & GRAD_PX) ? qxp : 0.0, s.qcx,
                                          (fb & GRAD_MX) ? qxm : 0.0, P.idx[0]);
    const double dqy = one_sided_gradient(fb & GRAD_PY, fb & GRAD_MY, (fb & GRAD_PY) ? qyp : 0.0, s.qcy,
                                          (fb & GRAD_MY) ? qym : 0.0, P.idx[1]);
    const double dqz = one_sided_gradient(fb & GRAD_PZ, fb & GRAD_MZ, (fb & GRAD_PZ) ? qzp : 0.0, s.qcz,
                                          (fb & GRAD_MZ) ? qzm : 0.0, P.idx[2]);
    const Coll cc = collision_coefficients(s, mf, mg, dqx, dqy, dqz, P);
    const double omega = fluid ? cc.omega : 0.0;
    // this cell's contributions to rows j-1 (TA), j (TB), j+1 (TC): [plane c][rho, jx], then e2
    double TA[3][2] = {}, TB[3][2] = {}, TC[3][2] = {};
    double EA[3] = {}, EB[3] = {}, EC[3] = {};
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        constexpr int d = ez(Q) + 1;
        const double v = f[Q] + omega * (feq_q<Q>(cc) - f[Q]);
        if (own) stb(A.fout[Q], c, v);
        double t = v;
        if constexpr (ex(Q) == 1) t = __shfl_up_sync(FULL, v, 1);
        if constexpr (ex(Q) == -1) t = __shfl_down_sync(FULL, v, 1);
        double(&Tt)[3][2] = ey(Q) == -1 ? TA : ey(Q) == 0 ? TB : TC;
        Tt[d][0] += t;
        if constexpr (ex(Q) == 1) Tt[d][1] += t;
        if constexpr (ex(Q) == -1) Tt[d][1] -= t;
    });
    static_for<0, NQ>([&](auto qc_) {
        constexpr int Q = decltype(qc_)::value;
        constexpr int d = ez(Q) + 1;
        const double gq = sg[Q * T];
        const double v = gq + omega * (geq_q<Q>(cc) - gq);
        if (own) stb(A.gout[Q], c, v);
        double t = v;
        if constexpr (ex(Q) == 1) t = __shfl_up_sync(FULL, v, 1);
        if constexpr (ex(Q) == -1) t = __shfl_down_sync(FULL, v, 1);
        double(&E)[3] = ey(Q) == -1 ? EA : ey(Q) == 0 ? EB : EC;
        E[d] += t;
    });
    // y exchange: what this row sends down (TA, EA: slots 0..8) and up (TC, EC: slots 9..17), in the g slots;
    // the first row's "down" and the last row's "up" leave the CTA through the edge arrays
    const unsigned ce = (unsigned)(i + OX) * 8u + (unsigned)blockIdx.y * px8 + (unsigned)(k + GZ) * (unsigned)C.esz8;
    const bool first = w == 0, last = w == rows - 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (first) {
            if (own) {
                stb(A.edge[3 * d + 0], ce, TA[d][0]);
                stb(A.edge[3 * d + 1], ce, TA[d][1]);
                stb(A.edge[3 * d + 2], ce, EA[d]);
            }
        } else {
            sg[(3 * d + 0) * T] = TA[d][0];
            sg[(3 * d + 1) * T] = TA[d][1];
            sg[(3 * d + 2) * T] = EA[d];
        }
        if (last) {
            if (own) {
                stb(A.edge[9 + 3 * d + 0], ce, TC[d][0]);
                stb(A.edge[9 + 3 * d + 1], ce, TC[d][1]);
                stb(A.edge[9 + 3 * d + 2], ce, EC[d]);
            }
        } else {
            sg[(9 + 3 * d + 0) * T] = TC[d][0];
            sg[(9 + 3 * d + 1) * T] = TC[d][1];
            sg[(9 + 3 * d + 2) * T] = EC[d];
        }
    }
    __syncthreads();
    if (own) {
        const double* up = sg + 32;   // row j+1: its e_y = -1 terms arrive here
        const double* dn = sg - 32;   // row j-1: its e_y = +1 terms
        double r[3], x[3], y[3], e[3];  // per destination plane k-1, k, k+1: rho, jx, jy, e2 (y-complete up to the edges)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double ua = last ? 0.0 : up[(3 * d + 0) * T], dc = first ? 0.0 : dn[(9 + 3 * d + 0) * T];
            const double ux = last ? 0.0 : up[(3 * d + 1) * T], dx = first ? 0.0 : dn[(9 + 3 * d + 1) * T];
            const double ue = last ? 0.0 : up[(3 * d + 2) * T], de = first ? 0.0 : dn[(9 + 3 * d + 2) * T];
            r[d] = TB[d][0] + ua + dc;
            x[d] = TB[d][1] + ux + dx;
            y[d] = dc - ua;
            e[d] = EB[d] + ue + de;
        }
        if (kk == 0) {
            // even plane: what goes down leaves now; its own sums and what goes up wait for the partner plane
            stb(A.part[5], c, r[0]);
            stb(A.part[6], c, x[0]);
            stb(A.part[7], c, y[0]);
            stb(A.part[8], c, e[0]);
            sk[0 * T] = r[1], sk[1 * T] = x[1], sk[2 * T] = y[1], sk[3 * T] = e[1];
            sk[4 * T] = r[2], sk[5 * T] = x[2], sk[6 * T] = y[2], sk[7 * T] = e[2];
            c_first = c;
        } else {
            // odd plane: own = its e_z = 0 sums + the even plane's e_z = +1 sums; what it sends up leaves
            stb(A.part[0], c, r[1] + sk[4 * T]);
            stb(A.part[1], c, x[1] + sk[5 * T]);
            stb(A.part[2], c, y[1] + sk[6 * T]);
            stb(A.part[3], c, sk[4 * T]);            // jz: e_z = +1 terms
            stb(A.part[4], c, e[1] + sk[7 * T]);
            stb(A.part[5], c, r[2]);
            stb(A.part[6], c, x[2]);
            stb(A.part[7], c, y[2]);
            stb(A.part[8], c, e[2]);
            // the even plane: its own e_z = 0 sums + this plane's e_z = -1 sums
            stb(A.part[0], c_first, sk[0 * T] + r[0]);
            stb(A.part[1], c_first, sk[1 * T] + x[0]);
            stb(A.part[2], c_first, sk[2 * T] + y[0]);
            stb(A.part[3], c_first, -r[0]);          // jz: e_z = -1 terms
            stb(A.part[4], c_first, sk[3 * T] + e[0]);
        }
    }
    __syncthreads();  // the exchange slots are the next plane's g slots
    }
}

// q-corrections from the 9-word layout of k_collide_tile_pair
__global__ void __launch_bounds__(128, 6) k_qcorr_combine_pair(const double* __restrict__ fin, const double* __restrict__ gin,
                                                               const uint32_t* __restrict__ nbr, const double* __restrict__ part,
                                                               const double* __restrict__ edge, int W, long long esz,
                                                               double* __restrict__ qc, const __grid_constant__ Layout L,
                                                               const __grid_constant__ Phys P, int k0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.nx) return;
    const int j = blockIdx.y, k = blockIdx.z + k0;
    const long long c = L.cell(i, j, k);
    const long long n = L.sq;
    const uint32_t m = nbr[c];
    if (!(m & 1u)) return;
    const bool inner = (L.wrap[0] || (i > 0 && i < L.nx - 1)) && (L.wrap[1] || (j > 0 && j < L.ny - 1)) &&
                       (L.wrap[2] || (k > 0 && k < L.nz - 1)) && k >= 0 && k < L.nz;
    if (!(inner && m == ALL_FLUID)) {
        qcorr_cell<true>(fin, gin, nbr, qc, L, P, i, j, k);
        return;
    }
    const int km = k == 0 ? L.nz - 1 : k - 1, kp = k == L.nz - 1 ? 0 : k + 1;
    // even plane: the plane below (odd) sent up; odd plane: the plane above (even) sent down
    const bool even = (k & 1) == 0;
    const long long co = L.cell(i, j, even ? km : kp);
    const double ro = part[5 * n + co];
    double rho = part[0 * n + c] + ro;
    double jx = part[1 * n + c] + part[6 * n + co];
    double jy = part[2 * n + c] + part[7 * n + co];
    double jz = part[3 * n + c] + (even ? ro : -ro);
    double e2 = part[4 * n + c] + part[8 * n + co];
    {
        const int jr = j / W, w = j - jr * W, nyr = (L.ny + W - 1) / W;
        const int rows = min(W, L.ny - jr * W);
        const long long en = esz * (L.nz + 2 * GZ);
        auto add_edge = [&](int side, int jrs, double sgn) {
            const long long e0 = (long long)(i + OX) + (long long)jrs * L.px;
            const double am = edge[(side * 9 + 6) * en + e0 + (long long)(km + GZ) * esz];
            const double a0 = edge[(side * 9 + 3) * en + e0 + (long long)(k + GZ) * esz];
            const double ap = edge[(side * 9 + 0) * en + e0 + (long long)(kp + GZ) * esz];
            rho += am + a0 + ap;
            jz += am - ap;
            jy += sgn * (am + a0 + ap);
            jx += edge[(side * 9 + 7) * en + e0 + (long long)(km + GZ) * esz] +
                  edge[(side * 9 + 4) * en + e0 + (long long)(k + GZ) * esz] +
                  edge[(side * 9 + 1) * en + e0 + (long long)(kp + GZ) * esz];
            e2 += edge[(side * 9 + 8) * en + e0 + (long long)(km + GZ) * esz] +
                  edge[(side * 9 + 5) * en + e0 + (long long)(k + GZ) * esz] +
                  edge[(side * 9 + 2) * en + e0 + (long long)(kp + GZ) * esz];
        };
        if (w == 0) add_edge(1, jr == 0 ? nyr - 1 : jr - 1, 1.0);
        if (w == rows - 1) add_edge(0, jr == nyr - 1 ? 0 : jr + 1, -1.0);
    }
    const Prim s = primitives(rho, jx, jy, jz, e2, P);
    qc[c] = s.qcx;
    qc[n + c] = s.qcy;
    qc[2 * n + c] = s.qcz;
}

// ---------------------------------------------------------------------------
// Carry step with the z sum completed ON CHIP (variant 9): k_collide_tile marching through the ZM planes of a z-chunk.
// A thread keeps, for its cell column, the sums plane k-1 has received so far (own + from below) and what plane k-1
// sends up, in a private shared-memory column; when plane k has been collided, plane k-1 is complete in z.  For a cell
// whose row is not the first / last of the CTA (complete in y too) and whose 27 sources are all collided cells of the
// box, the thread finishes the job the combine kernel did -- rho, j, 2rhoE -> QCorr -- and stores the THREE words
// into the q-correction array of the next step (a second array: this step's is still being read).  Every other cell
// stores its z-complete sums (5 words: rho, jx, jy, jz, 2rhoE); the first / last plane of a chunk additionally store
// what they send down / up (4 words), exactly the layout of the plane-pair kernel (ZM = 2).  k_qcorr_combine_march
// finishes those cells and skips the finished ones.  Carried words per cell at W = 6, ZM = 8: 8 written + 7.5 read
// instead of 15 + 15.  zpos[k]: 0 interior plane of its chunk, 1 first, 2 last.
// ---------------------------------------------------------------------------
constexpr int MARCH_KEEP_SLOTS = 9;  // P: rho, jx, jy, jz, e2 of plane k-1 so far;  U: rho, jx, jy, e2 it sends up
// PIPE: the populations of plane k+1 travel global -> shared (cp.async: f into 27 slots of its own, g into the second of
// two g buffers) while plane k is collided, and the mask / flag / QCorr neighbours of plane k+1 wait in registers: the
// loads of a CTA overlap its own arithmetic instead of relying on the other resident CTA.  90 slots per thread.
constexpr int march_slots(bool pipe) { return pipe ? 3 * NQ + MARCH_KEEP_SLOTS : NQ + MARCH_KEEP_SLOTS; }
// ABL (MBL_EXPERIMENTS, timing only -- results are wrong): 1 no x shuffles, 2 no y exchange and no barriers, 4 no
// carried stores (the sums become dead code)
// ORD (MBL_EXPERIMENTS, plain path): order in which a plane's loads are issued -- 0: g (cp.async), f, mask / QCorr;
// 1: f, mask / QCorr, g;  2: mask / QCorr, f, g.  Results identical.
template <int W, bool PIPE, int ABL = 0, int ORD = 0>
__global__ void __launch_bounds__(32 * W, PIPE ? (W <= 4 ? 2 : 1) : (W <= 4 ? 3 : W <= 8 ? 2 : 1))
    k_collide_tile_march(const __grid_constant__ CarryPtrs A, const __grid_constant__ MarchOut Q,
                         const uint32_t* __restrict__ nbr, const uint8_t* __restrict__ flag,
                         const __grid_constant__ Layout L, const __grid_constant__ Phys P,
                         const __grid_constant__ CarryPlan C, int ka, int kb, int zm)
{
    constexpr int T = 32 * W;
    extern __shared__ double smem[];
    double* const sg0 = smem + threadIdx.x;  // sg0[slot * T]; PIPE: g buffers at slots 0 and NQ, f at 2 NQ
    const unsigned sg0_addr = (unsigned)__cvta_generic_to_shared(sg0);
    double* const sk = smem + (PIPE ? 3 * NQ : NQ) * T + threadIdx.x;  // sk[slot * T]: carried sums of plane k-1
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int xc = blockIdx.x;
    const int y0 = blockIdx.y * W;
    const int rows = min(W, L.ny - y0);
    const int i = xc * C.own - C.halo + lane;
    const bool own = lane >= C.halo && lane < C.halo + C.own && i < L.nx && w < rows;
    int is = i, j = min(y0 + w, L.ny - 1);
    if (is < 0) is = L.wrap[0] ? is + L.nx : 0;
    if (is >= L.nx) is = (L.wrap[0] && is - L.nx < L.nx) ? is - L.nx : L.nx - 1;
    const unsigned FULL = 0xffffffffu;
    const unsigned px8 = (unsigned)L.px * 8u, sz8 = (unsigned)L.sz * 8u;
    unsigned xo[3], yo[3], zo[3];
    xo[1] = yo[1] = zo[1] = 0u;
    xo[2] = (L.wrap[0] && is == 0) ? (unsigned)(L.nx - 1) * 8u : 0u - 8u;
    xo[0] = (L.wrap[0] && is == L.nx - 1) ? 0u - (unsigned)(L.nx - 1) * 8u : 8u;
    yo[2] = (L.wrap[1] && j == 0) ? (unsigned)(L.ny - 1) * px8 : 0u - px8;
    yo[0] = (L.wrap[1] && j == L.ny - 1) ? 0u - (unsigned)(L.ny - 1) * px8 : px8;
    const bool first = w == 0, last = w == rows - 1;
    // cells this thread may finish itself: complete in y inside the CTA, not on a non-wrapped x / y face of the box
    const bool xy_inner = own && !first && !last && (L.wrap[0] || (i > 0 && i < L.nx - 1)) &&
                          (L.wrap[1] || (j > 0 && j < L.ny - 1));
    const int k0c = ka + (int)blockIdx.z * zm;
    const int nk = min(zm, kb - k0c);
    unsigned c_prev = 0u;
    uint32_t m_prev = 0u;
    auto ldb = [](const double* base, unsigned off) { return *(const double*)((const char*)base + off); };
    auto stb = [](double* base, unsigned off, double v) { *(double*)((char*)base + off) = v; };
    auto stc = [&](double* base, unsigned off, double v) {  // carried words
        if constexpr (!(ABL & 4)) *(double*)((char*)base + off) = v;
    };
    const unsigned cxy = (unsigned)(is + OX) * 8u + (unsigned)(j + GY) * px8;
    // populations of plane k: g -> g buffer `buf`, f -> registers (plain) or the f slots (PIPE)
    auto pull = [&](int k, int buf, double* f, int which = 3) {  // which: 1 = g, 2 = f
        zo[2] = (L.wrap[2] && k == 0) ? (unsigned)(L.nz - 1) * sz8 : 0u - sz8;
        zo[0] = (L.wrap[2] && k == L.nz - 1) ? 0u - (unsigned)(L.nz - 1) * sz8 : sz8;
        const unsigned c = cxy + (unsigned)(k + GZ) * sz8;
        unsigned cyz[3][3];
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int d = 0; d < 3; ++d) cyz[b][d] = c + yo[b] + zo[d];
        if (which & 1)
            static_for<0, NQ>([&](auto qc_) {
                constexpr int Qd = decltype(qc_)::value;
                cp_async8(sg0_addr + (buf * NQ + Qd) * T * 8, (const char*)A.gin[Qd] + (cyz[ey(Qd) + 1][ez(Qd) + 1] + xo[ex(Qd) + 1]));
            });
        if (which & 2)
            static_for<0, NQ>([&](auto qc_) {
                constexpr int Qd = decltype(qc_)::value;
                if constexpr (PIPE)
                    cp_async8(sg0_addr + (2 * NQ + Qd) * T * 8, (const char*)A.fin[Qd] + (cyz[ey(Qd) + 1][ez(Qd) + 1] + xo[ex(Qd) + 1]));
                else
                    f[Qd] = ldb(A.fin[Qd], cyz[ey(Qd) + 1][ez(Qd) + 1] + xo[ex(Qd) + 1]);
            });
    };
    // mask, gradient flags and the six QCorr neighbours of plane k
    struct Small {
        uint32_t m;
        unsigned fb;
        double qxp, qxm, qyp, qym, qzp, qzm;
    };
    auto pull_small = [&](int k) {
        const unsigned c = cxy + (unsigned)(k + GZ) * sz8;
        Small S;
        S.m = *(const uint32_t*)((const char*)nbr + (c >> 1));
        S.fb = flag[c >> 3];
        S.qxp = ldb(A.qc[0], c + 8u), S.qxm = ldb(A.qc[0], c - 8u);
        S.qyp = ldb(A.qc[1], c + px8), S.qym = ldb(A.qc[1], c - px8);
        S.qzp = ldb(A.qc[2], c + sz8), S.qzm = ldb(A.qc[2], c - sz8);
        return S;
    };
    Small Sn;
    if constexpr (PIPE) {
        pull(k0c, 0, nullptr);
        Sn = pull_small(k0c);
    }
#pragma unroll 1
    for (int kk = 0; kk < nk; ++kk) {
        const int k = k0c + kk;
        const unsigned c = cxy + (unsigned)(k + GZ) * sz8;
        double* const sg = PIPE ? sg0 + (kk & 1) * NQ * T : sg0;  // this plane's g slots
        double f[NQ];
        Small S;
        if constexpr (PIPE) {
            S = Sn;
            cp_async_wait_all();
#pragma unroll
            for (int q = 0; q < NQ; ++q) f[q] = sg0[(2 * NQ + q) * T];
        } else if constexpr (ORD == 1) {
            pull(k, 0, f, 2);
            S = pull_small(k);
            pull(k, 0, f, 1);
        } else if constexpr (ORD == 2) {
            S = pull_small(k);
            pull(k, 0, f, 2);
            pull(k, 0, f, 1);
        } else {
            pull(k, 0, f);
            S = pull_small(k);
        }
        const uint32_t m = S.m;
        const unsigned fb = S.fb;
        const double qxp = S.qxp, qxm = S.qxm, qyp = S.qyp, qym = S.qym, qzp = S.qzp, qzm = S.qzm;
        const bool fluid = m & 1u;
        if (m != ALL_FLUID) {
            if (fluid) {
                static_for<1, NQ>([&](auto qc_) {
                    constexpr int Qd = decltype(qc_)::value;
                    if (!((m >> Qd) & 1u)) f[Qd] = ldb(A.fin[opp(Qd)], c);
                });
            } else {
#pragma unroll
                for (int q = 0; q < NQ; ++q) f[q] = -1.0;
            }
        }
        MomF mf = moments_f([&](int q) { return f[q]; });
        if constexpr (PIPE) {
            // the f slots have been read (mf depends on all of them): plane k+1 may land there and in the other g buffer
            if (kk + 1 < nk) {
                asm volatile("" : "+d"(mf.rho)::"memory");
                pull(k + 1, (kk + 1) & 1, nullptr);
                Sn = pull_small(k + 1);
            }
        } else {
            cp_async_wait_all_after(mf.rho);
        }
        if (m != ALL_FLUID) {
            if (fluid) {
                static_for<1, NQ>([&](auto qc_) {
                    constexpr int Qd = decltype(qc_)::value;
                    if (!((m >> Qd) & 1u)) sg[Qd * T] = ldb(A.gin[opp(Qd)], c);
                });
            } else {
#pragma unroll
                for (int q = 0; q < NQ; ++q) sg[q * T] = -1.0;
            }
        }
        const MomG mg = moments_g([&](int q) { return sg[q * T]; });
        const Prim s = primitives(mf.rho, mf.jx, mf.jy, mf.jz, mg.e2, P);
        const double dqx = one_sided_gradient(fb & GRAD_PX, fb & GRAD_MX, (fb & GRAD_PX) ? qxp : 0.0, s.qcx,
                                              (fb & GRAD_MX) ? qxm : 0.0, P.idx[0]);
        const double dqy = one_sided_gradient(fb & GRAD_PY, fb & GRAD_MY, (fb & GRAD_PY) ? qyp : 0.0, s.qcy,
                                              (fb & GRAD_MY) ? qym : 0.0, P.idx[1]);
        const double dqz = one_sided_gradient(fb & GRAD_PZ, fb & GRAD_MZ, (fb & GRAD_PZ) ? qzp : 0.0, s.qcz,
                                              (fb & GRAD_MZ) ? qzm : 0.0, P.idx[2]);
        const Coll cc = collision_coefficients(s, mf, mg, dqx, dqy, dqz, P);
        const double omega = fluid ? cc.omega : 0.0;
        double TA[3][2] = {}, TB[3][2] = {}, TC[3][2] = {};
        double EA[3] = {}, EB[3] = {}, EC[3] = {};
        static_for<0, NQ>([&](auto qc_) {
            constexpr int Qd = decltype(qc_)::value;
            constexpr int d = ez(Qd) + 1;
            const double v = f[Qd] + omega * (feq_q<Qd>(cc) - f[Qd]);
            if (own) stb(A.fout[Qd], c, v);
            double t = v;
            if constexpr (!(ABL & 1)) {
                if constexpr (ex(Qd) == 1) t = __shfl_up_sync(FULL, v, 1);
                if constexpr (ex(Qd) == -1) t = __shfl_down_sync(FULL, v, 1);
            }
            double(&Tt)[3][2] = ey(Qd) == -1 ? TA : ey(Qd) == 0 ? TB : TC;
            Tt[d][0] += t;
            if constexpr (ex(Qd) == 1) Tt[d][1] += t;
            if constexpr (ex(Qd) == -1) Tt[d][1] -= t;
        });
        static_for<0, NQ>([&](auto qc_) {
            constexpr int Qd = decltype(qc_)::value;
            constexpr int d = ez(Qd) + 1;
            const double gq = sg[Qd * T];
            const double v = gq + omega * (geq_q<Qd>(cc) - gq);
            if (own) stb(A.gout[Qd], c, v);
            double t = v;
            if constexpr (!(ABL & 1)) {
                if constexpr (ex(Qd) == 1) t = __shfl_up_sync(FULL, v, 1);
                if constexpr (ex(Qd) == -1) t = __shfl_down_sync(FULL, v, 1);
            }
            double(&E)[3] = ey(Qd) == -1 ? EA : ey(Qd) == 0 ? EB : EC;
            E[d] += t;
        });
        // y exchange through the g slots; the CTA's first / last row send through the edge arrays
        const unsigned ce = (unsigned)(i + OX) * 8u + (unsigned)blockIdx.y * px8 + (unsigned)(k + GZ) * (unsigned)C.esz8;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (first) {
                if (own && !(ABL & 4)) {
                    stb(A.edge[3 * d + 0], ce, TA[d][0]);
                    stb(A.edge[3 * d + 1], ce, TA[d][1]);
                    stb(A.edge[3 * d + 2], ce, EA[d]);
                }
            } else if (!(ABL & 2)) {
                sg[(3 * d + 0) * T] = TA[d][0];
                sg[(3 * d + 1) * T] = TA[d][1];
                sg[(3 * d + 2) * T] = EA[d];
            }
            if (last) {
                if (own && !(ABL & 4)) {
                    stb(A.edge[9 + 3 * d + 0], ce, TC[d][0]);
                    stb(A.edge[9 + 3 * d + 1], ce, TC[d][1]);
                    stb(A.edge[9 + 3 * d + 2], ce, EC[d]);
                }
            } else if (!(ABL & 2)) {
                sg[(9 + 3 * d + 0) * T] = TC[d][0];
                sg[(9 + 3 * d + 1) * T] = TC[d][1];
                sg[(9 + 3 * d + 2) * T] = EC[d];
            }
        }
        if constexpr (!(ABL & 2)) __syncthreads();
        if (own) {
            const double* up = sg + 32;   // row j+1: its e_y = -1 terms arrive here
            const double* dn = sg - 32;   // row j-1: its e_y = +1 terms
            double r[3], x[3], y[3], e[3];  // per destination plane k-1, k, k+1: rho, jx, jy, e2 (y-complete up to the CTA edges)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const bool nu = last || (ABL & 2), nd = first || (ABL & 2);
                const double ua = nu ? 0.0 : up[(3 * d + 0) * T], dc = nd ? 0.0 : dn[(9 + 3 * d + 0) * T];
                const double ux = nu ? 0.0 : up[(3 * d + 1) * T], dx = nd ? 0.0 : dn[(9 + 3 * d + 1) * T];
                const double ue = nu ? 0.0 : up[(3 * d + 2) * T], de = nd ? 0.0 : dn[(9 + 3 * d + 2) * T];
                r[d] = TB[d][0] + ua + dc;
                x[d] = TB[d][1] + ux + dx;
                y[d] = dc - ua;
                e[d] = EB[d] + ue + de;
            }
            if (kk == 0) {
                // first plane of the chunk: what it sends down leaves now (another chunk's last plane waits for it)
                stc(A.part[5], c, r[0]);
                stc(A.part[6], c, x[0]);
                stc(A.part[7], c, y[0]);
                stc(A.part[8], c, e[0]);
                sk[0 * T] = r[1], sk[1 * T] = x[1], sk[2 * T] = y[1], sk[3 * T] = 0.0, sk[4 * T] = e[1];
            } else {
                // plane k-1 is complete in z: its sums so far + this plane's e_z = -1 terms
                const double Sr = sk[0 * T] + r[0], Sx = sk[1 * T] + x[0], Sy = sk[2 * T] + y[0];
                const double Sz = sk[3 * T] - r[0], Se = sk[4 * T] + e[0];
                const int kp = k - 1;
                const bool z_inner = kk >= 2 && (L.wrap[2] || (kp > 0 && kp < L.nz - 1));
                if (xy_inner && z_inner && m_prev == ALL_FLUID) {
                    const Prim sp = primitives(Sr, Sx, Sy, Sz, Se, P);
                    stc(Q.qcn[0], c_prev, sp.qcx);
                    stc(Q.qcn[1], c_prev, sp.qcy);
                    stc(Q.qcn[2], c_prev, sp.qcz);
                } else {
                    stc(A.part[0], c_prev, Sr);
                    stc(A.part[1], c_prev, Sx);
                    stc(A.part[2], c_prev, Sy);
                    stc(A.part[3], c_prev, Sz);
                    stc(A.part[4], c_prev, Se);
                }
                // plane k so far: what plane k-1 sent up + its own e_z = 0 terms
                const double ur = sk[5 * T];
                sk[0 * T] = ur + r[1], sk[1 * T] = sk[6 * T] + x[1], sk[2 * T] = sk[7 * T] + y[1], sk[3 * T] = ur,
                       sk[4 * T] = sk[8 * T] + e[1];
            }
            sk[5 * T] = r[2], sk[6 * T] = x[2], sk[7 * T] = y[2], sk[8 * T] = e[2];
            if (kk == nk - 1) {
                // last plane of the chunk: incomplete (the next chunk's first plane sends down); its sums so far and
                // what it sends up leave
                stc(A.part[0], c, sk[0 * T]);
                stc(A.part[1], c, sk[1 * T]);
                stc(A.part[2], c, sk[2 * T]);
                stc(A.part[3], c, sk[3 * T]);
                stc(A.part[4], c, sk[4 * T]);
                if (kk > 0) {  // (a one-plane chunk keeps its send-down words; the launcher never makes one)
                    stc(A.part[5], c, r[2]);
                    stc(A.part[6], c, x[2]);
                    stc(A.part[7], c, y[2]);
                    stc(A.part[8], c, e[2]);
                }
            }
        }
        c_prev = c;
        m_prev = m;
        if constexpr (!(ABL & 2)) __syncthreads();  // the exchange slots are the next plane's g slots
    }
}

// finishes what k_collide_tile_march left: cells on the first / last row of a CTA, on the first / last plane of a
// chunk, next to a non-wrapped face, or with a solid source (those are pulled, as in k_qcorr)
__global__ void __launch_bounds__(128, 6) k_qcorr_combine_march(const double* __restrict__ fin, const double* __restrict__ gin,
                                                                const uint32_t* __restrict__ nbr, const double* __restrict__ part,
                                                                const double* __restrict__ edge, const signed char* __restrict__ zpos,
                                                                int W, long long esz, double* __restrict__ qc,
                                                                const __grid_constant__ Layout L, const __grid_constant__ Phys P, int k0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.nx) return;
    const int j = blockIdx.y, k = blockIdx.z + k0;
    const long long c = L.cell(i, j, k);
    const long long n = L.sq;
    const uint32_t m = nbr[c];
    if (!(m & 1u)) return;
    const bool inner = (L.wrap[0] || (i > 0 && i < L.nx - 1)) && (L.wrap[1] || (j > 0 && j < L.ny - 1)) &&
                       (L.wrap[2] || (k > 0 && k < L.nz - 1)) && k >= 0 && k < L.nz;
    if (!(inner && m == ALL_FLUID)) {
        qcorr_cell<true>(fin, gin, nbr, qc, L, P, i, j, k);
        return;
    }
    const int zp = zpos[k + GZ];
    const int jr = j / W, w = j - jr * W, nyr = (L.ny + W - 1) / W;
    const int rows = min(W, L.ny - jr * W);
    const bool edge_row = w == 0 || w == rows - 1;
    if (zp == 0 && !edge_row) return;  // finished by the collide kernel
    const int km = k == 0 ? L.nz - 1 : k - 1, kp = k == L.nz - 1 ? 0 : k + 1;
    double rho = part[0 * n + c], jx = part[1 * n + c], jy = part[2 * n + c], jz = part[3 * n + c], e2 = part[4 * n + c];
    if (zp != 0) {
        // first plane: the plane below (another chunk's last) sent up; last plane: the plane above sent down
        const long long co = L.cell(i, j, zp == 1 ? km : kp);
        const double ro = part[5 * n + co];
        rho += ro;
        jx += part[6 * n + co];
        jy += part[7 * n + co];
        jz += zp == 1 ? ro : -ro;
        e2 += part[8 * n + co];
    }
    if (edge_row) {
        const long long en = esz * (L.nz + 2 * GZ);
        auto add_edge = [&](int side, int jrs, double sgn) {
            const long long e0 = (long long)(i + OX) + (long long)jrs * L.px;
            const double am = edge[(side * 9 + 6) * en + e0 + (long long)(km + GZ) * esz];
            const double a0 = edge[(side * 9 + 3) * en + e0 + (long long)(k + GZ) * esz];
            const double ap = edge[(side * 9 + 0) * en + e0 + (long long)(kp + GZ) * esz];
            rho += am + a0 + ap;
            jz += am - ap;
            jy += sgn * (am + a0 + ap);
            jx += edge[(side * 9 + 7) * en + e0 + (long long)(km + GZ) * esz] +
                  edge[(side * 9 + 4) * en + e0 + (long long)(k + GZ) * esz] +
                  edge[(side * 9 + 1) * en + e0 + (long long)(kp + GZ) * esz];
            e2 += edge[(side * 9 + 8) * en + e0 + (long long)(km + GZ) * esz] +
                  edge[(side * 9 + 5) * en + e0 + (long long)(k + GZ) * esz] +
                  edge[(side * 9 + 2) * en + e0 + (long long)(kp + GZ) * esz];
        };
        if (w == 0) add_edge(1, jr == 0 ? nyr - 1 : jr - 1, 1.0);
        if (w == rows - 1) add_edge(0, jr == nyr - 1 ? 0 : jr + 1, -1.0);
    }
    const Prim s = primitives(rho, jx, jy, jz, e2, P);
    qc[c] = s.qcx;
    qc[n + c] = s.qcy;
    qc[2 * n + c] = s.qcz;
}

// q-corrections of the post-stream state from the carried plane sums (interior cells) or by pulling the
// populations (everything k_collide_carry could not serve)
__global__ void __launch_bounds__(128, 6) k_qcorr_combine(const double* __restrict__ fin, const double* __restrict__ gin,
                                                          const uint32_t* __restrict__ nbr, const double* __restrict__ part,
                                                          const double* __restrict__ edge, int W, long long esz,
                                                          double* __restrict__ qc, const __grid_constant__ Layout L,
                                                          const __grid_constant__ Phys P, int k0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.nx) return;
    const int j = blockIdx.y, k = blockIdx.z + k0;
    const long long c = L.cell(i, j, k);
    const long long n = L.sq;
    const uint32_t m = nbr[c];
    const bool inner = (L.wrap[0] || (i > 0 && i < L.nx - 1)) && (L.wrap[1] || (j > 0 && j < L.ny - 1)) &&
                       (L.wrap[2] || (k > 0 && k < L.nz - 1)) && k >= 0 && k < L.nz;
    // the twelve part words are loaded before the mask is looked at (they are always inside the padded box):
    // one memory latency per cell instead of two
    const int kc = min(max(k, 0), L.nz - 1);
    const int km = kc == 0 ? L.nz - 1 : kc - 1, kp = kc == L.nz - 1 ? 0 : kc + 1;
    const long long c0 = L.cell(i, j, kc);
    const long long cm = L.cell(i, j, km);      // source plane k-1 sends with c = +1
    const long long cp = L.cell(i, j, kp);      // source plane k+1 sends with c = -1
    const double rm = part[8 * n + cm], r0 = part[4 * n + c0], rp = part[0 * n + cp];
    const double xm = part[9 * n + cm], x0 = part[5 * n + c0], xp = part[1 * n + cp];
    const double ym = part[10 * n + cm], y0 = part[6 * n + c0], yp = part[2 * n + cp];
    const double em = part[11 * n + cm], e0 = part[7 * n + c0], ep = part[3 * n + cp];
    if (!(m & 1u)) return;
    if (!(inner && m == ALL_FLUID)) {
        qcorr_cell<true>(fin, gin, nbr, qc, L, P, i, j, k);
        return;
    }
    double rho = rm + r0 + rp;
    double jz = rm - rp;
    double jx = xm + x0 + xp;
    double jy = ym + y0 + yp;
    double e2 = em + e0 + ep;
    if (W > 0) {
        // k_collide_tile: the rows at the edges of a CTA lack what the neighbouring CTA's adjacent row sent
        // (edge arrays: [side 0 = first row's e_y = -1 sums | side 1 = last row's e_y = +1 sums][c][rho, jx, e2])
        const int jr = j / W, w = j - jr * W, nyr = (L.ny + W - 1) / W;
        const int rows = min(W, L.ny - jr * W);
        const long long en = esz * (L.nz + 2 * GZ);
        auto add_edge = [&](int side, int jrs, double sgn) {
            const long long e0 = (long long)(i + OX) + (long long)jrs * L.px;
            const double am = edge[(side * 9 + 6) * en + e0 + (long long)(km + GZ) * esz];
            const double a0 = edge[(side * 9 + 3) * en + e0 + (long long)(k + GZ) * esz];
            const double ap = edge[(side * 9 + 0) * en + e0 + (long long)(kp + GZ) * esz];
            rho += am + a0 + ap;
            jz += am - ap;
            jy += sgn * (am + a0 + ap);
            jx += edge[(side * 9 + 7) * en + e0 + (long long)(km + GZ) * esz] +
                  edge[(side * 9 + 4) * en + e0 + (long long)(k + GZ) * esz] +
                  edge[(side * 9 + 1) * en + e0 + (long long)(kp + GZ) * esz];
            e2 += edge[(side * 9 + 8) * en + e0 + (long long)(km + GZ) * esz] +
                  edge[(side * 9 + 5) * en + e0 + (long long)(k + GZ) * esz] +
                  edge[(side * 9 + 2) * en + e0 + (long long)(kp + GZ) * esz];
        };
        if (w == 0) add_edge(1, jr == 0 ? nyr - 1 : jr - 1, 1.0);          // row j-1 is another CTA's last row
        if (w == rows - 1) add_edge(0, jr == nyr - 1 ? 0 : jr + 1, -1.0);  // row j+1 is another CTA's first row
    }
    const Prim s = primitives(rho, jx, jy, jz, e2, P);
    qc[c] = s.qcx;
    qc[n + c] = s.qcy;
    qc[2 * n + c] = s.qcz;
}

// ---------------------------------------------------------------------------
// stream only (pull form of LBM.cpp:558-604): lat = blockIdx.y selects f / g via pointer arrays
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_stream(const double* __restrict__ fin, const double* __restrict__ gin,
                                                double* __restrict__ fout, double* __restrict__ gout,
                                                const uint32_t* __restrict__ nbr, Layout L)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int k = blockIdx.z;
    if (i >= L.nx) return;
    const long long c = L.cell(i, j, k);
    const long long n = L.sq;
    const uint32_t m = nbr[c];
    if (!(m & 1u)) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            fout[q * n + c] = -1.0;
            gout[q * n + c] = -1.0;
        }
        return;
    }
    const PullOffsets o = pull_offsets(L, i, j, k);
    gather27<true, false>(fin, c, m, L, o, [&](auto qc_, double v) { fout[(long long)decltype(qc_)::value * n + c] = v; });
    gather27<true, false>(gin, c, m, L, o, [&](auto qc_, double v) { gout[(long long)decltype(qc_)::value * n + c] = v; });
}

// f_to_macrodata on the current (already streamed) state, valid cells (LBM.cpp:810-906)
__global__ void __launch_bounds__(128) k_macrodata(const double* __restrict__ f, const double* __restrict__ g,
                                                   const uint8_t* __restrict__ flag, double* __restrict__ macro,
                                                   Layout L, Phys P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int k = blockIdx.z;
    if (i >= L.nx) return;
    const long long c = L.cell(i, j, k);
    const long long n = L.sq;
    if (!(flag[c] & FLAG_FLUID)) return;
    MomF mf = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    MomG mg = {0, 0, 0, 0};
    const PullOffsets o = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    gather27<false, true>(f, c, 0u, L, o, [&](auto qc_, double v) { acc_f<decltype(qc_)::value>(mf, v); });
    gather27<false, true>(g, c, 0u, L, o, [&](auto qc_, double v) { acc_g<decltype(qc_)::value>(mg, v); });
    const Prim s = primitives(mf.rho, mf.jx, mf.jy, mf.jz, mg.e2, P);
    macro[0 * n + c] = s.rho;
    macro[1 * n + c] = s.u;
    macro[2 * n + c] = s.v;
    macro[3 * n + c] = s.w;
    macro[4 * n + c] = sqrt(s.u * s.u + s.v * s.v + s.w * s.w);
    macro[5 * n + c] = mg.e2;
    macro[6 * n + c] = s.qcx;
    macro[7 * n + c] = s.qcy;
    macro[8 * n + c] = s.qcz;
    macro[9 * n + c] = mf.pxx;
    macro[10 * n + c] = mf.pyy;
    macro[11 * n + c] = mf.pzz;
    macro[12 * n + c] = mf.pxy;
    macro[13 * n + c] = mf.pxz;
    macro[14 * n + c] = mf.pyz;
    macro[15 * n + c] = mg.qx;
    macro[16 * n + c] = mg.qy;
    macro[17 * n + c] = mg.qz;
    macro[18 * n + c] = s.T;
}

// compute_derived (LBM.cpp:909-955): vorticity from the velocity macrodata of valid cells.
// Needs macrodata of the face neighbours inside the domain (valid cells of this box; the
// z-neighbours across a rank boundary are treated as unusable -- plot-only quantity).
__global__ void __launch_bounds__(128) k_derived(const uint8_t* __restrict__ flag, const double* __restrict__ macro,
                                                 double* __restrict__ derived, Layout L, Phys P, int with_dq, int zghost)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int k = blockIdx.z;
    if (i >= L.nx) return;
    const long long c = L.cell(i, j, k);
    const long long n = L.sq;
    const unsigned fb = flag[c];
    if (!(fb & FLAG_FLUID)) return;
    const long long step[3] = {1, L.px, L.sz};
    const unsigned bp[3] = {GRAD_PX, GRAD_PY, GRAD_PZ}, bm[3] = {GRAD_MX, GRAD_MY, GRAD_MZ};
    // zghost bit 0 / 1: the macrodata ghost plane below / above holds the neighbouring rank's plane
    const bool zin_p = (k + 1 < L.nz) || (zghost & 2), zin_m = (k - 1 >= 0) || (zghost & 1);
    auto grad = [&](int dir, int comp) {
        bool okp = fb & bp[dir], okm = fb & bm[dir];
        if (dir == 2) {
            okp = okp && zin_p;
            okm = okm && zin_m;
        }
        const double* a = macro + (long long)comp * n;
        const double dp = okp ? a[c + step[dir]] : 0.0, dm = okm ? a[c - step[dir]] : 0.0;
        return one_sided_gradient(okp, okm, dp, a[c], dm, P.idx[dir]);
    };
    const double vx = grad(0, 2), wx = grad(0, 3), uy = grad(1, 1), wy = grad(1, 3), uz = grad(2, 1), vz = grad(2, 2);
    derived[0 * n + c] = wy - vz;
    derived[1 * n + c] = uz - wx;
    derived[2 * n + c] = vx - uy;
    derived[3 * n + c] = sqrt((wy - vz) * (wy - vz) + (uz - wx) * (uz - wx) + (vx - uy) * (vx - uy));
    if (with_dq) {
        // compute_q_corrections (LBM.cpp:959-991) from the stored QCorr: only when no collide has produced them
        // (macrodata of an initial or restarted state, LBM.cpp:1196-1198, 1903-1914)
        derived[4 * n + c] = grad(0, 6);
        derived[5 * n + c] = grad(1, 7);
        derived[6 * n + c] = grad(2, 8);
    }
}

// compute_eb_forces (LBM.cpp:994-1044), single level: momentum exchange over solid cells that
// touch fluid (flag bit1); one double atomicAdd per block and direction
__global__ void __launch_bounds__(128) k_eb_forces(const double* __restrict__ f, const uint8_t* __restrict__ flag,
                                                   double* __restrict__ out3, Layout L)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int k = blockIdx.z;
    double fs[3] = {0.0, 0.0, 0.0};
    if (i < L.nx) {
        const long long c = L.cell(i, j, k);
        if (flag[c] & FLAG_EBB) {
            for (int q = 0; q < NQ; ++q) {
                const int o = c_dir.opp[q];
                const long long r = c + c_dir.ex[o] + c_dir.ey[o] * L.px + c_dir.ez[o] * L.sz;
                if (flag[r] & FLAG_FLUID) {
                    const double v = 2.0 * f[(long long)q * L.sq + r];
                    fs[0] += c_dir.ex[q] * v;
                    fs[1] += c_dir.ey[q] * v;
                    fs[2] += c_dir.ez[q] * v;
                }
            }
        }
    }
    __shared__ double sh[3][4];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double v = fs[d];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sh[d][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) v += sh[threadIdx.x][w];
        if (v != 0.0) atomicAdd(out3 + threadIdx.x, v);
    }
}

// ---------------------------------------------------------------------------
// ghost shell enumeration: the padded box minus the valid box as six slabs
// ---------------------------------------------------------------------------
struct Shell {
    long long off[7];
    int nxp, nyp;  // padded extents in x (ghost only, no alignment pad) and y
};
static Shell make_shell(const Layout& L)
{
    Shell s;
    s.nxp = L.nx + 2 * GX;
    s.nyp = L.ny + 2 * GY;
    const long long zslab = (long long)GZ * s.nxp * s.nyp;
    const long long yslab = (long long)GY * s.nxp * L.nz;
    const long long xslab = (long long)GX * L.ny * L.nz;
    s.off[0] = 0;
    s.off[1] = zslab;
    s.off[2] = 2 * zslab;
    s.off[3] = s.off[2] + yslab;
    s.off[4] = s.off[3] + yslab;
    s.off[5] = s.off[4] + xslab;
    s.off[6] = s.off[5] + xslab;
    return s;
}
__device__ __forceinline__ void shell_decode(const Shell& s, const Layout& L, long long t, int& i, int& j, int& k)
{
    if (t < s.off[2]) {  // z slabs: all i, j
        const bool hi = t >= s.off[1];
        t -= hi ? s.off[1] : 0;
        i = (int)(t % s.nxp) - GX;
        t /= s.nxp;
        j = (int)(t % s.nyp) - GY;
        const int kk = (int)(t / s.nyp);
        k = hi ? L.nz + kk : kk - GZ;
    } else if (t < s.off[4]) {  // y slabs: valid k, all i
        const bool hi = t >= s.off[3];
        t -= hi ? s.off[3] : s.off[2];
        i = (int)(t % s.nxp) - GX;
        t /= s.nxp;
        k = (int)(t % L.nz);
        const int jj = (int)(t / L.nz);
        j = hi ? L.ny + jj : jj - GY;
    } else {  // x slabs: valid j, k
        const bool hi = t >= s.off[5];
        t -= hi ? s.off[5] : s.off[4];
        j = (int)(t % L.ny);
        t /= L.ny;
        k = (int)(t % L.nz);
        const int ii = (int)(t / L.nz);
        i = hi ? L.nx + ii : ii - GX;
    }
}

__device__ __forceinline__ bool in_dom(const Layout& L, int d, int local) { return local + L.lo[d] >= L.dlo[d] && local + L.lo[d] <= L.dhi[d]; }
__device__ __forceinline__ bool in_padded(const Layout& L, int i, int j, int k)
{
    return i >= -GX && i <= L.nx - 1 + GX && j >= -GY && j <= L.ny - 1 + GY && k >= -GZ && k <= L.nz - 1 + GZ;
}

// periodic wrap in z on a box that spans the domain in z (FillBoundary, local copies)
__global__ void __launch_bounds__(128) k_zwrap(double* __restrict__ f, double* __restrict__ g, Layout L)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int kk = blockIdx.z;  // 0 .. 2*GZ-1
    if (i >= L.nx) return;
    const int k = kk < GZ ? kk - GZ : L.nz + (kk - GZ);
    int ks = k % L.nz;
    if (ks < 0) ks += L.nz;
    const long long dst = L.cell(i, j, k), src = L.cell(i, j, ks);
    for (int q = 0; q < NQ; ++q) {
        f[q * L.sq + dst] = f[q * L.sq + src];
        g[q * L.sq + dst] = g[q * L.sq + src];
    }
}

// K6 pre-pass (FillPatchOps.H:92-108) restricted to the ghosts the periodic fill will not overwrite
__global__ void __launch_bounds__(128) k_prepass(double* __restrict__ f, double* __restrict__ g, Layout L, Shell S,
                                                 BcInfo B)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S.off[6]) return;
    int i, j, k;
    shell_decode(S, L, t, i, j, k);
    const int iv[3] = {i, j, k};
    bool outside_np = false;
    for (int d = 0; d < 3; ++d) outside_np |= (!B.periodic[d] && !in_dom(L, d, iv[d]));
    if (!outside_np) return;
    const long long c = L.cell(i, j, k);
    for (int q = 0; q < NQ; ++q) {
        const int in = i + c_dir.ex[q], jn = j + c_dir.ey[q], kn = k + c_dir.ez[q];
        if (in_dom(L, 0, in) && in_dom(L, 1, jn) && in_dom(L, 2, kn) && in_padded(L, in, jn, kn)) {
            const long long s = L.cell(in, jn, kn) + (long long)c_dir.opp[q] * L.sq;
            f[q * L.sq + c] = f[s];
            g[q * L.sq + c] = g[s];
        }
    }
}

// periodic wrap in x and y for every ghost cell that FillBoundary would fill
__global__ void __launch_bounds__(128) k_xywrap(double* __restrict__ f, double* __restrict__ g, Layout L, Shell S,
                                                BcInfo B)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S.off[6]) return;
    int i, j, k;
    shell_decode(S, L, t, i, j, k);
    const bool xin = i >= 0 && i < L.nx, yin = j >= 0 && j < L.ny;
    if (xin && yin) return;  // z ghost of valid (i,j): filled by k_zwrap / halo exchange
    if (!(xin || B.periodic[0]) || !(yin || B.periodic[1]) || !(in_dom(L, 2, k) || B.periodic[2])) return;
    const int is = xin ? i : (i < 0 ? i + L.nx : i - L.nx);
    const int js = yin ? j : (j < 0 ? j + L.ny : j - L.ny);
    const long long dst = L.cell(i, j, k), src = L.cell(is, js, k);
    for (int q = 0; q < NQ; ++q) {
        f[q * L.sq + dst] = f[q * L.sq + src];
        g[q * L.sq + dst] = g[q * L.sq + src];
    }
}

struct Region {
    int lo[3], hi[3];     // cells of the region (local indices, inclusive)
    int in_lo[3], in_hi[3];  // the `inside` box of BC.H:374-380
};

// BCFill::operator() (BC.H:345-471) for the cells of one face / edge / corner region;
// blockIdx.y = lattice (0: f, 1: g = energy lattice)
__global__ void __launch_bounds__(64) k_bc_region(double* __restrict__ f, double* __restrict__ g, Layout L, BcInfo B,
                                                  Region Rg)
{
    const int n0 = Rg.hi[0] - Rg.lo[0] + 1, n1 = Rg.hi[1] - Rg.lo[1] + 1, n2 = Rg.hi[2] - Rg.lo[2] + 1;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n0 * n1 * n2) return;
    const int i = Rg.lo[0] + (int)(t % n0);
    t /= n0;
    const int j = Rg.lo[1] + (int)(t % n1);
    const int k = Rg.lo[2] + (int)(t / n1);
    const bool energy = blockIdx.y == 1;
    double* __restrict__ data = energy ? g : f;
    const long long c = L.cell(i, j, k);
    const int iv[3] = {i, j, k};
    for (int idir = 0; idir < 3; ++idir) {
        for (int lohi = 0; lohi < 2; ++lohi) {
            const int g_iv = iv[idir] + L.lo[idir];
            if (!((lohi == 0 && g_iv < L.dlo[idir]) || (lohi == 1 && g_iv > L.dhi[idir]))) continue;
            const int ndir = lohi == 0 ? 1 : -1;
            const int bc = B.bc[idir + 3 * lohi];
            // quantities that do not depend on q
            double vel[3] = {0.0, 0.0, 0.0}, R = 1.0, T = 1.0 / 3.0, gamma = 5.0 / 3.0;
            double rho_bc = (bc == 3) ? 1.0 : 0.0;
            if (bc == 2 || bc == 3) vel_bc_op(B, L.dhi, i + L.lo[0], j + L.lo[1], k + L.lo[2], rho_bc, vel, R, T, gamma);
            for (int q = 0; q < NQ; ++q) {
                const int e[3] = {c_dir.ex[q], c_dir.ey[q], c_dir.ez[q]};
                const int in = i + e[0], jn = j + e[1], kn = k + e[2];
                const bool inside = in >= Rg.in_lo[0] && in <= Rg.in_hi[0] && jn >= Rg.in_lo[1] && jn <= Rg.in_hi[1] &&
                                    kn >= Rg.in_lo[2] && kn <= Rg.in_hi[2];
                double* dst = data + (long long)q * L.sq + c;
                if (!inside) {
                    *dst = -1.0;  // BC.H:463-466
                    continue;
                }
                const long long cn = L.cell(in, jn, kn);
                if (bc == 1) {  // NOSLIPWALL, BC.H:81
                    *dst = data[(long long)c_dir.opp[q] * L.sq + cn];
                } else if (bc == 2) {  // VELOCITY, BC.H:394-419
                    *dst = energy ? geq_state(rho_bc, vel, T, R, gamma, q) : feq_std(rho_bc, vel, R * T, q);
                } else if (bc == 3) {  // PRESSURE, BC.H:421-450, 242-268, 298-320
                    double vb[3] = {0.0, 0.0, 0.0};
                    if (energy) {
                        vb[idir] = ndir * (1.0 - 1.0 / rho_bc);
                        *dst = geq_state(rho_bc, vb, T, R, gamma, q);
                    } else {
                        double rho_out = 0.0, rho_tan = 0.0;
                        for (int qq = 0; qq < NQ; ++qq) {
                            const int bq = c_dir.opp[qq];
                            const int ei[3] = {c_dir.ex[bq], c_dir.ey[bq], c_dir.ez[bq]};
                            const long long cs = cn + c_dir.ex[qq] + c_dir.ey[qq] * L.px + c_dir.ez[qq] * L.sz;
                            if (ei[idir] == -ndir) {
                                rho_out += 2.0 * data[(long long)bq * L.sq + cs];
                            } else if (ei[idir] == 0) {
                                rho_tan += data[(long long)bq * L.sq + cs];
                            }
                        }
                        vb[idir] = ndir * (1.0 - (rho_out + rho_tan) / rho_bc);
                        *dst = feq_std(rho_bc, vb, R * T, q);
                    }
                } else if (bc == 5) {  // OUTFLOW_ZEROTH_ORDER, BC.H:339-341
                    const long long step = idir == 0 ? 1 : idir == 1 ? L.px : L.sz;
                    *dst = data[(long long)q * L.sq + c + ndir * step];
                } else if (bc == 6) {  // SLIPWALLXNORMAL, BC.H:100
                    *dst = data[(long long)c_dir.mx[q] * L.sq + cn];
                } else if (bc == 7) {
                    *dst = data[(long long)c_dir.my[q] * L.sq + cn];
                } else if (bc == 8) {
                    *dst = data[(long long)c_dir.mz[q] * L.sq + cn];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// flags
// ---------------------------------------------------------------------------
__device__ __forceinline__ int fab_fluid(const int32_t* fab, const Layout& L, int ng, int i, int j, int k)
{
    // FAB layout over the valid box grown by ng; cells beyond it count as fluid
    if (i < -ng || i >= L.nx + ng || j < -ng || j >= L.ny + ng || k < -ng || k >= L.nz + ng) return 1;
    const long long sx = L.nx + 2 * ng, sy = L.ny + 2 * ng;
    return fab[(i + ng) + (j + ng) * sx + (long long)(k + ng) * sx * sy];
}

__global__ void __launch_bounds__(128) k_flags(const int32_t* __restrict__ fab, int ng, uint32_t* __restrict__ nbr,
                                               uint8_t* __restrict__ flag, Layout L, BcInfo B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x - GX;
    const int j = blockIdx.y - GY;
    const int k = blockIdx.z - GZ;
    if (i > L.nx - 1 + GX) return;
    const long long c = L.cell(i, j, k);
    uint32_t m = 0;
    for (int q = 0; q < NQ; ++q) {
        const int fl = fab ? fab_fluid(fab, L, ng, i - c_dir.ex[q], j - c_dir.ey[q], k - c_dir.ez[q]) : 1;
        m |= (fl == 1 ? 1u : 0u) << q;
    }
    unsigned fb = (m & 1u) ? FLAG_FLUID : 0u;
    const int iv[3] = {i, j, k};
    const unsigned bp[3] = {GRAD_PX, GRAD_PY, GRAD_PZ}, bm[3] = {GRAD_MX, GRAD_MY, GRAD_MZ};
    bool any_fluid_face = false;
    for (int d = 0; d < 3; ++d) {
        int p[3] = {i, j, k}, mm[3] = {i, j, k};
        p[d] += 1;
        mm[d] -= 1;
        const int fp = fab ? fab_fluid(fab, L, ng, p[0], p[1], p[2]) : 1;
        const int fm = fab ? fab_fluid(fab, L, ng, mm[0], mm[1], mm[2]) : 1;
        // gradient(): neighbour inside the DOMAIN box (not periodic-grown) and fluid
        // (Utilities.H:292-309 with dbox = geom.Domain(), LBM.cpp:968)
        if (in_dom(L, d, iv[d] + 1) && in_dom(L, (d + 1) % 3, iv[(d + 1) % 3]) && in_dom(L, (d + 2) % 3, iv[(d + 2) % 3]) &&
            fp == 1)
            fb |= bp[d];
        if (in_dom(L, d, iv[d] - 1) && in_dom(L, (d + 1) % 3, iv[(d + 1) % 3]) && in_dom(L, (d + 2) % 3, iv[(d + 2) % 3]) &&
            fm == 1)
            fb |= bm[d];
        any_fluid_face |= (fp != 0) || (fm != 0);
    }
    // eb_boundary: solid cell with a fluid face neighbour (LBM.cpp:1236-1259)
    if (!(m & 1u) && any_fluid_face) fb |= FLAG_EBB;
    nbr[c] = m;
    flag[c] = (uint8_t)fb;
}

// ---------------------------------------------------------------------------
// is_fluid of the analytic bodies of the shipped decks, evaluated on the device: a cell is
// solid when all 8 of its corners lie inside the body (EB2's covered-cell rule for these implicit functions; the
// same rule as the host mirror marbles_b200/geometry.py, which tests/ check against the reference's own is_fluid).
// Written in FAB layout over the box grown by ng, so that k_flags consumes it like an uploaded m_is_fluid: ghost
// cells of periodic directions take the periodic image (m_is_fluid.FillBoundary), ghost cells beyond a non-periodic
// face the geometry evaluated beyond the domain (LBM.cpp:1222-1234).  Explicit roundings: no FMA, same bits as numpy.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool node_in_body(const BodyInfo& G, double x, double y, double z)
{
    const double p[3] = {__dsub_rn(x, G.a[0]), __dsub_rn(y, G.a[1]), __dsub_rn(z, G.a[2])};
    bool inside;
    if (G.kind == 1) {  // sphere: a = centre, r
        inside = __dadd_rn(__dadd_rn(__dmul_rn(p[0], p[0]), __dmul_rn(p[1], p[1])), __dmul_rn(p[2], p[2])) < __dmul_rn(G.r, G.r);
    } else if (G.kind == 2) {  // cylinder: a = centre, r, height h along `axis`
        double rad2 = 0.0;
        for (int d = 0; d < 3; ++d)
            if (d != G.axis) rad2 = __dadd_rn(rad2, __dmul_rn(p[d], p[d]));
        inside = rad2 < __dmul_rn(G.r, G.r);
        if (G.h > 0.0) inside = inside && fabs(p[G.axis]) < __dmul_rn(0.5, G.h);
    } else {  // box: a = lo, b = hi
        inside = x > G.a[0] && x < G.b[0] && y > G.a[1] && y < G.b[1] && z > G.a[2] && z < G.b[2];
    }
    return G.fluid_inside ? !inside : inside;  // true: the node is solid
}

__global__ void __launch_bounds__(128) k_body_is_fluid(int32_t* __restrict__ fab, int ng, Layout L, BcInfo B, BodyInfo G)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x - ng;
    const int j = blockIdx.y - ng;
    const int k = blockIdx.z - ng;
    if (i >= L.nx + ng) return;
    int gidx[3] = {i + L.lo[0], j + L.lo[1], k + L.lo[2]};
    for (int d = 0; d < 3; ++d)
        if (B.periodic[d]) {
            const int n = L.dhi[d] - L.dlo[d] + 1;
            gidx[d] = L.dlo[d] + (((gidx[d] - L.dlo[d]) % n) + n) % n;
        }
    bool all_solid = true;
    for (int c = 0; c < 8; ++c) {
        const double x = __dadd_rn(B.prob_lo[0], __dmul_rn((double)(gidx[0] + (c & 1)), B.dx[0]));
        const double y = __dadd_rn(B.prob_lo[1], __dmul_rn((double)(gidx[1] + ((c >> 1) & 1)), B.dx[1]));
        const double z = __dadd_rn(B.prob_lo[2], __dmul_rn((double)(gidx[2] + ((c >> 2) & 1)), B.dx[2]));
        all_solid = all_solid && node_in_body(G, x, y, z);
    }
    const long long sx = L.nx + 2 * ng, sy = L.ny + 2 * ng;
    fab[(i + ng) + (j + ng) * sx + (long long)(k + ng) * sx * sy] = all_solid ? 0 : 1;
}

__global__ void k_fill(double* __restrict__ p, long long n, double v)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < n; t += stride) p[t] = v;
}

// ---------------------------------------------------------------------------
// initial conditions (IC.H), valid cells + ghost cells of the padded box
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_initialize(double* __restrict__ f, double* __restrict__ g,
                                                    const uint8_t* __restrict__ flag, Layout L, BcInfo B, IcInfo I)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x - GX;
    const int j = blockIdx.y - GY;
    const int k = blockIdx.z - GZ;
    if (i > L.nx - 1 + GX) return;
    const long long c = L.cell(i, j, k);
    const int gi = i + L.lo[0], gj = j + L.lo[1], gk = k + L.lo[2];
    double rho, vel[3], T, R, gamma;
    ic_state(I, B, gi, gj, gk, rho, vel, T, R, gamma);
    const bool solid = !(flag[c] & FLAG_FLUID);
    for (int q = 0; q < NQ; ++q) {
        f[(long long)q * L.sq + c] = solid ? 0.0 : feq_std(rho, vel, R * T, q);
        g[(long long)q * L.sq + c] = solid ? 0.0 : geq_state(rho, vel, T, R, gamma, q);
    }
}

// ---------------------------------------------------------------------------
// z-halo pack / unpack: GZ whole padded planes per component are contiguous
// ---------------------------------------------------------------------------
__global__ void k_halo_copy(double* __restrict__ f, double* __restrict__ g, double* __restrict__ buf, Layout L, int k0,
                            int to_buf)
{
    const long long chunk = (long long)GZ * L.sz;
    const long long total = 2LL * NQ * chunk;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long base = (long long)(k0 + GZ) * L.sz;
    for (; t < total; t += stride) {
        const long long w = t % chunk;
        const int q = (int)((t / chunk) % NQ);
        const int lat = (int)(t / (chunk * NQ));
        double* p = (lat ? g : f) + (long long)q * L.sq + base + w;
        if (to_buf)
            buf[t] = *p;
        else
            *p = buf[t];
    }
}

// ===========================================================================
// launchers
// ===========================================================================
static inline dim3 grid3(const Layout& L, int bx, int ex_ = 0, int ey_ = 0, int ez_ = 0)
{
    return dim3((L.nx + 2 * ex_ + bx - 1) / bx, L.ny + 2 * ey_, L.nz + 2 * ez_);
}
static inline int block_x(const Layout& L) { return L.nx >= 128 ? 128 : L.nx > 32 ? 64 : 32; }

int launch_flags(const Layout& L, const BcInfo& B, const int32_t* fab, int ng, uint32_t* nbr, uint8_t* flag,
                 cudaStream_t st)
{
    const int bx = block_x(L);
    k_flags<<<grid3(L, bx, GX, GY, GZ), bx, 0, st>>>(fab, ng, nbr, flag, L, B);
    return 1;
}
int launch_flags_all_fluid(const Layout& L, const BcInfo& B, uint32_t* nbr, uint8_t* flag, cudaStream_t st)
{
    return launch_flags(L, B, nullptr, 0, nbr, flag, st);
}

int launch_body_is_fluid(const Layout& L, const BcInfo& B, const BodyInfo& G, int32_t* fab, int ng, cudaStream_t st)
{
    const int bx = block_x(L);
    k_body_is_fluid<<<dim3((L.nx + 2 * ng + bx - 1) / bx, L.ny + 2 * ng, L.nz + 2 * ng), bx, 0, st>>>(fab, ng, L, B, G);
    return 1;
}

int launch_fill(double* p, long long n, double v, cudaStream_t st)
{
    k_fill<<<148 * 8, 256, 0, st>>>(p, n, v);
    return 1;
}

int launch_initialize(const Layout& L, const BcInfo& B, const IcInfo& I, const uint8_t* flag, double* f, double* g,
                      cudaStream_t st)
{
    const int bx = block_x(L);
    k_initialize<<<grid3(L, bx, GX, GY, GZ), bx, 0, st>>>(f, g, flag, L, B, I);
    return 1;
}

static bool side_range(const Layout& L, const BcInfo& B, int d, int side, int& a0, int& a1)
{
    const int n[3] = {L.nx, L.ny, L.nz};
    const int gw[3] = {GX, GY, GZ};
    const int plo = -gw[d], phi = n[d] - 1 + gw[d];
    const int big = 1 << 28;
    const int glo = B.periodic[d] ? -big : L.dlo[d] - L.lo[d];
    const int ghi = B.periodic[d] ? big : L.dhi[d] - L.lo[d];
    if (side < 0) {
        a0 = plo;
        a1 = glo - 1 < phi ? glo - 1 : phi;
    } else if (side > 0) {
        a0 = ghi + 1 > plo ? ghi + 1 : plo;
        a1 = phi;
    } else {
        a0 = glo > plo ? glo : plo;
        a1 = ghi < phi ? ghi : phi;
    }
    return a0 <= a1;
}

static int launch_bc_region(const Layout& L, const BcInfo& B, double* f, double* g, int sx, int sy, int sz,
                            cudaStream_t st)
{
    Region R;
    const int s[3] = {sx, sy, sz};
    for (int d = 0; d < 3; ++d)
        if (!side_range(L, B, d, s[d], R.lo[d], R.hi[d])) return 0;
    const int n[3] = {L.nx, L.ny, L.nz};
    for (int d = 0; d < 3; ++d) {
        R.in_lo[d] = 0;
        R.in_hi[d] = n[d] - 1;
    }
    // the z-ghost planes that belong to a neighbouring rank are valid cells of the level: BC values
    // that stream into them are needed for the q-correction of the first ghost plane
    if (L.lo[2] > L.dlo[2]) R.in_lo[2] = -GZ;
    if (L.lo[2] + L.nz - 1 < L.dhi[2]) R.in_hi[2] = L.nz - 1 + GZ;
    const long long cells = (long long)(R.hi[0] - R.lo[0] + 1) * (R.hi[1] - R.lo[1] + 1) * (R.hi[2] - R.lo[2] + 1);
    dim3 grid((unsigned)((cells + 63) / 64), 2, 1);
    k_bc_region<<<grid, 64, 0, st>>>(f, g, L, B, R);
    return 1;
}

int launch_ghost_fill(const Layout& L, const BcInfo& B, double* f, double* g, bool local_z, bool do_prepass,
                      bool do_periodic, cudaStream_t st)
{
    int nl = 0;
    const int bx = block_x(L);
    const Shell S = make_shell(L);
    const bool all_periodic = B.periodic[0] && B.periodic[1] && B.periodic[2];
    if (do_periodic && local_z && B.periodic[2]) {
        k_zwrap<<<dim3((L.nx + bx - 1) / bx, L.ny, 2 * GZ), bx, 0, st>>>(f, g, L);
        ++nl;
    }
    const unsigned sb = (unsigned)((S.off[6] + 127) / 128);
    if (do_prepass && !all_periodic) {
        k_prepass<<<sb, 128, 0, st>>>(f, g, L, S, B);
        ++nl;
    }
    if (do_periodic && (B.periodic[0] || B.periodic[1])) {
        k_xywrap<<<sb, 128, 0, st>>>(f, g, L, S, B);
        ++nl;
    }
    if (all_periodic) return nl;  // PhysBCFunct::operator() returns early (AMReX_PhysBCFunct.H:202)
    // faces (xlo ylo zlo xhi yhi zhi), edges, corners: AMReX_PhysBCFunct.H:593-678
    nl += launch_bc_region(L, B, f, g, -1, 0, 0, st);
    nl += launch_bc_region(L, B, f, g, 0, -1, 0, st);
    nl += launch_bc_region(L, B, f, g, 0, 0, -1, st);
    nl += launch_bc_region(L, B, f, g, +1, 0, 0, st);
    nl += launch_bc_region(L, B, f, g, 0, +1, 0, st);
    nl += launch_bc_region(L, B, f, g, 0, 0, +1, st);
    for (int s1 = -1; s1 <= 1; s1 += 2)
        for (int s0 = -1; s0 <= 1; s0 += 2) nl += launch_bc_region(L, B, f, g, s0, s1, 0, st);
    for (int s1 = -1; s1 <= 1; s1 += 2)
        for (int s0 = -1; s0 <= 1; s0 += 2) nl += launch_bc_region(L, B, f, g, s0, 0, s1, st);
    for (int s1 = -1; s1 <= 1; s1 += 2)
        for (int s0 = -1; s0 <= 1; s0 += 2) nl += launch_bc_region(L, B, f, g, 0, s0, s1, st);
    for (int s2 = -1; s2 <= 1; s2 += 2)
        for (int s1 = -1; s1 <= 1; s1 += 2)
            for (int s0 = -1; s0 <= 1; s0 += 2) nl += launch_bc_region(L, B, f, g, s0, s1, s2, st);
    return nl;
}

int launch_qcorr(const Layout& L, const Phys& P, const double* fin, const double* gin, const uint32_t* nbr, double* qc,
                 bool pull, cudaStream_t st, int ka, int kb)
{
    const int bx = block_x(L);
    // q-corrections are needed on the valid cells and, where the box borders another rank in z,
    // on the first ghost plane (recomputed from the two exchanged planes instead of a second exchange)
    int k0 = (L.lo[2] > L.dlo[2]) ? -1 : 0;
    int k1 = (L.lo[2] + L.nz - 1 < L.dhi[2]) ? L.nz : L.nz - 1;
    if (kb > ka) k0 = ka, k1 = kb - 1;  // explicit plane range [ka, kb)
    dim3 grid((L.nx + bx - 1) / bx, L.ny, k1 - k0 + 1);
    if (pull)
        k_qcorr<true><<<grid, bx, 0, st>>>(fin, gin, nbr, qc, L, P, k0);
    else
        k_qcorr<false><<<grid, bx, 0, st>>>(fin, gin, nbr, qc, L, P, k0);
    return 1;
}

int launch_collide(const Layout& L, const Phys& P, const double* fin, const double* gin, double* fout, double* gout,
                   const uint32_t* nbr, const uint8_t* flag, const double* qc, double* macro, bool pull,
                   cudaStream_t st, int ka, int kb)
{
    // the hot case -- fused pull + collide without macrodata output -- runs the lean kernel (bit-identical
    // results, 7 % faster: g staged through shared memory, 32-bit addressing); MBL_PLAIN_COLLIDE=1 keeps k_collide
    const char* pc = getenv("MBL_PLAIN_COLLIDE");
    const bool plain_only = pc && atoi(pc) != 0;
    if (pull && !macro && !plain_only && L.sq * 8 < (1LL << 32))
        return launch_collide_lean(L, P, 3, fin, gin, fout, gout, nbr, flag, qc, st, ka, kb);
    const int bx = block_x(L);
    dim3 grid = grid3(L, bx);
    int k0 = 0;
    if (kb > ka) k0 = ka, grid.z = kb - ka;  // explicit plane range [ka, kb)
    if (pull) {
        if (macro)
            k_collide<true, true><<<grid, bx, 0, st>>>(fin, gin, fout, gout, nbr, flag, qc, macro, L, P, k0);
        else
            k_collide<true, false><<<grid, bx, 0, st>>>(fin, gin, fout, gout, nbr, flag, qc, macro, L, P, k0);
    } else {
        if (macro)
            k_collide<false, true><<<grid, bx, 0, st>>>(fin, gin, fout, gout, nbr, flag, qc, macro, L, P, k0);
        else
            k_collide<false, false><<<grid, bx, 0, st>>>(fin, gin, fout, gout, nbr, flag, qc, macro, L, P, k0);
    }
    return 1;
}

CarryPlan make_carry_plan(const Layout& L, int own, int ky)
{
    CarryPlan C;
    C.own = (own == 28) ? 28 : 30;
#ifdef MBL_EXPERIMENTS
    if (own == 32) C.own = 32;  // no halo lanes: timing only, the x sums are wrong at the strip edges
#endif
    C.halo = (32 - C.own) / 2;
    C.ky = ky < 1 ? 1 : (ky > L.ny ? L.ny : ky);
    C.nxc = (L.nx + C.own - 1) / C.own;
    C.prefetch = 0;
    C.esz8 = 0;
    if (const char* e = getenv("MBL_PREFETCH")) C.prefetch = atoi(e);
    return C;
}

int launch_collide_lean(const Layout& L, const Phys& P, int min_blocks, const double* fin, const double* gin,
                        double* fout, double* gout, const uint32_t* nbr, const uint8_t* flag, const double* qc,
                        cudaStream_t st, int ka, int kb)
{
    if (L.sq * 8 >= (1LL << 32)) return -1;  // 32-bit byte offsets inside a component
    CarryPtrs A;
    for (int q = 0; q < NQ; ++q) {
        A.fin[q] = fin + (long long)q * L.sq;
        A.gin[q] = gin + (long long)q * L.sq;
        A.fout[q] = fout + (long long)q * L.sq;
        A.gout[q] = gout + (long long)q * L.sq;
    }
    for (int d = 0; d < 3; ++d) A.qc[d] = qc + (long long)d * L.sq;
    for (int w = 0; w < CARRY_WORDS; ++w) A.part[w] = nullptr;
    for (int e = 0; e < CARRY_EDGE_WORDS; ++e) A.edge[e] = nullptr;
    dim3 grid((L.nx + 127) / 128, L.ny, L.nz);
    int k0 = 0;
    if (kb > ka) k0 = ka, grid.z = kb - ka;
    const size_t sm = (size_t)NQ * 128 * 8;
    if (min_blocks >= 5)
        k_collide_lean<5><<<grid, 128, sm, st>>>(A, nbr, flag, L, P, k0);
    else if (min_blocks == 4)
        k_collide_lean<4><<<grid, 128, sm, st>>>(A, nbr, flag, L, P, k0);
    else
        k_collide_lean<3><<<grid, 128, sm, st>>>(A, nbr, flag, L, P, k0);
    return 1;
}

long long carry_edge_plane(const Layout& L, int W) { return L.px * (long long)((L.ny + W - 1) / W); }
int carry_tile_rows(int rows) { return rows == 4 ? 4 : rows == 8 ? 8 : rows == 12 ? 12 : 6; }

int launch_collide_tile(const Layout& L, const Phys& P, const CarryPlan& C, int rows, const double* fin,
                        const double* gin, double* fout, double* gout, const uint32_t* nbr, const uint8_t* flag,
                        const double* qc, double* part, double* edge, cudaStream_t st, int ka, int kb)
{
    if (L.sq * 8 >= (1LL << 32)) return -1;  // 32-bit byte offsets inside a component
    CarryPtrs A;
    for (int q = 0; q < NQ; ++q) {
        A.fin[q] = fin + (long long)q * L.sq;
        A.gin[q] = gin + (long long)q * L.sq;
        A.fout[q] = fout + (long long)q * L.sq;
        A.gout[q] = gout + (long long)q * L.sq;
    }
    for (int d = 0; d < 3; ++d) A.qc[d] = qc + (long long)d * L.sq;
    for (int w = 0; w < CARRY_WORDS; ++w) A.part[w] = part + (long long)w * L.sq;
    const int W = carry_tile_rows(rows);
    CarryPlan Ce = C;
    const long long esz = carry_edge_plane(L, W);
    Ce.esz8 = (unsigned)(esz * 8);
    for (int e = 0; e < CARRY_EDGE_WORDS; ++e) A.edge[e] = edge + (long long)e * esz * (L.nz + 2 * GZ);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_collide_tile<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, NQ * 32 * 4 * 8);
        cudaFuncSetAttribute(k_collide_tile<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, NQ * 32 * 6 * 8);
        cudaFuncSetAttribute(k_collide_tile<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, NQ * 32 * 8 * 8);
        cudaFuncSetAttribute(k_collide_tile<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, NQ * 32 * 12 * 8);
        attr_done = true;
    }
    dim3 grid(C.nxc, (L.ny + W - 1) / W, L.nz);
    int k0 = 0;
    if (kb > ka) k0 = ka, grid.z = kb - ka;  // explicit plane range [ka, kb)
    const size_t sm = (size_t)NQ * 32 * W * 8;
    if (W == 4)
        k_collide_tile<4><<<grid, 32 * 4, sm, st>>>(A, nbr, flag, L, P, Ce, k0);
    else if (W == 6)
        k_collide_tile<6><<<grid, 32 * 6, sm, st>>>(A, nbr, flag, L, P, Ce, k0);
    else if (W == 12)
        k_collide_tile<12><<<grid, 32 * 12, sm, st>>>(A, nbr, flag, L, P, Ce, k0);
    else
        k_collide_tile<8><<<grid, 32 * 8, sm, st>>>(A, nbr, flag, L, P, Ce, k0);
    return 1;
}

int launch_collide_tile_pair(const Layout& L, const Phys& P, const CarryPlan& C, const double* fin, const double* gin,
                             double* fout, double* gout, const uint32_t* nbr, const uint8_t* flag, const double* qc,
                             double* part, double* edge, cudaStream_t st, int ka, int kb)
{
    if (L.sq * 8 >= (1LL << 32)) return -1;
    constexpr int W = 6;
    if (kb <= ka) ka = 0, kb = L.nz;
    if ((ka & 1) || ((kb - ka) & 1)) return -2;  // pairs (2b, 2b+1)
    CarryPtrs A;
    for (int q = 0; q < NQ; ++q) {
        A.fin[q] = fin + (long long)q * L.sq;
        A.gin[q] = gin + (long long)q * L.sq;
        A.fout[q] = fout + (long long)q * L.sq;
        A.gout[q] = gout + (long long)q * L.sq;
    }
    for (int d = 0; d < 3; ++d) A.qc[d] = qc + (long long)d * L.sq;
    for (int w = 0; w < CARRY_WORDS; ++w) A.part[w] = part + (long long)w * L.sq;
    CarryPlan Ce = C;
    const long long esz = carry_edge_plane(L, W);
    Ce.esz8 = (unsigned)(esz * 8);
    for (int e = 0; e < CARRY_EDGE_WORDS; ++e) A.edge[e] = edge + (long long)e * esz * (L.nz + 2 * GZ);
    const size_t sm = (size_t)(NQ + PAIR_KEEP_SLOTS) * 32 * W * 8;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_collide_tile_pair<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        attr_done = true;
    }
    const dim3 grid(C.nxc, (L.ny + W - 1) / W, (kb - ka) / 2);
    k_collide_tile_pair<6><<<grid, 32 * W, sm, st>>>(A, nbr, flag, L, P, Ce, ka);
    return 1;
}

// variant 9.  Returns the number of kernels, -1 if a component exceeds 4 GB, -2 if a chunk would hold one plane
template <int W, bool PIPE, int ABL = 0, int ORD = 0>
static void run_collide_tile_march(dim3 grid, cudaStream_t st, const CarryPtrs& A, const MarchOut& Q, const uint32_t* nbr,
                                   const uint8_t* flag, const Layout& L, const Phys& P, const CarryPlan& Ce, int ka, int kb, int zm)
{
    const size_t sm = (size_t)march_slots(PIPE) * 32 * W * 8;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_collide_tile_march<W, PIPE, ABL, ORD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        attr_done = true;
    }
    k_collide_tile_march<W, PIPE, ABL, ORD><<<grid, 32 * W, sm, st>>>(A, Q, nbr, flag, L, P, Ce, ka, kb, zm);
}

// rows per CTA the z-march kernels are built for: 6 (plain); 4 and 8 (pipelined: a measured negative result,
// 30.5 / 29.1 ms against 25.8 ms at 512^3 with 8 instead of 12 warps per SM -- built only with MBL_EXPERIMENTS)
bool march_rows_supported(int W, bool pipe)
{
#ifdef MBL_EXPERIMENTS
    return pipe ? (W == 4 || W == 8) : W == 6;
#else
    return !pipe && W == 6;
#endif
}

int launch_collide_tile_march(const Layout& L, const Phys& P, const CarryPlan& C, int W, bool pipe, int zm, const double* fin,
                              const double* gin, double* fout, double* gout, const uint32_t* nbr, const uint8_t* flag,
                              const double* qc, double* qc_next, double* part, double* edge, cudaStream_t st, int ka, int kb)
{
    if (L.sq * 8 >= (1LL << 32)) return -1;
    if (!march_rows_supported(W, pipe)) return -3;
    if (kb <= ka) ka = 0, kb = L.nz;
    if (zm < 2) zm = 2;
    if ((kb - ka) % zm == 1 || kb - ka < 2) return -2;  // a one-plane chunk would need both send words
    CarryPtrs A;
    MarchOut Q;
    for (int q = 0; q < NQ; ++q) {
        A.fin[q] = fin + (long long)q * L.sq;
        A.gin[q] = gin + (long long)q * L.sq;
        A.fout[q] = fout + (long long)q * L.sq;
        A.gout[q] = gout + (long long)q * L.sq;
    }
    for (int d = 0; d < 3; ++d) A.qc[d] = qc + (long long)d * L.sq, Q.qcn[d] = qc_next + (long long)d * L.sq;
    for (int w = 0; w < CARRY_WORDS; ++w) A.part[w] = part + (long long)w * L.sq;
    CarryPlan Ce = C;
    const long long esz = carry_edge_plane(L, W);
    Ce.esz8 = (unsigned)(esz * 8);
    for (int e = 0; e < CARRY_EDGE_WORDS; ++e) A.edge[e] = edge + (long long)e * esz * (L.nz + 2 * GZ);
    const dim3 grid(C.nxc, (L.ny + W - 1) / W, (kb - ka + zm - 1) / zm);
#ifdef MBL_EXPERIMENTS
    static const int abl = getenv("MBL_ABLATE") ? atoi(getenv("MBL_ABLATE")) : 0;  // timing only: wrong results
    if (!pipe && abl == 1) { run_collide_tile_march<6, false, 1>(grid, st, A, Q, nbr, flag, L, P, Ce, ka, kb, zm); return 1; }
    if (!pipe && abl == 2) { run_collide_tile_march<6, false, 2>(grid, st, A, Q, nbr, flag, L, P, Ce, ka, kb, zm); return 1; }
    if (!pipe && abl == 4) { run_collide_tile_march<6, false, 4>(grid, st, A, Q, nbr, flag, L, P, Ce, ka, kb, zm); return 1; }
    if (!pipe && abl == 6) { run_collide_tile_march<6, false, 6>(grid, st, A, Q, nbr, flag, L, P, Ce, ka, kb, zm); return 1; }
    static const int ord = getenv("MBL_LOADORDER") ? atoi(getenv("MBL_LOADORDER")) : 0;  // results identical
    if (!pipe && ord == 1) { run_collide_tile_march<6, false, 0, 1>(grid, st, A, Q, nbr, flag, L, P, Ce, ka, kb, zm); return 1; }
    if (!pipe && ord == 2) { run_collide_tile_march<6, false, 0, 2>(grid, st, A, Q, nbr, flag, L, P, Ce, ka, kb, zm); return 1; }
#endif
    if (!pipe) run_collide_tile_march<6, false>(grid, st, A, Q, nbr, flag, L, P, Ce, ka, kb, zm);
#ifdef MBL_EXPERIMENTS
    else if (W == 4) run_collide_tile_march<4, true>(grid, st, A, Q, nbr, flag, L, P, Ce, ka, kb, zm);
    else run_collide_tile_march<8, true>(grid, st, A, Q, nbr, flag, L, P, Ce, ka, kb, zm);
#endif
    return 1;
}

int launch_qcorr_combine_march(const Layout& L, const Phys& P, const double* fin, const double* gin, const uint32_t* nbr,
                               const double* part, const double* edge, int W, const signed char* zpos, double* qc,
                               cudaStream_t st, int ka, int kb)
{
    const int bx = block_x(L);
    int k0 = (L.lo[2] > L.dlo[2]) ? -1 : 0;
    int k1 = (L.lo[2] + L.nz - 1 < L.dhi[2]) ? L.nz : L.nz - 1;
    if (kb > ka) k0 = ka, k1 = kb - 1;
    dim3 grid((L.nx + bx - 1) / bx, L.ny, k1 - k0 + 1);
    k_qcorr_combine_march<<<grid, bx, 0, st>>>(fin, gin, nbr, part, edge, zpos, W, carry_edge_plane(L, W), qc, L, P, k0);
    return 1;
}

int launch_qcorr_combine_pair(const Layout& L, const Phys& P, const double* fin, const double* gin, const uint32_t* nbr,
                              const double* part, const double* edge, double* qc, cudaStream_t st, int ka, int kb)
{
    const int bx = block_x(L);
    int k0 = (L.lo[2] > L.dlo[2]) ? -1 : 0;
    int k1 = (L.lo[2] + L.nz - 1 < L.dhi[2]) ? L.nz : L.nz - 1;
    if (kb > ka) k0 = ka, k1 = kb - 1;
    dim3 grid((L.nx + bx - 1) / bx, L.ny, k1 - k0 + 1);
    k_qcorr_combine_pair<<<grid, bx, 0, st>>>(fin, gin, nbr, part, edge, 6, carry_edge_plane(L, 6), qc, L, P, k0);
    return 1;
}

int launch_qcorr_combine(const Layout& L, const Phys& P, const double* fin, const double* gin, const uint32_t* nbr,
                         const double* part, const double* edge, int edge_rows, double* qc, cudaStream_t st, int ka,
                         int kb)
{
    const int bx = block_x(L);
    int k0 = (L.lo[2] > L.dlo[2]) ? -1 : 0;
    int k1 = (L.lo[2] + L.nz - 1 < L.dhi[2]) ? L.nz : L.nz - 1;
    if (kb > ka) k0 = ka, k1 = kb - 1;  // explicit plane range [ka, kb)
    dim3 grid((L.nx + bx - 1) / bx, L.ny, k1 - k0 + 1);
    const int W = edge ? carry_tile_rows(edge_rows) : 0;
    k_qcorr_combine<<<grid, bx, 0, st>>>(fin, gin, nbr, part, edge, W, W ? carry_edge_plane(L, W) : 0, qc, L, P, k0);
    return 1;
}

int launch_stream(const Layout& L, const double* fin, const double* gin, double* fout, double* gout,
                  const uint32_t* nbr, cudaStream_t st)
{
    const int bx = block_x(L);
    k_stream<<<grid3(L, bx), bx, 0, st>>>(fin, gin, fout, gout, nbr, L);
    return 1;
}

int launch_macrodata(const Layout& L, const Phys& P, const double* f, const double* g, const uint8_t* flag,
                     double* macro, cudaStream_t st)
{
    const int bx = block_x(L);
    k_macrodata<<<grid3(L, bx), bx, 0, st>>>(f, g, flag, macro, L, P);
    return 1;
}

int launch_derived(const Layout& L, const Phys& P, const uint8_t* flag, const double* macro, double* derived,
                   cudaStream_t st, int with_dq, int zghost)
{
    const int bx = block_x(L);
    k_derived<<<grid3(L, bx), bx, 0, st>>>(flag, macro, derived, L, P, with_dq, zghost);
    return 1;
}

int launch_eb_forces(const Layout& L, const double* f, const uint8_t* flag, double* d_out3, cudaStream_t st)
{
    const int bx = block_x(L);
    cudaMemsetAsync(d_out3, 0, 3 * sizeof(double), st);
    k_eb_forces<<<grid3(L, bx), bx, 0, st>>>(f, flag, d_out3, L);
    return 1;
}

// Lean z-halo: only the populations a neighbour's pull can reach.  Towards the neighbour on `side` (e_z = -1 for the
// low side, +1 for the high side) the plane next to it sends the 18 populations with e_z in {towards, 0} (the pull of
// its first owned plane and the q-correction of its first ghost plane), the second plane the 9 with e_z = towards
// (q-correction of the first ghost plane): 27 instead of 54 plane-components per lattice.  Valid where no bounce-back
// and no boundary ghost value can ask for the others (all-fluid, all-periodic levels); the caller decides.
struct LeanList {
    signed char q[NQ];      // component
    signed char plane[NQ];  // 0: the plane next to the neighbour, 1: the second plane
};
static LeanList make_lean_list(int side)
{
    const int towards = side == 0 ? -1 : 1;
    LeanList t;
    int n = 0;
    for (int q = 0; q < NQ; ++q)
        if (ez(q) == towards || ez(q) == 0) t.q[n] = (signed char)q, t.plane[n++] = 0;
    for (int q = 0; q < NQ; ++q)
        if (ez(q) == towards) t.q[n] = (signed char)q, t.plane[n++] = 1;
    return t;
}
__global__ void k_halo_copy_lean(double* __restrict__ f, double* __restrict__ g, double* __restrict__ buf, Layout L, int k_near,
                                 int k_step, LeanList T, int to_buf)
{
    const long long total = 2LL * NQ * L.sz;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        const long long w = t % L.sz;
        const int item = (int)((t / L.sz) % NQ);
        const int lat = (int)(t / (L.sz * NQ));
        const int k = k_near + k_step * T.plane[item];
        double* p = (lat ? g : f) + (long long)T.q[item] * L.sq + (long long)(k + GZ) * L.sz + w;
        if (to_buf)
            buf[t] = *p;
        else
            *p = buf[t];
    }
}

int launch_halo_pack(const Layout& L, const double* f, const double* g, int side, double* buf, cudaStream_t st, bool lean)
{
    if (lean) {
        // side 0: my planes 0 (near the lower neighbour), 1; side 1: my planes nz-1, nz-2
        k_halo_copy_lean<<<148 * 4, 256, 0, st>>>(const_cast<double*>(f), const_cast<double*>(g), buf, L,
                                                  side == 0 ? 0 : L.nz - 1, side == 0 ? 1 : -1, make_lean_list(side), 1);
        return 1;
    }
    const int k0 = side == 0 ? 0 : L.nz - GZ;
    k_halo_copy<<<148 * 4, 256, 0, st>>>(const_cast<double*>(f), const_cast<double*>(g), buf, L, k0, 1);
    return 1;
}
int launch_halo_unpack(const Layout& L, double* f, double* g, int side, const double* buf, cudaStream_t st, bool lean)
{
    if (lean) {
        // my low ghost planes -1 (near), -2 hold what the lower neighbour packed for ITS high side, and vice versa
        k_halo_copy_lean<<<148 * 4, 256, 0, st>>>(f, g, const_cast<double*>(buf), L, side == 0 ? -1 : L.nz,
                                                  side == 0 ? -1 : 1, make_lean_list(1 - side), 0);
        return 1;
    }
    const int k0 = side == 0 ? -GZ : L.nz;
    k_halo_copy<<<148 * 4, 256, 0, st>>>(f, g, const_cast<double*>(buf), L, k0, 0);
    return 1;
}

#ifdef MBL_EXPERIMENTS
// negative-result kernels (variants 3-4), kept out of the shipped library: DESIGN.md section 3
#include "experiments/experiments.cuh"
#endif

}  // namespace mbl
