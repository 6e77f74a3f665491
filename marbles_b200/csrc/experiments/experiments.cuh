// experiments.cuh -- negative-result kernels of round 1, NOT part of the shipped library (build with
// MBL_EXPERIMENTS=1 to get variants 1-4 of mbl_set_variant back): the marching carry kernel (variant 4) and the
// persistent warp-autonomous two-job kernel with plain loads (variant 3).  Included at the end of kernels.cu,
// inside namespace mbl, because they reuse its per-cell device functions.  DESIGN.md section 3 has the numbers.
#pragma once

// Shared memory of k_collide_carry, private per thread (slot-major, so a warp's access to one slot is one
// conflict-free 256-byte row): the 27 pulled g populations of the current cell (they arrive by cp.async
// and never occupy registers: the register file holds f, the collision coefficients and this cell's
// partial sums, which is what lets three or four CTAs share an SM) and the two ring rows of sums.
constexpr int CARRY_G_SLOTS = NQ;       // g[q]
constexpr int CARRY_RING_SLOTS = 21;    // A: rho,jx,jy,e2 x 3 planes;  B: rho,jx,e2 x 3 planes
constexpr int CARRY_SMEM_BYTES = (CARRY_G_SLOTS + CARRY_RING_SLOTS) * 128 * 8;

template <int MINB>
__global__ void __launch_bounds__(128, MINB)
    k_collide_carry(const __grid_constant__ CarryPtrs A, const uint32_t* __restrict__ nbr,
                    const uint8_t* __restrict__ flag, const __grid_constant__ Layout L, const __grid_constant__ Phys P,
                    const __grid_constant__ CarryPlan C)
{
    extern __shared__ double smem[];
    double* const sg = smem + threadIdx.x;                          // sg[q * 128]
    double* const ring = smem + CARRY_G_SLOTS * 128 + threadIdx.x;  // ring[slot * 128]
    const unsigned sg_addr = (unsigned)__cvta_generic_to_shared(sg);
    const int lane = threadIdx.x & 31;
    const int xc = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (xc >= C.nxc) return;  // whole warp
    const int y0 = blockIdx.y * C.ky;
    const int ky = min(C.ky, L.ny - y0);
    const int k = blockIdx.z;
    const int i = xc * C.own - C.halo + lane;
    const bool own_lane = lane >= C.halo && lane < C.halo + C.own && i < L.nx;
    // column this lane collides: its own cell, the periodic image for a halo lane that hangs over a wrapped
    // edge, otherwise clamped (the lane then only keeps the shuffles uniform; whatever it contributes lands
    // in cells k_qcorr_combine does not take from the carried sums)
    int is = i;
    if (is < 0) is = L.wrap[0] ? is + L.nx : 0;
    if (is >= L.nx) is = (L.wrap[0] && is - L.nx < L.nx) ? is - L.nx : L.nx - 1;
    const unsigned FULL = 0xffffffffu;
    // All addressing is a per-component base pointer (kernel parameter space, uniform) plus an UNSIGNED 32-bit
    // byte offset: unsigned arithmetic keeps the compiler from widening the index sums to 64 bits, so an
    // access costs one 32-bit add and one 64-bit base add instead of a 64-bit multiply-add chain.  Pull
    // offsets as in pull_offsets(); the x and z ones do not change along the march.
    const unsigned px8 = (unsigned)L.px * 8u, sz8 = (unsigned)L.sz * 8u;
    unsigned xo[3], zo[3];
    xo[1] = zo[1] = 0u;
    xo[2] = (L.wrap[0] && is == 0) ? (unsigned)(L.nx - 1) * 8u : 0u - 8u;
    xo[0] = (L.wrap[0] && is == L.nx - 1) ? 0u - (unsigned)(L.nx - 1) * 8u : 8u;
    zo[2] = (L.wrap[2] && k == 0) ? (unsigned)(L.nz - 1) * sz8 : 0u - sz8;
    zo[0] = (L.wrap[2] && k == L.nz - 1) ? 0u - (unsigned)(L.nz - 1) * sz8 : sz8;
    const unsigned ccol = (unsigned)(is + OX) * 8u + (unsigned)(k + GZ) * sz8;
    auto ldb = [](const double* base, unsigned off) { return *(const double*)((const char*)base + off); };
    auto stb = [](double* base, unsigned off, double v) { *(double*)((char*)base + off) = v; };

    // ring: sums destined for row j-1 (A, complete after this row: slots 0..11 = [plane c][rho,jx,jy,e2]) and
    // for row j (B, slots 12..20 = [plane c][rho,jx,e2]; its jy is its rho as long as it holds e_y = +1 terms only)
#pragma unroll
    for (int t = 0; t < CARRY_RING_SLOTS; ++t) ring[t * 128] = 0.0;

    for (int jj = -1; jj <= ky; ++jj) {
        int j = y0 + jj;
        bool row_ok = true;
        if (j < 0) {
            row_ok = L.wrap[1];
            j += L.ny;
        } else if (j >= L.ny) {
            row_ok = L.wrap[1];
            j -= L.ny;
        }
        // this row's contributions to rows j-1 (TA), j (TB), j+1 (TC): [plane c][rho, jx] now, e2 in the g phase
        double TA[3][2] = {}, TB[3][2] = {}, TC[3][2] = {};
        double EA[3] = {}, EB[3] = {}, EC[3] = {};
        if (row_ok) {
            const unsigned c = ccol + (unsigned)(j + GY) * px8;
            unsigned yo[3];
            yo[1] = 0u;
            yo[2] = (L.wrap[1] && j == 0) ? (unsigned)(L.ny - 1) * px8 : 0u - px8;
            yo[0] = (L.wrap[1] && j == L.ny - 1) ? 0u - (unsigned)(L.ny - 1) * px8 : px8;
            unsigned cyz[3][3];
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int d = 0; d < 3; ++d) cyz[b][d] = c + yo[b] + zo[d];
            // g: global -> shared, asynchronously, no registers
            static_for<0, NQ>([&](auto qc_) {
                constexpr int Q = decltype(qc_)::value;
                cp_async8(sg_addr + Q * 128 * 8, (const char*)A.gin[Q] + (cyz[ey(Q) + 1][ez(Q) + 1] + xo[ex(Q) + 1]));
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
                            sg[Q * 128] = ldb(A.gin[opp(Q)], c);
                        }
                    });
                } else {
                    // solid cell: the streamed value is the -1 sentinel (LBM.cpp:565, 582) and collide skips it;
                    // with omega = 0 below the "relaxed" value is exactly -1 again
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        f[q] = -1.0;
                        sg[q * 128] = -1.0;
                    }
                }
            }
            const MomF mf = moments_f([&](int q) { return f[q]; });
            const MomG mg = moments_g([&](int q) { return sg[q * 128]; });
            const Prim s = primitives(mf.rho, mf.jx, mf.jy, mf.jz, mg.e2, P);
            const double dqx = one_sided_gradient(fb & GRAD_PX, fb & GRAD_MX, (fb & GRAD_PX) ? qxp : 0.0, s.qcx,
                                                  (fb & GRAD_MX) ? qxm : 0.0, P.idx[0]);
            const double dqy = one_sided_gradient(fb & GRAD_PY, fb & GRAD_MY, (fb & GRAD_PY) ? qyp : 0.0, s.qcy,
                                                  (fb & GRAD_MY) ? qym : 0.0, P.idx[1]);
            const double dqz = one_sided_gradient(fb & GRAD_PZ, fb & GRAD_MZ, (fb & GRAD_PZ) ? qzp : 0.0, s.qcz,
                                                  (fb & GRAD_MZ) ? qzm : 0.0, P.idx[2]);
            const Coll cc = collision_coefficients(s, mf, mg, dqx, dqy, dqz, P);
            const double omega = fluid ? cc.omega : 0.0;
            const bool st = own_lane && jj >= 0 && jj < ky;
            // relax, store, and hand the new population to the cell it will be pulled by
            static_for<0, NQ>([&](auto qc_) {
                constexpr int Q = decltype(qc_)::value;
                constexpr int d = ez(Q) + 1;
                const double v = f[Q] + omega * (feq_q<Q>(cc) - f[Q]);
                if (st) stb(A.fout[Q], c, v);
                double t = v;
                if constexpr (ex(Q) == 1) t = __shfl_up_sync(FULL, v, 1);
                if constexpr (ex(Q) == -1) t = __shfl_down_sync(FULL, v, 1);
                double(&T)[3][2] = ey(Q) == -1 ? TA : ey(Q) == 0 ? TB : TC;
                T[d][0] += t;
                if constexpr (ex(Q) == 1) T[d][1] += t;
                if constexpr (ex(Q) == -1) T[d][1] -= t;
            });
            // a later row of this column: into L2 while this one is finished (one bulk prefetch per component
            // row segment, issued by lane 0 for the warp's 32 cells; the y offsets of the current row are close
            // enough at a wrapped edge)
            if (C.prefetch > 0 && jj + C.prefetch <= ky && lane == 0) {
                int jp = j + C.prefetch;
                if (jp >= L.ny) jp -= L.ny;
                const unsigned dp = (unsigned)(jp - j) * px8;
                static_for<0, NQ>([&](auto qc_) {
                    constexpr int Q = decltype(qc_)::value;
                    const unsigned off = (cyz[ey(Q) + 1][ez(Q) + 1] + dp - 8u) & ~15u;
                    prefetch_l2_bulk((const char*)A.fin[Q] + off, 288u);
                    prefetch_l2_bulk((const char*)A.gin[Q] + off, 288u);
                });
            }
            static_for<0, NQ>([&](auto qc_) {
                constexpr int Q = decltype(qc_)::value;
                constexpr int d = ez(Q) + 1;
                const double gq = sg[Q * 128];
                const double v = gq + omega * (geq_q<Q>(cc) - gq);
                if (st) stb(A.gout[Q], c, v);
                double t = v;
                if constexpr (ex(Q) == 1) t = __shfl_up_sync(FULL, v, 1);
                if constexpr (ex(Q) == -1) t = __shfl_down_sync(FULL, v, 1);
                double(&E)[3] = ey(Q) == -1 ? EA : ey(Q) == 0 ? EB : EC;
                E[d] += t;
            });
        }
        // merge with the ring.  Row j-1 of this chunk is complete: every plane sum goes out once.
        const bool out = own_lane && jj >= 1;
        const unsigned cd = (unsigned)(i + OX) * 8u + (unsigned)(y0 + jj - 1 + GY) * px8 + (unsigned)(k + GZ) * sz8;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double a_rho = ring[(4 * d + 0) * 128] + TA[d][0];
            const double a_jx = ring[(4 * d + 1) * 128] + TA[d][1];
            const double a_jy = ring[(4 * d + 2) * 128] - TA[d][0];  // e_y = -1 terms
            const double a_e2 = ring[(4 * d + 3) * 128] + EA[d];
            if (out) {
                stb(A.part[4 * d + 0], cd, a_rho);
                stb(A.part[4 * d + 1], cd, a_jx);
                stb(A.part[4 * d + 2], cd, a_jy);
                stb(A.part[4 * d + 3], cd, a_e2);
            }
            // rotate: B (+ this row's e_y = 0 terms) becomes A, this row's e_y = +1 terms become B
            const double b_rho = ring[(12 + 3 * d + 0) * 128];
            ring[(4 * d + 0) * 128] = b_rho + TB[d][0];
            ring[(4 * d + 1) * 128] = ring[(12 + 3 * d + 1) * 128] + TB[d][1];
            ring[(4 * d + 2) * 128] = b_rho;  // B held e_y = +1 terms only
            ring[(4 * d + 3) * 128] = ring[(12 + 3 * d + 2) * 128] + EB[d];
            ring[(12 + 3 * d + 0) * 128] = TC[d][0];
            ring[(12 + 3 * d + 1) * 128] = TC[d][1];
            ring[(12 + 3 * d + 2) * 128] = EC[d];
        }
    }
}

// ---------------------------------------------------------------------------
// both passes in ONE persistent launch (plain loads): a global ticket counter hands out 128-cell row
// jobs, ticket 2n = q-correction job n, ticket 2n+1 = collide job n - LAG, slab-major order (fused.cu
// explains the order and the completion counters).  The collide job re-reads from L2 what the
// q-correction job of the same row pulled from HBM a few hundred tickets earlier.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_i32(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

constexpr int FP_UW = 32;  // a job is one warp wide: every warp is an autonomous worker, no CTA barrier anywhere
template <bool MACRO>
__global__ void __launch_bounds__(128, 3)
    k_fused_plain(const double* __restrict__ fin, const double* __restrict__ gin, double* __restrict__ fout,
                  double* __restrict__ gout, const uint32_t* __restrict__ nbr, const uint8_t* __restrict__ flag,
                  double* __restrict__ qc, double* __restrict__ macro, const __grid_constant__ Layout L,
                  const __grid_constant__ Phys P, const __grid_constant__ FusedPlan F, int* __restrict__ counters)
{
    int* tickets = counters;
    int* done = counters + 1;
    const int lane = threadIdx.x & 31;
    const uint64_t pol_keep = make_policy(HINT_KEEP), pol_last = make_policy(HINT_LAST);
    int next = 0, pending = -1;
    if (lane == 0) next = atomicAdd(tickets, 1);
    // A finished q-correction job is published (release-increment of its slab counter) one job later, when
    // its stores have long drained and the fence is cheap; a warp that is about to wait publishes first.
    auto publish = [&]() {
        if (pending >= 0) {
            if (lane == 0) {
                __threadfence();
                atomicAdd(done + pending, 1);
            }
            pending = -1;
        }
    };
    auto target = [&](int bb) { return min(F.B + 1, L.ny - bb * F.B) * F.UPR; };
    for (;;) {
        const long long ticket = __shfl_sync(0xffffffffu, next, 0);
        if (ticket >= F.total_tickets) break;
        if (lane == 0) next = atomicAdd(tickets, 1);  // used one job later: its latency is hidden
        // decode (same enumeration as fused.cu)
        int type;
        long long jn;
        if (F.mode == 2) {
            type = (int)(ticket & 1);
            jn = (ticket >> 1) - (type == 1 ? F.LAG : 0);
        } else {
            type = F.mode;
            jn = ticket;
        }
        if (jn < 0 || jn >= F.NJ) continue;
        const int slab = (int)(jn / F.JPS), r = (int)(jn % F.JPS);
        const int b = slab / F.NK, kk = slab % F.NK, k = F.kq0 + kk;
        const int row = r / F.UPR, i = (r % F.UPR) * FP_UW + lane, j = b * F.B + row;
        if (j >= L.ny) continue;
        if (type == 1 && (row >= F.B || k < 0 || k >= L.nz)) continue;
        if (type == 0) {
            if (i < L.nx) qcorr_cell<true, HINT_KEEP>(fin, gin, nbr, qc, L, P, i, j, k, pol_keep);
            publish();     // the previous q-correction job of this warp
            __syncwarp();  // orders every lane's QCorr stores before lane 0's later fence + increment
            pending = slab;
        } else {
            if (F.mode == 2) {
                if (lane == 0) {
                    auto ok = [&]() {
                        bool o = ld_acquire_i32(done + slab) >= target(b);
                        if (o && kk > 0) o = ld_acquire_i32(done + slab - 1) >= target(b);
                        if (o && kk < F.NK - 1) o = ld_acquire_i32(done + slab + 1) >= target(b);
                        if (o && b > 0) o = ld_acquire_i32(done + slab - F.NK) >= target(b - 1);
                        return o;
                    };
                    if (!ok()) {
                        if (pending >= 0) {
                            __threadfence();
                            atomicAdd(done + pending, 1);
                        }
                        const long long t0 = clock64();
                        while (!ok()) {
                            __nanosleep(100);
                            if (clock64() - t0 > 4000000000LL) {
                                printf("marbles_b200: k_fused_plain dependency wait timed out (block %d slab %d)\n", blockIdx.x, slab);
                                __trap();
                            }
                        }
                        pending = -2;  // published above
                    }
                }
                if (__shfl_sync(0xffffffffu, pending, 0) == -2) pending = -1;
            }
            if (i < L.nx) collide_cell<true, MACRO, HINT_LAST>(fin, gin, fout, gout, nbr, flag, qc, macro, L, P, i, j, k, pol_last);
            publish();
        }
    }
    publish();
}

int launch_collide_carry(const Layout& L, const Phys& P, const CarryPlan& C, int min_blocks, const double* fin,
                         const double* gin, double* fout, double* gout, const uint32_t* nbr, const uint8_t* flag,
                         const double* qc, double* part, cudaStream_t st)
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
    for (int e = 0; e < CARRY_EDGE_WORDS; ++e) A.edge[e] = nullptr;
    const dim3 grid((C.nxc + 3) / 4, (L.ny + C.ky - 1) / C.ky, L.nz);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_collide_carry<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CARRY_SMEM_BYTES);
        cudaFuncSetAttribute(k_collide_carry<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, CARRY_SMEM_BYTES);
        cudaFuncSetAttribute(k_collide_carry<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, CARRY_SMEM_BYTES);
        attr_done = true;
    }
    if (min_blocks >= 4)
        k_collide_carry<4><<<grid, 128, CARRY_SMEM_BYTES, st>>>(A, nbr, flag, L, P, C);
    else if (min_blocks == 3)
        k_collide_carry<3><<<grid, 128, CARRY_SMEM_BYTES, st>>>(A, nbr, flag, L, P, C);
    else
        k_collide_carry<2><<<grid, 128, CARRY_SMEM_BYTES, st>>>(A, nbr, flag, L, P, C);
    return 1;
}

int launch_fused_plain(const Layout& L, const Phys& P, int band_rows, int lag_quarters, int mode, int sm_count,
                       const double* fin, const double* gin, double* fout, double* gout, const uint32_t* nbr,
                       const uint8_t* flag, double* qc, double* macro, int* counters, cudaStream_t st)
{
    const int grid = 3 * sm_count, workers = grid * 4;
    // collide job n follows q-correction job n + LAG: two slabs (the z+1 neighbour's q-correction must be
    // complete) plus lag_quarters/4 of the jobs the grid has in flight
    FusedPlan F = make_fused_plan(L, FP_UW, band_rows, mode, 0, 0);
    F.LAG = 2LL * F.JPS + (long long)lag_quarters * workers / 4;
    F.total_tickets = mode == 2 ? 2 * (F.NJ + F.LAG) : F.NJ;
    const size_t ints = mode == 1 ? 1 : 1 + (size_t)F.NB * F.NK;
    cudaMemsetAsync(counters, 0, ints * sizeof(int), st);
    if (macro)
        k_fused_plain<true><<<grid, 128, 0, st>>>(fin, gin, fout, gout, nbr, flag, qc, macro, L, P, F, counters);
    else
        k_fused_plain<false><<<grid, 128, 0, st>>>(fin, gin, fout, gout, nbr, flag, qc, macro, L, P, F, counters);
    return 1;
}

