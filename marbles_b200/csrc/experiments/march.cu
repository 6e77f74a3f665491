// march.cu -- (round-2 experiment, variant 8; built only with MBL_EXPERIMENTS=1) ONE kernel per step: pull-stream + q-correction + collide with every population touched once.
//
// The collision of a cell needs the gradient of QCorr, a function of the POST-STREAM moments of its six face
// neighbours (LBM.cpp:893-901, 959-991).  The two-kernel step reads every population twice for that, the carry
// step (kernels.cu) sends partial sums through DRAM.  Here a CTA owns a column strip of WP rows x 30 cells and
// MARCHES through the planes of a z-chunk.  Per plane every thread pulls the 54 populations of its cell ONCE,
// by cp.async into a private shared-memory column, where they stay for one iteration:
//
//   iteration k:   C(k)    collide plane k   (populations pulled two iterations ago; needs QCorr of k-1, k, k+1)
//                  A(k+2)  pull plane k+2 into the slots C(k) has just freed        (overlaps C(k)'s arithmetic)
//                  M(k+2)  rho, j, 2rhoE -> QCorr of plane k+2
//
// QCorr of the face neighbours: z from the thread's own registers (planes k-1, k+1), x from the neighbouring
// lanes (shuffle; a warp is 30 owned cells + one halo lane on either side), y from the neighbouring rows through
// a small shared-memory ring; the rows just below and above the CTA's WP rows are pulled by two extra warps that
// only compute moments (no collide, no store).  Those halo pulls are re-reads of lines the neighbouring CTA
// pulls at the same time (L2 hits); DRAM sees 54 reads + 54 writes per cell, plus two priming planes per chunk.
// Nothing else is stored: no QCorr array, no partial sums, no second kernel.
//
// MEASURED (profiles/r02/march_variants_512.json, traffic_512_march.csv): parity-green, DRAM traffic 147 GB per
// 512^3 step instead of 161 GB, but 49.7 ms per step against 28.8 ms for the tile carry step: two planes of
// populations per cell in shared memory leave room for 8 warps per SM (202 KB per CTA), the collide arithmetic of
// 1.5 productive warps per scheduler is latency-bound (issue slots 19 % busy, DRAM 33 %), and the halo re-reads miss
// L2 (CTAs of one wave drift apart by more planes than L2 holds).  Kept as the record of why the populations
// cannot simply stay on chip between the two touches.
//
// Every cell's QCorr comes from the exact pull (bounce-back through the 27-bit mask, ghost cells filled by the
// ghost kernels on non-periodic levels), so walls, EB cells and slab edges take the same path as the interior.
#include "../kernels.cuh"

namespace mbl {

namespace {

__device__ __forceinline__ void cp_async8(unsigned smem_addr, const void* gptr)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct QC3 {
    double x, y, z;
};

}  // namespace

template <int WP, int PIPE>
__global__ void __launch_bounds__(32 * (WP + 2), 1)
    k_march(const __grid_constant__ MarchPtrs A, const uint32_t* __restrict__ nbr, const uint8_t* __restrict__ flag,
            const __grid_constant__ Layout L, const __grid_constant__ Phys P, const __grid_constant__ MarchPlan C)
{
    constexpr int T = 32 * (WP + 2);   // threads: WP productive rows + the row below + the row above
    constexpr int TP = 32 * WP;        // threads of the productive rows (they keep a second buffer)
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double smem[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int j0 = blockIdx.y * WP;
    const int rows = min(WP, L.ny - j0);
    const bool prod_warp = w >= 1 && w <= rows;
    const bool halo_warp = w == 0 || w == rows + 1;
    // two buffers of 54 slots (slot-major: a warp's access to one slot is one conflict-free 256-byte row);
    // the halo rows never keep a plane, their two buffers are the same one
    double* bufX = smem + tid;
    double* bufY = prod_warp ? smem + 2 * NQ * T + (tid - 32) : bufX;
    constexpr int SX = T, SY_PROD = TP;
    int strideX = SX, strideY = prod_warp ? SY_PROD : SX;
    double* ring = smem + 2 * NQ * T + 2 * NQ * TP;  // ring[slot(4)][T]: QCorr_y of the thread's cell

    // the cell column of this thread: its own, the periodic image over a wrapped edge, or none
    const int i = blockIdx.x * C.own - C.halo + lane;
    int is = i;
    bool xact = true;
    if (is < 0) {
        xact = L.wrap[0] && is + L.nx >= 0;
        is += L.nx;
    } else if (is >= L.nx) {
        xact = L.wrap[0] && is - L.nx < L.nx;
        is -= L.nx;
    }
    int j = j0 - 1 + w;
    bool yact = prod_warp || halo_warp;
    if (j < 0) {
        yact = yact && L.wrap[1];
        j += L.ny;
    } else if (j >= L.ny) {
        yact = yact && L.wrap[1] && j - L.ny < L.ny;
        j -= L.ny;
    }
    const bool act = xact && yact;
    const bool own = prod_warp && lane >= C.halo && lane < C.halo + C.own && i < L.nx;
    if (!act) is = 0, j = 0;  // keep addresses inside the box; nothing is loaded or stored

    const unsigned px8 = (unsigned)L.px * 8u, sz8 = (unsigned)L.sz * 8u;
    unsigned xo[3], yo[3];
    xo[1] = yo[1] = 0u;
    xo[2] = (L.wrap[0] && is == 0) ? (unsigned)(L.nx - 1) * 8u : 0u - 8u;
    xo[0] = (L.wrap[0] && is == L.nx - 1) ? 0u - (unsigned)(L.nx - 1) * 8u : 8u;
    yo[2] = (L.wrap[1] && j == 0) ? (unsigned)(L.ny - 1) * px8 : 0u - px8;
    yo[0] = (L.wrap[1] && j == L.ny - 1) ? 0u - (unsigned)(L.ny - 1) * px8 : px8;
    const unsigned cxy = (unsigned)(is + OX) * 8u + (unsigned)(j + GY) * px8;

    // plane index -> plane inside the padded box (periodic image when the kernel wraps z itself)
    auto plane_of = [&](int kp) {
        if (L.wrap[2]) {
            if (kp < 0) kp += L.nz;
            if (kp >= L.nz) kp -= L.nz;
        }
        return kp;
    };
    auto cell_of = [&](int kp) { return cxy + (unsigned)(plane_of(kp) + GZ) * sz8; };
    auto mask_of = [&](int kp) -> uint32_t {
        return act ? *(const uint32_t*)((const char*)nbr + (cell_of(kp) >> 1)) : 0u;
    };

    // A(kp): pull the populations of plane kp into buf.  part 0: f, part 1: g, part 2: both
    auto issue = [&](int kp, uint32_t m, double* buf, int stride, int part) {
        if (!(m & 1u)) return;  // inactive thread or solid cell: nothing to pull
        const int kw = plane_of(kp);
        const unsigned c = cxy + (unsigned)(kw + GZ) * sz8;
        unsigned zo[3];
        zo[1] = 0u;
        zo[2] = (L.wrap[2] && kw == 0) ? (unsigned)(L.nz - 1) * sz8 : 0u - sz8;
        zo[0] = (L.wrap[2] && kw == L.nz - 1) ? 0u - (unsigned)(L.nz - 1) * sz8 : sz8;
        unsigned cyz[3][3];
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int d = 0; d < 3; ++d) cyz[b][d] = c + yo[b] + zo[d];
        const unsigned sa = (unsigned)__cvta_generic_to_shared(buf);
        const unsigned st8 = (unsigned)stride * 8u;
        if (m == ALL_FLUID) {
            static_for<0, NQ>([&](auto qc_) {
                constexpr int Q = decltype(qc_)::value;
                const unsigned off = cyz[ey(Q) + 1][ez(Q) + 1] + xo[ex(Q) + 1];
                if (part != 1) cp_async8(sa + Q * st8, (const char*)A.fin[Q] + off);
                if (part != 0) cp_async8(sa + (NQ + Q) * st8, (const char*)A.gin[Q] + off);
            });
        } else {
            // halfway bounce-back: a solid source gives the cell's own opposite population (LBM.cpp:590-595, pull form)
            static_for<0, NQ>([&](auto qc_) {
                constexpr int Q = decltype(qc_)::value;
                const bool fl = (m >> Q) & 1u;
                const unsigned off = fl ? cyz[ey(Q) + 1][ez(Q) + 1] + xo[ex(Q) + 1] : c;
                if (part != 1) cp_async8(sa + Q * st8, (const char*)(fl ? A.fin[Q] : A.fin[opp(Q)]) + off);
                if (part != 0) cp_async8(sa + (NQ + Q) * st8, (const char*)(fl ? A.gin[Q] : A.gin[opp(Q)]) + off);
            });
        }
    };

    // M: QCorr of the post-stream state in buf (LBM.cpp:841-901); zero where nothing was pulled
    auto moments_qc = [&](uint32_t m, const double* buf, int stride) {
        QC3 r = {0.0, 0.0, 0.0};
        if (m & 1u) {
            const MomG ml = moments_g([&](int q) { return buf[q * stride]; });  // {sum, x, y, z} = rho, jx, jy, jz
            double a[3];
#pragma unroll
            for (int z = 0; z < 3; ++z) {
                double s = 0.0;
#pragma unroll
                for (int t = 0; t < 9; ++t) s += buf[(NQ + 9 * z + t) * stride];
                a[z] = s;
            }
            const double e2 = (a[0] + a[1]) + a[2];
            const Prim s = primitives(ml.e2, ml.qx, ml.qy, ml.qz, e2, P);
            r.x = s.qcx, r.y = s.qcy, r.z = s.qcz;
        }
        return r;
    };

    const int kb0 = C.ka + (int)blockIdx.z * C.zm;
    const int kb1 = min(kb0 + C.zm, C.kb);

    // prologue: QCorr of planes kb0-1, kb0, kb0+1; plane kb0 stays in X, plane kb0+1 in Y
    uint32_t m0, m1, m2;
    QC3 qm, q0, q1;
    {
        const uint32_t mm = mask_of(kb0 - 1);
        issue(kb0 - 1, mm, bufX, strideX, 2);
        cp_async_commit();
        m0 = mask_of(kb0);
        m1 = mask_of(kb0 + 1);
        cp_async_wait0();
        qm = moments_qc(mm, bufX, strideX);
        issue(kb0, m0, bufX, strideX, 2);
        cp_async_commit();
        cp_async_wait0();
        q0 = moments_qc(m0, bufX, strideX);
        if (halo_warp || prod_warp) {
            issue(kb0 + 1, m1, bufY, strideY, 2);
            cp_async_commit();
            cp_async_wait0();
            q1 = moments_qc(m1, bufY, strideY);
        } else {
            q1 = {0.0, 0.0, 0.0};
        }
        ring[(kb0 & 3) * T + tid] = q0.y;
        ring[((kb0 + 1) & 3) * T + tid] = q1.y;
    }
    m2 = kb0 + 2 <= kb1 ? mask_of(kb0 + 2) : 0u;

    double* cur = bufX;
    double* nxt = bufY;
    int scur = strideX, snxt = strideY;

    auto ldb = [](const double* base, unsigned off) { return *(const double*)((const char*)base + off); };
    auto stb = [](double* base, unsigned off, double v) { *(double*)((char*)base + off) = v; };
    (void)ldb;

#pragma unroll 1
    for (int k = kb0; k < kb1; ++k) {
        __syncthreads();  // QCorr_y of planes k and k+1 of every row is in the ring
        const bool more = k + 2 <= kb1;  // plane k+2 is needed (its QCorr, and its populations if k+2 < kb1)
        const uint32_t m3 = (k + 3 <= kb1) ? mask_of(k + 3) : 0u;  // consumed by the next iteration's pull
        QC3 q2 = {0.0, 0.0, 0.0};
        if (prod_warp) {
            const unsigned c = cell_of(k);
            const unsigned fb = act ? flag[c >> 3] : 0u;
            // the six face neighbours' QCorr (LBM.cpp:979-986): x by shuffle, y from the ring, z from registers
            const double qxp = __shfl_down_sync(FULL, q0.x, 1), qxm = __shfl_up_sync(FULL, q0.x, 1);
            const double qyp = ring[(k & 3) * T + tid + 32], qym = ring[(k & 3) * T + tid - 32];
            const double qzp = q1.z, qzm = qm.z;
            double f[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) f[q] = cur[q * scur];
            if (PIPE && more) {  // the f slots are free: start pulling plane k+2 behind the arithmetic
                issue(k + 2, m2, cur, scur, 0);
                cp_async_commit();
            }
            double* const sg = cur + NQ * scur;
            const bool fluid = m0 & 1u;
            if (!fluid) {
                // solid cell: the streamed value is the -1 sentinel (LBM.cpp:565, 582) and collide skips it;
                // with omega = 0 below the "relaxed" value is exactly -1 again
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    f[q] = -1.0;
                    sg[q * scur] = -1.0;
                }
            }
            const MomF mf = moments_f([&](int q) { return f[q]; });
            const MomG mg = moments_g([&](int q) { return sg[q * scur]; });
            const Prim s = primitives(mf.rho, mf.jx, mf.jy, mf.jz, mg.e2, P);
            const double dqx = one_sided_gradient(fb & GRAD_PX, fb & GRAD_MX, (fb & GRAD_PX) ? qxp : 0.0, s.qcx,
                                                  (fb & GRAD_MX) ? qxm : 0.0, P.idx[0]);
            const double dqy = one_sided_gradient(fb & GRAD_PY, fb & GRAD_MY, (fb & GRAD_PY) ? qyp : 0.0, s.qcy,
                                                  (fb & GRAD_MY) ? qym : 0.0, P.idx[1]);
            const double dqz = one_sided_gradient(fb & GRAD_PZ, fb & GRAD_MZ, (fb & GRAD_PZ) ? qzp : 0.0, s.qcz,
                                                  (fb & GRAD_MZ) ? qzm : 0.0, P.idx[2]);
            const Coll cc = collision_coefficients(s, mf, mg, dqx, dqy, dqz, P);
            const double omega = fluid ? cc.omega : 0.0;
            // relax_f_to_equilibrium (LBM.cpp:799-801): g first, its slots are then free for plane k+2
            static_for<0, NQ>([&](auto qc_) {
                constexpr int Q = decltype(qc_)::value;
                const double gq = sg[Q * scur];
                if (own) stb(A.gout[Q], c, gq + omega * (geq_q<Q>(cc) - gq));
                asm volatile("" ::: "memory");  // keep the relax / store pairs in order: fewer values live at once
            });
            if (more) {
                issue(k + 2, m2, cur, scur, PIPE ? 1 : 2);
                cp_async_commit();
            }
            static_for<0, NQ>([&](auto qc_) {
                constexpr int Q = decltype(qc_)::value;
                if (own) stb(A.fout[Q], c, f[Q] + omega * (feq_q<Q>(cc) - f[Q]));
                asm volatile("" ::: "memory");
            });
            if (more) {
                cp_async_wait0();
                q2 = moments_qc(m2, cur, scur);
            }
        } else if (halo_warp) {
            if (more) {
                issue(k + 2, m2, cur, scur, 2);
                cp_async_commit();
                cp_async_wait0();
                q2 = moments_qc(m2, cur, scur);
            }
        }
        if (more) ring[((k + 2) & 3) * T + tid] = q2.y;
        qm = q0, q0 = q1, q1 = q2;
        m0 = m1, m1 = m2, m2 = m3;
        double* t = cur;
        cur = nxt, nxt = t;
        const int ts = scur;
        scur = snxt, snxt = ts;
    }
}

MarchPlan make_march_plan(const Layout& L, int zm, int ka, int kb)
{
    MarchPlan C;
    C.own = 30;
    C.halo = 1;
    C.nxc = (L.nx + C.own - 1) / C.own;
    if (kb <= ka) ka = 0, kb = L.nz;
    C.ka = ka, C.kb = kb;
    C.zm = zm < 1 ? 64 : zm;
    if (C.zm > kb - ka) C.zm = kb - ka;
    return C;
}

size_t march_smem_bytes(int wp)
{
    const size_t T = 32 * (wp + 2), TP = 32 * wp;
    return (2 * NQ * T + 2 * NQ * TP + 4 * T) * sizeof(double);
}

template <int WP, int PIPE>
static int launch_march_t(const MarchPtrs& A, const uint32_t* nbr, const uint8_t* flag, const Layout& L, const Phys& P,
                          const MarchPlan& C, cudaStream_t st)
{
    static bool attr_done = false;
    const size_t sm = march_smem_bytes(WP);
    if (!attr_done) {
        if (cudaFuncSetAttribute(k_march<WP, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
            return -3;
        attr_done = true;
    }
    const dim3 grid(C.nxc, (L.ny + WP - 1) / WP, (C.kb - C.ka + C.zm - 1) / C.zm);
    k_march<WP, PIPE><<<grid, 32 * (WP + 2), sm, st>>>(A, nbr, flag, L, P, C);
    return 1;
}

int launch_march(const Layout& L, const Phys& P, int rows, int zm, int pipe, const double* fin, const double* gin,
                 double* fout, double* gout, const uint32_t* nbr, const uint8_t* flag, cudaStream_t st, int ka, int kb)
{
    if (L.sq * 8 >= (1LL << 32)) return -1;  // 32-bit byte offsets inside a component
    MarchPtrs A;
    for (int q = 0; q < NQ; ++q) {
        A.fin[q] = fin + (long long)q * L.sq;
        A.gin[q] = gin + (long long)q * L.sq;
        A.fout[q] = fout + (long long)q * L.sq;
        A.gout[q] = gout + (long long)q * L.sq;
    }
    const MarchPlan C = make_march_plan(L, zm, ka, kb);
    if (C.kb <= C.ka) return 0;
    if (rows == 4) return launch_march_t<4, 1>(A, nbr, flag, L, P, C, st);  // small CTAs: the ragged-box tests
    return pipe ? launch_march_t<6, 1>(A, nbr, flag, L, P, C, st) : launch_march_t<6, 0>(A, nbr, flag, L, P, C, st);
}

}  // namespace mbl
