// fused.cu -- the whole lattice update of one level in ONE persistent kernel (sm_100a).
//
// Why: collide needs the gradient of the q-correction, which is built from the POST-stream moments of
// the six face neighbours (LBM.cpp:959-991), so every population has to be touched twice: once to form
// moments (job type P1: pull f,g -> rho,u,T -> QCorr, LBM.cpp:810-906) and once to relax (job type P2:
// pull, moments, grad QCorr, feq/geq, BGK relax, store; LBM.cpp:558-604 + 607-807).  As two kernels the
// second touch comes from HBM again (165 words per cell instead of 108).  Here both job types run in one
// grid, ordered so that P2 of a row follows P1 of the same row by a few hundred jobs: the second touch is
// an L2 hit and HBM sees each population once.
//
// Structure (one CTA = UW consumer threads + one producer warp, persistent, 1-2 CTAs per SM):
//   * jobs = (type, 128- or 256-cell row segment); a global ticket counter hands them out in an order
//     in which everything a job depends on has a smaller ticket (so waiting can never deadlock):
//     the level is cut into y-bands of B rows; inside a band jobs run plane by plane; ticket 2n is P1
//     job n, ticket 2n+1 is P2 job n-LAG.
//   * the producer warp stages a job's inputs in shared memory with 1-D bulk TMA copies
//     (cp.async.bulk ... mbarrier::complete_tx): for each of the 27 directions the pulled row segment
//     x - e_q is ONE contiguous run of the SoA component plane, so a pull is 27 row copies per lattice;
//     plus the pull-mask / flag bytes / five QCorr rows.  Ring of 3 population slots + 2 aux slots,
//     full/empty mbarriers; loads of the next job are in flight while the consumers compute.
//   * consumers: thread t owns cell i0+t.  P1: sum the staged rows -> QCorr -> global, then
//     release-increment the slab's completion counter.  P2: f to registers (slot released at once),
//     g read in place, collide, store.
//   * the producer checks the completion counters of the slabs a P2 job reads QCorr from (acquire) before
//     it issues that job's QCorr copies.
// Bounce-back (source cell solid) falls back to a direct global load of the cell's own opposite
// population.  The staged rows are read un-wrapped, so this kernel needs every ghost cell (periodic images
// included) filled by the ghost kernels (kernels.cu) beforehand.
#include "../kernels.cuh"

#include <cstdio>

namespace mbl {

namespace {

constexpr int JOB_P1 = 0, JOB_P2 = 1, JOB_EXIT = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Every wait in this kernel is bounded: a protocol error must surface as a launch failure
// (cudaErrorLaunchFailure from __trap), never as a hung GPU.
constexpr long long SPIN_LIMIT_CYCLES = 4000000000LL;  // ~2 s
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > SPIN_LIMIT_CYCLES) {
            printf("marbles_b200: k_fused barrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_normal()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ int ld_acquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int UW>
struct Cfg {
    static constexpr int ROWD = UW + 4;  // staged cells per row: i0-2 .. i0+UW+1
    static constexpr int SLOT_BYTES = NQ * ROWD * 8;
    static constexpr int NSLOT = 3;
    static constexpr int NAUX = 2;
    static constexpr int DESC_BYTES = 64;
    static constexpr int NBR_BYTES = ROWD * 4;
    static constexpr int FLAG_BYTES = UW + 16;
    static constexpr int QC_BYTES = ROWD * 8;
    static constexpr int AUX_BYTES = DESC_BYTES + NBR_BYTES + FLAG_BYTES + 5 * QC_BYTES;
    static constexpr int BAR_BYTES = 128;
    static constexpr int SMEM_BYTES = NSLOT * SLOT_BYTES + NAUX * AUX_BYTES + BAR_BYTES;
    static constexpr int NCW = UW / 32;  // consumer warps
    static constexpr int THREADS = UW;
    static_assert(SLOT_BYTES % 16 == 0 && AUX_BYTES % 16 == 0 && NBR_BYTES % 16 == 0 && FLAG_BYTES % 16 == 0, "TMA alignment");
};

struct JobDesc {
    int type, i0, j, k, slab;
    int pad[11];
};
static_assert(sizeof(JobDesc) == 64, "descriptor slot");

struct Job {
    int type, i0, j, k, slab, b, kk;
    bool valid;
};

// ticket -> job.  Both job types are enumerated slab-major: n = slab * JPS + row * UPR + seg with
// slab = band * NK + plane; P1 covers B+1 rows per band (the extra row feeds the y-gradient of the band's
// last row), P2 covers B rows and only valid planes.
__device__ __forceinline__ Job decode(long long ticket, const FusedPlan& F, const Layout& L, int UW)
{
    Job J;
    J.valid = false;
    long long n;
    if (F.mode == 2) {
        J.type = (int)(ticket & 1);
        n = (ticket >> 1) - (J.type == JOB_P2 ? F.LAG : 0);
    } else {
        J.type = F.mode;
        n = ticket;
    }
    if (n < 0 || n >= F.NJ) {
        J.type = (ticket >= F.total_tickets) ? JOB_EXIT : J.type;
        return J;
    }
    J.slab = (int)(n / F.JPS);
    const int r = (int)(n % F.JPS);
    J.b = J.slab / F.NK;
    J.kk = J.slab % F.NK;
    J.k = F.kq0 + J.kk;
    const int row = r / F.UPR;
    J.i0 = (r % F.UPR) * UW;
    J.j = J.b * F.B + row;
    if (J.j >= L.ny) return J;
    if (J.type == JOB_P2 && (row >= F.B || J.k < 0 || J.k >= L.nz)) return J;
    J.valid = true;
    return J;
}

__device__ __forceinline__ int slab_target(int b, const FusedPlan& F, const Layout& L, int ncw)
{
    const int rows = min(F.B + 1, L.ny - b * F.B);
    return rows * F.UPR * ncw;
}

}  // namespace

// One thread (thread 0) also plays producer at two fixed points of every job: (A) when the f slot of the
// current job has been drained it refills that slot with g of the next job, (B) at the end of the job it
// draws the job after next (the ticket was requested one job earlier, so the atomic's latency is hidden),
// checks its dependencies and issues its aux rows and f rows.  Loads therefore run one job ahead of the
// arithmetic; a separate producer warp would leave the 4 SM sub-partitions unevenly loaded (5 warps per
// CTA) and caps the kernel at 168 registers, which spills.
template <int UW, bool MACRO>
__global__ void __launch_bounds__(Cfg<UW>::THREADS, UW == 128 ? 2 : 1)
    k_fused(const double* __restrict__ fin, const double* __restrict__ gin, double* __restrict__ fout,
            double* __restrict__ gout, const uint32_t* __restrict__ nbr, const uint8_t* __restrict__ flag,
            double* __restrict__ qc, double* __restrict__ macro, Layout L, Phys P, FusedPlan F, int* __restrict__ counters)
{
    using C = Cfg<UW>;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* pop_base = smem;
    unsigned char* aux_base = smem + C::NSLOT * C::SLOT_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(aux_base + C::NAUX * C::AUX_BYTES);
    uint64_t* pop_full = bars;                 // [NSLOT]
    uint64_t* pop_empty = bars + C::NSLOT;     // [NSLOT]
    uint64_t* aux_full = bars + 2 * C::NSLOT;  // [NAUX]
    uint64_t* aux_empty = aux_full + C::NAUX;  // [NAUX]

    // element offset of the row a direction pulls from, relative to the job's own row (the x shift is
    // applied when the staged row is read)
    __shared__ long long s_rowoff[NQ];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    if (tid == 0) {
        static_for<0, NQ>([&](auto qc_) {
            constexpr int Q = decltype(qc_)::value;
            s_rowoff[Q] = (long long)Q * L.sq - ((long long)ey(Q) * L.px + (long long)ez(Q) * L.sz);
        });
        for (int s = 0; s < C::NSLOT; ++s) {
            mbar_init(pop_full + s, 1);
            mbar_init(pop_empty + s, C::NCW);
        }
        for (int a = 0; a < C::NAUX; ++a) {
            mbar_init(aux_full + a, 1);
            mbar_init(aux_empty + a, C::NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    int* tickets = counters;
    int* done = counters + 1;
    const long long n = L.sq;

    // A finished q-correction job is published (release-increment of its slab counter) lazily: right after
    // its stores the fence would wait for them to drain (~1 us per job); one job later they are long
    // complete.  Anything that might block on another job publishes first, so laziness cannot deadlock.
    int pending_slab = -1;
    auto publish_pending = [&]() {
        if (pending_slab >= 0) {
            if (lane == 0) {
                __threadfence();
                atomicAdd(done + pending_slab, 1);
            }
            pending_slab = -1;
        }
    };
    auto wait_full = [&](uint64_t* bar, uint32_t parity) {
        if (!mbar_try_wait(bar, parity)) {
            publish_pending();
            mbar_wait(bar, parity);
        }
    };

    // ------------------------------------------------------------------ producer state (thread 0 only)
    uint32_t ps = 0, pphase = 0, pas = 0, paphase = 0;  // next population / aux slot to fill
    int prefetched_ticket = 0;
    bool producing = (tid == 0);  // false once the EXIT descriptor has been queued
    Job pend;                     // job whose g rows are still to be issued
    pend.valid = false;
    uint64_t pol_keep = 0, pol_last = 0, pol_norm = 0;
    if (tid == 0) {
        // P1 reads are re-read by the P2 job a few hundred tickets later: keep them; P2 reads are the last use
        pol_keep = policy_evict_last(), pol_last = policy_evict_first(), pol_norm = policy_normal();
        prefetched_ticket = atomicAdd(tickets, 1);
    }
    auto issue_pop = [&](const Job& J, int lat) {
        const long long c0 = L.cell(J.i0 - 2, J.j, J.k);
        const uint32_t len_d = (uint32_t)min(C::ROWD, (int)(L.px - J.i0)) * 8u;
        const uint64_t pol = (F.mode == 2) ? (J.type == JOB_P1 ? pol_keep : pol_last) : pol_norm;
        const double* src = lat ? gin : fin;
        unsigned char* slot = pop_base + ps * C::SLOT_BYTES;
        mbar_wait(pop_empty + ps, pphase ^ 1);
        mbar_expect_tx(pop_full + ps, NQ * len_d);
#pragma unroll 9
        for (int q = 0; q < NQ; ++q) bulk_g2s(slot + q * C::ROWD * 8, src + c0 + s_rowoff[q], len_d, pop_full + ps, pol);
        if (++ps == C::NSLOT) ps = 0, pphase ^= 1;
    };
    auto produce_g = [&]() {
        if (pend.valid) {
            issue_pop(pend, 1);
            pend.valid = false;
        }
    };
    auto deps_ok = [&](const Job& J) {
        const int tgt = slab_target(J.b, F, L, C::NCW);
        bool ok = ld_acquire(done + J.slab) >= tgt;
        if (ok && J.kk > 0) ok = ld_acquire(done + J.slab - 1) >= tgt;
        if (ok && J.kk < F.NK - 1) ok = ld_acquire(done + J.slab + 1) >= tgt;
        if (ok && J.b > 0) ok = ld_acquire(done + J.slab - F.NK) >= slab_target(J.b - 1, F, L, C::NCW);
        return ok;
    };
    // Draw tickets until a real job (or the end) and queue its descriptor + aux rows + f rows.  A collide job
    // whose q-corrections are not complete yet is only waited for when this CTA has nothing else queued
    // (may_block): otherwise it is parked and retried after the next job, because the jobs this CTA has
    // queued may be exactly what another CTA's parked job is waiting for.
    int queued = 0;
    bool has_parked = false;
    Job parked;
    parked.valid = false;
    auto try_produce = [&](bool may_block) -> bool {
        produce_g();
        Job J;
        if (has_parked) {
            J = parked;
        } else {
            for (;;) {
                const long long ticket = prefetched_ticket;
                prefetched_ticket = atomicAdd(tickets, 1);  // consumed one job later
                J = decode(ticket, F, L, UW);
                if (J.valid || J.type == JOB_EXIT) break;
            }
        }
        if (J.type == JOB_P2 && F.mode == 2) {
            // QCorr of this row's face neighbours must be complete (and visible to the async proxy)
            if (!deps_ok(J)) {
                if (!may_block) {
                    parked = J;
                    has_parked = true;
                    return false;
                }
                const long long t0 = clock64();
                do {
                    publish_pending();  // the awaited q-correction job may be this warp's own
                    __nanosleep(100);
                    if (clock64() - t0 > SPIN_LIMIT_CYCLES) {
                        printf("marbles_b200: k_fused dependency wait timed out (block %d slab %d)\n", blockIdx.x, J.slab);
                        __trap();
                    }
                } while (!deps_ok(J));
            }
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        has_parked = false;
        unsigned char* aux = aux_base + pas * C::AUX_BYTES;
        mbar_wait(aux_empty + pas, paphase ^ 1);
        JobDesc* d = reinterpret_cast<JobDesc*>(aux);
        d->type = J.type;
        if (J.type == JOB_EXIT) {
            mbar_arrive(aux_full + pas);
            producing = false;
            return false;
        }
        d->i0 = J.i0, d->j = J.j, d->k = J.k, d->slab = J.slab;
        const long long c0 = L.cell(J.i0 - 2, J.j, J.k);  // first staged cell of the row
        const int avail = (int)(L.px - J.i0);              // cells left in the padded row from c0
        const uint32_t len_d = (uint32_t)min(C::ROWD, avail) * 8u;
        const uint32_t len_n = (uint32_t)min(C::ROWD, avail) * 4u;
        const uint32_t len_f = (uint32_t)min(C::FLAG_BYTES, avail);
        unsigned char* a_nbr = aux + C::DESC_BYTES;
        unsigned char* a_flag = a_nbr + C::NBR_BYTES;
        unsigned char* a_qc = a_flag + C::FLAG_BYTES;
        mbar_expect_tx(aux_full + pas, J.type == JOB_P2 ? len_n + len_f + 5 * len_d : len_n);
        bulk_g2s(a_nbr, nbr + c0, len_n, aux_full + pas, pol_norm);
        if (J.type == JOB_P2) {
            bulk_g2s(a_flag, flag + c0, len_f, aux_full + pas, pol_norm);
            bulk_g2s(a_qc + 0 * C::QC_BYTES, qc + c0, len_d, aux_full + pas, pol_norm);
            bulk_g2s(a_qc + 1 * C::QC_BYTES, qc + n + c0 - L.px, len_d, aux_full + pas, pol_norm);
            bulk_g2s(a_qc + 2 * C::QC_BYTES, qc + n + c0 + L.px, len_d, aux_full + pas, pol_norm);
            bulk_g2s(a_qc + 3 * C::QC_BYTES, qc + 2 * n + c0 - L.sz, len_d, aux_full + pas, pol_norm);
            bulk_g2s(a_qc + 4 * C::QC_BYTES, qc + 2 * n + c0 + L.sz, len_d, aux_full + pas, pol_norm);
        }
        if (++pas == C::NAUX) pas = 0, paphase ^= 1;
        issue_pop(J, 0);
        pend = J;
        ++queued;
        return true;
    };
    // keep two jobs queued: the one being computed and the one whose rows are in flight
    auto top_up = [&]() {
        while (producing && queued < 2)
            if (!try_produce(queued == 0)) break;
    };
    if (tid == 0) top_up();

    // ---------------------------------------------------------------------- all warps consume
    uint32_t cs = 0, cphase = 0, as = 0, aphase = 0;
    auto next_slot = [&]() {
        if (++cs == C::NSLOT) cs = 0, cphase ^= 1;
    };
    auto release_slot = [&](uint64_t* bar) {
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
    };
    for (;;) {
        unsigned char* aux = aux_base + as * C::AUX_BYTES;
        wait_full(aux_full + as, aphase);
        const JobDesc* d = reinterpret_cast<const JobDesc*>(aux);
        const int type = d->type;
        if (type == JOB_EXIT) break;
        const int i0 = d->i0, j = d->j, k = d->k, slab = d->slab;
        const uint32_t* s_nbr = reinterpret_cast<const uint32_t*>(aux + C::DESC_BYTES);
        const uint8_t* s_flag = aux + C::DESC_BYTES + C::NBR_BYTES;
        const double* s_qc = reinterpret_cast<const double*>(aux + C::DESC_BYTES + C::NBR_BYTES + C::FLAG_BYTES);
        const int i = i0 + tid;
        const bool in_row = i < L.nx;
        const uint32_t m = in_row ? s_nbr[tid + 2] : 0u;
        const bool fluid = (m & 1u) != 0;
        const bool fast = __all_sync(0xffffffffu, !fluid || m == ALL_FLUID);
        const long long c = L.cell(i, j, k);
        // value of direction Q pulled into this thread's cell from a staged slot
        auto pulled = [&](const double* slot, const double* __restrict__ src, auto qc_) {
            constexpr int Q = decltype(qc_)::value;
            double v = slot[Q * C::ROWD + tid + 2 - ex(Q)];
            if (!fast && fluid && !((m >> Q) & 1u)) v = src[(long long)opp(Q) * n + c];  // halfway bounce-back
            return v;
        };

        if (type == JOB_P1) {
            release_slot(aux_empty + as);  // only the pull mask was needed
            MomL ml = {0.0, 0.0, 0.0, 0.0};
            double e2 = 0.0;
            {
                const double* slot = reinterpret_cast<const double*>(pop_base + cs * C::SLOT_BYTES);
                wait_full(pop_full + cs, cphase);
                if (fluid) static_for<0, NQ>([&](auto qc_) { acc_l<decltype(qc_)::value>(ml, pulled(slot, fin, qc_)); });
                release_slot(pop_empty + cs);
                next_slot();
                if (tid == 0) produce_g();
            }
            {
                const double* slot = reinterpret_cast<const double*>(pop_base + cs * C::SLOT_BYTES);
                wait_full(pop_full + cs, cphase);
                if (fluid) static_for<0, NQ>([&](auto qc_) { e2 += pulled(slot, gin, qc_); });
                release_slot(pop_empty + cs);
                next_slot();
            }
            publish_pending();
            if (fluid) {
                const Prim s = primitives(ml.rho, ml.jx, ml.jy, ml.jz, e2, P);
                qc[c] = s.qcx;
                qc[n + c] = s.qcy;
                qc[2 * n + c] = s.qcz;
            }
            __syncwarp();  // orders every lane's stores before lane 0's later fence + increment
            pending_slab = slab;
        } else {
            double f[NQ];
            MomF mf = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            MomG mg = {0, 0, 0, 0};
            {
                const double* slot = reinterpret_cast<const double*>(pop_base + cs * C::SLOT_BYTES);
                wait_full(pop_full + cs, cphase);
                static_for<0, NQ>([&](auto qc_) {
                    constexpr int Q = decltype(qc_)::value;
                    f[Q] = fluid ? pulled(slot, fin, qc_) : 0.0;
                    acc_f<Q>(mf, f[Q]);
                });
                release_slot(pop_empty + cs);  // f lives in registers from here on
                next_slot();
                if (tid == 0) produce_g();
            }
            const double* gslot = reinterpret_cast<const double*>(pop_base + cs * C::SLOT_BYTES);
            uint64_t* gbar = pop_empty + cs;
            wait_full(pop_full + cs, cphase);
            next_slot();
            if (fluid) static_for<0, NQ>([&](auto qc_) { acc_g<decltype(qc_)::value>(mg, pulled(gslot, gin, qc_)); });
            const Prim s = primitives(mf.rho, mf.jx, mf.jy, mf.jz, mg.e2, P);
            // grad of the q-correction (LBM.cpp:959-991, Utilities.H:279-312) from the staged rows
            const unsigned fb = s_flag[tid + 2];
            double dqx, dqy, dqz;
            {
                const bool okp = fb & GRAD_PX, okm = fb & GRAD_MX;
                const double dp = okp ? s_qc[tid + 3] : 0.0, dm = okm ? s_qc[tid + 1] : 0.0;
                dqx = one_sided_gradient(okp, okm, dp, s.qcx, dm, P.idx[0]);
            }
            {
                const bool okp = fb & GRAD_PY, okm = fb & GRAD_MY;
                const double dp = okp ? s_qc[2 * C::ROWD + tid + 2] : 0.0, dm = okm ? s_qc[1 * C::ROWD + tid + 2] : 0.0;
                dqy = one_sided_gradient(okp, okm, dp, s.qcy, dm, P.idx[1]);
            }
            {
                const bool okp = fb & GRAD_PZ, okm = fb & GRAD_MZ;
                const double dp = okp ? s_qc[4 * C::ROWD + tid + 2] : 0.0, dm = okm ? s_qc[3 * C::ROWD + tid + 2] : 0.0;
                dqz = one_sided_gradient(okp, okm, dp, s.qcz, dm, P.idx[2]);
            }
            release_slot(aux_empty + as);
            publish_pending();
            if (in_row && !fluid) {
                // solid cell: the streamed value is the -1 sentinel (LBM.cpp:565, 582); collide skips it
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    fout[q * n + c] = -1.0;
                    gout[q * n + c] = -1.0;
                }
            }
            if (fluid) {
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
                    macro[23 * n + c] = dqx;
                    macro[24 * n + c] = dqy;
                    macro[25 * n + c] = dqz;
                }
                const Coll cc = collision_coefficients(s, mf, mg, dqx, dqy, dqz, P);
                // relax_f_to_equilibrium (LBM.cpp:799-801) + the FillBoundary of f, g that follows it
                static_for<0, NQ>([&](auto qc_) {
                    constexpr int Q = decltype(qc_)::value;
                    fout[(long long)Q * n + c] = f[Q] + cc.omega * (feq_q<Q>(cc) - f[Q]);
                });
                static_for<0, NQ>([&](auto qc_) {
                    constexpr int Q = decltype(qc_)::value;
                    const double gq = pulled(gslot, gin, qc_);
                    gout[(long long)Q * n + c] = gq + cc.omega * (geq_q<Q>(cc) - gq);
                });
            }
            release_slot(gbar);
        }
        if (++as == C::NAUX) as = 0, aphase ^= 1;
        if (tid == 0) {
            --queued;
            top_up();
        }
    }
    publish_pending();
}

// ===========================================================================
// host side
// ===========================================================================
FusedPlan make_fused_plan(const Layout& L, int uw, int band_rows, int mode, int grid, int lag_per_cta)
{
    FusedPlan F;
    F.B = band_rows < L.ny ? band_rows : L.ny;
    F.NB = (L.ny + F.B - 1) / F.B;
    // q-corrections are needed on the valid planes and, where the box borders another rank in z, on the
    // first ghost plane (recomputed from the two exchanged planes instead of a second exchange)
    F.kq0 = (L.lo[2] > L.dlo[2]) ? -1 : 0;
    const int kq1 = (L.lo[2] + L.nz - 1 < L.dhi[2]) ? L.nz : L.nz - 1;
    F.NK = kq1 - F.kq0 + 1;
    F.UPR = (L.nx + uw - 1) / uw;
    F.JPS = (F.B + 1) * F.UPR;
    F.NJ = (long long)F.NB * F.NK * F.JPS;
    // P2 of a slab needs P1 of the next slab: two slabs of lag plus the jobs the grid keeps in flight
    F.LAG = 2LL * F.JPS + (long long)lag_per_cta * grid;
    F.mode = mode;
    F.total_tickets = mode == 2 ? 2 * (F.NJ + F.LAG) : F.NJ;
    return F;
}

size_t fused_counter_ints(const Layout& L)
{
    // ticket + one counter per slab; bands are at least 1 row high
    return 1 + (size_t)L.ny * (size_t)(L.nz + 2);
}

template <int UW>
static int launch_fused_t(const Layout& L, const Phys& P, const FusedPlan& F, int grid, const double* fin,
                          const double* gin, double* fout, double* gout, const uint32_t* nbr, const uint8_t* flag,
                          double* qc, double* macro, int* counters, cudaStream_t st)
{
    using C = Cfg<UW>;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_fused<UW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        cudaFuncSetAttribute(k_fused<UW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        attr = true;
    }
    if (macro)
        k_fused<UW, true><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(fin, gin, fout, gout, nbr, flag, qc, macro, L, P, F, counters);
    else
        k_fused<UW, false><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(fin, gin, fout, gout, nbr, flag, qc, macro, L, P, F, counters);
    return 1;
}

int fused_grid(int uw, int sm_count) { return uw == 128 ? 2 * sm_count : sm_count; }

int launch_fused(const Layout& L, const Phys& P, int uw, int band_rows, int lag_per_cta, int mode, int sm_count,
                 const double* fin, const double* gin, double* fout, double* gout, const uint32_t* nbr,
                 const uint8_t* flag, double* qc, double* macro, int* counters, cudaStream_t st)
{
    const int grid = fused_grid(uw, sm_count);
    const FusedPlan F = make_fused_plan(L, uw, band_rows, mode, grid, lag_per_cta);
    // ticket counter always restarts; the completion counters restart whenever P1 jobs run
    const size_t ints = mode == 1 ? 1 : 1 + (size_t)F.NB * F.NK;
    cudaMemsetAsync(counters, 0, ints * sizeof(int), st);
    if (uw == 128)
        return launch_fused_t<128>(L, P, F, grid, fin, gin, fout, gout, nbr, flag, qc, macro, counters, st);
    return launch_fused_t<256>(L, P, F, grid, fin, gin, fout, gout, nbr, flag, qc, macro, counters, st);
}

}  // namespace mbl
