// patch.cu -- kernels of the multi-box / multi-level ("AMR-exact") mode: the reference's un-fused operator
// sequence on AMReX-shaped FABs (valid box + 3 ghost cells), several boxes per launch (blockIdx.y = box).
//
//   k_patch_copy      FabArray::FillBoundary / ParallelCopy as a list of rectangular copy tags (api side builds them)
//   k_patch_prepass   K6 no-slip pre-pass on out-of-domain ghosts          (FillPatchOps.H:92-108)
//   k_patch_bc        BCFill::operator() per face / edge / corner region   (BC.H:345-471, AMReX_PhysBCFunct.H:593-678)
//   k_patch_stream    LBM::stream on the GROWN box, pull form, -1 sentinel (LBM.cpp:558-604)
//   k_patch_qcorr     f_to_macrodata on the valid box grown by 1           (LBM.cpp:810-906)
//   k_patch_collide   compute_q_corrections + macrodata_to_equilibrium + relax_f_to_equilibrium on valid cells
//                     (LBM.cpp:959-991, 621-807), in place
//   k_patch_avgdown   masked_avgdown incl. the coarse ghost ring           (Utilities.H:315-350, Utilities.cpp:5-28)
//   k_patch_interp    CellConservativeLinear without linear limiting = cell_cons_interp
//                     (AMReX_Interpolater.cpp:41, 833-982; AMReX_MFInterp_3D_C.H:176-249)
//
// These are O(cells) one-thread-per-cell kernels with x fastest; the roofline path is the fused single-level
// step (kernels.cu).  Arithmetic of the collision is shared with it (lattice.cuh).
#include "patch.cuh"
#include "bcic.cuh"

namespace mbl {

void patch_init_tables()
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

namespace {

constexpr int PT = 256;  // threads per block

// "keep what the destination has": what k_patch_avgdown writes where all eight fine values are masked, and what a copy
// with skip_keep set does not store.  A quiet NaN with a payload no arithmetic produces.
constexpr long long KEEP_BITS = 0x7ff84d424c4b4550LL;

__device__ __forceinline__ double* parr(const PBox& b, int arr, int cur)
{
    return arr == PA_F ? b.f[cur] : arr == PA_G ? b.g[cur] : arr == PA_QC ? b.qc : b.macro;
}
__device__ __forceinline__ bool in_dom(const PGeom& G, int i, int j, int k)
{
    return i >= G.dlo[0] && i <= G.dhi[0] && j >= G.dlo[1] && j <= G.dhi[1] && k >= G.dlo[2] && k <= G.dhi[2];
}
// cell index of the grown box -> (i, j, k)
__device__ __forceinline__ void decode(const PBox& b, long long t, int& i, int& j, int& k)
{
    i = b.glo[0] + (int)(t % b.n[0]);
    t /= b.n[0];
    j = b.glo[1] + (int)(t % b.n[1]);
    k = b.glo[2] + (int)(t / b.n[1]);
}

__global__ void __launch_bounds__(PT) k_patch_copy(const PBox* __restrict__ dtab, int dcur, const PBox* __restrict__ stab,
                                                   int scur, const CopyTag* __restrict__ tags, int darr, int sarr, int ncomp,
                                                   int skip_keep)
{
    const CopyTag T = tags[blockIdx.y];
    const PBox D = dtab[T.dbox], S = stab[T.sbox];
    double* __restrict__ dst = parr(D, darr, dcur);
    const double* __restrict__ src = parr(S, sarr, scur);
    const long long cells = (long long)T.n[0] * T.n[1] * T.n[2];
    for (long long t = (long long)blockIdx.x * PT + threadIdx.x; t < cells; t += (long long)gridDim.x * PT) {
        const int a = (int)(t % T.n[0]);
        long long r = t / T.n[0];
        const int b = (int)(r % T.n[1]), c = (int)(r / T.n[1]);
        const long long dc = D.cell(T.d[0] + a, T.d[1] + b, T.d[2] + c), sc = S.cell(T.s[0] + a, T.s[1] + b, T.s[2] + c);
        for (int q = 0; q < ncomp; ++q) {
            const double v = src[q * S.sq + sc];
            if (!skip_keep || __double_as_longlong(v) != KEEP_BITS) dst[q * D.sq + dc] = v;
        }
    }
}

// pack (to_buf) the source region of every tag into the message, or unpack the message into the destination regions
__global__ void __launch_bounds__(PT) k_patch_pack(const PBox* __restrict__ tab, int cur, const CopyTag* __restrict__ tags,
                                                   int arr, int ncomp, double* __restrict__ buf, long long base, int to_buf,
                                                   int skip_keep)
{
    const CopyTag T = tags[blockIdx.y];
    const PBox B = tab[to_buf ? T.sbox : T.dbox];
    double* __restrict__ a = parr(B, arr, cur);
    const long long cells = (long long)T.n[0] * T.n[1] * T.n[2];
    double* __restrict__ m = buf + base + T.boff * ncomp;
    const int* o = to_buf ? T.s : T.d;
    for (long long t = (long long)blockIdx.x * PT + threadIdx.x; t < cells; t += (long long)gridDim.x * PT) {
        const int x = (int)(t % T.n[0]);
        long long r = t / T.n[0];
        const int y = (int)(r % T.n[1]), z = (int)(r / T.n[1]);
        const long long c = B.cell(o[0] + x, o[1] + y, o[2] + z);
        if (to_buf)
            for (int q = 0; q < ncomp; ++q) m[q * cells + t] = a[q * B.sq + c];
        else
            for (int q = 0; q < ncomp; ++q) {
                const double v = m[q * cells + t];
                if (!skip_keep || __double_as_longlong(v) != KEEP_BITS) a[q * B.sq + c] = v;
            }
    }
}

__global__ void __launch_bounds__(PT) k_patch_fill(const PBox* __restrict__ tab, int cur, int arr, int ncomp, double v)
{
    const PBox B = tab[blockIdx.y];
    double* __restrict__ a = parr(B, arr, cur);
    const long long total = B.sq * ncomp;
    for (long long t = (long long)blockIdx.x * PT + threadIdx.x; t < total; t += (long long)gridDim.x * PT) a[t] = v;
}

// Initializer<ICOp>::initialize on the grown box + fill_f_inside_eb (IC.H:474-519, LBM.cpp:1287-1295)
__global__ void __launch_bounds__(PT) k_patch_initialize(const PBox* __restrict__ tab, int cur, BcInfo Bc, IcInfo I)
{
    const PBox B = tab[blockIdx.y];
    for (long long t = (long long)blockIdx.x * PT + threadIdx.x; t < B.sq; t += (long long)gridDim.x * PT) {
        int i, j, k;
        decode(B, t, i, j, k);
        double rho, vel[3], T, R, gamma;
        ic_state(I, Bc, i, j, k, rho, vel, T, R, gamma);
        const bool solid = B.isfl[t] == 0;
        for (int q = 0; q < NQ; ++q) {
            B.f[cur][q * B.sq + t] = solid ? 0.0 : feq_std(rho, vel, R * T, q);
            B.g[cur][q * B.sq + t] = solid ? 0.0 : geq_state(rho, vel, T, R, gamma, q);
        }
    }
}

// fill_f_inside_eb (LBM.cpp:1278-1298): f = g = 0 in solid cells of the grown box
__global__ void __launch_bounds__(PT) k_patch_zero_solid(const PBox* __restrict__ tab, int cur)
{
    const PBox B = tab[blockIdx.y];
    for (long long t = (long long)blockIdx.x * PT + threadIdx.x; t < B.sq; t += (long long)gridDim.x * PT) {
        if (B.isfl[t] != 0) continue;
        for (int q = 0; q < NQ; ++q) {
            B.f[cur][q * B.sq + t] = 0.0;
            B.g[cur][q * B.sq + t] = 0.0;
        }
    }
}

// K6: every out-of-domain ghost takes the no-slip value of the in-domain cell it faces (FillPatchOps.H:92-108)
__global__ void __launch_bounds__(PT) k_patch_prepass(const PBox* __restrict__ tab, int cur, PGeom G)
{
    const PBox B = tab[blockIdx.y];
    for (long long t = (long long)blockIdx.x * PT + threadIdx.x; t < B.sq; t += (long long)gridDim.x * PT) {
        int i, j, k;
        decode(B, t, i, j, k);
        if (in_dom(G, i, j, k)) continue;
        for (int q = 0; q < NQ; ++q) {
            const int in = i + c_dir.ex[q], jn = j + c_dir.ey[q], kn = k + c_dir.ez[q];
            if (in_dom(G, in, jn, kn) && B.in_grown(in, jn, kn)) {
                const long long s = B.cell(in, jn, kn) + (long long)c_dir.opp[q] * B.sq;
                B.f[cur][q * B.sq + t] = B.f[cur][s];
                B.g[cur][q * B.sq + t] = B.g[cur][s];
            }
        }
    }
}

// range of the grown box in dimension d on one side of the (periodically grown) domain: side -1 below,
// 0 inside, +1 above (AMReX_PhysBCFunct.H:416-421)
__device__ __forceinline__ bool side_range(const PBox& B, const PGeom& G, int d, int side, int& a0, int& a1)
{
    const int blo = B.glo[d], bhi = B.glo[d] + B.n[d] - 1;
    const int glo = G.periodic[d] ? G.dlo[d] - B.n[d] : G.dlo[d];
    const int ghi = G.periodic[d] ? G.dhi[d] + B.n[d] : G.dhi[d];
    if (side < 0) {
        a0 = blo;
        a1 = min(glo - 1, bhi);
    } else if (side > 0) {
        a0 = max(ghi + 1, blo);
        a1 = bhi;
    } else {
        a0 = max(glo, blo);
        a1 = min(ghi, bhi);
    }
    return a0 <= a1;
}

// BCFill::operator() (BC.H:345-471) for the cells of one face / edge / corner region of every box;
// blockIdx.z = lattice (0: f, 1: g = energy lattice).  `inside` (BC.H:374-380) is the valid box of the FAB.
__global__ void __launch_bounds__(64) k_patch_bc(const PBox* __restrict__ tab, int cur, PGeom G, BcInfo Bc, int sx, int sy,
                                                 int sz)
{
    const PBox B = tab[blockIdx.y];
    int r0[3], r1[3];
    if (!side_range(B, G, 0, sx, r0[0], r1[0]) || !side_range(B, G, 1, sy, r0[1], r1[1]) ||
        !side_range(B, G, 2, sz, r0[2], r1[2]))
        return;
    const int n0 = r1[0] - r0[0] + 1, n1 = r1[1] - r0[1] + 1, n2 = r1[2] - r0[2] + 1;
    const bool energy = blockIdx.z == 1;
    double* __restrict__ data = energy ? B.g[cur] : B.f[cur];
    const long long cells = (long long)n0 * n1 * n2;
    for (long long t = (long long)blockIdx.x * 64 + threadIdx.x; t < cells; t += (long long)gridDim.x * 64) {
        const int i = r0[0] + (int)(t % n0);
        const long long r = t / n0;
        const int j = r0[1] + (int)(r % n1), k = r0[2] + (int)(r / n1);
        const long long c = B.cell(i, j, k);
        const int iv[3] = {i, j, k};
        for (int idir = 0; idir < 3; ++idir) {
            for (int lohi = 0; lohi < 2; ++lohi) {
                if (!((lohi == 0 && iv[idir] < G.dlo[idir]) || (lohi == 1 && iv[idir] > G.dhi[idir]))) continue;
                const int ndir = lohi == 0 ? 1 : -1;
                const int bc = Bc.bc[idir + 3 * lohi];
                double vel[3] = {0.0, 0.0, 0.0}, R = 1.0, T = 1.0 / 3.0, gamma = 5.0 / 3.0;
                double rho_bc = (bc == 3) ? 1.0 : 0.0;
                if (bc == 2 || bc == 3) vel_bc_op(Bc, G.dhi, i, j, k, rho_bc, vel, R, T, gamma);
                for (int q = 0; q < NQ; ++q) {
                    const int in = i + c_dir.ex[q], jn = j + c_dir.ey[q], kn = k + c_dir.ez[q];
                    double* dst = data + (long long)q * B.sq + c;
                    if (!B.in_valid(in, jn, kn)) {
                        *dst = -1.0;  // BC.H:463-466
                        continue;
                    }
                    const long long cn = B.cell(in, jn, kn);
                    if (bc == 1) {  // NOSLIPWALL, BC.H:81
                        *dst = data[(long long)c_dir.opp[q] * B.sq + cn];
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
                                const long long cs = cn + c_dir.ex[qq] + c_dir.ey[qq] * B.sy + c_dir.ez[qq] * B.sz;
                                if (ei[idir] == -ndir)
                                    rho_out += 2.0 * data[(long long)bq * B.sq + cs];
                                else if (ei[idir] == 0)
                                    rho_tan += data[(long long)bq * B.sq + cs];
                            }
                            vb[idir] = ndir * (1.0 - (rho_out + rho_tan) / rho_bc);
                            *dst = feq_std(rho_bc, vb, R * T, q);
                        }
                    } else if (bc == 5) {
                        // OUTFLOW_ZEROTH_ORDER, BC.H:339-341.  The reference copies the inward neighbour, which for the
                        // ghost layers beyond the first is another ghost cell of the same region (its value then depends
                        // on the CPU loop order).  Only the first layer is ever read; deeper layers take the in-domain
                        // cell on their normal here, which is race-free.
                        int s[3] = {i, j, k};
                        s[idir] = lohi == 0 ? G.dlo[idir] : G.dhi[idir];
                        *dst = data[(long long)q * B.sq + B.cell(s[0], s[1], s[2])];
                    } else if (bc == 6) {  // SLIPWALLXNORMAL, BC.H:100
                        *dst = data[(long long)c_dir.mx[q] * B.sq + cn];
                    } else if (bc == 7) {
                        *dst = data[(long long)c_dir.my[q] * B.sq + cn];
                    } else if (bc == 8) {
                        *dst = data[(long long)c_dir.mz[q] * B.sq + cn];
                    }
                }
            }
        }
    }
}

// one population of LBM::stream in pull form (LBM.cpp:565-599): -1 where the cell is not fluid or the
// source lies outside the FAB, the cell's own opposite population where the source is solid (halfway bounce-back)
__device__ __forceinline__ void patch_pull(const PBox& B, const double* __restrict__ fin, const double* __restrict__ gin,
                                           long long t, int i, int j, int k, bool fluid, int q, double& vf, double& vg)
{
    vf = -1.0, vg = -1.0;  // f_star.setVal(-1), LBM.cpp:565
    const int is = i - c_dir.ex[q], js = j - c_dir.ey[q], ks = k - c_dir.ez[q];
    if (fluid && B.in_grown(is, js, ks)) {
        const long long s = B.cell(is, js, ks);
        const int fs = B.isfl[s];
        if (fs == 1) {
            vf = fin[q * B.sq + s];
            vg = gin[q * B.sq + s];
        } else if (fs == 0) {  // LBM.cpp:590-595
            vf = fin[(long long)c_dir.opp[q] * B.sq + t];
            vg = gin[(long long)c_dir.opp[q] * B.sq + t];
        }
    }
}

// The same for all 27 populations of a cell, in two phases so that no load waits for a branch: first the 27 source
// flags -> element offsets (component included; -1 = sentinel), then the 54 values
__device__ __forceinline__ void patch_pull_all(const PBox& B, const double* __restrict__ fin, const double* __restrict__ gin,
                                               long long t, int i, int j, int k, bool fluid, double (&fv)[NQ], double (&gv)[NQ])
{
    int off[NQ];
    const int n = (int)B.sq, tt = (int)t;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int is = i - c_dir.ex[q], js = j - c_dir.ey[q], ks = k - c_dir.ez[q];
        const bool ok = fluid && B.in_grown(is, js, ks);
        const int s = ok ? (int)B.cell(is, js, ks) : tt;
        const int fs = ok ? B.isfl[s] : -1;
        off[q] = fs == 1 ? q * n + s : fs == 0 ? c_dir.opp[q] * n + tt : -1;
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int o = off[q] < 0 ? tt : off[q];
        const double a = fin[o], b = gin[o];
        fv[q] = off[q] < 0 ? -1.0 : a;
        gv[q] = off[q] < 0 ? -1.0 : b;
    }
}

// LBM::stream on the grown box in pull form: buffers [cur] -> [1 - cur]
__global__ void __launch_bounds__(PT) k_patch_stream(const PBox* __restrict__ tab, int cur)
{
    const PBox B = tab[blockIdx.y];
    const double* __restrict__ fin = B.f[cur];
    const double* __restrict__ gin = B.g[cur];
    double* __restrict__ fout = B.f[1 - cur];
    double* __restrict__ gout = B.g[1 - cur];
    for (long long t = (long long)blockIdx.x * PT + threadIdx.x; t < B.sq; t += (long long)gridDim.x * PT) {
        int i, j, k;
        decode(B, t, i, j, k);
        const bool fluid = B.isfl[t] == 1;
        for (int q = 0; q < NQ; ++q) {
            double vf, vg;
            patch_pull(B, fin, gin, t, i, j, k, fluid, q, vf, vg);
            fout[q * B.sq + t] = vf;
            gout[q * B.sq + t] = vg;
        }
    }
}

// f_to_macrodata on the valid box grown by 1 (LBM.cpp:823-903): QCorr always, all 19 fields on request
// PULL: from the state BEFORE the stream (the populations are pulled as LBM::stream would move them): the first half of
// the fused advance of a finest level
template <bool MACRO, bool PULL>
__global__ void __launch_bounds__(PT) k_patch_qcorr(const PBox* __restrict__ tab, int cur, Phys P)
{
    const PBox B = tab[blockIdx.y];
    const int m0 = B.n[0] - 2 * (PNG - 1), m1 = B.n[1] - 2 * (PNG - 1), m2 = B.n[2] - 2 * (PNG - 1);
    const long long cells = (long long)m0 * m1 * m2;
    const double* __restrict__ f = B.f[cur];
    const double* __restrict__ g = B.g[cur];
    for (long long t = (long long)blockIdx.x * PT + threadIdx.x; t < cells; t += (long long)gridDim.x * PT) {
        const int i = B.lo[0] - 1 + (int)(t % m0);
        const long long r = t / m0;
        const int j = B.lo[1] - 1 + (int)(r % m1), k = B.lo[2] - 1 + (int)(r / m1);
        const long long c = B.cell(i, j, k);
        if (B.isfl[c] != 1) continue;
        const long long n = B.sq;
        MomF mf;
        MomG mg;
        if constexpr (PULL) {
            double fv[NQ], gv[NQ];
            patch_pull_all(B, f, g, c, i, j, k, true, fv, gv);
            mf = moments_f([&](int q) { return fv[q]; });
            mg = moments_g([&](int q) { return gv[q]; });
        } else {
            mf = moments_f([&](int q) { return f[q * n + c]; });
            mg = moments_g([&](int q) { return g[q * n + c]; });
        }
        const Prim s = primitives(mf.rho, mf.jx, mf.jy, mf.jz, mg.e2, P);
        B.qc[c] = s.qcx;
        B.qc[n + c] = s.qcy;
        B.qc[2 * n + c] = s.qcz;
        if constexpr (MACRO) {
            double* __restrict__ macro = B.macro;
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
    }
}

// gradient() of Utilities.H:279-312: neighbour usable = inside the DOMAIN box and fluid
__device__ __forceinline__ double patch_gradient(const PBox& B, const PGeom& G, const double* __restrict__ a, long long c,
                                                 int i, int j, int k, int dir, double idx)
{
    const int e[3] = {dir == 0, dir == 1, dir == 2};
    const long long step = dir == 0 ? 1 : dir == 1 ? B.sy : B.sz;
    const bool okp = in_dom(G, i + e[0], j + e[1], k + e[2]) && B.isfl[c + step] == 1;
    const bool okm = in_dom(G, i - e[0], j - e[1], k - e[2]) && B.isfl[c - step] == 1;
    return one_sided_gradient(okp, okm, okp ? a[c + step] : 0.0, a[c], okm ? a[c - step] : 0.0, idx);
}

template <bool MACRO>
__global__ void __launch_bounds__(128) k_patch_collide(const PBox* __restrict__ tab, int cur, PGeom G, Phys P)
{
    const PBox B = tab[blockIdx.y];
    const int m0 = B.hi[0] - B.lo[0] + 1, m1 = B.hi[1] - B.lo[1] + 1, m2 = B.hi[2] - B.lo[2] + 1;
    const long long cells = (long long)m0 * m1 * m2;
    double* __restrict__ f = B.f[cur];
    double* __restrict__ g = B.g[cur];
    const long long n = B.sq;
    for (long long t = (long long)blockIdx.x * 128 + threadIdx.x; t < cells; t += (long long)gridDim.x * 128) {
        const int i = B.lo[0] + (int)(t % m0);
        const long long r = t / m0;
        const int j = B.lo[1] + (int)(r % m1), k = B.lo[2] + (int)(r / m1);
        const long long c = B.cell(i, j, k);
        if (B.isfl[c] != 1) continue;
        double fv[NQ], gv[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            fv[q] = f[q * n + c];
            gv[q] = g[q * n + c];
        }
        const MomF mf = moments_f([&](int q) { return fv[q]; });
        const MomG mg = moments_g([&](int q) { return gv[q]; });
        const Prim s = primitives(mf.rho, mf.jx, mf.jy, mf.jz, mg.e2, P);
        const double dqx = patch_gradient(B, G, B.qc, c, i, j, k, 0, P.idx[0]);
        const double dqy = patch_gradient(B, G, B.qc + n, c, i, j, k, 1, P.idx[1]);
        const double dqz = patch_gradient(B, G, B.qc + 2 * n, c, i, j, k, 2, P.idx[2]);
        if constexpr (MACRO) {
            B.macro[23 * n + c] = dqx;
            B.macro[24 * n + c] = dqy;
            B.macro[25 * n + c] = dqz;
        }
        const Coll cc = collision_coefficients(s, mf, mg, dqx, dqy, dqz, P);
        static_for<0, NQ>([&](auto qc_) {
            constexpr int Q = decltype(qc_)::value;
            f[Q * n + c] = fv[Q] + cc.omega * (feq_q<Q>(cc) - fv[Q]);
            g[Q * n + c] = gv[Q] + cc.omega * (geq_q<Q>(cc) - gv[Q]);
        });
    }
}

// Fused LBM::advance of a FINEST level (stream + collide, LBM.cpp:523-544 without the average-down in between):
// every cell of the grown box pulls its populations as LBM::stream would move them; a valid fluid cell collides them in
// registers (q-corrections of the streamed state from k_patch_qcorr<., true>), every other cell stores the streamed
// values -- the ghost cells a neighbour box does not cover keep them, as in the reference, the others are overwritten
// by the FillBoundary that follows.  Buffers [cur] -> [1 - cur].  Same device functions, same order: bit-identical to
// k_patch_stream + k_patch_collide.
template <bool MACRO>
__global__ void __launch_bounds__(128) k_patch_advance(const PBox* __restrict__ tab, int cur, PGeom G, Phys P)
{
    const PBox B = tab[blockIdx.y];
    const double* __restrict__ fin = B.f[cur];
    const double* __restrict__ gin = B.g[cur];
    double* __restrict__ fout = B.f[1 - cur];
    double* __restrict__ gout = B.g[1 - cur];
    const long long n = B.sq;
    for (long long t = (long long)blockIdx.x * 128 + threadIdx.x; t < n; t += (long long)gridDim.x * 128) {
        int i, j, k;
        decode(B, t, i, j, k);
        const bool fluid = B.isfl[t] == 1;
        double fv[NQ], gv[NQ];
        patch_pull_all(B, fin, gin, t, i, j, k, fluid, fv, gv);
        const bool valid = i >= B.lo[0] && i <= B.hi[0] && j >= B.lo[1] && j <= B.hi[1] && k >= B.lo[2] && k <= B.hi[2];
        if (!(valid && fluid)) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                fout[q * n + t] = fv[q];
                gout[q * n + t] = gv[q];
            }
            continue;
        }
        const MomF mf = moments_f([&](int q) { return fv[q]; });
        const MomG mg = moments_g([&](int q) { return gv[q]; });
        const Prim s = primitives(mf.rho, mf.jx, mf.jy, mf.jz, mg.e2, P);
        const double dqx = patch_gradient(B, G, B.qc, t, i, j, k, 0, P.idx[0]);
        const double dqy = patch_gradient(B, G, B.qc + n, t, i, j, k, 1, P.idx[1]);
        const double dqz = patch_gradient(B, G, B.qc + 2 * n, t, i, j, k, 2, P.idx[2]);
        if constexpr (MACRO) {
            B.macro[23 * n + t] = dqx;
            B.macro[24 * n + t] = dqy;
            B.macro[25 * n + t] = dqz;
        }
        const Coll cc = collision_coefficients(s, mf, mg, dqx, dqy, dqz, P);
        static_for<0, NQ>([&](auto qc_) {
            constexpr int Q = decltype(qc_)::value;
            fout[Q * n + t] = fv[Q] + cc.omega * (feq_q<Q>(cc) - fv[Q]);
            gout[Q * n + t] = gv[Q] + cc.omega * (geq_q<Q>(cc) - gv[Q]);
        });
    }
}

// compute_eb_forces (LBM.cpp:994-1044) on the valid cells of every local box: momentum exchange over the solid cells that
// have a fluid face neighbour (is_fluid comp 1, LBM.cpp:1236-1259), reading the populations of the neighbour cells
// (ghost cells included: FillBoundary'd after the collision).  One double atomicAdd per block and direction.
__global__ void __launch_bounds__(128) k_patch_eb_forces(const PBox* __restrict__ tab, int cur, double* __restrict__ out3)
{
    const PBox B = tab[blockIdx.y];
    const int m0 = B.hi[0] - B.lo[0] + 1, m1 = B.hi[1] - B.lo[1] + 1, m2 = B.hi[2] - B.lo[2] + 1;
    const long long cells = (long long)m0 * m1 * m2;
    const double* __restrict__ f = B.f[cur];
    double fs[3] = {0.0, 0.0, 0.0};
    for (long long t = (long long)blockIdx.x * 128 + threadIdx.x; t < cells; t += (long long)gridDim.x * 128) {
        const int i = B.lo[0] + (int)(t % m0);
        const long long r = t / m0;
        const int j = B.lo[1] + (int)(r % m1), k = B.lo[2] + (int)(r / m1);
        const long long c = B.cell(i, j, k);
        if (B.isfl[c] == 1) continue;
        const bool all_covered = B.isfl[c - 1] == 0 && B.isfl[c + 1] == 0 && B.isfl[c - B.sy] == 0 && B.isfl[c + B.sy] == 0 &&
                                 B.isfl[c - B.sz] == 0 && B.isfl[c + B.sz] == 0;
        if (all_covered) continue;
        for (int q = 0; q < NQ; ++q) {
            const int o = c_dir.opp[q];
            const long long rr = c + c_dir.ex[o] + c_dir.ey[o] * B.sy + c_dir.ez[o] * B.sz;
            const double v = 2.0 * f[(long long)q * B.sq + rr] * B.isfl[rr];
            fs[0] += c_dir.ex[q] * v;
            fs[1] += c_dir.ey[q] * v;
            fs[2] += c_dir.ez[q] * v;
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
        const double v = sh[threadIdx.x][0] + sh[threadIdx.x][1] + sh[threadIdx.x][2] + sh[threadIdx.x][3];
        if (v != 0.0) atomicAdd(out3 + threadIdx.x, v);
    }
}

// compute_derived (LBM.cpp:909-955): vorticity from the velocity macrodata (1 ghost cell, FillBoundary'd)
__global__ void __launch_bounds__(PT) k_patch_derived(const PBox* __restrict__ tab, PGeom G, Phys P, int with_dq)
{
    const PBox B = tab[blockIdx.y];
    const int m0 = B.hi[0] - B.lo[0] + 1, m1 = B.hi[1] - B.lo[1] + 1, m2 = B.hi[2] - B.lo[2] + 1;
    const long long cells = (long long)m0 * m1 * m2;
    const long long n = B.sq;
    double* __restrict__ macro = B.macro;
    for (long long t = (long long)blockIdx.x * PT + threadIdx.x; t < cells; t += (long long)gridDim.x * PT) {
        const int i = B.lo[0] + (int)(t % m0);
        const long long r = t / m0;
        const int j = B.lo[1] + (int)(r % m1), k = B.lo[2] + (int)(r / m1);
        const long long c = B.cell(i, j, k);
        if (B.isfl[c] != 1) continue;
        auto grad = [&](int dir, int comp) { return patch_gradient(B, G, macro + (long long)comp * n, c, i, j, k, dir, P.idx[dir]); };
        const double vx = grad(0, 2), wx = grad(0, 3), uy = grad(1, 1), wy = grad(1, 3), uz = grad(2, 1), vz = grad(2, 2);
        macro[19 * n + c] = wy - vz;
        macro[20 * n + c] = uz - wx;
        macro[21 * n + c] = vx - uy;
        macro[22 * n + c] = sqrt((wy - vz) * (wy - vz) + (uz - wx) * (uz - wx) + (vx - uy) * (vx - uy));
        if (with_dq) {
            macro[23 * n + c] = grad(0, 6);
            macro[24 * n + c] = grad(1, 7);
            macro[25 * n + c] = grad(2, 8);
        }
    }
}

// masked_avgdown (Utilities.H:315-350): ctab[b] is the coarsened fine box b grown by ng.  Where all eight fine values
// carry the -1 sentinel the reference keeps the coarse level's value (it copied the coarse level into these boxes
// first): here the cell gets the KEEP marker and the copy back into the coarse level (skip_keep) leaves the
// destination alone -- the same result without moving the coarse level into the boxes
__global__ void __launch_bounds__(PT) k_patch_avgdown(const PBox* __restrict__ ftab, int fcur, const PBox* __restrict__ ctab)
{
    const PBox F = ftab[blockIdx.y], Cb = ctab[blockIdx.y];
    const double small_num = 2.220446049250313e-16 * 1e10;  // constants::SMALL_NUM, Constants.H:57-58
    for (long long t = (long long)blockIdx.x * PT + threadIdx.x; t < Cb.sq; t += (long long)gridDim.x * PT) {
        int i, j, k;
        decode(Cb, t, i, j, k);
        for (int lat = 0; lat < 2; ++lat) {
            const double* __restrict__ src = lat ? F.g[fcur] : F.f[fcur];
            double* __restrict__ dst = lat ? Cb.g[0] : Cb.f[0];
            for (int q = 0; q < NQ; ++q) {
                double c = 0.0, vol = 0.0;
                for (int kr = 0; kr < 2; ++kr)
                    for (int jr = 0; jr < 2; ++jr)
                        for (int ir = 0; ir < 2; ++ir) {
                            const double fv = src[q * F.sq + F.cell(2 * i + ir, 2 * j + jr, 2 * k + kr)];
                            if (fabs(fv - (-1.0)) > small_num) {
                                c += fv;
                                vol += 1.0;
                            }
                        }
                dst[q * Cb.sq + t] = vol > 0.0 ? c / vol : __longlong_as_double(KEEP_BITS);
            }
        }
    }
}

// one population of one fine cell by CellConservativeLinear with the monotonized-central slopes of
// mf_cell_cons_lin_interp_mcslope (AMReX_MFInterp_3D_C.H:176-232) and mf_cell_cons_lin_interp (:234-249), ratio 2.
// Products and sums are written with explicit roundings so that no FMA changes the reference's result.
__device__ __forceinline__ double cons_lin_interp(const double* __restrict__ u, const PBox& Cb, int ic, int jc, int kc,
                                                  double xoff, double yoff, double zoff)
{
    const long long c = Cb.cell(ic, jc, kc);
    const double u0 = u[c];
    auto slope = [&](long long step) {
        const double up = u[c + step], um = u[c - step];
        const double dc = __dmul_rn(0.5, __dsub_rn(up, um));
        const double df = __dmul_rn(2.0, __dsub_rn(up, u0));
        const double db = __dmul_rn(2.0, __dsub_rn(u0, um));
        double s = (__dmul_rn(df, db) >= 0.0) ? fmin(fabs(df), fabs(db)) : 0.0;
        return copysign(1.0, dc) * fmin(s, fabs(dc));
    };
    double sx = slope(1), sy = slope(Cb.sy), sz = slope(Cb.sz);
    double alpha = 1.0;
    if (sx != 0.0 || sy != 0.0 || sz != 0.0) {
        const double dumax = __dadd_rn(__dadd_rn(__dmul_rn(fabs(sx), 0.25), __dmul_rn(fabs(sy), 0.25)), __dmul_rn(fabs(sz), 0.25));
        double umax = u0, umin = u0;
        for (int dk = -1; dk <= 1; ++dk)
            for (int dj = -1; dj <= 1; ++dj)
                for (int di = -1; di <= 1; ++di) {
                    const double v = u[c + di + dj * Cb.sy + dk * Cb.sz];
                    umin = fmin(umin, v);
                    umax = fmax(umax, v);
                }
        if (__dmul_rn(dumax, alpha) > __dsub_rn(umax, u0)) alpha = __ddiv_rn(__dsub_rn(umax, u0), dumax);
        if (__dmul_rn(dumax, alpha) > __dsub_rn(u0, umin)) alpha = __ddiv_rn(__dsub_rn(u0, umin), dumax);
    }
    sx = __dmul_rn(sx, alpha), sy = __dmul_rn(sy, alpha), sz = __dmul_rn(sz, alpha);
    return __dadd_rn(__dadd_rn(__dadd_rn(u0, __dmul_rn(xoff, sx)), __dmul_rn(yoff, sy)), __dmul_rn(zoff, sz));
}

// fine ghost regions from the coarse patch ctab[box] = coarsen(grown fine box) grown by 1
__global__ void __launch_bounds__(PT) k_patch_interp(const PBox* __restrict__ ftab, int fcur, const PBox* __restrict__ ctab,
                                                     const RegionTag* __restrict__ regs)
{
    const RegionTag R = regs[blockIdx.y];
    const PBox F = ftab[R.box], Cb = ctab[R.box];
    const long long cells = (long long)R.n[0] * R.n[1] * R.n[2];
    for (long long t = (long long)blockIdx.x * PT + threadIdx.x; t < cells; t += (long long)gridDim.x * PT) {
        const int i = R.lo[0] + (int)(t % R.n[0]);
        const long long r = t / R.n[0];
        const int j = R.lo[1] + (int)(r % R.n[1]), k = R.lo[2] + (int)(r / R.n[1]);
        const int ic = i >> 1, jc = j >> 1, kc = k >> 1;  // arithmetic shift = floor for negative indices
        const double xoff = (i - 2 * ic) ? 0.25 : -0.25, yoff = (j - 2 * jc) ? 0.25 : -0.25, zoff = (k - 2 * kc) ? 0.25 : -0.25;
        const long long c = F.cell(i, j, k);
        for (int q = 0; q < NQ; ++q) {
            F.f[fcur][q * F.sq + c] = cons_lin_interp(Cb.f[0] + q * Cb.sq, Cb, ic, jc, kc, xoff, yoff, zoff);
            F.g[fcur][q * F.sq + c] = cons_lin_interp(Cb.g[0] + q * Cb.sq, Cb, ic, jc, kc, xoff, yoff, zoff);
        }
    }
}

inline unsigned blocks_for(long long cells, int threads)
{
    long long b = (cells + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > 16384) b = 16384;
    return (unsigned)b;
}

}  // namespace

int launch_patch_copy(const PBox* dtab, int dcur, const PBox* stab, int scur, const CopyTag* tags, int ntags, int darr,
                      int sarr, int ncomp, long long max_cells, cudaStream_t st, bool skip_keep)
{
    if (ntags <= 0) return 0;
    for (int t0 = 0; t0 < ntags; t0 += 65535) {
        const int nt = ntags - t0 < 65535 ? ntags - t0 : 65535;
        k_patch_copy<<<dim3(blocks_for(max_cells, PT), nt), PT, 0, st>>>(dtab, dcur, stab, scur, tags + t0, darr, sarr, ncomp,
                                                                         skip_keep ? 1 : 0);
    }
    return (ntags + 65534) / 65535;
}

int launch_patch_pack(const PBox* tab, int cur, const CopyTag* tags, int ntags, int arr, int ncomp, long long max_cells,
                      double* buf, long long base, bool to_buf, cudaStream_t st, bool skip_keep)
{
    if (ntags <= 0) return 0;
    for (int t0 = 0; t0 < ntags; t0 += 65535) {
        const int nt = ntags - t0 < 65535 ? ntags - t0 : 65535;
        k_patch_pack<<<dim3(blocks_for(max_cells, PT), nt), PT, 0, st>>>(tab, cur, tags + t0, arr, ncomp, buf, base, to_buf ? 1 : 0,
                                                                         skip_keep ? 1 : 0);
    }
    return (ntags + 65534) / 65535;
}

int launch_patch_fill(const PBox* tab, int nb, long long max_cells, int cur, int arr, int ncomp, double v, cudaStream_t st)
{
    if (nb <= 0) return 0;  // this rank holds no box of the level
    k_patch_fill<<<dim3(blocks_for(max_cells * ncomp, PT), nb), PT, 0, st>>>(tab, cur, arr, ncomp, v);
    return 1;
}

int launch_patch_initialize(const PBox* tab, int nb, long long max_cells, int cur, const BcInfo& B, const IcInfo& I,
                            cudaStream_t st)
{
    if (nb <= 0) return 0;  // this rank holds no box of the level
    k_patch_initialize<<<dim3(blocks_for(max_cells, PT), nb), PT, 0, st>>>(tab, cur, B, I);
    return 1;
}

int launch_patch_zero_solid(const PBox* tab, int nb, long long max_cells, int cur, cudaStream_t st)
{
    if (nb <= 0) return 0;  // this rank holds no box of the level
    k_patch_zero_solid<<<dim3(blocks_for(max_cells, PT), nb), PT, 0, st>>>(tab, cur);
    return 1;
}

int launch_patch_prepass(const PBox* tab, int nb, long long max_cells, int cur, const PGeom& G, cudaStream_t st)
{
    if (nb <= 0) return 0;  // this rank holds no box of the level
    k_patch_prepass<<<dim3(blocks_for(max_cells, PT), nb), PT, 0, st>>>(tab, cur, G);
    return 1;
}

int launch_patch_physbc(const PBox* tab, int nb, long long max_face, int cur, const PGeom& G, const BcInfo& B,
                        cudaStream_t st)
{
    if (nb <= 0) return 0;  // this rank holds no box of the level
    if (G.periodic[0] && G.periodic[1] && G.periodic[2]) return 0;  // PhysBCFunct::operator() returns early
    int nl = 0;
    // a region has at most the cells of a 3-cell-thick face slab of the grown box (the kernel strides if it has more)
    const dim3 grid(blocks_for(max_face + 64, 64), nb, 2);
    auto region = [&](int sx, int sy, int sz) {
        // a region outside a periodic direction's grown domain is empty for every box: skip the launch
        if ((sx && G.periodic[0]) || (sy && G.periodic[1]) || (sz && G.periodic[2])) return;
        k_patch_bc<<<grid, 64, 0, st>>>(tab, cur, G, B, sx, sy, sz);
        ++nl;
    };
    // faces (xlo ylo zlo xhi yhi zhi), edges, corners: AMReX_PhysBCFunct.H:593-678
    region(-1, 0, 0), region(0, -1, 0), region(0, 0, -1), region(+1, 0, 0), region(0, +1, 0), region(0, 0, +1);
    for (int s1 = -1; s1 <= 1; s1 += 2)
        for (int s0 = -1; s0 <= 1; s0 += 2) region(s0, s1, 0);
    for (int s1 = -1; s1 <= 1; s1 += 2)
        for (int s0 = -1; s0 <= 1; s0 += 2) region(s0, 0, s1);
    for (int s1 = -1; s1 <= 1; s1 += 2)
        for (int s0 = -1; s0 <= 1; s0 += 2) region(0, s0, s1);
    for (int s2 = -1; s2 <= 1; s2 += 2)
        for (int s1 = -1; s1 <= 1; s1 += 2)
            for (int s0 = -1; s0 <= 1; s0 += 2) region(s0, s1, s2);
    return nl;
}

int launch_patch_stream(const PBox* tab, int nb, long long max_cells, int cur, cudaStream_t st)
{
    if (nb <= 0) return 0;  // this rank holds no box of the level
    k_patch_stream<<<dim3(blocks_for(max_cells, PT), nb), PT, 0, st>>>(tab, cur);
    return 1;
}

int launch_patch_qcorr(const PBox* tab, int nb, long long max_cells, int cur, const Phys& P, int want_macro, cudaStream_t st,
                       bool pull)
{
    if (nb <= 0) return 0;  // this rank holds no box of the level
    const dim3 grid(blocks_for(max_cells, PT), nb);
    if (pull) {
        if (want_macro)
            k_patch_qcorr<true, true><<<grid, PT, 0, st>>>(tab, cur, P);
        else
            k_patch_qcorr<false, true><<<grid, PT, 0, st>>>(tab, cur, P);
    } else if (want_macro)
        k_patch_qcorr<true, false><<<grid, PT, 0, st>>>(tab, cur, P);
    else
        k_patch_qcorr<false, false><<<grid, PT, 0, st>>>(tab, cur, P);
    return 1;
}

int launch_patch_advance(const PBox* tab, int nb, long long max_cells, int cur, const PGeom& G, const Phys& P, int want_macro,
                         cudaStream_t st)
{
    if (nb <= 0) return 0;  // this rank holds no box of the level
    const dim3 grid(blocks_for(max_cells, 128), nb);
    if (want_macro)
        k_patch_advance<true><<<grid, 128, 0, st>>>(tab, cur, G, P);
    else
        k_patch_advance<false><<<grid, 128, 0, st>>>(tab, cur, G, P);
    return 1;
}

int launch_patch_collide(const PBox* tab, int nb, long long max_cells, int cur, const PGeom& G, const Phys& P, int want_macro,
                         cudaStream_t st)
{
    if (nb <= 0) return 0;  // this rank holds no box of the level
    const dim3 grid(blocks_for(max_cells, 128), nb);
    if (want_macro)
        k_patch_collide<true><<<grid, 128, 0, st>>>(tab, cur, G, P);
    else
        k_patch_collide<false><<<grid, 128, 0, st>>>(tab, cur, G, P);
    return 1;
}

int launch_patch_eb_forces(const PBox* tab, int nb, long long max_cells, int cur, double* d_out3, cudaStream_t st)
{
    cudaMemsetAsync(d_out3, 0, 3 * sizeof(double), st);
    if (nb <= 0) return 0;
    k_patch_eb_forces<<<dim3(blocks_for(max_cells, 128), nb), 128, 0, st>>>(tab, cur, d_out3);
    return 1;
}

int launch_patch_derived(const PBox* tab, int nb, long long max_cells, const PGeom& G, const Phys& P, int with_dq,
                         cudaStream_t st)
{
    if (nb <= 0) return 0;  // this rank holds no box of the level
    k_patch_derived<<<dim3(blocks_for(max_cells, PT), nb), PT, 0, st>>>(tab, G, P, with_dq);
    return 1;
}

int launch_patch_avgdown(const PBox* ftab, int fcur, const PBox* ctab, int nb, long long max_cells, int /*ng*/,
                         cudaStream_t st)
{
    if (nb <= 0) return 0;  // this rank holds no box of the level
    k_patch_avgdown<<<dim3(blocks_for(max_cells, PT), nb), PT, 0, st>>>(ftab, fcur, ctab);
    return 1;
}

int launch_patch_interp(const PBox* ftab, int fcur, const PBox* ctab, const RegionTag* regs, int nregs, long long max_cells,
                        cudaStream_t st)
{
    if (nregs <= 0) return 0;
    for (int t0 = 0; t0 < nregs; t0 += 65535) {
        const int nt = nregs - t0 < 65535 ? nregs - t0 : 65535;
        k_patch_interp<<<dim3(blocks_for(max_cells, PT), nt), PT, 0, st>>>(ftab, fcur, ctab, regs + t0);
    }
    return (nregs + 65534) / 65535;
}

}  // namespace mbl
