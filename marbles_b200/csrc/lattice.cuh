// lattice.cuh -- D3Q27 tables, device layout and per-cell device math shared by
// every kernel of the marbles_b200 hot path (sm_100a).
//
// Reference semantics (cited per function): Source/Stencil.H:49-167 (tables),
// Source/Utilities.H:11-312 (equilibria, moments, gradient), Source/LBM.cpp:621-906
// (moments, q-correction, relaxation).  The arithmetic is re-derived for the GPU:
// all per-cell scalars are hoisted out of the direction loop, directions are
// compile-time constants so lattice velocities fold into adds/subs, and divisions
// are shared.  Results agree with the reference to round-off (tests/: <= 1e-12 of
// the field scale per step); they are not bit-identical to the CPU build because
// nvcc contracts a*b+c into FMA.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mbl {

constexpr int NQ = 27;
constexpr double THETA0 = 1.0 / 3.0;  // Source/Stencil.H:42

// lattice velocities in the reference's order: rest, 6 axis, 12 face diagonals, 8 body
// diagonals, each (e, -e) pair adjacent (Source/Stencil.H:49-84)
__host__ __device__ constexpr int ex(int q)
{
    constexpr int t[NQ] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1};
    return t[q];
}
__host__ __device__ constexpr int ey(int q)
{
    constexpr int t[NQ] = {0, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1, -1, 1, 1, -1};
    return t[q];
}
__host__ __device__ constexpr int ez(int q)
{
    constexpr int t[NQ] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1, 1, -1, 1, -1, -1, 1, -1, 1};
    return t[q];
}
// opposite direction: pairs are adjacent, so q^1 shifted by the rest state (Stencil.H:126-134)
__host__ __device__ constexpr int opp(int q) { return q == 0 ? 0 : ((q - 1) ^ 1) + 1; }
__host__ __device__ constexpr int find_dir(int x, int y, int z)
{
    for (int q = 0; q < NQ; ++q)
        if (ex(q) == x && ey(q) == y && ez(q) == z) return q;
    return -1;
}
// mirror tables used by the slip walls (Stencil.H:136-167)
__host__ __device__ constexpr int mirror_x(int q) { return find_dir(-ex(q), ey(q), ez(q)); }
__host__ __device__ constexpr int mirror_y(int q) { return find_dir(ex(q), -ey(q), ez(q)); }
__host__ __device__ constexpr int mirror_z(int q) { return find_dir(ex(q), ey(q), -ez(q)); }
// weights (Stencil.H:80-90); w0 is computed the way the reference computes it
__host__ __device__ constexpr double weight(int q)
{
    const int s = (ex(q) != 0) + (ey(q) != 0) + (ez(q) != 0);
    return s == 0 ? 1.0 - (6.0 * (2.0 / 27.0) + 12.0 * (1.0 / 54.0) + 8.0 * (1.0 / 216.0))
         : s == 1 ? 2.0 / 27.0
         : s == 2 ? 1.0 / 54.0
                  : 1.0 / 216.0;
}

// ---------------------------------------------------------------------------
// device layout: structure of arrays [q][k][j][i] over the valid box grown by
// one ghost cell in x and y and GZ ghost planes in z; rows are padded to a
// multiple of 16 doubles and valid cell i=0 sits at an even offset so that
// 128-bit accesses of aligned cell pairs are legal.
// ---------------------------------------------------------------------------
constexpr int GX = 1, GY = 1, GZ = 2, OX = 2;

struct Layout {
    int nx, ny, nz;   // valid cells of the local box
    int lo[3];        // global index of local cell (0,0,0)
    int dlo[3], dhi[3];  // level domain
    long long px, sz, sq; // row pitch, plane stride, component stride (doubles)
    int wrap[3];      // the kernels wrap source indices in this direction themselves (set only when every
                      // direction is periodic: then no ghost cell of the box is ever read)
    __host__ __device__ long long cell(int i, int j, int k) const
    {
        return (long long)(i + OX) + (long long)(j + GY) * px + (long long)(k + GZ) * sz;
    }
    __host__ __device__ int nyp() const { return ny + 2 * GY; }
    __host__ __device__ int nzp() const { return nz + 2 * GZ; }
};

inline Layout make_layout(const int lo[3], const int hi[3], const int dlo[3], const int dhi[3])
{
    Layout L;
    L.nx = hi[0] - lo[0] + 1;
    L.ny = hi[1] - lo[1] + 1;
    L.nz = hi[2] - lo[2] + 1;
    for (int d = 0; d < 3; ++d) {
        L.lo[d] = lo[d];
        L.dlo[d] = dlo[d];
        L.dhi[d] = dhi[d];
    }
    L.px = ((long long)(L.nx + OX + GX) + 15) / 16 * 16;
    L.sz = L.px * (L.ny + 2 * GY);
    L.sq = L.sz * (L.nz + 2 * GZ);
    L.wrap[0] = L.wrap[1] = L.wrap[2] = 0;
    return L;
}

// Offsets (in doubles) from a cell to the cell one step back along each axis, for e = -1, 0, +1
// (index e + 1): the pull source of direction q is  c + xo[ex+1] + yo[ey+1] + zo[ez+1].  Without wrap
// these are -e, -e*px, -e*sz (the neighbour may be a ghost cell filled by the ghost kernels); with wrap
// the first / last cell of a periodic direction reads the opposite end of the box instead, which is what
// the reference's FillBoundary would have copied into that ghost cell (LBM.cpp:603, 805-806).
struct PullOffsets {
    long long xo[3], yo[3], zo[3];
};
__device__ __forceinline__ PullOffsets pull_offsets(const Layout& L, int i, int j, int k)
{
    PullOffsets o;
    o.xo[1] = o.yo[1] = o.zo[1] = 0;
    o.xo[2] = (L.wrap[0] && i == 0) ? (long long)(L.nx - 1) : -1LL;           // e = +1: source i - 1
    o.xo[0] = (L.wrap[0] && i == L.nx - 1) ? -(long long)(L.nx - 1) : 1LL;    // e = -1: source i + 1
    o.yo[2] = ((L.wrap[1] && j == 0) ? (long long)(L.ny - 1) : -1LL) * L.px;
    o.yo[0] = ((L.wrap[1] && j == L.ny - 1) ? -(long long)(L.ny - 1) : 1LL) * L.px;
    o.zo[2] = ((L.wrap[2] && k == 0) ? (long long)(L.nz - 1) : -1LL) * L.sz;
    o.zo[0] = ((L.wrap[2] && k == L.nz - 1) ? -(long long)(L.nz - 1) : 1LL) * L.sz;
    return o;
}

struct Phys {
    double nu, alpha, R, cv, gamma, dt, mesh_speed;
    double idx[3];  // geom.InvCellSizeArray()
};

// flag byte (the "EB flag byte-field"): bit0 fluid, bit1 eb_boundary,
// bits 2..7: neighbour usable by gradient() in +x,-x,+y,-y,+z,-z
constexpr unsigned FLAG_FLUID = 1u, FLAG_EBB = 2u;
constexpr unsigned GRAD_PX = 4u, GRAD_MX = 8u, GRAD_PY = 16u, GRAD_MY = 32u, GRAD_PZ = 64u, GRAD_MZ = 128u;
constexpr uint32_t ALL_FLUID = 0x7FFFFFFu;  // pull mask: bit q = cell x - e_q is fluid

// ---------------------------------------------------------------------------
// moments (Source/LBM.cpp:841-886): lattice velocities are compile-time, so the
// products e*f fold into signed adds
// ---------------------------------------------------------------------------
struct MomF {
    double rho, jx, jy, jz, pxx, pyy, pzz, pxy, pxz, pyz;
};
struct MomG {
    double e2, qx, qy, qz;
};

template <int Q>
__device__ __forceinline__ void acc_f(MomF& m, double v)
{
    m.rho += v;
    if (ex(Q) == 1) m.jx += v;
    if (ex(Q) == -1) m.jx -= v;
    if (ey(Q) == 1) m.jy += v;
    if (ey(Q) == -1) m.jy -= v;
    if (ez(Q) == 1) m.jz += v;
    if (ez(Q) == -1) m.jz -= v;
    if (ex(Q) != 0) m.pxx += v;
    if (ey(Q) != 0) m.pyy += v;
    if (ez(Q) != 0) m.pzz += v;
    if (ex(Q) * ey(Q) == 1) m.pxy += v;
    if (ex(Q) * ey(Q) == -1) m.pxy -= v;
    if (ex(Q) * ez(Q) == 1) m.pxz += v;
    if (ex(Q) * ez(Q) == -1) m.pxz -= v;
    if (ey(Q) * ez(Q) == 1) m.pyz += v;
    if (ey(Q) * ez(Q) == -1) m.pyz -= v;
}
template <int Q>
__device__ __forceinline__ void acc_g(MomG& m, double v)
{
    m.e2 += v;
    if (ex(Q) == 1) m.qx += v;
    if (ex(Q) == -1) m.qx -= v;
    if (ey(Q) == 1) m.qy += v;
    if (ey(Q) == -1) m.qy -= v;
    if (ez(Q) == 1) m.qz += v;
    if (ez(Q) == -1) m.qz -= v;
}
// light version for the q-correction pass: rho and momentum only
struct MomL {
    double rho, jx, jy, jz;
};
template <int Q>
__device__ __forceinline__ void acc_l(MomL& m, double v)
{
    m.rho += v;
    if (ex(Q) == 1) m.jx += v;
    if (ex(Q) == -1) m.jx -= v;
    if (ey(Q) == 1) m.jy += v;
    if (ey(Q) == -1) m.jy -= v;
    if (ez(Q) == 1) m.jz += v;
    if (ez(Q) == -1) m.jz -= v;
}

// ---------------------------------------------------------------------------
// The same moments by separable reduction over the tensor-product lattice (x, then y, then z): 66 adds for the
// ten f moments instead of ~170 and short dependency chains; `get(q)` returns population q.  Used by the
// collide kernels, which hold all 27 values anyway (k_qcorr accumulates on the fly with acc_l instead, to stay
// at 40 registers).  Sums in another order than acc_f / acc_g: round-off level differences.
// ---------------------------------------------------------------------------
template <typename G>
__device__ __forceinline__ MomF moments_f(G&& get)
{
    double A[3], Ay[3], Ayy[3], D[3], Dy[3], S[3];
#pragma unroll
    for (int z = 0; z < 3; ++z) {
        double a[3], d[3], sx[3];
#pragma unroll
        for (int y = 0; y < 3; ++y) {
            const double fm = get(find_dir(-1, y - 1, z - 1)), f0 = get(find_dir(0, y - 1, z - 1)),
                         fp = get(find_dir(1, y - 1, z - 1));
            sx[y] = fp + fm;
            d[y] = fp - fm;
            a[y] = sx[y] + f0;
        }
        const double t = a[2] + a[0];
        A[z] = t + a[1];
        Ay[z] = a[2] - a[0];
        Ayy[z] = t;
        D[z] = (d[2] + d[0]) + d[1];
        Dy[z] = d[2] - d[0];
        S[z] = (sx[2] + sx[0]) + sx[1];
    }
    MomF m;
    const double t = A[2] + A[0];
    m.rho = t + A[1];
    m.jz = A[2] - A[0];
    m.pzz = t;
    m.jy = (Ay[2] + Ay[0]) + Ay[1];
    m.pyz = Ay[2] - Ay[0];
    m.pyy = (Ayy[2] + Ayy[0]) + Ayy[1];
    m.jx = (D[2] + D[0]) + D[1];
    m.pxz = D[2] - D[0];
    m.pxy = (Dy[2] + Dy[0]) + Dy[1];
    m.pxx = (S[2] + S[0]) + S[1];
    return m;
}
template <typename G>
__device__ __forceinline__ MomG moments_g(G&& get)
{
    double A[3], Ay[3], D[3];
#pragma unroll
    for (int z = 0; z < 3; ++z) {
        double a[3], d[3];
#pragma unroll
        for (int y = 0; y < 3; ++y) {
            const double gm = get(find_dir(-1, y - 1, z - 1)), g0 = get(find_dir(0, y - 1, z - 1)),
                         gp = get(find_dir(1, y - 1, z - 1));
            d[y] = gp - gm;
            a[y] = (gp + gm) + g0;
        }
        A[z] = (a[2] + a[0]) + a[1];
        Ay[z] = a[2] - a[0];
        D[z] = (d[2] + d[0]) + d[1];
    }
    MomG m;
    m.e2 = (A[2] + A[0]) + A[1];
    m.qz = A[2] - A[0];
    m.qy = (Ay[2] + Ay[0]) + Ay[1];
    m.qx = (D[2] + D[0]) + D[1];
    return m;
}

// compile-time loop
template <int Q, int N, typename F>
__device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (Q < N) {
        f(std::integral_constant<int, Q>{});
        static_for<Q + 1, N>(f);
    }
}

// primitive state from the conserved moments (LBM.cpp:863-901, Utilities.H:187-196)
struct Prim {
    double rho, inv_rho, u, v, w, T, RT;
    double qcx, qcy, qcz;  // QCorr_a = rho u_a ((1 - 3RT) - u_a^2)
};
__device__ __forceinline__ Prim primitives(double rho, double jx, double jy, double jz, double e2, const Phys& P)
{
    Prim s;
    s.rho = rho;
    s.inv_rho = 1.0 / rho;
    const double c = P.mesh_speed * s.inv_rho;
    s.u = jx * c;
    s.v = jy * c;
    s.w = jz * c;
    s.T = (0.5 / P.cv) * (e2 * s.inv_rho - (s.u * s.u + s.v * s.v + s.w * s.w));
    s.RT = P.R * s.T;
    const double a = 1.0 - 3.0 * s.RT;
    s.qcx = rho * s.u * (a - s.u * s.u);
    s.qcy = rho * s.v * (a - s.v * s.v);
    s.qcz = rho * s.w * (a - s.w * s.w);
    return s;
}

// gradient() of Source/Utilities.H:279-312 for one direction, given the usable bits
__device__ __forceinline__ double one_sided_gradient(bool okp, bool okm, double dp, double d0, double dm, double idx)
{
    double vp = 0.0, vc = 0.0, vm = 0.0;
    if (okp && okm) {
        vp = 0.5 * dp;
        vm = -0.5 * dm;
    } else if (okm) {
        vc = d0;
        vm = -dm;
    } else if (okp) {
        vp = dp;
        vc = -d0;
    }
    return (vp + vc + vm) * idx;
}

// ---------------------------------------------------------------------------
// collision coefficients: everything that does not depend on the direction
// (LBM.cpp:675-758, Utilities.H:42-63, 65-163, 245-277).
// ---------------------------------------------------------------------------
struct Coll {
    double omega;
    double rho;
    double phx[3], phy[3], phz[3];  // phi_a(e) for e = -1, 0, +1 (index e+1)
    double g0;                       // E2 - 0.5 theta0 (a2xx + a2yy + a2zz)
    double a1x, a1y, a1z;            // q*_a / theta0
    double hxx, hyy, hzz;            // 0.5 a2_aa
    double axy, axz, ayz;            // a2_ab (already times 2 * 0.5)
};

__device__ __forceinline__ Coll collision_coefficients(const Prim& s, const MomF& mf, const MomG& mg, double dqx,
                                                       double dqy, double dqz, const Phys& P)
{
    Coll c;
    const double rho = s.rho, u = s.u, v = s.v, w = s.w;
    const double x = 1.0 / (s.RT * P.dt);
    const double omega = 1.0 / (P.nu * x + 0.5);          // LBM.cpp:675-677
    const double omega_one = 1.0 / (P.alpha * x + 0.5);   // LBM.cpp:678-680
    const double r = omega_one / omega;                    // LBM.cpp:681
    const double omega_corr = (2.0 - omega) / (2.0 * omega * rho);  // LBM.cpp:682-683
    c.omega = omega;
    c.rho = rho;
    const double k = P.dt * omega_corr;
    const double pix = u * u + s.RT + k * dqx;  // LBM.cpp:685-694
    const double piy = v * v + s.RT + k * dqy;
    const double piz = w * w + s.RT + k * dqz;
    // phi(e) = e u/2 + |e| (1.5 Pi - 1) - Pi + 1   (Utilities.H:53-59)
    c.phx[1] = 1.0 - pix;
    c.phx[2] = 0.5 * u + (1.5 * pix - 1.0) - pix + 1.0;
    c.phx[0] = -0.5 * u + (1.5 * pix - 1.0) - pix + 1.0;
    c.phy[1] = 1.0 - piy;
    c.phy[2] = 0.5 * v + (1.5 * piy - 1.0) - piy + 1.0;
    c.phy[0] = -0.5 * v + (1.5 * piy - 1.0) - piy + 1.0;
    c.phz[1] = 1.0 - piz;
    c.phz[2] = 0.5 * w + (1.5 * piz - 1.0) - piz + 1.0;
    c.phz[0] = -0.5 * w + (1.5 * piz - 1.0) - piz + 1.0;

    // energy lattice (Utilities.H:245-277); the temperature here is the RealVect overload
    // of get_temperature whose macro expansion subtracts only u^2 and ADDS v^2, w^2
    // (Utilities.H:204-206) -- reproduced on purpose
    const double e2 = mg.e2;
    const double Tq = (0.5 / P.cv) * (e2 * s.inv_rho - u * u + v * v + w * w);
    const double p = rho * P.R * Tq;
    const double p_rho = p * s.inv_rho;
    const double h = e2 * (0.5 * s.inv_rho) + p_rho;
    const double H = h + p_rho;
    const double two_rho = 2.0 * rho;
    const double qex = two_rho * u * h, qey = two_rho * v * h, qez = two_rho * w * h;
    const double two_ph = 2.0 * p * h;
    const double rH = two_rho * H;
    const double rxx = rH * u * u + two_ph, ryy = rH * v * v + two_ph, rzz = rH * w * w + two_ph;
    const double rxy = rH * u * v, rxz = rH * u * w, ryz = rH * v * w;
    // MRT heat flux (LBM.cpp:727-748)
    const double omr = 1.0 - r;
    const double qsx = r * qex + omr * (mg.qx - 2.0 * u * mf.pxx - 2.0 * v * mf.pxy - 2.0 * w * mf.pxz - u * P.dt * dqx);
    const double qsy = r * qey + omr * (mg.qy - 2.0 * u * mf.pxy - 2.0 * v * mf.pyy - 2.0 * w * mf.pyz - v * P.dt * dqy);
    const double qsz = r * qez + omr * (mg.qz - 2.0 * u * mf.pxz - 2.0 * v * mf.pyz - 2.0 * w * mf.pzz - w * P.dt * dqz);
    // Grad expansion with frame velocity 0, s = 1 (Utilities.H:109-160)
    constexpr double it = 1.0 / THETA0;
    c.a1x = qsx * it;
    c.a1y = qsy * it;
    c.a1z = qsz * it;
    const double a2xx = (rxx - e2 * THETA0) * it * it;
    const double a2yy = (ryy - e2 * THETA0) * it * it;
    const double a2zz = (rzz - e2 * THETA0) * it * it;
    c.hxx = 0.5 * a2xx;
    c.hyy = 0.5 * a2yy;
    c.hzz = 0.5 * a2zz;
    c.axy = rxy * it * it;
    c.axz = rxz * it * it;
    c.ayz = ryz * it * it;
    c.g0 = e2 - THETA0 * (c.hxx + c.hyy + c.hzz);
    return c;
}

template <int Q>
__device__ __forceinline__ double feq_q(const Coll& c)
{
    return c.rho * c.phx[ex(Q) + 1] * c.phy[ey(Q) + 1] * c.phz[ez(Q) + 1];
}
template <int Q>
__device__ __forceinline__ double geq_q(const Coll& c)
{
    double v = c.g0;
    if (ex(Q) == 1) v += c.a1x;
    if (ex(Q) == -1) v -= c.a1x;
    if (ey(Q) == 1) v += c.a1y;
    if (ey(Q) == -1) v -= c.a1y;
    if (ez(Q) == 1) v += c.a1z;
    if (ez(Q) == -1) v -= c.a1z;
    if (ex(Q) != 0) v += c.hxx;
    if (ey(Q) != 0) v += c.hyy;
    if (ez(Q) != 0) v += c.hzz;
    if (ex(Q) * ey(Q) == 1) v += c.axy;
    if (ex(Q) * ey(Q) == -1) v -= c.axy;
    if (ex(Q) * ez(Q) == 1) v += c.axz;
    if (ex(Q) * ez(Q) == -1) v -= c.axz;
    if (ey(Q) * ez(Q) == 1) v += c.ayz;
    if (ey(Q) * ez(Q) == -1) v -= c.ayz;
    return weight(Q) * v;
}

// ---------------------------------------------------------------------------
// run-time-direction versions for the (tiny) boundary / initial-condition kernels
// ---------------------------------------------------------------------------
struct DirTables {
    int ex[NQ], ey[NQ], ez[NQ], opp[NQ], mx[NQ], my[NQ], mz[NQ];
    double w[NQ];
};
static __constant__ DirTables c_dir;  // one copy per translation unit (only kernels.cu reads it)

// set_equilibrium_value, Utilities.H:11-40
__device__ __forceinline__ double feq_std(double rho, const double vel[3], double rt, int q)
{
    double phi[3];
    const int e[3] = {c_dir.ex[q], c_dir.ey[q], c_dir.ez[q]};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double pi = vel[d] * vel[d] + rt;
        phi[d] = e[d] * 0.5 * vel[d] + (e[d] != 0 ? 1.0 : 0.0) * (1.5 * pi - 1.0) - pi + 1.0;
    }
    return rho * phi[0] * phi[1] * phi[2];
}

// g equilibrium from a primitive state: get_energy + get_equilibrium_moments +
// set_extended_grad_expansion_generic (IC.H:493-518, BC.H:169-188, 298-320)
__device__ __forceinline__ double geq_state(double rho, const double vel[3], double T, double R, double gamma, int q)
{
    const double cv = R / (gamma - 1.0);
    const double u = vel[0], v = vel[1], w = vel[2];
    const double e2 = rho * (2.0 * cv * T + u * u + v * v + w * w);
    const double Tq = (0.5 / cv) * (e2 / rho - u * u + v * v + w * w);
    const double p = rho * R * Tq;
    const double h = e2 / (2.0 * rho) + p / rho;
    const double H = h + p / rho;
    const double qe[3] = {2.0 * rho * u * h, 2.0 * rho * v * h, 2.0 * rho * w * h};
    const double rxx = 2.0 * rho * u * u * H + 2.0 * p * h;
    const double ryy = 2.0 * rho * v * v * H + 2.0 * p * h;
    const double rzz = 2.0 * rho * w * w * H + 2.0 * p * h;
    const double rxy = 2.0 * rho * u * v * H, rxz = 2.0 * rho * u * w * H, ryz = 2.0 * rho * v * w * H;
    constexpr double it = 1.0 / THETA0;
    const int e0 = c_dir.ex[q], e1 = c_dir.ey[q], e2i = c_dir.ez[q];
    double f = e2 + qe[0] * it * e0 + qe[1] * it * e1 + qe[2] * it * e2i;
    f += 0.5 * ((e0 * e0 - THETA0) * ((rxx - e2 * THETA0) * it * it) + (e1 * e1 - THETA0) * ((ryy - e2 * THETA0) * it * it) +
                2.0 * (e0 * e1) * (rxy * it * it) + (e2i * e2i - THETA0) * ((rzz - e2 * THETA0) * it * it) +
                2.0 * (e0 * e2i) * (rxz * it * it) + 2.0 * (e1 * e2i) * (ryz * it * it));
    return f * c_dir.w[q];
}

}  // namespace mbl
