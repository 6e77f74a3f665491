/*
 * marbles_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see marbles_oracle.h).
 *
 * Restates, function by function, the reference's D3Q27 f+g lattice update.
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).  The floating-point expressions keep the reference's
 * association order so the result is comparable bit for bit with the reference
 * executable (no -ffast-math, no FMA contraction: build with -ffp-contract=off).
 */
#include "marbles_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* index helpers (FAB layout: AMReX_Array4.H:60-94)                    */
/* ------------------------------------------------------------------ */
typedef struct {
    int lo[3];  /* lower corner of the grown box */
    long s[3];  /* strides: 1, nx, nx*ny          */
    long nc;    /* cells per component            */
    int n[3];
} orc_box;

static orc_box grown(const orc_params* p, int ng)
{
    orc_box b;
    for (int d = 0; d < 3; ++d) {
        b.lo[d] = p->lo[d] - ng;
        b.n[d] = p->hi[d] - p->lo[d] + 1 + 2 * ng;
    }
    b.s[0] = 1;
    b.s[1] = b.n[0];
    b.s[2] = (long)b.n[0] * b.n[1];
    b.nc = b.s[2] * b.n[2];
    return b;
}

long orc_ncell_grown(const orc_params* p, int ng) { return grown(p, ng).nc; }

#define IDX(b, i, j, k) (((long)((i) - (b).lo[0])) + ((long)((j) - (b).lo[1])) * (b).s[1] + ((long)((k) - (b).lo[2])) * (b).s[2])
#define AT(a, b, i, j, k, c) ((a)[IDX(b, i, j, k) + (long)(c) * (b).nc])

static int in_box(const int lo[3], const int hi[3], int i, int j, int k)
{
    return i >= lo[0] && i <= hi[0] && j >= lo[1] && j <= hi[1] && k >= lo[2] && k <= hi[2];
}

static int in_grown(const orc_box* b, int i, int j, int k)
{
    return i >= b->lo[0] && i < b->lo[0] + b->n[0] && j >= b->lo[1] && j < b->lo[1] + b->n[1] &&
           k >= b->lo[2] && k < b->lo[2] + b->n[2];
}

/* ------------------------------------------------------------------ */
/* D3Q27 tables -- Source/Stencil.H:49-167                             */
/* ------------------------------------------------------------------ */
static int EV[27][3];
static double WT[27];
static int BOUNCE[27], BOUNCE_X[27], BOUNCE_Y[27], BOUNCE_Z[27];
static int tables_ready = 0;

static int find_dir(int x, int y, int z)
{
    for (int q = 0; q < 27; ++q)
        if (EV[q][0] == x && EV[q][1] == y && EV[q][2] == z) return q;
    return -1;
}

static void build_tables(void)
{
    if (tables_ready) return;
    /* velocity ordering of Stencil.H:49-84: rest, 6 axis, 12 face-diagonal, 8 body-diagonal;
       inside each group pairs (e, -e) are adjacent */
    static const int ev[27][3] = {
        {0, 0, 0},  {1, 0, 0},   {-1, 0, 0},  {0, 1, 0},   {0, -1, 0},  {0, 0, 1},  {0, 0, -1},
        {1, 1, 0},  {-1, -1, 0}, {1, -1, 0},  {-1, 1, 0},  {1, 0, 1},   {-1, 0, -1}, {1, 0, -1},
        {-1, 0, 1}, {0, 1, 1},   {0, -1, -1}, {0, 1, -1},  {0, -1, 1},  {1, 1, 1},  {-1, -1, -1},
        {1, -1, 1}, {-1, 1, -1}, {1, -1, -1}, {-1, 1, 1},  {1, 1, -1},  {-1, -1, 1}};
    memcpy(EV, ev, sizeof(ev));
    const double wsc = 2.0 / 27.0, wfcc = 1.0 / 54.0, wbcc = 1.0 / 216.0;
    const double w0 = 1.0 - (6.0 * (wsc) + 12.0 * (wfcc) + 8.0 * (wbcc)); /* Stencil.H:84-85 */
    for (int q = 0; q < 27; ++q) {
        int s = abs(EV[q][0]) + abs(EV[q][1]) + abs(EV[q][2]);
        WT[q] = (s == 0) ? w0 : (s == 1) ? wsc : (s == 2) ? wfcc : wbcc;
        BOUNCE[q] = find_dir(-EV[q][0], -EV[q][1], -EV[q][2]); /* Stencil.H:126-134 */
        BOUNCE_X[q] = find_dir(-EV[q][0], EV[q][1], EV[q][2]); /* Stencil.H:136-145 */
        BOUNCE_Y[q] = find_dir(EV[q][0], -EV[q][1], EV[q][2]); /* Stencil.H:147-156 */
        BOUNCE_Z[q] = find_dir(EV[q][0], EV[q][1], -EV[q][2]); /* Stencil.H:158-167 */
    }
    tables_ready = 1;
}

void orc_stencil(int evs[27][3], double w[27], int bounce[27], int bx[27], int by[27], int bz[27])
{
    build_tables();
    memcpy(evs, EV, sizeof(EV));
    memcpy(w, WT, sizeof(WT));
    memcpy(bounce, BOUNCE, sizeof(BOUNCE));
    memcpy(bx, BOUNCE_X, sizeof(BOUNCE_X));
    memcpy(by, BOUNCE_Y, sizeof(BOUNCE_Y));
    memcpy(bz, BOUNCE_Z, sizeof(BOUNCE_Z));
}

/* Source/Stencil.cpp:5-62 */
int orc_check_stencil(void)
{
    build_tables();
    const double small = 2.220446049250313e-16 * 1e10; /* Constants.H:57-58 */
    for (int q = 0; q < 27; ++q) {
        const int b = BOUNCE[q];
        const int sum = abs(EV[q][0]) + abs(EV[q][1]) + abs(EV[q][2]);
        if (EV[q][0] + EV[b][0] != 0 || EV[q][1] + EV[b][1] != 0 || EV[q][2] + EV[b][2] != 0) return 1;
        const double want = sum == 3 ? 1.0 / 216.0 : sum == 2 ? 1.0 / 54.0 : sum == 1 ? 2.0 / 27.0 : 8.0 / 27.0;
        if (fabs(WT[q] - want) > small) return 2;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* device math -- Source/Utilities.H                                   */
/* ------------------------------------------------------------------ */

/* Utilities.H:53-62 (extended) and :29-39 (standard: pass pxx = u^2 + RT) */
static double feq_product(double rho, const double vel[3], double pxx, double pyy, double pzz, const int ev[3])
{
    const double phix = ev[0] * 0.5 * vel[0] + abs(ev[0]) * (1.50 * pxx - 1.0) - pxx + 1.0;
    const double phiy = ev[1] * 0.5 * vel[1] + abs(ev[1]) * (1.50 * pyy - 1.0) - pyy + 1.0;
    const double phiz = ev[2] * 0.5 * vel[2] + abs(ev[2]) * (1.50 * pzz - 1.0) - pzz + 1.0;
    return rho * phix * phiy * phiz;
}

/* Utilities.H:11-40 */
static double set_equilibrium_value(double rho, const double vel[3], double rt, const int ev[3])
{
    const double pxx = vel[0] * vel[0] + rt;
    const double pyy = vel[1] * vel[1] + rt;
    const double pzz = vel[2] * vel[2] + rt;
    return feq_product(rho, vel, pxx, pyy, pzz, ev);
}

/* Utilities.H:65-163 with frame_velocity = 0 and s = 1 kept symbolic so the
   rounding sequence is the reference's */
static double grad_expansion(double rho, const double mom[3], const double flux[6], double wt, const int ev[3],
                             double theta0)
{
    const double ux = 0.0, uy = 0.0, uz = 0.0, s = 1.0;
    const double jx = mom[0], jy = mom[1], jz = mom[2];
    const double pxx = flux[0], pyy = flux[1], pzz = flux[2], pxy = flux[3], pxz = flux[4], pyz = flux[5];
    const double stheta0 = s * theta0;
    const double ib = 1.0 / stheta0;
    const double a1x = ((jx)-rho * ux) * ib;
    const double a1y = ((jy)-rho * uy) * ib;
    const double a1z = ((jz)-rho * uz) * ib;
    const double a2xx = ((pxx)-rho * s * stheta0 - rho * ux * ux - ux * ((jx)-rho * ux) - ux * ((jx)-rho * ux)) * ib * ib;
    const double a2yy = ((pyy)-rho * s * stheta0 - rho * uy * uy - uy * ((jy)-rho * uy) - uy * ((jy)-rho * uy)) * ib * ib;
    const double a2zz = ((pzz)-rho * s * stheta0 - rho * uz * uz - uz * ((jz)-rho * uz) - uz * ((jz)-rho * uz)) * ib * ib;
    const double a2xy = ((pxy)-0 - rho * ux * uy - ux * ((jy)-rho * uy) - uy * ((jx)-rho * ux)) * ib * ib;
    const double a2xz = ((pxz)-0 - rho * ux * uz - ux * ((jz)-rho * uz) - uz * ((jx)-rho * ux)) * ib * ib;
    const double a2yz = ((pyz)-0 - rho * uy * uz - uy * ((jz)-rho * uz) - uz * ((jy)-rho * uy)) * ib * ib;

    double f = rho + a1x * ev[0] + a1y * ev[1] + a1z * ev[2];
    f += 0.5 * ((ev[0] * ev[0] - theta0) * a2xx + (ev[1] * ev[1] - theta0) * a2yy + 2.0 * (ev[0] * ev[1] - 0) * a2xy +
                (ev[2] * ev[2] - theta0) * a2zz + 2.0 * (ev[0] * ev[2] - 0) * a2xz + 2.0 * (ev[1] * ev[2] - 0) * a2yz);
    f *= wt;
    return f;
}

/* Utilities.H:176-185 */
static double get_energy(double T, double rho, const double vel[3], double cv)
{
    /* RealVect overload: the dimension macro expands without parentheses */
    return rho * (2.0 * cv * T + vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
}

/* Utilities.H:187-196 (scalar overload, correct) */
static double get_temperature(double two_rho_e, double rho, double ux, double uy, double uz, double cv)
{
    return (0.50 / cv) * ((two_rho_e / rho) - (ux * ux + uy * uy + uz * uz));
}

/* Utilities.H:198-207 (RealVect overload: the macro expands WITHOUT parentheses,
   so only u_x^2 is subtracted and u_y^2, u_z^2 are added -- reproduced) */
static double get_temperature_vec(double two_rho_e, double rho, const double vel[3], double cv)
{
    return (0.50 / cv) * ((two_rho_e / rho) - vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
}

/* Utilities.H:245-277 */
static void get_equilibrium_moments(double rho, const double vel[3], double total_energy, double cv, double R,
                                    double qeq[3], double req[6])
{
    const double energy = total_energy / (2.0 * rho);
    const double temperature = get_temperature_vec(total_energy, rho, vel, cv);
    const double p = rho * R * temperature;
    const double h = energy + (p / rho);
    qeq[0] = 2.0 * rho * vel[0] * h;
    qeq[1] = 2.0 * rho * vel[1] * h;
    req[0] = 2.0 * rho * vel[0] * vel[0] * (h + (p / rho)) + 2.0 * p * h;
    req[1] = 2.0 * rho * vel[1] * vel[1] * (h + (p / rho)) + 2.0 * p * h;
    req[3] = 2.0 * rho * vel[0] * vel[1] * (h + (p / rho)) + 0;
    qeq[2] = 2.0 * rho * vel[2] * h;
    req[2] = 2.0 * rho * vel[2] * vel[2] * (h + (p / rho)) + 2.0 * p * h;
    req[4] = 2.0 * rho * vel[0] * vel[2] * (h + (p / rho)) + 0;
    req[5] = 2.0 * rho * vel[1] * vel[2] * (h + (p / rho)) + 0;
}

/* thermal equilibrium used by IC.H:493-518, BC.H:169-188 and BC.H:298-320 */
static double geq_from_state(double rho, const double vel[3], double T, double R, double gamma, double wt,
                             const int ev[3], double theta0)
{
    const double cv = R / (gamma - 1.0);
    const double two_rho_e = get_energy(T, rho, vel, cv);
    double q[3] = {0.0, 0.0, 0.0}, r[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    get_equilibrium_moments(rho, vel, two_rho_e, cv, R, q, r);
    return grad_expansion(two_rho_e, q, r, wt, ev, theta0);
}

/* Utilities.H:279-312 */
static double gradient(const orc_params* p, int dir, int comp, int i, int j, int k, const int* is_fluid,
                       const orc_box* fb, const double* data, const orc_box* mb)
{
    int ip[3] = {i, j, k}, im[3] = {i, j, k};
    ip[dir] += 1;
    im[dir] -= 1;
    const int bad_p = (!in_box(p->dom_lo, p->dom_hi, ip[0], ip[1], ip[2])) || (AT(is_fluid, *fb, ip[0], ip[1], ip[2], 0) != 1);
    const int bad_m = (!in_box(p->dom_lo, p->dom_hi, im[0], im[1], im[2])) || (AT(is_fluid, *fb, im[0], im[1], im[2], 0) != 1);
    double vp = 0.0, vc = 0.0, vm = 0.0;
    if (bad_p && bad_m) {
        vp = 0.0;
        vc = 0.0;
        vm = 0.0;
    } else if (bad_p) {
        vp = 0.0;
        vc = AT(data, *mb, i, j, k, comp);
        vm = -AT(data, *mb, im[0], im[1], im[2], comp);
    } else if (bad_m) {
        vp = AT(data, *mb, ip[0], ip[1], ip[2], comp);
        vc = -AT(data, *mb, i, j, k, comp);
        vm = 0.0;
    } else {
        vp = 0.5 * AT(data, *mb, ip[0], ip[1], ip[2], comp);
        vc = 0.0;
        vm = -0.5 * AT(data, *mb, im[0], im[1], im[2], comp);
    }
    return (vp + vc + vm) * p->inv_dx[dir];
}

/* ------------------------------------------------------------------ */
/* initial conditions -- Source/IC.H                                   */
/* ------------------------------------------------------------------ */
static void ic_eval(const orc_params* p, const orc_ic* ic, int i, int j, int k, double* rho, double vel[3], double* T,
                    double* R, double* gamma)
{
    const double x05 = p->prob_lo[0] + (i + 0.5) * p->dx[0];
    const double y05 = p->prob_lo[1] + (j + 0.5) * p->dx[1];
    const double z05 = p->prob_lo[2] + (k + 0.5) * p->dx[2];
    switch (ic->kind) {
    case ORC_IC_CONSTANT: /* IC.H:42-57 */
        *rho = ic->density;
        vel[0] = ic->velocity[0];
        vel[1] = ic->velocity[1];
        vel[2] = ic->velocity[2];
        *T = ic->T0;
        *R = ic->R;
        *gamma = ic->gamma;
        break;
    case ORC_IC_TAYLORGREEN: { /* IC.H:119-153; L = 1/pi (IC.H:77) */
        const double L = 1.0 / M_PI;
        const double o0 = ic->omega[0], o1 = ic->omega[1], o2 = ic->omega[2];
        const double rho0 = ic->density, v0 = ic->v0;
        *rho = rho0 + rho0 * v0 * v0 / 16.0 * (cos(2.0 * o0 * x05 / L) + cos(2.0 * o1 * y05 / L)) *
                          (cos(2.0 * o2 * z05 / L) + 2.0);
        vel[0] = v0 * sin(o0 * x05 / L) * cos(o1 * y05 / L) * cos(o2 * z05 / L);
        vel[1] = -v0 * cos(o0 * x05 / L) * sin(o1 * y05 / L) * cos(o2 * z05 / L);
        vel[2] = 0.0;
        *T = ic->T0;
        *R = ic->R;
        *gamma = 5.0 / 3.0; /* const member, IC.H:80 */
        break;
    }
    case ORC_IC_VISCOSITY: { /* IC.H:213-240; note (iv + 0.5*0.0) */
        const double y = p->prob_lo[1] + (j + 0.5 * 0.0) * p->dx[1];
        *rho = ic->density;
        const double a0 = 0.010 * ic->c_s;
        vel[0] = ic->velocity[0] + a0 * sin(2.0 * M_PI * y / ic->wave_length);
        vel[1] = ic->velocity[1];
        vel[2] = ic->velocity[2];
        *T = ic->T0;
        *R = ic->R;
        *gamma = ic->gamma;
        break;
    }
    case ORC_IC_THERMALDIFF: { /* IC.H:302-333 */
        const double y = p->prob_lo[1] + (j + 0.5 * 0.0) * p->dx[1];
        const double a0 = 0.0010 * ic->T0;
        *R = ic->R;
        const double pressure = ic->density * *R * ic->T0;
        *rho = ic->density + a0 * sin(2.0 * M_PI * y / ic->wave_length);
        vel[0] = ic->velocity[0];
        vel[1] = ic->velocity[1];
        vel[2] = ic->velocity[2];
        *gamma = ic->gamma;
        *T = pressure / (*rho * *R);
        break;
    }
    case ORC_IC_SOD: { /* IC.H:394-428 */
        const double x = p->prob_lo[0] + (i + 0.5 * 0.0) * p->dx[0];
        *R = ic->R;
        vel[0] = ic->velocity[0];
        vel[1] = ic->velocity[1];
        vel[2] = ic->velocity[2];
        *gamma = ic->gamma;
        *rho = ic->density +
               0.5 * (1.0 + tanh((x - ic->x_discontinuity) * 3.0)) * (ic->density_ratio * ic->density - ic->density);
        *T = ic->T0 + 0.5 * (1.0 + tanh((x - ic->x_discontinuity) * 3.0)) * (ic->temperature_ratio * ic->T0 - ic->T0);
        break;
    }
    default:
        *rho = 1.0;
        vel[0] = vel[1] = vel[2] = 0.0;
        *T = 1.0 / 3.0;
        *R = 1.0;
        *gamma = 1.667;
    }
}

/* IC.H:474-519 on the grown box, then LBM.cpp:1287-1295 */
void orc_initialize(const orc_params* p, const orc_ic* ic, const int* is_fluid, double* f, double* g)
{
    build_tables();
    const orc_box b = grown(p, p->ng);
    const double theta0 = 1.0 / 3.0;
    for (int k = b.lo[2]; k < b.lo[2] + b.n[2]; ++k)
        for (int j = b.lo[1]; j < b.lo[1] + b.n[1]; ++j)
            for (int i = b.lo[0]; i < b.lo[0] + b.n[0]; ++i) {
                double rho = 1.0, vel[3] = {0.0, 0.0, 0.0}, T = 1.0 / 3.0, R = 1.0, gamma = 1.667;
                ic_eval(p, ic, i, j, k, &rho, vel, &T, &R, &gamma);
                const int solid = (AT(is_fluid, b, i, j, k, 0) == 0);
                for (int q = 0; q < 27; ++q) {
                    AT(f, b, i, j, k, q) = solid ? 0.0 : set_equilibrium_value(rho, vel, R * T, EV[q]);
                    AT(g, b, i, j, k, q) = solid ? 0.0 : geq_from_state(rho, vel, T, R, gamma, WT[q], EV[q], theta0);
                }
            }
}

/* ------------------------------------------------------------------ */
/* ghost fill                                                          */
/* ------------------------------------------------------------------ */

/* FillPatchOps.H:92-108 (K6): every out-of-domain ghost gets the no-slip value
   of the in-domain cell it faces */
void orc_prepass(const orc_params* p, double* f)
{
    build_tables();
    const orc_box b = grown(p, p->ng);
    for (int q = 0; q < 27; ++q)
        for (int k = b.lo[2]; k < b.lo[2] + b.n[2]; ++k)
            for (int j = b.lo[1]; j < b.lo[1] + b.n[1]; ++j)
                for (int i = b.lo[0]; i < b.lo[0] + b.n[0]; ++i) {
                    const int ie = i + EV[q][0], je = j + EV[q][1], ke = k + EV[q][2];
                    if (!in_box(p->dom_lo, p->dom_hi, i, j, k) && in_box(p->dom_lo, p->dom_hi, ie, je, ke) &&
                        in_grown(&b, ie, je, ke)) {
                        AT(f, b, i, j, k, q) = AT(f, b, ie, je, ke, BOUNCE[q]);
                    }
                }
}

static int wrap(int i, int lo, int hi)
{
    const int n = hi - lo + 1;
    int r = (i - lo) % n;
    if (r < 0) r += n;
    return lo + r;
}

/* FabArray::FillBoundary(periodicity) restricted to one box: a ghost cell whose
   periodic image (shifts only in periodic directions) is a valid cell of THIS box
   receives that cell's value (AMReX_FabArrayCommI.H:8-169, local copies) */
#define FILL_PERIODIC_BODY(TYPE)                                                          \
    const orc_box b = grown(p, ng);                                                       \
    for (int c = 0; c < ncomp; ++c)                                                       \
        for (int k = b.lo[2]; k < b.lo[2] + b.n[2]; ++k)                                  \
            for (int j = b.lo[1]; j < b.lo[1] + b.n[1]; ++j)                              \
                for (int i = b.lo[0]; i < b.lo[0] + b.n[0]; ++i) {                        \
                    if (in_box(p->lo, p->hi, i, j, k)) continue;                          \
                    const int ii = p->periodic[0] ? wrap(i, p->dom_lo[0], p->dom_hi[0]) : i; \
                    const int jj = p->periodic[1] ? wrap(j, p->dom_lo[1], p->dom_hi[1]) : j; \
                    const int kk = p->periodic[2] ? wrap(k, p->dom_lo[2], p->dom_hi[2]) : k; \
                    if (!in_box(p->lo, p->hi, ii, jj, kk)) continue;                      \
                    AT(a, b, i, j, k, c) = AT(a, b, ii, jj, kk, c);                       \
                }

void orc_fill_periodic(const orc_params* p, double* a, int ncomp, int ng) { FILL_PERIODIC_BODY(double) }
void orc_fill_periodic_int(const orc_params* p, int* a, int ncomp, int ng) { FILL_PERIODIC_BODY(int) }

/* inlet functors -- Source/VelocityBC.H:44-190 (evaluated at the literal ghost index) */
static void vel_bc_op(const orc_params* p, int i, int j, int k, double* rho, double vel[3], double* R, double* T,
                      double* gamma)
{
    const int iv[3] = {i, j, k};
    switch (p->vbc_kind) {
    case ORC_VBC_CONSTANT: /* VelocityBC.H:60-77 */
        *rho = p->vbc_rho;
        vel[p->vbc_dir] = p->vbc_u;
        *R = p->vbc_R;
        *T = p->vbc_T;
        *gamma = p->vbc_gamma;
        break;
    case ORC_VBC_CHANNEL: { /* VelocityBC.H:106-128 */
        *rho = p->vbc_rho;
        const double c1 = iv[1] * (p->dom_hi[1] - iv[1]);
        const double c2 = iv[2] * (p->dom_hi[2] - iv[2]);
        vel[0] = 16.0 * p->vbc_u * c1 * c2 / pow((double)(p->dom_hi[1] + 1), 4);
        *R = p->vbc_R;
        *T = p->vbc_T;
        *gamma = p->vbc_gamma;
        break;
    }
    case ORC_VBC_PARABOLIC: { /* VelocityBC.H:160-190 */
        *rho = p->vbc_rho;
        const int nd = p->vbc_normal_dir;
        const double height = p->prob_hi[nd] - p->prob_lo[nd];
        const double x = p->prob_lo[nd] + (iv[nd] + 0.5) * p->dx[nd];
        vel[p->vbc_tangential_dir] = 4.0 * p->vbc_u * x * (height - x) / (height * height);
        *R = p->vbc_R;
        *T = p->vbc_T;
        *gamma = p->vbc_gamma;
        break;
    }
    default: /* NoOp, VelocityBC.H:13-31 */
        break;
    }
}

/* BCFill::operator() -- Source/BC.H:345-471 for one ghost cell */
static void bc_fill_cell(const orc_params* p, const orc_box* b, double* data, int is_energy, int i, int j, int k)
{
    const int iv[3] = {i, j, k};
    const double theta0 = 1.0 / 3.0;
    for (int idir = 0; idir < 3; ++idir) {
        for (int lohi = 0; lohi < 2; ++lohi) {
            const int ndir = lohi == 0 ? 1 : -1;
            const int bc = p->bc_type[idir + lohi * 3];
            if (!((lohi == 0 && iv[idir] < p->dom_lo[idir]) || (lohi == 1 && iv[idir] > p->dom_hi[idir]))) continue;
            for (int q = 0; q < 27; ++q) {
                const int* ev = EV[q];
                const int in = i + ev[0], jn = j + ev[1], kn = k + ev[2];
                /* `inside` = data box shrunk by ng = valid box of this FAB (BC.H:374-390) */
                if (in_box(p->lo, p->hi, in, jn, kn)) {
                    if (bc == ORC_BC_NOSLIP) { /* BC.H:81 */
                        AT(data, *b, i, j, k, q) = AT(data, *b, in, jn, kn, BOUNCE[q]);
                    } else if (bc == ORC_BC_VELOCITY) { /* BC.H:394-419 */
                        double vel[3] = {0.0, 0.0, 0.0}, R = 1.0, T = 1.0 / 3.0, gamma = 5.0 / 3.0, rho_bc = 0.0;
                        vel_bc_op(p, i, j, k, &rho_bc, vel, &R, &T, &gamma);
                        if (is_energy) {
                            AT(data, *b, i, j, k, q) = geq_from_state(rho_bc, vel, T, R, gamma, WT[q], ev, theta0);
                        } else {
                            AT(data, *b, i, j, k, q) = set_equilibrium_value(rho_bc, vel, R * T, ev);
                        }
                    } else if (bc == ORC_BC_PRESSURE) { /* BC.H:421-450 */
                        double vel0[3] = {0.0, 0.0, 0.0}, R = 1.0, T = 1.0 / 3.0, gamma = 5.0 / 3.0, rho_bc = 1.0;
                        vel_bc_op(p, i, j, k, &rho_bc, vel0, &R, &T, &gamma);
                        if (is_energy) { /* BC.H:298-320 with rho_bc_out = 1 */
                            double vel[3] = {0.0, 0.0, 0.0};
                            vel[idir] = ndir * (1.0 - (1.0) / rho_bc);
                            AT(data, *b, i, j, k, q) = geq_from_state(rho_bc, vel, T, R, gamma, WT[q], ev, theta0);
                        } else { /* BC.H:242-268 */
                            double rho_out = 0.0, rho_tan = 0.0;
                            for (int qq = 0; qq < 27; ++qq) {
                                const int bq = BOUNCE[qq];
                                const int* eo = EV[qq];
                                const int* ei = EV[bq];
                                if (ei[idir] == -ndir) {
                                    rho_out += 2.0 * AT(data, *b, in + eo[0], jn + eo[1], kn + eo[2], bq);
                                } else if (ei[idir] == 0) {
                                    rho_tan += 1.0 * AT(data, *b, in + eo[0], jn + eo[1], kn + eo[2], bq);
                                }
                            }
                            double vel[3] = {0.0, 0.0, 0.0};
                            vel[idir] = ndir * (1.0 - (rho_out + rho_tan) / rho_bc);
                            AT(data, *b, i, j, k, q) = set_equilibrium_value(rho_bc, vel, R * T, ev);
                        }
                    } else if (bc == ORC_BC_OUTFLOW) { /* BC.H:339-341 */
                        int s[3] = {i, j, k};
                        s[idir] += 1 * ndir;
                        AT(data, *b, i, j, k, q) = AT(data, *b, s[0], s[1], s[2], q);
                    } else if (bc == ORC_BC_SLIP_X) { /* BC.H:100 */
                        AT(data, *b, i, j, k, q) = AT(data, *b, in, jn, kn, BOUNCE_X[q]);
                    } else if (bc == ORC_BC_SLIP_Y) { /* BC.H:119 */
                        AT(data, *b, i, j, k, q) = AT(data, *b, in, jn, kn, BOUNCE_Y[q]);
                    } else if (bc == ORC_BC_SLIP_Z) { /* BC.H:138 */
                        AT(data, *b, i, j, k, q) = AT(data, *b, in, jn, kn, BOUNCE_Z[q]);
                    }
                } else {
                    AT(data, *b, i, j, k, q) = -1.0; /* BC.H:463-466 */
                }
            }
        }
    }
}

/* side: -1 = below gdomain, 0 = inside gdomain, +1 = above; returns the range
   [a0, a1] of cells of the grown box in dimension d on that side */
static int side_range(const orc_params* p, const orc_box* b, int d, int side, int* a0, int* a1)
{
    const int blo = b->lo[d], bhi = b->lo[d] + b->n[d] - 1;
    /* gdomain = domain grown by the box length in periodic directions
       (AMReX_PhysBCFunct.H:417-423) */
    const int glo = p->periodic[d] ? p->dom_lo[d] - b->n[d] : p->dom_lo[d];
    const int ghi = p->periodic[d] ? p->dom_hi[d] + b->n[d] : p->dom_hi[d];
    if (side < 0) {
        *a0 = blo;
        *a1 = (glo - 1 < bhi) ? glo - 1 : bhi;
    } else if (side > 0) {
        *a0 = (ghi + 1 > blo) ? ghi + 1 : blo;
        *a1 = bhi;
    } else {
        *a0 = (glo > blo) ? glo : blo;
        *a1 = (ghi < bhi) ? ghi : bhi;
    }
    return *a0 <= *a1;
}

static void bc_region(const orc_params* p, const orc_box* b, double* data, int is_energy, int sx, int sy, int sz)
{
    int i0, i1, j0, j1, k0, k1;
    if (!side_range(p, b, 0, sx, &i0, &i1)) return;
    if (!side_range(p, b, 1, sy, &j0, &j1)) return;
    if (!side_range(p, b, 2, sz, &k0, &k1)) return;
    for (int k = k0; k <= k1; ++k)
        for (int j = j0; j <= j1; ++j)
            for (int i = i0; i <= i1; ++i) bc_fill_cell(p, b, data, is_energy, i, j, k);
}

/* PhysBCFunct + GpuBndryFuncFab::ccfcdoit, CPU branch: faces, edges, corners
   (AMReX_PhysBCFunct.H:199-240, 593-678) */
void orc_physbc(const orc_params* p, double* f, int is_energy_lattice, double time)
{
    (void)time;
    build_tables();
    if (p->periodic[0] && p->periodic[1] && p->periodic[2]) return; /* PhysBCFunct.H:202 */
    const orc_box b = grown(p, p->ng);
    /* faces: xlo ylo zlo xhi yhi zhi */
    bc_region(p, &b, f, is_energy_lattice, -1, 0, 0);
    bc_region(p, &b, f, is_energy_lattice, 0, -1, 0);
    bc_region(p, &b, f, is_energy_lattice, 0, 0, -1);
    bc_region(p, &b, f, is_energy_lattice, +1, 0, 0);
    bc_region(p, &b, f, is_energy_lattice, 0, +1, 0);
    bc_region(p, &b, f, is_energy_lattice, 0, 0, +1);
    /* edges: xy (4), xz (4), yz (4); first index fastest */
    for (int s1 = -1; s1 <= 1; s1 += 2)
        for (int s0 = -1; s0 <= 1; s0 += 2) bc_region(p, &b, f, is_energy_lattice, s0, s1, 0);
    for (int s1 = -1; s1 <= 1; s1 += 2)
        for (int s0 = -1; s0 <= 1; s0 += 2) bc_region(p, &b, f, is_energy_lattice, s0, 0, s1);
    for (int s1 = -1; s1 <= 1; s1 += 2)
        for (int s0 = -1; s0 <= 1; s0 += 2) bc_region(p, &b, f, is_energy_lattice, 0, s0, s1);
    /* corners, x fastest */
    for (int s2 = -1; s2 <= 1; s2 += 2)
        for (int s1 = -1; s1 <= 1; s1 += 2)
            for (int s0 = -1; s0 <= 1; s0 += 2) bc_region(p, &b, f, is_energy_lattice, s0, s1, s2);
}

/* FillPatchOps::fillpatch, lev 0 -- FillPatchOps.H:75-132 */
void orc_fillpatch(const orc_params* p, double* f, int is_energy_lattice, double time)
{
    orc_prepass(p, f);
    orc_fill_periodic(p, f, 27, p->ng);
    orc_physbc(p, f, is_energy_lattice, time);
}

/* ------------------------------------------------------------------ */
/* stream -- Source/LBM.cpp:558-604                                    */
/* ------------------------------------------------------------------ */
void orc_stream(const orc_params* p, const int* is_fluid, double* f, int fill_boundary)
{
    build_tables();
    const orc_box b = grown(p, p->ng);
    double* fs = (double*)malloc(sizeof(double) * 27 * b.nc);
    for (long n = 0; n < 27 * b.nc; ++n) fs[n] = -1.0; /* LBM.cpp:565 */
    for (int q = 0; q < 27; ++q)
        for (int k = b.lo[2]; k < b.lo[2] + b.n[2]; ++k)
            for (int j = b.lo[1]; j < b.lo[1] + b.n[1]; ++j)
                for (int i = b.lo[0]; i < b.lo[0] + b.n[0]; ++i) {
                    if (AT(is_fluid, b, i, j, k, 0) != 1) continue;
                    const int in = i + EV[q][0], jn = j + EV[q][1], kn = k + EV[q][2];
                    if (!in_grown(&b, in, jn, kn)) continue;
                    if (AT(is_fluid, b, in, jn, kn, 0) != 0) {
                        AT(fs, b, in, jn, kn, q) = AT(f, b, i, j, k, q);
                    } else {
                        AT(fs, b, i, j, k, BOUNCE[q]) = AT(f, b, i, j, k, q);
                    }
                }
    memcpy(f, fs, sizeof(double) * 27 * b.nc); /* LBM.cpp:601 */
    free(fs);
    if (fill_boundary) orc_fill_periodic(p, f, 27, p->ng); /* LBM.cpp:603 */
}

/* ------------------------------------------------------------------ */
/* collide                                                             */
/* ------------------------------------------------------------------ */

/* LBM.cpp:810-906 */
void orc_f_to_macrodata(const orc_params* p, const int* is_fluid, const double* f, const double* g, double* macro,
                        int fill_boundary)
{
    build_tables();
    const orc_box fb = grown(p, p->ng);
    const orc_box mb = grown(p, 1);
    const double R = p->R;
    const double cv = R / (p->gamma - 1.0);
    for (int k = mb.lo[2]; k < mb.lo[2] + mb.n[2]; ++k)
        for (int j = mb.lo[1]; j < mb.lo[1] + mb.n[1]; ++j)
            for (int i = mb.lo[0]; i < mb.lo[0] + mb.n[0]; ++i) {
                if (AT(is_fluid, fb, i, j, k, 0) != 1) continue;
                double rho = 0.0, u = 0.0, v = 0.0, w = 0.0;
                double pxx = 0.0, pyy = 0.0, pzz = 0.0, pxy = 0.0, pxz = 0.0, pyz = 0.0;
                double two_rho_e = 0.0, qx = 0.0, qy = 0.0, qz = 0.0;
                for (int q = 0; q < 27; ++q) {
                    const double fq = AT(f, fb, i, j, k, q);
                    const double gq = AT(g, fb, i, j, k, q);
                    const int* ev = EV[q];
                    rho += fq;
                    u += ev[0] * fq;
                    v += ev[1] * fq;
                    w += ev[2] * fq;
                    pxx += ev[0] * ev[0] * fq;
                    pyy += ev[1] * ev[1] * fq;
                    pxy += ev[0] * ev[1] * fq;
                    pzz += ev[2] * ev[2] * fq;
                    pxz += ev[0] * ev[2] * fq;
                    pyz += ev[1] * ev[2] * fq;
                    two_rho_e += gq;
                    qx += ev[0] * gq;
                    qy += ev[1] * gq;
                    qz += ev[2] * gq;
                }
                u *= p->mesh_speed / rho;
                v *= p->mesh_speed / rho;
                w *= p->mesh_speed / rho;
                AT(macro, mb, i, j, k, 0) = rho;
                AT(macro, mb, i, j, k, 1) = u;
                AT(macro, mb, i, j, k, 2) = v;
                AT(macro, mb, i, j, k, 3) = w;
                AT(macro, mb, i, j, k, 4) = sqrt(u * u + v * v + w * w);
                AT(macro, mb, i, j, k, 9) = pxx;
                AT(macro, mb, i, j, k, 10) = pyy;
                AT(macro, mb, i, j, k, 11) = pzz;
                AT(macro, mb, i, j, k, 12) = pxy;
                AT(macro, mb, i, j, k, 13) = pxz;
                AT(macro, mb, i, j, k, 14) = pyz;
                AT(macro, mb, i, j, k, 5) = two_rho_e;
                AT(macro, mb, i, j, k, 15) = qx;
                AT(macro, mb, i, j, k, 16) = qy;
                AT(macro, mb, i, j, k, 17) = qz;
                const double T = get_temperature(two_rho_e, rho, u, v, w, cv);
                AT(macro, mb, i, j, k, 18) = T;
                AT(macro, mb, i, j, k, 6) = rho * u * ((1.0 - 3.0 * R * T) - u * u);
                AT(macro, mb, i, j, k, 7) = rho * v * ((1.0 - 3.0 * R * T) - v * v);
                AT(macro, mb, i, j, k, 8) = rho * w * ((1.0 - 3.0 * R * T) - w * w);
            }
    if (fill_boundary) orc_fill_periodic(p, macro, ORC_NMACRO, 1); /* LBM.cpp:905 */
}

/* LBM.cpp:959-991 */
void orc_compute_q_corrections(const orc_params* p, const int* is_fluid, const double* macro, double* derived)
{
    const orc_box fb = grown(p, p->ng);
    const orc_box mb = grown(p, 1);
    const orc_box db = grown(p, 0);
    for (int k = p->lo[2]; k <= p->hi[2]; ++k)
        for (int j = p->lo[1]; j <= p->hi[1]; ++j)
            for (int i = p->lo[0]; i <= p->hi[0]; ++i) {
                if (AT(is_fluid, fb, i, j, k, 0) != 1) continue;
                AT(derived, db, i, j, k, 4) = gradient(p, 0, 6, i, j, k, is_fluid, &fb, macro, &mb);
                AT(derived, db, i, j, k, 5) = gradient(p, 1, 7, i, j, k, is_fluid, &fb, macro, &mb);
                AT(derived, db, i, j, k, 6) = gradient(p, 2, 8, i, j, k, is_fluid, &fb, macro, &mb);
            }
}

/* LBM.cpp:909-955 */
void orc_compute_derived(const orc_params* p, const int* is_fluid, const double* macro, double* derived)
{
    const orc_box fb = grown(p, p->ng);
    const orc_box mb = grown(p, 1);
    const orc_box db = grown(p, 0);
    for (int k = p->lo[2]; k <= p->hi[2]; ++k)
        for (int j = p->lo[1]; j <= p->hi[1]; ++j)
            for (int i = p->lo[0]; i <= p->hi[0]; ++i) {
                if (AT(is_fluid, fb, i, j, k, 0) != 1) continue;
                const double vx = gradient(p, 0, 2, i, j, k, is_fluid, &fb, macro, &mb);
                const double wx = gradient(p, 0, 3, i, j, k, is_fluid, &fb, macro, &mb);
                const double uy = gradient(p, 1, 1, i, j, k, is_fluid, &fb, macro, &mb);
                const double wy = gradient(p, 1, 3, i, j, k, is_fluid, &fb, macro, &mb);
                const double uz = gradient(p, 2, 1, i, j, k, is_fluid, &fb, macro, &mb);
                const double vz = gradient(p, 2, 2, i, j, k, is_fluid, &fb, macro, &mb);
                AT(derived, db, i, j, k, 0) = wy - vz;
                AT(derived, db, i, j, k, 1) = uz - wx;
                AT(derived, db, i, j, k, 2) = vx - uy;
                AT(derived, db, i, j, k, 3) =
                    sqrt((wy - vz) * (wy - vz) + (uz - wx) * (uz - wx) + (vx - uy) * (vx - uy));
            }
}

/* LBM.cpp:621-762 */
void orc_macrodata_to_equilibrium(const orc_params* p, const int* is_fluid, const double* macro,
                                  const double* derived, double* eq, double* eq_g)
{
    build_tables();
    const orc_box fb = grown(p, p->ng);
    const orc_box mb = grown(p, 1);
    const orc_box db = grown(p, 0);
    const double R = p->R;
    const double cv = R / (p->gamma - 1.0);
    const double nu = p->nu, alpha = p->alpha, dt = p->dt;
    const double theta0 = 1.0 / 3.0;
    for (int q = 0; q < 27; ++q)
        for (int k = p->lo[2]; k <= p->hi[2]; ++k)
            for (int j = p->lo[1]; j <= p->hi[1]; ++j)
                for (int i = p->lo[0]; i <= p->hi[0]; ++i) {
                    if (AT(is_fluid, fb, i, j, k, 0) != 1) continue;
                    const double rho = AT(macro, mb, i, j, k, 0);
                    const double vel[3] = {AT(macro, mb, i, j, k, 1), AT(macro, mb, i, j, k, 2),
                                           AT(macro, mb, i, j, k, 3)};
                    const double two_rho_e = AT(macro, mb, i, j, k, 5);
                    const double wt = WT[q];
                    const int* ev = EV[q];
                    const double T = AT(macro, mb, i, j, k, 18);
                    const double omega = 1.0 / (nu / (R * T * dt) + 0.5);
                    const double omega_one = 1.0 / (alpha / (R * T * dt) + 0.5);
                    const double omega_one_by_omega = omega_one / omega;
                    const double omega_corr = (2.0 - omega) / (2.0 * omega * rho);
                    const double dqx = AT(derived, db, i, j, k, 4);
                    const double dqy = AT(derived, db, i, j, k, 5);
                    const double dqz = AT(derived, db, i, j, k, 6);
                    const double pxx_ext = vel[0] * vel[0] + R * T + dt * (omega_corr)*dqx;
                    const double pyy_ext = vel[1] * vel[1] + R * T + dt * (omega_corr)*dqy;
                    const double pzz_ext = vel[2] * vel[2] + R * T + dt * (omega_corr)*dqz;
                    AT(eq, db, i, j, k, q) = feq_product(rho, vel, pxx_ext, pyy_ext, pzz_ext, ev);

                    double heat_flux[3] = {0.0, 0.0, 0.0};
                    double r[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                    get_equilibrium_moments(rho, vel, two_rho_e, cv, R, heat_flux, r);
                    double qx_eq = heat_flux[0], qy_eq = heat_flux[1], qz_eq = heat_flux[2];
                    const double pxx = AT(macro, mb, i, j, k, 9);
                    const double pyy = AT(macro, mb, i, j, k, 10);
                    const double pzz = AT(macro, mb, i, j, k, 11);
                    const double pxy = AT(macro, mb, i, j, k, 12);
                    const double pxz = AT(macro, mb, i, j, k, 13);
                    const double pyz = AT(macro, mb, i, j, k, 14);
                    const double qx = AT(macro, mb, i, j, k, 15);
                    const double qy = AT(macro, mb, i, j, k, 16);
                    const double qz = AT(macro, mb, i, j, k, 17);
                    qx_eq *= omega_one_by_omega;
                    qy_eq *= omega_one_by_omega;
                    qz_eq *= omega_one_by_omega;
                    qx_eq += (1.0 - omega_one_by_omega) *
                             (qx - 2.0 * vel[0] * pxx - 2.0 * vel[1] * pxy - 2.0 * vel[2] * pxz - vel[0] * dt * dqx);
                    qy_eq += (1.0 - omega_one_by_omega) *
                             (qy - 2.0 * vel[0] * pxy - 2.0 * vel[1] * pyy - 2.0 * vel[2] * pyz - vel[1] * dt * dqy);
                    qz_eq += (1.0 - omega_one_by_omega) *
                             (qz - 2.0 * vel[0] * pxz - 2.0 * vel[1] * pyz - 2.0 * vel[2] * pzz - vel[2] * dt * dqz);
                    const double mrt[3] = {qx_eq, qy_eq, qz_eq};
                    AT(eq_g, db, i, j, k, q) = grad_expansion(two_rho_e, mrt, r, wt, ev, theta0);
                }
}

/* LBM.cpp:765-807 */
void orc_relax(const orc_params* p, const int* is_fluid, const double* macro, const double* eq, const double* eq_g,
               double* f, double* g, int fill_boundary)
{
    const orc_box fb = grown(p, p->ng);
    const orc_box mb = grown(p, 1);
    const orc_box db = grown(p, 0);
    for (int q = 0; q < 27; ++q)
        for (int k = p->lo[2]; k <= p->hi[2]; ++k)
            for (int j = p->lo[1]; j <= p->hi[1]; ++j)
                for (int i = p->lo[0]; i <= p->hi[0]; ++i) {
                    if (AT(is_fluid, fb, i, j, k, 0) != 1) continue;
                    const double T = AT(macro, mb, i, j, k, 18);
                    const double omega = 1.0 / (p->nu / (p->R * T * p->dt) + 0.5);
                    AT(f, fb, i, j, k, q) += omega * (AT(eq, db, i, j, k, q) - AT(f, fb, i, j, k, q));
                    AT(g, fb, i, j, k, q) += omega * (AT(eq_g, db, i, j, k, q) - AT(g, fb, i, j, k, q));
                }
    if (fill_boundary) {
        orc_fill_periodic(p, f, 27, p->ng); /* LBM.cpp:805-806 */
        orc_fill_periodic(p, g, 27, p->ng);
    }
}

/* LBM.cpp:607-618 */
void orc_collide(const orc_params* p, const int* is_fluid, double* f, double* g, double* macro, double* derived,
                 double* eq, double* eq_g, int fill_boundary)
{
    orc_f_to_macrodata(p, is_fluid, f, g, macro, fill_boundary);
    orc_compute_q_corrections(p, is_fluid, macro, derived);
    orc_macrodata_to_equilibrium(p, is_fluid, macro, derived, eq, eq_g);
    orc_relax(p, is_fluid, macro, eq, eq_g, f, g, fill_boundary);
}

/* LBM.cpp:1236-1261: comp 1 = solid cell with at least one fluid face neighbour,
   on the box grown by ng-1 */
void orc_eb_boundary(const orc_params* p, int* is_fluid)
{
    const orc_box fb = grown(p, p->ng);
    const orc_box ib = grown(p, p->ng - 1);
    for (int k = ib.lo[2]; k < ib.lo[2] + ib.n[2]; ++k)
        for (int j = ib.lo[1]; j < ib.lo[1] + ib.n[1]; ++j)
            for (int i = ib.lo[0]; i < ib.lo[0] + ib.n[0]; ++i) {
                int all_covered = 1;
                all_covered &= (AT(is_fluid, fb, i - 1, j, k, 0) == 0) && (AT(is_fluid, fb, i + 1, j, k, 0) == 0);
                all_covered &= (AT(is_fluid, fb, i, j - 1, k, 0) == 0) && (AT(is_fluid, fb, i, j + 1, k, 0) == 0);
                all_covered &= (AT(is_fluid, fb, i, j, k - 1, 0) == 0) && (AT(is_fluid, fb, i, j, k + 1, 0) == 0);
                AT(is_fluid, fb, i, j, k, 1) = (all_covered || AT(is_fluid, fb, i, j, k, 0) == 1) ? 0 : 1;
            }
}

/* LBM.cpp:994-1044, single level (mask == 0 everywhere) */
void orc_eb_forces(const orc_params* p, const int* is_fluid, const double* f, double forces[3])
{
    build_tables();
    const orc_box fb = grown(p, p->ng);
    forces[0] = forces[1] = forces[2] = 0.0;
    for (int k = p->lo[2]; k <= p->hi[2]; ++k)
        for (int j = p->lo[1]; j <= p->hi[1]; ++j)
            for (int i = p->lo[0]; i <= p->hi[0]; ++i) {
                if (AT(is_fluid, fb, i, j, k, 1) != 1) continue;
                double fs[3] = {0.0, 0.0, 0.0};
                for (int q = 0; q < 27; ++q) {
                    const int* ev = EV[q];
                    const int* eb = EV[BOUNCE[q]];
                    const int ir = i + eb[0], jr = j + eb[1], kr = k + eb[2];
                    for (int d = 0; d < 3; ++d)
                        fs[d] += 2.0 * ev[d] * AT(f, fb, ir, jr, kr, q) * AT(is_fluid, fb, ir, jr, kr, 0);
                }
                forces[0] += fs[0];
                forces[1] += fs[1];
                forces[2] += fs[2];
            }
}

/* LBM.cpp:416-422 and 523-544 for a single level */
void orc_step(const orc_params* p, const int* is_fluid, double* f, double* g, double* macro, double* derived,
              double* eq, double* eq_g, double time)
{
    orc_fillpatch(p, f, 0, time);
    orc_fillpatch(p, g, 1, time);
    orc_stream(p, is_fluid, f, 1);
    orc_stream(p, is_fluid, g, 1);
    orc_collide(p, is_fluid, f, g, macro, derived, eq, eq_g, 1);
}
