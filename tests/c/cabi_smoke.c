/* Plain-C caller of the marbles_b200 C ABI (no Python, no torch): what a host application links against.
 * Without a GPU it checks the loud failure; with one it runs a few steps of a small periodic box and checks
 * mass conservation.  Built and run by tests/test_cabi.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "marbles_b200.h"

int main(void)
{
    if (mbl_version() < 100) {
        printf("FAIL version %d\n", mbl_version());
        return 1;
    }
    mbl_params p;
    memset(&p, 0, sizeof(p));
    p.nu = 0.01, p.alpha = 0.01, p.R = 1.0, p.gamma = 5.0 / 3.0, p.mesh_speed = 1.0;
    for (int d = 0; d < 3; ++d) p.periodic[d] = 1;
    mbl_ctx* ctx = NULL;
    if (mbl_create(&p, 0, &ctx) != 0) {
        const char* msg = mbl_last_error();
        printf("NO_DEVICE %s\n", msg);
        return (msg && strstr(msg, "no CPU path")) ? 0 : 1; /* must fail loudly, not fall back */
    }
    const int n = 16;
    mbl_level_geom g;
    memset(&g, 0, sizeof(g));
    for (int d = 0; d < 3; ++d) {
        g.dom_lo[d] = g.lo[d] = 0, g.dom_hi[d] = g.hi[d] = n - 1;
        g.dx[d] = 2.0 / n, g.inv_dx[d] = n / 2.0, g.prob_lo[d] = -1.0, g.prob_hi[d] = 1.0;
    }
    g.dt = 1.0;
    if (mbl_level_define(ctx, 0, &g, NULL)) return 2;
    double ic[16] = {0};
    ic[0] = 1.0; /* density */
    ic[4] = 0.1; /* v0 */
    ic[5] = ic[6] = 1.0, ic[7] = 0.0; /* omega */
    ic[8] = 1.0 / 3.141592653589793; /* wave_length */
    ic[9] = 1.0 / 3.0, ic[10] = 5.0 / 3.0, ic[11] = 1.0, ic[12] = 1.0; /* T0, gamma, R, c_s */
    if (mbl_initialize(ctx, 0, 1 /* taylorgreen */, ic, 16)) return 3;
    const size_t cells = (size_t)n * n * n;
    double* f = (double*)malloc(27 * cells * sizeof(double));
    if (mbl_download(ctx, 0, MBL_F, f, 0)) return 4;
    double m0 = 0.0;
    for (size_t i = 0; i < 27 * cells; ++i) m0 += f[i];
    if (mbl_step(ctx, 0, 8, 0.0, 0) || mbl_sync(ctx)) {
        printf("FAIL step: %s\n", mbl_last_error());
        return 5;
    }
    if (mbl_download(ctx, 0, MBL_F, f, 0)) return 6;
    double m1 = 0.0;
    for (size_t i = 0; i < 27 * cells; ++i) m1 += f[i];
    printf("DEVICE mass %.15g -> %.15g, %lld kernel launches\n", m0, m1, (long long)mbl_launch_count(ctx));
    free(f);
    mbl_destroy(ctx);
    return fabs(m1 - m0) <= 1e-11 * m0 ? 0 : 7;
}
