// bcic.cuh -- inlet functors and initial-condition functors shared by the single-box kernels (kernels.cu) and the
// multi-box / multi-level kernels (patch.cu).  Reference: Source/VelocityBC.H:44-190, Source/IC.H:42-428.
#pragma once
#include "kernels.cuh"

namespace mbl {

// inlet functors (VelocityBC.H:44-190) at the literal (un-wrapped) global ghost index
__device__ __forceinline__ void vel_bc_op(const BcInfo& B, const int dhi[3], int gi, int gj, int gk, double& rho,
                                          double vel[3], double& R, double& T, double& gamma)
{
    const int iv[3] = {gi, gj, gk};
    if (B.vbc_kind == 1) {
        rho = B.vbc_rho;
        vel[B.vbc_dir] = B.vbc_u;
    } else if (B.vbc_kind == 2) {
        rho = B.vbc_rho;
        const double c1 = (double)(iv[1] * (dhi[1] - iv[1]));
        const double c2 = (double)(iv[2] * (dhi[2] - iv[2]));
        const double d = (double)(dhi[1] + 1);
        vel[0] = 16.0 * B.vbc_u * c1 * c2 / (d * d * d * d);
    } else if (B.vbc_kind == 3) {
        rho = B.vbc_rho;
        const int nd = B.vbc_normal_dir;
        const double height = B.prob_hi[nd] - B.prob_lo[nd];
        const double x = B.prob_lo[nd] + (iv[nd] + 0.5) * B.dx[nd];
        vel[B.vbc_tangential_dir] = 4.0 * B.vbc_u * x * (height - x) / (height * height);
    } else {
        return;
    }
    R = B.vbc_R;
    T = B.vbc_T;
    gamma = B.vbc_gamma;
}

// ic::{Constant, TaylorGreen, ViscosityTest, ThermalDiffusivityTest, SodTest} at global cell (gi, gj, gk)
__device__ __forceinline__ void ic_state(const IcInfo& I, const BcInfo& B, int gi, int gj, int gk, double& rho,
                                         double vel[3], double& T, double& R, double& gamma)
{
    rho = 1.0, vel[0] = vel[1] = vel[2] = 0.0, T = 1.0 / 3.0, R = 1.0, gamma = 1.667;
    const double PI = 3.14159265358979323846;
    if (I.kind == 0) {  // IC.H:42-57
        rho = I.density;
        vel[0] = I.vel[0], vel[1] = I.vel[1], vel[2] = I.vel[2];
        T = I.T0, R = I.R, gamma = I.gamma;
    } else if (I.kind == 1) {  // IC.H:119-153
        const double x = B.prob_lo[0] + (gi + 0.5) * B.dx[0];
        const double y = B.prob_lo[1] + (gj + 0.5) * B.dx[1];
        const double z = B.prob_lo[2] + (gk + 0.5) * B.dx[2];
        const double Lc = 1.0 / PI;
        rho = I.density + I.density * I.v0 * I.v0 / 16.0 * (cos(2.0 * I.omega[0] * x / Lc) + cos(2.0 * I.omega[1] * y / Lc)) *
                              (cos(2.0 * I.omega[2] * z / Lc) + 2.0);
        vel[0] = I.v0 * sin(I.omega[0] * x / Lc) * cos(I.omega[1] * y / Lc) * cos(I.omega[2] * z / Lc);
        vel[1] = -I.v0 * cos(I.omega[0] * x / Lc) * sin(I.omega[1] * y / Lc) * cos(I.omega[2] * z / Lc);
        vel[2] = 0.0;
        T = I.T0, R = I.R, gamma = 5.0 / 3.0;
    } else if (I.kind == 2) {  // IC.H:213-240
        const double y = B.prob_lo[1] + (gj + 0.5 * 0.0) * B.dx[1];
        rho = I.density;
        vel[0] = I.vel[0] + 0.010 * I.c_s * sin(2.0 * PI * y / I.wave_length);
        vel[1] = I.vel[1], vel[2] = I.vel[2];
        T = I.T0, R = I.R, gamma = I.gamma;
    } else if (I.kind == 3) {  // IC.H:302-333
        const double y = B.prob_lo[1] + (gj + 0.5 * 0.0) * B.dx[1];
        R = I.R;
        const double pressure = I.density * R * I.T0;
        rho = I.density + 0.0010 * I.T0 * sin(2.0 * PI * y / I.wave_length);
        vel[0] = I.vel[0], vel[1] = I.vel[1], vel[2] = I.vel[2];
        gamma = I.gamma;
        T = pressure / (rho * R);
    } else if (I.kind == 4) {  // IC.H:394-428
        const double x = B.prob_lo[0] + (gi + 0.5 * 0.0) * B.dx[0];
        R = I.R, gamma = I.gamma;
        vel[0] = I.vel[0], vel[1] = I.vel[1], vel[2] = I.vel[2];
        const double s = 0.5 * (1.0 + tanh((x - I.x_disc) * 3.0));
        rho = I.density + s * (I.density_ratio * I.density - I.density);
        T = I.T0 + s * (I.temperature_ratio * I.T0 - I.T0);
    }
}

}  // namespace mbl
