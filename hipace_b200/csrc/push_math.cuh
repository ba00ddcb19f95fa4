// Plasma momentum derivative and the second-order sub-step built on dual numbers, host + device:
// src/particles/pusher/PushPlasmaParticles.H:39-75, src/utils/DualNumbers.H:13-43,
// src/particles/pusher/PlasmaParticleAdvance.cpp:152-166.  Used by the order-2 kernel of particles.cu
// and the generic-order kernels of generic_order.cu; tests/test_device_math_host.py runs it on the CPU
// against the reference's own headers.
#pragma once
#include "shapes.cuh"

struct Dual { double v, e; };
HPB_HD Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.e + b.e}; }
HPB_HD Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.e - b.e}; }
HPB_HD Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.e * b.v + a.v * b.e}; }
HPB_HD Dual operator*(Dual a, double b) { return {a.v * b, a.e * b}; }
HPB_HD Dual operator*(double a, Dual b) { return {a * b.v, a * b.e}; }
HPB_HD Dual operator+(Dual a, double b) { return {a.v + b, a.e}; }
HPB_HD Dual operator+(double a, Dual b) { return {a + b.v, b.e}; }
HPB_HD Dual operator-(Dual a, double b) { return {a.v - b, a.e}; }

struct PushFields { double ExmBy, EypBx, Ez, Bx_c, By_c, Bz; };

// Aabssq_norm, AabssqDx_norm, AabssqDy_norm of PushPlasmaParticles.H:32-34
struct PushLaser { double A, ADx, ADy; };

// PlasmaMomentumPush<T>, PushPlasmaParticles.H:39-75 (LASER = false: the laser terms are zero)
template <class T, bool LASER>
HPB_HD void momentum_push(const T &ux, const T &uy, const T &psi_inv,
                                              const PushFields &f, double clight_inv, double qmc,
                                              T &dz_ux, T &dz_uy, T &dz_psi, const PushLaser &las)
{
    const double c2 = clight_inv * clight_inv;
    if (LASER) {
        const T gamma_psi = 0.5 * psi_inv * psi_inv * ((1.0 + las.A) + ux * ux * c2 + uy * uy * c2) + 0.5;
        dz_ux = qmc * (gamma_psi * f.ExmBy + f.By_c + (uy * f.Bz) * psi_inv) - las.ADx * psi_inv;
        dz_uy = qmc * (gamma_psi * f.EypBx - f.Bx_c - (ux * f.Bz) * psi_inv) - las.ADy * psi_inv;
    } else {
        const T gamma_psi = 0.5 * psi_inv * psi_inv * (1.0 + ux * ux * c2 + uy * uy * c2) + 0.5;
        dz_ux = qmc * (gamma_psi * f.ExmBy + f.By_c + (uy * f.Bz) * psi_inv);
        dz_uy = qmc * (gamma_psi * f.EypBx - f.Bx_c - (ux * f.Bz) * psi_inv);
    }
    dz_psi = (qmc * clight_inv) * ((ux * f.ExmBy + uy * f.EypBx) * clight_inv * psi_inv - f.Ez);
}

template <bool LASER>
HPB_HD void push_substep(double &ux, double &uy, double &psi,
                                             const PushFields &f, double clight_inv, double qmc,
                                             double sdz, const PushLaser &las)
{
    const double psi_inv = 1.0 / psi;
    double dz_ux, dz_uy, dz_psi;
    momentum_push<double, LASER>(ux, uy, psi_inv, f, clight_inv, qmc, dz_ux, dz_uy, dz_psi, las);
    const Dual ux_d{ux, dz_ux}, uy_d{uy, dz_uy}, pi_d{psi_inv, -psi_inv * psi_inv * dz_psi};
    Dual d_ux, d_uy, d_psi;
    momentum_push<Dual, LASER>(ux_d, uy_d, pi_d, f, clight_inv, qmc, d_ux, d_uy, d_psi, las);
    ux += sdz * dz_ux + 0.5 * sdz * sdz * d_ux.e;
    uy += sdz * dz_uy + 0.5 * sdz * sdz * d_uy.e;
    psi += sdz * dz_psi + 0.5 * sdz * sdz * d_psi.e;
}

