// Per-particle arithmetic of the particle kernels for ANY hipace.depos_order_xy (0..3) and
// hipace.depos_derivative_type (0..2), host + device.  The order-2 / centred-derivative default has
// its own warp-aggregated, shared-memory staged kernels in particles.cu; every other combination
// (and, with HPB_GENERIC=1, the default as a cross-check) runs the plain one-thread-per-particle
// kernels of generic_order.cu built on these functions.
//
// Each function is the body of one particle of
//   ::DepositCurrent            src/particles/deposition/PlasmaDepositCurrent.cpp:155-246
//   DepositCurrentSlice (beam)  src/particles/deposition/BeamDepositCurrent.cpp:136-194
//   ::ExplicitDeposition        src/particles/deposition/ExplicitDeposition.cpp:140-261
//   doGatherShapeN              src/particles/particles_utils/FieldGather.H:45-96
//   doLaserGatherShapeN         src/particles/particles_utils/FieldGather.H:236-330
//   AdvancePlasmaParticles      src/particles/pusher/PlasmaParticleAdvance.cpp:92-217
// with the scatter abstracted as a functor (fp64 RED on the device, a plain += in the host harness
// csrc/host_check.cu that tests/test_device_math_host.py drives against the oracle and the
// reference's own headers).
#pragma once
#include "common.cuh"
#include "push_math.cuh"

struct GenGrid { double x_off, y_off, dx_inv, dy_inv; };

template <int ORDER, bool DERIV>
HPB_HD void gen_laser_gather(const SliceView &a, int c_aabs, const GenGrid &gr, double xp, double yp,
                             double &A, double &ADx, double &ADy)
{
    double sx[ORDER + 1], sy[ORDER + 1];
    const int i0 = hpb_shape<ORDER>((xp - gr.x_off) * gr.dx_inv, sx);
    const int j0 = hpb_shape<ORDER>((yp - gr.y_off) * gr.dy_inv, sy);
    const double *ab = a.comp(c_aabs);
    const long js = a.jstride;
    A = 0.; ADx = 0.; ADy = 0.;
#pragma unroll
    for (int iy = 0; iy <= ORDER; ++iy) {
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix) {
            const long o = a.idx(i0 + ix, j0 + iy);
            const double w = sx[ix] * sy[iy];
            A += w * ab[o];
            if (DERIV) {
                ADx += w * 0.5 * gr.dx_inv * (ab[o + 1] - ab[o - 1]);
                ADy += w * 0.5 * gr.dy_inv * (ab[o + js] - ab[o - js]);
            }
        }
    }
}

// Plasma current deposition of one valid particle.  c[] = {jx, jy, jz, rho, chi, rhomjz} (-1: skip;
// jz is what the predictor-corrector solver needs, PlasmaDepositCurrent.cpp:223).
// Returns false when the particle violates the quasi-static limits (nothing deposited; the caller
// zeroes its weight, invalidates the id and counts it, PlasmaDepositCurrent.cpp:197-204).
struct GenDepositPar {
    double clightinv, charge_invvol, charge_mu0_mass_ratio, max_qsa, laser_norm;
    int c_aabs;
    double clight;
};
template <int ORDER, bool LASER, class Add>
HPB_HD bool gen_deposit_current(const SliceView &a, const int c[6], const GenGrid &gr,
                                const GenDepositPar &p, double xp, double yp, double w, double ux,
                                double uy, double psi, const Add &add)
{
    const double psi_inv = 1.0 / psi;
    const double vx_c = ux * psi_inv, vy_c = uy * psi_inv;
    double Aabssqp = 0.;
    if (LASER) {
        double ADx, ADy;
        gen_laser_gather<ORDER, false>(a, p.c_aabs, gr, xp, yp, Aabssqp, ADx, ADy);
        Aabssqp *= p.laser_norm;
    }
    const double gamma_psi = 0.5 * ((1.0 + 0.5 * Aabssqp) * psi_inv * psi_inv
                                    + vx_c * vx_c * p.clightinv * p.clightinv
                                    + vy_c * vy_c * p.clightinv * p.clightinv + 1.0);
    if (gamma_psi < 0.0 || gamma_psi > p.max_qsa || psi_inv < 0.0) return false;
    double sx[ORDER + 1], sy[ORDER + 1];
    const int i0 = hpb_shape<ORDER>((xp - gr.x_off) * gr.dx_inv, sx);
    const int j0 = hpb_shape<ORDER>((yp - gr.y_off) * gr.dy_inv, sy);
    const double q_invvol = p.charge_invvol * w;
    const double chi_fac = p.charge_mu0_mass_ratio * psi_inv;
#pragma unroll
    for (int iy = 0; iy <= ORDER; ++iy) {
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix) {
            const long o = a.idx(i0 + ix, j0 + iy);
            const double cd = q_invvol * sx[ix] * sy[iy];
            if (c[0] >= 0) { add(a.comp(c[0]) + o, cd * vx_c); add(a.comp(c[1]) + o, cd * vy_c); }
            if (c[2] >= 0) add(a.comp(c[2]) + o, cd * (gamma_psi - 1.0) * p.clight);
            if (c[3] >= 0) add(a.comp(c[3]) + o, cd * gamma_psi);
            if (c[4] >= 0) add(a.comp(c[4]) + o, cd * chi_fac);
            if (c[5] >= 0) add(a.comp(c[5]) + o, cd);
        }
    }
    return true;
}

template <int ORDER, class Add>
HPB_HD void gen_beam_deposit(const SliceView &a, int c_jx, int c_jy, int c_jz, const GenGrid &gr,
                             double clightsq, double q_invvol, double xp, double yp, double w,
                             double ux, double uy, double uz, const Add &add)
{
    const double gaminv = 1.0 / sqrt(1.0 + ux * ux * clightsq + uy * uy * clightsq + uz * uz * clightsq);
    const double wq = q_invvol * w;
    const double wqx = wq * (ux * gaminv), wqy = wq * (uy * gaminv), wqz = wq * (uz * gaminv);
    double sx[ORDER + 1], sy[ORDER + 1];
    const int i0 = hpb_shape<ORDER>((xp - gr.x_off) * gr.dx_inv, sx);
    const int j0 = hpb_shape<ORDER>((yp - gr.y_off) * gr.dy_inv, sy);
#pragma unroll
    for (int iy = 0; iy <= ORDER; ++iy) {
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix) {
            const long o = a.idx(i0 + ix, j0 + iy);
            const double s = sx[ix] * sy[iy];
            if (c_jx >= 0) { add(a.comp(c_jx) + o, s * wqx); add(a.comp(c_jy) + o, s * wqy); }
            if (c_jz >= 0) add(a.comp(c_jz) + o, s * wqz);
        }
    }
}

// Explicit (Sy, Sx) deposition of one valid particle
struct GenExplicitPar {
    int c_sy, c_sx, c_bz, c_ez, c_exmby, c_eypbx, c_aabs;
    double clight, clight_inv, charge_invvol_mu0, q_mass_ratio, laser_fac;
};
template <int ORDER, int DTYPE, bool LASER, class Add>
HPB_HD void gen_explicit_deposition(const SliceView &a, const GenGrid &gr, const GenExplicitPar &p,
                                    double xp, double yp, double w, double ux, double uy, double psi,
                                    const Add &add)
{
    constexpr int NS = ORDER + DTYPE + 1;
    const double psi_inv = 1.0 / psi;
    const double vx = ux * psi_inv * p.clight_inv, vy = uy * psi_inv * p.clight_inv;
    const double cdm = p.charge_invvol_mu0 * w;
    const double qmr = p.q_mass_ratio;
    double Aabssqp = 0.;
    if (LASER) {      // fully gathered first, unlike the per-cell derivatives below (:167-175)
        double t0, t1;
        gen_laser_gather<ORDER, false>(a, p.c_aabs, gr, xp, yp, Aabssqp, t0, t1);
        Aabssqp *= p.laser_fac * qmr * qmr;
    }
    const double gamma_psi = 0.5 * ((1.0 + 0.5 * Aabssqp) * psi_inv * psi_inv + vx * vx + vy * vy + 1.0);
    double sx[NS], dsx[NS], sy[NS], dsy[NS];
    const int i0 = hpb_dshape<DTYPE, ORDER>((xp - gr.x_off) * gr.dx_inv, sx, dsx);
    const int j0 = hpb_dshape<DTYPE, ORDER>((yp - gr.y_off) * gr.dy_inv, sy, dsy);
    const double *Bz = a.comp(p.c_bz), *Ez = a.comp(p.c_ez);
    const double *ExmBy = a.comp(p.c_exmby), *EypBx = a.comp(p.c_eypbx);
    const double *ab = LASER ? a.comp(p.c_aabs) : nullptr;
    double *Sy = a.comp(p.c_sy), *Sx = a.comp(p.c_sx);
    const long js = a.jstride;
#pragma unroll
    for (int iy = 0; iy < NS; ++iy) {
#pragma unroll
        for (int ix = 0; ix < NS; ++ix) {
            // the corners of the centred-derivative stencil carry zero weight (:193-198)
            if (DTYPE == 2 && (ix == 0 || ix == NS - 1) && (iy == 0 || iy == NS - 1)) continue;
            const long o = a.idx(i0 + ix, j0 + iy);
            const double shx = sx[ix], shdx = dsx[ix], shy = sy[iy], shdy = dsy[iy];
            const double Bz_v = Bz[o], Ez_v = Ez[o], ExmBy_v = ExmBy[o], EypBx_v = EypBx[o];
            double ADx = 0., ADy = 0.;
            if (LASER && shx * shy != 0.) {          // "avoid going outside of the domain" (:215-226)
                ADx = (ab[o + 1] - ab[o - 1]) * 0.5 * gr.dx_inv * p.laser_fac * p.clight;
                ADy = (ab[o + js] - ab[o - js]) * 0.5 * gr.dy_inv * p.laser_fac * p.clight;
            }
            add(Sy + o, cdm * (
                - shx * shy * (
                    - Bz_v * vx
                    + (Ez_v * vy + ExmBy_v * (-vx * vy) + EypBx_v * (gamma_psi - vy * vy)) * p.clight_inv
                    - 0.25 * ADy * qmr * psi_inv
                  ) * qmr * psi_inv
                + (- shdx * shy * gr.dx_inv * (-vx * vy)
                   - shx * shdy * gr.dy_inv * (gamma_psi - vy * vy - 1.0)) * p.clight));
            add(Sx + o, cdm * (
                + shx * shy * (
                    + Bz_v * vy
                    + (Ez_v * vx + ExmBy_v * (gamma_psi - vx * vx) + EypBx_v * (-vx * vy)) * p.clight_inv
                    - 0.25 * ADx * qmr * psi_inv
                  ) * qmr * psi_inv
                + (+ shdx * shy * gr.dx_inv * (gamma_psi - vx * vx - 1.0)
                   + shx * shdy * gr.dy_inv * (-vx * vy)) * p.clight));
        }
    }
}

// doGatherShapeN<ORDER>: always the nodal derivative shapes, (ORDER+2)^2 cells
template <int ORDER>
HPB_HD GatheredFields gen_gather(const SliceView &a, int c_psi, int c_ez, int c_bx, int c_by, int c_bz,
                                 const GenGrid &gr, double xp, double yp)
{
    constexpr int NS = ORDER + 2;
    double sx[NS], dsx[NS], sy[NS], dsy[NS];
    const int i0 = hpb_dshape<1, ORDER>((xp - gr.x_off) * gr.dx_inv, sx, dsx);
    const int j0 = hpb_dshape<1, ORDER>((yp - gr.y_off) * gr.dy_inv, sy, dsy);
    const double *Psi = a.comp(c_psi), *Ez = a.comp(c_ez), *Bx = a.comp(c_bx);
    const double *By = a.comp(c_by), *Bz = a.comp(c_bz);
    GatheredFields f = {0., 0., 0., 0., 0., 0.};
#pragma unroll
    for (int iy = 0; iy < NS; ++iy) {
#pragma unroll
        for (int ix = 0; ix < NS; ++ix) {
            const long o = a.idx(i0 + ix, j0 + iy);
            const double psi_v = Psi[o];
            f.ExmBy += (dsx[ix] * sy[iy]) * psi_v * gr.dx_inv;
            f.EypBx += (sx[ix] * dsy[iy]) * psi_v * gr.dy_inv;
            const double w = sx[ix] * sy[iy];
            f.Ez += w * Ez[o];
            f.Bx += w * Bx[o];
            f.By += w * By[o];
            f.Bz += w * Bz[o];
        }
    }
    return f;
}

// AdvancePlasmaParticles of one valid particle.  st[] in: x_prev, y_prev, ux_half, uy_half,
// psi_half; out[] = x, y, ux, uy, psi and (unless temp_slice) the updated st[].  Returns false if
// the particle left through an absorbing boundary (caller: w = 0, id invalid).
struct GenPushPar {
    int c_psi, c_ez, c_bx, c_by, c_bz, c_aabs;
    double clight, qmc, dz;          // qmc = charge / (mass c); dz already divided by n_subcycles
    int n_subcycles, temp_slice, bc;
    double lox, loy, hix, hiy, laser_norm;
};
template <int ORDER, bool LASER>
HPB_HD bool gen_advance_plasma(const SliceView &a, const GenGrid &gr, const GenPushPar &p,
                               double st[5], double out[5])
{
    const double clight_inv = 1.0 / p.clight;
    double xp0 = st[0], yp0 = st[1], ux0 = st[2], uy0 = st[3], psi0 = st[4];
    double xp = xp0, yp = yp0, ux = ux0, uy = uy0, psi = psi0;
    for (int isc = 0; isc < p.n_subcycles; ++isc) {
        xp = xp0; yp = yp0;
        const GatheredFields g = gen_gather<ORDER>(a, p.c_psi, p.c_ez, p.c_bx, p.c_by, p.c_bz, gr, xp, yp);
        PushFields f = {g.ExmBy, g.EypBx, g.Ez, g.Bx * p.clight, g.By * p.clight, g.Bz};
        PushLaser las = {0., 0., 0.};
        if (LASER) {        // PlasmaParticleAdvance.cpp:123-133
            gen_laser_gather<ORDER, true>(a, p.c_aabs, gr, xp, yp, las.A, las.ADx, las.ADy);
            las.A *= 0.5 * p.laser_norm;
            las.ADx *= 0.25 * p.clight * p.laser_norm;
            las.ADy *= 0.25 * p.clight * p.laser_norm;
        }
        constexpr int nsub = 4;
        const double sdz = p.dz / nsub;
        ux = ux0; uy = uy0; psi = psi0;
        for (int isub = 0; isub < nsub; ++isub) push_substep<LASER>(ux, uy, psi, f, clight_inv, p.qmc, sdz, las);
        xp += p.dz * clight_inv * (ux * (1.0 / psi));
        yp += p.dz * clight_inv * (uy * (1.0 / psi));
        if (enforce_particle_bc(xp, yp, ux, uy, p.bc, p.lox, p.loy, p.hix, p.hiy)) return false;
        out[0] = xp; out[1] = yp;
        if (!p.temp_slice) {
            st[0] = xp; st[1] = yp; st[2] = ux; st[3] = uy; st[4] = psi;
            xp0 = xp; yp0 = yp; ux0 = ux; uy0 = uy; psi0 = psi;
        }
        for (int isub = 0; isub < nsub / 2; ++isub) push_substep<LASER>(ux, uy, psi, f, clight_inv, p.qmc, sdz, las);
        out[2] = ux; out[3] = uy; out[4] = psi;
    }
    return true;
}
