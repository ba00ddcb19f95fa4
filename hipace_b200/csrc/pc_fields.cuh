// Per-cell arithmetic of the predictor-corrector Bx/By solver and of the open field boundary, host +
// device (csrc/host_check.cu runs it on the CPU against the oracle):
//   Fields::SolvePoissonBxBy            src/fields/Fields.cpp:1008-1078   (right-hand sides)
//   Fields::SetBoundaryCondition (Open) src/fields/Fields.cpp:685-738, SetDirichletBoundaries :628-673
//   the multipole expansion             src/fields/OpenBoundary.H:39-156
#pragma once
#include "common.cuh"
#include "shapes.cuh"

struct BxByRhsPar {
    int c_jz, c_prev_jx, c_prev_jy, c_next_jx, c_next_jy;
    double mu0, dx_inv_half, dy_inv_half, dz_inv_half;
};
// rhs(Bx) = mu0 (-d_y jz + d_z jy),  rhs(By) = mu0 (d_x jz - d_z jx),  d_z f = (f_prev - f_next) / (2 dz)
HPB_HD void bxby_rhs_cell(const SliceView &a, const BxByRhsPar &p, int i, int j, double &rbx, double &rby)
{
    const long o = a.idx(i, j), js = a.jstride;
    const double *jz = a.comp(p.c_jz);
    const double dy_jz = (jz[o + js] - jz[o - js]) * p.dy_inv_half;
    const double dx_jz = (jz[o + 1] - jz[o - 1]) * p.dx_inv_half;
    const double dz_jy = (a.comp(p.c_prev_jy)[o] - a.comp(p.c_next_jy)[o]) * p.dz_inv_half;
    const double dz_jx = (a.comp(p.c_prev_jx)[o] - a.comp(p.c_next_jx)[o]) * p.dz_inv_half;
    rbx = -p.mu0 * dy_jz + p.mu0 * dz_jy;
    rby = p.mu0 * dx_jz + (-p.mu0) * dz_jx;
}

struct PsiEzBzRhsPar { int c_rhomjz, c_jx, c_jy; double f_psi, f_ez, mu0, dx_inv_half, dy_inv_half; };
// the three right-hand sides of Fields.cpp:886-912 (same expressions as poisson.cu's fused load)
HPB_HD void psi_ez_bz_rhs_cell(const SliceView &a, const PsiEzBzRhsPar &p, int i, int j, double r[3])
{
    const long o = a.idx(i, j), js = a.jstride;
    const double *jx = a.comp(p.c_jx), *jy = a.comp(p.c_jy);
    r[0] = p.f_psi * a.comp(p.c_rhomjz)[o];
    r[1] = p.f_ez * ((jx[o + 1] - jx[o - 1]) * p.dx_inv_half) + p.f_ez * ((jy[o + js] - jy[o - js]) * p.dy_inv_half);
    r[2] = p.mu0 * ((jx[o + js] - jx[o - js]) * p.dy_inv_half) + (-p.mu0) * ((jy[o + 1] - jy[o - 1]) * p.dx_inv_half);
}

// ---- open boundary -----------------------------------------------------------------------------
// The free-space potential of the sources s inside 95 % of the largest centred circle,
//   phi(r) = dx dy / (4 pi) sum_s s ln |r - r_s|^2,
// expanded about the origin with z = x + i y and the moments M_k = sum_s s z_s^k, k = 0..18:
//   phi = dx dy / (4 pi) ( M_0 ln |z|^2 - 2 sum_{k >= 1} Re(M_k / z^k) / k )
// -- the 37 real coefficients and the polynomial table of OpenBoundary.H are this series written
// out in x and y.  Coordinates are scaled by 3 / (box diagonal) like the reference's.
constexpr int kMultipoleOrder = 18;
constexpr int kMultipoleN = 2 * (kMultipoleOrder + 1);      // Re, Im of M_0..M_18

// the contribution s z^k of one source cell, t[2k], t[2k+1] = Re, Im
HPB_HD void multipole_terms(double s, double x, double y, double t[kMultipoleN])
{
    double pr = s, pi = 0.;
    for (int k = 0; k <= kMultipoleOrder; ++k) {
        t[2 * k] = pr; t[2 * k + 1] = pi;
        const double nr = pr * x - pi * y, ni = pr * y + pi * x;
        pr = nr; pi = ni;
    }
}
// M_0 ln|z|^2 - 2 sum_k Re(M_k / z^k) / k at the (scaled) point (x, y)
HPB_HD double multipole_value(const double M[kMultipoleN], double x, double y)
{
    const double r2 = x * x + y * y;
    const double ix = x / r2, iy = -y / r2;        // 1 / z
    double phi = M[0] * log(r2);
    double qr = 1., qi = 0.;                       // 1 / z^k
    for (int k = 1; k <= kMultipoleOrder; ++k) {
        const double nr = qr * ix - qi * iy, ni = qr * iy + qi * ix;
        qr = nr; qi = ni;
        phi -= 2.0 * (M[2 * k] * qr - M[2 * k + 1] * qi) / k;
    }
    return phi;
}

// geometry of the expansion (Fields.cpp:695-712)
struct OpenBcPar { int nx, ny; double dx, dy, off_x, off_y, scale, cutoff_sq, dxdy_div_4pi; };
inline bool open_bc_par(int nx, int ny, double dx, double dy, double lo_x, double hi_x, double lo_y,
                        double hi_y, OpenBcPar &p)
{
    const double lx = hi_x - lo_x, ly = hi_y - lo_y;
    p.nx = nx; p.ny = ny; p.dx = dx; p.dy = dy;
    p.scale = 3.0 / sqrt(lx * lx + ly * ly);
    const double radius = fmin(fmin(fabs(lo_x), fabs(hi_x)), fmin(fabs(lo_y), fabs(hi_y)));
    p.cutoff_sq = (0.95 * radius * p.scale) * (0.95 * radius * p.scale);
    p.off_x = 0.5 * (lo_x + hi_x - dx * (nx - 1));           // GetPosOffset of the valid box
    p.off_y = 0.5 * (lo_y + hi_y - dy * (ny - 1));
    p.dxdy_div_4pi = dx * dy / (4.0 * 3.14159265358979323846);
    return radius > 0.;
}
// source cell c = j nx + i: its multipole terms, or false outside the cut-off circle
HPB_HD bool open_bc_source(const OpenBcPar &p, long c, double s, double t[kMultipoleN])
{
    const int j = (int)(c / p.nx), i = (int)(c - (long)j * p.nx);
    const double x = (i * p.dx + p.off_x) * p.scale, y = (j * p.dy + p.off_y) * p.scale;
    if (x * x + y * y > p.cutoff_sq) return false;
    multipole_terms(s, x, y, t);
    return true;
}
// edge slot e in [0, 2 (nx + ny)): the cell it belongs to and the value SetDirichletBoundaries
// (:628-673, offset = factor = 1) adds to it; corners are hit by two slots
HPB_HD double open_bc_edge(const OpenBcPar &p, int e, const double M[kMultipoleN], long &cell)
{
    int i, j; double xi, yj, dd;
    if (e < 2 * p.nx) {                     // j_lo / j_hi edges, i changes
        i = e % p.nx; const bool hi = e >= p.nx;
        j = hi ? p.ny - 1 : 0;
        xi = i; yj = hi ? p.ny : -1; dd = p.dy * p.dy;
    } else {                                // i_lo / i_hi edges
        const int q = e - 2 * p.nx;
        j = q % p.ny; const bool hi = q >= p.ny;
        i = hi ? p.nx - 1 : 0;
        xi = hi ? p.nx : -1; yj = j; dd = p.dx * p.dx;
    }
    cell = (long)j * p.nx + i;
    const double x = (xi * p.dx + p.off_x) * p.scale, y = (yj * p.dy + p.off_y) * p.scale;
    return -(p.dxdy_div_4pi * multipole_value(M, x, y)) / dd;
}
