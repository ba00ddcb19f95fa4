// Per-cell arithmetic of the laser envelope ADVANCE with the fft solver (second part),
// host + device: MultiLaser::AdvanceSliceFFT (src/laser/MultiLaser.cpp:609-801), InterpolateChi
// (:334-407), UpdateLaserAabs (:214-291) from a stored envelope slice.  The envelope equation in the
// Benedetti et al. (2017) discretisation: time levels n-1, n, n+1 and slices j, j+1, j+2.
#pragma once
#include "common.cuh"
#include "shapes.cuh"

struct hpb_c2 { double re, im; };
HPB_HD hpb_c2 c2(double re, double im) { hpb_c2 z; z.re = re; z.im = im; return z; }
HPB_HD hpb_c2 operator+(hpb_c2 a, hpb_c2 b) { return c2(a.re + b.re, a.im + b.im); }
HPB_HD hpb_c2 operator-(hpb_c2 a, hpb_c2 b) { return c2(a.re - b.re, a.im - b.im); }
HPB_HD hpb_c2 operator*(hpb_c2 a, hpb_c2 b) { return c2(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
HPB_HD hpb_c2 operator*(double a, hpb_c2 b) { return c2(a * b.re, a * b.im); }

// compute_shape_factor<order> for the laser <-> field interpolation (lasers.interp_order 0..2)
HPB_HD int laser_interp_shape(double xmid, int order, double w[3])
{
    w[0] = w[1] = w[2] = 0.;
    if (order == 0) return hpb_shape<0>(xmid, w);
    if (order == 1) return hpb_shape<1>(xmid, w);
    const double xfloor = floor(xmid + 0.5), xint = xmid - xfloor;      // the literal order-2 polynomials
    w[0] = 0.5 * (0.5 - xint) * (0.5 - xint);
    w[1] = 0.75 - xint * xint;
    w[2] = 0.5 * (0.5 + xint) * (0.5 + xint);
    return (int)xfloor - 1;
}

// on-axis phase terms (:651-685) from the sums h0, h1, h2 of n00j00, n00jp1, n00jp2 over the centre cells
struct LaserPhase { hpb_c2 exp1, exp2; double djn; };
HPB_HD LaserPhase laser_phase(hpb_c2 h0, hpb_c2 h1, hpb_c2 h2, double dz, int use_phase)
{
    const double pi = 3.14159265358979323846;
    double tj00 = 0., tjp1 = 0., tjp2 = 0.;
    if (use_phase) { tj00 = atan2(h0.im, h0.re); tjp1 = atan2(h1.im, h1.re); tjp2 = atan2(h2.im, h2.re); }
    double dt1 = tj00 - tjp1, dt2 = tjp1 - tjp2;
    if (dt1 < -1.5 * pi) dt1 += 2. * pi;
    if (dt1 > 1.5 * pi) dt1 -= 2. * pi;
    if (dt2 < -1.5 * pi) dt2 += 2. * pi;
    if (dt2 > 1.5 * pi) dt2 -= 2. * pi;
    LaserPhase p;
    p.exp1 = c2(cos(tj00 - tjp1), sin(tj00 - tjp1));
    p.exp2 = c2(cos(tj00 - tjp2), sin(tj00 - tjp2));
    p.djn = (-3. * dt1 + dt2) / (2. * dz);
    return p;
}

// the sums of the three n-level slices over the centre cell(s) of the grid (:651-664): two cells per
// direction for an even size, one for an odd size
HPB_HD void laser_axis_sums(const hpb_c2 *n00j00, const hpb_c2 *n00jp1, const hpb_c2 *n00jp2, int nx, int ny,
                            hpb_c2 h[3])
{
    const int imid = (nx + 1) / 2, jmid = (ny + 1) / 2;
    const int i_lo = nx % 2 == 0 ? imid - 1 : imid, j_lo = ny % 2 == 0 ? jmid - 1 : jmid;
    h[0] = h[1] = h[2] = c2(0., 0.);
    for (int j = j_lo; j <= jmid; ++j)
        for (int i = i_lo; i <= imid; ++i) {
            const long o = (long)j * nx + i;
            h[0] = h[0] + n00j00[o]; h[1] = h[1] + n00jp1[o]; h[2] = h[2] + n00jp2[o];
        }
}

// |A| of the diagnostic at column i: the y = mid-domain line of an xz diagnostic (order-1
// interpolation: mean of the two central rows for even ny, diagnostics/Diagnostic.cpp:393-407)
HPB_HD double laser_diag_xz_abs(const hpb_c2 *env, int i, int nx, int ny)
{
    hpb_c2 v = env[(long)(ny / 2) * nx + i];
    if (ny % 2 == 0) v = 0.5 * (env[(long)(ny / 2 - 1) * nx + i] + v);
    return sqrt(v.re * v.re + v.im * v.im);
}

// MultiLaser::InSituComputeDiags (src/laser/MultiLaser.cpp:923-1001), one laser cell: |a|^2 (for the
// maximum and the sum), its first and second moments in x and y, and the cell's share of the on-axis
// value (the centre cell, or the mean of the two / four centre cells)
HPB_HD void insitu_laser_terms(const hpb_c2 *env, int i, int j, int nx, int ny, double dx, double dy,
                               double x_off, double y_off, double out[8])
{
    const hpb_c2 a = env[(long)j * nx + i];
    const double aabssq = a.re * a.re + a.im * a.im;
    const double x = i * dx + x_off, y = j * dy + y_off;
    const int xlo = (nx - 1) / 2, xhi = nx / 2, ylo = (ny - 1) / 2, yhi = ny / 2;
    const bool on_axis = (i == xlo || i == xhi) && (j == ylo || j == yhi);
    out[0] = aabssq; out[1] = aabssq; out[2] = aabssq * x; out[3] = aabssq * x * x;
    out[4] = aabssq * y; out[5] = aabssq * y * y;
    out[6] = on_axis ? a.re : 0.; out[7] = on_axis ? a.im : 0.;
}

struct LaserAdvPar {
    int nx, ny, step0;                 // step0: first time step (only two time levels exist)
    double dx, dy, dz, c, dt, k0;
    double dkx, dky;                   // 2 pi / box length
};

// five-point Laplacian, zero on the edge cells (:703-722)
HPB_HD hpb_c2 laser_lap(const hpb_c2 *a, int i, int j, const LaserAdvPar &p)
{
    if (i == 0 || j == 0 || i == p.nx - 1 || j == p.ny - 1) return c2(0., 0.);
    const long o = (long)j * p.nx + i;
    const hpb_c2 c = a[o];
    const hpb_c2 xx = a[o + 1] + a[o - 1] - 2.0 * c, yy = a[o + p.nx] + a[o - p.nx] - 2.0 * c;
    return (1.0 / (p.dx * p.dx)) * xx + (1.0 / (p.dy * p.dy)) * yy;
}

// right-hand side of the envelope equation at cell (i, j) (:724-763)
struct LaserPlanes { const hpb_c2 *nm1j00, *nm1jp1, *nm1jp2, *n00j00, *n00jp1, *n00jp2, *np1jp1, *np1jp2; };
HPB_HD hpb_c2 laser_rhs_cell(const LaserPlanes &L, const double *chi, int i, int j, const LaserAdvPar &p,
                             const LaserPhase &ph)
{
    const long o = (long)j * p.nx + i;
    const double cdt = p.c * p.dt, cdtdz = p.c * p.dt * p.dz;
    const hpb_c2 a00 = L.n00j00[o];
    if (p.step0) {
        const hpb_c2 t1 = (8.0 / cdtdz) * ((L.n00jp1[o] - L.np1jp1[o]) * ph.exp1);
        const hpb_c2 t2 = (2.0 / cdtdz) * ((L.np1jp2[o] - L.n00jp2[o]) * ph.exp2);
        const hpb_c2 t3 = (2.0 * chi[o]) * a00;
        const hpb_c2 t4 = laser_lap(L.n00j00, i, j, p);
        const hpb_c2 f = c2(-6.0 / cdtdz, 4.0 * ph.djn / cdt + 4.0 * p.k0 / cdt);
        return t1 + t2 + t3 - t4 + f * a00;
    }
    const double c2dt2 = p.c * p.c * p.dt * p.dt;
    const hpb_c2 am1 = L.nm1j00[o];
    const hpb_c2 t1 = (4.0 / cdtdz) * ((L.nm1jp1[o] - L.np1jp1[o]) * ph.exp1);
    const hpb_c2 t2 = (1.0 / cdtdz) * ((L.np1jp2[o] - L.nm1jp2[o]) * ph.exp2);
    const hpb_c2 t3 = (-4.0 / c2dt2) * a00 + (2.0 * chi[o]) * a00;
    const hpb_c2 t4 = laser_lap(L.nm1j00, i, j, p);
    const hpb_c2 f = c2(-3.0 / cdtdz + 2.0 / c2dt2, 2.0 * ph.djn / cdt + 2.0 * p.k0 / cdt);
    return t1 + t2 + t3 - t4 + f * am1;
}

// right-hand side and real coefficient of the multigrid variant (MultiLaser::AdvanceSliceMG :536-592):
// chi enters the operator (do_avg_rhs) and the right-hand side carries chi * A once, or the fft
// variant's 2 chi A^n (MG_average_rhs = 0)
HPB_HD hpb_c2 laser_rhs_mg_cell(const LaserPlanes &L, const double *chi, int i, int j, const LaserAdvPar &p,
                                const LaserPhase &ph, int do_avg_rhs, double &acf_real)
{
    const long o = (long)j * p.nx + i;
    const double cdt = p.c * p.dt, cdtdz = p.c * p.dt * p.dz;
    const hpb_c2 a00 = L.n00j00[o];
    const double acoeff_real = p.step0 ? 6.0 / cdtdz : 3.0 / cdtdz + 2.0 / (p.c * p.c * p.dt * p.dt);
    acf_real = do_avg_rhs ? acoeff_real + chi[o] : acoeff_real;
    if (p.step0) {
        const hpb_c2 t1 = (8.0 / cdtdz) * ((L.n00jp1[o] - L.np1jp1[o]) * ph.exp1);
        const hpb_c2 t2 = (2.0 / cdtdz) * ((L.np1jp2[o] - L.n00jp2[o]) * ph.exp2);
        const hpb_c2 t4 = laser_lap(L.n00j00, i, j, p);
        const hpb_c2 f = c2(-6.0 / cdtdz, 4.0 * ph.djn / cdt + 4.0 * p.k0 / cdt);
        const hpb_c2 r = t1 + t2 - t4 + f * a00;
        return r + (do_avg_rhs ? chi[o] : 2.0 * chi[o]) * a00;
    }
    const double c2dt2 = p.c * p.c * p.dt * p.dt;
    const hpb_c2 am1 = L.nm1j00[o];
    const hpb_c2 t1 = (4.0 / cdtdz) * ((L.nm1jp1[o] - L.np1jp1[o]) * ph.exp1);
    const hpb_c2 t2 = (1.0 / cdtdz) * ((L.np1jp2[o] - L.nm1jp2[o]) * ph.exp2);
    const hpb_c2 t3 = (-4.0 / c2dt2) * a00;
    const hpb_c2 t4 = laser_lap(L.nm1j00, i, j, p);
    const hpb_c2 f = c2(-3.0 / cdtdz + 2.0 / c2dt2, 2.0 * ph.djn / cdt + 2.0 * p.k0 / cdt);
    const hpb_c2 r = t1 + t2 + t3 - t4 + f * am1;
    return r + (do_avg_rhs ? chi[o] * am1 : (2.0 * chi[o]) * a00);
}

// spectral solve (:765-795): -rhs_f / (kx^2 + ky^2 + acoeff), times 1 / (nx ny) for the unnormalised
// inverse transform
HPB_HD hpb_c2 laser_spectral_cell(hpb_c2 rhs_f, int i, int j, const LaserAdvPar &p, const LaserPhase &ph)
{
    const int imid = (p.nx + 1) / 2, jmid = (p.ny + 1) / 2;
    const double kx = i < imid ? p.dkx * i : p.dkx * (i - p.nx);
    const double ky = j < jmid ? p.dky * j : p.dky * (j - p.ny);
    const double cdt = p.c * p.dt, cdtdz = p.c * p.dt * p.dz;
    hpb_c2 ac;
    if (p.step0) ac = c2(6.0 / cdtdz, -4.0 * (p.k0 + ph.djn) / cdt);
    else ac = c2(3.0 / cdtdz + 2.0 / (p.c * p.c * p.dt * p.dt), -2.0 * (p.k0 + ph.djn) / cdt);
    const hpb_c2 den = c2(kx * kx + ky * ky + ac.re, ac.im);
    const double d2 = den.re * den.re + den.im * den.im;
    if (d2 == 0.) return c2(0., 0.);
    const hpb_c2 inv = c2(den.re / d2, -den.im / d2);
    const double nrm = -1.0 / ((double)p.nx * (double)p.ny);
    return nrm * (rhs_f * inv);
}

// InterpolateChi for coinciding grids (:334-407): chi of the field slice inside the field box shrunk by
// two guard widths, the initial chi elsewhere
HPB_HD double laser_chi_cell(const SliceView &a, int c_chi, const double *chi_initial, int i, int j, int nx,
                             int ny, int g, double dx, double dy, double x_off, double y_off, int order)
{
    if (i < g || i > nx - 1 - g || j < g || j > ny - 1 - g) return chi_initial[(long)j * nx + i];
    const double xmid = ((i * dx + x_off) - x_off) * (1.0 / dx), ymid = ((j * dy + y_off) - y_off) * (1.0 / dy);
    double wx[3], wy[3];
    const int i0 = laser_interp_shape(xmid, order, wx), j0 = laser_interp_shape(ymid, order, wy);
    const double *chi = a.comp(c_chi);
    double v = 0.;
    for (int iy = 0; iy <= order; ++iy)
        for (int ix = 0; ix <= order; ++ix) v += (wy[iy] * wx[ix]) * chi[a.idx(i0 + ix, j0 + iy)];
    return v;
}

// UpdateLaserAabs from a stored envelope slice (:214-291): |a|^2 at field cell (i, j) of the grown box
HPB_HD double laser_aabs_cell(const hpb_c2 *env, int i, int j, int nx, int ny, double dx, double dy,
                              double x_off, double y_off, int order)
{
    const double xmid = ((i * dx + x_off) - x_off) * (1.0 / dx), ymid = ((j * dy + y_off) - y_off) * (1.0 / dy);
    double wx[3], wy[3];
    const int i0 = laser_interp_shape(xmid, order, wx), j0 = laser_interp_shape(ymid, order, wy);
    double v = 0.;
    for (int iy = 0; iy <= order; ++iy)
        for (int ix = 0; ix <= order; ++ix) {
            const int cx = i0 + ix, cy = j0 + iy;
            if (cx >= 0 && cx <= nx - 1 && cy >= 0 && cy <= ny - 1) {
                const hpb_c2 e = env[(long)cy * nx + cx];
                v += (wy[iy] * wx[ix]) * (e.re * e.re + e.im * e.im);
            }
        }
    return v;
}
