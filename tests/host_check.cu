// HOST harness for the __host__ __device__ arithmetic of the CUDA path (shapes.cuh, push_math.cuh,
// generic_order.cuh): the very functions the kernels call, run on the CPU with a plain `+=` in place
// of the fp64 RED.  Built into build/libhpb200_hostcheck.so by hipace_b200.build.build_host_check()
// and driven by tests/test_device_math_host.py against the reference's own headers
// (oracle/ref_headers.cpp) and the NumPy oracle.  This is a TEST artefact: it is not part of
// libhpb200.so and nothing in the product loads it.
#include "generic_order.cuh"
#include "insitu.cuh"
#include "pc_fields.cuh"
#include "laser_advance.cuh"

namespace {
struct HostAdd {
    __host__ __device__ void operator()(double *p, double v) const { *p += v; }
};
SliceView view(double *planes, int nx_tot, int ny_tot, int g)
{
    SliceView v;
    v.p = planes; v.lo_x = -g; v.lo_y = -g; v.jstride = nx_tot; v.nstride = (long)nx_tot * ny_tot;
    return v;
}
}

#define BY_ORDER(order, CALL)                                  \
    switch (order) {                                           \
    case 0: { constexpr int O = 0; CALL; } break;              \
    case 1: { constexpr int O = 1; CALL; } break;              \
    case 2: { constexpr int O = 2; CALL; } break;              \
    case 3: { constexpr int O = 3; CALL; } break;              \
    default: return 1;                                         \
    }

template <int O>
static void shape_all(long n, const double *xmid, double *s, long *cell)
{
    for (long p = 0; p < n; ++p) {
        double w[O + 1];
        cell[p] = hpb_shape<O>(xmid[p], w);
        for (int k = 0; k <= O; ++k) s[k * n + p] = w[k];
    }
}
extern "C" int hc_shape(int order, long n, const double *xmid, double *s, long *cell)
{
    BY_ORDER(order, shape_all<O>(n, xmid, s, cell))
    return 0;
}

template <int D, int O>
static void dshape_all(long n, const double *xmid, double *s, double *ds, long *cell)
{
    for (long p = 0; p < n; ++p) {
        double a[O + D + 1], b[O + D + 1];
        cell[p] = hpb_dshape<D, O>(xmid[p], a, b);
        for (int k = 0; k <= O + D; ++k) { s[k * n + p] = a[k]; ds[k * n + p] = b[k]; }
    }
}
extern "C" int hc_dshape(int dtype, int order, long n, const double *xmid, double *s, double *ds, long *cell)
{
    if (dtype == 0) { BY_ORDER(order, (dshape_all<0, O>(n, xmid, s, ds, cell))) }
    else if (dtype == 1) { BY_ORDER(order, (dshape_all<1, O>(n, xmid, s, ds, cell))) }
    else if (dtype == 2) { BY_ORDER(order, (dshape_all<2, O>(n, xmid, s, ds, cell))) }
    else return 1;
    return 0;
}

// in: ux uy psi_inv ExmBy EypBx Ez Bx_c By_c Bz A ADx ADy; eps: epsilon parts of ux uy psi_inv.
// out[9][n]: plain derivative (3), dual value (3), dual epsilon (3)
extern "C" void hc_momentum_push(long n, const double *const *in, const double *const *eps, int laser,
                                 double clight_inv, double qmc, double *out)
{
    for (long p = 0; p < n; ++p) {
        const PushFields f = {in[3][p], in[4][p], in[5][p], in[6][p], in[7][p], in[8][p]};
        const PushLaser las = {in[9][p], in[10][p], in[11][p]};
        double a, b, c;
        Dual da, db, dc;
        const Dual ux{in[0][p], eps[0][p]}, uy{in[1][p], eps[1][p]}, pi{in[2][p], eps[2][p]};
        if (laser) {
            momentum_push<double, true>(in[0][p], in[1][p], in[2][p], f, clight_inv, qmc, a, b, c, las);
            momentum_push<Dual, true>(ux, uy, pi, f, clight_inv, qmc, da, db, dc, las);
        } else {
            momentum_push<double, false>(in[0][p], in[1][p], in[2][p], f, clight_inv, qmc, a, b, c, las);
            momentum_push<Dual, false>(ux, uy, pi, f, clight_inv, qmc, da, db, dc, las);
        }
        out[p] = a; out[n + p] = b; out[2 * n + p] = c;
        out[3 * n + p] = da.v; out[4 * n + p] = db.v; out[5 * n + p] = dc.v;
        out[6 * n + p] = da.e; out[7 * n + p] = db.e; out[8 * n + p] = dc.e;
    }
}

struct HcGrid { int nx_tot, ny_tot, g; double x_off, y_off, dx_inv, dy_inv; };

extern "C" int hc_gather(int order, long n, const double *xp, const double *yp, double *planes,
                         const HcGrid *hg, const int *comps, double *out)
{
    const SliceView a = view(planes, hg->nx_tot, hg->ny_tot, hg->g);
    const GenGrid gr = {hg->x_off, hg->y_off, hg->dx_inv, hg->dy_inv};
    for (long p = 0; p < n; ++p) {
        GatheredFields f;
        BY_ORDER(order, (f = gen_gather<O>(a, comps[0], comps[1], comps[2], comps[3], comps[4], gr, xp[p], yp[p])))
        out[p] = f.ExmBy; out[n + p] = f.EypBx; out[2 * n + p] = f.Ez;
        out[3 * n + p] = f.Bx; out[4 * n + p] = f.By; out[5 * n + p] = f.Bz;
    }
    return 0;
}

extern "C" int hc_laser_gather(int order, long n, const double *xp, const double *yp, double *plane,
                               const HcGrid *hg, double *out)
{
    const SliceView a = view(plane, hg->nx_tot, hg->ny_tot, hg->g);
    const GenGrid gr = {hg->x_off, hg->y_off, hg->dx_inv, hg->dy_inv};
    for (long p = 0; p < n; ++p) {
        double A, Ax, Ay, A0, t0, t1;
        BY_ORDER(order, (gen_laser_gather<O, true>(a, 0, gr, xp[p], yp[p], A, Ax, Ay),
                         gen_laser_gather<O, false>(a, 0, gr, xp[p], yp[p], A0, t0, t1)))
        out[p] = A; out[n + p] = Ax; out[2 * n + p] = Ay; out[3 * n + p] = A0;
    }
    return 0;
}

// plasma SoA: r[11] in PlasmaIdx order (hpb200.h), valid[] in/out (0/1)
extern "C" long hc_deposit_current(int order, long n, double *const *r, unsigned char *valid,
                                   double *planes, const HcGrid *hg, const int *c6,
                                   const GenDepositPar *par)
{
    const SliceView a = view(planes, hg->nx_tot, hg->ny_tot, hg->g);
    const GenGrid gr = {hg->x_off, hg->y_off, hg->dx_inv, hg->dy_inv};
    long n_bad = 0;
    for (long p = 0; p < n; ++p) {
        if (!valid[p]) continue;
        bool ok = true;
        if (par->c_aabs >= 0) {
            BY_ORDER(order, (ok = gen_deposit_current<O, true>(a, c6, gr, *par, r[HPB_X][p], r[HPB_Y][p], r[HPB_W][p],
                                                               r[HPB_UX][p], r[HPB_UY][p], r[HPB_PSI][p], HostAdd())))
        } else {
            BY_ORDER(order, (ok = gen_deposit_current<O, false>(a, c6, gr, *par, r[HPB_X][p], r[HPB_Y][p], r[HPB_W][p],
                                                                r[HPB_UX][p], r[HPB_UY][p], r[HPB_PSI][p], HostAdd())))
        }
        if (!ok) { ++n_bad; r[HPB_W][p] = 0.; valid[p] = 0; }
    }
    return n_bad;
}

extern "C" int hc_beam_deposit(int order, long n, const double *const *b7, const unsigned char *valid,
                               double *planes, const HcGrid *hg, int c_jx, int c_jy, int c_jz,
                               double clightsq, double q_invvol)
{
    const SliceView a = view(planes, hg->nx_tot, hg->ny_tot, hg->g);
    const GenGrid gr = {hg->x_off, hg->y_off, hg->dx_inv, hg->dy_inv};
    for (long p = 0; p < n; ++p) {
        if (!valid[p]) continue;
        // b7 = x y z w ux uy uz
        BY_ORDER(order, (gen_beam_deposit<O>(a, c_jx, c_jy, c_jz, gr, clightsq, q_invvol, b7[0][p], b7[1][p],
                                             b7[3][p], b7[4][p], b7[5][p], b7[6][p], HostAdd())))
    }
    return 0;
}

template <int O, int D>
static void explicit_all(long n, double *const *r, const unsigned char *valid, const SliceView &a,
                         const GenGrid &gr, const GenExplicitPar &par)
{
    for (long p = 0; p < n; ++p) {
        if (!valid[p]) continue;
        if (par.c_aabs >= 0)
            gen_explicit_deposition<O, D, true>(a, gr, par, r[HPB_X][p], r[HPB_Y][p], r[HPB_W][p], r[HPB_UX][p],
                                                r[HPB_UY][p], r[HPB_PSI][p], HostAdd());
        else
            gen_explicit_deposition<O, D, false>(a, gr, par, r[HPB_X][p], r[HPB_Y][p], r[HPB_W][p], r[HPB_UX][p],
                                                 r[HPB_UY][p], r[HPB_PSI][p], HostAdd());
    }
}
extern "C" int hc_explicit_deposition(int order, int dtype, long n, double *const *r,
                                      const unsigned char *valid, double *planes, const HcGrid *hg,
                                      const GenExplicitPar *par)
{
    const SliceView a = view(planes, hg->nx_tot, hg->ny_tot, hg->g);
    const GenGrid gr = {hg->x_off, hg->y_off, hg->dx_inv, hg->dy_inv};
    if (order == 0 && dtype == 0) return 1;
    if (dtype == 0) { BY_ORDER(order, (explicit_all<(O == 0 ? 1 : O), 0>(n, r, valid, a, gr, *par))) }
    else if (dtype == 1) { BY_ORDER(order, (explicit_all<O, 1>(n, r, valid, a, gr, *par))) }
    else if (dtype == 2) { BY_ORDER(order, (explicit_all<O, 2>(n, r, valid, a, gr, *par))) }
    else return 1;
    return 0;
}

extern "C" int hc_advance_plasma(int order, long n, double *const *r, unsigned char *valid,
                                 double *planes, const HcGrid *hg, const GenPushPar *par)
{
    const SliceView a = view(planes, hg->nx_tot, hg->ny_tot, hg->g);
    const GenGrid gr = {hg->x_off, hg->y_off, hg->dx_inv, hg->dy_inv};
    for (long p = 0; p < n; ++p) {
        if (!valid[p]) continue;
        double st[5] = {r[HPB_X_PREV][p], r[HPB_Y_PREV][p], r[HPB_UX_HALF][p], r[HPB_UY_HALF][p], r[HPB_PSI_HALF][p]};
        double out[5] = {r[HPB_X][p], r[HPB_Y][p], r[HPB_UX][p], r[HPB_UY][p], r[HPB_PSI][p]};
        bool alive = true;
        if (par->c_aabs >= 0) { BY_ORDER(order, (alive = gen_advance_plasma<O, true>(a, gr, *par, st, out))) }
        else { BY_ORDER(order, (alive = gen_advance_plasma<O, false>(a, gr, *par, st, out))) }
        r[HPB_X][p] = out[0]; r[HPB_Y][p] = out[1]; r[HPB_UX][p] = out[2]; r[HPB_UY][p] = out[3]; r[HPB_PSI][p] = out[4];
        if (!par->temp_slice) {
            r[HPB_X_PREV][p] = st[0]; r[HPB_Y_PREV][p] = st[1]; r[HPB_UX_HALF][p] = st[2];
            r[HPB_UY_HALF][p] = st[3]; r[HPB_PSI_HALF][p] = st[4];
        }
        if (!alive) { r[HPB_W][p] = 0.; valid[p] = 0; }
    }
    return 0;
}

// the 23 raw in-situ sums of a beam slice (insitu.cuh), b7 = x y z w ux uy uz
extern "C" void hc_beam_insitu(long n, const double *const *b7, const unsigned char *valid, double clight_inv,
                               double radius_sq, double *out23)
{
    for (int k = 0; k < 23; ++k) out23[k] = 0.;
    for (long p = 0; p < n; ++p) {
        double t[23];
        if (insitu_beam_terms(valid[p] != 0, b7[0][p], b7[1][p], b7[2][p], b7[4][p], b7[5][p], b7[6][p],
                              b7[3][p], clight_inv, radius_sq, t))
            for (int k = 0; k < 23; ++k) out23[k] += t[k];
    }
}

// predictor-corrector right-hand sides into stage[2][ny][nx] (pc_fields.cuh)
extern "C" void hc_bxby_rhs(double *planes, const HcGrid *hg, const BxByRhsPar *par, int nx, int ny, double *stage)
{
    const SliceView a = view(planes, hg->nx_tot, hg->ny_tot, hg->g);
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            double rbx, rby;
            bxby_rhs_cell(a, *par, i, j, rbx, rby);
            stage[(long)j * nx + i] = rbx;
            stage[(long)nx * ny + (long)j * nx + i] = rby;
        }
}

// Open boundary on one staging plane in place: moments, then the edge values
extern "C" int hc_open_boundary(double *rhs, int nx, int ny, double dx, double dy, double lo_x, double hi_x,
                                double lo_y, double hi_y, int monopole)
{
    OpenBcPar p;
    if (!open_bc_par(nx, ny, dx, dy, lo_x, hi_x, lo_y, hi_y, p)) return 1;
    double M[kMultipoleN];
    for (int k = 0; k < kMultipoleN; ++k) M[k] = 0.;
    for (long c = 0; c < (long)nx * ny; ++c) {
        double t[kMultipoleN];
        if (open_bc_source(p, c, rhs[c], t))
            for (int k = 0; k < kMultipoleN; ++k) M[k] += t[k];
    }
    if (!monopole) M[0] = 0.;
    for (int e = 0; e < 2 * (nx + ny); ++e) {
        long cell;
        const double v = open_bc_edge(p, e, M, cell);
        rhs[cell] += v;
    }
    return 0;
}

// the 15 raw in-situ sums of a plasma species (insitu.cuh); r[11] in PlasmaIdx order
extern "C" void hc_plasma_insitu(long n, const double *const *r, const unsigned char *valid, double clight_inv,
                                 double radius_sq, double *out15)
{
    for (int k = 0; k < 15; ++k) out15[k] = 0.;
    for (long p = 0; p < n; ++p) {
        double t[15];
        if (insitu_plasma_terms(valid[p] != 0, r[HPB_X][p], r[HPB_Y][p], r[HPB_UX][p], r[HPB_UY][p], r[HPB_PSI][p],
                                r[HPB_W][p], clight_inv, radius_sq, t))
            for (int k = 0; k < 15; ++k) out15[k] += t[k];
    }
}

// the 10 raw field in-situ sums over the valid box; comps7 = ExmBy EypBx Ez Bx By Bz jz_beam
extern "C" void hc_field_insitu(double *planes, const HcGrid *hg, const int *comps7, int nx, int ny,
                                double clight, double *out10)
{
    const SliceView a = view(planes, hg->nx_tot, hg->ny_tot, hg->g);
    for (int k = 0; k < 10; ++k) out10[k] = 0.;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const long o = a.idx(i, j);
            double t[10];
            insitu_field_terms(a.comp(comps7[0])[o], a.comp(comps7[1])[o], a.comp(comps7[2])[o], a.comp(comps7[3])[o],
                               a.comp(comps7[4])[o], a.comp(comps7[5])[o], a.comp(comps7[6])[o], clight, t);
            for (int k = 0; k < 10; ++k) out10[k] += t[k];
        }
}

// ---- laser envelope advance (laser_advance.cuh) ------------------------------------------------
// planes8: nm1j00 nm1jp1 nm1jp2 n00j00 n00jp1 n00jp2 np1jp1 np1jp2 (complex, [ny][nx]); h3: the three
// on-axis sums (re, im pairs); out: rhs[ny][nx] complex; phase_out: exp1 (2), exp2 (2), djn
extern "C" void hc_laser_rhs(const hpb_c2 *const *planes8, const double *chi, const LaserAdvPar *par,
                             const double *h3, int use_phase, hpb_c2 *rhs, double *phase_out)
{
    const LaserPlanes L = {planes8[0], planes8[1], planes8[2], planes8[3], planes8[4], planes8[5], planes8[6],
                           planes8[7]};
    // h3 == NULL: the on-axis sums come from the device function too
    hpb_c2 h[3];
    if (h3) { h[0] = c2(h3[0], h3[1]); h[1] = c2(h3[2], h3[3]); h[2] = c2(h3[4], h3[5]); }
    else laser_axis_sums(planes8[3], planes8[4], planes8[5], par->nx, par->ny, h);
    const LaserPhase ph = laser_phase(h[0], h[1], h[2], par->dz, use_phase);
    phase_out[0] = ph.exp1.re; phase_out[1] = ph.exp1.im; phase_out[2] = ph.exp2.re; phase_out[3] = ph.exp2.im;
    phase_out[4] = ph.djn;
    for (int j = 0; j < par->ny; ++j)
        for (int i = 0; i < par->nx; ++i) rhs[(long)j * par->nx + i] = laser_rhs_cell(L, chi, i, j, *par, ph);
}
extern "C" void hc_laser_spectral(hpb_c2 *rhs_f, const LaserAdvPar *par, const double *phase5)
{
    LaserPhase ph;
    ph.exp1 = c2(phase5[0], phase5[1]); ph.exp2 = c2(phase5[2], phase5[3]); ph.djn = phase5[4];
    for (int j = 0; j < par->ny; ++j)
        for (int i = 0; i < par->nx; ++i) {
            const long o = (long)j * par->nx + i;
            rhs_f[o] = laser_spectral_cell(rhs_f[o], i, j, *par, ph);
        }
}
extern "C" void hc_laser_chi_aabs(double *planes, const HcGrid *hg, int c_chi, const double *chi_initial,
                                  const hpb_c2 *env, int nx, int ny, double dx, double dy, int order,
                                  double *chi_out, double *aabs_out)
{
    const SliceView a = view(planes, hg->nx_tot, hg->ny_tot, hg->g);
    const int g = hg->g;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i)
            chi_out[(long)j * nx + i] = laser_chi_cell(a, c_chi, chi_initial, i, j, nx, ny, g, dx, dy, hg->x_off,
                                                       hg->y_off, order);
    for (int j = -g; j < ny + g; ++j)
        for (int i = -g; i < nx + g; ++i)
            aabs_out[(long)(j + g) * hg->nx_tot + (i + g)] = laser_aabs_cell(env, i, j, nx, ny, dx, dy, hg->x_off,
                                                                             hg->y_off, order);
}

extern "C" double hc_laser_diag_xz_sum(const hpb_c2 *env, int nx, int ny)
{
    double acc = 0.;
    for (int i = 0; i < nx; ++i) acc += laser_diag_xz_abs(env, i, nx, ny);
    return acc;
}

// the 8 raw laser in-situ values of one stored slice (laser_advance.cuh)
extern "C" void hc_laser_insitu(const hpb_c2 *env, int nx, int ny, double dx, double dy, double x_off,
                                double y_off, double *out8)
{
    for (int k = 0; k < 8; ++k) out8[k] = 0.;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            double t[8];
            insitu_laser_terms(env, i, j, nx, ny, dx, dy, x_off, y_off, t);
            out8[0] = out8[0] > t[0] ? out8[0] : t[0];
            for (int k = 1; k < 8; ++k) out8[k] += t[k];
        }
}
