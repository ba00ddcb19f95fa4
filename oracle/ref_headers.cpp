// C entry points around the reference's OWN header-only arithmetic, compiled from where the
// headers lie under /root/reference/src (never copied): ShapeFactors.H (every order 0..3 and
// derivative type 0..2), FieldGather.H (field gather and laser gather, every order), fields/OpenBoundary.H (the multipole
// expansion of the open field boundary), PushPlasmaParticles.H + DualNumbers.H (the momentum derivative for
// plain reals and for dual numbers).  Built by oracle/refhdr.py into oracle/_ref/ and used by
// tests/test_oracle_refheaders.py to pin the NumPy restatement to the reference itself.
// TEST INFRASTRUCTURE ONLY.
#include "AMReX_Config.H"
#include "particles/particles_utils/ShapeFactors.H"
#include "utils/DualNumbers.H"
#include "particles/pusher/PushPlasmaParticles.H"
#define HIPACE_GPUUTIL_H_   /* Array3 comes from the shim, see there */
#include "particles/particles_utils/FieldGather.H"
#include "fields/OpenBoundary.H"

namespace {
template <int ORDER>
void shape_all(long n, const double *xmid, double *s, long *cell)
{
    for (long p = 0; p < n; ++p) {
        double w[ORDER + 1];
        cell[p] = compute_shape_factor<ORDER>(w, xmid[p]);
        for (int k = 0; k <= ORDER; ++k) {
            s[k * n + p] = w[k];
            // the two "single" variants must agree with the array variant cell by cell
            auto a = compute_single_shape_factor<false, ORDER>(xmid[p], k);
            auto b = compute_single_shape_factor<true, ORDER>(xmid[p], k);
            if (a.cell != cell[p] + k || b.cell != cell[p] + k) cell[p] = -(1L << 40);
            s[(ORDER + 1 + k) * n + p] = a.factor;
            s[(2 * (ORDER + 1) + k) * n + p] = b.factor;
        }
    }
}
template <int DTYPE, int ORDER>
void dshape_all(long n, const double *xmid, double *s, double *ds, long *cell)
{
    for (long p = 0; p < n; ++p)
        for (int k = 0; k <= ORDER + DTYPE; ++k) {
            auto r = single_derivative_shape_factor<DTYPE, ORDER>(xmid[p], k);
            s[k * n + p] = r.factor;
            ds[k * n + p] = r.dx_factor;
            if (k == 0) cell[p] = r.cell;
            else if (r.cell != cell[p] + k) cell[p] = -(1L << 40);
        }
}
}

extern "C" int ref_shape(int order, long n, const double *xmid, double *s, long *cell)
{
    switch (order) {
    case 0: shape_all<0>(n, xmid, s, cell); return 0;
    case 1: shape_all<1>(n, xmid, s, cell); return 0;
    case 2: shape_all<2>(n, xmid, s, cell); return 0;
    case 3: shape_all<3>(n, xmid, s, cell); return 0;
    }
    return 1;
}

extern "C" int ref_dshape(int dtype, int order, long n, const double *xmid, double *s, double *ds,
                          long *cell)
{
#define HPB_CASE(D, O) if (dtype == D && order == O) { dshape_all<D, O>(n, xmid, s, ds, cell); return 0; }
    HPB_CASE(0, 1) HPB_CASE(0, 2) HPB_CASE(0, 3) HPB_CASE(0, 0)
    HPB_CASE(1, 0) HPB_CASE(1, 1) HPB_CASE(1, 2) HPB_CASE(1, 3)
    HPB_CASE(2, 0) HPB_CASE(2, 1) HPB_CASE(2, 2) HPB_CASE(2, 3)
#undef HPB_CASE
    return 1;
}

// in: 14 arrays of length n (ux, uy, psi_inv, ExmBy, EypBx, Ez, Bx_c, By_c, Bz, A, ADx, ADy) and
// two scalars; out: d[3][n] plain derivative
extern "C" void ref_momentum_push(long n, const double *const *in, double clight_inv, double qmc,
                                  double *out)
{
    for (long p = 0; p < n; ++p) {
        auto r = PlasmaMomentumPush<double>(in[0][p], in[1][p], in[2][p], in[3][p], in[4][p],
                                            in[5][p], in[6][p], in[7][p], in[8][p], in[9][p],
                                            in[10][p], in[11][p], clight_inv, qmc);
        out[p] = r.dz_ux; out[n + p] = r.dz_uy; out[2 * n + p] = r.dz_psi;
    }
}

// dual-number evaluation: eps[3][n] are the epsilon parts of (ux, uy, psi_inv); out: value[3][n]
// then epsilon[3][n]
extern "C" void ref_momentum_push_dual(long n, const double *const *in, const double *const *eps,
                                       double clight_inv, double qmc, double *out)
{
    for (long p = 0; p < n; ++p) {
        auto r = PlasmaMomentumPush<DualNumber>(
            DualNumber{in[0][p], eps[0][p]}, DualNumber{in[1][p], eps[1][p]},
            DualNumber{in[2][p], eps[2][p]}, in[3][p], in[4][p], in[5][p], in[6][p], in[7][p],
            in[8][p], in[9][p], in[10][p], in[11][p], clight_inv, qmc);
        out[p] = r.dz_ux.value; out[n + p] = r.dz_uy.value; out[2 * n + p] = r.dz_psi.value;
        out[3 * n + p] = r.dz_ux.epsilon; out[4 * n + p] = r.dz_uy.epsilon;
        out[5 * n + p] = r.dz_psi.epsilon;
    }
}

// planes: ncomp planes of (ny + 2g) x (nx + 2g) doubles, x fastest, cell (i, j) of the grown box
// [-g, n-1+g]; comps = {psi, ez, bx, by, bz}; out[6][n] = ExmBy EypBx Ez Bx By Bz
extern "C" int ref_gather(int order, long n, const double *xp, const double *yp,
                          const double *planes, int nx_tot, int ny_tot, int g, const int *comps,
                          double dx_inv, double dy_inv, double x_off, double y_off, double *out)
{
    Array3<const double> arr{planes, nx_tot, (long)nx_tot * ny_tot, (long)g + (long)g * nx_tot};
    if (order < 0 || order > 3) return 1;
    for (long p = 0; p < n; ++p) {
        double v[6] = {0, 0, 0, 0, 0, 0};
        doGatherShapeN(xp[p], yp[p], v[0], v[1], v[2], v[3], v[4], v[5], arr, comps[0], comps[1],
                       comps[2], comps[3], comps[4], dx_inv, dy_inv, x_off, y_off, order);
        for (int k = 0; k < 6; ++k) out[k * n + p] = v[k];
    }
    return 0;
}

namespace {
template <int ORDER>
void laser_gather_all(long n, const double *xp, const double *yp, const Array3<const double> &arr,
                      double dx_inv, double dy_inv, double x_off, double y_off, double *out)
{
    for (long p = 0; p < n; ++p) {
        double a = 0, ax = 0, ay = 0, a_only = 0;
        doLaserGatherShapeN<ORDER>(xp[p], yp[p], a, ax, ay, arr, 0, dx_inv, dy_inv, x_off, y_off);
        doLaserGatherShapeN<ORDER>(xp[p], yp[p], a_only, arr, 0, dx_inv, dy_inv, x_off, y_off);
        out[p] = a; out[n + p] = ax; out[2 * n + p] = ay; out[3 * n + p] = a_only;
    }
}
}

// out[4][n] = Aabssq, its x and y derivative, and Aabssq of the value-only overload
extern "C" int ref_laser_gather(int order, long n, const double *xp, const double *yp,
                                const double *plane, int nx_tot, int ny_tot, int g, double dx_inv,
                                double dy_inv, double x_off, double y_off, double *out)
{
    Array3<const double> arr{plane, nx_tot, (long)nx_tot * ny_tot, (long)g + (long)g * nx_tot};
    switch (order) {
    case 0: laser_gather_all<0>(n, xp, yp, arr, dx_inv, dy_inv, x_off, y_off, out); return 0;
    case 1: laser_gather_all<1>(n, xp, yp, arr, dx_inv, dy_inv, x_off, y_off, out); return 0;
    case 2: laser_gather_all<2>(n, xp, yp, arr, dx_inv, dy_inv, x_off, y_off, out); return 0;
    case 3: laser_gather_all<3>(n, xp, yp, arr, dx_inv, dy_inv, x_off, y_off, out); return 0;
    }
    return 1;
}

// Open boundary (fields/OpenBoundary.H): the 37 coefficients summed over n sources (s, x, y) and the
// field value they give at m points (xd, yd) -- all coordinates already scaled
namespace {
template <std::size_t... I>
void tuple_add(MultipoleTuple &acc, const MultipoleTuple &t, std::index_sequence<I...>)
{
    ((amrex::get<I>(acc) += amrex::get<I>(t)), ...);
}
}
extern "C" void ref_open_boundary(long n, const double *s, const double *x, const double *y, int monopole,
                                  long m, const double *xd, const double *yd, double *out)
{
    MultipoleTuple acc{};
    for (long k = 0; k < n; ++k) tuple_add(acc, GetMultipoleCoeffs(s[k], x[k], y[k]), std::make_index_sequence<37>{});
    if (!monopole) amrex::get<0>(acc) = 0.;
    for (long k = 0; k < m; ++k) out[k] = GetFieldMultipole(acc, xd[k], yd[k]);
}
