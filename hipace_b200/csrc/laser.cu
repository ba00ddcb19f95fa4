// Laser envelope at time step 0 (first part): the analytic gaussian envelope of
// MultiLaser::InitLaserSlice (src/laser/MultiLaser.cpp:881-917) evaluated on the field grid (the
// default laser geometry of MakeLaserGeometry, :58-118) and MultiLaser::UpdateLaserAabs
// (:214-291): |a|^2 interpolated onto the grown field slice.  The envelope is never stored: every
// field cell evaluates the (<= 9) laser cells it interpolates from -- a few hundred flops per cell
// once per slice, against planes of HBM traffic for a stored complex slice.
// The envelope ADVANCE over time steps (AdvanceSliceMG, hpmg type 2) is not implemented.
#include "common.cuh"
#include <cuda/std/complex>

namespace {

using cplx = cuda::std::complex<double>;

struct LaserSet { int n; hpb_laser L[HPB_MAX_LASERS]; double k0; };

// :881-917, literally (note that the CEP enters the exponent as written there)
__device__ cplx laser_envelope(const LaserSet &ls, double xg, double yg, double zg)
{
    const cplx I(0., 1.);
    cplx env(0., 0.);
    for (int l = 0; l < ls.n; ++l) {
        const hpb_laser &L = ls.L[l];
        const double x = xg - L.position_mean[0], y = yg - L.position_mean[1], z = zg - L.position_mean[2];
        const double ang = L.propagation_angle_yz + (L.pft_yz - 1.5707963267948966);
        const double yp = cos(ang) * y - sin(ang) * z;
        const double zp = sin(ang) * y + cos(ang) * z;
        const cplx diffract = 1.0 + I * ((zp - L.focal_distance + L.position_mean[2] * cos(L.propagation_angle_yz))
                                         * 2.0 / (ls.k0 * L.w0 * L.w0));
        const cplx inv_complex_waist_2 = 1.0 / (L.w0 * L.w0 * diffract);
        const cplx prefactor = L.a0 / diffract;
        const cplx stcfactor = prefactor * exp(-(zp * zp / (L.L0 * L.L0)));
        const cplx exp_argument = -(x * x + yp * yp) * inv_complex_waist_2;
        env += stcfactor * cuda::std::exp(exp_argument)
               * cuda::std::exp(I * (yp * ls.k0 * L.propagation_angle_yz) + L.cep);
    }
    return env;
}

// compute_shape_factor<order>, ShapeFactors.H:40-117 (orders 0..2): weights, leftmost cell
__device__ __forceinline__ int interp_shape(double xmid, int order, double w[3])
{
    if (order == 0) { w[0] = 1.; w[1] = w[2] = 0.; return (int)floor(xmid + 0.5); }
    if (order == 1) {
        const double j = floor(xmid), f = xmid - j;
        w[0] = 1. - f; w[1] = f; w[2] = 0.;
        return (int)j;
    }
    return shape2(xmid, w);
}

template <int G>       // guard cells of the field slice
__global__ void __launch_bounds__(256)
k_laser_aabs(SliceView a, int c_aabs, LaserSet ls, int nx, int ny, double dx, double dy, double x_off,
             double y_off, int order, double z, double *abs_sum)
{
    hpb_pdl_prologue();
    __shared__ double red[256];
    const int i = blockIdx.x * blockDim.x + threadIdx.x - G;
    const int j = (int)blockIdx.y - G;
    double mag = 0.;
    if (i < nx + G) {
        // field cell -> position -> laser-grid coordinate (the two grids coincide)
        const double xmid = ((i * dx + x_off) - x_off) * (1.0 / dx);
        const double ymid = ((j * dy + y_off) - y_off) * (1.0 / dy);
        double wx[3], wy[3];
        const int i0 = interp_shape(xmid, order, wx), j0 = interp_shape(ymid, order, wy);
        double aabs = 0.;
        for (int iy = 0; iy <= order; ++iy) {
            for (int ix = 0; ix <= order; ++ix) {
                const int cx = i0 + ix, cy = j0 + iy;
                if (cx >= 0 && cx <= nx - 1 && cy >= 0 && cy <= ny - 1) {
                    const cplx e = laser_envelope(ls, cx * dx + x_off, cy * dy + y_off, z);
                    aabs += wx[ix] * wy[iy] * (e.real() * e.real() + e.imag() * e.imag());
                }
            }
        }
        a.comp(c_aabs)[a.idx(i, j)] = aabs;
        if (abs_sum && i >= 0 && i < nx && j >= 0 && j < ny)
            mag = cuda::std::abs(laser_envelope(ls, i * dx + x_off, j * dy + y_off, z));
    }
    if (!abs_sum) return;
    red[threadIdx.x] = mag;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0 && red[0] != 0.) atomicAdd(abs_sum, red[0]);
}

}  // namespace

extern "C" int hpb_laser_update_aabs(hpb_ctx *ctx, hpb_slice sl, int c_aabs, const hpb_laser *lasers,
                                     int nlasers, double lambda0, int interp_order, double z_slice,
                                     double *d_envelope_abs_sum)
{
    if (!ctx || c_aabs < 0 || !lasers || nlasers < 1 || lambda0 <= 0.) return HPB_ERR_ARG;
    if (nlasers > HPB_MAX_LASERS) { hpb_set_error("at most %d lasers", HPB_MAX_LASERS); return HPB_ERR_UNSUPPORTED; }
    if (interp_order < 0 || interp_order > 2) {
        hpb_set_error("lasers.interp_order = %d is not supported (0..2)", interp_order);
        return HPB_ERR_UNSUPPORTED;
    }
    const hpb_geom &g = ctx->g;
    LaserSet ls;
    ls.n = nlasers;
    for (int l = 0; l < nlasers; ++l) ls.L[l] = lasers[l];
    ls.k0 = 2.0 * 3.14159265358979323846 / lambda0;
    const int ng = -sl.lo_x;
    if (ng < 1 || ng > 3 || sl.lo_y != sl.lo_x) { hpb_set_error("slice with %d guard cells", ng); return HPB_ERR_ARG; }
    dim3 grid((g.nx + 2 * ng + 255) / 256, g.ny + 2 * ng);
#define HPB_AABS(G) hpb_launch(k_laser_aabs<G>, grid, 256, 0, ctx->stream, make_view(sl), c_aabs, ls, g.nx, \
                               g.ny, g.dx, g.dy, g.x_off, g.y_off, interp_order, z_slice, d_envelope_abs_sum)
    if (ng == 2) HPB_AABS(2); else if (ng == 1) HPB_AABS(1); else HPB_AABS(3);
#undef HPB_AABS
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}
