// Laser envelope at time step 0 (first part): the analytic gaussian envelope of
// MultiLaser::InitLaserSlice (src/laser/MultiLaser.cpp:881-917) evaluated on the field grid (the
// default laser geometry of MakeLaserGeometry, :58-118) and MultiLaser::UpdateLaserAabs
// (:214-291): |a|^2 interpolated onto the grown field slice.  The envelope is never stored: every
// field cell evaluates the (<= 9) laser cells it interpolates from -- a few hundred flops per cell
// once per slice, against planes of HBM traffic for a stored complex slice.
// The envelope ADVANCE over time steps (second half of this file) stores the slices and solves with the
// fft solver (lasers.solver_type = fft, MultiLaser::AdvanceSliceFFT); hpmg type 2 is not implemented.
#include "common.cuh"
#include "laser_advance.cuh"

// fft2d.cu
struct hpb_fft2d;
int hpb_fft2d_create(hpb_fft2d **out, int nx, int ny);
void hpb_fft2d_destroy(hpb_fft2d *f);
int hpb_fft2d_exec(hpb_fft2d *f, hpb_ctx *ctx, const double2 *in, double2 *out, int dir);
#include <cuda/std/complex>
#include <vector>

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

// =================================================================================================
// Envelope ADVANCE over time steps, fft solver (second part)
//   MultiLaser::AdvanceSliceFFT  src/laser/MultiLaser.cpp:609-801
//   InterpolateChi :334-407, UpdateLaserAabs :214-291, ShiftLaserSlices :180-212,
//   the hand-over of A^{n+1}, A^n to the next time step: src/utils/MultiBuffer.cpp:840-851, 913-923
// Per-cell arithmetic: laser_advance.cuh (verified on the host against the oracle).  The 2-D complex
// FFT of the laser grid is our own shared-memory transform (fft2d.cu), not cuFFT.
// =================================================================================================
namespace {

constexpr int kLT = 256;

__global__ void __launch_bounds__(kLT)
k_laser_init_slice(LaserSet ls, hpb_c2 *__restrict__ dst, int nx, int ny, double dx, double dy, double x_off,
                   double y_off, double z)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nx) return;
    const cplx e = laser_envelope(ls, i * dx + x_off, j * dy + y_off, z);
    dst[(long)j * nx + i] = c2(e.real(), e.imag());
}

__global__ void __launch_bounds__(kLT)
k_laser_aabs_slice(SliceView a, int c_aabs, const hpb_c2 *__restrict__ env, int nx, int ny, int g, double dx,
                   double dy, double x_off, double y_off, int order)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x - g, j = (int)blockIdx.y - g;
    if (i >= nx + g) return;
    a.comp(c_aabs)[a.idx(i, j)] = laser_aabs_cell(env, i, j, nx, ny, dx, dy, x_off, y_off, order);
}

// sum |A| of the diagnostic (xyz: the valid box; xz: the y = mid-domain line, order-1 interpolation)
__global__ void __launch_bounds__(kLT)
k_laser_abs_sum(const hpb_c2 *__restrict__ env, int nx, int ny, int xz, double *out)
{
    hpb_pdl_prologue();
    __shared__ double sm[kLT];
    double acc = 0.;
    if (xz) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x)
            acc += laser_diag_xz_abs(env, i, nx, ny);
    } else {
        const long n = (long)nx * ny;
        for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long)gridDim.x * blockDim.x)
            acc += sqrt(env[q].re * env[q].re + env[q].im * env[q].im);
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int st = kLT / 2; st > 0; st >>= 1) {
        if (threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0 && sm[0] != 0.) atomicAdd(out, sm[0]);
}

__global__ void __launch_bounds__(kLT)
k_laser_chi(SliceView a, int c_chi, const double *__restrict__ chi_initial, double *__restrict__ chi_out, int nx,
            int ny, int g, double dx, double dy, double x_off, double y_off, int order)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nx) return;
    chi_out[(long)j * nx + i] = laser_chi_cell(a, c_chi, chi_initial, i, j, nx, ny, g, dx, dy, x_off, y_off, order);
}

// the three on-axis sums and the phase terms: one thread
__global__ void k_laser_phase(const hpb_c2 *n00j00, const hpb_c2 *n00jp1, const hpb_c2 *n00jp2, int nx, int ny,
                              double dz, int use_phase, LaserPhase *out)
{
    hpb_pdl_prologue();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    hpb_c2 h[3];
    laser_axis_sums(n00j00, n00jp1, n00jp2, nx, ny, h);
    *out = laser_phase(h[0], h[1], h[2], dz, use_phase);
}

__global__ void __launch_bounds__(kLT)
k_laser_rhs(LaserPlanes L, const double *__restrict__ chi, LaserAdvPar par, const LaserPhase *__restrict__ ph,
            hpb_c2 *__restrict__ rhs)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= par.nx) return;
    rhs[(long)j * par.nx + i] = laser_rhs_cell(L, chi, i, j, par, *ph);
}

// multigrid variant: planar right-hand side [2][ny][nx], real coefficient plane, and the initial guess
// (whatever np1j00 holds: the previous slice's solution, MultiLaser.cpp:598-606) in planar form
__global__ void __launch_bounds__(kLT)
k_laser_rhs_mg(LaserPlanes L, const double *__restrict__ chi, LaserAdvPar par, const LaserPhase *__restrict__ ph,
               int do_avg_rhs, const hpb_c2 *__restrict__ guess, double *__restrict__ rhs2, double *__restrict__ acf_r,
               double *__restrict__ sol2)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= par.nx) return;
    const long o = (long)j * par.nx + i, plane = (long)par.nx * par.ny;
    double ar;
    const hpb_c2 r = laser_rhs_mg_cell(L, chi, i, j, par, *ph, do_avg_rhs, ar);
    rhs2[o] = r.re; rhs2[plane + o] = r.im;
    acf_r[o] = ar;
    const hpb_c2 g = guess[o];
    sol2[o] = g.re; sol2[plane + o] = g.im;
}
__global__ void __launch_bounds__(kLT)
k_laser_from_planar(const double *__restrict__ sol2, hpb_c2 *__restrict__ dst, int nx, int ny)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nx) return;
    const long o = (long)j * nx + i, plane = (long)nx * ny;
    dst[o] = c2(sol2[o], sol2[plane + o]);
}

__global__ void __launch_bounds__(kLT)
k_laser_spectral(hpb_c2 *__restrict__ rhs_f, LaserAdvPar par, const LaserPhase *__restrict__ ph)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= par.nx) return;
    const long o = (long)j * par.nx + i;
    rhs_f[o] = laser_spectral_cell(rhs_f[o], i, j, par, *ph);
}

__device__ __forceinline__ void atomic_max_double(double *addr, double v)
{
    unsigned long long *a = (unsigned long long *)addr, old = *a;
    while (__longlong_as_double((long long)old) < v) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
// rec[k * stride]: k = 0 maximum of |a|^2, 1..5 sums, 6, 7 the on-axis sum (re, im)
__global__ void __launch_bounds__(kLT)
k_laser_insitu(const hpb_c2 *__restrict__ env, int nx, int ny, double dx, double dy, double x_off, double y_off,
               double *__restrict__ rec, long stride)
{
    hpb_pdl_prologue();
    __shared__ double red[kLT / 32][8];
    double acc[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
    const long n = (long)nx * ny;
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long)gridDim.x * blockDim.x) {
        const int j = (int)(q / nx), i = (int)(q - (long)j * nx);
        double t[8];
        insitu_laser_terms(env, i, j, nx, ny, dx, dy, x_off, y_off, t);
        acc[0] = fmax(acc[0], t[0]);
#pragma unroll
        for (int k = 1; k < 8; ++k) acc[k] += t[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double u = __shfl_xor_sync(0xffffffffu, v, o);
            v = k == 0 ? fmax(v, u) : v + u;
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const int k = threadIdx.x;
        double v = red[0][k];
        for (int w = 1; w < kLT / 32; ++w) v = k == 0 ? fmax(v, red[w][k]) : v + red[w][k];
        if (k == 0) atomic_max_double(rec, v);
        else if (v != 0.) atomicAdd(rec + k * stride, v);
    }
}

}  // namespace

// the nine complex work slices (time levels n-1, n, n+1 at slices j, j+1, j+2: MultiLaser.H:24-48), the
// per-step storage of A^n and A^{n-1} for every slice, and the solver scratch
struct hpb_laser_state {
    int nx = 0, ny = 0, nz = 0, interp_order = 1, use_phase = 1;
    LaserSet ls;
    hpb_c2 *work = nullptr;             // 9 planes
    hpb_c2 *w[9] = {};                  // nm1j00 nm1jp1 nm1jp2 n00j00 n00jp1 n00jp2 np1j00 np1jp1 np1jp2
    hpb_c2 *store[4] = {};              // this step's {A^n, A^{n-1}}, next step's {A^{n+1}, A^n}: nz planes each
    hpb_c2 *rhs = nullptr;
    double *chi = nullptr, *chi_initial = nullptr;
    LaserPhase *phase = nullptr;
    hpb_fft2d *fft = nullptr;           // 2-D complex transform of the laser grid (fft2d.cu)
    // lasers.solver_type = multigrid (the reference's default): hpmg type 2 (mg.cu: hpb_mg_solve2)
    int use_mg = 0, mg_avg_rhs = 1;
    double mg_tol_rel = 1e-4, mg_tol_abs = 0.;
    double *mg_rhs2 = nullptr, *mg_sol2 = nullptr, *mg_acf = nullptr;    // planar [2][ny][nx], [ny][nx]
    long mg_vcycles = 0;
    int pipelined = 0;                  // the stored slices travel between ranks (pipeline.cu)
};
enum { L_NM1J00 = 0, L_NM1JP1, L_NM1JP2, L_N00J00, L_N00JP1, L_N00JP2, L_NP1J00, L_NP1JP1, L_NP1JP2 };

extern "C" void hpb_laser_state_destroy(hpb_laser_state *st)
{
    if (!st) return;
    hpb_fft2d_destroy(st->fft);
    cudaFree(st->work); cudaFree(st->rhs); cudaFree(st->chi); cudaFree(st->chi_initial); cudaFree(st->phase);
    cudaFree(st->mg_rhs2); cudaFree(st->mg_sol2); cudaFree(st->mg_acf);
    for (auto p : st->store) cudaFree(p);
    delete st;
}

extern "C" int hpb_laser_state_create(hpb_laser_state **out, hpb_ctx *ctx, int nz, const hpb_laser *lasers,
                                      int nlasers, double lambda0, int interp_order, int use_phase)
{
    if (!out || !ctx || nz < 1 || !lasers || nlasers < 1 || nlasers > HPB_MAX_LASERS || lambda0 <= 0.
        || interp_order < 0 || interp_order > 2) return HPB_ERR_ARG;
    hpb_laser_state *st = new hpb_laser_state();
    st->nx = ctx->g.nx; st->ny = ctx->g.ny; st->nz = nz; st->interp_order = interp_order; st->use_phase = use_phase;
    st->ls.n = nlasers;
    for (int l = 0; l < nlasers; ++l) st->ls.L[l] = lasers[l];
    st->ls.k0 = 2.0 * 3.14159265358979323846 / lambda0;
    const size_t plane = (size_t)st->nx * st->ny;
    bool ok = cudaMalloc(&st->work, 9 * plane * sizeof(hpb_c2)) == cudaSuccess
        && cudaMalloc(&st->rhs, plane * sizeof(hpb_c2)) == cudaSuccess
        && cudaMalloc(&st->chi, plane * sizeof(double)) == cudaSuccess
        && cudaMalloc(&st->chi_initial, plane * sizeof(double)) == cudaSuccess
        && cudaMalloc(&st->phase, sizeof(LaserPhase)) == cudaSuccess;
    for (int k = 0; ok && k < 4; ++k)
        ok = cudaMalloc(&st->store[k], (size_t)nz * plane * sizeof(hpb_c2)) == cudaSuccess
             && cudaMemset(st->store[k], 0, (size_t)nz * plane * sizeof(hpb_c2)) == cudaSuccess;
    if (ok && hpb_fft2d_create(&st->fft, st->nx, st->ny) != HPB_OK) ok = false;
    if (!ok) {
        hpb_set_error("laser envelope advance: allocation or FFT plan failed (%zu bytes per slice plane, %d slices)",
                      plane * sizeof(hpb_c2), nz);
        hpb_laser_state_destroy(st);
        return HPB_ERR_CUDA;
    }
    for (int k = 0; k < 9; ++k) st->w[k] = st->work + k * plane;
    *out = st;
    return HPB_OK;
}

// start of a time step: empty work slices (ResetAllQuantities), the unperturbed chi on the laser grid
// (MultiLaser::SetInitialChi :293-332, computed by the caller from the plasma density profiles)
extern "C" int hpb_laser_begin_step(hpb_laser_state *st, hpb_ctx *ctx, const double *h_chi_initial)
{
    if (!st || !ctx || !h_chi_initial) return HPB_ERR_ARG;
    const size_t plane = (size_t)st->nx * st->ny;
    HPB_CUDA_CHECK(cudaMemsetAsync(st->work, 0, 9 * plane * sizeof(hpb_c2), ctx->stream));
    for (int k = 0; k < 9; ++k) st->w[k] = st->work + k * plane;
    HPB_CUDA_CHECK(cudaMemcpyAsync(st->chi_initial, h_chi_initial, plane * sizeof(double), cudaMemcpyHostToDevice,
                                   ctx->stream));
    return HPB_OK;
}

// the envelope of slice `islice` arrives (time step 0: the analytic pulse, InitLaserSlice; later: A^n
// and A^{n-1} stored by the previous step), |a|^2 goes onto the field slice, sum |A| into the checksum
extern "C" int hpb_laser_get_slice(hpb_laser_state *st, hpb_ctx *ctx, hpb_slice sl, int c_aabs, int islice,
                                   int step, double z_slice, double *d_env_abs_sum, int diag_xz)
{
    if (!st || !ctx || c_aabs < 0 || islice < 0 || islice >= st->nz) return HPB_ERR_ARG;
    const hpb_geom &g = ctx->g;
    const size_t plane = (size_t)st->nx * st->ny;
    const int ng = -sl.lo_x;
    dim3 grid((st->nx + kLT - 1) / kLT, st->ny);
    if (step == 0) {
        hpb_launch(k_laser_init_slice, grid, kLT, 0, ctx->stream, st->ls, st->w[L_N00J00], st->nx, st->ny, g.dx,
                   g.dy, g.x_off, g.y_off, z_slice);
        hpb_count_launch(ctx);
    } else {
        HPB_CUDA_CHECK(cudaMemcpyAsync(st->w[L_N00J00], st->store[0] + (size_t)islice * plane, plane * sizeof(hpb_c2),
                                       cudaMemcpyDeviceToDevice, ctx->stream));
        HPB_CUDA_CHECK(cudaMemcpyAsync(st->w[L_NM1J00], st->store[1] + (size_t)islice * plane, plane * sizeof(hpb_c2),
                                       cudaMemcpyDeviceToDevice, ctx->stream));
    }
    dim3 gridg((st->nx + 2 * ng + kLT - 1) / kLT, st->ny + 2 * ng);
    hpb_launch(k_laser_aabs_slice, gridg, kLT, 0, ctx->stream, make_view(sl), c_aabs,
               (const hpb_c2 *)st->w[L_N00J00], st->nx, st->ny, ng, g.dx, g.dy, g.x_off, g.y_off, st->interp_order);
    hpb_count_launch(ctx);
    if (d_env_abs_sum) {
        hpb_launch(k_laser_abs_sum, diag_xz ? 1u : 64u, kLT, 0, ctx->stream, (const hpb_c2 *)st->w[L_N00J00],
                   st->nx, st->ny, diag_xz, d_env_abs_sum);
        hpb_count_launch(ctx);
    }
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

// AdvanceSliceFFT once chi of the slice is known; A^{n+1}, A^n of the slice are kept for the next step
extern "C" int hpb_laser_advance_slice(hpb_laser_state *st, hpb_ctx *ctx, hpb_slice sl, int c_chi, int islice,
                                       double dt, int step, double prob_len_x, double prob_len_y)
{
    if (!st || !ctx || c_chi < 0 || islice < 0 || islice >= st->nz) return HPB_ERR_ARG;
    if (dt == 0.) return HPB_OK;                                      // MultiLaser.cpp:418
    const hpb_geom &g = ctx->g;
    const size_t plane = (size_t)st->nx * st->ny;
    const int ng = -sl.lo_x;
    dim3 grid((st->nx + kLT - 1) / kLT, st->ny);
    hpb_launch(k_laser_chi, grid, kLT, 0, ctx->stream, make_view(sl), c_chi, (const double *)st->chi_initial,
               st->chi, st->nx, st->ny, ng, g.dx, g.dy, g.x_off, g.y_off, st->interp_order);
    hpb_launch(k_laser_phase, 1, 32, 0, ctx->stream, (const hpb_c2 *)st->w[L_N00J00],
               (const hpb_c2 *)st->w[L_N00JP1], (const hpb_c2 *)st->w[L_N00JP2], st->nx, st->ny, g.dz,
               st->use_phase, st->phase);
    const LaserPlanes L = {st->w[L_NM1J00], st->w[L_NM1JP1], st->w[L_NM1JP2], st->w[L_N00J00], st->w[L_N00JP1],
                           st->w[L_N00JP2], st->w[L_NP1JP1], st->w[L_NP1JP2]};
    const double pi2 = 2.0 * 3.14159265358979323846;
    const LaserAdvPar par = {st->nx, st->ny, step == 0 ? 1 : 0, g.dx, g.dy, g.dz, g.c, dt, st->ls.k0,
                             pi2 / prob_len_x, pi2 / prob_len_y};
    if (st->use_mg) {
        // MultiLaser::AdvanceSliceMG (:429-607)
        if (!st->mg_rhs2) {
            HPB_CUDA_CHECK(cudaMalloc(&st->mg_rhs2, 2 * plane * sizeof(double)));
            HPB_CUDA_CHECK(cudaMalloc(&st->mg_sol2, 2 * plane * sizeof(double)));
            HPB_CUDA_CHECK(cudaMalloc(&st->mg_acf, plane * sizeof(double)));
        }
        hpb_launch(k_laser_rhs_mg, grid, kLT, 0, ctx->stream, L, (const double *)st->chi, par,
                   (const LaserPhase *)st->phase, st->mg_avg_rhs, (const hpb_c2 *)st->w[L_NP1J00], st->mg_rhs2,
                   st->mg_acf, st->mg_sol2);
        hpb_count_launch(ctx, 3);
        // acoeff_imag (:518-519) needs djn: one small read-back per slice (the V-cycle loop synchronises
        // on its convergence test anyway)
        LaserPhase hph;
        HPB_CUDA_CHECK(cudaMemcpyAsync(&hph, st->phase, sizeof(hph), cudaMemcpyDeviceToHost, ctx->stream));
        HPB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        const double acoeff_imag = (step == 0 ? -4.0 : -2.0) * (st->ls.k0 + hph.djn) / (g.c * dt);
        int iters = 0;
        int rcm = hpb_mg_solve2(ctx, st->mg_sol2, st->mg_rhs2, st->mg_acf, acoeff_imag, st->mg_tol_rel, st->mg_tol_abs,
                                200, &iters);
        if (rcm) return rcm;
        st->mg_vcycles += iters;
        hpb_launch(k_laser_from_planar, grid, kLT, 0, ctx->stream, (const double *)st->mg_sol2, st->w[L_NP1J00],
                   st->nx, st->ny);
        hpb_count_launch(ctx);
        HPB_CUDA_CHECK(cudaMemcpyAsync(st->store[2] + (size_t)islice * plane, st->w[L_NP1J00], plane * sizeof(hpb_c2),
                                       cudaMemcpyDeviceToDevice, ctx->stream));
        HPB_CUDA_CHECK(cudaMemcpyAsync(st->store[3] + (size_t)islice * plane, st->w[L_N00J00], plane * sizeof(hpb_c2),
                                       cudaMemcpyDeviceToDevice, ctx->stream));
        HPB_CUDA_CHECK(cudaGetLastError());
        return HPB_OK;
    }
    hpb_launch(k_laser_rhs, grid, kLT, 0, ctx->stream, L, (const double *)st->chi, par,
               (const LaserPhase *)st->phase, st->rhs);
    hpb_count_launch(ctx, 3);
    HPB_CUDA_CHECK(cudaGetLastError());
    int rc = hpb_fft2d_exec(st->fft, ctx, (const double2 *)st->rhs, (double2 *)st->rhs, -1);
    if (rc) return rc;
    hpb_launch(k_laser_spectral, grid, kLT, 0, ctx->stream, st->rhs, par, (const LaserPhase *)st->phase);
    hpb_count_launch(ctx);
    if ((rc = hpb_fft2d_exec(st->fft, ctx, (const double2 *)st->rhs, (double2 *)st->w[L_NP1J00], +1))) return rc;
    // MultiBuffer::pack_data: A^{n+1} and A^n of this slice go to the next time step
    HPB_CUDA_CHECK(cudaMemcpyAsync(st->store[2] + (size_t)islice * plane, st->w[L_NP1J00], plane * sizeof(hpb_c2),
                                   cudaMemcpyDeviceToDevice, ctx->stream));
    HPB_CUDA_CHECK(cudaMemcpyAsync(st->store[3] + (size_t)islice * plane, st->w[L_N00J00], plane * sizeof(hpb_c2),
                                   cudaMemcpyDeviceToDevice, ctx->stream));
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

// ShiftLaserSlices (:180-212): j+1 -> j+2, j -> j+1 on every time level (pointer rotation)
// lasers.solver_type / MG_tolerance_rel / MG_tolerance_abs / MG_average_rhs (MultiLaser.cpp:40-56)
extern "C" int hpb_laser_set_solver(hpb_laser_state *st, int use_multigrid, double tol_rel, double tol_abs,
                                    int average_rhs)
{
    if (!st) return HPB_ERR_ARG;
    st->use_mg = use_multigrid != 0; st->mg_tol_rel = tol_rel; st->mg_tol_abs = tol_abs; st->mg_avg_rhs = average_rhs != 0;
    return HPB_OK;
}
extern "C" long hpb_laser_mg_vcycles(hpb_laser_state *st) { return st ? st->mg_vcycles : -1; }

extern "C" int hpb_laser_shift_slices(hpb_laser_state *st)
{
    if (!st) return HPB_ERR_ARG;
    for (int lvl = 0; lvl < 3; ++lvl) {
        hpb_c2 *j00 = st->w[3 * lvl], *jp1 = st->w[3 * lvl + 1], *jp2 = st->w[3 * lvl + 2];
        st->w[3 * lvl + 2] = jp1; st->w[3 * lvl + 1] = j00; st->w[3 * lvl] = jp2;     // jp2's memory is reused for j00
    }
    return HPB_OK;
}

// end of a time step: what the advance stored becomes the next step's input
extern "C" int hpb_laser_end_step(hpb_laser_state *st)
{
    if (!st) return HPB_ERR_ARG;
    // single rank: this step's {A^{n+1}, A^n} are the next step's {A^n, A^{n-1}}.  In a multi-rank
    // pipeline the next owned step gets them slice by slice from the upstream rank instead (below).
    if (!st->pipelined) {
        std::swap(st->store[0], st->store[2]);
        std::swap(st->store[1], st->store[3]);
    }
    return HPB_OK;
}

// The laser part of the per-slice hand-over between time steps (MultiBuffer::pack_data /
// unpack_data, src/utils/MultiBuffer.cpp:444-490, 840-851, 913-923): A^{n+1} and A^n of slice
// `islice` leave from send[0..1], the upstream rank's pair lands in recv[0..1] (this step's A^n, A^{n-1}).
void hpb_laser_packet(hpb_laser_state *st, int islice, void *recv[2], void *send[2], size_t *bytes)
{
    const size_t plane = (size_t)st->nx * st->ny;
    recv[0] = st->store[0] + (size_t)islice * plane; recv[1] = st->store[1] + (size_t)islice * plane;
    send[0] = st->store[2] + (size_t)islice * plane; send[1] = st->store[3] + (size_t)islice * plane;
    *bytes = plane * sizeof(hpb_c2);
    st->pipelined = 1;
}

// MultiLaser::InSituComputeDiags (:923-1001) of the current slice's A^n: d_record[k * stride], k = 0..7
extern "C" int hpb_laser_insitu_slice(hpb_laser_state *st, hpb_ctx *ctx, double *d_record, long stride)
{
    if (!st || !ctx || !d_record || stride < 1) return HPB_ERR_ARG;
    const hpb_geom &g = ctx->g;
    hpb_launch(k_laser_insitu, 64, kLT, 0, ctx->stream, (const hpb_c2 *)st->w[L_N00J00], st->nx, st->ny, g.dx, g.dy,
               g.x_off, g.y_off, d_record, stride);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}
