// Field kernels of the predictor-corrector Bx/By solver and of boundary.field = Open:
// right-hand sides into the Poisson staging area, the multipole moments of a right-hand side and the
// non-zero Dirichlet values folded into its edge cells, the relative B-field error, and the linear
// combinations of the (Bx, By) plane pairs.  Per-cell arithmetic: pc_fields.cuh.
//   Hipace::PredictorCorrectorLoopToSolveBxBy   src/Hipace.cpp:935-1031
//   Fields::SolvePoissonBxBy / InitialBfieldGuess / MixAndShiftBfields / ComputeRelBFieldError
//                                               src/fields/Fields.cpp:1008-1078, 1149-1286
//   Fields::SetBoundaryCondition, SetDirichletBoundaries   src/fields/Fields.cpp:617-738
#include "pc_fields.cuh"

namespace {
constexpr int kT = 256;

__global__ void __launch_bounds__(kT)
k_bxby_rhs(SliceView a, BxByRhsPar p, int nx, int ny, double *__restrict__ stage)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nx) return;
    double rbx, rby;
    bxby_rhs_cell(a, p, i, j, rbx, rby);
    const long o = (long)j * nx + i;
    stage[o] = rbx;
    stage[(long)nx * ny + o] = rby;
}

__global__ void __launch_bounds__(kT)
k_psi_ez_bz_rhs(SliceView a, PsiEzBzRhsPar p, int nx, int ny, double *__restrict__ stage)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nx) return;
    double r[3];
    psi_ez_bz_rhs_cell(a, p, i, j, r);
    const long o = (long)j * nx + i, plane = (long)nx * ny;
    stage[o] = r[0]; stage[plane + o] = r[1]; stage[2 * plane + o] = r[2];
}

// block-wide sum of NV values per thread -> NV atomics per block
template <int NV>
__device__ __forceinline__ void block_sum_to(double (&v)[NV], double *__restrict__ out)
{
    __shared__ double red[kT / 32][NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double x = 0.;
        for (int w = 0; w < kT / 32; ++w) x += red[w][threadIdx.x];
        if (x != 0.) atomicAdd(out + threadIdx.x, x);
    }
}

// moments of one staging plane over the cells inside the cut-off circle (Fields.cpp:713-727)
__global__ void __launch_bounds__(kT)
k_multipole_moments(const double *__restrict__ rhs, OpenBcPar p, double *__restrict__ M)
{
    hpb_pdl_prologue();
    double acc[kMultipoleN];
#pragma unroll
    for (int k = 0; k < kMultipoleN; ++k) acc[k] = 0.;
    const long n = (long)p.nx * p.ny;
    for (long c = (long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long)gridDim.x * blockDim.x) {
        double t[kMultipoleN];
        if (!open_bc_source(p, c, rhs[c], t)) continue;
#pragma unroll
        for (int k = 0; k < kMultipoleN; ++k) acc[k] += t[k];
    }
    block_sum_to<kMultipoleN>(acc, M);
}

// one thread per edge slot
__global__ void __launch_bounds__(kT)
k_open_edges(double *__restrict__ rhs, OpenBcPar p, const double *__restrict__ M, int monopole)
{
    hpb_pdl_prologue();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * (p.nx + p.ny)) return;
    double Mloc[kMultipoleN];
#pragma unroll
    for (int k = 0; k < kMultipoleN; ++k) Mloc[k] = M[k];
    if (!monopole) Mloc[0] = 0.;            // Ez, Bz: no monopole (:729-733)
    long cell;
    const double v = open_bc_edge(p, e, Mloc, cell);
    atomicAdd(rhs + cell, v);               // the corners receive two contributions
}

// ComputeRelBFieldError (:1227-1286): out[0] += sum |B_a|, out[1] += sum |B_a - B_b| (valid box)
__global__ void __launch_bounds__(kT)
k_rel_b_error(SliceView a, int ca, int cb, int nx, int ny, double *__restrict__ out)
{
    hpb_pdl_prologue();
    double acc[2] = {0., 0.};
    const long n = (long)nx * ny;
    for (long c = (long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long)gridDim.x * blockDim.x) {
        const int j = (int)(c / nx), i = (int)(c - (long)j * nx);
        const long o = a.idx(i, j);
        const double ax = a.comp(ca)[o], ay = a.comp(ca + 1)[o];
        const double bx = a.comp(cb)[o], by = a.comp(cb + 1)[o];
        acc[0] += sqrt(ax * ax + ay * ay);
        acc[1] += sqrt((ax - bx) * (ax - bx) + (ay - by) * (ay - by));
    }
    block_sum_to<2>(acc, out);
}

// MultiFab::LinComb on two adjacent components over the grown box: dst = fa * A + fb * B
__global__ void __launch_bounds__(kT)
k_lincomb2(SliceView a, int c_dst, double fa, int c_a, double fb, int c_b, long ntot)
{
    hpb_pdl_prologue();
    const long o = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= ntot) return;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double va = fa == 0. ? 0. : fa * a.comp(c_a + k)[o];
        const double vb = fb == 0. ? 0. : fb * a.comp(c_b + k)[o];
        a.comp(c_dst + k)[o] = va + vb;
    }
}
}  // namespace

extern "C" int hpb_fields_bxby_rhs(hpb_ctx *ctx, hpb_slice sl, const int *comps, double *d_stage)
{
    if (!ctx || !comps || !d_stage) return HPB_ERR_ARG;
    const hpb_geom &g = ctx->g;
    const BxByRhsPar p = {comps[HPB_C_JZ], comps[HPB_C_PREV_JX], comps[HPB_C_PREV_JY], comps[HPB_C_NEXT_JX],
                          comps[HPB_C_NEXT_JY], g.mu0, 0.5 * (1.0 / g.dx), 0.5 * (1.0 / g.dy), 0.5 * (1.0 / g.dz)};
    if (p.c_jz < 0 || p.c_prev_jx < 0 || p.c_next_jx < 0) return HPB_ERR_ARG;
    dim3 grid((g.nx + kT - 1) / kT, g.ny);
    hpb_launch(k_bxby_rhs, grid, kT, 0, ctx->stream, make_view(sl), p, g.nx, g.ny, d_stage);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_fields_psi_ez_bz_rhs(hpb_ctx *ctx, hpb_slice sl, const int *comps, double *d_stage)
{
    if (!ctx || !comps || !d_stage) return HPB_ERR_ARG;
    const hpb_geom &g = ctx->g;
    const PsiEzBzRhsPar p = {comps[HPB_C_RHOMJZ], comps[HPB_C_JX], comps[HPB_C_JY], -1.0 / g.ep0,
                             1.0 / (g.ep0 * g.c), g.mu0, 0.5 * (1.0 / g.dx), 0.5 * (1.0 / g.dy)};
    dim3 grid((g.nx + kT - 1) / kT, g.ny);
    hpb_launch(k_psi_ez_bz_rhs, grid, kT, 0, ctx->stream, make_view(sl), p, g.nx, g.ny, d_stage);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

// d_rhs: one nx*ny staging plane, modified in place; d_moments: 38 doubles of scratch
extern "C" int hpb_fields_open_boundary(hpb_ctx *ctx, double *d_rhs, int monopole, double prob_lo_x,
                                        double prob_hi_x, double prob_lo_y, double prob_hi_y,
                                        double *d_moments)
{
    if (!ctx || !d_rhs || !d_moments) return HPB_ERR_ARG;
    const hpb_geom &g = ctx->g;
    OpenBcPar p;
    if (!open_bc_par(g.nx, g.ny, g.dx, g.dy, prob_lo_x, prob_hi_x, prob_lo_y, prob_hi_y, p)) {
        hpb_set_error("open boundaries: x = 0, y = 0 must be inside the box (the point of expansion)");
        return HPB_ERR_ARG;
    }
    HPB_CUDA_CHECK(cudaMemsetAsync(d_moments, 0, sizeof(double) * kMultipoleN, ctx->stream));
    unsigned nb = (unsigned)(((long)g.nx * g.ny + kT - 1) / kT);
    if (nb > 296) nb = 296;
    hpb_launch(k_multipole_moments, nb, kT, 0, ctx->stream, (const double *)d_rhs, p, d_moments);
    hpb_launch(k_open_edges, (unsigned)((2 * (g.nx + g.ny) + kT - 1) / kT), kT, 0, ctx->stream, d_rhs, p,
               (const double *)d_moments, monopole);
    hpb_count_launch(ctx, 2);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_fields_rel_b_error(hpb_ctx *ctx, hpb_slice sl, int c_bx_a, int c_bx_b, double *d_out2)
{
    if (!ctx || c_bx_a < 0 || c_bx_b < 0 || !d_out2) return HPB_ERR_ARG;
    const hpb_geom &g = ctx->g;
    HPB_CUDA_CHECK(cudaMemsetAsync(d_out2, 0, 2 * sizeof(double), ctx->stream));
    unsigned nb = (unsigned)(((long)g.nx * g.ny + kT - 1) / kT);
    if (nb > 296) nb = 296;
    hpb_launch(k_rel_b_error, nb, kT, 0, ctx->stream, make_view(sl), c_bx_a, c_bx_b, g.nx, g.ny, d_out2);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_fields_lincomb2(hpb_ctx *ctx, hpb_slice sl, int c_dst, double fa, int c_a, double fb,
                                   int c_b)
{
    if (!ctx || c_dst < 0 || c_a < 0 || c_b < 0) return HPB_ERR_ARG;
    const long ntot = (long)sl.jstride * sl.ny_tot;
    hpb_launch(k_lincomb2, (unsigned)((ntot + kT - 1) / kT), kT, 0, ctx->stream, make_view(sl), c_dst, fa, c_a, fb,
               c_b, ntot);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}
