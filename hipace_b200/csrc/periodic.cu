// boundary.field = Periodic and fields.poisson_solver = FFTPeriodic.
//   hpb_fields_enforce_periodic  <- Fields::EnforcePeriodic (src/fields/Fields.cpp:1117-1145): AMReX
//       SumBoundary / FillBoundary of the single slice box under the level-0 periodicity
//   hpb_poisson_solve_periodic   <- FFTPoissonSolverPeriodic::SolvePoissonEquation
//       (src/fields/fft_poisson_solver/FFTPoissonSolverPeriodic.cpp:111-149, inv_k2 of :69-92)
// The reference transforms one real staging area at a time (R2C / C2R, 2 FFT plans, 2 helper kernels per
// solve).  Here two right-hand sides ride in the Re / Im lanes of ONE complex 2-D transform (the filter
// -1/k^2 is real and even, so the two solutions separate exactly) on the shared-memory FFT of fft2d.cu:
// three solves = 2 forward + 2 inverse complex transforms and 3 helper kernels.
#include "common.cuh"
#include <math.h>

struct hpb_fft2d;
int hpb_fft2d_create(hpb_fft2d **out, int nx, int ny);
void hpb_fft2d_destroy(hpb_fft2d *f);
int hpb_fft2d_exec(hpb_fft2d *f, hpb_ctx *ctx, const double2 *in, double2 *out, int dir);

namespace {

struct CompList { int c[12]; int n; };

// SumBoundary: every valid cell within g cells of an edge receives the guard cells that are its periodic
// images -- (i +- nx, j), (i, j +- ny), (i +- nx, j +- ny) -- in this fixed order (AMReX adds the shifted
// copies in box-list order; the sum is the same up to the association of at most three additions).  The
// guard cells themselves keep their values (the destination of SumBoundary is the valid region only).
__global__ void k_periodic_sum(SliceView v, CompList cl, int nx, int ny, int g)
{
    hpb_pdl_prologue();
    // edge-cell enumeration: the 2g rows at the y edges are full rows; the remaining rows contribute
    // their 2g edge columns
    const long nfull = 2L * g * nx;
    const long nside = 2L * g * (ny - 2 * g);
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nfull + nside) return;
    int i, j;
    if (e < nfull) {
        const int r = (int)(e / nx);
        i = (int)(e - (long)r * nx);
        j = r < g ? r : ny - 2 * g + r;
    } else {
        const long q = e - nfull;
        const int r = (int)(q / (2 * g)), cidx = (int)(q - (long)r * 2 * g);
        j = g + r;
        i = cidx < g ? cidx : nx - 2 * g + cidx;
    }
    const int ii = i < g ? i + nx : (i >= nx - g ? i - nx : i);
    const int jj = j < g ? j + ny : (j >= ny - g ? j - ny : j);
    for (int k = 0; k < cl.n; ++k) {
        double *p = v.comp(cl.c[k]);
        double a = p[v.idx(i, j)];
        if (ii != i) a += p[v.idx(ii, j)];
        if (jj != j) a += p[v.idx(i, jj)];
        if (ii != i && jj != j) a += p[v.idx(ii, jj)];
        p[v.idx(i, j)] = a;
    }
}

// FillBoundary: every guard cell takes the value of its periodic image in the valid region
__global__ void k_periodic_fill(SliceView v, CompList cl, int nx, int ny, int g)
{
    hpb_pdl_prologue();
    const int nxt = nx + 2 * g;
    const long nfull = 2L * g * nxt;                  // the guard rows below and above
    const long nside = 2L * g * ny;                   // guard columns of the valid rows
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nfull + nside) return;
    int i, j;
    if (e < nfull) {
        const int r = (int)(e / nxt);
        i = (int)(e - (long)r * nxt) - g;
        j = r < g ? r - g : ny + (r - g);
    } else {
        const long q = e - nfull;
        const int r = (int)(q / (2 * g)), cidx = (int)(q - (long)r * 2 * g);
        j = r;
        i = cidx < g ? cidx - g : nx + (cidx - g);
    }
    const int ii = i < 0 ? i + nx : (i >= nx ? i - nx : i);
    const int jj = j < 0 ? j + ny : (j >= ny ? j - ny : j);
    for (int k = 0; k < cl.n; ++k) {
        double *p = v.comp(cl.c[k]);
        p[v.idx(i, j)] = p[v.idx(ii, jj)];
    }
}

// z[pair][idx] = (rhs[2 pair][idx], rhs[2 pair + 1][idx] or 0)
__global__ void k_per_pack(const double *__restrict__ rhs, long plane, int nbatch, double2 *__restrict__ z)
{
    hpb_pdl_prologue();
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= plane) return;
    const int pair = blockIdx.y;
    const double a = rhs[(2L * pair) * plane + idx];
    const double b = 2 * pair + 1 < nbatch ? rhs[(2L * pair + 1) * plane + idx] : 0.;
    z[pair * plane + idx] = make_double2(a, b);
}

// tmp_cmplx_arr(i,j) *= -inv_k2_arr(i,j)  (:119-131) on the full spectrum: the reference stores the
// half spectrum i = 0 .. nx/2 with kx = dkx i, ky = dky j (j < (ny+1)/2) or dky (j - ny), and sets
// inv_k2 = 0 wherever i == 0 OR j == 0 (:84-89, not only at the origin); the mirrored half i > nx/2 is
// the Hermitian partner (nx - i, (ny - j) mod ny), whose k^2 is that of |kx| = dkx (nx - i)
__global__ void k_per_filter(double2 *__restrict__ z, int nx, int ny, double dkx, double dky)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nx) return;
    const long o = (long)blockIdx.z * nx * ny + (long)j * nx + i;
    double f = 0.;
    if (i != 0 && j != 0) {
        const int ih = i <= nx / 2 ? i : nx - i;
        const int jh = i <= nx / 2 ? j : ny - j;                 // the stored partner's row
        const double kx = dkx * ih;
        const double ky = (jh < (ny + 1) / 2) ? dky * jh : dky * (jh - ny);
        f = -(1.0 / (kx * kx + ky * ky));
    }
    const double2 a = z[o];
    z[o] = make_double2(a.x * f, a.y * f);
}

// lhs_arr(i,j) = inv_N * tmp_real_arr(i,j)  (:136-148); pair 0 carries solves 0 (Re) and 1 (Im), pair 1
// carries solve 2 in its Re lane
__global__ void k_per_unpack(const double2 *__restrict__ z, SliceView v, int c0, int c1, int c2, int nx, int ny,
                             double inv_n)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nx) return;
    const double2 a = z[(long)blockIdx.z * nx * ny + (long)j * nx + i];
    if (blockIdx.z == 0) {
        v.comp(c0)[v.idx(i, j)] = inv_n * a.x;
        if (c1 >= 0) v.comp(c1)[v.idx(i, j)] = inv_n * a.y;
    } else {
        v.comp(c2)[v.idx(i, j)] = inv_n * a.x;
    }
}

struct PerState {
    hpb_fft2d *fft = nullptr;
    double2 *z = nullptr;       // 2 complex planes
};

}  // namespace

void hpb_periodic_free(hpb_ctx *ctx)
{
    PerState *st = (PerState *)ctx->periodic;
    if (!st) return;
    hpb_fft2d_destroy(st->fft);
    cudaFree(st->z);
    delete st;
    ctx->periodic = nullptr;
}

extern "C" int hpb_fields_enforce_periodic(hpb_ctx *ctx, hpb_slice sl, int do_sum, const int *comp_list, int n)
{
    if (!ctx || !sl.p || !comp_list || n < 1 || n > 12) return HPB_ERR_ARG;
    const int nx = ctx->g.nx, ny = ctx->g.ny, g = -sl.lo_x;
    if (g < 1 || sl.lo_y != sl.lo_x || nx < 2 * g || ny < 2 * g) {
        hpb_set_error("enforce_periodic: needs nx, ny >= 2 * guard width (%d)", g);
        return HPB_ERR_ARG;
    }
    CompList cl;
    cl.n = 0;
    for (int k = 0; k < n; ++k) {
        if (comp_list[k] < 0) continue;
        if (comp_list[k] >= sl.ncomp) return HPB_ERR_ARG;
        cl.c[cl.n++] = comp_list[k];
    }
    if (cl.n == 0) return HPB_OK;
    const SliceView v = make_view(sl);
    if (do_sum) {
        const long ne = 2L * g * nx + 2L * g * (ny - 2 * g);
        hpb_launch(k_periodic_sum, (unsigned)((ne + 255) / 256), 256, 0, ctx->stream, v, cl, nx, ny, g);
    } else {
        const long ne = 2L * g * (nx + 2 * g) + 2L * g * ny;
        hpb_launch(k_periodic_fill, (unsigned)((ne + 255) / 256), 256, 0, ctx->stream, v, cl, nx, ny, g);
    }
    hpb_count_launch(ctx, 1);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_poisson_solve_periodic(hpb_ctx *ctx, const double *d_rhs, hpb_slice sl, const int *c_lhs,
                                          int nbatch)
{
    if (!ctx || !d_rhs || !c_lhs || nbatch < 1 || nbatch > 3) return HPB_ERR_ARG;
    const int nx = ctx->g.nx, ny = ctx->g.ny;
    const long plane = (long)nx * ny;
    PerState *st = (PerState *)ctx->periodic;
    if (!st) {
        st = new PerState();
        int rc = hpb_fft2d_create(&st->fft, nx, ny);
        if (rc == HPB_OK && cudaMalloc(&st->z, sizeof(double2) * 2 * plane) != cudaSuccess) rc = HPB_ERR_CUDA;
        if (rc != HPB_OK) {
            hpb_fft2d_destroy(st->fft);
            delete st;
            return rc;
        }
        ctx->periodic = st;
    }
    const int npair = (nbatch + 1) / 2;
    const double pi = 3.14159265358979323846;
    const double dkx = 2 * pi / (nx * ctx->g.dx), dky = 2 * pi / (ny * ctx->g.dy);
    hpb_launch(k_per_pack, dim3((unsigned)((plane + 255) / 256), npair), 256, 0, ctx->stream, d_rhs, plane, nbatch, st->z);
    hpb_count_launch(ctx, 1);
    int rc;
    for (int p = 0; p < npair; ++p)
        if ((rc = hpb_fft2d_exec(st->fft, ctx, st->z + p * plane, st->z + p * plane, -1))) return rc;
    hpb_launch(k_per_filter, dim3((nx + 127) / 128, ny, npair), 128, 0, ctx->stream, st->z, nx, ny, dkx, dky);
    hpb_count_launch(ctx, 1);
    for (int p = 0; p < npair; ++p)
        if ((rc = hpb_fft2d_exec(st->fft, ctx, st->z + p * plane, st->z + p * plane, +1))) return rc;
    const SliceView v = make_view(sl);
    hpb_launch(k_per_unpack, dim3((nx + 127) / 128, ny, npair), 128, 0, ctx->stream, (const double2 *)st->z, v,
               c_lhs[0], nbatch > 1 ? c_lhs[1] : -1, nbatch > 2 ? c_lhs[2] : -1, nx, ny, 1.0 / (double)plane);
    hpb_count_launch(ctx, 1);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}
