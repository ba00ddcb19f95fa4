// MEASUREMENT ARM, not the product path: the reference's own GPU algorithm for the transverse Poisson
// solve, restated and compiled for sm_100a so that it can be timed on the same box beside ours
// ("the reference's GPU algorithm on B200" is the bar to beat).
//
//   FFTPoissonSolverDirichletFast::SolvePoissonEquation   src/fields/fft_poisson_solver/
//       FFTPoissonSolverDirichletFast.cpp:286-328: per right-hand side FOUR batched 1-D Z2D cuFFT
//       transforms (x, y, y, x; plans of WrapCuFFT.cpp:110-137, executed :227-232) and FIVE helper
//       kernels (ToComplex, ToSine_Transpose_ToComplex, ToSine_Mult_ToComplex,
//       ToSine_Transpose_ToComplex, ToSine, :32-190), eigenvalue table of :214-235.
//
// A DST-I of n reals is taken from a length-(n+1) complex-to-real FFT: with the odd extension
// e(-2) = -f(0), e(-1) = 0, e(n) = 0, e(n+1) = -f(n-1) the complex input is
// c(i) = e(2i) - e(2i-2) + i e(2i-1), i = 0 .. (n+1)/2, and the sine coefficients are recovered from
// the real output r as s(i) = (r(n-i) - r(i+1) + (r(i+1) + r(n-i)) / (2 sin(pi (i+1)/(n+1)))) / 2.
// Selected with hpb_set_option("poisson_impl", 1) (bench.py --impl cufft_ref / naive); cuFFT is bound
// with dlopen like in laser.cu.  Results agree with the product solver to round-off (tests).
#include "common.cuh"
#include <dlfcn.h>
#include <math.h>

int hpb_launch_exmby_eypbx(hpb_ctx *ctx, const hpb_slice &sl, const int *comps);
extern "C" int hpb_fields_psi_ez_bz_rhs(hpb_ctx *ctx, hpb_slice sl, const int *comps, double *d_stage);

namespace {

struct Cufft {
    void *h = nullptr;
    int (*plan1d)(int *, int, int, int) = nullptr;
    int (*exec_z2d)(int, void *, void *) = nullptr;
    int (*set_stream)(int, cudaStream_t) = nullptr;
    int (*destroy)(int) = nullptr;
};
Cufft *cufft()
{
    static Cufft api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        for (const char *name : {"libcufft.so.11", "libcufft.so.12", "libcufft.so"}) {
            api.h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (api.h) break;
        }
        if (api.h) {
            api.plan1d = (int (*)(int *, int, int, int))dlsym(api.h, "cufftPlan1d");
            api.exec_z2d = (int (*)(int, void *, void *))dlsym(api.h, "cufftExecZ2D");
            api.set_stream = (int (*)(int, cudaStream_t))dlsym(api.h, "cufftSetStream");
            api.destroy = (int (*)(int))dlsym(api.h, "cufftDestroy");
            if (!api.plan1d || !api.exec_z2d || !api.set_stream || !api.destroy) api.h = nullptr;
        }
    }
    return api.h ? &api : nullptr;
}
constexpr int kCufftZ2D = 0x6c;

struct RefArm {
    int nx = 0, ny = 0;
    int plan_x = -1, plan_y = -1;         // Z2D of length nx + 1 (batch ny) / ny + 1 (batch nx)
    double *d_real = nullptr;             // max((nx+1) ny, (ny+1) nx) reals: FFT output
    double2 *d_cplx = nullptr;            // max(((nx+1)/2+1) ny, ((ny+1)/2+1) nx): FFT input
    double *d_eig = nullptr;              // [nx][ny] eigenvalue table in the transposed layout
    double *d_sfx = nullptr, *d_sfy = nullptr;   // 1 / (2 sin(pi (i+1) / (n+1)))
    double *d_stage = nullptr;            // 3 right-hand sides
};

constexpr int kT = 256;

// the odd extension described above, read from a row of n reals
__device__ __forceinline__ double ext(const double *f, int k, int n)
{
    if (k >= 0 && k < n) return f[k];
    if (k == -2) return -f[0];
    if (k == n + 1) return -f[n - 1];
    return 0.;          // k == -1, k == n
}

// rows of n reals (stride ld) -> rows of n/2-ish + 1 complex inputs
__global__ void __launch_bounds__(kT)
k_ref_to_complex(const double *__restrict__ in, long ld, int n, int nbatch, double2 *__restrict__ out)
{
    const int nh = (n + 1) / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i > nh || b >= nbatch) return;
    const double *f = in + (long)b * ld;
    out[(long)b * (nh + 1) + i] = make_double2(ext(f, 2 * i, n) - ext(f, 2 * i - 2, n), ext(f, 2 * i - 1, n));
}

__device__ __forceinline__ double sine_of(const double *r, int i, int n, const double *sf)
{
    const double a = r[i + 1], b = r[n - i];
    return 0.5 * (b - a + (a + b) * sf[i]);
}

// FFT output rows r[b][0..n] (b < nbatch) -> sine coefficients S[b][i] -> transposed T[i][b] -> complex
// input rows of the transform along b (length nbatch): out[i][0 .. (nbatch+1)/2].  One 32 x 32 tile of
// T per block through shared memory so that both the reads (along i) and the writes (along b) are
// contiguous.  MULT: the sine coefficients are multiplied by eig[b][i] first (the spectral solve) and
// NOT transposed (out[b][...] along i) -- the reference's ToSine_Mult_ToComplex.
__global__ void __launch_bounds__(kT)
k_ref_sine_transpose_complex(const double *__restrict__ r, int n, int nbatch, const double *__restrict__ sf,
                             double2 *__restrict__ out)
{
    __shared__ double tile[32][67];                // [i local][b local - 2 .. b local + 64]
    const int i0 = blockIdx.x * 32, c0 = blockIdx.y * 32;        // 32 values of i, 32 complex outputs
    const int nh = (nbatch + 1) / 2;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 8 rows of threads
    // T[i][k] for k = 2 c0 - 2 .. 2 c0 + 63 (k is the batch index of the input)
    for (int kk = ty; kk < 66; kk += 8) {
        const int k = 2 * c0 - 2 + kk, i = i0 + tx;
        double v = 0.;
        if (i < n) {
            if (k >= 0 && k < nbatch) v = sine_of(r + (long)k * (n + 1), i, n, sf);
            else if (k == -2) v = -sine_of(r, i, n, sf);
            else if (k == nbatch + 1) v = -sine_of(r + (long)(nbatch - 1) * (n + 1), i, n, sf);
        }
        tile[tx][kk] = v;
    }
    __syncthreads();
    for (int ii = ty; ii < 32; ii += 8) {
        const int i = i0 + ii, c = c0 + tx;
        if (i < n && c <= nh) {
            const double *t = &tile[ii][2 * tx];                  // t[0] = T[2c-2], t[1] = T[2c-1], t[2] = T[2c]
            out[(long)i * (nh + 1) + c] = make_double2(t[2] - t[0], t[1]);
        }
    }
}

__global__ void __launch_bounds__(kT)
k_ref_sine_mult_complex(const double *__restrict__ r, int n, int nbatch, const double *__restrict__ sf,
                        const double *__restrict__ eig, double2 *__restrict__ out)
{
    const int nh = (n + 1) / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i > nh || b >= nbatch) return;
    const double *row = r + (long)b * (n + 1);
    const double *e = eig + (long)b * n;
    auto m = [&](int k) -> double {
        if (k >= 0 && k < n) return e[k] * sine_of(row, k, n, sf);
        if (k == -2) return -e[0] * sine_of(row, 0, n, sf);
        if (k == n + 1) return -e[n - 1] * sine_of(row, n - 1, n, sf);
        return 0.;
    };
    out[(long)b * (nh + 1) + i] = make_double2(m(2 * i) - m(2 * i - 2), m(2 * i - 1));
}

__global__ void __launch_bounds__(kT)
k_ref_to_sine(const double *__restrict__ r, int n, int nbatch, const double *__restrict__ sf, double *out, long ld)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i >= n || b >= nbatch) return;
    out[(long)b * ld + i] = sine_of(r + (long)b * (n + 1), i, n, sf);
}

__global__ void k_ref_tables(double *eig, double *sfx, double *sfy, int nx, int ny, double dx, double dy)
{
    const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const double pi = 3.14159265358979323846;
    if (q < (long)nx * ny) {
        // transposed layout: row = x mode i, column = y mode j (what ToSine_Mult_ToComplex sees)
        const int i = (int)(q / ny), j = (int)(q - (long)i * ny);
        const double sxs = sin((i + 1) * pi / (2. * (nx + 1))), sys = sin((j + 1) * pi / (2. * (ny + 1)));
        const double norm = 0.5 / (2. * ((double)(nx + 1) * (ny + 1)));
        eig[q] = norm / (-4.0 * (sxs * sxs / (dx * dx) + sys * sys / (dy * dy)));
    }
    if (q < nx) sfx[q] = 1.0 / (2.0 * sinpi((q + 1.0) / (nx + 1.0)));
    if (q < ny) sfy[q] = 1.0 / (2.0 * sinpi((q + 1.0) / (ny + 1.0)));
}

RefArm *arm_of(hpb_ctx *ctx)
{
    if (ctx->ref_arm) return (RefArm *)ctx->ref_arm;
    Cufft *F = cufft();
    if (!F) { hpb_set_error("reference arm: libcufft could not be loaded"); return nullptr; }
    RefArm *a = new RefArm();
    a->nx = ctx->g.nx; a->ny = ctx->g.ny;
    const int nx = a->nx, ny = a->ny;
    const size_t nreal = (size_t)std::max((long)(nx + 1) * ny, (long)(ny + 1) * nx);
    const size_t ncplx = (size_t)std::max((long)((nx + 1) / 2 + 1) * ny, (long)((ny + 1) / 2 + 1) * nx);
    bool ok = cudaMalloc(&a->d_real, nreal * sizeof(double)) == cudaSuccess
              && cudaMalloc(&a->d_cplx, ncplx * sizeof(double2)) == cudaSuccess
              && cudaMalloc(&a->d_eig, (size_t)nx * ny * sizeof(double)) == cudaSuccess
              && cudaMalloc(&a->d_sfx, nx * sizeof(double)) == cudaSuccess
              && cudaMalloc(&a->d_sfy, ny * sizeof(double)) == cudaSuccess
              && cudaMalloc(&a->d_stage, 3 * (size_t)nx * ny * sizeof(double)) == cudaSuccess;
    ok = ok && F->plan1d(&a->plan_x, nx + 1, kCufftZ2D, ny) == 0 && F->plan1d(&a->plan_y, ny + 1, kCufftZ2D, nx) == 0;
    if (!ok) { hpb_set_error("reference arm: allocation or cuFFT plan failed"); delete a; return nullptr; }
    F->set_stream(a->plan_x, ctx->stream);
    F->set_stream(a->plan_y, ctx->stream);
    k_ref_tables<<<(unsigned)(((long)nx * ny + 255) / 256), 256, 0, ctx->stream>>>(a->d_eig, a->d_sfx, a->d_sfy, nx, ny,
                                                                                  ctx->g.dx, ctx->g.dy);
    ctx->ref_arm = a;
    return a;
}

}  // namespace

void hpb_ref_arm_free(hpb_ctx *ctx)
{
    RefArm *a = (RefArm *)ctx->ref_arm;
    if (!a) return;
    if (Cufft *F = cufft()) { if (a->plan_x >= 0) F->destroy(a->plan_x); if (a->plan_y >= 0) F->destroy(a->plan_y); }
    cudaFree(a->d_real); cudaFree(a->d_cplx); cudaFree(a->d_eig); cudaFree(a->d_sfx); cudaFree(a->d_sfy);
    cudaFree(a->d_stage);
    delete a;
    ctx->ref_arm = nullptr;
}

// Psi, Ez, Bz of one slice the reference's way: staged right-hand sides, then per solve
// ToComplex, FFT x, ToSine+Transpose+ToComplex, FFT y, ToSine*eig+ToComplex, FFT y,
// ToSine+Transpose+ToComplex, FFT x, ToSine; finally ExmBy / EypBx.
int hpb_ref_arm_solve_psi_ez_bz(hpb_ctx *ctx, hpb_slice sl, const int *comps)
{
    RefArm *a = arm_of(ctx);
    Cufft *F = cufft();
    if (!a || !F) return HPB_ERR_UNSUPPORTED;
    const int nx = a->nx, ny = a->ny;
    int rc = hpb_fields_psi_ez_bz_rhs(ctx, sl, comps, a->d_stage);
    if (rc) return rc;
    SliceView v = make_view(sl);
    const int lhs[3] = {comps[HPB_C_PSI], comps[HPB_C_EZ], comps[HPB_C_BZ]};
    const int nhx = (nx + 1) / 2, nhy = (ny + 1) / 2;
    for (int k = 0; k < 3; ++k) {
        const double *rhs = a->d_stage + (size_t)k * nx * ny;
        k_ref_to_complex<<<dim3((nhx + kT) / kT, ny), kT, 0, ctx->stream>>>(rhs, nx, nx, ny, a->d_cplx);
        if (F->exec_z2d(a->plan_x, a->d_cplx, a->d_real)) return HPB_ERR_CUDA;
        k_ref_sine_transpose_complex<<<dim3((nx + 31) / 32, (nhy + 32) / 32), kT, 0, ctx->stream>>>(
            a->d_real, nx, ny, a->d_sfx, a->d_cplx);
        if (F->exec_z2d(a->plan_y, a->d_cplx, a->d_real)) return HPB_ERR_CUDA;
        k_ref_sine_mult_complex<<<dim3((nhy + kT) / kT, nx), kT, 0, ctx->stream>>>(a->d_real, ny, nx, a->d_sfy, a->d_eig,
                                                                                   a->d_cplx);
        if (F->exec_z2d(a->plan_y, a->d_cplx, a->d_real)) return HPB_ERR_CUDA;
        k_ref_sine_transpose_complex<<<dim3((ny + 31) / 32, (nhx + 32) / 32), kT, 0, ctx->stream>>>(
            a->d_real, ny, nx, a->d_sfy, a->d_cplx);
        if (F->exec_z2d(a->plan_x, a->d_cplx, a->d_real)) return HPB_ERR_CUDA;
        k_ref_to_sine<<<dim3((nx + kT - 1) / kT, ny), kT, 0, ctx->stream>>>(a->d_real, nx, ny, a->d_sfx,
                                                                           v.comp(lhs[k]) + v.idx(0, 0), sl.jstride);
    }
    hpb_count_launch(ctx, 3 * 9);
    HPB_CUDA_CHECK(cudaGetLastError());
    return hpb_launch_exmby_eypbx(ctx, sl, comps);
}
