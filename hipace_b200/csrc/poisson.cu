// Dirichlet Poisson solver for one transverse slice:  laplace(phi) = rhs,  phi = 0 at the first
// guard-cell centre -- the exact inverse of the 5-point Laplacian that the reference computes as
// DST2D(DST2D(rhs) * eigenvalue) (src/fields/fft_poisson_solver/FFTPoissonSolverDirichletFast.cpp:
// 224-248, 286-328; identical maths in ...DirichletDirect.cpp:87-139).
//
// B200-native formulation (not the reference's 4 cuFFT calls + 5 helper kernels):
//   1. rows:    Rhat[j][k] = DST-I_x(rhs[j][:])            one CTA per PAIR of rows; the two real
//               rows ride in the real/imaginary lanes of ONE complex FFT of length N = nx+1 held
//               in shared memory (mixed-radix Stockham, any N: 1025 = 5*5*41, 1024 = 4^5, ...)
//   2. columns: for every x-mode k solve the constant-coefficient tridiagonal system in y
//               (phi[j-1] - 2 phi[j] + phi[j+1])/dy^2 + lambda_k phi[j] = Rhat[j][k]
//               lambda_k = -4 sin^2(pi (k+1) / (2 (nx+1))) / dx^2      (Thomas, pivots tabulated)
//   3. rows:    phi[j][:] = DST-I_x(phihat[j][:]) / (2 (nx+1))   written straight into the slice
// Two 1-D transform passes instead of four, no transposes, no dependence on ny+1 being smooth.
// The three solves of a slice (Psi, Ez, Bz) are batched into each launch (grid.y = batch).
#include "common.cuh"
#include <math.h>
#include <vector>

namespace {

constexpr int kFftThreads = 256;
constexpr int kMaxRad = 32;

struct FftPlan { int N; int nrad; int rad[kMaxRad]; };
struct OutPtrs { double *p[4]; };

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// One Stockham stage of radix r: every thread produces output elements
//   out[(j-k) r + k + b Ns] = sum_t in[j + t N/r] * w_N^{ t (k + b Ns) N/(Ns r) }
__device__ __forceinline__ void fft_stage_generic(const double2 *__restrict__ in,
                                                  double2 *__restrict__ out,
                                                  const double2 *__restrict__ root, int N, int Ns,
                                                  int r)
{
    const int Nr = N / r;
    const int blk = Ns * r;
    const int tw = N / blk;
    for (int e = threadIdx.x; e < N; e += blockDim.x) {
        const int q = e / blk, rem = e - q * blk;
        const int b = rem / Ns, k = rem - b * Ns;
        const int j = q * Ns + k;
        const int step = (k + b * Ns) * tw;
        double2 acc = in[j];
        int idx = step;
        for (int t = 1; t < r; ++t) {
            const double2 w = __ldg(&root[idx]);
            const double2 v = in[j + t * Nr];
            acc.x += v.x * w.x - v.y * w.y;
            acc.y += v.x * w.y + v.y * w.x;
            idx += step;
            if (idx >= N) idx -= N;
        }
        out[e] = acc;
    }
}

// radix-4 stage, one butterfly per thread iteration (4 inputs read once)
__device__ __forceinline__ void fft_stage_r4(const double2 *__restrict__ in,
                                             double2 *__restrict__ out,
                                             const double2 *__restrict__ root, int N, int Ns)
{
    const int Nr = N >> 2;
    const int tw = N / (Ns * 4);
    for (int j = threadIdx.x; j < Nr; j += blockDim.x) {
        const int k = j % Ns;
        double2 v0 = in[j], v1 = in[j + Nr], v2 = in[j + 2 * Nr], v3 = in[j + 3 * Nr];
        if (k) {
            v1 = cmul(v1, __ldg(&root[k * tw]));
            v2 = cmul(v2, __ldg(&root[2 * k * tw]));
            v3 = cmul(v3, __ldg(&root[3 * k * tw]));
        }
        // DFT-4 with w = -i
        const double2 s02 = make_double2(v0.x + v2.x, v0.y + v2.y);
        const double2 d02 = make_double2(v0.x - v2.x, v0.y - v2.y);
        const double2 s13 = make_double2(v1.x + v3.x, v1.y + v3.y);
        const double2 d13 = make_double2(v1.x - v3.x, v1.y - v3.y);
        const int o = (j - k) * 4 + k;
        out[o] = make_double2(s02.x + s13.x, s02.y + s13.y);
        out[o + Ns] = make_double2(d02.x + d13.y, d02.y - d13.x);       // d02 - i d13
        out[o + 2 * Ns] = make_double2(s02.x - s13.x, s02.y - s13.y);
        out[o + 3 * Ns] = make_double2(d02.x - d13.y, d02.y + d13.x);   // d02 + i d13
    }
}

// DST-I along x of two rows per CTA.
//   in : rows of length nx, row stride in_rs, batch stride in_bs
//   out: rows of length nx at out.p[batch] + row*out_rs (first element = cell 0)
__global__ void __launch_bounds__(kFftThreads)
k_dst_rows(const double *__restrict__ in, long in_rs, long in_bs, OutPtrs out, long out_rs,
           int nx, int ny, FftPlan plan, const double2 *__restrict__ root,
           const double *__restrict__ sinx, double scale)
{
    extern __shared__ double2 smem[];
    const int N = plan.N;
    double2 *buf0 = smem;
    double2 *buf1 = smem + N;
    __shared__ double2 partial[kFftThreads];

    const int ja = 2 * blockIdx.x;
    const int jb = ja + 1;
    const bool has_b = jb < ny;
    const double *row_a = in + (long)blockIdx.y * in_bs + (long)ja * in_rs;
    const double *row_b = row_a + in_rs;

    // stage a_j (a_0 = 0, a_j = x_{j-1})
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        double2 v = make_double2(0., 0.);
        if (j > 0) {
            v.x = row_a[j - 1];
            if (has_b) v.y = row_b[j - 1];
        }
        buf1[j] = v;
    }
    __syncthreads();
    // auxiliary sequence y_j = sin(pi j/N) (a_j + a_{N-j}) + (a_j - a_{N-j})/2
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        double2 y = make_double2(0., 0.);
        if (j > 0) {
            const double2 a = buf1[j], ar = buf1[N - j];
            const double s = __ldg(&sinx[j]);
            y.x = s * (a.x + ar.x) + 0.5 * (a.x - ar.x);
            y.y = s * (a.y + ar.y) + 0.5 * (a.y - ar.y);
        }
        buf0[j] = y;
    }
    __syncthreads();
    // complex FFT of length N (forward, e^{-2 pi i jk/N})
    double2 *src = buf0, *dst = buf1;
    int Ns = 1;
    for (int s = 0; s < plan.nrad; ++s) {
        const int r = plan.rad[s];
        if (r == 4) fft_stage_r4(src, dst, root, N, Ns);
        else fft_stage_generic(src, dst, root, N, Ns, r);
        __syncthreads();
        double2 *t = src; src = dst; dst = t;
        Ns *= r;
    }
    // split the two real transforms:  Ya = (Z_k + conj Z_{N-k})/2, Yb = (Z_k - conj Z_{N-k})/(2i)
    // F_{2k} = -Im Y_k ; F_{2k+1} = F_{2k-1} + Re Y_k, F_1 = Re Y_0 / 2
    // dst[k]      <- (Re Ya_k, Re Yb_k)   (to be prefix-summed),  k = 0..M-1
    // src reuse is not possible (still read) so evens go to registers -> written after the scan
    const int M = (N + 1) / 2;             // number of odd outputs F_1, F_3, ...  (2k+1 <= N-1)
    const int chunk = (M + blockDim.x - 1) / blockDim.x;
    const int k0 = threadIdx.x * chunk;
    double2 run = make_double2(0., 0.);
    for (int c = 0; c < chunk; ++c) {
        const int k = k0 + c;
        if (k < M) {
            const double2 Z = src[k];
            const double2 Zr = src[k == 0 ? 0 : N - k];
            double2 re = make_double2(0.5 * (Z.x + Zr.x), 0.5 * (Z.y + Zr.y));  // Re Ya, Re Yb
            if (k == 0) { re.x *= 0.5; re.y *= 0.5; }
            run.x += re.x;
            run.y += re.y;
            dst[k] = run;                  // chunk-local inclusive scan
        }
    }
    partial[threadIdx.x] = run;
    __syncthreads();
    // Hillis-Steele inclusive scan over the per-thread totals
    for (int off = 1; off < (int)blockDim.x; off <<= 1) {
        double2 add = make_double2(0., 0.);
        if ((int)threadIdx.x >= off) add = partial[threadIdx.x - off];
        __syncthreads();
        partial[threadIdx.x].x += add.x;
        partial[threadIdx.x].y += add.y;
        __syncthreads();
    }
    const double2 base = threadIdx.x == 0 ? make_double2(0., 0.) : partial[threadIdx.x - 1];
    double *out_a = out.p[blockIdx.y] + (long)ja * out_rs;
    double *out_b = out_a + out_rs;
    const double sc = 2.0 * scale;
    // odd F_{2k+1} -> output index m = 2k
    for (int c = 0; c < chunk; ++c) {
        const int k = k0 + c;
        if (k < M && 2 * k < nx) {
            const double2 v = dst[k];
            out_a[2 * k] = sc * (v.x + base.x);
            if (has_b) out_b[2 * k] = sc * (v.y + base.y);
        }
    }
    // even F_{2k} = -Im Y_k -> output index m = 2k-1,  k = 1..(N-1)/2
    for (int k = threadIdx.x + 1; 2 * k <= N - 1; k += blockDim.x) {
        const double2 Z = src[k];
        const double2 Zr = src[N - k];
        // Im Ya = (Z.y - Zr.y)/2 ; Yb = (Z - conj Zr)/(2i): Im Yb = -(Z.x - Zr.x)/2
        out_a[2 * k - 1] = sc * (-0.5 * (Z.y - Zr.y));
        if (has_b) out_b[2 * k - 1] = sc * (0.5 * (Z.x - Zr.x));
    }
}

// Thomas solve along y for every x-mode; in place on spec[batch][j][k]
__global__ void __launch_bounds__(128)
k_tridiag_y(double *__restrict__ spec, const double *__restrict__ tm,
            const double *__restrict__ tc, int nx, int ny, double a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nx) return;
    double *d = spec + (long)blockIdx.y * nx * ny + k;
    const double *m = tm + k, *c = tc + k;
    constexpr int U = 8;
    double prev = 0.0;
    int j = 0;
    for (; j + U <= ny; j += U) {
        double rv[U], mv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { rv[u] = d[(long)(j + u) * nx]; mv[u] = __ldg(&m[(long)(j + u) * nx]); }
#pragma unroll
        for (int u = 0; u < U; ++u) { prev = (rv[u] - a * prev) * mv[u]; d[(long)(j + u) * nx] = prev; }
    }
    for (; j < ny; ++j) { prev = (d[(long)j * nx] - a * prev) * __ldg(&m[(long)j * nx]); d[(long)j * nx] = prev; }
    // back substitution: phi_j = d_j - c_j phi_{j+1}
    double phi = prev;     // j = ny-1
    j = ny - 2;
    for (; j - (U - 1) >= 0; j -= U) {
        double dv[U], cv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { dv[u] = d[(long)(j - u) * nx]; cv[u] = __ldg(&c[(long)(j - u) * nx]); }
#pragma unroll
        for (int u = 0; u < U; ++u) { phi = dv[u] - cv[u] * phi; d[(long)(j - u) * nx] = phi; }
    }
    for (; j >= 0; --j) { phi = d[(long)j * nx] - __ldg(&c[(long)j * nx]) * phi; d[(long)j * nx] = phi; }
}

void factorize(int N, FftPlan &plan)
{
    plan.N = N;
    plan.nrad = 0;
    int n = N;
    const int pref[] = {4, 2, 3, 5, 7};
    for (int r : pref)
        while (n % r == 0) { plan.rad[plan.nrad++] = r; n /= r; }
    for (int p = 11; n > 1; p += 2)
        while (n % p == 0) { plan.rad[plan.nrad++] = p; n /= p; }
}

}  // namespace

int hpb_launch_poisson_rhs(hpb_ctx *ctx, const hpb_slice &sl, const int *comps);
int hpb_launch_exmby_eypbx(hpb_ctx *ctx, const hpb_slice &sl, const int *comps);

int hpb_poisson_init(hpb_ctx *ctx)
{
    const hpb_geom &g = ctx->g;
    const int nx = g.nx, ny = g.ny, N = nx + 1;
    ctx->fftN = N;
    FftPlan plan;
    factorize(N, plan);
    ctx->nrad = plan.nrad;
    for (int i = 0; i < plan.nrad; ++i) ctx->radices[i] = plan.rad[i];
    std::vector<double2> root(N);
    std::vector<double> sinx(N);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int t = 0; t < N; ++t) {
        root[t].x = (double)cosl(-2.0L * pi * t / N);
        root[t].y = (double)sinl(-2.0L * pi * t / N);
        sinx[t] = (double)sinl(pi * t / N);
    }
    HPB_CUDA_CHECK(cudaMalloc(&ctx->d_root, sizeof(double2) * N));
    HPB_CUDA_CHECK(cudaMalloc(&ctx->d_sinx, sizeof(double) * N));
    HPB_CUDA_CHECK(cudaMemcpy(ctx->d_root, root.data(), sizeof(double2) * N, cudaMemcpyHostToDevice));
    HPB_CUDA_CHECK(cudaMemcpy(ctx->d_sinx, sinx.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
    // Thomas pivots: b_k = -2/dy^2 + lambda_k, a = 1/dy^2
    const double a = 1.0 / (g.dy * g.dy);
    std::vector<double> tm((size_t)nx * ny), tc((size_t)nx * ny);
    for (int k = 0; k < nx; ++k) {
        const long double s = sinl(pi * (k + 1) / (2.0L * (nx + 1)));
        const double b = (double)(-2.0L / ((long double)g.dy * g.dy)
                                  - 4.0L * s * s / ((long double)g.dx * g.dx));
        double cp = 0.0;
        for (int j = 0; j < ny; ++j) {
            const double m = 1.0 / (b - a * cp);
            cp = a * m;
            tm[(size_t)j * nx + k] = m;
            tc[(size_t)j * nx + k] = cp;
        }
    }
    const size_t bytes = sizeof(double) * (size_t)nx * ny;
    HPB_CUDA_CHECK(cudaMalloc(&ctx->d_tri_m, bytes));
    HPB_CUDA_CHECK(cudaMalloc(&ctx->d_tri_c, bytes));
    HPB_CUDA_CHECK(cudaMemcpy(ctx->d_tri_m, tm.data(), bytes, cudaMemcpyHostToDevice));
    HPB_CUDA_CHECK(cudaMemcpy(ctx->d_tri_c, tc.data(), bytes, cudaMemcpyHostToDevice));
    HPB_CUDA_CHECK(cudaMalloc(&ctx->d_spec, 3 * bytes));
    HPB_CUDA_CHECK(cudaMalloc(&ctx->d_stage, 3 * bytes));
    const size_t smem = 2 * sizeof(double2) * (size_t)N;
    if (smem > 200 * 1024) {
        hpb_set_error("poisson: nx = %d too large for the shared-memory row FFT", nx);
        return HPB_ERR_UNSUPPORTED;
    }
    HPB_CUDA_CHECK(cudaFuncSetAttribute(k_dst_rows, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    return HPB_OK;
}

void hpb_poisson_free(hpb_ctx *ctx)
{
    cudaFree(ctx->d_root); cudaFree(ctx->d_sinx); cudaFree(ctx->d_tri_m); cudaFree(ctx->d_tri_c);
    cudaFree(ctx->d_spec); cudaFree(ctx->d_stage);
}

extern "C" int hpb_poisson_solve(hpb_ctx *ctx, const double *d_rhs, hpb_slice sl, const int *c_lhs,
                                 int nbatch)
{
    if (!ctx || !d_rhs || !c_lhs || nbatch < 1 || nbatch > 3) return HPB_ERR_ARG;
    const hpb_geom &g = ctx->g;
    const int nx = g.nx, ny = g.ny, N = ctx->fftN;
    FftPlan plan;
    plan.N = N; plan.nrad = ctx->nrad;
    for (int i = 0; i < plan.nrad; ++i) plan.rad[i] = ctx->radices[i];
    const size_t smem = 2 * sizeof(double2) * (size_t)N;
    const long plane = (long)nx * ny;
    dim3 grid((ny + 1) / 2, nbatch);
    OutPtrs o1;
    for (int b = 0; b < 4; ++b) o1.p[b] = ctx->d_spec + (b < nbatch ? b : 0) * plane;
    k_dst_rows<<<grid, kFftThreads, smem, ctx->stream>>>(d_rhs, nx, plane, o1, nx, nx, ny, plan,
                                                         ctx->d_root, ctx->d_sinx, 1.0);
    dim3 gridt((nx + 127) / 128, nbatch);
    k_tridiag_y<<<gridt, 128, 0, ctx->stream>>>(ctx->d_spec, ctx->d_tri_m, ctx->d_tri_c, nx, ny,
                                                1.0 / (g.dy * g.dy));
    SliceView v = make_view(sl);
    OutPtrs o2;
    for (int b = 0; b < 4; ++b) o2.p[b] = v.comp(c_lhs[b < nbatch ? b : 0]) + v.idx(0, 0);
    k_dst_rows<<<grid, kFftThreads, smem, ctx->stream>>>(ctx->d_spec, nx, plane, o2, sl.jstride,
                                                         nx, ny, plan, ctx->d_root, ctx->d_sinx,
                                                         1.0 / (2.0 * N));
    hpb_count_launch(ctx, 3);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_fields_solve_psi_ez_bz(hpb_ctx *ctx, hpb_slice sl, const int *comps)
{
    if (!ctx || !comps) return HPB_ERR_ARG;
    int rc = hpb_launch_poisson_rhs(ctx, sl, comps);
    if (rc) return rc;
    const int lhs[3] = {comps[HPB_C_PSI], comps[HPB_C_EZ], comps[HPB_C_BZ]};
    rc = hpb_poisson_solve(ctx, ctx->d_stage, sl, lhs, 3);
    if (rc) return rc;
    return hpb_launch_exmby_eypbx(ctx, sl, comps);
}
