// Dirichlet Poisson solver for one transverse slice:  laplace(phi) = rhs,  phi = 0 at the first
// guard-cell centre -- the exact inverse of the 5-point Laplacian that the reference computes as
// DST2D(DST2D(rhs) * eigenvalue) (src/fields/fft_poisson_solver/FFTPoissonSolverDirichletFast.cpp:
// 224-248, 286-328; identical maths in ...DirichletDirect.cpp:87-139).
//
// B200-native formulation (not the reference's 4 cuFFT calls + 5 helper kernels per solve):
//   K1 rows:   Rhat[j][k] = DST-I_x(rhs[j][:]).  One CTA per PAIR of rows: the two real rows ride
//              in the real / imaginary lanes of ONE complex FFT of length N = nx+1 held in shared
//              memory (Stockham, radix 4/2/3/5 + any odd prime factor: 1025 = 5*5*41), with the
//              pre-differencing / sine-factor post-processing of the (n+1)-point DST trick.  The
//              RHS assembly of Fields::SolvePoissonPsiExmByEypBxEzBz is fused into the row load.
//   K2,K3:     for every x-mode k the constant-coefficient tridiagonal system in y
//                (phi[j-1] - 2 phi[j] + phi[j+1])/dy^2 + lambda_k phi[j] = Rhat[j][k]
//              is solved by the partition method: independent Thomas solves on chunks of rows
//              (nx * nchunk * 3 threads, pivots tabulated), then a small reduced system for the
//              chunk-interface values; the interface correction is applied in the load of K4.
//   K4 rows:   phi[j][:] = DST-I_x(phihat[j][:]) / (2 (nx+1)), written straight into the slice.
// Two transform passes instead of four, no transposes, and the three solves of a slice
// (Psi, Ez, Bz) are batched into every launch (grid.y = batch).
#include "common.cuh"
#include "fft_smem.cuh"
#include <math.h>
#include <stdlib.h>
#include <vector>

namespace {

struct OutPtrs { double *p[4]; };

// ---- row sources ---------------------------------------------------------------------------
struct SrcStage {            // plain rows: in[b][j][i]
    const double *in; long rs, bs;
    __device__ __forceinline__ double get(int b, int j, int i) const { return in[(long)b * bs + (long)j * rs + i]; }
};
struct SrcFields {           // RHS of the three Poisson equations (Fields.cpp:886-912)
    SliceView a; int c_rhomjz, c_jx, c_jy;
    double f_psi, f_ez, mu0, dx_inv_half, dy_inv_half;
    __device__ __forceinline__ double get(int b, int j, int i) const
    {
        const long o = a.idx(i, j);
        if (b == 0) return f_psi * a.comp(c_rhomjz)[o];
        const double *jx = a.comp(c_jx), *jy = a.comp(c_jy);
        const long js = a.jstride;
        if (b == 1) {
            const double dx_jx = (jx[o + 1] - jx[o - 1]) * dx_inv_half;
            const double dy_jy = (jy[o + js] - jy[o - js]) * dy_inv_half;
            return f_ez * dx_jx + f_ez * dy_jy;
        }
        const double dy_jx = (jx[o + js] - jx[o - js]) * dy_inv_half;
        const double dx_jy = (jy[o + 1] - jy[o - 1]) * dx_inv_half;
        return mu0 * dy_jx + (-mu0) * dx_jy;
    }
};
struct SrcSpec {             // chunk-local Thomas solution + interface correction
    const double *y; long rs, bs;             // spec[b][j][k]
    const double *xl, *xr;                    // [b][chunk][k]
    const double *tp, *tq;                    // spike tables [table row][k]
    int L, C, nx, last_base;                  // table row of the last chunk's first row
    __device__ __forceinline__ double get(int b, int j, int i) const
    {
        const int c = j / L, r = j - c * L;
        const int tr = (c == C - 1) ? last_base + r : r;
        const long ci = ((long)b * C + c) * nx + i;
        return y[(long)b * bs + (long)j * rs + i] + xl[ci] * __ldg(&tp[(long)tr * nx + i])
               + xr[ci] * __ldg(&tq[(long)tr * nx + i]);
    }
};

// DST-I along x of two rows per CTA (unnormalised, FFTW RODFT00 convention, times `scale`).
template <class Src, int NTHR, int MINB, bool BLUE>
__global__ void __launch_bounds__(NTHR, MINB)
k_dst_rows(Src src, OutPtrs out, long out_rs, int nx, int ny, FftPlan plan,
           const double2 *__restrict__ root, const double *__restrict__ sinf, double scale)
{
    hpb_pdl_prologue();
    extern __shared__ double2 smem[];
    const int N = plan.N, n = nx;
    double2 *buf0 = smem;
    double2 *buf1 = smem + (BLUE ? plan.M + 4 : N);      // (N: the layout the radix plans were tuned with)
    const int b = blockIdx.y;
    const int ja = 2 * blockIdx.x, jb = ja + 1;
    const bool has_b = jb < ny;

    // stage the two real rows: xs[q] = (x_a[q], x_b[q]), q = 0..n-1, padded by two zero entries on
    // each side so that the differencing below needs no bounds checks
    double2 *xs = buf1 + 2;
    for (int q = threadIdx.x; q < n; q += blockDim.x)
        xs[q] = make_double2(src.get(b, ja, q), has_b ? src.get(b, jb, q) : 0.);
    if (threadIdx.x < 2) { buf1[threadIdx.x] = make_double2(0., 0.); }
    // N = n + 1 entries in buf1: indices n+2.. are beyond; the reads below touch xs[-2..n-1] only
    __syncthreads();
    // W[i] = (x[2i] - x[2i-2]) + i x[2i-1]  (x[-1] = x[-2]... = 0 handled explicitly), i = 0..nh;
    // V = W_a + i W_b with Hermitian extension; buf0 <- conj(V) so that the forward FFT below is
    // the backward (C2R) transform of the (n+1)-point DST trick
    const int nh = (n + 1) / 2;
    for (int i = threadIdx.x; i <= nh; i += blockDim.x) {
        double2 re, im;      // .x: row a, .y: row b
        if (i == 0) {
            re = make_double2(2. * xs[0].x, 2. * xs[0].y);
            im = make_double2(0., 0.);
        } else if (i == nh) {
            if (n & 1) {
                re = make_double2(-2. * xs[2 * i - 2].x, -2. * xs[2 * i - 2].y);
                im = make_double2(0., 0.);
            } else {
                re = make_double2(-xs[2 * i - 2].x, -xs[2 * i - 2].y);
                im = xs[2 * i - 1];
            }
        } else {
            re = csub(xs[2 * i], xs[2 * i - 2]);
            im = xs[2 * i - 1];
        }
        // W_a = re.x + i im.x, W_b = re.y + i im.y
        buf0[i] = make_double2(re.x - im.y, -(im.x + re.y));               // conj(V[i])
        if (i > 0 && N - i > nh) buf0[N - i] = make_double2(re.x + im.y, im.x - re.y);   // conj(V[N-i])
    }
    __syncthreads();
    // (the prime-stage table sits behind the two buffers: 2 (L + 4) complex elements)
    double *s_tab = reinterpret_cast<double *>(smem + 2 * (plan.buf_len() + 4));
    const double2 *F = fft_smem<BLUE>(buf0, buf1, plan, root, s_tab);
    // z_a = Re F, z_b = -Im F;  out[i] = 0.5 (z[n-i] - z[i+1] + (z[i+1] + z[n-i]) sinf[i])
    double *out_a = out.p[b] + (long)ja * out_rs;
    double *out_b = out_a + out_rs;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double2 za = F[i + 1], zb = F[n - i];
        const double sf = __ldg(&sinf[i]);
        out_a[i] = scale * (0.5 * (zb.x - za.x + (za.x + zb.x) * sf));
        if (has_b) out_b[i] = scale * (0.5 * (-zb.y + za.y - (za.y + zb.y) * sf));
    }
}

// K2: chunk-local Thomas solve along y for every x-mode; in place on spec[batch][j][k].
// Thread = (k, chunk, batch).  Writes the first / last local value of each chunk to yf / ye.
__global__ void __launch_bounds__(128)
k_thomas_local(double *__restrict__ spec, const double *__restrict__ tm, const double *__restrict__ tc,
               double *__restrict__ yf, double *__restrict__ ye, int nx, int ny, int L, int C,
               int last_base, double a)
{
    hpb_pdl_prologue();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nx) return;
    const int c = blockIdx.y, b = blockIdx.z;
    const int j0 = c * L, len = min(L, ny - j0);
    const int tr0 = (c == C - 1) ? last_base : 0;
    double *d = spec + (long)b * nx * ny + (long)j0 * nx + k;
    const double *m = tm + (long)tr0 * nx + k, *cc = tc + (long)tr0 * nx + k;
    constexpr int U = 8;
    double prev = 0.0;
    int r = 0;
    for (; r + U <= len; r += U) {
        double rv[U], mv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { rv[u] = d[(long)(r + u) * nx]; mv[u] = __ldg(&m[(long)(r + u) * nx]); }
#pragma unroll
        for (int u = 0; u < U; ++u) { prev = (rv[u] - a * prev) * mv[u]; d[(long)(r + u) * nx] = prev; }
    }
    for (; r < len; ++r) { prev = (d[(long)r * nx] - a * prev) * __ldg(&m[(long)r * nx]); d[(long)r * nx] = prev; }
    double phi = prev;
    const long ci = ((long)b * C + c) * nx + k;
    ye[ci] = phi;
    r = len - 2;
    for (; r - (U - 1) >= 0; r -= U) {
        double dv[U], cv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { dv[u] = d[(long)(r - u) * nx]; cv[u] = __ldg(&cc[(long)(r - u) * nx]); }
#pragma unroll
        for (int u = 0; u < U; ++u) { phi = dv[u] - cv[u] * phi; d[(long)(r - u) * nx] = phi; }
    }
    for (; r >= 0; --r) { phi = d[(long)r * nx] - __ldg(&cc[(long)r * nx]) * phi; d[(long)r * nx] = phi; }
    yf[ci] = phi;
}

// K3: reduced system for the chunk-interface values.  Thread = (k, batch).
//   in : yf[c] = first, ye[c] = last chunk-local value
//   out: xl[c] = value of the row just below chunk c (0 for c = 0), xr[c] = row just above (0 for last)
// (yf is overwritten by xl's data flow: xl/xr alias separate arrays)
__global__ void __launch_bounds__(128)
k_thomas_reduced(const double *__restrict__ yf, const double *__restrict__ ye, double *__restrict__ xl,
                 double *__restrict__ xr, const double *__restrict__ t_pe, const double *__restrict__ t_pf,
                 const double *__restrict__ t_b, const double *__restrict__ t_inv,
                 const double *__restrict__ t_del, int nx, int C)
{
    hpb_pdl_prologue();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nx) return;
    const int b = blockIdx.y;
    const long base = (long)b * C * nx + k;
    // Both recurrences are short (C - 1 <= 31 steps) but every step needs six values from global
    // memory: they are fetched U steps ahead (independent loads) so that the dependent chain
    // only sees arithmetic.
    constexpr int U = 8;
    // forward: A_c -> xl[c+1], gamma_c -> xr[c]
    double alpha = 0.0;
    for (int c0 = 0; c0 < C - 1; c0 += U) {
        double v_ye[U], v_yf[U], v_pe[U], v_pf[U], v_inv[U], v_b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + u;
            if (c < C - 1) {
                const long t = (long)c * nx + k;
                v_ye[u] = ye[base + (long)c * nx];
                v_yf[u] = yf[base + (long)(c + 1) * nx];
                v_pe[u] = __ldg(&t_pe[t]); v_pf[u] = __ldg(&t_pf[t]);
                v_inv[u] = __ldg(&t_inv[t]); v_b[u] = __ldg(&t_b[t]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + u;
            if (c < C - 1) {
                const double A = v_ye[u] + v_pe[u] * alpha;
                const double gam = (v_yf[u] + v_pf[u] * A) * v_inv[u];
                alpha = A + v_b[u] * gam;
                xl[base + (long)(c + 1) * nx] = A;
                xr[base + (long)c * nx] = gam;
            }
        }
    }
    // backward: v_c = gamma_c + delta_c v_{c+1};  u_c = A_c + B_c v_c
    double v = 0.0;
    xr[base + (long)(C - 1) * nx] = 0.0;
    xl[base] = 0.0;
    for (int c0 = C - 2; c0 >= 0; c0 -= U) {
        double v_g[U], v_a[U], v_del[U], v_b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 - u;
            if (c >= 0) {
                const long t = (long)c * nx + k;
                v_g[u] = xr[base + (long)c * nx];
                v_a[u] = xl[base + (long)(c + 1) * nx];
                v_del[u] = __ldg(&t_del[t]); v_b[u] = __ldg(&t_b[t]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 - u;
            if (c >= 0) {
                v = v_g[u] + v_del[u] * v;
                xr[base + (long)c * nx] = v;
                xl[base + (long)(c + 1) * nx] = v_a[u] + v_b[u] * v;
            }
        }
    }
}

void factorize(int N, FftPlan &plan)
{
    plan.N = N;
    plan.M = fft_factorize(N, plan.rad, plan.nrad);      // M > 0: Bluestein, the radices factor M
    plan.chirp = plan.bhat = nullptr;
}

template <class T>
int upload(T **dptr, const std::vector<T> &h)
{
    HPB_CUDA_CHECK(cudaMalloc(dptr, sizeof(T) * h.size()));
    HPB_CUDA_CHECK(cudaMemcpy(*dptr, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    return HPB_OK;
}

FftPlan make_plan(const hpb_ctx *ctx)
{
    FftPlan plan;
    plan.N = ctx->fftN; plan.nrad = ctx->nrad;
    plan.M = ctx->fftM; plan.chirp = ctx->d_chirp; plan.bhat = ctx->d_bhat;
    for (int i = 0; i < plan.nrad; ++i) {
        plan.rad[i] = ctx->radices[i];
        plan.cs_cos[i] = ctx->d_cs_cos[i];
        plan.cs_sin[i] = ctx->d_cs_sin[i];
        // fft_variant 2: the scalar prime stage (the A/B partner of the tensor-core one)
        plan.cs_frag[i] = ctx->tune_fft_variant == 2 ? nullptr : ctx->d_cs_frag[i];
    }
    return plan;
}

template <class Src, int NTHR, int MINB>
int launch_rows_v(hpb_ctx *ctx, const Src &src, const OutPtrs &out, long out_rs, int nbatch, double scale)
{
    const int nx = ctx->g.nx, ny = ctx->g.ny;
    const int L = ctx->fftM > 0 ? ctx->fftM : ctx->fftN;
    const FftPlan pl_ = make_plan(ctx);
    const size_t smem = 2 * sizeof(double2) * (size_t)(L + 4) + 2 * sizeof(double) * (size_t)pl_.max_prime();
    static bool attr_set = false;       // per instantiation; the limit covers every supported N
    if (!attr_set) {
        HPB_CUDA_CHECK(cudaFuncSetAttribute(k_dst_rows<Src, NTHR, MINB, false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        HPB_CUDA_CHECK(cudaFuncSetAttribute(k_dst_rows<Src, NTHR, MINB, true>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    dim3 grid((ny + 1) / 2, nbatch);
    if (ctx->fftM > 0)
        hpb_launch(k_dst_rows<Src, NTHR, MINB, true>, grid, NTHR, smem, ctx->stream, src, out, out_rs, nx, ny,
                   make_plan(ctx), ctx->d_root, ctx->d_sinf, scale);
    else
        hpb_launch(k_dst_rows<Src, NTHR, MINB, false>, grid, NTHR, smem, ctx->stream, src, out, out_rs, nx, ny,
                   make_plan(ctx), ctx->d_root, ctx->d_sinf, scale);
    hpb_count_launch(ctx);
    return HPB_OK;
}

template <class Src>
int launch_rows(hpb_ctx *ctx, const Src &src, const OutPtrs &out, long out_rs, int nbatch, double scale)
{
    const int variant = ctx->tune_fft_variant;
    // 5 CTAs / SM (48 registers, no spills) since the prime stage runs on the tensor cores: Poisson stage
    // 0.138 -> 0.129 ms (profiles/r02/r02F_tune.txt).  With the scalar prime stage the kernel was bound by the
    // shared-memory pipe and 6 CTAs / SM (40 registers, two waves of 148 x 6 for the 1536 row pairs of a
    // 1024^2 batch) measured best: fft_variant 1 keeps that geometry, 2 also the scalar stage.
    // long Bluestein transforms (M >= 2048: 2 x 64 KB and more of shared memory, one CTA per SM): 1024 threads
    // per row pair instead of 256 -- the only parallelism an SM has then is inside its one CTA
    // (2048^2, M = 4096: the configs[4] shape; fft_variant 3 keeps 256 threads for the A/B)
    if (ctx->fftM >= 2048 && variant != 3) return launch_rows_v<Src, 1024, 1>(ctx, src, out, out_rs, nbatch, scale);
    if (variant == 1 || variant == 2) return launch_rows_v<Src, kFftThreads, 6>(ctx, src, out, out_rs, nbatch, scale);
    return launch_rows_v<Src, kFftThreads, 5>(ctx, src, out, out_rs, nbatch, scale);
}

}  // namespace

int hpb_launch_exmby_eypbx(hpb_ctx *ctx, const hpb_slice &sl, const int *comps);

extern "C" long hpb_debug_fft_prime_table(int p, double *out, long out_len)
{
    if (p < 3 || (p & 1) == 0) return -1;
    const std::vector<double> t = fft_prime_frag_table(p);
    if (out) for (long k = 0; k < out_len && k < (long)t.size(); ++k) out[k] = t[k];
    return (long)t.size();
}

int hpb_poisson_init(hpb_ctx *ctx)
{
    const hpb_geom &g = ctx->g;
    const int nx = g.nx, ny = g.ny, N = nx + 1;
    ctx->fftN = N;
    FftPlan plan;
    factorize(N, plan);
    ctx->nrad = plan.nrad;
    ctx->fftM = plan.M;
    const int Lfft = plan.M > 0 ? plan.M : N;            // length of the root table / of the stages
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < plan.nrad; ++i) {
        ctx->radices[i] = plan.rad[i];
        ctx->d_cs_cos[i] = ctx->d_cs_sin[i] = ctx->d_cs_frag[i] = nullptr;
        const int p = plan.rad[i];
        if (p > 5) {
            const int h = (p - 1) / 2;
            std::vector<double> tc((size_t)(h + 1) * h), ts((size_t)(h + 1) * h);
            for (int b = 0; b <= h; ++b)
                for (int t = 1; t <= h; ++t) {
                    const long tb = ((long)t * b) % p;
                    tc[(size_t)b * h + t - 1] = (double)cosl(2.0L * pi * tb / p);
                    ts[(size_t)b * h + t - 1] = (double)sinl(2.0L * pi * tb / p);
                }
            int rc = upload(&ctx->d_cs_cos[i], tc); if (rc) return rc;
            rc = upload(&ctx->d_cs_sin[i], ts); if (rc) return rc;
            rc = upload(&ctx->d_cs_frag[i], fft_prime_frag_table(p)); if (rc) return rc;
        }
    }
    std::vector<double2> root(Lfft);
    std::vector<double> sinf(nx);
    for (int t = 0; t < Lfft; ++t) {
        root[t].x = (double)cosl(-2.0L * pi * t / Lfft);
        root[t].y = (double)sinl(-2.0L * pi * t / Lfft);
    }
    if (plan.M > 0) {
        std::vector<double2> chirp(N), bhat(plan.M);
        fft_bluestein_tables(N, plan.M, chirp.data(), bhat.data());
        int rcb = upload(&ctx->d_chirp, chirp); if (rcb) return rcb;
        rcb = upload(&ctx->d_bhat, bhat); if (rcb) return rcb;
    }
    for (int i = 0; i < nx; ++i) sinf[i] = (double)(1.0L / (2.0L * sinl(pi * (i + 1) / N)));
    int rc = upload(&ctx->d_root, root); if (rc) return rc;
    rc = upload(&ctx->d_sinf, sinf); if (rc) return rc;

    // ---- partitioned Thomas tables ----------------------------------------------------------
    int L = (ny + 31) / 32;
    if (L < 4) L = 4;
    if (L > ny) L = ny;
    const int C = (ny + L - 1) / L;
    const int Llast = ny - (C - 1) * L;
    const bool two_types = (Llast != L);
    const int nrows = L + (two_types ? Llast : 0);
    ctx->th_L = L; ctx->th_C = C; ctx->th_last_base = two_types ? L : 0;
    const long double a = 1.0L / ((long double)g.dy * g.dy);
    std::vector<double> tm((size_t)nrows * nx), tc((size_t)nrows * nx), tp((size_t)nrows * nx), tq((size_t)nrows * nx);
    const int nif = C > 1 ? C - 1 : 1;
    std::vector<double> t_pe((size_t)nif * nx, 0.), t_pf((size_t)nif * nx, 0.), t_b((size_t)nif * nx, 0.),
        t_inv((size_t)nif * nx, 1.), t_del((size_t)nif * nx, 0.);
    std::vector<long double> m(L), cp(L), sp(L), sq(L);
    for (int k = 0; k < nx; ++k) {
        const long double s = sinl(pi * (k + 1) / (2.0L * (nx + 1)));
        const long double bdiag = -2.0L * a - 4.0L * s * s / ((long double)g.dx * g.dx);
        long double pE[2], qE[2], pF[2], qF[2];
        for (int type = 0; type < (two_types ? 2 : 1); ++type) {
            const int len = type == 0 ? L : Llast;
            const int base = type == 0 ? 0 : L;
            long double c_prev = 0.0L;
            for (int r = 0; r < len; ++r) {
                // pivots in double arithmetic order of the kernel: m = 1/(b - a c'), c' = a m
                m[r] = 1.0L / (bdiag - a * c_prev);
                cp[r] = a * m[r];
                c_prev = cp[r];
                tm[(size_t)(base + r) * nx + k] = (double)m[r];
                tc[(size_t)(base + r) * nx + k] = (double)cp[r];
            }
            // spikes: p = T^{-1}(-a e_0), q = T^{-1}(-a e_{len-1})
            long double prev = 0.0L;
            for (int r = 0; r < len; ++r) { const long double d = (r == 0 ? -a : 0.0L); prev = (d - a * prev) * m[r]; sp[r] = prev; }
            for (int r = len - 2; r >= 0; --r) sp[r] = sp[r] - cp[r] * sp[r + 1];
            prev = 0.0L;
            for (int r = 0; r < len; ++r) { const long double d = (r == len - 1 ? -a : 0.0L); prev = (d - a * prev) * m[r]; sq[r] = prev; }
            for (int r = len - 2; r >= 0; --r) sq[r] = sq[r] - cp[r] * sq[r + 1];
            for (int r = 0; r < len; ++r) {
                tp[(size_t)(base + r) * nx + k] = (double)sp[r];
                tq[(size_t)(base + r) * nx + k] = (double)sq[r];
            }
            pF[type] = sp[0]; qF[type] = sq[0]; pE[type] = sp[len - 1]; qE[type] = sq[len - 1];
        }
        if (!two_types) { pE[1] = pE[0]; qE[1] = qE[0]; pF[1] = pF[0]; qF[1] = qF[0]; }
        long double beta = 0.0L;
        for (int c = 0; c < C - 1; ++c) {
            // interface between chunk c (always a full chunk) and chunk c+1 (last one may be short)
            const int tn = (c + 1 == C - 1) ? 1 : 0;
            const long double B = qE[0] + pE[0] * beta;
            const long double inv = 1.0L / (1.0L - pF[tn] * B);
            const long double del = qF[tn] * inv;
            beta = B * del;
            const size_t t = (size_t)c * nx + k;
            t_pe[t] = (double)pE[0]; t_pf[t] = (double)pF[tn]; t_b[t] = (double)B;
            t_inv[t] = (double)inv; t_del[t] = (double)del;
        }
    }
    if ((rc = upload(&ctx->d_tri_m, tm))) return rc;
    if ((rc = upload(&ctx->d_tri_c, tc))) return rc;
    if ((rc = upload(&ctx->d_tri_p, tp))) return rc;
    if ((rc = upload(&ctx->d_tri_q, tq))) return rc;
    if ((rc = upload(&ctx->d_red_pe, t_pe))) return rc;
    if ((rc = upload(&ctx->d_red_pf, t_pf))) return rc;
    if ((rc = upload(&ctx->d_red_b, t_b))) return rc;
    if ((rc = upload(&ctx->d_red_inv, t_inv))) return rc;
    if ((rc = upload(&ctx->d_red_del, t_del))) return rc;
    const size_t bytes = sizeof(double) * (size_t)nx * ny;
    HPB_CUDA_CHECK(cudaMalloc(&ctx->d_spec, 3 * bytes));
    HPB_CUDA_CHECK(cudaMalloc(&ctx->d_iface, 4 * 3 * sizeof(double) * (size_t)C * nx));
    HPB_CUDA_CHECK(cudaMemset(ctx->d_iface, 0, 4 * 3 * sizeof(double) * (size_t)C * nx));
    const size_t smem = 2 * sizeof(double2) * (size_t)(Lfft + 4);
    if (smem > 220 * 1024) {
        hpb_set_error("poisson: nx = %d too large for the shared-memory row FFT", nx);
        return HPB_ERR_UNSUPPORTED;
    }
    return HPB_OK;
}

void hpb_poisson_free(hpb_ctx *ctx)
{
    cudaFree(ctx->d_root); cudaFree(ctx->d_sinf); cudaFree(ctx->d_tri_m); cudaFree(ctx->d_tri_c);
    cudaFree(ctx->d_tri_p); cudaFree(ctx->d_tri_q); cudaFree(ctx->d_red_pe); cudaFree(ctx->d_red_pf);
    cudaFree(ctx->d_red_b); cudaFree(ctx->d_red_inv); cudaFree(ctx->d_red_del);
    cudaFree(ctx->d_spec); cudaFree(ctx->d_iface); cudaFree(ctx->d_chirp); cudaFree(ctx->d_bhat);
    for (int i = 0; i < ctx->nrad; ++i) { cudaFree(ctx->d_cs_cos[i]); cudaFree(ctx->d_cs_sin[i]); cudaFree(ctx->d_cs_frag[i]); }
}

// K2 + K3 + K4 on ctx->d_spec (already holding the row transforms of nbatch right-hand sides)
static int poisson_finish(hpb_ctx *ctx, hpb_slice sl, const int *c_lhs, int nbatch)
{
    const hpb_geom &g = ctx->g;
    const int nx = g.nx, ny = g.ny, N = ctx->fftN, L = ctx->th_L, C = ctx->th_C;
    const long plane = (long)nx * ny;
    const long ifn = 3L * C * nx;
    double *yf = ctx->d_iface, *ye = yf + ifn, *xl = ye + ifn, *xr = xl + ifn;
    dim3 g2((nx + 127) / 128, C, nbatch);
    hpb_launch(k_thomas_local, g2, 128, 0, ctx->stream, ctx->d_spec, ctx->d_tri_m, ctx->d_tri_c, yf, ye, nx, ny, L, C,
                                                ctx->th_last_base, 1.0 / (g.dy * g.dy));
    dim3 g3((nx + 63) / 64, nbatch);
    hpb_launch(k_thomas_reduced, g3, 64, 0, ctx->stream, yf, ye, xl, xr, ctx->d_red_pe, ctx->d_red_pf, ctx->d_red_b,
                                                  ctx->d_red_inv, ctx->d_red_del, nx, C);
    hpb_count_launch(ctx, 2);
    SliceView v = make_view(sl);
    OutPtrs o2;
    for (int b = 0; b < 4; ++b) o2.p[b] = v.comp(c_lhs[b < nbatch ? b : 0]) + v.idx(0, 0);
    SrcSpec ss{ctx->d_spec, nx, plane, xl, xr, ctx->d_tri_p, ctx->d_tri_q, L, C, nx, ctx->th_last_base};
    int rc = launch_rows(ctx, ss, o2, sl.jstride, nbatch, 1.0 / (2.0 * N));
    if (rc) return rc;
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_poisson_solve(hpb_ctx *ctx, const double *d_rhs, hpb_slice sl, const int *c_lhs,
                                 int nbatch)
{
    if (!ctx || !d_rhs || !c_lhs || nbatch < 1 || nbatch > 3) return HPB_ERR_ARG;
    const int nx = ctx->g.nx, ny = ctx->g.ny;
    const long plane = (long)nx * ny;
    OutPtrs o1;
    for (int b = 0; b < 4; ++b) o1.p[b] = ctx->d_spec + (b < nbatch ? b : 0) * plane;
    SrcStage st{d_rhs, nx, plane};
    int rc = launch_rows(ctx, st, o1, nx, nbatch, 1.0);
    if (rc) return rc;
    return poisson_finish(ctx, sl, c_lhs, nbatch);
}

int hpb_ref_arm_solve_psi_ez_bz(hpb_ctx *ctx, hpb_slice sl, const int *comps);

extern "C" int hpb_fields_solve_psi_ez_bz(hpb_ctx *ctx, hpb_slice sl, const int *comps)
{
    if (!ctx || !comps) return HPB_ERR_ARG;
    if (ctx->tune_poisson_impl == 1) return hpb_ref_arm_solve_psi_ez_bz(ctx, sl, comps);    // measurement arm
    const hpb_geom &g = ctx->g;
    const long plane = (long)g.nx * g.ny;
    OutPtrs o1;
    for (int b = 0; b < 4; ++b) o1.p[b] = ctx->d_spec + (b < 3 ? b : 0) * plane;
    SrcFields sf{make_view(sl), comps[HPB_C_RHOMJZ], comps[HPB_C_JX], comps[HPB_C_JY],
                 -1.0 / g.ep0, 1.0 / (g.ep0 * g.c), g.mu0, 0.5 * (1.0 / g.dx), 0.5 * (1.0 / g.dy)};
    int rc = launch_rows(ctx, sf, o1, g.nx, 3, 1.0);
    if (rc) return rc;
    const int lhs[3] = {comps[HPB_C_PSI], comps[HPB_C_EZ], comps[HPB_C_BZ]};
    rc = poisson_finish(ctx, sl, lhs, 3);
    if (rc) return rc;
    return hpb_launch_exmby_eypbx(ctx, sl, comps);
}
