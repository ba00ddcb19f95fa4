// Shared-memory complex FFT of arbitrary length (Stockham autosort, radix 4 / 2 / 3 / 5 and a generic
// odd-prime stage with tabulated cos / sin): the building block of the Dirichlet Poisson solver's row
// transforms (poisson.cu) and of the laser envelope's 2-D complex transform (fft2d.cu).  One CTA
// transforms one sequence held in shared memory; `root` is exp(-2 pi i t / N), t = 0 .. N-1.
#pragma once
#include "common.cuh"
#include <math.h>
#include <vector>

namespace {

constexpr int kFftThreads = 256;
constexpr int kMaxRad = 32;

struct FftPlan {
    int N; int nrad; int rad[kMaxRad];
    const double *cs_cos[kMaxRad];     // per prime stage: cos(2 pi t b / p), [b = 0..h][t = 1..h]
    const double *cs_sin[kMaxRad];
    // per prime stage: the same coefficients in the operand-fragment order of mma.m8n8k4.f64
    // (fft_prime_frag_table); nullptr = the scalar inner product
    const double *cs_frag[kMaxRad];
    // Bluestein (chirp-z) for lengths with a large prime factor (257, 2049 = 3 * 683, ...): M = the
    // power of two >= 2N - 2 (enough because the chirp is even), rad[] then factor M, `root` passed to
    // fft_smem is exp(-2 pi i t / M); chirp[n] = exp(-i pi n^2 / N), bhat = FFT_M(conj chirp, wrapped) / M
    int M;
    const double2 *chirp, *bhat;
    __host__ __device__ int buf_len() const { return M > 0 ? M : N; }     // complex elements per buffer
    __host__ __device__ int max_prime() const                             // largest generic (> 5) radix, or 0
    {
        int q = 0;
        for (int i = 0; i < nrad; ++i) if (rad[i] > 5 && rad[i] > q) q = rad[i];
        return q;
    }
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// j / d and j % d for 0 <= j < 2^21 and a run-time divisor: a float reciprocal instead of the
// ~20-instruction integer division sequence (the stage loops are instruction-issue bound)
struct FastDiv {
    int d; float inv;
    __device__ __forceinline__ explicit FastDiv(int d_) : d(d_), inv(1.0f / (float)d_) {}
    __device__ __forceinline__ int div(int j) const
    {
        int q = (int)(((float)j + 0.5f) * inv);
        // the float estimate is off by at most one
        const int r = j - q * d;
        q += (r >= d) - (r < 0);
        return q;
    }
    __device__ __forceinline__ int mod(int j) const { return j - div(j) * d; }
};

// Stockham stage of radix r with Ns = product of the previous radices:
//   out[(j-k) r + k + b Ns] = sum_t ( in[j + t N/r] w_N^{t k N/(Ns r)} ) w_r^{t b},  k = j mod Ns

__device__ __forceinline__ void fft_stage_r2(const double2 *__restrict__ in, double2 *__restrict__ out,
                                             const double2 *__restrict__ root, int N, int Ns)
{
    const int Nr = N >> 1, tw = N / (Ns * 2);
    const FastDiv fd(Ns);
    for (int j = threadIdx.x; j < Nr; j += blockDim.x) {
        const int k = fd.mod(j);
        double2 v0 = in[j], v1 = in[j + Nr];
        if (k) v1 = cmul(v1, __ldg(&root[k * tw]));
        const int o = (j - k) * 2 + k;
        out[o] = cadd(v0, v1);
        out[o + Ns] = csub(v0, v1);
    }
}

__device__ __forceinline__ void fft_stage_r3(const double2 *__restrict__ in, double2 *__restrict__ out,
                                             const double2 *__restrict__ root, int N, int Ns)
{
    const int Nr = N / 3, tw = N / (Ns * 3);
    const double c1 = -0.5, s1 = 0.86602540378443864676;
    const FastDiv fd(Ns);
    for (int j = threadIdx.x; j < Nr; j += blockDim.x) {
        const int k = fd.mod(j);
        double2 v0 = in[j], v1 = in[j + Nr], v2 = in[j + 2 * Nr];
        if (k) { v1 = cmul(v1, __ldg(&root[k * tw])); v2 = cmul(v2, __ldg(&root[2 * k * tw])); }
        const double2 a = cadd(v1, v2), d = csub(v1, v2);
        const double2 P = make_double2(v0.x + c1 * a.x, v0.y + c1 * a.y);
        const double2 Q = make_double2(s1 * d.x, s1 * d.y);
        const int o = (j - k) * 3 + k;
        out[o] = cadd(v0, a);
        out[o + Ns] = make_double2(P.x + Q.y, P.y - Q.x);          // P - iQ
        out[o + 2 * Ns] = make_double2(P.x - Q.y, P.y + Q.x);      // P + iQ
    }
}

// radix-4 stage, one butterfly per thread iteration (4 inputs read once)
__device__ __forceinline__ void fft_stage_r4(const double2 *__restrict__ in, double2 *__restrict__ out,
                                             const double2 *__restrict__ root, int N, int Ns)
{
    const int Nr = N >> 2;
    const int tw = N / (Ns * 4);
    const FastDiv fd(Ns);
    for (int j = threadIdx.x; j < Nr; j += blockDim.x) {
        const int k = fd.mod(j);
        double2 v0 = in[j], v1 = in[j + Nr], v2 = in[j + 2 * Nr], v3 = in[j + 3 * Nr];
        if (k) {
            v1 = cmul(v1, __ldg(&root[k * tw]));
            v2 = cmul(v2, __ldg(&root[2 * k * tw]));
            v3 = cmul(v3, __ldg(&root[3 * k * tw]));
        }
        const double2 s02 = cadd(v0, v2), d02 = csub(v0, v2), s13 = cadd(v1, v3), d13 = csub(v1, v3);
        const int o = (j - k) * 4 + k;
        out[o] = cadd(s02, s13);
        out[o + Ns] = make_double2(d02.x + d13.y, d02.y - d13.x);       // d02 - i d13
        out[o + 2 * Ns] = csub(s02, s13);
        out[o + 3 * Ns] = make_double2(d02.x - d13.y, d02.y + d13.x);   // d02 + i d13
    }
}

__device__ __forceinline__ void fft_stage_r5(const double2 *__restrict__ in, double2 *__restrict__ out,
                                             const double2 *__restrict__ root, int N, int Ns)
{
    const int Nr = N / 5, tw = N / (Ns * 5);
    const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
    const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
    const FastDiv fd(Ns);
    for (int j = threadIdx.x; j < Nr; j += blockDim.x) {
        const int k = fd.mod(j);
        double2 v0 = in[j], v1 = in[j + Nr], v2 = in[j + 2 * Nr], v3 = in[j + 3 * Nr], v4 = in[j + 4 * Nr];
        if (k) {
            v1 = cmul(v1, __ldg(&root[k * tw]));
            v2 = cmul(v2, __ldg(&root[2 * k * tw]));
            v3 = cmul(v3, __ldg(&root[3 * k * tw]));
            v4 = cmul(v4, __ldg(&root[4 * k * tw]));
        }
        const double2 a1 = cadd(v1, v4), a2 = cadd(v2, v3), d1 = csub(v1, v4), d2 = csub(v2, v3);
        const double2 P1 = make_double2(v0.x + c1 * a1.x + c2 * a2.x, v0.y + c1 * a1.y + c2 * a2.y);
        const double2 P2 = make_double2(v0.x + c2 * a1.x + c1 * a2.x, v0.y + c2 * a1.y + c1 * a2.y);
        const double2 Q1 = make_double2(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y);
        const double2 Q2 = make_double2(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y);
        const int o = (j - k) * 5 + k;
        out[o] = make_double2(v0.x + a1.x + a2.x, v0.y + a1.y + a2.y);
        out[o + Ns] = make_double2(P1.x + Q1.y, P1.y - Q1.x);
        out[o + 4 * Ns] = make_double2(P1.x - Q1.y, P1.y + Q1.x);
        out[o + 2 * Ns] = make_double2(P2.x + Q2.y, P2.y - Q2.x);
        out[o + 3 * Ns] = make_double2(P2.x - Q2.y, P2.y + Q2.x);
    }
}

// any odd prime radix p = 2h+1.  Two sub-steps through `tmp` (N entries):
//   A: x_t = in[j + t Nr] w^{t k tw};  U[t][j] = x_t + x_{p-t},  V[t][j] = x_t - x_{p-t},  X0[j]
//   B: X_b = x0 + sum_t U_t cos(2 pi t b/p) - i sum_t V_t sin(2 pi t b/p),  X_{p-b} = conj-sign
// Result is written back into `in` (no buffer swap for this stage).
// s_tab: 2 p doubles of shared memory.  The p-th roots of unity cos / sin(2 pi m / p), m = 0 .. p-1, are
// staged there once per stage and indexed with m = (t b) mod p kept incrementally: the inner loop's two
// table loads per step used to be global (L1) loads -- 40 % of the row kernel's stall samples were
// long-scoreboard waits on them (profiles: r01e source page) -- and are shared-memory broadcasts now.
// D (8 x 8) += A (8 x 4, row) B (4 x 8, col) on the fp64 tensor-core path: lane l holds A[l / 4][l % 4],
// B[l % 4][l / 4] and D[l / 4][2 (l % 4) + {0, 1}]
__device__ __forceinline__ void dmma_884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void fft_stage_prime(double2 *__restrict__ in, double2 *__restrict__ tmp,
                                                const double2 *__restrict__ root, int N, int Ns, int p,
                                                const double *__restrict__ tcos,
                                                const double *__restrict__ tsin, double *__restrict__ s_tab,
                                                const double *__restrict__ frag)
{
    const int Nr = N / p, h = (p - 1) / 2, tw = N / (Ns * p);
    const FastDiv fNr(Nr), fNs(Ns);
    double *s_c = s_tab, *s_s = s_tab + p;
    for (int m = threadIdx.x; m < (frag ? 0 : p); m += blockDim.x) {
        // row b = 1 of the (b, t) tables holds cos / sin(2 pi t / p), t = 1 .. h
        const int mm = m <= h ? m : p - m;
        s_c[m] = m == 0 ? 1.0 : __ldg(&tcos[h + mm - 1]);
        s_s[m] = m == 0 ? 0.0 : (m <= h ? __ldg(&tsin[h + mm - 1]) : -__ldg(&tsin[h + mm - 1]));
    }
    for (int e = threadIdx.x; e < Nr * h; e += blockDim.x) {
        const int tq = fNr.div(e);
        const int t = tq + 1, j = e - tq * Nr;
        const int k = fNs.mod(j);
        double2 a = in[j + t * Nr], b = in[j + (p - t) * Nr];
        if (k) {
            // t k tw < p Ns N / (Ns p) = N: the twiddle index needs no reduction
            a = cmul(a, __ldg(&root[t * k * tw]));
            b = cmul(b, __ldg(&root[(p - t) * k * tw]));
        }
        tmp[(t - 1) * Nr + j] = cadd(a, b);
        tmp[(h + t - 1) * Nr + j] = csub(a, b);
    }
    for (int j = threadIdx.x; j < Nr; j += blockDim.x) tmp[2 * h * Nr + j] = in[j];
    __syncthreads();
    if (frag) {
        // Sub-step B as two small dense products on the tensor cores: P = C U and Q = S V with
        // C / S[b][t] = cos / sin(2 pi t b / p) ((h+1) x h, padded to 8 x 4 tiles with zeros) and U, V the
        // h x 2 Nr real matrices tmp already holds (row t, columns = Re / Im of the Nr sequences: the double2
        // layout IS the row-major real matrix with leading dimension 2 Nr).  The scalar form of this step
        // issued 2 LDS.128 + 2 conflicting LDS.64 per 4 DFMA and made the row kernels shared-memory-pipe
        // bound (l1tex 83 % of peak, profiles/r01e); a warp-level mma reads each operand element once per
        // 8 x 8 tile.  One output tile (8 values of b x 4 sequences) per warp and iteration.
        const int KT = (h + 3) >> 2, MT = (h + 8) >> 3, NTl = (2 * Nr + 7) >> 3;
        const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
        const int ld = 2 * Nr;
        const double *T = reinterpret_cast<const double *>(tmp);
        for (int tile = threadIdx.x >> 5; tile < MT * NTl; tile += (int)(blockDim.x >> 5)) {
            const int nt = tile / MT, mt = tile - nt * MT;
            const int n = nt * 8 + gid;                          // column of the B operand held by this lane
            double p0 = 0., p1 = 0., q0 = 0., q1 = 0.;
            const double *fr = frag + (size_t)mt * KT * 64 + lane;
            for (int kt = 0; kt < KT; ++kt) {
                const int t = kt * 4 + tig;
                const bool okb = t < h && n < ld;                // (padding must be a true 0, not 0 x garbage)
                const double bu = okb ? T[t * ld + n] : 0.;
                const double bv = okb ? T[(h + t) * ld + n] : 0.;
                dmma_884(p0, p1, __ldg(fr + kt * 64), bu);
                dmma_884(q0, q1, __ldg(fr + kt * 64 + 32), bv);
            }
            const int b = mt * 8 + gid, j = nt * 4 + tig;        // D fragment: (Re, Im) of sequence j, output b
            if (b <= h && j < Nr) {
                const double2 x0 = tmp[2 * h * Nr + j];
                const double Px = x0.x + p0, Py = x0.y + p1;
                const int k = fNs.mod(j);
                const int o = (j - k) * p + k;
                in[o + b * Ns] = make_double2(Px + q1, Py - q0);                    // P - iQ
                if (b) in[o + (p - b) * Ns] = make_double2(Px - q1, Py + q0);       // P + iQ
            }
        }
        return;
    }
    // BB outputs b per work item share the shared-memory reads of U_t, V_t (BB > 1 measured slower:
    // the extra accumulators spill at 6 CTAs / SM)
    constexpr int BB = 1;
    const int nbg = (h + BB) / BB;              // ceil((h + 1) / BB)
    for (int e = threadIdx.x; e < Nr * nbg; e += blockDim.x) {
        const int bg = fNr.div(e), j = e - bg * Nr;
        const int b0 = bg * BB;
        const int k = fNs.mod(j);
        const double2 x0 = tmp[2 * h * Nr + j];
        double2 P[BB], Q[BB];
#pragma unroll
        for (int bb = 0; bb < BB; ++bb) { P[bb] = x0; Q[bb] = make_double2(0., 0.); }
        int m[BB];
#pragma unroll
        for (int bb = 0; bb < BB; ++bb) m[bb] = 0;
#pragma unroll 4
        for (int t = 0; t < h; ++t) {
            const double2 u = tmp[t * Nr + j], v = tmp[(h + t) * Nr + j];
#pragma unroll
            for (int bb = 0; bb < BB; ++bb) {
                const int b = min(b0 + bb, h);          // (a clamped duplicate is simply not stored)
                m[bb] += b;                              // ((t + 1) b) mod p
                if (m[bb] >= p) m[bb] -= p;
                const double c = s_c[m[bb]], sn = s_s[m[bb]];
                P[bb].x += c * u.x; P[bb].y += c * u.y;
                Q[bb].x += sn * v.x; Q[bb].y += sn * v.y;
            }
        }
        const int o = (j - k) * p + k;
#pragma unroll
        for (int bb = 0; bb < BB; ++bb) {
            const int b = b0 + bb;
            if (b > h) break;
            in[o + b * Ns] = make_double2(P[bb].x + Q[bb].y, P[bb].y - Q[bb].x);                    // P - iQ
            if (b) in[o + (p - b) * Ns] = make_double2(P[bb].x - Q[bb].y, P[bb].y + Q[bb].x);       // P + iQ
        }
    }
}

// the Stockham stages of a plan on a sequence of length L held in src (scratch dst); returns the
// buffer holding the result
__device__ __forceinline__ double2 *fft_stages(double2 *src, double2 *dst, const FftPlan &plan, int L,
                                               const double2 *__restrict__ root, double *s_tab)
{
    int Ns = 1;
    for (int s = 0; s < plan.nrad; ++s) {
        const int r = plan.rad[s];
        bool swap = true;
        if (r == 4) fft_stage_r4(src, dst, root, L, Ns);
        else if (r == 2) fft_stage_r2(src, dst, root, L, Ns);
        else if (r == 5) fft_stage_r5(src, dst, root, L, Ns);
        else if (r == 3) fft_stage_r3(src, dst, root, L, Ns);
        else { fft_stage_prime(src, dst, root, L, Ns, r, plan.cs_cos[s], plan.cs_sin[s], s_tab, plan.cs_frag[s]); swap = false; }
        __syncthreads();
        if (swap) { double2 *t = src; src = dst; dst = t; }
        Ns *= r;
    }
    return src;
}

// forward complex FFT of length N on shared memory (both buffers hold plan.buf_len() elements, the
// input is src[0 .. N-1]); returns the buffer holding the result
// BLUE is a compile-time switch: the radix-only instantiation keeps the register budget it was tuned
// with (the chirp-z branch costs k_dst_rows 170 bytes of spills at 40 registers when it is merely present)
template <bool BLUE>
__device__ __forceinline__ double2 *fft_smem(double2 *src, double2 *dst, const FftPlan &plan,
                                             const double2 *__restrict__ root, double *s_tab)
{
    if (!BLUE) return fft_stages(src, dst, plan, plan.N, root, s_tab);
    // Bluestein: X[k] = c[k] sum_n (x[n] c[n]) conj(c)[k - n],  c[n] = exp(-i pi n^2 / N): a circular
    // convolution of length M through two power-of-two FFTs
    const int N = plan.N, M = plan.M;
    for (int n = threadIdx.x; n < M; n += blockDim.x)
        dst[n] = n < N ? cmul(src[n], __ldg(&plan.chirp[n])) : make_double2(0., 0.);
    __syncthreads();
    double2 *A = fft_stages(dst, src, plan, M, root, s_tab);
    double2 *B = (A == dst) ? src : dst;
    // times bhat; conjugated so that the second forward FFT is the inverse transform
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const double2 v = cmul(A[m], __ldg(&plan.bhat[m]));
        A[m] = make_double2(v.x, -v.y);
    }
    __syncthreads();
    double2 *Cv = fft_stages(A, B, plan, M, root, s_tab);
    double2 *out = (Cv == A) ? B : A;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const double2 v = make_double2(Cv[k].x, -Cv[k].y);
        out[k] = cmul(v, __ldg(&plan.chirp[k]));
    }
    __syncthreads();
    return out;
}

// ---- host: plan construction shared by poisson.cu and fft2d.cu ----------------------------------------
// Factorisation of N; lengths whose largest prime factor exceeds hpb_bluestein_min_prime() (default 64,
// process-wide option "bluestein_min_prime") switch to Bluestein (returns M > 0 and factors M instead).
inline int fft_factorize(int N, int *rad, int &nrad)
{
    auto factor = [&](int n) {
        nrad = 0;
        const int pref[] = {4, 2, 3, 5};
        for (int r : pref) while (n % r == 0) { rad[nrad++] = r; n /= r; }
        int largest = nrad ? 5 : 1;
        for (int q = 7; n > 1; q += 2) while (n % q == 0) { rad[nrad++] = q; n /= q; largest = q; }
        return largest;
    };
    if (factor(N) <= hpb_bluestein_min_prime()) return 0;
    int M = 1;
    while (M < 2 * N - 2) M <<= 1;
    factor(M);
    return M;
}

// Coefficients of the prime stage p = 2h + 1 in mma.m8n8k4 A-fragment order:
//   tab[((mt KT + kt) 2 + cs) 32 + lane] = cos (cs = 0) / sin (cs = 1) of 2 pi t b / p,  b = 8 mt + lane / 4,
//   t = 4 kt + lane % 4 + 1;  0 where b > h or t > h
inline std::vector<double> fft_prime_frag_table(int p)
{
    const int h = (p - 1) / 2, KT = (h + 3) / 4, MT = (h + 8) / 8;
    const long double pi = 3.14159265358979323846264338327950288L;
    std::vector<double> tab((size_t)MT * KT * 64, 0.0);
    for (int mt = 0; mt < MT; ++mt)
        for (int kt = 0; kt < KT; ++kt)
            for (int lane = 0; lane < 32; ++lane) {
                const int b = 8 * mt + lane / 4, t = 4 * kt + lane % 4 + 1;
                if (b > h || t > h) continue;
                const long tb = ((long)t * b) % p;
                tab[((size_t)(mt * KT + kt) * 2 + 0) * 32 + lane] = (double)cosl(2.0L * pi * tb / p);
                tab[((size_t)(mt * KT + kt) * 2 + 1) * 32 + lane] = (double)sinl(2.0L * pi * tb / p);
            }
    return tab;
}

// chirp[n] = exp(-i pi n^2 / N) (n < N) and bhat = FFT_M(b) / M with b[m] = exp(+i pi m^2 / N) wrapped
// (b[M - m] = b[m]); long-double arithmetic, O(M^2) on the host once per plan
inline void fft_bluestein_tables(int N, int M, double2 *chirp, double2 *bhat)
{
    const long double pi = 3.14159265358979323846264338327950288L;
    std::vector<long double> br(M, 0.0L), bi(M, 0.0L);
    for (int n = 0; n < N; ++n) {
        const long long q = ((long long)n * n) % (2LL * N);          // n^2 mod 2N keeps the argument small
        const long double a = pi * (long double)q / (long double)N;
        chirp[n].x = (double)cosl(a); chirp[n].y = (double)-sinl(a);
        br[n] = cosl(a); bi[n] = sinl(a);
        if (n > 0) { br[M - n] = cosl(a); bi[M - n] = sinl(a); }
    }
    std::vector<long double> wr(M), wi(M);
    for (int t = 0; t < M; ++t) { wr[t] = cosl(-2.0L * pi * t / M); wi[t] = sinl(-2.0L * pi * t / M); }
    for (int k = 0; k < M; ++k) {
        long double sr = 0.0L, si = 0.0L;
        for (int m = 0; m < M; ++m) {
            if (br[m] == 0.0L && bi[m] == 0.0L) continue;
            const int t = (int)(((long long)k * m) % M);
            sr += br[m] * wr[t] - bi[m] * wi[t];
            si += br[m] * wi[t] + bi[m] * wr[t];
        }
        bhat[k].x = (double)(sr / M); bhat[k].y = (double)(si / M);
    }
}

}  // namespace
