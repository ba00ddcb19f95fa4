// 2-D complex-to-complex FFT of the laser grid (the transform MultiLaser::AdvanceSliceFFT needs,
// src/laser/MultiLaser.cpp:610-800; the reference calls AnyFFT / cufftExecZ2Z, WrapCuFFT.cpp:126-137).
// Hand-written for sm_100a on the shared-memory Stockham FFT of fft_smem.cuh: one CTA per row, then
// one CTA per column (the whole sequence lives in shared memory, every global element is read and
// written once per pass; a 512 x 512 complex grid is 4 MB and stays in the L2 between the passes).
// Same conventions as cuFFT: unnormalised, forward = exp(-i ...), inverse = exp(+i ...); in place or
// out of place.
#include "fft_smem.cuh"
#include <math.h>
#include <vector>

namespace {

struct Plan1D {
    FftPlan plan;
    double2 *d_root = nullptr, *d_chirp = nullptr, *d_bhat = nullptr;
    double *d_cos[kMaxRad] = {}, *d_sin[kMaxRad] = {}, *d_frag[kMaxRad] = {};
};

template <class T>
bool to_device(T **dptr, const std::vector<T> &h)
{
    return cudaMalloc(dptr, sizeof(T) * h.size()) == cudaSuccess
           && cudaMemcpy(*dptr, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice) == cudaSuccess;
}

bool build_plan(int N, Plan1D &p)
{
    p.plan.N = N;
    p.plan.M = fft_factorize(N, p.plan.rad, p.plan.nrad);
    p.plan.chirp = p.plan.bhat = nullptr;
    const int L = p.plan.M > 0 ? p.plan.M : N;
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < p.plan.nrad; ++i) {
        p.plan.cs_cos[i] = p.plan.cs_sin[i] = p.plan.cs_frag[i] = nullptr;
        const int q = p.plan.rad[i];
        if (q <= 5) continue;
        const int h = (q - 1) / 2;
        std::vector<double> tc((size_t)(h + 1) * h), ts((size_t)(h + 1) * h);
        for (int b = 0; b <= h; ++b)
            for (int t = 1; t <= h; ++t) {
                const long tb = ((long)t * b) % q;
                tc[(size_t)b * h + t - 1] = (double)cosl(2.0L * pi * tb / q);
                ts[(size_t)b * h + t - 1] = (double)sinl(2.0L * pi * tb / q);
            }
        if (!to_device(&p.d_cos[i], tc) || !to_device(&p.d_sin[i], ts)
            || !to_device(&p.d_frag[i], fft_prime_frag_table(q))) return false;
        p.plan.cs_cos[i] = p.d_cos[i]; p.plan.cs_sin[i] = p.d_sin[i]; p.plan.cs_frag[i] = p.d_frag[i];
    }
    std::vector<double2> root(L);
    for (int t = 0; t < L; ++t) {
        root[t].x = (double)cosl(-2.0L * pi * t / L);
        root[t].y = (double)sinl(-2.0L * pi * t / L);
    }
    if (p.plan.M > 0) {
        std::vector<double2> chirp(N), bhat(p.plan.M);
        fft_bluestein_tables(N, p.plan.M, chirp.data(), bhat.data());
        if (!to_device(&p.d_chirp, chirp) || !to_device(&p.d_bhat, bhat)) return false;
        p.plan.chirp = p.d_chirp; p.plan.bhat = p.d_bhat;
    }
    return to_device(&p.d_root, root);
}
void free_plan(Plan1D &p)
{
    cudaFree(p.d_root); cudaFree(p.d_chirp); cudaFree(p.d_bhat);
    for (int i = 0; i < kMaxRad; ++i) { cudaFree(p.d_cos[i]); cudaFree(p.d_sin[i]); cudaFree(p.d_frag[i]); }
}

// One sequence of N complex numbers per CTA: element e of sequence s is at in[s * seq_stride + e *
// elem_stride] (rows: seq_stride = nx, elem_stride = 1; columns: seq_stride = 1, elem_stride = nx).
// INV: conj in, forward transform, conj out = the inverse (unnormalised) transform.
template <bool INV, bool BLUE>
__global__ void __launch_bounds__(kFftThreads)
k_fft_seq(const double2 *__restrict__ in, double2 *__restrict__ out, long seq_stride, long elem_stride,
          FftPlan plan, const double2 *__restrict__ root)
{
    hpb_pdl_prologue();
    extern __shared__ double2 fsm[];
    const int N = plan.N;
    double2 *b0 = fsm, *b1 = fsm + plan.buf_len();
    const long base = (long)blockIdx.x * seq_stride;
    for (int e = threadIdx.x; e < N; e += blockDim.x) {
        double2 v = in[base + (long)e * elem_stride];
        if (INV) v.y = -v.y;
        b0[e] = v;
    }
    __syncthreads();
    const double2 *F = fft_smem<BLUE>(b0, b1, plan, root, reinterpret_cast<double *>(fsm + 2 * plan.buf_len()));
    for (int e = threadIdx.x; e < N; e += blockDim.x) {
        double2 v = F[e];
        if (INV) v.y = -v.y;
        out[base + (long)e * elem_stride] = v;
    }
}

}  // namespace

struct hpb_fft2d {
    int nx = 0, ny = 0;
    Plan1D px, py;
};

int hpb_fft2d_create(hpb_fft2d **out, int nx, int ny)
{
    if (!out || nx < 2 || ny < 2) return HPB_ERR_ARG;
    hpb_fft2d *f = new hpb_fft2d();
    f->nx = nx; f->ny = ny;
    if (!build_plan(nx, f->px) || !build_plan(ny, f->py)
        || 2 * sizeof(double2) * (size_t)std::max(f->px.plan.buf_len(), f->py.plan.buf_len()) > 220 * 1024) {
        hpb_set_error("fft2d: %d x %d not supported by the shared-memory FFT", nx, ny);
        free_plan(f->px); free_plan(f->py);
        delete f;
        return HPB_ERR_UNSUPPORTED;
    }
    cudaFuncSetAttribute(k_fft_seq<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_fft_seq<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_fft_seq<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_fft_seq<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    *out = f;
    return HPB_OK;
}

void hpb_fft2d_destroy(hpb_fft2d *f)
{
    if (!f) return;
    free_plan(f->px); free_plan(f->py);
    delete f;
}

// data[j * nx + i]; dir < 0 forward, dir > 0 inverse; in == out allowed
int hpb_fft2d_exec(hpb_fft2d *f, hpb_ctx *ctx, const double2 *in, double2 *out, int dir)
{
    if (!f || !ctx || !in || !out) return HPB_ERR_ARG;
    const int nx = f->nx, ny = f->ny;
    const size_t smx = 2 * sizeof(double2) * (size_t)f->px.plan.buf_len() + 2 * sizeof(double) * (size_t)f->px.plan.max_prime();
    const size_t smy = 2 * sizeof(double2) * (size_t)f->py.plan.buf_len() + 2 * sizeof(double) * (size_t)f->py.plan.max_prime();
    auto pass = [&](const double2 *src, double2 *dst, unsigned nseq, size_t smem, long seq_stride, long elem_stride,
                    const Plan1D &p) {
        const bool blue = p.plan.M > 0;
        if (dir < 0) {
            if (blue) hpb_launch(k_fft_seq<false, true>, nseq, kFftThreads, smem, ctx->stream, src, dst, seq_stride, elem_stride, p.plan, (const double2 *)p.d_root);
            else hpb_launch(k_fft_seq<false, false>, nseq, kFftThreads, smem, ctx->stream, src, dst, seq_stride, elem_stride, p.plan, (const double2 *)p.d_root);
        } else {
            if (blue) hpb_launch(k_fft_seq<true, true>, nseq, kFftThreads, smem, ctx->stream, src, dst, seq_stride, elem_stride, p.plan, (const double2 *)p.d_root);
            else hpb_launch(k_fft_seq<true, false>, nseq, kFftThreads, smem, ctx->stream, src, dst, seq_stride, elem_stride, p.plan, (const double2 *)p.d_root);
        }
    };
    pass(in, out, (unsigned)ny, smx, (long)nx, 1L, f->px);                       // rows
    pass((const double2 *)out, out, (unsigned)nx, smy, 1L, (long)nx, f->py);     // columns
    hpb_count_launch(ctx, 2);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}
