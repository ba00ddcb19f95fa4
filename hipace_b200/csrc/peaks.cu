// Measured fp64 FMA peak of the device this library runs on: the second roof (beside the HBM copy
// bandwidth of MEASURED_PEAKS.json) that the particle kernels are reported against -- the push and
// the explicit deposition execute ~700 fp64 instructions per particle and are closer to this roof
// than to the memory one.  Measurement infrastructure: called by
// bench.py, never by the slice loop.
#include "common.cuh"

namespace {

// 8 independent DFMA chains per thread, 4096 dependent steps each: 16 warps per SM sub-partition keep
// the fp64 pipe saturated; the result is stored so that nothing is optimised away
__global__ void __launch_bounds__(256)
k_dfma_peak(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3., x4 = x0 + 4., x5 = x0 + 5., x6 = x0 + 6.,
           x7 = x0 + 7.;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(long)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace

// Best of `reps` timed launches (CUDA events on a private stream, after one warm-up launch).
// tflops: 2 flops per DFMA.  Returns HPB_OK or HPB_ERR_CUDA.
extern "C" int hpb_measure_fp64_peak(int device, int reps, double *tflops)
{
    if (!tflops || reps < 1) return HPB_ERR_ARG;
    HPB_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    HPB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2048;
    double *d_out = nullptr;
    HPB_CUDA_CHECK(cudaMalloc(&d_out, sizeof(double) * blocks * threads));
    cudaStream_t st;
    HPB_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    HPB_CUDA_CHECK(cudaEventCreate(&e0));
    HPB_CUDA_CHECK(cudaEventCreate(&e1));
    double best_ms = 1e30;
    for (int r = 0; r <= reps; ++r) {
        HPB_CUDA_CHECK(cudaEventRecord(e0, st));
        k_dfma_peak<<<blocks, threads, 0, st>>>(d_out, iters, 0.999999, 1e-9);
        HPB_CUDA_CHECK(cudaEventRecord(e1, st));
        HPB_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        HPB_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best_ms) best_ms = ms;
    }
    const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;      // 8 chains x 8 unrolled = 64 DFMA / iteration
    *tflops = flops / (best_ms * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st); cudaFree(d_out);
    return HPB_OK;
}
