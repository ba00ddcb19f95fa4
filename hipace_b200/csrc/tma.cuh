// Tensor Memory Accelerator helpers: mbarrier + cp.async.bulk(.tensor) wrappers (inline PTX for
// sm_100a; SASS: UTMALDG / UBLKCP / SYNCS) and the host-side tensor-map encoder.
//
// Used to stage field patches of the slice array (a 3-D tensor x, y, component of fp64) into shared
// memory with ONE instruction per field instead of one 8-byte LDGSTS per lane and cell, and to
// stream contiguous rows (Poisson solver) with bulk copies.  cuTensorMapEncodeTiled is resolved at
// run time through cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hpb200.h"

// Tensor map over a slice array for (box_w x box_h x 1)-boxes of one component.  Returns false
// if the array cannot be described to the TMA unit (base not 16-byte aligned, a stride that is not
// a multiple of 16 bytes -- e.g. an odd row length in doubles -- or no driver entry point): the
// callers then use their cp.async (LDGSTS) staging instead.
bool hpb_encode_slice_tmap(const hpb_slice &sl, int box_w, int box_h, CUtensorMap *out);

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t hpb_smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void hpb_mbar_init(uint64_t *bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hpb_smem_u32(bar)), "r"(arrivals) : "memory");
    // make the initialised barrier visible to the async proxy (the TMA unit completes on it)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one arrival that also announces `bytes` of asynchronous copies that will complete on the barrier
__device__ __forceinline__ void hpb_mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(hpb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool hpb_mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(hpb_smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
// wait for the phase with the given parity.  try_wait suspends the thread in hardware until the
// phase completes or a time limit passes; a copy that never completes (a bad tensor map) traps after
// about two seconds instead of hanging the device.
__device__ __forceinline__ void hpb_mbar_wait(uint64_t *bar, uint32_t parity)
{
    if (hpb_mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!hpb_mbar_try_wait(bar, parity))
        if (clock64() - t0 > 4000000000ll) __trap();
}
// generic-proxy accesses to shared memory (the threads' loads of the previous tile) are ordered
// before the async-proxy writes of the next TMA copy into the same buffer
__device__ __forceinline__ void hpb_fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 3-D tiled load: box at element coordinates (x, y, z) -> dense shared-memory tile; completes
// `box bytes` on the barrier.  Out-of-bounds elements are filled with zeros.
__device__ __forceinline__ void hpb_tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar,
                                                int x, int y, int z)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(hpb_smem_u32(smem_dst)), "l"((uint64_t)map), "r"(hpb_smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}
__device__ __forceinline__ void hpb_tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
// 1-D bulk copy global -> shared (bytes: multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void hpb_bulk_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                                 uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(hpb_smem_u32(smem_dst)), "l"((uint64_t)gmem_src), "r"(bytes), "r"(hpb_smem_u32(bar))
        : "memory");
}
#endif
