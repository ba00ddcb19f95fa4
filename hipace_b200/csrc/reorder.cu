// Plasma particle reordering: a counting sort of the particle SoA by transverse cell, the role of
//   PlasmaParticleContainer::ReorderParticles   src/particles/plasma/PlasmaParticleContainer.cpp:196-208
//   (amrex::ParticleContainer::SortParticlesForDeposition with <plasma>.reorder_idx_type)
// called from the slice loop every <plasma>.reorder_period slices (src/Hipace.cpp:595).
//
// Three passes over the particles, none of them through the host:
//   k_reorder_key      cell key of every particle + its rank inside the cell (one integer atomic
//                      per particle on the L2-resident histogram)
//   k_scan_*           exclusive prefix sum of the histogram (block scan, scan of the block sums,
//                      add back) -- hand-written, 1 M bins take three short launches
//   k_reorder_scatter  every SoA stream is read once (coalesced) and written once to its sorted slot
// Invalid particles go to a last bin, so the valid ones end up contiguous and cell-ordered (x fastest).
// The rank inside a cell is the arrival order of the atomics: the particle ORDER inside a cell is
// not reproducible run to run (neither is AMReX's), the particle SET and every per-particle value are
// bit-exact, and the slice loop's sums do not depend on the order beyond fp64 round-off.
#include "common.cuh"

namespace {

constexpr int kT = 256;
constexpr int kScanBlock = 1024;        // elements per scan block (256 threads x 4)

struct ReorderGeom {
    double x_lo, y_lo, dx_inv, dy_inv, shift_x, shift_y;      // shift: 0 (cell) or 0.5 (node), reorder_idx_type
    int nbx, nby;
};

__global__ void __launch_bounds__(kT)
k_reorder_key(const double *__restrict__ x, const double *__restrict__ y, const uint64_t *__restrict__ idcpu,
              long np, ReorderGeom g, unsigned *__restrict__ key, unsigned *__restrict__ rank,
              unsigned *__restrict__ hist)
{
    hpb_pdl_prologue();
    const long ip = (long)blockIdx.x * kT + threadIdx.x;
    if (ip >= np) return;
    unsigned k = (unsigned)g.nbx * g.nby;                     // the bin of the invalid particles
    if (hpb_is_valid(idcpu[ip])) {
        int i = (int)floor((x[ip] - g.x_lo) * g.dx_inv + g.shift_x);
        int j = (int)floor((y[ip] - g.y_lo) * g.dy_inv + g.shift_y);
        i = i < 0 ? 0 : (i >= g.nbx ? g.nbx - 1 : i);
        j = j < 0 ? 0 : (j >= g.nby ? g.nby - 1 : j);
        k = (unsigned)j * g.nbx + i;
    }
    key[ip] = k;
    rank[ip] = atomicAdd(&hist[k], 1u);
}

// exclusive scan of one block of kScanBlock elements (in place), block total -> sums[block]
__global__ void __launch_bounds__(kT)
k_scan_blocks(unsigned *__restrict__ a, long n, unsigned *__restrict__ sums)
{
    hpb_pdl_prologue();
    __shared__ unsigned wsum[kT / 32];
    const long base = (long)blockIdx.x * kScanBlock + threadIdx.x * 4;
    unsigned v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = base + k < n ? a[base + k] : 0u;
    const unsigned t = v[0] + v[1] + v[2] + v[3];
    unsigned inc = t;                                         // inclusive scan of the thread totals
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned s = lane < kT / 32 ? wsum[lane] : 0u;
#pragma unroll
        for (int o = 1; o < kT / 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += up;
        }
        if (lane < kT / 32) wsum[lane] = s;                   // inclusive over the warps
    }
    __syncthreads();
    unsigned run = inc - t + (w > 0 ? wsum[w - 1] : 0u);      // exclusive prefix of this thread
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < n) a[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == kT - 1) sums[blockIdx.x] = wsum[kT / 32 - 1];
}

// exclusive scan of the block sums by ONE block (any count: chunks of kT with a running carry)
__global__ void __launch_bounds__(kT)
k_scan_sums(unsigned *__restrict__ sums, int nblocks)
{
    hpb_pdl_prologue();
    __shared__ unsigned wsum[kT / 32];
    __shared__ unsigned carry_s;
    if (threadIdx.x == 0) carry_s = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < nblocks; b0 += kT) {
        const int i = b0 + threadIdx.x;
        const unsigned v = i < nblocks ? sums[i] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        unsigned pre = 0u;
        for (int q = 0; q < w; ++q) pre += wsum[q];
        const unsigned carry = carry_s;
        if (i < nblocks) sums[i] = carry + pre + inc - v;
        __syncthreads();
        if (threadIdx.x == kT - 1) carry_s = carry + pre + inc;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kT)
k_scan_add(unsigned *__restrict__ a, long n, const unsigned *__restrict__ sums)
{
    hpb_pdl_prologue();
    const long base = (long)blockIdx.x * kScanBlock + threadIdx.x * 4;
    const unsigned add = sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (base + k < n) a[base + k] += add;
}

struct SoA { double *r[HPB_PLASMA_NREAL]; uint64_t *idcpu; };

__global__ void __launch_bounds__(kT)
k_reorder_scatter(SoA in, SoA out, long np, const unsigned *__restrict__ key, const unsigned *__restrict__ rank,
                  const unsigned *__restrict__ offs)
{
    hpb_pdl_prologue();
    const long ip = (long)blockIdx.x * kT + threadIdx.x;
    if (ip >= np) return;
    const long dst = (long)offs[key[ip]] + rank[ip];
    double v[HPB_PLASMA_NREAL];
#pragma unroll
    for (int k = 0; k < HPB_PLASMA_NREAL; ++k) v[k] = __ldcs(&in.r[k][ip]);
    const unsigned long long id = __ldcs((const unsigned long long *)&in.idcpu[ip]);
#pragma unroll
    for (int k = 0; k < HPB_PLASMA_NREAL; ++k) out.r[k][dst] = v[k];
    out.idcpu[dst] = (uint64_t)id;
}

}  // namespace

// Sorts `in` into `out` (same capacity, caller-owned; the two must not alias).  idx_type: 0 = by
// cell, 1 = by node (bins shifted by half a cell), per direction (reorder_idx_type).  Scratch of the
// context grows on demand.  prob_lo: physical lower corner of the box.
extern "C" int hpb_plasma_reorder(hpb_ctx *ctx, hpb_plasma in, hpb_plasma out, double prob_lo_x, double prob_lo_y,
                                  int idx_type_x, int idx_type_y)
{
    if (!ctx || in.np != out.np) return HPB_ERR_ARG;
    const long np = in.np;
    if (np == 0) return HPB_OK;
    const hpb_geom &g = ctx->g;
    ReorderGeom rg;
    rg.x_lo = prob_lo_x; rg.y_lo = prob_lo_y; rg.dx_inv = 1.0 / g.dx; rg.dy_inv = 1.0 / g.dy;
    rg.shift_x = idx_type_x ? 0.5 : 0.; rg.shift_y = idx_type_y ? 0.5 : 0.;
    rg.nbx = g.nx + (idx_type_x ? 1 : 0); rg.nby = g.ny + (idx_type_y ? 1 : 0);
    const long nbins = (long)rg.nbx * rg.nby + 1;
    const int nblk = (int)((nbins + kScanBlock - 1) / kScanBlock);
    if (np > ctx->reorder_np_cap) {
        cudaFree(ctx->d_reorder_key); cudaFree(ctx->d_reorder_rank);
        ctx->d_reorder_key = ctx->d_reorder_rank = nullptr; ctx->reorder_np_cap = 0;
        HPB_CUDA_CHECK(cudaMalloc(&ctx->d_reorder_key, sizeof(unsigned) * np));
        HPB_CUDA_CHECK(cudaMalloc(&ctx->d_reorder_rank, sizeof(unsigned) * np));
        ctx->reorder_np_cap = np;
    }
    if (nbins > ctx->reorder_bins_cap) {
        cudaFree(ctx->d_reorder_hist); cudaFree(ctx->d_reorder_sums);
        ctx->d_reorder_hist = ctx->d_reorder_sums = nullptr; ctx->reorder_bins_cap = 0;
        HPB_CUDA_CHECK(cudaMalloc(&ctx->d_reorder_hist, sizeof(unsigned) * nbins));
        HPB_CUDA_CHECK(cudaMalloc(&ctx->d_reorder_sums, sizeof(unsigned) * (nblk + 1)));
        ctx->reorder_bins_cap = nbins;
    }
    HPB_CUDA_CHECK(cudaMemsetAsync(ctx->d_reorder_hist, 0, sizeof(unsigned) * nbins, ctx->stream));
    const unsigned nbp = (unsigned)((np + kT - 1) / kT);
    hpb_launch(k_reorder_key, nbp, kT, 0, ctx->stream, (const double *)in.r[HPB_X], (const double *)in.r[HPB_Y],
               (const uint64_t *)in.idcpu, np, rg, ctx->d_reorder_key, ctx->d_reorder_rank, ctx->d_reorder_hist);
    hpb_launch(k_scan_blocks, (unsigned)nblk, kT, 0, ctx->stream, ctx->d_reorder_hist, nbins, ctx->d_reorder_sums);
    hpb_launch(k_scan_sums, 1u, kT, 0, ctx->stream, ctx->d_reorder_sums, nblk);
    hpb_launch(k_scan_add, (unsigned)nblk, kT, 0, ctx->stream, ctx->d_reorder_hist, nbins,
               (const unsigned *)ctx->d_reorder_sums);
    SoA a, b;
    for (int k = 0; k < HPB_PLASMA_NREAL; ++k) { a.r[k] = in.r[k]; b.r[k] = out.r[k]; }
    a.idcpu = in.idcpu; b.idcpu = out.idcpu;
    hpb_launch(k_reorder_scatter, nbp, kT, 0, ctx->stream, a, b, np, (const unsigned *)ctx->d_reorder_key,
               (const unsigned *)ctx->d_reorder_rank, (const unsigned *)ctx->d_reorder_hist);
    hpb_count_launch(ctx, 5);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

void hpb_reorder_free(hpb_ctx *ctx)
{
    cudaFree(ctx->d_reorder_key); cudaFree(ctx->d_reorder_rank);
    cudaFree(ctx->d_reorder_hist); cudaFree(ctx->d_reorder_sums);
}
