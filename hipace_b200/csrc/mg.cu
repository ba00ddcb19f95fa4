// Geometric multigrid for the explicit Bx/By solve:  laplace(B) - chi*B = S  for the two
// components (Bx,By) <- (Sy,Sx) sharing the coefficient chi, homogeneous Dirichlet.
// Same algorithm, smoother count/order, transfer operators and stopping rule as
// hpmg::MultiGrid system type 1 (src/mg_solver/HpMultiGrid.cpp: gs1 :265-292, residual1 :184-190,
// restrict :29-52, interp :88-121, solve_doit :1307-1427, vcycle :1429-1512, bottomsolve
// :1514-1594) because the answer depends on where the iteration stops (tol_rel = 1e-4).
// Compiled with -fmad=false so that every level operator is bit-identical to the oracle.
//
// Kernels: one shared-memory tile kernel does 4 red-black Gauss-Seidel half-sweeps (+ residual)
// per launch: tile 64x32 with a 4/5-cell halo so the shrunken interior is exact.
#include "common.cuh"
#include <math.h>
#include <float.h>

namespace {

struct V2 {              // 2-component view over a level box, origin = level index (0,0)
    double *p;
    long rs, cs;         // row stride, component stride
    __device__ __forceinline__ double &at(int i, int j, int n) const { return p[i + (long)j * rs + n * cs]; }
};

struct LevelGeom {
    int nx, ny;          // points of the level box
    int vlo, vhix, vhiy; // valid index range [vlo, vhix] x [vlo, vhiy] (cc: 0..n-1, nodal: 1..n-2)
    int cc;
    double facx, facy;
};

constexpr int CX = 64, CY = 32;       // compute region of a tile
constexpr int AX = CX + 2, AY = CY + 2;
constexpr int NT = 256;

__device__ __forceinline__ void gs1(double *ph, int li, int lj, int i, int j, const LevelGeom &g,
                                    double rhs, double acf)
{
    // ph: smem plane [AY][AX]; (li, lj) local index of cell (i, j)
    double lap;
    double c0 = -(acf + 2.0 * (g.facx + g.facy));
    const double *c = ph + lj * AX + li;
    if (g.cc && i == g.vlo) {
        lap = g.facx * (4. / 3.) * c[1];
        c0 -= 2.0 * g.facx;
    } else if (g.cc && i == g.vhix) {
        lap = g.facx * (4. / 3.) * c[-1];
        c0 -= 2.0 * g.facx;
    } else {
        lap = g.facx * (c[-1] + c[1]);
    }
    if (g.cc && j == g.vlo) {
        lap += g.facy * (4. / 3.) * c[AX];
        c0 -= 2.0 * g.facy;
    } else if (g.cc && j == g.vhiy) {
        lap += g.facy * (4. / 3.) * c[-AX];
        c0 -= 2.0 * g.facy;
    } else {
        lap += g.facy * (c[-AX] + c[AX]);
    }
    const double c0_inv = 1.0 / c0;
    ph[lj * AX + li] = (rhs - lap) * c0_inv;
}

__device__ __forceinline__ double residual1(const double *ph, int li, int lj, int i, int j,
                                            const LevelGeom &g, double rhs, double acf)
{
    const double *c = ph + lj * AX + li;
    double lap = -2.0 * (g.facx + g.facy) * c[0];
    if (g.cc && i == g.vlo) lap += g.facx * ((4. / 3.) * c[1] - 2.0 * c[0]);
    else if (g.cc && i == g.vhix) lap += g.facx * ((4. / 3.) * c[-1] - 2.0 * c[0]);
    else lap += g.facx * (c[-1] + c[1]);
    if (g.cc && j == g.vlo) lap += g.facy * ((4. / 3.) * c[AX] - 2.0 * c[0]);
    else if (g.cc && j == g.vhiy) lap += g.facy * ((4. / 3.) * c[-AX] - 2.0 * c[0]);
    else lap += g.facy * (c[-AX] + c[AX]);
    return rhs + acf * c[0] - lap;
}

// phi_out = GSRB^4(phi_in or 0); optionally res = rhs + acf*phi - lap(phi)
template <bool ZERO_INIT, bool DO_RES>
__global__ void __launch_bounds__(NT)
k_gsrb4(LevelGeom g, V2 phi_in, V2 rhs, const double *__restrict__ acf, long acf_rs, V2 phi_out,
        V2 res, int nbx)
{
    constexpr int EO = DO_RES ? 4 : 3;            // edge offset (HpMultiGrid.cpp:427)
    constexpr int FX = CX - 2 * EO, FY = CY - 2 * EO;
    __shared__ double sm[2][AY * AX];

    const int bx = blockIdx.x % nbx, by = blockIdx.x / nbx;
    // local (1,1) is compute cell (0,0) of the tile; final region starts EO further in
    const int ox = bx * FX - EO + g.vlo;          // global index of compute cell 0
    const int oy = by * FY - EO + g.vlo;

    for (int s = threadIdx.x; s < AX * AY; s += NT) {
        const int lj = s / AX, li = s - lj * AX;
        const int i = ox - 1 + li, j = oy - 1 + lj;
        double v0 = 0., v1 = 0.;
        if (!ZERO_INIT && i >= g.vlo && i <= g.vhix && j >= g.vlo && j <= g.vhiy) {
            v0 = phi_in.at(i, j, 0);
            v1 = phi_in.at(i, j, 1);
        }
        sm[0][s] = v0;
        sm[1][s] = v1;
    }
    // every thread owns 4 vertical pairs: columns tx, tx+32; row pairs 2*(ty + 8b)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    double r0[4][2], r1[4][2], ac[4][2];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int ci = tx + 32 * (p & 1), cj = 2 * (ty + 8 * (p >> 1));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = ox + ci, j = oy + cj + h;
            const bool ok = i >= g.vlo && i <= g.vhix && j >= g.vlo && j <= g.vhiy;
            r0[p][h] = ok ? rhs.at(i, j, 0) : 0.;
            r1[p][h] = ok ? rhs.at(i, j, 1) : 0.;
            ac[p][h] = ok ? acf[i + (long)j * acf_rs] : 0.;
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int icolor = 0; icolor < 4; ++icolor) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int ci = tx + 32 * (p & 1), cj = 2 * (ty + 8 * (p >> 1));
            const int i = ox + ci;
            const int sh = (i + oy + cj + icolor) & 1;
            const int j = oy + cj + sh;
            if (i >= g.vlo && i <= g.vhix && j >= g.vlo && j <= g.vhiy) {
                const double a = sh ? ac[p][1] : ac[p][0];
                gs1(sm[0], ci + 1, cj + sh + 1, i, j, g, sh ? r0[p][1] : r0[p][0], a);
                gs1(sm[1], ci + 1, cj + sh + 1, i, j, g, sh ? r1[p][1] : r1[p][0], a);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int ci = tx + 32 * (p & 1), cj = 2 * (ty + 8 * (p >> 1));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = ox + ci, j = oy + cj + h;
            if (i >= g.vlo && i <= g.vhix && j >= g.vlo && j <= g.vhiy &&
                ci >= EO && ci < CX - EO && cj + h >= EO && cj + h < CY - EO) {
                if (DO_RES) {
                    res.at(i, j, 0) = residual1(sm[0], ci + 1, cj + h + 1, i, j, g, r0[p][h], ac[p][h]);
                    res.at(i, j, 1) = residual1(sm[1], ci + 1, cj + h + 1, i, j, g, r1[p][h], ac[p][h]);
                }
                phi_out.at(i, j, 0) = sm[0][(cj + h + 1) * AX + ci + 1];
                phi_out.at(i, j, 1) = sm[1][(cj + h + 1) * AX + ci + 1];
            }
        }
    }
}

// plain red-black half-sweeps in global memory by a single CTA (coarsest level: <= 5x5 points)
__global__ void k_bottom(LevelGeom g, V2 phi, V2 rhs, const double *acf, int nsweeps)
{
    const int nvx = g.vhix - g.vlo + 1, nvy = g.vhiy - g.vlo + 1;
    for (int s = threadIdx.x; s < g.nx * g.ny; s += blockDim.x) {
        phi.p[s] = 0.;
        phi.p[s + phi.cs] = 0.;
    }
    __syncthreads();
    for (int is = 0; is < nsweeps; ++is) {
        for (int s = threadIdx.x; s < nvx * nvy; s += blockDim.x) {
            const int j = s / nvx + g.vlo, i = s % nvx + g.vlo;
            if (((i + j + is) & 1) == 0) {
                const double a = acf[i + (long)j * g.nx];
                for (int n = 0; n < 2; ++n) {
                    auto P = [&](int ii, int jj) -> double {
                        // nodal boundary nodes are stored (and stay) zero; cc never reads outside
                        return phi.at(ii, jj, n);
                    };
                    double lap;
                    double c0 = -(a + 2.0 * (g.facx + g.facy));
                    if (g.cc && i == g.vlo) { lap = g.facx * (4. / 3.) * P(i + 1, j); c0 -= 2.0 * g.facx; }
                    else if (g.cc && i == g.vhix) { lap = g.facx * (4. / 3.) * P(i - 1, j); c0 -= 2.0 * g.facx; }
                    else lap = g.facx * (P(i - 1, j) + P(i + 1, j));
                    if (g.cc && j == g.vlo) { lap += g.facy * (4. / 3.) * P(i, j + 1); c0 -= 2.0 * g.facy; }
                    else if (g.cc && j == g.vhiy) { lap += g.facy * (4. / 3.) * P(i, j - 1); c0 -= 2.0 * g.facy; }
                    else lap += g.facy * (P(i, j - 1) + P(i, j + 1));
                    const double c0_inv = 1.0 / c0;
                    phi.at(i, j, n) = (rhs.at(i, j, n) - lap) * c0_inv;
                }
            }
        }
        __syncthreads();
    }
}

// crse = R(fine) over the valid coarse points, ncomp components
__global__ void k_restrict(LevelGeom gc, V2 crse, V2 fine, int ncomp)
{
    const int nvx = gc.vhix - gc.vlo + 1;
    const long nv = (long)nvx * (gc.vhiy - gc.vlo + 1);
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nv) return;
    const int j = (int)(s / nvx) + gc.vlo, i = (int)(s % nvx) + gc.vlo;
    for (int n = 0; n < ncomp; ++n) {
        if (gc.cc) {
            crse.at(i, j, n) = 0.25 * (fine.at(2 * i, 2 * j, n) + fine.at(2 * i + 1, 2 * j, n)
                                       + fine.at(2 * i, 2 * j + 1, n) + fine.at(2 * i + 1, 2 * j + 1, n));
        } else {
            crse.at(i, j, n) = (1. / 16.) * (fine.at(2 * i - 1, 2 * j - 1, n)
                               + 2. * fine.at(2 * i, 2 * j - 1, n)
                               + fine.at(2 * i + 1, 2 * j - 1, n)
                               + 2. * fine.at(2 * i - 1, 2 * j, n)
                               + 4. * fine.at(2 * i, 2 * j, n)
                               + 2. * fine.at(2 * i + 1, 2 * j, n)
                               + fine.at(2 * i - 1, 2 * j + 1, n)
                               + 2. * fine.at(2 * i, 2 * j + 1, n)
                               + fine.at(2 * i + 1, 2 * j + 1, n));
        }
    }
}

// fine_out = fine_in + I(crse) over the valid fine points
__global__ void k_interp_add(LevelGeom gf, V2 fine_in, V2 crse, V2 fine_out)
{
    const int nvx = gf.vhix - gf.vlo + 1;
    const long nv = (long)nvx * (gf.vhiy - gf.vlo + 1);
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nv) return;
    const int j = (int)(s / nvx) + gf.vlo, i = (int)(s % nvx) + gf.vlo;
    const int ic = i >> 1, jc = j >> 1;
    for (int n = 0; n < 2; ++n) {
        double add;
        if (gf.cc) {
            add = crse.at(ic, jc, n);
        } else {
            const bool io = (ic * 2 != i), jo = (jc * 2 != j);
            if (io && jo) add = (crse.at(ic, jc, n) + crse.at(ic + 1, jc, n) + crse.at(ic, jc + 1, n)
                                 + crse.at(ic + 1, jc + 1, n)) * 0.25;
            else if (io) add = (crse.at(ic, jc, n) + crse.at(ic + 1, jc, n)) * 0.5;
            else if (jo) add = (crse.at(ic, jc, n) + crse.at(ic, jc + 1, n)) * 0.5;
            else add = crse.at(ic, jc, n);
        }
        fine_out.at(i, j, n) = fine_in.at(i, j, n) + add;
    }
}

__global__ void k_copy_acf(LevelGeom g, double *dst, const double *src, long src_rs)
{
    const int nvx = g.vhix - g.vlo + 1;
    const long nv = (long)nvx * (g.vhiy - g.vlo + 1);
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nv) return;
    const int j = (int)(s / nvx) + g.vlo, i = (int)(s % nvx) + g.vlo;
    dst[i + (long)j * g.nx] = src[i + (long)j * src_rs];
}

// max|a| (2 comps) and optionally max|b| over the valid box -> out[0], out[1] (uint64 atomicMax
// on the bit pattern of non-negative doubles: order-independent, exact)
__global__ void k_norms(LevelGeom g, V2 a, V2 b, int do_b, double *out)
{
    __shared__ double sa[256], sb[256];
    const int nvx = g.vhix - g.vlo + 1;
    const long nv = (long)nvx * (g.vhiy - g.vlo + 1);
    double ma = 0., mb = 0.;
    for (long s = (long)blockIdx.x * blockDim.x + threadIdx.x; s < nv; s += (long)gridDim.x * blockDim.x) {
        const int j = (int)(s / nvx) + g.vlo, i = (int)(s % nvx) + g.vlo;
        ma = fmax(ma, fmax(fabs(a.at(i, j, 0)), fabs(a.at(i, j, 1))));
        if (do_b) mb = fmax(mb, fmax(fabs(b.at(i, j, 0)), fabs(b.at(i, j, 1))));
    }
    sa[threadIdx.x] = ma; sb[threadIdx.x] = mb;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) {
            sa[threadIdx.x] = fmax(sa[threadIdx.x], sa[threadIdx.x + st]);
            sb[threadIdx.x] = fmax(sb[threadIdx.x], sb[threadIdx.x + st]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicMax((unsigned long long *)&out[0], (unsigned long long)__double_as_longlong(sa[0]));
        if (do_b) atomicMax((unsigned long long *)&out[1], (unsigned long long)__double_as_longlong(sb[0]));
    }
}

__global__ void k_copy2(LevelGeom g, V2 dst, V2 src)
{
    const int nvx = g.vhix - g.vlo + 1;
    const long nv = (long)nvx * (g.vhiy - g.vlo + 1);
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nv) return;
    const int j = (int)(s / nvx) + g.vlo, i = (int)(s % nvx) + g.vlo;
    dst.at(i, j, 0) = src.at(i, j, 0);
    dst.at(i, j, 1) = src.at(i, j, 1);
}

LevelGeom level_geom(const hpb_ctx *ctx, int l)
{
    LevelGeom g;
    g.nx = ctx->mg[l].nx; g.ny = ctx->mg[l].ny;
    g.cc = ctx->mg_cc;
    g.vlo = g.cc ? 0 : 1;
    g.vhix = g.cc ? g.nx - 1 : g.nx - 2;
    g.vhiy = g.cc ? g.ny - 1 : g.ny - 2;
    const double dx = ctx->g.dx * (double)(1 << l), dy = ctx->g.dy * (double)(1 << l);
    g.facx = 1.0 / (dx * dx);
    g.facy = 1.0 / (dy * dy);
    return g;
}

V2 lvl_view(const hpb_ctx *ctx, int l, double *p)
{
    V2 v; v.p = p; v.rs = ctx->mg[l].nx; v.cs = (long)ctx->mg[l].nx * ctx->mg[l].ny; return v;
}

template <bool Z, bool R>
void launch_gsrb4(hpb_ctx *ctx, const LevelGeom &g, V2 in, V2 rhs, const double *acf, long acf_rs,
                  V2 out, V2 res)
{
    constexpr int EO = R ? 4 : 3;
    const int FX = CX - 2 * EO, FY = CY - 2 * EO;
    const int nvx = g.vhix - g.vlo + 1, nvy = g.vhiy - g.vlo + 1;
    const int nbx = (nvx + FX - 1) / FX, nby = (nvy + FY - 1) / FY;
    k_gsrb4<Z, R><<<nbx * nby, NT, 0, ctx->stream>>>(g, in, rhs, acf, acf_rs, out, res, nbx);
    hpb_count_launch(ctx);
}

inline unsigned nb(long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

int hpb_mg_init(hpb_ctx *ctx)
{
    const int nx = ctx->g.nx, ny = ctx->g.ny;
    if ((nx % 2) != (ny % 2)) {
        hpb_set_error("hpmg: nx and ny must have the same parity (HpMultiGrid.cpp:1051-1052)");
        return HPB_ERR_UNSUPPORTED;
    }
    ctx->mg_cc = (nx % 2 == 0);
    int w = ctx->mg_cc ? nx : nx + 2, h = ctx->mg_cc ? ny : ny + 2;
    int nl = 0;
    while (true) {
        ctx->mg[nl].nx = w; ctx->mg[nl].ny = h;
        ++nl;
        bool ok;
        if (ctx->mg_cc) ok = (w % 2 == 0 && h % 2 == 0 && w >= 4 && h >= 4);   // coarsenable(2, min 2)
        else ok = ((w - 1) % 2 == 0 && (h - 1) % 2 == 0 && w >= 8 && h >= 8);  // nodal, min width 4
        if (!ok || nl >= 30) break;
        if (ctx->mg_cc) { w /= 2; h /= 2; } else { w = (w - 1) / 2 + 1; h = (h - 1) / 2 + 1; }
    }
    ctx->mg_nlev = nl;
    for (int l = 0; l < nl; ++l) {
        const size_t n = (size_t)ctx->mg[l].nx * ctx->mg[l].ny;
        HPB_CUDA_CHECK(cudaMalloc(&ctx->mg[l].acf, n * sizeof(double)));
        HPB_CUDA_CHECK(cudaMalloc(&ctx->mg[l].res, 2 * n * sizeof(double)));
        HPB_CUDA_CHECK(cudaMalloc(&ctx->mg[l].cor, 2 * n * sizeof(double)));
        HPB_CUDA_CHECK(cudaMalloc(&ctx->mg[l].rescor, 2 * n * sizeof(double)));
        HPB_CUDA_CHECK(cudaMemset(ctx->mg[l].acf, 0, n * sizeof(double)));
        HPB_CUDA_CHECK(cudaMemset(ctx->mg[l].res, 0, 2 * n * sizeof(double)));
        HPB_CUDA_CHECK(cudaMemset(ctx->mg[l].cor, 0, 2 * n * sizeof(double)));
        HPB_CUDA_CHECK(cudaMemset(ctx->mg[l].rescor, 0, 2 * n * sizeof(double)));
    }
    HPB_CUDA_CHECK(cudaMalloc(&ctx->d_mg_norm, 2 * sizeof(double)));
    HPB_CUDA_CHECK(cudaMallocHost(&ctx->h_mg_norm, 2 * sizeof(double)));
    return HPB_OK;
}

void hpb_mg_free(hpb_ctx *ctx)
{
    for (int l = 0; l < ctx->mg_nlev; ++l) {
        cudaFree(ctx->mg[l].acf); cudaFree(ctx->mg[l].res); cudaFree(ctx->mg[l].cor);
        cudaFree(ctx->mg[l].rescor);
    }
    cudaFree(ctx->d_mg_norm);
    cudaFreeHost(ctx->h_mg_norm);
}

static int mg_vcycle(hpb_ctx *ctx, V2 sol, V2 rhs0)
{
    const int nl = ctx->mg_nlev;
    for (int l = 0; l < nl - 1; ++l) {
        const LevelGeom g = level_geom(ctx, l);
        if (l > 0) {
            launch_gsrb4<true, true>(ctx, g, V2{}, lvl_view(ctx, l, ctx->mg[l].res), ctx->mg[l].acf,
                                     g.nx, lvl_view(ctx, l, ctx->mg[l].cor),
                                     lvl_view(ctx, l, ctx->mg[l].rescor));
        }
        const LevelGeom gc = level_geom(ctx, l + 1);
        const long nv = (long)(gc.vhix - gc.vlo + 1) * (gc.vhiy - gc.vlo + 1);
        k_restrict<<<nb(nv), 256, 0, ctx->stream>>>(gc, lvl_view(ctx, l + 1, ctx->mg[l + 1].res),
                                                    lvl_view(ctx, l, ctx->mg[l].rescor), 2);
        hpb_count_launch(ctx);
    }
    {
        const int l = nl - 1;
        const LevelGeom g = level_geom(ctx, l);
        int nsw = 16;
        const int mx = g.nx > g.ny ? g.nx : g.ny;
        if ((mx + 1) / 2 * 2 > nsw) nsw = (mx + 1) / 2 * 2;       // HpMultiGrid.cpp:1587
        k_bottom<<<1, 64, 0, ctx->stream>>>(g, lvl_view(ctx, l, ctx->mg[l].cor),
                                            lvl_view(ctx, l, ctx->mg[l].res), ctx->mg[l].acf, nsw);
        hpb_count_launch(ctx);
    }
    for (int l = nl - 2; l >= 0; --l) {
        const LevelGeom g = level_geom(ctx, l);
        const long nv = (long)(g.vhix - g.vlo + 1) * (g.vhiy - g.vlo + 1);
        k_interp_add<<<nb(nv), 256, 0, ctx->stream>>>(g, lvl_view(ctx, l, ctx->mg[l].cor),
                                                      lvl_view(ctx, l + 1, ctx->mg[l + 1].cor),
                                                      lvl_view(ctx, l, ctx->mg[l].rescor));
        hpb_count_launch(ctx);
        if (l == 0)
            launch_gsrb4<false, false>(ctx, g, lvl_view(ctx, 0, ctx->mg[0].rescor), rhs0,
                                       ctx->mg[0].acf, g.nx, sol, V2{});
        else
            launch_gsrb4<false, false>(ctx, g, lvl_view(ctx, l, ctx->mg[l].rescor),
                                       lvl_view(ctx, l, ctx->mg[l].res), ctx->mg[l].acf, g.nx,
                                       lvl_view(ctx, l, ctx->mg[l].cor), V2{});
    }
    const LevelGeom g0 = level_geom(ctx, 0);
    launch_gsrb4<false, true>(ctx, g0, sol, rhs0, ctx->mg[0].acf, g0.nx,
                              lvl_view(ctx, 0, ctx->mg[0].cor), lvl_view(ctx, 0, ctx->mg[0].rescor));
    return HPB_OK;
}

extern "C" int hpb_mg_solve1(hpb_ctx *ctx, hpb_slice sl, int c_sol, int c_rhs, int c_acf,
                             double tol_rel, double tol_abs, int max_iters, int *h_iters)
{
    if (!ctx || c_sol < 0 || c_rhs < 0 || c_acf < 0) return HPB_ERR_ARG;
    SliceView v = make_view(sl);
    // center_box (HpMultiGrid.H:168-175): cc level index (0,0) = cell (0,0);
    // nodal level index (0,0) = cell (-1,-1)
    const int sh = ctx->mg_cc ? 0 : -1;
    V2 sol{v.comp(c_sol) + v.idx(sh, sh), v.jstride, v.nstride};
    V2 rhs{v.comp(c_rhs) + v.idx(sh, sh), v.jstride, v.nstride};
    const double *chi = v.comp(c_acf) + v.idx(sh, sh);
    const int nl = ctx->mg_nlev;
    const LevelGeom g0 = level_geom(ctx, 0);
    const long nv0 = (long)(g0.vhix - g0.vlo + 1) * (g0.vhiy - g0.vlo + 1);

    // acf[0] <- chi, then average down (solve1 :1177-1187, average_down_acoef :1640-1700)
    k_copy_acf<<<nb(nv0), 256, 0, ctx->stream>>>(g0, ctx->mg[0].acf, chi, v.jstride);
    hpb_count_launch(ctx);
    for (int l = 1; l < nl; ++l) {
        const LevelGeom gc = level_geom(ctx, l);
        const long nv = (long)(gc.vhix - gc.vlo + 1) * (gc.vhiy - gc.vlo + 1);
        V2 c{ctx->mg[l].acf, gc.nx, 0}, f{ctx->mg[l - 1].acf, ctx->mg[l - 1].nx, 0};
        k_restrict<<<nb(nv), 256, 0, ctx->stream>>>(gc, c, f, 1);
        hpb_count_launch(ctx);
    }
    // cor0 = GSRB^4(sol), rescor0 = rhs - L(cor0)   (:1326-1327)
    launch_gsrb4<false, true>(ctx, g0, sol, rhs, ctx->mg[0].acf, g0.nx,
                              lvl_view(ctx, 0, ctx->mg[0].cor), lvl_view(ctx, 0, ctx->mg[0].rescor));
    HPB_CUDA_CHECK(cudaMemsetAsync(ctx->d_mg_norm, 0, 2 * sizeof(double), ctx->stream));
    unsigned nbn = nb(nv0); if (nbn > 592) nbn = 592;
    k_norms<<<nbn, 256, 0, ctx->stream>>>(g0, lvl_view(ctx, 0, ctx->mg[0].rescor), rhs, 1, ctx->d_mg_norm);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaMemcpyAsync(ctx->h_mg_norm, ctx->d_mg_norm, 2 * sizeof(double),
                                   cudaMemcpyDeviceToHost, ctx->stream));
    HPB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    const double resnorm0 = ctx->h_mg_norm[0], rhsnorm0 = ctx->h_mg_norm[1];
    const double max_norm = rhsnorm0 >= resnorm0 ? rhsnorm0 : resnorm0;
    const double res_target = fmax(tol_abs, fmax(tol_rel, 1.e-16) * max_norm);      // :1361
    int iters = 0;
    if (!(resnorm0 <= res_target)) {
        bool converged = false;
        for (int it = 0; it < max_iters; ++it) {
            mg_vcycle(ctx, sol, rhs);
            iters = it + 1;
            HPB_CUDA_CHECK(cudaMemsetAsync(ctx->d_mg_norm, 0, sizeof(double), ctx->stream));
            k_norms<<<nbn, 256, 0, ctx->stream>>>(g0, lvl_view(ctx, 0, ctx->mg[0].rescor), rhs, 0,
                                                  ctx->d_mg_norm);
            hpb_count_launch(ctx);
            HPB_CUDA_CHECK(cudaMemcpyAsync(ctx->h_mg_norm, ctx->d_mg_norm, sizeof(double),
                                           cudaMemcpyDeviceToHost, ctx->stream));
            HPB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            const double norminf = ctx->h_mg_norm[0];
            if (norminf <= res_target) { converged = true; break; }
            if (norminf > 1.e20 * max_norm || norminf != norminf) {
                hpb_set_error("hpmg failing so lets stop here (resid/max_norm = %g)", norminf / max_norm);
                return HPB_ERR_MG_DIVERGED;
            }
        }
        if (!converged) {
            hpb_set_error("hpmg failed to converge after %d iterations", max_iters);
            return HPB_ERR_MG_DIVERGED;
        }
    }
    // sol <- cor0 on the valid box (:1419-1426)
    k_copy2<<<nb(nv0), 256, 0, ctx->stream>>>(g0, sol, lvl_view(ctx, 0, ctx->mg[0].cor));
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    if (h_iters) *h_iters = iters;
    return HPB_OK;
}
