// Geometric multigrid for the explicit Bx/By solve:  laplace(B) - chi*B = S  for the two
// components (Bx,By) <- (Sy,Sx) sharing the coefficient chi, homogeneous Dirichlet.
// Same algorithm, smoother count/order, transfer operators and stopping rule as
// hpmg::MultiGrid system type 1 (src/mg_solver/HpMultiGrid.cpp: gs1 :265-292, residual1 :184-190,
// restrict :29-52, interp :88-121, solve_doit :1307-1427, vcycle :1429-1512, bottomsolve
// :1514-1594) because the answer depends on where the iteration stops (tol_rel = 1e-4).
// Compiled with -fmad=false so that every level operator rounds exactly like the oracle.
//
// B200 formulation: every level operator pair of the V-cycle is ONE shared-memory tile kernel
//   down:  cor = GSRB^4(0; res)            + residual + restriction -> res[l+1]      (k_smooth<0,1>)
//   up:    cor = GSRB^4(cor + I(cor[l+1]))  (interpolation fused into the tile load)   (k_smooth<2,0>)
//   top:   cor0 = GSRB^4(sol) + residual + restriction -> res[1] + max-norm            (k_smooth<1,1>)
// with a 4/5/6-cell halo so that the shrunken interior is exact, and all levels of <= 34x34
// points run inside a single 1024-thread CTA (k_coarse).  chi is used in place as the level-0
// coefficient.  A V-cycle is ~2 + 2*(tile levels) + 1 launches instead of ~40, and the level-0
// traffic drops from 24 to 15 planes.
#include "common.cuh"
#include <math.h>
#include <float.h>
#include <stdlib.h>

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
constexpr int NT = 512;
constexpr int kCoarseMax = 34;        // levels with <= 34 x 34 points run in k_coarse
constexpr int kCoarseThreads = 1024;

// residual1 (:184-190) with the boundary-modified laplacian (:163-182); ph: smem plane
__device__ __forceinline__ double residual_smem(const double *ph, int li, int lj, int i, int j,
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

// CG: load through the L2 only (ld.global.cg).  The persistent mid-level kernel (k_mid) reads arrays that
// CTAs on OTHER SMs wrote earlier in the same launch; the L1 is not coherent across SMs.
template <bool CG>
__device__ __forceinline__ double ldv(const double *p) { return CG ? __ldcg(p) : *p; }

// I(crse)(i, j): interpcpy_cc / interpcpy_nd (:88-121)
template <bool CG = false>
__device__ __forceinline__ double interp_at(const V2 &crse, int i, int j, int n, int cc)
{
    const int ic = i >> 1, jc = j >> 1;
    auto at = [&](int a, int b) { return ldv<CG>(&crse.at(a, b, n)); };
    if (cc) return at(ic, jc);
    const bool io = (ic * 2 != i), jo = (jc * 2 != j);
    if (io && jo) return (at(ic, jc) + at(ic + 1, jc) + at(ic, jc + 1) + at(ic + 1, jc + 1)) * 0.25;
    if (io) return (at(ic, jc) + at(ic + 1, jc)) * 0.5;
    if (jo) return (at(ic, jc) + at(ic, jc + 1)) * 0.5;
    return at(ic, jc);
}

// One tile: phi = GSRB^4(init); optional residual -> restricted into res_c (+ max-norms).
//   INIT 0: phi = 0;  1: phi = phi_in;  2: phi = phi_in + I(crse)
//   RES: rescor = rhs + acf*phi - lap(phi) is restricted on the fly (restrict_cc / restrict_nd)
//        into res_c; norm[0] = max|rescor| and norm[1] = max|rhs| if norm != nullptr.
// 512 threads per 64x32 tile: thread (tx, ty) owns the vertical cell pairs (tx, 2ty), (tx, 2ty+1)
// and the same 32 columns further right.  Of each pair exactly one cell has the colour of a
// half-sweep, and which one is fixed per thread (parity of tx + tile origin): the per-cell
// invariants (rhs, 1/c0, boundary weights, shared-memory offset) are sorted once into an "A"
// set (updated by colours 0, 2) and a "B" set (colours 1, 3), so that a half-sweep is 8 shared
// loads, 14 flops and 2 stores per cell and nothing else -- the kernel is instruction-issue
// bound, not bandwidth bound.
struct SmCell {
    double r0, r1, cinv, wx, wy;
    int o;              // shared-memory offset of the cell
    bool ok;
};

// RH: the tile is 64 x (32 RH) cells (RH = 2: 8 cells per thread -- the per-thread set-up cost
//     is amortised over twice the work).  NS: number of GSRB^4 groups applied back to back
//     (NS = 2 fuses the last smoother of a V-cycle, sol = GSRB^4(cor0 + I(cor1)), with the first
//     one of the next, cor0 = GSRB^4(sol), at the price of a 4-cell deeper halo).
// NTHR: 512 (two columns per thread) or 1024 (one column per thread: half the dependent work per
//       thread -- used on the small levels, whose launches are latency bound, not throughput bound)
// tile: index of the tile in the level's nbx x nby tiling (blockIdx.x of k_smooth; the persistent k_mid loops
// over tiles); sm_dyn: 2 x AYr x AX doubles of shared memory; is_done: the solve has converged (no-op)
template <int INIT, bool RES, int RH, int NS, int NTHR, bool CG = false>
__device__ __forceinline__ void
smooth_tile(const LevelGeom &g, const V2 &phi_in, const V2 &crse, const V2 &rhs, const double *__restrict__ acf,
            long acf_rs, const double *__restrict__ c0i_in, double *__restrict__ c0i_out, const V2 &phi_out,
            const LevelGeom &gc, const V2 &res_c, double *norm, int nbx, int EO, int is_done, int lean,
            double *sm_dyn, int tile)
{
    // c0i_in: plane of 1 / c0 (row stride g.nx) written by an earlier launch of this solve -- the
    // four fp64 divisions per thread are then loads; c0i_out: where the first launch stores it
    constexpr int TXN = NTHR / 16;              // threads along x: 32 or 64
    constexpr int CH = CX / TXN;                // column halves per thread: 2 or 1
    constexpr int CYr = CY * RH, AYr = CYr + 2, NP = CH * RH;
    double *const sm0 = sm_dyn, *const sm1 = sm_dyn + AYr * AX;
    const int FX = CX - 2 * EO, FY = CYr - 2 * EO;
    const int bx = tile % nbx, by = tile / nbx;
    const int ox = bx * FX - EO + g.vlo;          // level index of compute cell (0, 0)
    const int oy = by * FY - EO + g.vlo;
    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;     // ty = 0..15
    // the whole tile (with its ring) lies strictly inside the valid range: no bounds or
    // boundary-stencil cases anywhere (block-uniform)
    const bool inner = ox - 1 > g.vlo && ox + CX < g.vhix && oy - 1 > g.vlo && oy + CYr < g.vhiy;

    // ---- restriction of the residual held in shared memory (owned region) + the max-norms: the tail both
    //      tile paths share
    auto restrict_and_norm = [&](double nres, double nrhs) {
        // restriction of the owned region [X0, X0 + FX) x [Y0, Y0 + FY) (level indices)
        const int X0 = ox + EO, Y0 = oy + EO;
        // coarse points owned by this tile: cc: fine cells (2I, 2I+1); nodal: fine node 2I
        const int I0 = (X0 + 1) >> 1, J0 = (Y0 + 1) >> 1;
        const int I1 = (X0 + FX + 1) >> 1, J1 = (Y0 + FY + 1) >> 1;     // exclusive
        // thread -> coarse points (I0 + lane, J0 + warp + m NTHR/32): FX/2 <= 29 columns
        {
            const int I = I0 + (tid & 31);
            for (int J = J0 + (tid >> 5); J < J1; J += NTHR / 32) {
                if (I < I1 && I >= gc.vlo && I <= gc.vhix && J >= gc.vlo && J <= gc.vhiy) {
                    const int li = 2 * I - ox + 1, lj = 2 * J - oy + 1;     // smem index of fine (2I, 2J)
                    const long oc = I + (long)J * res_c.rs;
#pragma unroll
                    for (int n = 0; n < 2; ++n) {
                        const double *f = (n ? sm1 : sm0) + lj * AX + li;
                        double v;
                        if (g.cc) {
                            v = 0.25 * (f[0] + f[1] + f[AX] + f[AX + 1]);
                        } else {
                            v = (1. / 16.) * (f[-AX - 1] + 2. * f[-AX] + f[-AX + 1]
                                              + 2. * f[-1] + 4. * f[0] + 2. * f[1]
                                              + f[AX - 1] + 2. * f[AX] + f[AX + 1]);
                        }
                        res_c.p[oc + n * res_c.cs] = v;
                    }
                }
            }
        }
        if (norm) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                nres = fmax(nres, __shfl_xor_sync(0xffffffffu, nres, o));
                nrhs = fmax(nrhs, __shfl_xor_sync(0xffffffffu, nrhs, o));
            }
            if ((tid & 31) == 0 && (nres > 0. || nrhs > 0.)) {
                // bit pattern of non-negative doubles is monotone: exact, order-independent max
                atomicMax((unsigned long long *)&norm[0], (unsigned long long)__double_as_longlong(nres));
                atomicMax((unsigned long long *)&norm[1], (unsigned long long)__double_as_longlong(nrhs));
            }
        }
    };

    // ---- lean path for tiles strictly inside the valid range (85 % of the level-0 tiles): no bounds or
    //      boundary-stencil cases, the per-cell invariants are loaded straight into the colour sets (the
    //      general path below loads them per row and re-selects them with FSEL pairs in every half-sweep
    //      because both orders do not fit 64 registers), one warp per tile row in the load.  Every cell goes
    //      through exactly the operations of the general path in the same order: bit-identical results.
    if (RH == 1 && NS == 1 && NTHR == 512 && lean && inner) {
        // thread (lane = tx, warp = ty) owns rows 2 ty, 2 ty + 1 of columns tx and tx + 32; of each vertical
        // pair the cell with (i + j) even is updated by colours 0 and 2 (set A), the other by 1 and 3 (set B)
        const int hA = (ox + tx + oy) & 1;
        double rA0[2], rA1[2], cA[2], aA[2], rB0[2], rB1[2], cB[2], aB[2];
        const double f2 = 2.0 * (g.facx + g.facy);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int i = ox + tx + 32 * p;
            const int jA = oy + 2 * ty + hA, jB = oy + 2 * ty + (hA ^ 1);
            const long oa = i + (long)jA * rhs.rs, ob = i + (long)jB * rhs.rs;
            rA0[p] = ldv<CG>(rhs.p + oa); rA1[p] = ldv<CG>(rhs.p + oa + rhs.cs);
            rB0[p] = ldv<CG>(rhs.p + ob); rB1[p] = ldv<CG>(rhs.p + ob + rhs.cs);
            if (c0i_in) {
                cA[p] = c0i_in[i + (long)jA * g.nx]; cB[p] = c0i_in[i + (long)jB * g.nx];
                aA[p] = RES ? acf[i + (long)jA * acf_rs] : 0.;
                aB[p] = RES ? acf[i + (long)jB * acf_rs] : 0.;
            } else {
                aA[p] = acf[i + (long)jA * acf_rs]; aB[p] = acf[i + (long)jB * acf_rs];
                cA[p] = 1.0 / (-(aA[p] + f2));                          // gs1 :265-292, interior cell
                cB[p] = 1.0 / (-(aB[p] + f2));
            }
        }
        // tile rows ty, ty + 16, ty + 32 (< 34): columns lane, lane + 32 and (lanes 0, 1) 64 + lane
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
            const int lj = ty + 16 * rr;
            if (lj < AYr) {
                const int j = oy - 1 + lj;
#pragma unroll
                for (int cg = 0; cg < 3; ++cg) {
                    const int li = tx + 32 * cg;
                    if (li < AX) {
                        double v0 = 0., v1 = 0.;
                        if (INIT != 0) {
                            const int i = ox - 1 + li;
                            const long o = i + (long)j * phi_in.rs;
                            v0 = ldv<CG>(phi_in.p + o);
                            v1 = ldv<CG>(phi_in.p + o + phi_in.cs);
                            if (INIT == 2) {
                                v0 = v0 + interp_at<CG>(crse, i, j, 0, g.cc);
                                v1 = v1 + interp_at<CG>(crse, i, j, 1, g.cc);
                            }
                        }
                        sm0[lj * AX + li] = v0; sm1[lj * AX + li] = v1;
                    }
                }
            }
        }
        if (is_done) return;            // block-uniform; nothing has been written to global memory yet
        int oA[2], oB[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            oA[p] = (2 * ty + hA + 1) * AX + tx + 32 * p + 1;
            oB[p] = (2 * ty + (hA ^ 1) + 1) * AX + tx + 32 * p + 1;
        }
        __syncthreads();
#pragma unroll
        for (int icolor = 0; icolor < 4; ++icolor) {
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const bool b = (icolor & 1) != 0;
                double *c = sm0 + (b ? oB[p] : oA[p]);
                double *e = sm1 + (b ? oB[p] : oA[p]);
                const double lap0 = g.facx * (c[-1] + c[1]) + g.facy * (c[-AX] + c[AX]);
                const double lap1 = g.facx * (e[-1] + e[1]) + g.facy * (e[-AX] + e[AX]);
                c[0] = ((b ? rB0[p] : rA0[p]) - lap0) * (b ? cB[p] : cA[p]);
                e[0] = ((b ? rB1[p] : rA1[p]) - lap1) * (b ? cB[p] : cA[p]);
            }
            __syncthreads();
        }
        double nres = 0., nrhs = 0.;
        double qs0[2][2], qs1[2][2];            // residuals of [set A | B][p]
#pragma unroll
        for (int s = 0; s < 2; ++s) {
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int cix = tx + 32 * p, cjh = 2 * ty + (s ? (hA ^ 1) : hA);
                const int so = s ? oB[p] : oA[p];
                const double r0v = s ? rB0[p] : rA0[p], r1v = s ? rB1[p] : rA1[p];
                const double acv = s ? aB[p] : aA[p], civ = s ? cB[p] : cA[p];
                qs0[s][p] = qs1[s][p] = 0.;
                if (RES && cix >= EO - 1 && cix < CX - EO + 1 && cjh >= EO - 1 && cjh < CYr - EO + 1) {
                    // residual1 (:184-190), interior form -- the operation order of the general path
                    const double *c = sm0 + so, *e = sm1 + so;
                    double lap = -2.0 * (g.facx + g.facy) * c[0];
                    lap += g.facx * (c[-1] + c[1]);
                    lap += g.facy * (c[-AX] + c[AX]);
                    qs0[s][p] = r0v + acv * c[0] - lap;
                    lap = -2.0 * (g.facx + g.facy) * e[0];
                    lap += g.facx * (e[-1] + e[1]);
                    lap += g.facy * (e[-AX] + e[AX]);
                    qs1[s][p] = r1v + acv * e[0] - lap;
                }
                if (cix >= EO && cix < CX - EO && cjh >= EO && cjh < CYr - EO) {
                    const int i = ox + cix, j = oy + cjh;
                    const long o = i + (long)j * phi_out.rs;
                    phi_out.p[o] = sm0[so];
                    phi_out.p[o + phi_out.cs] = sm1[so];
                    if (c0i_out) c0i_out[i + (long)j * g.nx] = civ;
                    if (RES && norm) {
                        nres = fmax(nres, fmax(fabs(qs0[s][p]), fabs(qs1[s][p])));
                        nrhs = fmax(nrhs, fmax(fabs(r0v), fabs(r1v)));
                    }
                }
            }
        }
        if (!RES) return;
        __syncthreads();                   // everyone is done reading phi from shared memory
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            sm0[oA[p]] = qs0[0][p]; sm1[oA[p]] = qs1[0][p];
            sm0[oB[p]] = qs0[1][p]; sm1[oB[p]] = qs1[1][p];
        }
        __syncthreads();
        restrict_and_norm(nres, nrhs);
        return;
    }

    // ---- tile load (with ring), linear over the AYr x AX shared array; all loads are independent
    constexpr int NLD = (AYr * AX + NTHR - 1) / NTHR;
    double v0[NLD], v1[NLD];
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
        v0[k] = 0.; v1[k] = 0.;
        if (INIT != 0) {
            const int e = tid + k * NTHR;
            const int lj = e / AX, li = e - lj * AX;
            const int i = ox - 1 + li, j = oy - 1 + lj;
            if (e < AYr * AX && (inner || (i >= g.vlo && i <= g.vhix && j >= g.vlo && j <= g.vhiy))) {
                const long o = i + (long)j * phi_in.rs;
                v0[k] = ldv<CG>(phi_in.p + o);
                v1[k] = ldv<CG>(phi_in.p + o + phi_in.cs);
                if (INIT == 2) {
                    v0[k] = v0[k] + interp_at<CG>(crse, i, j, 0, g.cc);
                    v1[k] = v1[k] + interp_at<CG>(crse, i, j, 1, g.cc);
                }
            }
        }
    }
    // ---- per-cell invariants of the owned cells: p = (row half, column half), h = row of the pair
    const double fx43 = g.facx * (4. / 3.), fy43 = g.facy * (4. / 3.);
    double r0[NP][2], r1[NP][2], ac[NP][2], ci[NP][2], wx[NP][2], wy[NP][2];
    bool okc[NP][2];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int i = ox + tx + TXN * (p % CH);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = oy + 32 * (p / CH) + 2 * ty + h;
            const bool ok = inner || (i >= g.vlo && i <= g.vhix && j >= g.vlo && j <= g.vhiy);
            okc[p][h] = ok;
            const long o = i + (long)j * rhs.rs;
            r0[p][h] = ok ? ldv<CG>(rhs.p + o) : 0.;
            r1[p][h] = ok ? ldv<CG>(rhs.p + o + rhs.cs) : 0.;
            const bool xb = !inner && g.cc && (i == g.vlo || i == g.vhix);
            const bool yb = !inner && g.cc && (j == g.vlo || j == g.vhiy);
            if (c0i_in) {
                ci[p][h] = ok ? c0i_in[i + (long)j * g.nx] : 0.;
                ac[p][h] = (RES && ok) ? acf[i + (long)j * acf_rs] : 0.;
            } else {
                const double a = ok ? acf[i + (long)j * acf_rs] : 0.;
                ac[p][h] = a;
                double c0 = -(a + 2.0 * (g.facx + g.facy));             // gs1 :265-292
                if (xb) c0 -= 2.0 * g.facx;
                if (yb) c0 -= 2.0 * g.facy;
                ci[p][h] = 1.0 / c0;
            }
            // boundary cells (cell-centred): one neighbour is the zero halo, the other is
            // weighted 4/3 -- identical to the branches of gs1
            wx[p][h] = xb ? fx43 : g.facx;
            wy[p][h] = yb ? fy43 : g.facy;
        }
    }
    // (the rhs / acf loads above are in flight together with the tile loads)
    if (is_done) return;            // block-uniform; nothing has been written yet
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
        const int e = tid + k * NTHR;
        if (e < AYr * AX) { sm0[e] = v0[k]; sm1[e] = v1[k]; }
    }
    // colour 0 updates the cells with (i + j) even: h = par0 for every p
    const int par0 = (ox + tx + oy) & 1;
    SmCell A[NP], B[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int base = (32 * (p / CH) + 2 * ty + 1) * AX + tx + TXN * (p % CH) + 1;
        A[p].r0 = par0 ? r0[p][1] : r0[p][0];   B[p].r0 = par0 ? r0[p][0] : r0[p][1];
        A[p].r1 = par0 ? r1[p][1] : r1[p][0];   B[p].r1 = par0 ? r1[p][0] : r1[p][1];
        A[p].cinv = par0 ? ci[p][1] : ci[p][0]; B[p].cinv = par0 ? ci[p][0] : ci[p][1];
        A[p].wx = par0 ? wx[p][1] : wx[p][0];   B[p].wx = par0 ? wx[p][0] : wx[p][1];
        A[p].wy = par0 ? wy[p][1] : wy[p][0];   B[p].wy = par0 ? wy[p][0] : wy[p][1];
        A[p].ok = par0 ? okc[p][1] : okc[p][0]; B[p].ok = par0 ? okc[p][0] : okc[p][1];
        A[p].o = base + par0 * AX;              B[p].o = base + (par0 ^ 1) * AX;
    }
    __syncthreads();
#pragma unroll 1
    for (int is = 0; is < NS; ++is) {
#pragma unroll
        for (int icolor = 0; icolor < 4; ++icolor) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const SmCell &q = (icolor & 1) ? B[p] : A[p];
                if (q.ok) {
                    double *c = sm0 + q.o;
                    double *e = sm1 + q.o;
                    const double lap0 = q.wx * (c[-1] + c[1]) + q.wy * (c[-AX] + c[AX]);
                    const double lap1 = q.wx * (e[-1] + e[1]) + q.wy * (e[-AX] + e[AX]);
                    c[0] = (q.r0 - lap0) * q.cinv;
                    e[0] = (q.r1 - lap1) * q.cinv;
                }
            }
            __syncthreads();
        }
    }
    double rs0[NP][2], rs1[NP][2];
    double nres = 0., nrhs = 0.;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int cix = tx + TXN * (p % CH);
        const bool own_x = cix >= EO && cix < CX - EO;
        const bool ring_x = cix >= EO - 1 && cix < CX - EO + 1;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int cjh = 32 * (p / CH) + 2 * ty + h;
            const int i = ox + cix, j = oy + cjh;
            const bool ok = okc[p][h];
            const int so = (cjh + 1) * AX + cix + 1;
            rs0[p][h] = rs1[p][h] = 0.;
            // the residual is exact one ring further out than the owned region (needed by the
            // nodal full-weighting restriction)
            if (RES && ok && ring_x && cjh >= EO - 1 && cjh < CYr - EO + 1) {
                if (inner) {
                    // residual1 (:184-190), interior form -- same operation order as residual_smem
                    const double *c = sm0 + so, *e = sm1 + so;
                    double lap = -2.0 * (g.facx + g.facy) * c[0];
                    lap += g.facx * (c[-1] + c[1]);
                    lap += g.facy * (c[-AX] + c[AX]);
                    rs0[p][h] = r0[p][h] + ac[p][h] * c[0] - lap;
                    lap = -2.0 * (g.facx + g.facy) * e[0];
                    lap += g.facx * (e[-1] + e[1]);
                    lap += g.facy * (e[-AX] + e[AX]);
                    rs1[p][h] = r1[p][h] + ac[p][h] * e[0] - lap;
                } else {
                    rs0[p][h] = residual_smem(sm0, cix + 1, cjh + 1, i, j, g, r0[p][h], ac[p][h]);
                    rs1[p][h] = residual_smem(sm1, cix + 1, cjh + 1, i, j, g, r1[p][h], ac[p][h]);
                }
            }
            if (ok && own_x && cjh >= EO && cjh < CYr - EO) {
                const long o = i + (long)j * phi_out.rs;
                phi_out.p[o] = sm0[so];
                phi_out.p[o + phi_out.cs] = sm1[so];
                if (c0i_out) c0i_out[i + (long)j * g.nx] = ci[p][h];
                if (RES && norm) {
                    nres = fmax(nres, fmax(fabs(rs0[p][h]), fabs(rs1[p][h])));
                    nrhs = fmax(nrhs, fmax(fabs(r0[p][h]), fabs(r1[p][h])));
                }
            }
        }
    }
    if (!RES) return;
    __syncthreads();                   // everyone is done reading phi from shared memory
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int cix = tx + TXN * (p % CH);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int so = (32 * (p / CH) + 2 * ty + h + 1) * AX + cix + 1;
            sm0[so] = rs0[p][h];
            sm1[so] = rs1[p][h];
        }
    }
    __syncthreads();
    restrict_and_norm(nres, nrhs);
}

template <int INIT, bool RES, int RH, int NS, int NTHR>
__global__ void __launch_bounds__(NTHR, (RH == 2 || NTHR == 1024) ? 1 : 2)
k_smooth(LevelGeom g, V2 phi_in, V2 crse, V2 rhs, const double *__restrict__ acf, long acf_rs,
         const double *__restrict__ c0i_in, double *__restrict__ c0i_out, V2 phi_out, LevelGeom gc,
         V2 res_c, double *norm, int nbx, int EO, const int *done, int lean)
{
    extern __shared__ double sm_dyn[];
    hpb_pdl_prologue();
    // converged: the speculatively enqueued V-cycle is a no-op.  The flag is fetched here but only
    // tested after the tile loads have been issued, so its L2 round trip is not serialised.
    const int is_done = done ? *(const volatile int *)done : 0;
    smooth_tile<INIT, RES, RH, NS, NTHR>(g, phi_in, crse, rhs, acf, acf_rs, c0i_in, c0i_out, phi_out, gc, res_c, norm,
                                         nbx, EO, is_done, lean, sm_dyn, (int)blockIdx.x);
}

// ---- single-CTA part of the V-cycle: all levels with <= 34 x 34 points -------------------------
struct CoarseLevel {
    LevelGeom g;
    double *acf, *c0i, *res, *cor, *rescor;      // res/cor/rescor: 2 comps, comp stride nx*ny
};
struct CoarseArgs { int nl; int nsweeps_bottom; CoarseLevel L[12]; };

__device__ void c_gsrb(const CoarseLevel &L, double *phi, const double *rhs, int nsweeps)
{
    const LevelGeom &g = L.g;
    const int nvx = g.vhix - g.vlo + 1, nvy = g.vhiy - g.vlo + 1;
    const int ncell = nvx * nvy;
    const long cs = (long)g.nx * g.ny;
    const double fx43 = g.facx * (4. / 3.), fy43 = g.facy * (4. / 3.);
    for (int is = 0; is < nsweeps; ++is) {
        for (int s = threadIdx.x; s < 2 * ncell; s += blockDim.x) {
            const int n = s >= ncell, cell = s - n * ncell;
            const int j = cell / nvx + g.vlo, i = cell % nvx + g.vlo;
            if (((i + j + is) & 1) == 0) {
                const long o = i + (long)j * g.nx;
                const double *c = phi + n * cs + o;
                double lap;
                if (g.cc && i == g.vlo) lap = fx43 * c[1];
                else if (g.cc && i == g.vhix) lap = fx43 * c[-1];
                else lap = g.facx * (c[-1] + c[1]);
                if (g.cc && j == g.vlo) lap += fy43 * c[g.nx];
                else if (g.cc && j == g.vhiy) lap += fy43 * c[-g.nx];
                else lap += g.facy * (c[-g.nx] + c[g.nx]);
                phi[n * cs + o] = (rhs[n * cs + o] - lap) * L.c0i[o];
            }
        }
        __syncthreads();
    }
}

__device__ void c_residual(const CoarseLevel &L, double *res, const double *phi, const double *rhs)
{
    const LevelGeom &g = L.g;
    const int nvx = g.vhix - g.vlo + 1, nvy = g.vhiy - g.vlo + 1;
    const int ncell = nvx * nvy;
    const long cs = (long)g.nx * g.ny;
    for (int s = threadIdx.x; s < 2 * ncell; s += blockDim.x) {
        const int n = s >= ncell, cell = s - n * ncell;
        const int j = cell / nvx + g.vlo, i = cell % nvx + g.vlo;
        const long o = i + (long)j * g.nx;
        const double *c = phi + n * cs + o;
        double lap = -2.0 * (g.facx + g.facy) * c[0];
        if (g.cc && i == g.vlo) lap += g.facx * ((4. / 3.) * c[1] - 2.0 * c[0]);
        else if (g.cc && i == g.vhix) lap += g.facx * ((4. / 3.) * c[-1] - 2.0 * c[0]);
        else lap += g.facx * (c[-1] + c[1]);
        if (g.cc && j == g.vlo) lap += g.facy * ((4. / 3.) * c[g.nx] - 2.0 * c[0]);
        else if (g.cc && j == g.vhiy) lap += g.facy * ((4. / 3.) * c[-g.nx] - 2.0 * c[0]);
        else lap += g.facy * (c[-g.nx] + c[g.nx]);
        res[n * cs + o] = rhs[n * cs + o] + L.acf[o] * c[0] - lap;
    }
    __syncthreads();
}

// crse(ncomp comps) = R(fine)
__device__ void c_restrict(const LevelGeom &gc, double *crse, const LevelGeom &gf, const double *fine,
                           int ncomp)
{
    const int nvx = gc.vhix - gc.vlo + 1, nvy = gc.vhiy - gc.vlo + 1;
    const int ncell = nvx * nvy;
    const long ccs = (long)gc.nx * gc.ny, fcs = (long)gf.nx * gf.ny;
    for (int s = threadIdx.x; s < ncomp * ncell; s += blockDim.x) {
        const int n = s / ncell, cell = s - n * ncell;
        const int j = cell / nvx + gc.vlo, i = cell % nvx + gc.vlo;
        const double *f = fine + n * fcs + (2 * i) + (long)(2 * j) * gf.nx;
        const int w = gf.nx;
        double v;
        if (gc.cc) v = 0.25 * (f[0] + f[1] + f[w] + f[w + 1]);
        else v = (1. / 16.) * (f[-w - 1] + 2. * f[-w] + f[-w + 1] + 2. * f[-1] + 4. * f[0] + 2. * f[1]
                               + f[w - 1] + 2. * f[w] + f[w + 1]);
        crse[n * ccs + i + (long)j * gc.nx] = v;
    }
    __syncthreads();
}

__device__ void c_interp_add(const CoarseLevel &Lf, double *fine, const CoarseLevel &Lc, double *crse)
{
    const LevelGeom &g = Lf.g;
    const int nvx = g.vhix - g.vlo + 1, nvy = g.vhiy - g.vlo + 1;
    const int ncell = nvx * nvy;
    const long cs = (long)g.nx * g.ny;
    V2 cv{crse, Lc.g.nx, (long)Lc.g.nx * Lc.g.ny};
    for (int s = threadIdx.x; s < 2 * ncell; s += blockDim.x) {
        const int n = s >= ncell, cell = s - n * ncell;
        const int j = cell / nvx + g.vlo, i = cell % nvx + g.vlo;
        const long o = n * cs + i + (long)j * g.nx;
        fine[o] = fine[o] + interp_at(cv, i, j, n, g.cc);
    }
    __syncthreads();
}

__device__ void c_zero(double *p, long n)
{
    for (long s = threadIdx.x; s < n; s += blockDim.x) p[s] = 0.;
    __syncthreads();
}

// levels L[0..nl-1] (L[0] = first single-CTA level): in res[0]; out cor[0].
// All level arrays live in shared memory for the duration of the kernel.  The valid points of a
// level are mapped compactly onto the first nvx*nvy threads (point = (t % nvx, t / nvx)), which
// keep their rhs, 1/c0 and stencil weights in registers for the duration of a level visit: a
// half-sweep is 4 shared-memory loads, 7 flops and a barrier per component.  This part of the
// V-cycle is pure latency (~60 dependent phases), so every phase synchronises only the warps
// that own points of the level: a named barrier over T_l = ceil32(nvx*nvy) threads, a
// __syncwarp() for levels of <= 32 points (4x4 and 2x2: 27 of the 60 phases).  Threads without
// points on a level skip it and wait at the barrier of the next finer level.
struct SLevel {
    int nx, ny, vlo, vhix, vhiy, n;
    int nvx, npts, T;              // valid width, valid points, participating threads (multiple of 32)
    double facx, facy;
    int o_acf, o_c0i, o_res, o_cor, o_rescor;      // offsets into the dynamic shared array
};

__device__ __forceinline__ void level_bar(int l, int T)
{
    if (T <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(l + 1), "r"(T) : "memory");
}

struct SCell {         // per-thread view of its point on one level
    bool ok, hl, hr, hd, hu;       // has a left / right / lower / upper neighbour inside the level box
    int i, j, o, par;
    double r0, r1, c0i, cxw, cyw;
};

__device__ __forceinline__ SCell s_cell(const SLevel &L, const double *sm, int cc)
{
    SCell c;
    const int t = threadIdx.x;
    const int q = t / L.nvx;
    c.i = L.vlo + (t - q * L.nvx); c.j = L.vlo + q;
    c.ok = t < L.npts;
    const int i = c.i, j = c.j;
    c.o = i + j * L.nx;
    c.par = (i + j) & 1;
    c.r0 = c.r1 = c.c0i = 0.;
    c.cxw = (cc && (i == L.vlo || i == L.vhix)) ? L.facx * (4. / 3.) : L.facx;
    c.cyw = (cc && (j == L.vlo || j == L.vhiy)) ? L.facy * (4. / 3.) : L.facy;
    // cell-centred level arrays carry no ghost ring: the missing neighbour counts as 0
    c.hl = !(cc && i == L.vlo); c.hr = !(cc && i == L.vhix);
    c.hd = !(cc && j == L.vlo); c.hu = !(cc && j == L.vhiy);
    if (c.ok) { c.r0 = sm[L.o_res + c.o]; c.r1 = sm[L.o_res + L.n + c.o]; c.c0i = sm[L.o_c0i + c.o]; }
    return c;
}

// half-sweeps is = first..last-1 on cor (boundary / exterior points hold 0: the zero neighbour and
// the 4/3 weight reproduce the branches of gs1 exactly)
__device__ __forceinline__ void s_sweeps(const SLevel &L, int l, const SCell &c, double *sm, int first, int last)
{
    double *p0 = sm + L.o_cor + c.o, *p1 = p0 + L.n;
    const int w = L.nx;
    for (int is = first; is < last; ++is) {
        if (c.ok && ((c.par + is) & 1) == 0) {
            const double l0 = c.hl ? p0[-1] : 0., r0 = c.hr ? p0[1] : 0.;
            const double d0 = c.hd ? p0[-w] : 0., u0 = c.hu ? p0[w] : 0.;
            const double l1 = c.hl ? p1[-1] : 0., r1 = c.hr ? p1[1] : 0.;
            const double d1 = c.hd ? p1[-w] : 0., u1 = c.hu ? p1[w] : 0.;
            const double lap0 = c.cxw * (l0 + r0) + c.cyw * (d0 + u0);
            const double lap1 = c.cxw * (l1 + r1) + c.cyw * (d1 + u1);
            p0[0] = (c.r0 - lap0) * c.c0i;
            p1[0] = (c.r1 - lap1) * c.c0i;
        }
        level_bar(l, L.T);
    }
}

// cor = GSRB^n(0): the first half-sweep of a zero field is rhs / c0 on its colour
__device__ __forceinline__ void s_sweeps_from_zero(const SLevel &L, int l, const SCell &c, double *sm, int n)
{
    if (c.ok) {
        const bool first = (c.par & 1) == 0;
        sm[L.o_cor + c.o] = first ? (c.r0 - 0.) * c.c0i : 0.;
        sm[L.o_cor + L.n + c.o] = first ? (c.r1 - 0.) * c.c0i : 0.;
    }
    level_bar(l, L.T);
    s_sweeps(L, l, c, sm, 1, n);
}

__device__ __forceinline__ void s_residual(const SLevel &L, int l, const SCell &c, double *sm, int cc)
{
    if (c.ok) {
        const int w = L.nx;
        const int i = c.i, j = c.j;
        const double a = sm[L.o_acf + c.o];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            const double *p = sm + L.o_cor + n * L.n + c.o;
            double lap = -2.0 * (L.facx + L.facy) * p[0];
            if (cc && i == L.vlo) lap += L.facx * ((4. / 3.) * p[1] - 2.0 * p[0]);
            else if (cc && i == L.vhix) lap += L.facx * ((4. / 3.) * p[-1] - 2.0 * p[0]);
            else lap += L.facx * (p[-1] + p[1]);
            if (cc && j == L.vlo) lap += L.facy * ((4. / 3.) * p[w] - 2.0 * p[0]);
            else if (cc && j == L.vhiy) lap += L.facy * ((4. / 3.) * p[-w] - 2.0 * p[0]);
            else lap += L.facy * (p[-w] + p[w]);
            sm[L.o_rescor + n * L.n + c.o] = (n ? c.r1 : c.r0) + a * p[0] - lap;
        }
    }
    level_bar(l, L.T);
}

// res[coarse] = R(rescor[fine]); executed by the threads of the coarse level
__device__ __forceinline__ void s_restrict(const SLevel &Lc, int lc, const SLevel &Lf, double *sm, int cc)
{
    const int t = threadIdx.x;
    if (t < Lc.npts) {
        const int q = t / Lc.nvx;
        const int i = Lc.vlo + (t - q * Lc.nvx), j = Lc.vlo + q;
        const int w = Lf.nx;
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            const double *f = sm + Lf.o_rescor + n * Lf.n + 2 * i + 2 * j * w;
            double v;
            if (cc) v = 0.25 * (f[0] + f[1] + f[w] + f[w + 1]);
            else v = (1. / 16.) * (f[-w - 1] + 2. * f[-w] + f[-w + 1] + 2. * f[-1] + 4. * f[0] + 2. * f[1]
                                   + f[w - 1] + 2. * f[w] + f[w + 1]);
            sm[Lc.o_res + n * Lc.n + i + j * Lc.nx] = v;
        }
    }
    level_bar(lc, Lc.T);
}

// cor[fine] += I(cor[coarse])
__device__ __forceinline__ void s_interp_add(const SLevel &Lf, int l, const SCell &c, const SLevel &Lc, double *sm, int cc)
{
    if (c.ok) {
        const int i = c.i, j = c.j;
        const int ic = i >> 1, jc = j >> 1, wc = Lc.nx;
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            const double *q = sm + Lc.o_cor + n * Lc.n + ic + jc * wc;
            double add;
            if (cc) add = q[0];
            else {
                const bool io = (ic * 2 != i), jo = (jc * 2 != j);
                if (io && jo) add = (q[0] + q[1] + q[wc] + q[wc + 1]) * 0.25;
                else if (io) add = (q[0] + q[1]) * 0.5;
                else if (jo) add = (q[0] + q[wc]) * 0.5;
                else add = q[0];
            }
            double *f = sm + Lf.o_cor + n * Lf.n + c.o;
            f[0] = f[0] + add;
        }
    }
    level_bar(l, Lf.T);
}

// all levels of A in one 1024-thread CTA (csm: the dynamic shared array holding every level).  CG: the
// restricted residual of the first level was written by other SMs earlier in the same launch (k_mid)
template <bool CG>
__device__ __forceinline__ void coarse_body(const CoarseArgs &A, double *csm)
{
    __shared__ SLevel S[12];
    const int nl = A.nl;
    const int cc = A.L[0].g.cc;
    const int tid = threadIdx.x;
    if (tid == 0) {
        int off = 0;
        for (int l = 0; l < nl; ++l) {
            const LevelGeom &g = A.L[l].g;
            SLevel &L = S[l];
            L.nx = g.nx; L.ny = g.ny; L.vlo = g.vlo; L.vhix = g.vhix; L.vhiy = g.vhiy;
            L.n = g.nx * g.ny; L.facx = g.facx; L.facy = g.facy;
            L.nvx = g.vhix - g.vlo + 1;
            L.npts = L.nvx * (g.vhiy - g.vlo + 1);
            L.T = (L.npts + 31) / 32 * 32;
            L.o_acf = off; off += L.n;
            L.o_c0i = off; off += L.n;
            L.o_res = off; off += 2 * L.n;
            L.o_cor = off; off += 2 * L.n;
            L.o_rescor = off; off += 2 * L.n;
        }
    }
    __syncthreads();
    for (int l = 0; l < nl; ++l) {
        const SLevel L = S[l];
        for (int s = tid; s < L.n; s += blockDim.x) {
            csm[L.o_acf + s] = A.L[l].acf[s];
            csm[L.o_c0i + s] = A.L[l].c0i[s];
            csm[L.o_res + s] = l == 0 ? ldv<CG>(A.L[0].res + s) : 0.;             // boundary nodes stay 0
            csm[L.o_res + L.n + s] = l == 0 ? ldv<CG>(A.L[0].res + L.n + s) : 0.;
            csm[L.o_cor + s] = 0.;    csm[L.o_cor + L.n + s] = 0.;
            csm[L.o_rescor + s] = 0.; csm[L.o_rescor + L.n + s] = 0.;
        }
    }
    __syncthreads();
    // down: a thread takes part in level l while tid < T_l (T is non-increasing with l)
#pragma unroll 1
    for (int l = 0; l < nl - 1; ++l) {
        const SLevel L = S[l];
        if (tid >= L.T) break;
        const SCell c = s_cell(L, csm, cc);
        s_sweeps_from_zero(L, l, c, csm, 4);
        s_residual(L, l, c, csm, cc);
        if (tid < S[l + 1].T) s_restrict(S[l + 1], l + 1, L, csm, cc);
    }
    if (tid < S[nl - 1].T) {
        const SLevel L = S[nl - 1];
        const SCell c = s_cell(L, csm, cc);
        s_sweeps_from_zero(L, nl - 1, c, csm, A.nsweeps_bottom);
    }
    // up: the threads that sat out level l+1 meet the ones that worked on it at level l's barrier
#pragma unroll 1
    for (int l = nl - 2; l >= 0; --l) {
        const SLevel L = S[l];
        if (tid >= L.T) continue;
        level_bar(l, L.T);
        const SCell c = s_cell(L, csm, cc);
        s_interp_add(L, l, c, S[l + 1], csm, cc);
        s_sweeps(L, l, c, csm, 0, 4);
    }
    __syncthreads();
    {
        const SLevel L = S[0];
        for (int s = tid; s < 2 * L.n; s += blockDim.x) A.L[0].cor[s] = csm[L.o_cor + s];
    }
}

__global__ void __launch_bounds__(kCoarseThreads) k_coarse(CoarseArgs A, const int *done)
{
    extern __shared__ double csm[];
    hpb_pdl_prologue();
    if (*done) return;
    coarse_body<false>(A, csm);
}

// ---- the mid levels of a V-cycle in ONE persistent launch --------------------------------------------
// Levels whose tiling has at most one tile per SM are latency chains: every launch (down-stroke smoother,
// the single-CTA coarse part, up-stroke smoother) costs 8-10 us whatever its size, 14 launches per V-cycle
// at 1024^2.  k_mid runs the phases  down(l0) .. down(lc-1), coarse, up(lc-1) .. up(l0)  back to back in
// one cooperative launch of (largest tile count) CTAs with a grid-wide barrier between the phases; the
// phases are the SAME device functions the separate launches run (smooth_tile, coarse_body), so the result
// is bit-identical.  Arrays written in one phase and read in a later one by other SMs are read through the
// L2 (CG).
struct GridBar { unsigned count, gen; };
__device__ __forceinline__ void grid_barrier(GridBar *bar)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned gen = *(volatile unsigned *)&bar->gen;
        if (atomicAdd(&bar->count, 1u) == gridDim.x - 1) {
            bar->count = 0;
            __threadfence();
            atomicAdd(&bar->gen, 1u);
        } else {
            while (*(volatile unsigned *)&bar->gen == gen) { }
        }
        __threadfence();
    }
    __syncthreads();
}

struct MidLevel {
    LevelGeom g, gc;
    V2 res, cor, rescor, res_c, up_crse;       // res_c: res of the next coarser level; up_crse: what the up-stroke interpolates
    const double *acf, *c0i;
    int nbx_d, nt_d, EO_d, nbx_u, nt_u, EO_u;  // tiling of the down / up smoother (different halo widths)
};
struct MidArgs { int nlev; MidLevel L[6]; CoarseArgs coarse; };

__global__ void __launch_bounds__(kCoarseThreads, 1) k_mid(MidArgs A, const int *done, GridBar *bar, int lean)
{
    extern __shared__ double msm[];
    hpb_pdl_prologue();
    if (*(const volatile int *)done) return;           // uniform over the grid: the barrier is never entered
    const V2 none{};
    const LevelGeom gnone{};
    for (int l = 0; l < A.nlev; ++l) {                  // down: cor = GSRB^4(0; res), res_c = R(residual)
        const MidLevel &L = A.L[l];
        for (int t = blockIdx.x; t < L.nt_d; t += gridDim.x) {
            smooth_tile<0, true, 1, 1, kCoarseThreads, true>(L.g, none, none, L.res, L.acf, L.g.nx, L.c0i, nullptr, L.cor,
                                                             L.gc, L.res_c, nullptr, L.nbx_d, L.EO_d, 0, lean, msm, t);
            __syncthreads();
        }
        grid_barrier(bar);
    }
    if (blockIdx.x == 0) coarse_body<true>(A.coarse, msm);
    grid_barrier(bar);
    for (int l = A.nlev - 1; l >= 0; --l) {             // up: rescor = GSRB^4(cor + I(up_crse))
        const MidLevel &L = A.L[l];
        for (int t = blockIdx.x; t < L.nt_u; t += gridDim.x) {
            smooth_tile<2, false, 1, 1, kCoarseThreads, true>(L.g, L.cor, L.up_crse, L.res, L.acf, L.g.nx, L.c0i, nullptr,
                                                              L.rescor, gnone, none, nullptr, L.nbx_u, L.EO_u, 0, lean, msm, t);
            __syncthreads();
        }
        if (l > 0) grid_barrier(bar);
    }
}

// fallback when the single-CTA levels do not fit in shared memory: same algorithm on the global
// arrays (coherent within one CTA after __syncthreads)
__global__ void __launch_bounds__(kCoarseThreads) k_coarse_global(CoarseArgs A, const int *done)
{
    hpb_pdl_prologue();
    if (*done) return;
    const int nl = A.nl;
    for (int l = 0; l < nl - 1; ++l) {
        const CoarseLevel &L = A.L[l];
        const long n = (long)L.g.nx * L.g.ny;
        c_zero(L.cor, 2 * n);
        c_gsrb(L, L.cor, L.res, 4);
        c_residual(L, L.rescor, L.cor, L.res);
        c_restrict(A.L[l + 1].g, A.L[l + 1].res, L.g, L.rescor, 2);
    }
    {
        const CoarseLevel &L = A.L[nl - 1];
        c_zero(L.cor, 2 * (long)L.g.nx * L.g.ny);
        c_gsrb(L, L.cor, L.res, A.nsweeps_bottom);
    }
    for (int l = nl - 2; l >= 0; --l) {
        c_interp_add(A.L[l], A.L[l].cor, A.L[l + 1], A.L[l + 1].cor);
        c_gsrb(A.L[l], A.L[l].cor, A.L[l].res, 4);
    }
}

// average_down_acoef (:1640-1700) for the single-CTA levels + their 1/c0 tables.
// L[0].acf must already hold the restriction from the last tile level (or be level 1's input).
__global__ void __launch_bounds__(kCoarseThreads) k_coarse_setup(CoarseArgs A, LevelGeom gfine,
                                                                  const double *acf_fine, long fine_rs)
{
    hpb_pdl_prologue();
    // restrict into L[0] from the finer (tile) level with an arbitrary row stride
    {
        const LevelGeom &gc = A.L[0].g;
        const int nvx = gc.vhix - gc.vlo + 1, nvy = gc.vhiy - gc.vlo + 1;
        for (int s = threadIdx.x; s < nvx * nvy; s += blockDim.x) {
            const int j = s / nvx + gc.vlo, i = s % nvx + gc.vlo;
            const double *f = acf_fine + (2 * i) + (long)(2 * j) * fine_rs;
            const long w = fine_rs;
            double v;
            if (gc.cc) v = 0.25 * (f[0] + f[1] + f[w] + f[w + 1]);
            else v = (1. / 16.) * (f[-w - 1] + 2. * f[-w] + f[-w + 1] + 2. * f[-1] + 4. * f[0] + 2. * f[1]
                                   + f[w - 1] + 2. * f[w] + f[w + 1]);
            A.L[0].acf[i + (long)j * gc.nx] = v;
        }
        __syncthreads();
        (void)gfine;
    }
    for (int l = 1; l < A.nl; ++l) c_restrict(A.L[l].g, A.L[l].acf, A.L[l - 1].g, A.L[l - 1].acf, 1);
    for (int l = 0; l < A.nl; ++l) {
        const LevelGeom &g = A.L[l].g;
        const int nvx = g.vhix - g.vlo + 1, nvy = g.vhiy - g.vlo + 1;
        for (int s = threadIdx.x; s < nvx * nvy; s += blockDim.x) {
            const int j = s / nvx + g.vlo, i = s % nvx + g.vlo;
            const long o = i + (long)j * g.nx;
            double c0 = -(A.L[l].acf[o] + 2.0 * (g.facx + g.facy));
            if (g.cc && (i == g.vlo || i == g.vhix)) c0 -= 2.0 * g.facx;
            if (g.cc && (j == g.vlo || j == g.vhiy)) c0 -= 2.0 * g.facy;
            A.L[l].c0i[o] = 1.0 / c0;
        }
    }
}

// crse = R(fine), 1 component, arbitrary fine row stride (coefficient average-down on tile levels)
__global__ void k_restrict_acf(LevelGeom gc, double *crse, double *c0i, const double *fine, long fine_rs)
{
    hpb_pdl_prologue();
    const int nvx = gc.vhix - gc.vlo + 1;
    const long nv = (long)nvx * (gc.vhiy - gc.vlo + 1);
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nv) return;
    const int j = (int)(s / nvx) + gc.vlo, i = (int)(s % nvx) + gc.vlo;
    const double *f = fine + (2 * i) + (long)(2 * j) * fine_rs;
    const long w = fine_rs;
    double v;
    if (gc.cc) v = 0.25 * (f[0] + f[1] + f[w] + f[w + 1]);
    else v = (1. / 16.) * (f[-w - 1] + 2. * f[-w] + f[-w + 1] + 2. * f[-1] + 4. * f[0] + 2. * f[1]
                           + f[w - 1] + 2. * f[w] + f[w + 1]);
    crse[i + (long)j * gc.nx] = v;
    // 1 / c0 of the coarse level (gs1 :265-292), used by every smoother launch of this solve
    double c0 = -(v + 2.0 * (gc.facx + gc.facy));
    if (gc.cc && (i == gc.vlo || i == gc.vhix)) c0 -= 2.0 * gc.facx;
    if (gc.cc && (j == gc.vlo || j == gc.vhiy)) c0 -= 2.0 * gc.facy;
    c0i[i + (long)j * gc.nx] = 1.0 / c0;
}

__global__ void k_copy2(LevelGeom g, V2 dst, V2 src)
{
    hpb_pdl_prologue();
    const int nvx = g.vhix - g.vlo + 1;
    const long nv = (long)nvx * (g.vhiy - g.vlo + 1);
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nv) return;
    const int j = (int)(s / nvx) + g.vlo, i = (int)(s % nvx) + g.vlo;
    dst.at(i, j, 0) = src.at(i, j, 0);
    dst.at(i, j, 1) = src.at(i, j, 1);
}

// convergence bookkeeping on the device (solve_doit :1354-1416)
//   st[0] = res_target, st[1] = max_norm, st[2] = last norm;  ist[0] = done, ist[1] = V-cycles,
//   ist[2] = failed.  norm[] is reset for the next accumulation.
__global__ void k_mg_check(int mode, double *norm, double *st, int *ist, double tol_rel, double tol_abs)
{
    hpb_pdl_prologue();
    if (mode == 0) {
        const double resnorm0 = norm[0], rhsnorm0 = norm[1];
        const double max_norm = rhsnorm0 >= resnorm0 ? rhsnorm0 : resnorm0;
        st[0] = fmax(tol_abs, fmax(tol_rel, 1.e-16) * max_norm);      // :1361
        st[1] = max_norm;
        st[2] = resnorm0;
        ist[0] = (resnorm0 <= st[0]) ? 1 : 0;
        ist[1] = 0;
        ist[2] = 0;
    } else if (!ist[0]) {
        const double norminf = norm[0];
        st[2] = norminf;
        ist[1] += 1;
        if (norminf <= st[0]) ist[0] = 1;
        else if (norminf > 1.e20 * st[1] || norminf != norminf) { ist[0] = 1; ist[2] = 1; }
    }
    norm[0] = 0.;
    norm[1] = 0.;
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

template <int INIT, bool RES, int RH, int NS, int NTHR>
int launch_smooth_t(hpb_ctx *ctx, const LevelGeom &g, V2 in, V2 crse, V2 rhs, const double *acf,
                    long acf_rs, const double *c0i_in, double *c0i_out, V2 out, const LevelGeom &gc,
                    V2 res_c, double *norm, const int *done, int nbx, int nby, int EO)
{
    const size_t smem = 2 * sizeof(double) * (size_t)(CY * RH + 2) * AX;
    static bool attr_set = false;       // per instantiation
    if (!attr_set) {
        HPB_CUDA_CHECK(cudaFuncSetAttribute(k_smooth<INIT, RES, RH, NS, NTHR>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    hpb_launch(k_smooth<INIT, RES, RH, NS, NTHR>, nbx * nby, NTHR, smem, ctx->stream, g, in, crse, rhs,
               acf, acf_rs, c0i_in, c0i_out, out, gc, res_c, norm, nbx, EO, done, ctx->tune_mg_lean);
    hpb_count_launch(ctx);
    return HPB_OK;
}

template <int INIT, bool RES, int RH = 1, int NS = 1>
int launch_smooth(hpb_ctx *ctx, const LevelGeom &g, V2 in, V2 crse, V2 rhs, const double *acf,
                  long acf_rs, const double *c0i_in, double *c0i_out, V2 out, const LevelGeom &gc,
                  V2 res_c, double *norm, const int *done)
{
    // owned region must start on even indices relative to vlo for the fused restriction;
    // cc needs 1 extra ring for the residual (EO 4), nodal full weighting one more (EO 5);
    // every further GSRB^4 group costs 4 more rings
    // (FX, FY are even for any EO, so tile origins stay even)
    const int EO = (RES ? (g.cc ? 4 : 5) : 3) + 4 * (NS - 1);
    const int FX = CX - 2 * EO, FY = CY * RH - 2 * EO;
    const int nvx = g.vhix - g.vlo + 1, nvy = g.vhiy - g.vlo + 1;
    const int nbx = (nvx + FX - 1) / FX, nby = (nvy + FY - 1) / FY;
    const int wide = ctx->tune_mg_wide;
    // levels whose tiles do not even fill the GPU once are latency bound: 1024 threads per tile
    // (option "mg_wide", default on since round 2: 0.393 -> 0.367 ms per slice and +2 % slices/s in the
    // un-profiled slice loop, profiles/r02/r02r_tune.txt; round 1 had measured no gain)
    if (RH == 1 && NS == 1 && wide && nbx * nby <= 148)
        return launch_smooth_t<INIT, RES, 1, 1, 1024>(ctx, g, in, crse, rhs, acf, acf_rs, c0i_in, c0i_out, out,
                                                      gc, res_c, norm, done, nbx, nby, EO);
    return launch_smooth_t<INIT, RES, RH, NS, NT>(ctx, g, in, crse, rhs, acf, acf_rs, c0i_in, c0i_out, out, gc,
                                                  res_c, norm, done, nbx, nby, EO);
}

inline unsigned nb(long n) { return (unsigned)((n + 255) / 256); }

CoarseArgs coarse_args(hpb_ctx *ctx)
{
    CoarseArgs A;
    const int lc = ctx->mg_lc, nl = ctx->mg_nlev;
    A.nl = nl - lc;
    for (int l = lc; l < nl; ++l) {
        CoarseLevel &L = A.L[l - lc];
        L.g = level_geom(ctx, l);
        L.acf = ctx->mg[l].acf; L.c0i = ctx->mg[l].c0i; L.res = ctx->mg[l].res;
        L.cor = ctx->mg[l].cor; L.rescor = ctx->mg[l].rescor;
    }
    const LevelGeom gb = level_geom(ctx, nl - 1);
    int nsw = 16;
    const int mx = gb.nx > gb.ny ? gb.nx : gb.ny;
    if ((mx + 1) / 2 * 2 > nsw) nsw = (mx + 1) / 2 * 2;       // HpMultiGrid.cpp:1587
    A.nsweeps_bottom = nsw;
    return A;
}


// =================================================================================================
// hpmg system type 2: the complex Helmholtz system  lap(A) - (a_r + i a_i) A = rhs  of the laser
// envelope solver (MultiLaser::AdvanceSliceMG, src/laser/MultiLaser.cpp:429-607) as two coupled real
// components: gs2 (HpMultiGrid.cpp:296-334), residual2r / residual2i (:192-208), solve2 (:1192-1296).
// V-cycle, transfer operators and stopping rule are those of type 1 (solve_doit :1307-1427, vcycle
// :1429-1512, bottomsolve :1514-1594).  First implementation: one plain global-memory kernel per
// level operator (a red-black half-sweep, a residual, a restriction, an interpolation), the level
// arrays of the type-1 solver reused as scratch.  Correct against the oracle's MultiGrid2; not yet
// tile-fused like the type-1 path.
// =================================================================================================
struct MG2 {
    double *acf[32] = {};       // per level: 2 components (a_r, a_i) over the level box
    double *sol0 = nullptr, *rhs0 = nullptr;     // level-0 iterate and right-hand side (2 comps, level box)
};

// gs2 (HpMultiGrid.cpp:296-334) at point (i, j)
__device__ __forceinline__ void gs2_update(const LevelGeom &g, const V2 &phi, const V2 &rhs, const V2 &acf, int i, int j)
{
    double lap0, lap1;
    double c0 = -2.0 * (g.facx + g.facy);
    if (g.cc && i == g.vlo) {
        lap0 = g.facx * (4. / 3.) * phi.at(i + 1, j, 0); lap1 = g.facx * (4. / 3.) * phi.at(i + 1, j, 1);
        c0 -= 2.0 * g.facx;
    } else if (g.cc && i == g.vhix) {
        lap0 = g.facx * (4. / 3.) * phi.at(i - 1, j, 0); lap1 = g.facx * (4. / 3.) * phi.at(i - 1, j, 1);
        c0 -= 2.0 * g.facx;
    } else {
        lap0 = g.facx * (phi.at(i - 1, j, 0) + phi.at(i + 1, j, 0));
        lap1 = g.facx * (phi.at(i - 1, j, 1) + phi.at(i + 1, j, 1));
    }
    if (g.cc && j == g.vlo) {
        lap0 += g.facy * (4. / 3.) * phi.at(i, j + 1, 0); lap1 += g.facy * (4. / 3.) * phi.at(i, j + 1, 1);
        c0 -= 2.0 * g.facy;
    } else if (g.cc && j == g.vhiy) {
        lap0 += g.facy * (4. / 3.) * phi.at(i, j - 1, 0); lap1 += g.facy * (4. / 3.) * phi.at(i, j - 1, 1);
        c0 -= 2.0 * g.facy;
    } else {
        lap0 += g.facy * (phi.at(i, j - 1, 0) + phi.at(i, j + 1, 0));
        lap1 += g.facy * (phi.at(i, j - 1, 1) + phi.at(i, j + 1, 1));
    }
    double cr = c0 - acf.at(i, j, 0), ci = -acf.at(i, j, 1);
    const double cmag = 1.0 / (cr * cr + ci * ci);
    cr *= cmag; ci *= cmag;
    const double dr = rhs.at(i, j, 0) - lap0, di = rhs.at(i, j, 1) - lap1;
    phi.at(i, j, 0) = dr * cr + di * ci;
    phi.at(i, j, 1) = di * cr - dr * ci;
}

// one red-black half-sweep of gs2 on the valid points with (i + j + color) even
__global__ void k2_gsrb(LevelGeom g, V2 phi, V2 rhs, V2 acf, int color)
{
    hpb_pdl_prologue();
    const int nvx = g.vhix - g.vlo + 1;
    const long nv = (long)nvx * (g.vhiy - g.vlo + 1);
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nv) return;
    const int j = (int)(s / nvx) + g.vlo, i = (int)(s % nvx) + g.vlo;
    if (((i + j + color) & 1) != 0) return;
    gs2_update(g, phi, rhs, acf, i, j);
}

__device__ __forceinline__ double lap_global(const LevelGeom &g, const V2 &phi, int i, int j, int n)
{
    double lap = -2.0 * (g.facx + g.facy) * phi.at(i, j, n);
    if (g.cc && i == g.vlo) lap += g.facx * ((4. / 3.) * phi.at(i + 1, j, n) - 2.0 * phi.at(i, j, n));
    else if (g.cc && i == g.vhix) lap += g.facx * ((4. / 3.) * phi.at(i - 1, j, n) - 2.0 * phi.at(i, j, n));
    else lap += g.facx * (phi.at(i - 1, j, n) + phi.at(i + 1, j, n));
    if (g.cc && j == g.vlo) lap += g.facy * ((4. / 3.) * phi.at(i, j + 1, n) - 2.0 * phi.at(i, j, n));
    else if (g.cc && j == g.vhiy) lap += g.facy * ((4. / 3.) * phi.at(i, j - 1, n) - 2.0 * phi.at(i, j, n));
    else lap += g.facy * (phi.at(i, j - 1, n) + phi.at(i, j + 1, n));
    return lap;
}

// out = residual2(phi) on the valid points; norm (may be null): [0] = max|out|, [1] = max|rhs|
__global__ void k2_residual(LevelGeom g, V2 out, V2 phi, V2 rhs, V2 acf, double *norm)
{
    hpb_pdl_prologue();
    const int nvx = g.vhix - g.vlo + 1;
    const long nv = (long)nvx * (g.vhiy - g.vlo + 1);
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double nres = 0., nrhs = 0.;
    if (s < nv) {
        const int j = (int)(s / nvx) + g.vlo, i = (int)(s % nvx) + g.vlo;
        const double ar = acf.at(i, j, 0), ai = acf.at(i, j, 1);
        const double pr = phi.at(i, j, 0), pi = phi.at(i, j, 1);
        const double r0 = rhs.at(i, j, 0) + ar * pr - ai * pi - lap_global(g, phi, i, j, 0);
        const double r1 = rhs.at(i, j, 1) + ai * pr + ar * pi - lap_global(g, phi, i, j, 1);
        out.at(i, j, 0) = r0;
        out.at(i, j, 1) = r1;
        nres = fmax(fabs(r0), fabs(r1));
        nrhs = fmax(fabs(rhs.at(i, j, 0)), fabs(rhs.at(i, j, 1)));
    }
    if (!norm) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nres = fmax(nres, __shfl_xor_sync(0xffffffffu, nres, o));
        nrhs = fmax(nrhs, __shfl_xor_sync(0xffffffffu, nrhs, o));
    }
    if ((threadIdx.x & 31) == 0 && (nres > 0. || nrhs > 0.)) {
        atomicMax((unsigned long long *)&norm[0], (unsigned long long)__double_as_longlong(nres));
        atomicMax((unsigned long long *)&norm[1], (unsigned long long)__double_as_longlong(nrhs));
    }
}

// crse = R(fine) for 2 components on the valid coarse points (restrict_cc / restrict_nd :29-52)
__global__ void k2_restrict(LevelGeom gc, V2 crse, V2 fine)
{
    hpb_pdl_prologue();
    const int nvx = gc.vhix - gc.vlo + 1;
    const long nv = (long)nvx * (gc.vhiy - gc.vlo + 1);
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nv) return;
    const int j = (int)(s / nvx) + gc.vlo, i = (int)(s % nvx) + gc.vlo;
#pragma unroll
    for (int n = 0; n < 2; ++n) {
        const double *f = fine.p + n * fine.cs + (2 * i) + (long)(2 * j) * fine.rs;
        const long w = fine.rs;
        double v;
        if (gc.cc) v = 0.25 * (f[0] + f[1] + f[w] + f[w + 1]);
        else v = (1. / 16.) * (f[-w - 1] + 2. * f[-w] + f[-w + 1] + 2. * f[-1] + 4. * f[0] + 2. * f[1]
                               + f[w - 1] + 2. * f[w] + f[w + 1]);
        crse.at(i, j, n) = v;
    }
}

// out = fine + I(crse) on the valid fine points (interpcpy_cc / interpcpy_nd :88-121); out may be fine
__global__ void k2_interp_add(LevelGeom g, V2 out, V2 fine, V2 crse)
{
    hpb_pdl_prologue();
    const int nvx = g.vhix - g.vlo + 1;
    const long nv = (long)nvx * (g.vhiy - g.vlo + 1);
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nv) return;
    const int j = (int)(s / nvx) + g.vlo, i = (int)(s % nvx) + g.vlo;
    out.at(i, j, 0) = fine.at(i, j, 0) + interp_at(crse, i, j, 0, g.cc);
    out.at(i, j, 1) = fine.at(i, j, 1) + interp_at(crse, i, j, 1, g.cc);
}

// level-0 embedding: planar valid-box arrays [2][ny][nx] <-> the level box (offset vlo)
__global__ void k2_embed(LevelGeom g, V2 lvl, const double *valid2, const double *valid_r, double scalar_i, int nxv,
                         int nyv, int mode)
{
    hpb_pdl_prologue();
    const long nv = (long)nxv * nyv;
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nv) return;
    const int jv = (int)(s / nxv), iv = (int)(s % nxv);
    const int i = iv + g.vlo, j = jv + g.vlo;
    if (mode == 0) {            // two planar components in
        lvl.at(i, j, 0) = valid2[s];
        lvl.at(i, j, 1) = valid2[nv + s];
    } else {                    // coefficient: array real part, scalar imaginary part
        lvl.at(i, j, 0) = valid_r[s];
        lvl.at(i, j, 1) = scalar_i;
    }
}
__global__ void k2_extract(LevelGeom g, V2 lvl, double *valid2, int nxv, int nyv)
{
    hpb_pdl_prologue();
    const long nv = (long)nxv * nyv;
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nv) return;
    const int jv = (int)(s / nxv), iv = (int)(s % nxv);
    valid2[s] = lvl.at(iv + g.vlo, jv + g.vlo, 0);
    valid2[nv + s] = lvl.at(iv + g.vlo, jv + g.vlo, 1);
}

// ---- all levels of <= kSmall2 points in ONE 1024-thread CTA (the role k_coarse plays for type 1): down
// through the small levels, the bottom solve, and back up, with block barriers between the operators.
// The arrays stay in global memory (they are L1 / L2 resident: <= 64 x 64 points per level).
constexpr int kSmall2 = 66 * 66;
struct Small2Level { LevelGeom g; V2 acf, res, cor, rescor; };
struct Small2Args { int nl; int nsweeps_bottom; Small2Level L[16]; };

__device__ void s2_gsrb(const Small2Level &L, const V2 &phi, const V2 &rhs, int nsweeps)
{
    const LevelGeom &g = L.g;
    const int nvx = g.vhix - g.vlo + 1, nv = nvx * (g.vhiy - g.vlo + 1);
    for (int ic = 0; ic < nsweeps; ++ic) {
        for (int s = threadIdx.x; s < nv; s += blockDim.x) {
            const int j = s / nvx + g.vlo, i = s % nvx + g.vlo;
            if (((i + j + ic) & 1) == 0) gs2_update(g, phi, rhs, L.acf, i, j);
        }
        __syncthreads();
    }
}
__device__ void s2_zero(const Small2Level &L, const V2 &a)
{
    const long n = 2L * L.g.nx * L.g.ny;
    for (long s = threadIdx.x; s < n; s += blockDim.x) a.p[s] = 0.;
    __syncthreads();
}

__global__ void __launch_bounds__(1024) k2_small(Small2Args A, const int *done)
{
    hpb_pdl_prologue();
    if (done && *(const volatile int *)done) return;
    const int nl = A.nl;
    for (int l = 0; l < nl - 1; ++l) {
        const Small2Level &L = A.L[l];
        const LevelGeom &g = L.g;
        s2_zero(L, L.cor);
        s2_gsrb(L, L.cor, L.res, 4);
        const int nvx = g.vhix - g.vlo + 1, nv = nvx * (g.vhiy - g.vlo + 1);
        for (int s = threadIdx.x; s < nv; s += blockDim.x) {
            const int j = s / nvx + g.vlo, i = s % nvx + g.vlo;
            const double ar = L.acf.at(i, j, 0), ai = L.acf.at(i, j, 1);
            const double pr = L.cor.at(i, j, 0), pi = L.cor.at(i, j, 1);
            L.rescor.at(i, j, 0) = L.res.at(i, j, 0) + ar * pr - ai * pi - lap_global(g, L.cor, i, j, 0);
            L.rescor.at(i, j, 1) = L.res.at(i, j, 1) + ai * pr + ar * pi - lap_global(g, L.cor, i, j, 1);
        }
        __syncthreads();
        const Small2Level &C = A.L[l + 1];
        const LevelGeom &gc = C.g;
        const int cvx = gc.vhix - gc.vlo + 1, cv = cvx * (gc.vhiy - gc.vlo + 1);
        for (int s = threadIdx.x; s < 2 * cv; s += blockDim.x) {
            const int n = s >= cv, q = s - n * cv;
            const int j = q / cvx + gc.vlo, i = q % cvx + gc.vlo;
            const double *f = L.rescor.p + n * L.rescor.cs + (2 * i) + (long)(2 * j) * L.rescor.rs;
            const long w = L.rescor.rs;
            double v;
            if (gc.cc) v = 0.25 * (f[0] + f[1] + f[w] + f[w + 1]);
            else v = (1. / 16.) * (f[-w - 1] + 2. * f[-w] + f[-w + 1] + 2. * f[-1] + 4. * f[0] + 2. * f[1]
                                   + f[w - 1] + 2. * f[w] + f[w + 1]);
            C.res.at(i, j, n) = v;
        }
        __syncthreads();
    }
    {
        const Small2Level &B = A.L[nl - 1];
        s2_zero(B, B.cor);
        s2_gsrb(B, B.cor, B.res, A.nsweeps_bottom);
    }
    for (int l = nl - 2; l >= 0; --l) {
        const Small2Level &L = A.L[l];
        const LevelGeom &g = L.g;
        const V2 crse = A.L[l + 1].cor;
        const int nvx = g.vhix - g.vlo + 1, nv = nvx * (g.vhiy - g.vlo + 1);
        for (int s = threadIdx.x; s < nv; s += blockDim.x) {
            const int j = s / nvx + g.vlo, i = s % nvx + g.vlo;
            L.cor.at(i, j, 0) = L.cor.at(i, j, 0) + interp_at(crse, i, j, 0, g.cc);
            L.cor.at(i, j, 1) = L.cor.at(i, j, 1) + interp_at(crse, i, j, 1, g.cc);
        }
        __syncthreads();
        s2_gsrb(L, L.cor, L.res, 4);
    }
}

// ---- tile smoother for the large levels: phi_out = GSRB^4(init) on a 32 x 32 tile with a 5-cell halo held
// in shared memory (after four half-sweeps the values four cells inside the loaded region are exact, one
// more ring for the residual), the per-cell inverse coefficient (c0 - a_r, -a_i) / |.|^2 of gs2 computed
// once per tile, optional residual2 + max-norms.  One launch replaces four half-sweep launches and the
// residual launch.  INIT 0: phi = 0, 1: phi = phi_in, 2: phi = phi_in + I(crse).
constexpr int kT2 = 32, kH2 = 5, kR2 = kT2 + 2 * kH2;          // tile, halo, region side (42)
constexpr int kT2Threads = 256;
template <int INIT, bool RES>
__global__ void __launch_bounds__(kT2Threads, 2)
k2_smooth4(LevelGeom g, V2 phi_in, V2 crse, V2 rhs, V2 acf, V2 phi_out, V2 res_out, double *norm, int ntx)
{
    extern __shared__ double sm2[];
    double *p0 = sm2, *p1 = sm2 + kR2 * kR2, *r0 = sm2 + 2 * kR2 * kR2, *r1 = sm2 + 3 * kR2 * kR2;
    double *cr = sm2 + 4 * kR2 * kR2, *ci = sm2 + 5 * kR2 * kR2;
    hpb_pdl_prologue();
    const int tx = blockIdx.x % ntx, ty = blockIdx.x / ntx;
    const int ox = g.vlo + tx * kT2 - kH2, oy = g.vlo + ty * kT2 - kH2;       // level index of region cell (0, 0)
    const double fx43 = g.facx * (4. / 3.), fy43 = g.facy * (4. / 3.);
    for (int e = threadIdx.x; e < kR2 * kR2; e += kT2Threads) {
        const int lj = e / kR2, li = e - lj * kR2;
        const int i = ox + li, j = oy + lj;
        double a = 0., b = 0., q0 = 0., q1 = 0., c_r = 0., c_i = 0.;
        if (i >= g.vlo && i <= g.vhix && j >= g.vlo && j <= g.vhiy) {
            if (INIT >= 1) { a = phi_in.at(i, j, 0); b = phi_in.at(i, j, 1); }
            if (INIT == 2) { a = a + interp_at(crse, i, j, 0, g.cc); b = b + interp_at(crse, i, j, 1, g.cc); }
            q0 = rhs.at(i, j, 0); q1 = rhs.at(i, j, 1);
            double c0 = -2.0 * (g.facx + g.facy);
            if (g.cc && (i == g.vlo || i == g.vhix)) c0 -= 2.0 * g.facx;
            if (g.cc && (j == g.vlo || j == g.vhiy)) c0 -= 2.0 * g.facy;
            c_r = c0 - acf.at(i, j, 0); c_i = -acf.at(i, j, 1);
            const double cmag = 1.0 / (c_r * c_r + c_i * c_i);
            c_r *= cmag; c_i *= cmag;
        }
        p0[e] = a; p1[e] = b; r0[e] = q0; r1[e] = q1; cr[e] = c_r; ci[e] = c_i;
    }
    __syncthreads();
    for (int ic = 0; ic < 4; ++ic) {
        // interior of the region only (the outermost ring has no neighbours in shared memory)
        for (int e = threadIdx.x; e < (kR2 - 2) * (kR2 - 2); e += kT2Threads) {
            const int lj = e / (kR2 - 2) + 1, li = e - (lj - 1) * (kR2 - 2) + 1;
            const int i = ox + li, j = oy + lj;
            if (((i + j + ic) & 1) != 0 || i < g.vlo || i > g.vhix || j < g.vlo || j > g.vhiy) continue;
            const int o = lj * kR2 + li;
            const double wx = (g.cc && (i == g.vlo || i == g.vhix)) ? fx43 : g.facx;
            const double wy = (g.cc && (j == g.vlo || j == g.vhiy)) ? fy43 : g.facy;
            double lap0 = wx * (p0[o - 1] + p0[o + 1]);
            double lap1 = wx * (p1[o - 1] + p1[o + 1]);
            lap0 += wy * (p0[o - kR2] + p0[o + kR2]);
            lap1 += wy * (p1[o - kR2] + p1[o + kR2]);
            const double dr = r0[o] - lap0, di = r1[o] - lap1;
            p0[o] = dr * cr[o] + di * ci[o];
            p1[o] = di * cr[o] - dr * ci[o];
        }
        __syncthreads();
    }
    double nres = 0., nrhs = 0.;
    for (int e = threadIdx.x; e < kT2 * kT2; e += kT2Threads) {
        const int lj = e / kT2 + kH2, li = e - (lj - kH2) * kT2 + kH2;
        const int i = ox + li, j = oy + lj;
        if (i > g.vhix || j > g.vhiy) continue;
        const int o = lj * kR2 + li;
        phi_out.at(i, j, 0) = p0[o];
        phi_out.at(i, j, 1) = p1[o];
        if (RES) {
            double lap[2];
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const double *c = (n ? p1 : p0) + o;
                double l = -2.0 * (g.facx + g.facy) * c[0];
                if (g.cc && i == g.vlo) l += g.facx * ((4. / 3.) * c[1] - 2.0 * c[0]);
                else if (g.cc && i == g.vhix) l += g.facx * ((4. / 3.) * c[-1] - 2.0 * c[0]);
                else l += g.facx * (c[-1] + c[1]);
                if (g.cc && j == g.vlo) l += g.facy * ((4. / 3.) * c[kR2] - 2.0 * c[0]);
                else if (g.cc && j == g.vhiy) l += g.facy * ((4. / 3.) * c[-kR2] - 2.0 * c[0]);
                else l += g.facy * (c[-kR2] + c[kR2]);
                lap[n] = l;
            }
            const double ar = acf.at(i, j, 0), ai = acf.at(i, j, 1);
            const double v0 = r0[o] + ar * p0[o] - ai * p1[o] - lap[0];
            const double v1 = r1[o] + ai * p0[o] + ar * p1[o] - lap[1];
            res_out.at(i, j, 0) = v0;
            res_out.at(i, j, 1) = v1;
            nres = fmax(nres, fmax(fabs(v0), fabs(v1)));
            nrhs = fmax(nrhs, fmax(fabs(r0[o]), fabs(r1[o])));
        }
    }
    if (!RES || !norm) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nres = fmax(nres, __shfl_xor_sync(0xffffffffu, nres, o));
        nrhs = fmax(nrhs, __shfl_xor_sync(0xffffffffu, nrhs, o));
    }
    if ((threadIdx.x & 31) == 0 && (nres > 0. || nrhs > 0.)) {
        atomicMax((unsigned long long *)&norm[0], (unsigned long long)__double_as_longlong(nres));
        atomicMax((unsigned long long *)&norm[1], (unsigned long long)__double_as_longlong(nrhs));
    }
}

template <int INIT, bool RES>
int mg2_smooth4(hpb_ctx *ctx, const LevelGeom &g, V2 phi_in, V2 crse, V2 rhs, V2 acf, V2 phi_out, V2 res_out,
                double *norm)
{
    const size_t smem = 6 * sizeof(double) * kR2 * kR2;
    static bool attr_set = false;
    if (!attr_set) {
        HPB_CUDA_CHECK(cudaFuncSetAttribute(k2_smooth4<INIT, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int ntx = (g.vhix - g.vlo + kT2) / kT2, nty = (g.vhiy - g.vlo + kT2) / kT2;
    hpb_launch(k2_smooth4<INIT, RES>, (unsigned)(ntx * nty), kT2Threads, smem, ctx->stream, g, phi_in, crse, rhs, acf,
               phi_out, res_out, norm, ntx);
    hpb_count_launch(ctx);
    return HPB_OK;
}

int mg2_sweeps(hpb_ctx *ctx, const LevelGeom &g, V2 phi, V2 rhs, V2 acf, int nsweeps)
{
    const long nv = (long)(g.vhix - g.vlo + 1) * (g.vhiy - g.vlo + 1);
    for (int ic = 0; ic < nsweeps; ++ic) hpb_launch(k2_gsrb, nb(nv), 256, 0, ctx->stream, g, phi, rhs, acf, ic);
    hpb_count_launch(ctx, nsweeps);
    return HPB_OK;
}

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
    if (nl < 2) {
        hpb_set_error("hpmg: a %d x %d grid cannot be coarsened (HpMultiGrid.cpp:1085-1090)", nx, ny);
        return HPB_ERR_UNSUPPORTED;
    }
    // first level handled by the single-CTA kernel (>= 1: level 0 always uses the tile kernels)
    int lc = 1;
    while (lc < nl - 1 && (ctx->mg[lc].nx > kCoarseMax || ctx->mg[lc].ny > kCoarseMax)) ++lc;
    if (nl - lc > 12) lc = nl - 12;
    // (grids that stop coarsening early leave a large 'coarse' level: k_coarse loops over it)
    ctx->mg_lc = lc;
    {
        size_t bytes = 0;
        for (int l = lc; l < nl; ++l) bytes += 8 * sizeof(double) * (size_t)ctx->mg[l].nx * ctx->mg[l].ny;
        ctx->mg_coarse_smem = 0;
        bool fits = true;       // one thread per valid point of a level
        for (int l = lc; l < nl; ++l) {
            const int vx = ctx->mg_cc ? ctx->mg[l].nx : ctx->mg[l].nx - 2;
            const int vy = ctx->mg_cc ? ctx->mg[l].ny : ctx->mg[l].ny - 2;
            if (vx > 32 || vy > 32) fits = false;
        }
        if (fits && bytes <= 200 * 1024) {
            HPB_CUDA_CHECK(cudaFuncSetAttribute(k_coarse, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                200 * 1024));
            ctx->mg_coarse_smem = (int)bytes;
        }
    }
    for (int l = 0; l < nl; ++l) {
        const size_t n = (size_t)ctx->mg[l].nx * ctx->mg[l].ny;
        double **arrs[5] = {&ctx->mg[l].acf, &ctx->mg[l].c0i, &ctx->mg[l].res, &ctx->mg[l].cor,
                            &ctx->mg[l].rescor};
        const size_t mult[5] = {1, 1, 2, 2, 2};
        for (int k = 0; k < 5; ++k) {
            if (l == 0 && (k == 0 || k == 2)) { *arrs[k] = nullptr; continue; }  // chi / rhs in place
            HPB_CUDA_CHECK(cudaMalloc(arrs[k], mult[k] * n * sizeof(double)));
            HPB_CUDA_CHECK(cudaMemset(*arrs[k], 0, mult[k] * n * sizeof(double)));
        }
    }
    HPB_CUDA_CHECK(cudaMalloc(&ctx->d_mg_norm, 2 * sizeof(double)));
    HPB_CUDA_CHECK(cudaMemset(ctx->d_mg_norm, 0, 2 * sizeof(double)));
    HPB_CUDA_CHECK(cudaMalloc(&ctx->d_mg_state, 4 * sizeof(double)));
    HPB_CUDA_CHECK(cudaMalloc(&ctx->d_mg_istate, 4 * sizeof(int)));
    HPB_CUDA_CHECK(cudaMalloc(&ctx->d_mg_bar, 2 * sizeof(unsigned)));
    HPB_CUDA_CHECK(cudaMemset(ctx->d_mg_bar, 0, 2 * sizeof(unsigned)));
    HPB_CUDA_CHECK(cudaMemset(ctx->d_mg_istate, 0, 4 * sizeof(int)));
    HPB_CUDA_CHECK(cudaMallocHost(&ctx->h_mg_norm, 4 * sizeof(double)));
    HPB_CUDA_CHECK(cudaMallocHost(&ctx->h_mg_istate, 4 * sizeof(int)));
    ctx->mg_last_iters = 2;
    return HPB_OK;
}

void hpb_mg2_free(hpb_ctx *ctx);

void hpb_mg_free(hpb_ctx *ctx)
{
    hpb_mg2_free(ctx);
    for (int l = 0; l < ctx->mg_nlev; ++l) {
        cudaFree(ctx->mg[l].acf); cudaFree(ctx->mg[l].c0i); cudaFree(ctx->mg[l].res);
        cudaFree(ctx->mg[l].cor); cudaFree(ctx->mg[l].rescor);
    }
    cudaFree(ctx->d_mg_norm); cudaFree(ctx->d_mg_state); cudaFree(ctx->d_mg_istate); cudaFree(ctx->d_mg_bar);
    cudaFreeHost(ctx->h_mg_norm); cudaFreeHost(ctx->h_mg_istate);
}

// one V-cycle (:1429-1512).  On entry: cur = cor0 (level-0 iterate after 4 sweeps), res[1] holds
// the restricted residual of cor0.  On exit the same invariants hold for the new cor0, and
// d_mg_norm[0] holds max|rescor0|.  tmp is the second level-0 buffer.
// fused = true: the two level-0 smoothers are ONE launch (8 half-sweeps, 64 x 64 tiles) that
// reads cur and writes the new cor0 into tmp: the caller ping-pongs the two buffers.
// top_out: where the unfused path's last level-0 smoother writes the new cor0 (cur itself, or -- buffer
// rotation, option "mg_rotate" -- the caller's sol planes, so that no final copy is needed)
static int mg_vcycle(hpb_ctx *ctx, V2 cur, V2 tmp, V2 rhs0, const double *chi, long chi_rs,
                     double tol_rel, double tol_abs, bool fused, V2 top_out)
{
    const int lc = ctx->mg_lc;
    const V2 none{};
    const LevelGeom gnone{};
    const int *done = ctx->d_mg_istate;
    // the levels [lm, lc) whose tilings have at most one tile per SM + the single-CTA levels: one persistent
    // launch (k_mid) instead of 2 (lc - lm) + 1
    int lm = lc;
    auto tiles = [&](int l, int EO, int &nbx) {
        const LevelGeom g = level_geom(ctx, l);
        const int FX = CX - 2 * EO, FY = CY - 2 * EO;
        nbx = (g.vhix - g.vlo + 1 + FX - 1) / FX;
        return nbx * ((g.vhiy - g.vlo + 1 + FY - 1) / FY);
    };
    const int EOd = ctx->mg_cc ? 4 : 5, EOu = 3;
    if (ctx->tune_mg_persist && ctx->mg_coarse_smem > 0 && ctx->d_mg_bar) {
        int nbx;
        while (lm > 1 && lc - (lm - 1) <= 6 && tiles(lm - 1, EOd, nbx) <= 148 && tiles(lm - 1, EOu, nbx) <= 148) --lm;
    }
    for (int l = 1; l < lm; ++l) {                 // down, tile levels
        const LevelGeom g = level_geom(ctx, l), gc = level_geom(ctx, l + 1);
        launch_smooth<0, true>(ctx, g, none, none, lvl_view(ctx, l, ctx->mg[l].res), ctx->mg[l].acf,
                               g.nx, ctx->mg[l].c0i, nullptr, lvl_view(ctx, l, ctx->mg[l].cor), gc,
                               lvl_view(ctx, l + 1, ctx->mg[l + 1].res), nullptr, done);
    }
    if (lm < lc) {
        MidArgs A;
        A.nlev = lc - lm;
        int grid = 1;
        for (int l = lm; l < lc; ++l) {
            MidLevel &L = A.L[l - lm];
            L.g = level_geom(ctx, l); L.gc = level_geom(ctx, l + 1);
            L.res = lvl_view(ctx, l, ctx->mg[l].res); L.cor = lvl_view(ctx, l, ctx->mg[l].cor);
            L.rescor = lvl_view(ctx, l, ctx->mg[l].rescor);
            L.res_c = lvl_view(ctx, l + 1, ctx->mg[l + 1].res);
            L.up_crse = lvl_view(ctx, l + 1, l + 1 == lc ? ctx->mg[lc].cor : ctx->mg[l + 1].rescor);
            L.acf = ctx->mg[l].acf; L.c0i = ctx->mg[l].c0i;
            L.EO_d = EOd; L.nt_d = tiles(l, EOd, L.nbx_d);
            L.EO_u = EOu; L.nt_u = tiles(l, EOu, L.nbx_u);
            grid = std::max(grid, std::max(L.nt_d, L.nt_u));
        }
        A.coarse = coarse_args(ctx);
        const size_t smem = std::max((size_t)ctx->mg_coarse_smem, 2 * sizeof(double) * (size_t)(CY + 2) * AX);
        static bool attr_set = false;
        if (!attr_set) {
            HPB_CUDA_CHECK(cudaFuncSetAttribute(k_mid, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set = true;
        }
        // cooperative: the launch fails instead of dead-locking if the CTAs cannot all be resident
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kCoarseThreads); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
        cudaLaunchAttribute at[1];
        if (ctx->tune_mg_persist == 2) {
            // plain PDL launch: at most one CTA per SM and fewer CTAs than SMs, nothing else on the device
            // waits for this kernel, so all CTAs become resident as their predecessors drain
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = hpb_pdl_enabled() ? 1 : 0;
        } else {
            at[0].id = cudaLaunchAttributeCooperative;
            at[0].val.cooperative = 1;
        }
        cfg.attrs = at; cfg.numAttrs = 1;
        HPB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_mid, A, done, (GridBar *)ctx->d_mg_bar, ctx->tune_mg_lean));
        hpb_count_launch(ctx);
    }
    if (lm == lc) {
        if (ctx->mg_coarse_smem > 0)
            hpb_launch(k_coarse, 1, kCoarseThreads, ctx->mg_coarse_smem, ctx->stream, coarse_args(ctx), done);
        else
            hpb_launch(k_coarse_global, 1, kCoarseThreads, 0, ctx->stream, coarse_args(ctx), done);
        hpb_count_launch(ctx);
    }
    // up, tile levels: cor[l] <- GSRB^4(cor[l] + I(cor[l+1])); double-buffered through rescor[l]
    double *up_prev = lm == lc ? ctx->mg[lc].cor : ctx->mg[lm].rescor;
    for (int l = lm - 1; l >= 1; --l) {
        const LevelGeom g = level_geom(ctx, l);
        launch_smooth<2, false>(ctx, g, lvl_view(ctx, l, ctx->mg[l].cor), lvl_view(ctx, l + 1, up_prev),
                                lvl_view(ctx, l, ctx->mg[l].res), ctx->mg[l].acf, g.nx,
                                ctx->mg[l].c0i, nullptr, lvl_view(ctx, l, ctx->mg[l].rescor), gnone, none,
                                nullptr, done);
        up_prev = ctx->mg[l].rescor;
    }
    const LevelGeom g0 = level_geom(ctx, 0), g1 = level_geom(ctx, 1);
    if (fused) {
        // cor0' = GSRB^4(GSRB^4(cor0 + I(cor[1]))), rescor0 = rhs - L cor0' -> res[1], norm
        int rc = launch_smooth<2, true, 2, 2>(ctx, g0, cur, lvl_view(ctx, 1, up_prev), rhs0, chi, chi_rs,
                                              ctx->mg[0].c0i, nullptr, tmp, g1,
                                              lvl_view(ctx, 1, ctx->mg[1].res),
                                              ctx->d_mg_norm, done);
        if (rc) return rc;
    } else {
        // sol = GSRB^4(cor0 + I(cor[1]))  ->  tmp
        launch_smooth<2, false>(ctx, g0, cur, lvl_view(ctx, 1, up_prev), rhs0, chi, chi_rs,
                                ctx->mg[0].c0i, nullptr, tmp, gnone, none, nullptr, done);
        // cor0 = GSRB^4(sol), rescor0 = rhs - L cor0 -> res[1], norm   (:1501-1503)
        launch_smooth<1, true>(ctx, g0, tmp, none, rhs0, chi, chi_rs, ctx->mg[0].c0i, nullptr, top_out, g1,
                               lvl_view(ctx, 1, ctx->mg[1].res), ctx->d_mg_norm, done);
    }
    hpb_launch(k_mg_check, 1, 1, 0, ctx->stream, 1, ctx->d_mg_norm, ctx->d_mg_state, ctx->d_mg_istate, tol_rel, tol_abs);
    hpb_count_launch(ctx);
    return HPB_OK;
}

// average_down_acoef (:1640-1700): the coefficient on every coarser level (+ 1 / c0 of the level);
// level 0 uses chi in place (solve1 :1177-1187)
static void mg_average_down_acf(hpb_ctx *ctx, const double *chi, long chi_rs)
{
    const int lc = ctx->mg_lc;
    const double *fine = chi;
    long fine_rs = chi_rs;
    for (int l = 1; l < lc; ++l) {
        const LevelGeom gc = level_geom(ctx, l);
        const long nv = (long)(gc.vhix - gc.vlo + 1) * (gc.vhiy - gc.vlo + 1);
        hpb_launch(k_restrict_acf, nb(nv), 256, 0, ctx->stream, gc, ctx->mg[l].acf, ctx->mg[l].c0i, fine, fine_rs);
        hpb_count_launch(ctx);
        fine = ctx->mg[l].acf;
        fine_rs = gc.nx;
    }
    hpb_launch(k_coarse_setup, 1, kCoarseThreads, 0, ctx->stream, coarse_args(ctx), level_geom(ctx, lc - 1),
                                                          fine, fine_rs);
    hpb_count_launch(ctx);
}

// The coefficient hierarchy depends on chi only, which is final long before the Bx / By sources are: a
// driver may run this chain of five small launches on another stream beside the Poisson solve and the
// explicit deposition (the caller orders the streams); the next hpb_mg_solve1 with the same coefficient
// plane then starts at its first smoother.  The reference computes it inside solve1 (:1177-1187).
extern "C" int hpb_mg_prepare_acf(hpb_ctx *ctx, hpb_slice sl, int c_acf)
{
    if (!ctx || c_acf < 0 || c_acf >= sl.ncomp) return HPB_ERR_ARG;
    SliceView v = make_view(sl);
    const int sh = ctx->mg_cc ? 0 : -1;
    const double *chi = v.comp(c_acf) + v.idx(sh, sh);
    mg_average_down_acf(ctx, chi, v.jstride);
    ctx->mg_acf_ready = chi;
    HPB_CUDA_CHECK(cudaGetLastError());
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
    const long chi_rs = v.jstride;
    const int nl = ctx->mg_nlev, lc = ctx->mg_lc;
    const LevelGeom g0 = level_geom(ctx, 0), g1 = level_geom(ctx, 1);
    const long nv0 = (long)(g0.vhix - g0.vlo + 1) * (g0.vhiy - g0.vlo + 1);
    const V2 none{};
    (void)nl;

    // average_down_acoef (:1640-1700) unless hpb_mg_prepare_acf did it already for this coefficient
    if (ctx->mg_acf_ready != chi) mg_average_down_acf(ctx, chi, chi_rs);
    ctx->mg_acf_ready = nullptr;

    // cor0 = GSRB^4(sol), rescor0 = rhs - L(cor0)   (:1326-1327), fused with its restriction
    V2 cur = lvl_view(ctx, 0, ctx->mg[0].cor);
    launch_smooth<1, true>(ctx, g0, sol, none, rhs, chi, chi_rs, nullptr, ctx->mg[0].c0i, cur, g1,
                           lvl_view(ctx, 1, ctx->mg[1].res), ctx->d_mg_norm, nullptr);
    hpb_launch(k_mg_check, 1, 1, 0, ctx->stream, 0, ctx->d_mg_norm, ctx->d_mg_state, ctx->d_mg_istate, tol_rel, tol_abs);
    hpb_count_launch(ctx);
    // Speculative V-cycles: as many as the previous solve needed, enqueued without a host round
    // trip; each kernel is a no-op once the device-side test (:1391) has passed.  One
    // synchronisation then tells us whether more are needed.
    const bool fused = ctx->tune_mg_fuse != 0;
    // fused: V-cycle k (1-based) reads buf[(k-1) % 2] and writes buf[k % 2], buf = {cor0 buffer, sol};
    // the cycles that actually run are a prefix of the enqueued ones, so after `iters` executed
    // cycles the iterate lives in buf[iters % 2].  Unfused: sol is only the scratch iterate.
    const V2 buf[2] = {cur, sol};
    // unfused + "mg_rotate": cycle 1 reads cor0 from its level buffer, every later cycle from the caller's sol
    // planes, into which each cycle's last smoother writes; the up-stroke's output goes through the spare
    // level-0 buffer.  The executed cycles are a prefix of the enqueued ones, so after >= 1 executed cycle
    // the result already is in sol and the final copy disappears (it remains for 0 cycles).
    const bool rotate = !fused && ctx->tune_mg_rotate != 0;
    const V2 spare = lvl_view(ctx, 0, ctx->mg[0].rescor);
    int enq = 0;
    auto enqueue_cycle = [&]() -> int {
        int rc;
        if (fused) rc = mg_vcycle(ctx, buf[enq & 1], buf[(enq + 1) & 1], rhs, chi, chi_rs, tol_rel, tol_abs, true, none);
        else if (rotate) rc = mg_vcycle(ctx, enq == 0 ? cur : sol, spare, rhs, chi, chi_rs, tol_rel, tol_abs, false, sol);
        else rc = mg_vcycle(ctx, cur, sol, rhs, chi, chi_rs, tol_rel, tol_abs, false, cur);
        ++enq;
        return rc;
    };
    const int n_pred = ctx->mg_last_iters < 1 ? 1 : (ctx->mg_last_iters > max_iters ? max_iters : ctx->mg_last_iters);
    while (enq < n_pred) {
        int rc = enqueue_cycle();
        if (rc) return rc;
    }
    int iters = 0;
    while (true) {
        HPB_CUDA_CHECK(cudaMemcpyAsync(ctx->h_mg_istate, ctx->d_mg_istate, 4 * sizeof(int),
                                       cudaMemcpyDeviceToHost, ctx->stream));
        HPB_CUDA_CHECK(cudaMemcpyAsync(ctx->h_mg_norm, ctx->d_mg_state, 3 * sizeof(double),
                                       cudaMemcpyDeviceToHost, ctx->stream));
        HPB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        iters = ctx->h_mg_istate[1];
        if (ctx->h_mg_istate[2]) {
            hpb_set_error("hpmg failing so lets stop here (resid/max_norm = %g)",
                          ctx->h_mg_norm[2] / ctx->h_mg_norm[1]);
            return HPB_ERR_MG_DIVERGED;
        }
        if (ctx->h_mg_istate[0]) break;
        if (enq >= max_iters) {
            hpb_set_error("hpmg failed to converge after %d iterations", max_iters);
            return HPB_ERR_MG_DIVERGED;
        }
        int rc = enqueue_cycle();
        if (rc) return rc;
    }
    ctx->mg_last_iters = iters;
    // sol <- cor0 on the valid box (:1419-1426) unless the last executed cycle wrote it there
    if (rotate ? iters == 0 : (!fused || (iters & 1) == 0)) {
        hpb_launch(k_copy2, nb(nv0), 256, 0, ctx->stream, g0, sol, cur);
        hpb_count_launch(ctx);
    }
    HPB_CUDA_CHECK(cudaGetLastError());
    if (h_iters) *h_iters = iters;
    return HPB_OK;
}

void hpb_mg2_free(hpb_ctx *ctx)
{
    MG2 *m = (MG2 *)ctx->mg2;
    if (!m) return;
    for (int l = 0; l < 32; ++l) cudaFree(m->acf[l]);
    cudaFree(m->sol0); cudaFree(m->rhs0);
    delete m;
    ctx->mg2 = nullptr;
}

// hpmg::MultiGrid::solve2 with an array real coefficient and a scalar imaginary one (the laser case).
// d_sol2 / d_rhs2: planar [2][ny][nx] over the valid box (real, imaginary), d_sol2 holds the initial
// guess and receives the solution; d_acf_r: [ny][nx].
extern "C" int hpb_mg_solve2(hpb_ctx *ctx, double *d_sol2, const double *d_rhs2, const double *d_acf_r,
                             double acf_i, double tol_rel, double tol_abs, int max_iters, int *h_iters)
{
    if (!ctx || !d_sol2 || !d_rhs2 || !d_acf_r) return HPB_ERR_ARG;
    const int nl = ctx->mg_nlev;
    MG2 *m = (MG2 *)ctx->mg2;
    if (!m) {
        m = new MG2();
        ctx->mg2 = m;
        for (int l = 0; l < nl; ++l) {
            const size_t n = 2 * (size_t)ctx->mg[l].nx * ctx->mg[l].ny * sizeof(double);
            HPB_CUDA_CHECK(cudaMalloc(&m->acf[l], n));
            HPB_CUDA_CHECK(cudaMemset(m->acf[l], 0, n));
        }
        const size_t n0 = 2 * (size_t)ctx->mg[0].nx * ctx->mg[0].ny * sizeof(double);
        HPB_CUDA_CHECK(cudaMalloc(&m->sol0, n0));
        HPB_CUDA_CHECK(cudaMalloc(&m->rhs0, n0));
        HPB_CUDA_CHECK(cudaMemset(m->sol0, 0, n0));
        HPB_CUDA_CHECK(cudaMemset(m->rhs0, 0, n0));
    }
    const int nxv = ctx->g.nx, nyv = ctx->g.ny;
    const long nvv = (long)nxv * nyv;
    auto G = [&](int l) { return level_geom(ctx, l); };
    auto NV = [&](int l) { const LevelGeom g = G(l); return (long)(g.vhix - g.vlo + 1) * (g.vhiy - g.vlo + 1); };
    auto ACF = [&](int l) { return lvl_view(ctx, l, m->acf[l]); };
    auto bytes2 = [&](int l) { return 2 * (size_t)ctx->mg[l].nx * ctx->mg[l].ny * sizeof(double); };
    const V2 sol0 = lvl_view(ctx, 0, m->sol0), rhs0 = lvl_view(ctx, 0, m->rhs0);
    const V2 cor0 = lvl_view(ctx, 0, ctx->mg[0].cor), rescor0 = lvl_view(ctx, 0, ctx->mg[0].rescor);
    // (the type-1 solver keeps 1 / c0 planes and its iterate in the same arrays: nothing of it
    // survives a type-2 solve, and it rebuilds them on every solve)
    hpb_launch(k2_embed, nb(nvv), 256, 0, ctx->stream, G(0), ACF(0), (const double *)nullptr, d_acf_r, acf_i, nxv, nyv, 1);
    hpb_launch(k2_embed, nb(nvv), 256, 0, ctx->stream, G(0), sol0, (const double *)d_sol2, (const double *)nullptr, 0., nxv, nyv, 0);
    hpb_launch(k2_embed, nb(nvv), 256, 0, ctx->stream, G(0), rhs0, d_rhs2, (const double *)nullptr, 0., nxv, nyv, 0);
    for (int l = 1; l < nl; ++l) hpb_launch(k2_restrict, nb(NV(l)), 256, 0, ctx->stream, G(l), ACF(l), ACF(l - 1));
    hpb_count_launch(ctx, 3 + nl - 1);
    HPB_CUDA_CHECK(cudaMemsetAsync(ctx->d_mg_norm, 0, 2 * sizeof(double), ctx->stream));
    const V2 none{};
    // cor0 = GSRB^4(sol), rescor0 = residual(cor0)   (:1326-1327)
    { int rcs = mg2_smooth4<1, true>(ctx, G(0), sol0, none, rhs0, ACF(0), cor0, rescor0, ctx->d_mg_norm); if (rcs) return rcs; }
    hpb_launch(k_mg_check, 1, 1, 0, ctx->stream, 0, ctx->d_mg_norm, ctx->d_mg_state, ctx->d_mg_istate, tol_rel, tol_abs);
    hpb_count_launch(ctx);
    // ls: first level handled by the single-CTA kernel (>= 1; all levels from there down are small)
    int ls = 1;
    while (ls < nl - 1 && (long)ctx->mg[ls].nx * ctx->mg[ls].ny > kSmall2) ++ls;
    if (nl - ls > 16) ls = nl - 16;
    int iters = 0;
    while (true) {
        HPB_CUDA_CHECK(cudaMemcpyAsync(ctx->h_mg_istate, ctx->d_mg_istate, 4 * sizeof(int), cudaMemcpyDeviceToHost,
                                       ctx->stream));
        HPB_CUDA_CHECK(cudaMemcpyAsync(ctx->h_mg_norm, ctx->d_mg_state, 3 * sizeof(double), cudaMemcpyDeviceToHost,
                                       ctx->stream));
        HPB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        iters = ctx->h_mg_istate[1];
        if (ctx->h_mg_istate[2]) {
            hpb_set_error("hpmg (type 2) failing so lets stop here (resid/max_norm = %g)",
                          ctx->h_mg_norm[2] / ctx->h_mg_norm[1]);
            return HPB_ERR_MG_DIVERGED;
        }
        if (ctx->h_mg_istate[0]) break;
        if (iters >= max_iters) { hpb_set_error("hpmg (type 2) failed to converge after %d iterations", max_iters); return HPB_ERR_MG_DIVERGED; }
        // ---- one V-cycle (:1429-1512): tile smoothers on the large levels, one CTA for the small ones ----
        for (int l = 0; l < ls; ++l) {
            const V2 rescor = lvl_view(ctx, l, ctx->mg[l].rescor);
            if (l > 0) {
                int rcs = mg2_smooth4<0, true>(ctx, G(l), none, none, lvl_view(ctx, l, ctx->mg[l].res), ACF(l),
                                               lvl_view(ctx, l, ctx->mg[l].cor), rescor, nullptr);
                if (rcs) return rcs;
            }
            hpb_launch(k2_restrict, nb(NV(l + 1)), 256, 0, ctx->stream, G(l + 1), lvl_view(ctx, l + 1, ctx->mg[l + 1].res), rescor);
            hpb_count_launch(ctx);
        }
        {
            Small2Args A;
            A.nl = nl - ls;
            for (int l = ls; l < nl; ++l) {
                Small2Level &L = A.L[l - ls];
                L.g = G(l); L.acf = ACF(l);
                L.res = lvl_view(ctx, l, ctx->mg[l].res); L.cor = lvl_view(ctx, l, ctx->mg[l].cor);
                L.rescor = lvl_view(ctx, l, ctx->mg[l].rescor);
            }
            const LevelGeom gb = G(nl - 1);
            int nsw = 16;
            const int mx = gb.nx > gb.ny ? gb.nx : gb.ny;
            if ((mx + 1) / 2 * 2 > nsw) nsw = (mx + 1) / 2 * 2;       // :1587
            A.nsweeps_bottom = nsw;
            hpb_launch(k2_small, 1, 1024, 0, ctx->stream, A, (const int *)nullptr);
            hpb_count_launch(ctx);
        }
        // up: cor[l] <- GSRB^4(cor[l] + I(cor[l+1])), double-buffered through rescor[l] (a tile must not
        // overwrite the halo its neighbours still read)
        double *up_prev = ctx->mg[ls].cor;
        for (int l = ls - 1; l >= 0; --l) {
            const V2 crse = lvl_view(ctx, l + 1, up_prev);
            int rcs;
            if (l == 0) {
                rcs = mg2_smooth4<2, false>(ctx, G(0), cor0, crse, rhs0, ACF(0), sol0, none, nullptr);      // sol
            } else {
                rcs = mg2_smooth4<2, false>(ctx, G(l), lvl_view(ctx, l, ctx->mg[l].cor), crse,
                                            lvl_view(ctx, l, ctx->mg[l].res), ACF(l),
                                            lvl_view(ctx, l, ctx->mg[l].rescor), none, nullptr);
                up_prev = ctx->mg[l].rescor;
            }
            if (rcs) return rcs;
        }
        // cor0 = GSRB^4(sol), rescor0 = residual(cor0)   (:1501-1503)
        { int rcs = mg2_smooth4<1, true>(ctx, G(0), sol0, none, rhs0, ACF(0), cor0, rescor0, ctx->d_mg_norm); if (rcs) return rcs; }
        hpb_launch(k_mg_check, 1, 1, 0, ctx->stream, 1, ctx->d_mg_norm, ctx->d_mg_state, ctx->d_mg_istate, tol_rel, tol_abs);
        hpb_count_launch(ctx);
    }
    hpb_launch(k2_extract, nb(nvv), 256, 0, ctx->stream, G(0), cor0, d_sol2, nxv, nyv);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    if (h_iters) *h_iters = iters;
    return HPB_OK;
}
