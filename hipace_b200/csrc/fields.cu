// Slice housekeeping and stencil kernels (all HBM/L2-bandwidth bound, one thread per cell,
// x fastest => coalesced).  Restates src/fields/Fields.cpp:535-615 (InitializeSlices,
// ShiftSlices, AddRhoIons), :880-956 (RHS assembly, ExmBy/EypBx) and src/Hipace.cpp:744-790.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct CompList { int n; int c[12]; };
struct CompPairs { int n; int dst[8]; int src[8]; };

// grown box, flat
__global__ void k_zero(SliceView a, CompList cl, long ntot)
{
    hpb_pdl_prologue();
    const long o = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= ntot) return;
    for (int n = 0; n < cl.n; ++n) a.comp(cl.c[n])[o] = 0.0;
}

__global__ void k_copy(SliceView a, CompPairs cp, long ntot)
{
    hpb_pdl_prologue();
    const long o = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= ntot) return;
    double v[8];
    for (int n = 0; n < cp.n; ++n) v[n] = a.comp(cp.src[n])[o];
    for (int n = 0; n < cp.n; ++n) a.comp(cp.dst[n])[o] = v[n];
}

__global__ void k_add(SliceView a, CompPairs cp, long ntot)
{
    hpb_pdl_prologue();
    const long o = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= ntot) return;
    for (int n = 0; n < cp.n; ++n) a.comp(cp.dst[n])[o] += a.comp(cp.src[n])[o];
}

// ExmBy = -d/dx Psi, EypBx = -d/dy Psi on the box grown by g-1 (Fields.cpp:931-956).
// G = guard cells of the slice: (depos_order_xy + 1) / 2 + 1 = 1, 2 (default order) or 3
template <int G>
__global__ void k_exmby_eypbx(SliceView a, int c_psi, int c_exmby, int c_eypbx, int nx, int ny,
                              double dx_inv_half, double dy_inv_half)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x - (G - 1);
    const int j = (int)blockIdx.y - (G - 1);
    if (i >= nx + (G - 1)) return;
    const long o = a.idx(i, j);
    const double *psi = a.comp(c_psi);
    const long js = a.jstride;
    a.comp(c_exmby)[o] = -(psi[o + 1] - psi[o - 1]) * dx_inv_half;
    a.comp(c_eypbx)[o] = -(psi[o + js] - psi[o - js]) * dy_inv_half;
}

// Hipace::InitializeSxSyWithBeam (Hipace.cpp:775-788) on the valid box; the guard cells are set to
// zero (the state Fields::InitializeSlices leaves them in), so that Sx / Sy need no separate
// zero-fill pass before this kernel.
template <int G>
__global__ void k_sxsy_from_beam(SliceView a, int c_sx, int c_sy, int c_next_jxb, int c_next_jyb,
                                 int c_jzb, int c_prev_jxb, int c_prev_jyb, int nx, int ny,
                                 double mu0, double dx, double dy, double dz)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x - G;
    const int j = (int)blockIdx.y - G;
    if (i >= nx + G) return;
    const long o = a.idx(i, j);
    if (i < 0 || i >= nx || j < 0 || j >= ny) {
        a.comp(c_sy)[o] = 0.0;
        a.comp(c_sx)[o] = 0.0;
        return;
    }
    const long js = a.jstride;
    const double *jzb = a.comp(c_jzb);
    const double dx_jzb = (jzb[o + 1] - jzb[o - 1]) / (2.0 * dx);
    const double dy_jzb = (jzb[o + js] - jzb[o - js]) / (2.0 * dy);
    const double dz_jxb = (a.comp(c_prev_jxb)[o] - a.comp(c_next_jxb)[o]) / (2.0 * dz);
    const double dz_jyb = (a.comp(c_prev_jyb)[o] - a.comp(c_next_jyb)[o]) / (2.0 * dz);
    a.comp(c_sy)[o] = mu0 * (-dy_jzb + dz_jyb);
    a.comp(c_sx)[o] = -mu0 * (-dx_jzb + dz_jxb);
}

// GridCurrent::DepositCurrentSlice (utils/GridCurrent.cpp:25-70): an analytic gaussian current
// density added to jz_beam on the valid box; cell centres plo + (i + 1/2) dx
__global__ void k_grid_current(SliceView a, int c_jz, int nx, int ny, double plo_x, double plo_y,
                               double dx, double dy, double mx, double my, double sx, double sy,
                               double amp)
{
    hpb_pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= nx || j >= ny) return;
    const double x = plo_x + (i + 0.5) * dx, y = plo_y + (j + 0.5) * dy;
    const double ddx = (x - mx) / sx, ddy = (y - my) / sy;
    a.comp(c_jz)[a.idx(i, j)] += amp * exp(-0.5 * (ddx * ddx + ddy * ddy));
}

// ShiftSlices + InitializeSlices of the next slice in one pass over the grown box
struct ShiftInit {
    int n_zero; int zero[8];
    int n_copy; int dst[4]; int src[4];
};
__global__ void k_shift_init(SliceView a, ShiftInit si, long ntot)
{
    hpb_pdl_prologue();
    const long o = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= ntot) return;
    double v[4];
    for (int n = 0; n < si.n_copy; ++n) v[n] = a.comp(si.src[n])[o];
    for (int n = 0; n < si.n_zero; ++n) a.comp(si.zero[n])[o] = 0.0;
    for (int n = 0; n < si.n_copy; ++n) a.comp(si.dst[n])[o] = v[n];
}

// sum|Q| over the valid box; fixed-order block tree + one atomic per block
__global__ void k_abs_sum(SliceView a, int c, int nx, int ny, double *out)
{
    hpb_pdl_prologue();
    __shared__ double sm[kThreads];
    double acc = 0.0;
    const double *p = a.comp(c);
    for (long s = (long)blockIdx.x * blockDim.x + threadIdx.x; s < (long)nx * ny;
         s += (long)gridDim.x * blockDim.x) {
        const int j = (int)(s / nx), i = (int)(s - (long)j * nx);
        acc += fabs(p[a.idx(i, j)]);
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int st = kThreads / 2; st > 0; st >>= 1) {
        if (threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(out, sm[0]);
}

// sum|Q| of up to 24 components in ONE launch (blockIdx.y = entry): the driver's per-slice checksums were
// one launch per component (16 launches per slice, ~60 us of launch chain for 25 us of memory traffic)
struct AbsSumList { int c[24]; int slot[24]; };
__global__ void k_abs_sum_multi(SliceView a, AbsSumList l, int nx, int ny, double *out)
{
    hpb_pdl_prologue();
    __shared__ double sm[kThreads];
    double acc = 0.0;
    const double *p = a.comp(l.c[blockIdx.y]);
    for (int j = blockIdx.x; j < ny; j += gridDim.x) {
        const double *row = p + a.idx(0, j);
        for (int i = threadIdx.x; i < nx; i += blockDim.x) acc += fabs(row[i]);
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int st = kThreads / 2; st > 0; st >>= 1) {
        if (threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(out + l.slot[blockIdx.y], sm[0]);
}

// the same for an xz diagnostic (diagnostics/Diagnostic.cpp:393-407, Fields::Copy with order-1
// interpolation to y = mid-domain): the mean of the two central rows for even ny, else the central row
__global__ void k_abs_sum_xz(SliceView a, int c, int nx, int ny, double *out)
{
    hpb_pdl_prologue();
    __shared__ double sm[kThreads];
    double acc = 0.0;
    const double *p = a.comp(c);
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
        const double v = (ny % 2 == 0) ? 0.5 * (p[a.idx(i, ny / 2 - 1)] + p[a.idx(i, ny / 2)]) : p[a.idx(i, ny / 2)];
        acc += fabs(v);
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int st = kThreads / 2; st > 0; st >>= 1) {
        if (threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(out, sm[0]);
}

inline unsigned nb(long n) { return (unsigned)((n + kThreads - 1) / kThreads); }

}  // namespace

// exposed to poisson.cu (fused driver)
int hpb_launch_exmby_eypbx(hpb_ctx *ctx, const hpb_slice &sl, const int *comps)
{
    const hpb_geom &g = ctx->g;
    const int ng = -sl.lo_x;
    if (ng < 1 || ng > 3 || sl.lo_y != sl.lo_x) { hpb_set_error("slice with %d guard cells", ng); return HPB_ERR_ARG; }
    const int gx = g.nx + 2 * (ng - 1), gy = g.ny + 2 * (ng - 1);
    dim3 grid((gx + kThreads - 1) / kThreads, gy);
#define HPB_EXMBY(G) hpb_launch(k_exmby_eypbx<G>, grid, kThreads, 0, ctx->stream,                   \
        make_view(sl), comps[HPB_C_PSI], comps[HPB_C_EXMBY], comps[HPB_C_EYPBX], g.nx, g.ny,     \
        0.5 * (1.0 / g.dx), 0.5 * (1.0 / g.dy))
    if (ng == 2) HPB_EXMBY(2); else if (ng == 1) HPB_EXMBY(1); else HPB_EXMBY(3);
#undef HPB_EXMBY
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_fields_grid_current(hpb_ctx *ctx, hpb_slice sl, int c_jz_beam, double peak,
                                       const double mean[3], const double std[3], double plo_x,
                                       double plo_y, double z)
{
    if (!ctx || c_jz_beam < 0 || !mean || !std) return HPB_ERR_ARG;
    const hpb_geom &g = ctx->g;
    const double delta_z = (z - mean[2]) / std[2];
    const double long_pos_factor = exp(-0.5 * (delta_z * delta_z));
    dim3 grid((g.nx + kThreads - 1) / kThreads, g.ny);
    hpb_launch(k_grid_current, grid, kThreads, 0, ctx->stream, make_view(sl), c_jz_beam, g.nx, g.ny, plo_x,
               plo_y, g.dx, g.dy, mean[0], mean[1], std[0], std[1], peak * long_pos_factor);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

// setVal(0., ...) on a list of components (grown box)
extern "C" int hpb_fields_zero(hpb_ctx *ctx, hpb_slice sl, const int *comp_list, int n)
{
    if (!ctx || !comp_list || n < 0 || n > 12) return HPB_ERR_ARG;
    CompList cl;
    cl.n = 0;
    for (int k = 0; k < n; ++k) if (comp_list[k] >= 0) cl.c[cl.n++] = comp_list[k];
    if (cl.n == 0) return HPB_OK;
    const long ntot = (long)sl.jstride * sl.ny_tot;
    hpb_launch(k_zero, nb(ntot), kThreads, 0, ctx->stream, make_view(sl), cl, ntot);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_fields_initialize_slices(hpb_ctx *ctx, hpb_slice sl, const int *comps)
{
    if (!ctx || !comps) return HPB_ERR_ARG;
    // Fields::InitializeSlices, explicit branch (Fields.cpp:551-560, 578-580)
    CompList cl;
    cl.n = 0;
    const int names[] = {HPB_C_CHI, HPB_C_SY, HPB_C_SX, HPB_C_EXMBY, HPB_C_EYPBX, HPB_C_JZ_BEAM,
                         HPB_C_RHOMJZ, HPB_C_NEXT_JX_BEAM, HPB_C_NEXT_JY_BEAM, HPB_C_RHO};
    for (int k : names) if (comps[k] >= 0) cl.c[cl.n++] = comps[k];
    const long ntot = (long)sl.jstride * sl.ny_tot;
    hpb_launch(k_zero, nb(ntot), kThreads, 0, ctx->stream, make_view(sl), cl, ntot);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_fields_add_rho_ions(hpb_ctx *ctx, hpb_slice sl, const int *comps)
{
    if (!ctx || !comps) return HPB_ERR_ARG;
    if (comps[HPB_C_IONS_RHOMJZ] < 0) return HPB_OK;     // Fields.cpp:609
    CompPairs cp;
    cp.n = 0;
    cp.dst[cp.n] = comps[HPB_C_RHOMJZ]; cp.src[cp.n++] = comps[HPB_C_IONS_RHOMJZ];
    if (comps[HPB_C_RHO] >= 0) { cp.dst[cp.n] = comps[HPB_C_RHO]; cp.src[cp.n++] = comps[HPB_C_IONS_RHOMJZ]; }
    const long ntot = (long)sl.jstride * sl.ny_tot;
    hpb_launch(k_add, nb(ntot), kThreads, 0, ctx->stream, make_view(sl), cp, ntot);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_fields_shift_slices(hpb_ctx *ctx, hpb_slice sl, const int *comps)
{
    if (!ctx || !comps) return HPB_ERR_ARG;
    // Fields::ShiftSlices, explicit branch (Fields.cpp:596-599): Previous <- This (jx_beam,
    // jy_beam), then This{jx_beam,jy_beam,jx,jy} <- Next{jx_beam,jy_beam,jx_beam,jy_beam}.
    // One kernel: all sources are read before any destination is written.
    CompPairs cp;
    cp.n = 6;
    cp.dst[0] = comps[HPB_C_PREV_JX_BEAM]; cp.src[0] = comps[HPB_C_JX_BEAM];
    cp.dst[1] = comps[HPB_C_PREV_JY_BEAM]; cp.src[1] = comps[HPB_C_JY_BEAM];
    cp.dst[2] = comps[HPB_C_JX_BEAM];      cp.src[2] = comps[HPB_C_NEXT_JX_BEAM];
    cp.dst[3] = comps[HPB_C_JY_BEAM];      cp.src[3] = comps[HPB_C_NEXT_JY_BEAM];
    cp.dst[4] = comps[HPB_C_JX];           cp.src[4] = comps[HPB_C_NEXT_JX_BEAM];
    cp.dst[5] = comps[HPB_C_JY];           cp.src[5] = comps[HPB_C_NEXT_JY_BEAM];
    const long ntot = (long)sl.jstride * sl.ny_tot;
    hpb_launch(k_copy, nb(ntot), kThreads, 0, ctx->stream, make_view(sl), cp, ntot);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

// Our addition (no reference counterpart as one call): Fields::ShiftSlices (Fields.cpp:596-599) of
// this slice followed by Fields::InitializeSlices (:551-560, 578-580) and AddRhoIons (:606-615) of
// the next one, as ONE pass.  The three (Previous, This, Next) planes of jx_beam / jy_beam are
// rotated through the component table `comps` (in/out) instead of being copied; ExmBy / EypBx need
// no zero-fill (k_exmby_eypbx rewrites every cell it ever wrote) and Sx / Sy get theirs from
// k_sxsy_from_beam.  rhomjz starts from the ion background instead of 0 (+ AddRhoIons later).
extern "C" int hpb_fields_shift_and_initialize(hpb_ctx *ctx, hpb_slice sl, int *comps)
{
    if (!ctx || !comps) return HPB_ERR_ARG;
    // rotation: Previous <- This <- Next <- (old Previous, to be zeroed)
    const int old_prev_x = comps[HPB_C_PREV_JX_BEAM], old_prev_y = comps[HPB_C_PREV_JY_BEAM];
    comps[HPB_C_PREV_JX_BEAM] = comps[HPB_C_JX_BEAM]; comps[HPB_C_PREV_JY_BEAM] = comps[HPB_C_JY_BEAM];
    comps[HPB_C_JX_BEAM] = comps[HPB_C_NEXT_JX_BEAM]; comps[HPB_C_JY_BEAM] = comps[HPB_C_NEXT_JY_BEAM];
    comps[HPB_C_NEXT_JX_BEAM] = old_prev_x;           comps[HPB_C_NEXT_JY_BEAM] = old_prev_y;
    ShiftInit si;
    si.n_zero = 0; si.n_copy = 0;
    for (int k : {HPB_C_CHI, HPB_C_JZ_BEAM, HPB_C_NEXT_JX_BEAM, HPB_C_NEXT_JY_BEAM})
        si.zero[si.n_zero++] = comps[k];
    si.dst[si.n_copy] = comps[HPB_C_JX]; si.src[si.n_copy++] = comps[HPB_C_JX_BEAM];
    si.dst[si.n_copy] = comps[HPB_C_JY]; si.src[si.n_copy++] = comps[HPB_C_JY_BEAM];
    for (int k : {HPB_C_RHOMJZ, HPB_C_RHO}) {
        if (comps[k] < 0) continue;
        if (comps[HPB_C_IONS_RHOMJZ] >= 0) { si.dst[si.n_copy] = comps[k]; si.src[si.n_copy++] = comps[HPB_C_IONS_RHOMJZ]; }
        else si.zero[si.n_zero++] = comps[k];
    }
    const long ntot = (long)sl.jstride * sl.ny_tot;
    hpb_launch(k_shift_init, nb(ntot), kThreads, 0, ctx->stream, make_view(sl), si, ntot);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_fields_sxsy_from_beam(hpb_ctx *ctx, hpb_slice sl, const int *comps)
{
    if (!ctx || !comps) return HPB_ERR_ARG;
    const hpb_geom &g = ctx->g;
    const int ng = -sl.lo_x;
    if (ng < 1 || ng > 3 || sl.lo_y != sl.lo_x) { hpb_set_error("slice with %d guard cells", ng); return HPB_ERR_ARG; }
    dim3 grid((g.nx + 2 * ng + kThreads - 1) / kThreads, g.ny + 2 * ng);
#define HPB_SXSY(G) hpb_launch(k_sxsy_from_beam<G>, grid, kThreads, 0, ctx->stream,                 \
        make_view(sl), comps[HPB_C_SX], comps[HPB_C_SY], comps[HPB_C_NEXT_JX_BEAM],              \
        comps[HPB_C_NEXT_JY_BEAM], comps[HPB_C_JZ_BEAM], comps[HPB_C_PREV_JX_BEAM],              \
        comps[HPB_C_PREV_JY_BEAM], g.nx, g.ny, g.mu0, g.dx, g.dy, g.dz)
    if (ng == 2) HPB_SXSY(2); else if (ng == 1) HPB_SXSY(1); else HPB_SXSY(3);
#undef HPB_SXSY
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_abs_sum_xz(hpb_ctx *ctx, hpb_slice sl, int c, double *d_out)
{
    if (!ctx || c < 0 || !d_out) return HPB_ERR_ARG;
    hpb_launch(k_abs_sum_xz, 1, kThreads, 0, ctx->stream, make_view(sl), c, ctx->g.nx, ctx->g.ny, d_out);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_abs_sum_multi(hpb_ctx *ctx, hpb_slice sl, const int *comp_list, const int *slots, int n,
                                 double *d_out)
{
    if (!ctx || !comp_list || !slots || !d_out || n < 0) return HPB_ERR_ARG;
    const hpb_geom &g = ctx->g;
    for (int k0 = 0; k0 < n; k0 += 24) {
        AbsSumList l;
        const int m = n - k0 < 24 ? n - k0 : 24;
        for (int k = 0; k < m; ++k) {
            if (comp_list[k0 + k] < 0 || comp_list[k0 + k] >= sl.ncomp) return HPB_ERR_ARG;
            l.c[k] = comp_list[k0 + k]; l.slot[k] = slots[k0 + k];
        }
        const unsigned bx = g.ny < 74 ? g.ny : 74;          // 74 x 16 components = 8 CTAs per SM
        hpb_launch(k_abs_sum_multi, dim3(bx, m), kThreads, 0, ctx->stream, make_view(sl), l, g.nx, g.ny, d_out);
        hpb_count_launch(ctx);
    }
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_abs_sum(hpb_ctx *ctx, hpb_slice sl, int c, double *d_out)
{
    if (!ctx || c < 0 || !d_out) return HPB_ERR_ARG;
    const hpb_geom &g = ctx->g;
    const long n = (long)g.nx * g.ny;
    unsigned blocks = nb(n);
    if (blocks > 592) blocks = 592;
    hpb_launch(k_abs_sum, blocks, kThreads, 0, ctx->stream, make_view(sl), c, g.nx, g.ny, d_out);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}
