// Shared device helpers for the hpb200 hot path (sm_100a, fp64).
// Shape-factor arithmetic follows src/particles/particles_utils/ShapeFactors.H of the reference
// (order 2, the default hipace.depos_order_xy and the one every BASELINE config uses; the other
// orders and derivative types: shapes.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hpb200.h"

#define HPB_G HPB_NGUARD

#define HPB_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            hpb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            return HPB_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

void hpb_set_error(const char *fmt, ...);
void hpb_count_launch(hpb_ctx *ctx, int n = 1);

// ---- programmatic dependent launch -----------------------------------------------------------
// The slice loop is a strictly sequential chain of ~45 short kernels; with stream serialisation
// every kernel pays the launch latency and the ramp-up of its predecessor's tail.  All hot-path
// kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel signals
// `launch_dependents` as its first instruction, so the CTAs of the next kernel become resident
// while this one drains, and every kernel executes `griddepcontrol.wait` (= predecessor grid
// complete and its memory visible) before it touches global memory.  Without the attribute
// (HPB_PDL=0) both instructions are no-ops.
__device__ __forceinline__ void hpb_pdl_prologue()
{
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
bool hpb_pdl_enabled();
int hpb_bluestein_min_prime();     // context.cu: FFT lengths with a larger prime factor use Bluestein

template <class... KA, class... A>
inline cudaError_t hpb_launch(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, A &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = hpb_pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<A &&>(args)...);
}

// validity flag = top bit of idcpu (AMReX >= 24 ParticleIDWrapper packing)
#define HPB_ID_VALID_BIT (uint64_t(1) << 63)
__host__ __device__ inline bool hpb_is_valid(uint64_t idcpu) { return (idcpu >> 63) != 0; }
__host__ __device__ inline uint64_t hpb_make_invalid(uint64_t idcpu) { return idcpu & ~HPB_ID_VALID_BIT; }
__host__ __device__ inline uint64_t hpb_make_idcpu(uint64_t id, uint32_t cpu)
{
    return HPB_ID_VALID_BIT | (id << 24) | cpu;
}

// Array3 view: a(i,j,n) over the grown box
struct SliceView {
    double *p;
    int lo_x, lo_y;
    long jstride, nstride;
    __host__ __device__ inline double *comp(int n) const { return p + n * nstride; }
    __host__ __device__ inline long idx(int i, int j) const
    {
        return (long)(i - lo_x) + (long)(j - lo_y) * jstride;
    }
    __device__ inline double &operator()(int i, int j, int n) const { return p[idx(i, j) + n * nstride]; }
};
inline SliceView make_view(const hpb_slice &s)
{
    SliceView v;
    v.p = s.p; v.lo_x = s.lo_x; v.lo_y = s.lo_y; v.jstride = s.jstride; v.nstride = s.nstride;
    return v;
}

// ---- shape factors, order 2 ----------------------------------------------------------------

// compute_single_shape_factor<false,2>, ShapeFactors.H:165-174: returns leftmost cell
__device__ __forceinline__ int shape2(double xmid, double s[3])
{
    const double xfloor = floor(xmid + 0.5);
    const double xint = xmid - xfloor;
    s[0] = 0.5 * (0.5 - xint) * (0.5 - xint);
    s[1] = 0.75 - xint * xint;
    s[2] = 0.5 * (0.5 + xint) * (0.5 + xint);
    return (int)xfloor - 1;
}

// single_derivative_shape_factor<2,2> (centred derivative), ShapeFactors.H:405-430
// ds[] already carries the minus sign ("-sdx"). Returns leftmost cell of the 5-point stencil.
__device__ __forceinline__ int dshape2_centered(double xmid, double s[5], double ds[5])
{
    xmid += 0.5;
    const double xfloor = floor(xmid);
    const double xint = xmid - xfloor;
    const double xint_2 = xint * xint;
    s[0] = 0.;
    s[1] = 0.5 * xint_2 - xint + 0.5;
    s[2] = -xint_2 + xint + 0.5;
    s[3] = 0.5 * xint_2;
    s[4] = 0.;
    ds[0] = -(-0.25 * xint_2 + 0.5 * xint - 0.25);
    ds[1] = -(0.5 * xint_2 - 0.5 * xint - 0.25);
    ds[2] = -(0.25 - 0.5 * xint);
    ds[3] = -(-0.5 * xint_2 + 0.5 * xint + 0.25);
    ds[4] = -(0.25 * xint_2);
    return (int)xfloor - 2;
}

// single_derivative_shape_factor<1,2> (nodal derivative), ShapeFactors.H:305-329
__device__ __forceinline__ int dshape2_nodal(double xmid, double s[4], double ds[4])
{
    const double xfloor = floor(xmid);
    const double xint = xmid - xfloor;
    const double xint_2 = xint * xint;
    const bool lo = xint < 0.5;
    s[0] = lo ? 0.5 * xint_2 - 0.5 * xint + 0.125 : 0.;
    s[1] = lo ? 0.75 - xint_2 : 0.5 * xint_2 - 1.5 * xint + 1.125;
    s[2] = lo ? 0.5 * xint_2 + 0.5 * xint + 0.125 : -xint_2 + 2. * xint - 0.25;
    s[3] = lo ? 0. : 0.5 * xint_2 - 0.5 * xint + 0.125;
    ds[0] = -(-0.5 * xint_2 + xint - 0.5);
    ds[1] = -(1.5 * xint_2 - 2. * xint);
    ds[2] = -(-1.5 * xint_2 + xint + 0.5);
    ds[3] = -(0.5 * xint_2);
    return (int)xfloor - 1;
}

// doGatherShapeN<2> (src/particles/particles_utils/FieldGather.H:45-96): the six field values at
// (xp, yp) from Psi, Ez, Bx, By, Bz with nodal-derivative shapes over 4 x 4 cells
struct GatheredFields { double ExmBy, EypBx, Ez, Bx, By, Bz; };
__device__ __forceinline__ GatheredFields gather_order2(const SliceView &a, int c_psi, int c_ez,
                                                        int c_bx, int c_by, int c_bz, double x_off,
                                                        double y_off, double dx_inv, double dy_inv,
                                                        double xp, double yp)
{
    double sx[4], dsx[4], sy[4], dsy[4];
    const int i0 = dshape2_nodal((xp - x_off) * dx_inv, sx, dsx);
    const int j0 = dshape2_nodal((yp - y_off) * dy_inv, sy, dsy);
    const double *Psi = a.comp(c_psi), *Ez = a.comp(c_ez), *Bx = a.comp(c_bx);
    const double *By = a.comp(c_by), *Bz = a.comp(c_bz);
    GatheredFields f = {0., 0., 0., 0., 0., 0.};
#pragma unroll
    for (int iy = 0; iy < 4; ++iy) {
#pragma unroll
        for (int ix = 0; ix < 4; ++ix) {
            const long o = a.idx(i0 + ix, j0 + iy);
            const double psi_v = Psi[o];
            f.ExmBy += (dsx[ix] * sy[iy]) * psi_v * dx_inv;
            f.EypBx += (sx[ix] * dsy[iy]) * psi_v * dy_inv;
            const double w = sx[ix] * sy[iy];
            f.Ez += w * Ez[o];
            f.Bx += w * Bx[o];
            f.By += w * By[o];
            f.Bz += w * Bz[o];
        }
    }
    return f;
}

// doLaserGatherShapeN<2> (FieldGather.H:162-222, 236-283): |a|^2 at the particle with the plain
// order-2 shape; DERIV: also its centred x / y differences taken on the grid and gathered
template <bool DERIV>
__device__ __forceinline__ void laser_gather(const SliceView &a, int c_aabs, double x_off, double y_off,
                                             double dx_inv, double dy_inv, double xp, double yp,
                                             double &A, double &ADx, double &ADy)
{
    double sx[3], sy[3];
    const int i0 = shape2((xp - x_off) * dx_inv, sx);
    const int j0 = shape2((yp - y_off) * dy_inv, sy);
    const double *ab = a.comp(c_aabs);
    const long js = a.jstride;
    A = 0.; ADx = 0.; ADy = 0.;
#pragma unroll
    for (int iy = 0; iy < 3; ++iy) {
#pragma unroll
        for (int ix = 0; ix < 3; ++ix) {
            const long o = a.idx(i0 + ix, j0 + iy);
            const double w = sx[ix] * sy[iy];
            A += w * ab[o];
            if (DERIV) {
                ADx += w * 0.5 * dx_inv * (ab[o + 1] - ab[o - 1]);
                ADy += w * 0.5 * dy_inv * (ab[o + js] - ab[o - js]);
            }
        }
    }
}

// EnforceBC, src/particles/pusher/GetAndSetPosition.H:56-98.  Returns true if the particle
// became invalid (absorbing boundary).
__host__ __device__ __forceinline__ bool enforce_particle_bc(double &x, double &y, double &ux, double &uy,
                                                    int bc, double lox, double loy, double hix,
                                                    double hiy)
{
    if (x < lox || y < loy || x > hix || y > hiy) {
        const double len_x = hix - lox, len_y = hiy - loy;
        if (bc == HPB_BC_REFLECTING) {
            x = fmod(x - lox, 2 * len_x); if (x < 0) x += 2 * len_x; x += lox;
            if (x > hix) { x = 2 * hix - x; ux = -ux; }
            y = fmod(y - loy, 2 * len_y); if (y < 0) y += 2 * len_y; y += loy;
            if (y > hiy) { y = 2 * hiy - y; uy = -uy; }
        } else if (bc == HPB_BC_PERIODIC) {
            x = fmod(x - lox, len_x); if (x < 0) x += len_x; x += lox;
            y = fmod(y - loy, len_y); if (y < 0) y += len_y; y += loy;
        } else {
            return true;
        }
    }
    return false;
}

// fp64 reduction without return value (RED.E.ADD.F64 on sm_100a)
__device__ __forceinline__ void red_add(double *addr, double v) { atomicAdd(addr, v); }

// ---- the context ---------------------------------------------------------------------------
struct MGLevel {
    int nx, ny;            // points in the level box (cells if cc, nodes incl. boundary if nodal)
    double *acf, *c0i, *res, *cor, *rescor;   // res/cor/rescor: 2 comps, comp stride nx*ny
};

struct hpb_ctx {
    hpb_geom g;
    cudaStream_t stream;
    long n_launch;
    // Poisson (poisson.cu)
    int fftN;                 // nx + 1
    int fftM;                 // > 0: Bluestein length (nx + 1 has a prime factor > 64), fft_smem.cuh
    double2 *d_chirp, *d_bhat;
    int nrad; int radices[32];
    double *d_cs_cos[32], *d_cs_sin[32];         // per odd-prime stage: DFT-p cos / sin tables
    double *d_cs_frag[32];                       // ... and the same in mma.m8n8k4 fragment order (fft_smem.cuh)
    double2 *d_root;          // exp(-2 pi i t / N), t = 0..N-1
    double *d_sinf;           // 1 / (2 sin(pi (i+1) / N)), i = 0..nx-1
    int th_L, th_C, th_last_base;                // partitioned Thomas: chunk length, chunks
    double *d_tri_m, *d_tri_c, *d_tri_p, *d_tri_q;   // chunk-local pivots and spikes [row][nx]
    double *d_red_pe, *d_red_pf, *d_red_b, *d_red_inv, *d_red_del;   // reduced system [C-1][nx]
    double *d_spec;           // 3 * nx * ny spectral scratch
    double *d_iface;          // 4 * 3 * C * nx chunk-interface values
    // Multigrid (mg.cu)
    int mg_cc; int mg_nlev; int mg_lc; int mg_coarse_smem;   // mg_lc: first level solved by the single-CTA kernel
    MGLevel mg[32];
    double *d_mg_norm;        // [0]=res norm, [1]=rhs norm (atomicMax accumulators)
    double *d_mg_state;       // target, max_norm, last norm
    int *d_mg_istate;         // done, V-cycles, failed
    double *h_mg_norm;        // pinned copies
    int *h_mg_istate;
    int mg_last_iters;        // V-cycles of the previous solve = speculation depth of the next
    const double *mg_acf_ready;   // hpb_mg_prepare_acf ran for this coefficient plane (consumed by the next solve1)
    int tune_mg_lean;         // the lean interior-tile path of k_smooth
    int tune_mg_persist;      // the mid levels + the single-CTA levels of a V-cycle in one persistent launch (k_mid)
    void *d_mg_bar;           // its grid barrier
    int tune_mg_rotate;       // level-0 buffer rotation: the last smoother of a V-cycle writes into sol (no final copy)
    // misc
    int *d_scalar_i;          // scratch ints
    // particle order hint (hpb_set_plasma_lattice_hint): the next particle-kernel calls see np =
    // order_ppc passes of order_n lattice cells each (InitParticles order); 0 = unknown
    long order_n; int order_ppc;
    // hipace.depos_order_xy / hipace.depos_derivative_type (hpb_set_deposition_order; default 2 / 2:
    // the specialised kernels of particles.cu.  Anything else: generic_order.cu)
    int depos_order, depos_dtype;
    int force_generic;        // run the generic kernels for the default order too (cross-check)
    // kernel variants for A/B measurements (hpb_set_option; never read from the environment)
    int tune_order;           // bit mask of the particle kernels using the pass-interleaved map (default 1)
    int tune_expl_variant, tune_push_variant, tune_fft_variant, tune_mg_wide, tune_mg_fuse;
    // TMA tensor maps over the caller's slice array (tma.cuh), re-encoded when the view changes:
    // [0] the 40 x 6 gather patch of the push, [1] the patch of the explicit deposition
    void *mg2;                // mg.cu: level coefficient arrays of the type-2 (complex Helmholtz) solver
    void *periodic;           // periodic.cu: 2-D FFT plan and complex planes of the FFTPeriodic Poisson solver
    void *ref_arm;            // ref_gpu_arm.cu: cuFFT plans and buffers of the reference-algorithm arm
    int tune_poisson_impl;    // 0: product solver, 1: the reference's DirichletFast sequence on cuFFT
    // plasma reordering scratch (reorder.cu)
    unsigned *d_reorder_key, *d_reorder_rank, *d_reorder_hist, *d_reorder_sums;
    long reorder_np_cap, reorder_bins_cap;
    alignas(64) unsigned char tmap[3][128];      // [2]: the 136 x 6 patch of the row-tile push
    hpb_slice tmap_key[3];
    int tmap_ok[3];
};
// the cached tensor map `which` for boxes of box_w x box_h cells of sl, or nullptr (-> cp.async path)
const void *hpb_slice_tmap(hpb_ctx *ctx, int which, const hpb_slice &sl, int box_w, int box_h);
// true when the particle kernels must take the generic-order path
bool hpb_use_generic_order(const hpb_ctx *ctx);
// fields.cu: ExmBy / EypBx from Psi (Fields.cpp:931-956), shared by the fused and the staged Poisson paths
int hpb_launch_exmby_eypbx(hpb_ctx *ctx, const hpb_slice &sl, const int *comps);
// generic_order.cu: the entry points of particles.cu / beam.cu forward here in that case
int hpb_gen_deposit_current(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge, double mass,
                            int c_jx, int c_jy, int c_jz, int c_rho, int c_chi, int c_rhomjz, int c_aabs,
                            double max_qsa, int *d_n_qsa_violation);
int hpb_gen_beam_deposit(hpb_ctx *ctx, hpb_beam_slice bm, hpb_slice sl, double charge, int c_jx,
                         int c_jy, int c_jz);
int hpb_gen_explicit_deposition(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge, double mass,
                                const int *comps);
int hpb_gen_advance_plasma(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge, double mass,
                           int n_subcycles, int temp_slice, int particle_bc, const double bc_lo[2],
                           const double bc_hi[2], const int *comps);
