// Particle kernels for every hipace.depos_order_xy / depos_derivative_type other than the default
// 2 / 2 (and, with HPB_GENERIC=1, for the default too as a cross-check of the specialised kernels):
// one thread per particle, fp64 RED scatter, no warp aggregation or staging.  The per-particle
// arithmetic is generic_order.cuh; the reference dispatches the same way at compile time
// (amrex::CompileTimeOptions<0,1,2,3>, e.g. ExplicitDeposition.cpp:62-67, PlasmaParticleAdvance.cpp:88).
#include "generic_order.cuh"

namespace {

constexpr int kThreads = 128;
struct RedAdd {
    __device__ __forceinline__ void operator()(double *p, double v) const { atomicAdd(p, v); }
};
struct Comps6 { int c[6]; };       // jx jy jz rho chi rhomjz
struct PlasmaSoA { double *r[HPB_PLASMA_NREAL]; uint64_t *idcpu; long np; };
PlasmaSoA soa(const hpb_plasma &pl)
{
    PlasmaSoA p;
    for (int i = 0; i < HPB_PLASMA_NREAL; ++i) p.r[i] = pl.r[i];
    p.idcpu = pl.idcpu; p.np = pl.np;
    return p;
}
inline unsigned nblocks(long n) { return (unsigned)((n + kThreads - 1) / kThreads); }

template <int ORDER, bool LASER>
__global__ void __launch_bounds__(kThreads)
k_gen_deposit_current(PlasmaSoA pl, SliceView a, Comps6 c, GenGrid gr, GenDepositPar par,
                      int *n_qsa_violation)
{
    hpb_pdl_prologue();
    const long ip = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= pl.np) return;
    const uint64_t idcpu = pl.idcpu[ip];
    if (!hpb_is_valid(idcpu)) return;
    const bool ok = gen_deposit_current<ORDER, LASER>(a, c.c, gr, par, pl.r[HPB_X][ip], pl.r[HPB_Y][ip],
                                                      pl.r[HPB_W][ip], pl.r[HPB_UX][ip], pl.r[HPB_UY][ip],
                                                      pl.r[HPB_PSI][ip], RedAdd());
    if (!ok) {
        if (n_qsa_violation) atomicAdd(n_qsa_violation, 1);
        pl.r[HPB_W][ip] = 0.0;
        pl.idcpu[ip] = hpb_make_invalid(idcpu);
    }
}

template <int ORDER>
__global__ void __launch_bounds__(kThreads)
k_gen_beam_deposit(hpb_beam_slice b, SliceView a, int c_jx, int c_jy, int c_jz, GenGrid gr,
                   double clightsq, double q_invvol)
{
    hpb_pdl_prologue();
    const long ip = (long)blockIdx.x * blockDim.x + threadIdx.x;
    // getNumParticles: the slipped particles behind d_np[0] are not deposited (BeamDepositCurrent.cpp:100)
    if (ip >= b.np || (b.d_np && ip >= (long)b.d_np[0])) return;
    if (!hpb_is_valid(b.idcpu[ip])) return;
    gen_beam_deposit<ORDER>(a, c_jx, c_jy, c_jz, gr, clightsq, q_invvol, b.x[ip], b.y[ip], b.w[ip],
                            b.ux[ip], b.uy[ip], b.uz[ip], RedAdd());
}

template <int ORDER, int DTYPE, bool LASER>
__global__ void __launch_bounds__(kThreads)
k_gen_explicit_deposition(PlasmaSoA pl, SliceView a, GenGrid gr, GenExplicitPar par)
{
    hpb_pdl_prologue();
    const long ip = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= pl.np) return;
    if (!hpb_is_valid(pl.idcpu[ip])) return;
    gen_explicit_deposition<ORDER, DTYPE, LASER>(a, gr, par, pl.r[HPB_X][ip], pl.r[HPB_Y][ip],
                                                 pl.r[HPB_W][ip], pl.r[HPB_UX][ip], pl.r[HPB_UY][ip],
                                                 pl.r[HPB_PSI][ip], RedAdd());
}

template <int ORDER, bool LASER>
__global__ void __launch_bounds__(kThreads)
k_gen_advance_plasma(PlasmaSoA pl, SliceView a, GenGrid gr, GenPushPar par)
{
    hpb_pdl_prologue();
    const long ip = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= pl.np) return;
    const uint64_t idcpu = pl.idcpu[ip];
    if (!hpb_is_valid(idcpu)) return;
    double st[5] = {pl.r[HPB_X_PREV][ip], pl.r[HPB_Y_PREV][ip], pl.r[HPB_UX_HALF][ip],
                    pl.r[HPB_UY_HALF][ip], pl.r[HPB_PSI_HALF][ip]};
    double out[5] = {pl.r[HPB_X][ip], pl.r[HPB_Y][ip], pl.r[HPB_UX][ip], pl.r[HPB_UY][ip], pl.r[HPB_PSI][ip]};
    const bool alive = gen_advance_plasma<ORDER, LASER>(a, gr, par, st, out);
    // what a sub-cycle stored before a later one lost the particle stays stored, like the
    // reference's in-place updates (PlasmaParticleAdvance.cpp:183-217)
    pl.r[HPB_X][ip] = out[0]; pl.r[HPB_Y][ip] = out[1];
    pl.r[HPB_UX][ip] = out[2]; pl.r[HPB_UY][ip] = out[3]; pl.r[HPB_PSI][ip] = out[4];
    if (!par.temp_slice) {
        pl.r[HPB_X_PREV][ip] = st[0]; pl.r[HPB_Y_PREV][ip] = st[1];
        pl.r[HPB_UX_HALF][ip] = st[2]; pl.r[HPB_UY_HALF][ip] = st[3]; pl.r[HPB_PSI_HALF][ip] = st[4];
    }
    if (!alive) {
        pl.r[HPB_W][ip] = 0.0;
        pl.idcpu[ip] = hpb_make_invalid(idcpu);
    }
}

GenGrid grid_of(const hpb_geom &g) { return {g.x_off, g.y_off, 1.0 / g.dx, 1.0 / g.dy}; }
double invvol_of(const hpb_geom &g)
{
    return g.normalized ? 1.0 : (1.0 / g.dx) * (1.0 / g.dy) * (1.0 / g.dz);
}

// the slice must carry the guard cells the order needs (Fields.cpp:63-64)
int check_guards(const hpb_ctx *ctx, const hpb_slice &sl)
{
    const int need = HPB_NGUARD_OF(ctx->depos_order);
    if (-sl.lo_x < need || -sl.lo_y < need) {
        hpb_set_error("deposition order %d needs %d guard cells, the slice has %d / %d", ctx->depos_order,
                      need, -sl.lo_x, -sl.lo_y);
        return HPB_ERR_ARG;
    }
    return HPB_OK;
}

}  // namespace

#define HPB_BY_ORDER(order, CALL)                                                                  \
    switch (order) {                                                                               \
    case 0: { constexpr int O = 0; CALL; } break;                                                  \
    case 1: { constexpr int O = 1; CALL; } break;                                                  \
    case 2: { constexpr int O = 2; CALL; } break;                                                  \
    default: { constexpr int O = 3; CALL; } break;                                                 \
    }

int hpb_gen_deposit_current(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge, double mass,
                            int c_jx, int c_jy, int c_jz, int c_rho, int c_chi, int c_rhomjz, int c_aabs,
                            double max_qsa, int *d_n_qsa_violation)
{
    if (int rc = check_guards(ctx, sl)) return rc;
    const hpb_geom &g = ctx->g;
    const Comps6 c = {{c_jx, c_jy, c_jz, c_rho, c_chi, c_rhomjz}};
    GenDepositPar par = {1.0 / g.c, charge * invvol_of(g), charge * g.mu0 / mass, max_qsa,
                         (charge / g.q_e) * (g.m_e / mass) * (charge / g.q_e) * (g.m_e / mass), c_aabs, g.c};
#define HPB_CALL(LAS) hpb_launch(k_gen_deposit_current<O, LAS>, nblocks(pl.np), kThreads, 0, ctx->stream, \
                                 soa(pl), make_view(sl), c, grid_of(g), par, d_n_qsa_violation)
    if (c_aabs >= 0) { HPB_BY_ORDER(ctx->depos_order, HPB_CALL(true)) }
    else { HPB_BY_ORDER(ctx->depos_order, HPB_CALL(false)) }
#undef HPB_CALL
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

int hpb_gen_beam_deposit(hpb_ctx *ctx, hpb_beam_slice bm, hpb_slice sl, double charge, int c_jx,
                         int c_jy, int c_jz)
{
    if (int rc = check_guards(ctx, sl)) return rc;
    const hpb_geom &g = ctx->g;
#define HPB_CALL hpb_launch(k_gen_beam_deposit<O>, nblocks(bm.np), kThreads, 0, ctx->stream, bm,      \
                            make_view(sl), c_jx, c_jy, c_jz, grid_of(g), 1.0 / (g.c * g.c),            \
                            charge * invvol_of(g))
    HPB_BY_ORDER(ctx->depos_order, HPB_CALL)
#undef HPB_CALL
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

int hpb_gen_explicit_deposition(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge, double mass,
                                const int *comps)
{
    if (int rc = check_guards(ctx, sl)) return rc;
    const hpb_geom &g = ctx->g;
    const double laser_fac = (g.m_e / g.q_e) * (g.m_e / g.q_e);        // ExplicitDeposition.cpp:55
    GenExplicitPar par = {comps[HPB_C_SY], comps[HPB_C_SX], comps[HPB_C_BZ], comps[HPB_C_EZ],
                          comps[HPB_C_EXMBY], comps[HPB_C_EYPBX], comps[HPB_C_AABS], g.c, 1.0 / g.c,
                          charge * invvol_of(g) * g.mu0, charge / mass, laser_fac};
    const bool las = par.c_aabs >= 0;
#define HPB_CALL_(D, LAS) hpb_launch(k_gen_explicit_deposition<O, D, LAS>, nblocks(pl.np), kThreads, 0, \
                                     ctx->stream, soa(pl), make_view(sl), grid_of(g), par)
#define HPB_CALL(D) do { if (las) HPB_CALL_(D, true); else HPB_CALL_(D, false); } while (0)
    switch (ctx->depos_dtype) {
    case 0:     // (order 0 with the analytic derivative is refused by hpb_set_deposition_order)
        switch (ctx->depos_order) {
        case 1: { constexpr int O = 1; HPB_CALL(0); } break;
        case 2: { constexpr int O = 2; HPB_CALL(0); } break;
        default: { constexpr int O = 3; HPB_CALL(0); } break;
        }
        break;
    case 1: HPB_BY_ORDER(ctx->depos_order, HPB_CALL(1)) break;
    default: HPB_BY_ORDER(ctx->depos_order, HPB_CALL(2)) break;
    }
#undef HPB_CALL
#undef HPB_CALL_
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

int hpb_gen_advance_plasma(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge, double mass,
                           int n_subcycles, int temp_slice, int particle_bc, const double bc_lo[2],
                           const double bc_hi[2], const int *comps)
{
    if (int rc = check_guards(ctx, sl)) return rc;
    const hpb_geom &g = ctx->g;
    GenPushPar par = {comps[HPB_C_PSI], comps[HPB_C_EZ], comps[HPB_C_BX], comps[HPB_C_BY], comps[HPB_C_BZ],
                      comps[HPB_C_AABS], g.c, charge / (mass * g.c), g.dz / n_subcycles, n_subcycles,
                      temp_slice, particle_bc, bc_lo[0], bc_lo[1], bc_hi[0], bc_hi[1],
                      (charge / g.q_e) * (g.m_e / mass) * (charge / g.q_e) * (g.m_e / mass)};
#define HPB_CALL(LAS) hpb_launch(k_gen_advance_plasma<O, LAS>, nblocks(pl.np), kThreads, 0, ctx->stream, \
                                 soa(pl), make_view(sl), grid_of(g), par)
    if (par.c_aabs >= 0) { HPB_BY_ORDER(ctx->depos_order, HPB_CALL(true)) }
    else { HPB_BY_ORDER(ctx->depos_order, HPB_CALL(false)) }
#undef HPB_CALL
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}
