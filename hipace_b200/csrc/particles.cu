// Plasma / beam particle kernels of the slice loop: current deposition, explicit (Sx,Sy)
// deposition, field gather + leap-frog push.  One thread per particle over the SoA arrays
// (coalesced 8-byte streams); scatter goes through fp64 RED atomics into the L2-resident slice.
//
// Reference arithmetic restated from:
//   src/particles/deposition/PlasmaDepositCurrent.cpp:155-246
//   src/particles/deposition/ExplicitDeposition.cpp:140-261
//   src/particles/deposition/BeamDepositCurrent.cpp:136-194
//   src/particles/particles_utils/FieldGather.H:45-96
//   src/particles/pusher/PlasmaParticleAdvance.cpp:92-217, PushPlasmaParticles.H:39-75,
//   src/utils/DualNumbers.H:13-43, src/particles/pusher/GetAndSetPosition.H:29-99
#include "common.cuh"
#include "push_math.cuh"
#include "tma.cuh"
#include <string.h>

namespace {

constexpr int kThreads = 256;

struct PlasmaPtrs {
    double *r[HPB_PLASMA_NREAL];
    uint64_t *idcpu;
    long np;
    long lat_n; int lat_ppc;        // lattice hint: np = lat_ppc passes of lat_n cells (0: none)
    int lat_mode;                   // 1: passes interleaved warp by warp, 2: CTA by CTA (experiment)
};
PlasmaPtrs to_ptrs(const hpb_plasma &pl)
{
    PlasmaPtrs p;
    for (int i = 0; i < HPB_PLASMA_NREAL; ++i) p.r[i] = pl.r[i];
    p.idcpu = pl.idcpu; p.np = pl.np;
    p.lat_n = 0; p.lat_ppc = 1; p.lat_mode = 0;
    return p;
}
// with the lattice hint of the context (only if it describes exactly this particle array).
// which: 1 = explicit deposition, 2 = gather + push (+ deposit).  HPB_ORDER is the bit mask of the
// kernels that use the pass-interleaved map; measured on the 1024^2 ppc 4 deck: explicit
// deposition 0.270 -> 0.253 ms (the planes it reads are fetched once), push 0.346 -> 0.389 ms (its
// fused deposit then has four warps of a CTA reducing into the same cells at the same time) --
// so the default is 1.
PlasmaPtrs to_ptrs(const hpb_ctx *ctx, const hpb_plasma &pl, int which)
{
    PlasmaPtrs p = to_ptrs(pl);
    const int on = ctx->tune_order;
    // bits 4 / 8 (= which << 2): the same, interleaved CTA by CTA instead of warp by warp -- the next
    // thing to measure for the push (ROADMAP.md): the ppc passes of a cell group then run in
    // consecutive CTAs (co-resident on different SMs) instead of in the four warps of one CTA, so the
    // planes are still fetched once but the fused deposit does not reduce into the same cells from
    // within a CTA.  Not timed yet; results are identical by construction (a permutation of threads).
    const bool cta = (on & (which << 2)) != 0 && which == 2;
    if (((on & which) || cta) && ctx->order_n > 0 && ctx->order_ppc > 1 && ctx->order_n * ctx->order_ppc == pl.np) {
        p.lat_n = ctx->order_n; p.lat_ppc = ctx->order_ppc; p.lat_mode = cta ? 2 : 1;
    }
    return p;
}
// Pass-interleaved thread -> particle map: warp w works on pass (w % ppc) of cell group (w / ppc).
// `per_warp` consecutive cells are owned per warp, `lead` extra lanes in front of them (feed-only
// neighbours of the aggregation).  Returns the particle index and whether the lane holds one.
// LAT = false is the plain linear map (the code the kernels had before the hint existed).
// LAT = 2 (UPW = warps per CTA): the same with whole CTAs as the interleaving unit -- CTA c works on
// pass (c % ppc) of the UPW consecutive cell groups (c / ppc) * UPW ...
template <int LAT, int UPW = 1>
__host__ __device__ __forceinline__ long lattice_particle(const PlasmaPtrs &pl, long warp, int lane, int per_warp,
                                                           int lead, bool &in_range)
{
    if (LAT == 0) {
        const long ip = warp * per_warp - lead + lane;
        in_range = ip >= 0 && ip < pl.np;
        return ip;
    }
    const unsigned w = (unsigned)warp, ppc = (unsigned)pl.lat_ppc;      // 32-bit division only
    if (LAT == 2) {
        const unsigned unit = w / UPW, sub = w - unit * UPW;
        const unsigned group = unit / ppc, pass = unit - group * ppc;
        const long cell = ((long)group * UPW + sub) * per_warp - lead + lane;
        in_range = cell >= 0 && cell < pl.lat_n;
        return (long)pass * pl.lat_n + cell;
    }
    const unsigned group = w / ppc, pass = w - group * ppc;
    const long cell = (long)group * per_warp - lead + lane;
    in_range = cell >= 0 && cell < pl.lat_n;
    return (long)pass * pl.lat_n + cell;
}
// number of warps a launch needs (upw: warps per CTA, only used by the CTA-interleaved map)
inline long lattice_warps(const PlasmaPtrs &pl, int per_warp, int upw = 1)
{
    if (pl.lat_n <= 0) return (pl.np + per_warp - 1) / per_warp;
    const long groups = (pl.lat_n + per_warp - 1) / per_warp;
    if (pl.lat_mode == 2) return (groups + upw - 1) / upw * upw * pl.lat_ppc;
    return groups * pl.lat_ppc;
}

// -------------------------------------------------------------------------------------------
// DepositCurrent
// -------------------------------------------------------------------------------------------
// Warp-aggregated scatter.  Plasma particles keep the lattice order of InitParticles (cells
// x-fastest), so the 32 lanes of a warp usually sit in 32 consecutive cells of one row and their
// 3-wide stencils overlap: a lane whose left / right neighbour lane is exactly one cell to the
// left / right (and on the same rows) takes over that neighbour's contribution to its own centre
// column through warp shuffles.  An aligned particle then issues 3 fp64 reductions per component
// (its centre column) instead of 9; unaligned pairs fall back to their own reductions, so the
// result is the same sum for any particle order.
constexpr unsigned kFull = 0xffffffffu;

// The particle SoA streams are read / written exactly once per kernel (0.4 GB per pass, three
// times the L2): they are accessed with the streaming (evict-first) cache operator so that they
// do not push the slice planes -- which every particle gathers from and reduces into, and which
// fit the L2 several times over -- out of the cache.  (ncu before: 958 MB of DRAM traffic per
// push launch for 646 MB of algorithmic bytes; the planes were re-fetched once per ppc pass.)
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ uint64_t ld_stream(const uint64_t *p)
{
    return (uint64_t)__ldcs((const unsigned long long *)p);
}
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(uint64_t *p, uint64_t v)
{
    __stcs((unsigned long long *)p, (unsigned long long)v);
}
constexpr int kDepOwn = 30;      // particles deposited per warp (lanes 1..30)
constexpr int kExplOwn = 28;     // explicit deposition: lanes 2..29

// Scatter of one particle per lane with neighbour-lane aggregation.  Must be called by all 32
// lanes.  active: the lane holds a particle to deposit; owner: the lane issues reductions (the
// non-owner edge lanes of k_deposit_current only feed their neighbours).
template <bool JXY, bool RHO, bool CHI, bool RMJ>
__device__ __forceinline__ void
deposit_aggregated(const SliceView &a, int c_jx, int c_jy, int c_rho, int c_chi, int c_rhomjz,
                   bool active, bool owner, int lane, int i0, int j0, const double sx[3],
                   const double sy[3], double q_invvol, double vx_c, double vy_c, double gamma_psi,
                   double chi_fac)
{
    // centre column; inactive lanes get a sentinel that never aligns with a neighbour
    const int cc = active ? i0 + 1 : -(1 << 28) - 3 * lane;
    const int cL = __shfl_up_sync(kFull, cc, 1), jL = __shfl_up_sync(kFull, j0, 1);
    const int cR = __shfl_down_sync(kFull, cc, 1), jR = __shfl_down_sync(kFull, j0, 1);
    const bool L_ok = active && lane > 0 && cL == cc - 1 && jL == j0;
    const bool R_ok = active && lane < 31 && cR == cc + 1 && jR == j0;

    // per-row weights of the three stencil columns
    double P0[3], P1[3], P2[3], PL[3], PR[3];
#pragma unroll
    for (int iy = 0; iy < 3; ++iy) {
        P0[iy] = q_invvol * sx[0] * sy[iy];
        P1[iy] = q_invvol * sx[1] * sy[iy];
        P2[iy] = q_invvol * sx[2] * sy[iy];
        PL[iy] = __shfl_up_sync(kFull, P2[iy], 1);       // right column of the left neighbour
        PR[iy] = __shfl_down_sync(kFull, P0[iy], 1);     // left column of the right neighbour
        if (!L_ok) PL[iy] = 0.;
        if (!R_ok) PR[iy] = 0.;
    }
    // per-component multipliers: own, left neighbour's, right neighbour's
    const double vxL = __shfl_up_sync(kFull, vx_c, 1), vxR = __shfl_down_sync(kFull, vx_c, 1);
    const double vyL = __shfl_up_sync(kFull, vy_c, 1), vyR = __shfl_down_sync(kFull, vy_c, 1);
    const double gpL = RHO ? __shfl_up_sync(kFull, gamma_psi, 1) : 0., gpR = RHO ? __shfl_down_sync(kFull, gamma_psi, 1) : 0.;
    const double cfL = CHI ? __shfl_up_sync(kFull, chi_fac, 1) : 0., cfR = CHI ? __shfl_down_sync(kFull, chi_fac, 1) : 0.;
    if (!active || !owner) return;

    double *jx = JXY ? a.comp(c_jx) : nullptr, *jy = JXY ? a.comp(c_jy) : nullptr;
    double *rho = RHO ? a.comp(c_rho) : nullptr, *chi = CHI ? a.comp(c_chi) : nullptr;
    double *rmj = RMJ ? a.comp(c_rhomjz) : nullptr;
#pragma unroll
    for (int iy = 0; iy < 3; ++iy) {
        const long o = a.idx(cc, j0 + iy);
        if (JXY) {
            red_add(jx + o, P1[iy] * vx_c + PL[iy] * vxL + PR[iy] * vxR);
            red_add(jy + o, P1[iy] * vy_c + PL[iy] * vyL + PR[iy] * vyR);
        }
        if (RHO) red_add(rho + o, P1[iy] * gamma_psi + PL[iy] * gpL + PR[iy] * gpR);
        if (CHI) red_add(chi + o, P1[iy] * chi_fac + PL[iy] * cfL + PR[iy] * cfR);
        if (RMJ) red_add(rmj + o, P1[iy] + PL[iy] + PR[iy]);
        // a side column without an aligned neighbour lane is deposited by the particle itself
        if (!L_ok) {
            if (JXY) { red_add(jx + o - 1, P0[iy] * vx_c); red_add(jy + o - 1, P0[iy] * vy_c); }
            if (RHO) red_add(rho + o - 1, P0[iy] * gamma_psi);
            if (CHI) red_add(chi + o - 1, P0[iy] * chi_fac);
            if (RMJ) red_add(rmj + o - 1, P0[iy]);
        }
        if (!R_ok) {
            if (JXY) { red_add(jx + o + 1, P2[iy] * vx_c); red_add(jy + o + 1, P2[iy] * vy_c); }
            if (RHO) red_add(rho + o + 1, P2[iy] * gamma_psi);
            if (CHI) red_add(chi + o + 1, P2[iy] * chi_fac);
            if (RMJ) red_add(rmj + o + 1, P2[iy]);
        }
    }
}

template <bool JXY, bool RHO, bool CHI, bool RMJ, bool LASER>
__global__ void __launch_bounds__(kThreads)
k_deposit_current(PlasmaPtrs pl, SliceView a, int c_jx, int c_jy, int c_rho, int c_chi,
                  int c_rhomjz, double x_off, double y_off, double dx_inv, double dy_inv,
                  double clightinv, double charge_invvol, double charge_mu0_mass_ratio,
                  double max_qsa, int *n_qsa_violation, int c_aabs, double laser_norm)
{
    hpb_pdl_prologue();
    // warps overlap by one lane on each side: lanes 0 and 31 only feed their neighbours, so an
    // aligned run of particles never pays for warp-edge columns
    const int lane = threadIdx.x & 31;
    const long warp = (long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const long ip = warp * kDepOwn - 1 + lane;
    const bool owner = lane >= 1 && lane <= kDepOwn;
    bool active = ip >= 0 && ip < pl.np;
    uint64_t idcpu = 0;
    double psi_inv = 1., vx_c = 0., vy_c = 0., q_invvol = 0., gamma_psi = 1.;
    double sx[3] = {0., 0., 0.}, sy[3] = {0., 0., 0.};
    int i0 = 0, j0 = 0;
    if (active) {
        // all seven streams are requested before the first use (one HBM round trip, not two)
        idcpu = ld_stream(&pl.idcpu[ip]);
        const double psi = ld_stream(&pl.r[HPB_PSI][ip]);
        const double xp = ld_stream(&pl.r[HPB_X][ip]);
        const double yp = ld_stream(&pl.r[HPB_Y][ip]);
        const double ux = ld_stream(&pl.r[HPB_UX][ip]);
        const double uy = ld_stream(&pl.r[HPB_UY][ip]);
        const double w = ld_stream(&pl.r[HPB_W][ip]);
        active = hpb_is_valid(idcpu);
        psi_inv = 1.0 / psi;
        vx_c = ux * psi_inv;
        vy_c = uy * psi_inv;
        q_invvol = charge_invvol * w;
        if (LASER) {        // :182-195
            double A, ADx, ADy;
            laser_gather<false>(a, c_aabs, x_off, y_off, dx_inv, dy_inv, xp, yp, A, ADx, ADy);
            gamma_psi = 0.5 * ((1.0 + 0.5 * (A * laser_norm)) * psi_inv * psi_inv
                               + vx_c * vx_c * clightinv * clightinv
                               + vy_c * vy_c * clightinv * clightinv + 1.0);
        } else {
            gamma_psi = 0.5 * (psi_inv * psi_inv + vx_c * vx_c * clightinv * clightinv
                               + vy_c * vy_c * clightinv * clightinv + 1.0);
        }
        if (active && (gamma_psi < 0.0 || gamma_psi > max_qsa || psi_inv < 0.0)) {
            // QSA violation: discard the particle (PlasmaDepositCurrent.cpp:197-204)
            if (owner) {
                if (n_qsa_violation) atomicAdd(n_qsa_violation, 1);
                st_stream(&pl.r[HPB_W][ip], 0.0);
                st_stream(&pl.idcpu[ip], hpb_make_invalid(idcpu));
            }
            active = false;
        }
        if (active) {
            i0 = shape2((xp - x_off) * dx_inv, sx);
            j0 = shape2((yp - y_off) * dy_inv, sy);
        } else {
            q_invvol = 0.;
        }
    }
    deposit_aggregated<JXY, RHO, CHI, RMJ>(a, c_jx, c_jy, c_rho, c_chi, c_rhomjz, active, owner, lane,
                                           i0, j0, sx, sy, q_invvol, vx_c, vy_c, gamma_psi,
                                           charge_mu0_mass_ratio * psi_inv);
}

// -------------------------------------------------------------------------------------------
// beam DepositCurrentSlice
// -------------------------------------------------------------------------------------------
struct BeamPtrs { double *x, *y, *z, *w, *ux, *uy, *uz; uint64_t *idcpu; long np; const int64_t *d_np; };

__global__ void __launch_bounds__(kThreads)
k_beam_deposit(BeamPtrs b, SliceView a, int c_jx, int c_jy, int c_jz, double x_off, double y_off,
               double dx_inv, double dy_inv, double clightsq, double q_invvol)
{
    hpb_pdl_prologue();
    const long ip = (long)blockIdx.x * blockDim.x + threadIdx.x;
    // getNumParticles: the slipped particles behind d_np[0] are not deposited (BeamDepositCurrent.cpp:100)
    if (ip >= b.np || (b.d_np && ip >= (long)b.d_np[0])) return;
    if (!hpb_is_valid(b.idcpu[ip])) return;
    const double ux = b.ux[ip], uy = b.uy[ip], uz = b.uz[ip];
    const double gaminv = 1.0 / sqrt(1.0 + ux * ux * clightsq + uy * uy * clightsq + uz * uz * clightsq);
    const double wq = q_invvol * b.w[ip];       // q * w * invvol
    const double wqx = wq * (ux * gaminv), wqy = wq * (uy * gaminv), wqz = wq * (uz * gaminv);
    double sx[3], sy[3];
    const int i0 = shape2((b.x[ip] - x_off) * dx_inv, sx);
    const int j0 = shape2((b.y[ip] - y_off) * dy_inv, sy);
#pragma unroll
    for (int iy = 0; iy < 3; ++iy) {
#pragma unroll
        for (int ix = 0; ix < 3; ++ix) {
            const long o = a.idx(i0 + ix, j0 + iy);
            const double s = sx[ix] * sy[iy];
            if (c_jx >= 0) {
                red_add(a.comp(c_jx) + o, s * wqx);
                red_add(a.comp(c_jy) + o, s * wqy);
            }
            if (c_jz >= 0) red_add(a.comp(c_jz) + o, s * wqz);
        }
    }
}

// -------------------------------------------------------------------------------------------
// ExplicitDeposition (order 2, centred derivative: 5x5 stencil minus corners)
// -------------------------------------------------------------------------------------------
// Same warp aggregation, 5 columns wide: lane l owns the centre column of its particle for the 5
// stencil rows, loads the 4 field values of those 5 cells ONCE and adds the contributions of the
// particles of lanes l-2..l+2 whenever they are aligned (d cells to the side, same rows) using
// shuffled particle scalars; columns that no aligned neighbour owns are scattered by the particle
// itself.  An aligned particle costs 20 field loads + 10 reductions instead of 84 + 42.
// The per-cell expression is ExplicitDeposition.cpp:228-258 with the particle-only factors
// hoisted:  Sy += -ss hq Fy + sd1 gy1 + sd2 gy2,  Sx += ss hq Fx + sd1 gx1 + sd2 gx2.

// y part of single_derivative_shape_factor<2,2> from the in-cell offset (ShapeFactors.H:405-430)
__device__ __forceinline__ void dshape2_centered_frac(double xint, double s[5], double ds[5])
{
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
}

// particle-only factors for one stencil column (weights wx = shx[ix], wdx = shdx[ix]):
//   Sy(cell) += shy (A Fy + B) + shdy C,    Sx(cell) += shy (Ap Fx + Bp) + shdy Cp
//   Fy = -Bz vx + (Ez vy + ExmBy a3 + EypBx a4)/c,   Fx = Bz vy + (Ez vx + ExmBy a5 + EypBx a3)/c
struct ExplCol { double A, B, C, Ap, Bp, Cp; };
struct ExplPart { double vx, vy, a3, a4, a5, hq, gy1, gy2, gx1, gx2; };

__device__ __forceinline__ ExplPart expl_part(double vx, double vy, double gamma_psi, double cdm,
                                              double psi_inv, double q_mass_ratio, double a_clight)
{
    ExplPart e;
    e.vx = vx; e.vy = vy;
    e.a3 = -vx * vy;
    e.a4 = gamma_psi - vy * vy;
    e.a5 = gamma_psi - vx * vx;
    e.hq = cdm * (q_mass_ratio * psi_inv);
    const double cc = cdm * a_clight;
    e.gy1 = -e.a3 * cc;              // -shdx shy/dx (-vx vy) c
    e.gy2 = -(e.a4 - 1.0) * cc;      // -shx shdy/dy (gamma_psi - vy^2 - 1) c
    e.gx1 = (e.a5 - 1.0) * cc;       // +shdx shy/dx (gamma_psi - vx^2 - 1) c
    e.gx2 = e.a3 * cc;               // +shx shdy/dy (-vx vy) c
    return e;
}
__device__ __forceinline__ ExplCol expl_col(const ExplPart &e, double wx, double wdx, double dx_inv,
                                            double dy_inv)
{
    ExplCol k;
    const double wxd = wdx * dx_inv, wxy = wx * dy_inv;
    k.A = -wx * e.hq;  k.B = wxd * e.gy1;  k.C = wxy * e.gy2;
    k.Ap = wx * e.hq;  k.Bp = wxd * e.gx1; k.Cp = wxy * e.gx2;
    return k;
}
// lasy / lasx: the ponderomotive term inside the field bracket, 0.25 AabssqD{y,x}(cell)
// q_mass_ratio psi_inv(source) (ExplicitDeposition.cpp:234, 250); 0 without a laser
__device__ __forceinline__ void expl_cell(const ExplCol &k, double vx, double vy, double a3,
                                          double a4, double a5, double shy, double shdy, double Bz,
                                          double Ez, double ExmBy, double EypBx, double clight_inv,
                                          double &sy_out, double &sx_out, double lasy = 0.,
                                          double lasx = 0.)
{
    const double Fy = -Bz * vx + (Ez * vy + ExmBy * a3 + EypBx * a4) * clight_inv - lasy;
    const double Fx = Bz * vy + (Ez * vx + ExmBy * a5 + EypBx * a3) * clight_inv - lasx;
    sy_out += shy * (k.A * Fy + k.B) + shdy * k.C;
    sx_out += shy * (k.Ap * Fx + k.Bp) + shdy * k.Cp;
}

struct ExplLaser { int c_aabs; double fac_c, a_norm; };     // laser_fac * c, laser_fac * q_mass_ratio^2

// The particle-independent part of the explicit deposition of ONE group of lanes (one particle per lane,
// its seven stream values already in registers): everything of k_explicit_deposition after the loads.
// Must be called by all 32 lanes of the warp (shuffles inside).
struct ExplRaw { uint64_t idcpu; double psi, xp, yp, ux, uy, w; };
template <bool LASER>
__device__ __forceinline__ void
expl_group(const SliceView &a, int c_sy, int c_sx, int c_bz, int c_ez, int c_exmby, int c_eypbx, double x_off,
           double y_off, double dx_inv, double dy_inv, double a_clight, double clight_inv,
           double charge_invvol_mu0, double q_mass_ratio, const ExplLaser &las, int lane, bool active, bool owner,
           const ExplRaw &raw)
{
    double vx = 0., vy = 0., gamma_psi = 1., yint = 0.;
    double qp = 0.;             // 0.25 q_mass_ratio psi_inv (laser term factor of this particle)
    double sx[5] = {0., 0., 0., 0., 0.}, dsx[5] = {0., 0., 0., 0., 0.};
    ExplPart e = {};
    int i0 = 0, j0 = 0;
    if (active) {
        const uint64_t idcpu = raw.idcpu;
        const double psi = raw.psi, xp = raw.xp, yp = raw.yp, ux = raw.ux, uy = raw.uy, w = raw.w;
        active = hpb_is_valid(idcpu);
        if (active) {
            const double psi_inv = 1.0 / psi;
            vx = ux * psi_inv * clight_inv;
            vy = uy * psi_inv * clight_inv;
            const double cdm = charge_invvol_mu0 * w;
            if (LASER) {        // ExplicitDeposition.cpp:167-182
                double A, ADx, ADy;
                laser_gather<false>(a, las.c_aabs, x_off, y_off, dx_inv, dy_inv, xp, yp, A, ADx, ADy);
                gamma_psi = 0.5 * ((1.0 + 0.5 * (A * las.a_norm)) * psi_inv * psi_inv + vx * vx + vy * vy + 1.0);
                qp = 0.25 * q_mass_ratio * psi_inv;
            } else {
                gamma_psi = 0.5 * (psi_inv * psi_inv + vx * vx + vy * vy + 1.0);
            }
            i0 = dshape2_centered((xp - x_off) * dx_inv, sx, dsx);
            const double ym = (yp - y_off) * dy_inv + 0.5;
            const double yfl = floor(ym);
            yint = ym - yfl;
            j0 = (int)yfl - 2;
            e = expl_part(vx, vy, gamma_psi, cdm, psi_inv, q_mass_ratio, a_clight);
        }
    }
    const int cc = active ? i0 + 2 : -(1 << 28) - 7 * lane;      // centre column / sentinel
    const double *Bz = a.comp(c_bz), *Ez = a.comp(c_ez);
    const double *ExmBy = a.comp(c_exmby), *EypBx = a.comp(c_eypbx);
    double *Sy = a.comp(c_sy), *Sx = a.comp(c_sx);

    // The field-dependent part carries the plain shape factor shx*shy, which vanishes on the
    // outer ring of the 5x5 stencil (s[0] = s[4] = 0): only rows 1..3 of the owned column need
    // the fields, and only the sources one cell to the side contribute through them.
    double fBz[3], fEz[3], fEx[3], fEy[3];
    double fADx[3] = {0., 0., 0.}, fADy[3] = {0., 0., 0.};      // AabssqDx/Dy at the owned cells (:215-226)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        fBz[r] = fEz[r] = fEx[r] = fEy[r] = 0.;
        if (active && owner) {
            const long o = a.idx(cc, j0 + 1 + r);
            fBz[r] = Bz[o]; fEz[r] = Ez[o]; fEx[r] = ExmBy[o]; fEy[r] = EypBx[o];
            if (LASER) {
                const double *ab = a.comp(las.c_aabs);
                fADx[r] = (ab[o + 1] - ab[o - 1]) * 0.5 * dx_inv * las.fac_c;
                fADy[r] = (ab[o + a.jstride] - ab[o - a.jstride]) * 0.5 * dy_inv * las.fac_c;
            }
        }
    }
    double accy[5] = {0., 0., 0., 0., 0.}, accx[5] = {0., 0., 0., 0., 0.};
    unsigned absorbed = 0;      // bit ix set: my column ix is deposited by the lane that owns it
#pragma unroll
    for (int d = -2; d <= 2; ++d) {
        // source = lane + d; its column 2 - d is my centre column
        const int src = lane + d;
        const int srcl = src < 0 ? 0 : (src > 31 ? 31 : src);
        const ExplCol mine = expl_col(e, sx[2 - d], dsx[2 - d], dx_inv, dy_inv);
        if (d == -2 || d == 2) {
            // outer column of the source (shx = 0): only the shdx*shy terms survive, rows 1..3
            const int c_s = __shfl_sync(kFull, cc, srcl), j_s = __shfl_sync(kFull, j0, srcl);
            const double B = __shfl_sync(kFull, mine.B, srcl), Bp = __shfl_sync(kFull, mine.Bp, srcl);
            const double yint_s = __shfl_sync(kFull, yint, srcl);
            const bool ok = active && owner && src >= 0 && src <= 31 && c_s == cc + d && j_s == j0;
            if (ok) {
                absorbed |= 1u << (2 + d);
                double shy[5], shdy[5];
                dshape2_centered_frac(yint_s, shy, shdy);
#pragma unroll
                for (int r = 1; r <= 3; ++r) { accy[r] += shy[r] * B; accx[r] += shy[r] * Bp; }
            }
        } else {
            ExplCol k;
            double vx_s, vy_s, gp_s, yint_s, qp_s = qp;
            bool ok;
            if (d == 0) {
                k = mine; vx_s = vx; vy_s = vy; gp_s = gamma_psi; yint_s = yint; ok = active && owner;
            } else {
                if (LASER) qp_s = __shfl_sync(kFull, qp, srcl);
                const int c_s = __shfl_sync(kFull, cc, srcl), j_s = __shfl_sync(kFull, j0, srcl);
                k.A = __shfl_sync(kFull, mine.A, srcl);   k.B = __shfl_sync(kFull, mine.B, srcl);
                k.C = __shfl_sync(kFull, mine.C, srcl);   k.Ap = -k.A;
                k.Bp = __shfl_sync(kFull, mine.Bp, srcl); k.Cp = __shfl_sync(kFull, mine.Cp, srcl);
                vx_s = __shfl_sync(kFull, vx, srcl);      vy_s = __shfl_sync(kFull, vy, srcl);
                gp_s = __shfl_sync(kFull, gamma_psi, srcl);
                yint_s = __shfl_sync(kFull, yint, srcl);
                ok = active && owner && src >= 0 && src <= 31 && c_s == cc + d && j_s == j0;
                if (ok) absorbed |= 1u << (2 + d);
            }
            if (ok) {
                const double a3 = -vx_s * vy_s, a4 = gp_s - vy_s * vy_s, a5 = gp_s - vx_s * vx_s;
                double shy[5], shdy[5];
                dshape2_centered_frac(yint_s, shy, shdy);
                accy[0] += shdy[0] * k.C; accx[0] += shdy[0] * k.Cp;      // shy[0] = shy[4] = 0
                accy[4] += shdy[4] * k.C; accx[4] += shdy[4] * k.Cp;
#pragma unroll
                for (int r = 1; r <= 3; ++r)
                    expl_cell(k, vx_s, vy_s, a3, a4, a5, shy[r], shdy[r], fBz[r - 1], fEz[r - 1],
                              fEx[r - 1], fEy[r - 1], clight_inv, accy[r], accx[r],
                              LASER ? fADy[r - 1] * qp_s : 0., LASER ? fADx[r - 1] * qp_s : 0.);
            }
        }
    }
    if (!active || !owner) return;
#pragma unroll
    for (int iy = 0; iy < 5; ++iy) {
        const long o = a.idx(cc, j0 + iy);
        red_add(Sy + o, accy[iy]);
        red_add(Sx + o, accx[iy]);
    }
    if (absorbed == 0x1bu) return;       // columns 0, 1, 3, 4 all owned by aligned neighbours
    // scatter the columns nobody owns (same expressions, fields read at the target cells)
    double shy[5], shdy[5];
    dshape2_centered_frac(yint, shy, shdy);
#pragma unroll 1
    for (int ix = 0; ix < 5; ++ix) {
        if (ix == 2 || ((absorbed >> ix) & 1u)) continue;
        const double wx = ix == 0 ? sx[0] : ix == 1 ? sx[1] : ix == 3 ? sx[3] : sx[4];
        const double wdx = ix == 0 ? dsx[0] : ix == 1 ? dsx[1] : ix == 3 ? dsx[3] : dsx[4];
        const ExplCol k = expl_col(e, wx, wdx, dx_inv, dy_inv);
#pragma unroll
        for (int iy = 0; iy < 5; ++iy) {
            if ((ix == 0 || ix == 4) && (iy == 0 || iy == 4)) continue;
            const long o = a.idx(i0 + ix, j0 + iy);
            double vy_ = 0., vx_ = 0.;
            double lasy = 0., lasx = 0.;
            if (LASER && wx * shy[iy] != 0.) {      // "avoid going outside of domain", :217
                const double *ab = a.comp(las.c_aabs);
                lasx = (ab[o + 1] - ab[o - 1]) * 0.5 * dx_inv * las.fac_c * qp;
                lasy = (ab[o + a.jstride] - ab[o - a.jstride]) * 0.5 * dy_inv * las.fac_c * qp;
            }
            expl_cell(k, e.vx, e.vy, e.a3, e.a4, e.a5, shy[iy], shdy[iy], Bz[o], Ez[o], ExmBy[o],
                      EypBx[o], clight_inv, vy_, vx_, lasy, lasx);
            red_add(Sy + o, vy_);
            red_add(Sx + o, vx_);
        }
    }
}

template <int NTHR, int MINB, bool LAT, bool LASER>
__global__ void __launch_bounds__(NTHR, MINB)
k_explicit_deposition(PlasmaPtrs pl, SliceView a, int c_sy, int c_sx, int c_bz, int c_ez,
                      int c_exmby, int c_eypbx, double x_off, double y_off, double dx_inv,
                      double dy_inv, double a_clight, double clight_inv,
                      double charge_invvol_mu0, double q_mass_ratio, ExplLaser las)
{
    hpb_pdl_prologue();
    // warps overlap by two lanes on each side (see k_deposit_current)
    const int lane = threadIdx.x & 31;
    const long warp = (long)blockIdx.x * (NTHR / 32) + (threadIdx.x >> 5);
    bool active;
    const long ip = lattice_particle<LAT ? 1 : 0>(pl, warp, lane, kExplOwn, 2, active);
    const bool owner = lane >= 2 && lane < 2 + kExplOwn;
    ExplRaw raw = {0, 1., 0., 0., 0., 0., 0.};
    if (active) {
        raw.idcpu = ld_stream(&pl.idcpu[ip]);
        raw.psi = ld_stream(&pl.r[HPB_PSI][ip]);
        raw.xp = ld_stream(&pl.r[HPB_X][ip]);
        raw.yp = ld_stream(&pl.r[HPB_Y][ip]);
        raw.ux = ld_stream(&pl.r[HPB_UX][ip]);
        raw.uy = ld_stream(&pl.r[HPB_UY][ip]);
        raw.w = ld_stream(&pl.r[HPB_W][ip]);
    }
    expl_group<LASER>(a, c_sy, c_sx, c_bz, c_ez, c_exmby, c_eypbx, x_off, y_off, dx_inv, dy_inv, a_clight, clight_inv,
                      charge_invvol_mu0, q_mass_ratio, las, lane, active, owner, raw);
}

// Persistent, software-pipelined variant: a fixed grid of warps walks over the particle groups; the seven
// stream values of the NEXT group are requested before the current group is processed, so the DRAM latency
// of the particle streams (the round-1 kernel's dominant stall: long scoreboard 3.4 per issue at 26 %
// occupancy) is hidden behind ~1000 instructions of work instead of being waited for by every warp.
template <int NTHR, int MINB, bool LAT>
__global__ void __launch_bounds__(NTHR, MINB)
k_explicit_deposition_pipe(PlasmaPtrs pl, SliceView a, int c_sy, int c_sx, int c_bz, int c_ez,
                           int c_exmby, int c_eypbx, double x_off, double y_off, double dx_inv,
                           double dy_inv, double a_clight, double clight_inv,
                           double charge_invvol_mu0, double q_mass_ratio, long nwarps)
{
    hpb_pdl_prologue();
    const int lane = threadIdx.x & 31;
    const long stride = (long)gridDim.x * (NTHR / 32);
    const bool owner = lane >= 2 && lane < 2 + kExplOwn;
    const ExplLaser las = {-1, 0., 0.};
    auto fetch = [&](long warp, ExplRaw &raw) -> bool {
        bool in_range;
        const long ip = lattice_particle<LAT ? 1 : 0>(pl, warp, lane, kExplOwn, 2, in_range);
        raw.idcpu = 0; raw.psi = 1.; raw.xp = raw.yp = raw.ux = raw.uy = raw.w = 0.;
        if (in_range) {
            raw.idcpu = ld_stream(&pl.idcpu[ip]);
            raw.psi = ld_stream(&pl.r[HPB_PSI][ip]);
            raw.xp = ld_stream(&pl.r[HPB_X][ip]);
            raw.yp = ld_stream(&pl.r[HPB_Y][ip]);
            raw.ux = ld_stream(&pl.r[HPB_UX][ip]);
            raw.uy = ld_stream(&pl.r[HPB_UY][ip]);
            raw.w = ld_stream(&pl.r[HPB_W][ip]);
        }
        return in_range;
    };
    long warp = (long)blockIdx.x * (NTHR / 32) + (threadIdx.x >> 5);
    if (warp >= nwarps) return;
    ExplRaw cur, nxt;
    bool cur_in = fetch(warp, cur), nxt_in = false;
    for (; warp < nwarps; warp += stride) {
        const bool more = warp + stride < nwarps;
        if (more) nxt_in = fetch(warp + stride, nxt);
        expl_group<false>(a, c_sy, c_sx, c_bz, c_ez, c_exmby, c_eypbx, x_off, y_off, dx_inv, dy_inv, a_clight,
                          clight_inv, charge_invvol_mu0, q_mass_ratio, las, lane, cur_in, owner, cur);
        cur = nxt; cur_in = nxt_in;
    }
}

// -------------------------------------------------------------------------------------------
// gather + push
// -------------------------------------------------------------------------------------------
// EnforceBC, GetAndSetPosition.H:56-98.  Returns true if the particle became invalid.
__device__ __forceinline__ bool enforce_bc(double &x, double &y, double &ux, double &uy, int bc,
                                           double lox, double loy, double hix, double hiy)
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

// ---- field gather -----------------------------------------------------------------------------
// doGatherShapeN<2> (FieldGather.H:45-96) evaluated row by row: for each of the 4 stencil rows the
// x-sums  sum_ix dSx Psi,  sum_ix Sx Psi,  sum_ix Sx {Ez,Bx,By,Bz}  are formed first and then
// weighted with Sy / dSy -- 144 fused multiply-adds instead of the ~210 operations of the
// cell-by-cell form (same sum up to fp64 re-association).
// Ld(f, ix, iy) returns field f in {Psi, Ez, Bx, By, Bz} at stencil cell (ix, iy).
template <class Ld>
__device__ __forceinline__ PushFields gather_rows(const Ld &ld, const double sx[4], const double dsx[4],
                                                  const double sy[4], const double dsy[4],
                                                  double dx_inv, double dy_inv)
{
    PushFields f = {0., 0., 0., 0., 0., 0.};
#pragma unroll
    for (int iy = 0; iy < 4; ++iy) {
        double a = 0., b = 0., e = 0., p = 0., q = 0., r = 0.;
#pragma unroll
        for (int ix = 0; ix < 4; ++ix) {
            const double psi_v = ld(0, ix, iy);
            a += dsx[ix] * psi_v;
            b += sx[ix] * psi_v;
            e += sx[ix] * ld(1, ix, iy);
            p += sx[ix] * ld(2, ix, iy);
            q += sx[ix] * ld(3, ix, iy);
            r += sx[ix] * ld(4, ix, iy);
        }
        f.ExmBy += sy[iy] * a;
        f.EypBx += dsy[iy] * b;
        f.Ez += sy[iy] * e;
        f.Bx_c += sy[iy] * p;
        f.By_c += sy[iy] * q;
        f.Bz += sy[iy] * r;
    }
    f.ExmBy *= dx_inv;
    f.EypBx *= dy_inv;
    return f;
}

__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// DEPOSIT: our fusion of AdvancePlasmaParticles (this slice) with ::DepositCurrent of the NEXT
// slice (PlasmaDepositCurrent.cpp:155-246): the pushed x, y, ux, uy, psi are deposited straight
// from registers into jx, jy, chi, rhomjz (which the caller has already shifted / initialised for
// the next slice), saving the 56 B/particle re-read and one launch per slice.
struct DepositArgs {
    int c_jx, c_jy, c_chi, c_rhomjz;
    double clightinv, charge_invvol, charge_mu0_mass_ratio, max_qsa;
    int *n_qsa_violation;
};

// STAGE: warp-tile staging of the gathered fields.  The 32 particles of a warp are lattice
// neighbours for most of the box (InitParticles order, x fastest), so their 4x4 gather stencils
// cover a patch of ~35 x 4 cells.  Each warp finds the bounding box of its stencils
// (warp min / max reductions); if it fits a kTW x kTH tile, the five field patches are fetched
// with asynchronous copies (LDGSTS, no registers, all in flight at once: ONE L2 round trip
// instead of ~8 register-limited batches of dependent 8-byte loads) into the warp's private
// shared-memory tile and gathered from there.  Warps whose particles have dispersed (sheath,
// trajectory crossing) fall back to direct loads, so the result never depends on the order.
constexpr int kPushThreads = 128;
constexpr int kTW = 40, kTH = 6;

struct PushLaserArgs { int c_aabs; double norm; };      // (charge/q_e)^2 (m_e/mass)^2, PlasmaParticleAdvance.cpp:76-77

template <int MINB, bool DEPOSIT, bool STAGE, int LAT, bool LASER>
__global__ void __launch_bounds__(kPushThreads, MINB)
k_advance_plasma(PlasmaPtrs pl, SliceView a, int c_psi, int c_ez, int c_bx, int c_by, int c_bz,
                 double x_off, double y_off, double dx_inv, double dy_inv, double clight,
                 double qmc, double dz, int n_subcycles, int temp_slice, int bc, double lox,
                 double loy, double hix, double hiy, DepositArgs dep, PushLaserArgs lasa)
{
    __shared__ double s_tile[STAGE ? kPushThreads / 32 : 1][5][kTH][kTW];
    hpb_pdl_prologue();
    const int lane = threadIdx.x & 31;
    bool in_range;
    const long ip = lattice_particle<LAT, kPushThreads / 32>(pl, (long)blockIdx.x * (kPushThreads / 32) + (threadIdx.x >> 5),
                                                             lane, 32, 0, in_range);
    // request every input stream before the first use (one HBM round trip)
    uint64_t idcpu = 0;
    double xp0 = 0., yp0 = 0., ux0 = 0., uy0 = 0., psi0 = 1., wq = 0.;
    if (in_range) {
        idcpu = ld_stream(&pl.idcpu[ip]);
        xp0 = ld_stream(&pl.r[HPB_X_PREV][ip]);
        yp0 = ld_stream(&pl.r[HPB_Y_PREV][ip]);
        ux0 = ld_stream(&pl.r[HPB_UX_HALF][ip]);
        uy0 = ld_stream(&pl.r[HPB_UY_HALF][ip]);
        psi0 = ld_stream(&pl.r[HPB_PSI_HALF][ip]);
        if (DEPOSIT) wq = ld_stream(&pl.r[HPB_W][ip]);
    }
    bool valid = in_range && hpb_is_valid(idcpu);
    const double clight_inv = 1.0 / clight;
    const double *F0 = a.comp(c_psi), *F1 = a.comp(c_ez), *F2 = a.comp(c_bx);
    const double *F3 = a.comp(c_by), *F4 = a.comp(c_bz);
    double xp = xp0, yp = yp0, ux = ux0, uy = uy0, psi = psi0;
    double (*tile)[kTH][kTW] = s_tile[STAGE ? (threadIdx.x >> 5) : 0];

    // every lane runs the loop (warp collectives inside); `valid` predicates the particle work
    for (int isc = 0; isc < n_subcycles; ++isc) {
        xp = xp0; yp = yp0;
        double sx[4], dsx[4], sy[4], dsy[4];
        const int i0 = dshape2_nodal((xp - x_off) * dx_inv, sx, dsx);
        const int j0 = dshape2_nodal((yp - y_off) * dy_inv, sy, dsy);
        bool staged = false;
        int imin = 0, jmin = 0;
        if (STAGE) {
            const int big = 1 << 30;
            imin = __reduce_min_sync(kFull, valid ? i0 : big);
            jmin = __reduce_min_sync(kFull, valid ? j0 : big);
            const int imax = __reduce_max_sync(kFull, valid ? i0 : -big);
            const int jmax = __reduce_max_sync(kFull, valid ? j0 : -big);
            const int ncol = imax - imin + 4, nrow = jmax - jmin + 4;
            staged = imax >= imin && ncol <= kTW && nrow <= kTH;
            if (staged) {
                const long o0 = a.idx(imin, jmin);
#pragma unroll 1
                for (int r = 0; r < nrow; ++r) {
                    const long orow = o0 + (long)r * a.jstride;
                    for (int c = lane; c < ncol; c += 32) {
                        cp_async_8(&tile[0][r][c], F0 + orow + c);
                        cp_async_8(&tile[1][r][c], F1 + orow + c);
                        cp_async_8(&tile[2][r][c], F2 + orow + c);
                        cp_async_8(&tile[3][r][c], F3 + orow + c);
                        cp_async_8(&tile[4][r][c], F4 + orow + c);
                    }
                }
                cp_async_wait_all();
                __syncwarp();
            }
        }
        if (valid) {
            PushFields f;
            if (STAGE && staged) {
                const double *t0 = &tile[0][j0 - jmin][i0 - imin];
                f = gather_rows([&](int fi, int ix, int iy) { return t0[fi * (kTH * kTW) + iy * kTW + ix]; },
                                sx, dsx, sy, dsy, dx_inv, dy_inv);
            } else {
                const long o = a.idx(i0, j0);
                const long js = a.jstride;
                f = gather_rows([&](int fi, int ix, int iy) {
                        const double *F = fi == 0 ? F0 : fi == 1 ? F1 : fi == 2 ? F2 : fi == 3 ? F3 : F4;
                        return F[o + iy * js + ix];
                    }, sx, dsx, sy, dsy, dx_inv, dy_inv);
            }
            f.Bx_c *= clight;
            f.By_c *= clight;
            PushLaser las = {0., 0., 0.};
            if (LASER) {        // PlasmaParticleAdvance.cpp:123-133
                laser_gather<true>(a, lasa.c_aabs, x_off, y_off, dx_inv, dy_inv, xp, yp, las.A, las.ADx, las.ADy);
                las.A *= 0.5 * lasa.norm;
                las.ADx *= 0.25 * clight * lasa.norm;
                las.ADy *= 0.25 * clight * lasa.norm;
            }

            constexpr int nsub = 4;
            const double sdz = dz / nsub;
            ux = ux0; uy = uy0; psi = psi0;
#pragma unroll 1
            for (int isub = 0; isub < nsub; ++isub) push_substep<LASER>(ux, uy, psi, f, clight_inv, qmc, sdz, las);

            xp += dz * clight_inv * (ux * (1.0 / psi));
            yp += dz * clight_inv * (uy * (1.0 / psi));
            if (enforce_bc(xp, yp, ux, uy, bc, lox, loy, hix, hiy)) {
                st_stream(&pl.r[HPB_W][ip], 0.0);
                st_stream(&pl.idcpu[ip], hpb_make_invalid(idcpu));
                valid = false;
            } else {
                st_stream(&pl.r[HPB_X][ip], xp);
                st_stream(&pl.r[HPB_Y][ip], yp);
                if (!temp_slice) {
                    st_stream(&pl.r[HPB_UX_HALF][ip], ux);
                    st_stream(&pl.r[HPB_UY_HALF][ip], uy);
                    st_stream(&pl.r[HPB_PSI_HALF][ip], psi);
                    st_stream(&pl.r[HPB_X_PREV][ip], xp);
                    st_stream(&pl.r[HPB_Y_PREV][ip], yp);
                    xp0 = xp; yp0 = yp; ux0 = ux; uy0 = uy; psi0 = psi;
                }
#pragma unroll 1
                for (int isub = 0; isub < nsub / 2; ++isub) push_substep<LASER>(ux, uy, psi, f, clight_inv, qmc, sdz, las);
                st_stream(&pl.r[HPB_UX][ip], ux);
                st_stream(&pl.r[HPB_UY][ip], uy);
                st_stream(&pl.r[HPB_PSI][ip], psi);
            }
        }
        if (STAGE && isc + 1 < n_subcycles) __syncwarp();     // the tile is rewritten next round
    }
    if (!DEPOSIT) return;

    // ::DepositCurrent of the pushed particle (same expressions as k_deposit_current)
    bool active = valid;
    const double psi_inv = 1.0 / psi;
    const double vx_c = ux * psi_inv, vy_c = uy * psi_inv;
    double q_invvol = dep.charge_invvol * wq;
    const double gamma_psi = 0.5 * (psi_inv * psi_inv + vx_c * vx_c * dep.clightinv * dep.clightinv
                                    + vy_c * vy_c * dep.clightinv * dep.clightinv + 1.0);
    if (active && (gamma_psi < 0.0 || gamma_psi > dep.max_qsa || psi_inv < 0.0)) {
        if (dep.n_qsa_violation) atomicAdd(dep.n_qsa_violation, 1);
        st_stream(&pl.r[HPB_W][ip], 0.0);
        st_stream(&pl.idcpu[ip], hpb_make_invalid(idcpu));
        active = false;
    }
    double dsx3[3] = {0., 0., 0.}, dsy3[3] = {0., 0., 0.};
    int di0 = 0, dj0 = 0;
    if (active) {
        di0 = shape2((xp - x_off) * dx_inv, dsx3);
        dj0 = shape2((yp - y_off) * dy_inv, dsy3);
    } else {
        q_invvol = 0.;
    }
    deposit_aggregated<true, false, true, true>(a, dep.c_jx, dep.c_jy, -1, dep.c_chi, dep.c_rhomjz,
                                                active, true, lane, di0, dj0, dsx3, dsy3, q_invvol,
                                                vx_c, vy_c, gamma_psi,
                                                dep.charge_mu0_mass_ratio * psi_inv);
}


// -------------------------------------------------------------------------------------------
// CTA-tile gather + push + deposit (the default for lattice-ordered plasmas with ppc 4 or 9)
// -------------------------------------------------------------------------------------------
// The NW = ppc warps of a CTA work on the NW passes of the SAME 32 lattice cells (InitParticles
// order: pass p of cell c is particle p * lat_n + c), so until the plasma has been stirred all
// 32 * NW particles of a CTA sit in one ~35 x 5 cell patch.  Three things follow:
//   * the five gathered field patches are staged ONCE per CTA -- by five TMA box loads
//     (cp.async.bulk.tensor.3d, one elected thread, completion on an mbarrier) instead of ~50
//     LDGSTS per warp; that per-warp staging loop was 27 % of all executed instructions of the
//     round-1 kernel (ncu source page) in an issue-bound kernel;
//   * every slice plane is fetched from DRAM once per slice instead of once per ppc pass
//     (round 1: 1.46 x the algorithmic DRAM traffic);
//   * the deposits of the NW passes go to the same cells: after the warp-shuffle aggregation of
//     neighbouring lanes the NW warps combine their centre columns through shared memory (plain
//     stores and loads, no shared-memory atomics) and issue 12 / NW fp64 reductions per lane
//     instead of 12 -- the same-address RED contention that made the pass-interleaved map lose
//     in round 1 is gone.
// Lanes whose particle has left the patch (sheath, trajectory crossing) gather with direct loads
// and deposit with their own reductions: the result never depends on the particle order.
constexpr int kCW = 40, kCH = 6;
// Array coordinate (>= 0) of a patch that should start at cell `want`: clamped so that the whole box
// lies inside the array, and rounded down to an even column (a 16-byte aligned row start).  Boxes that
// hang over the upper edge of the array in y made the TMA copy never complete on the B200
// (gpurun_out/r02c_debug*.txt: the mbarrier wait timed out once particles near the wall had moved
// up by a cell); a clamped box still holds every cell a stencil inside the array can touch.
__device__ __forceinline__ int patch_origin(int want, int lo, int tot, int ext, bool even)
{
    int a = want - lo;
    a = a > tot - ext ? tot - ext : a;
    a = a < 0 ? 0 : a;
    if (even) a &= ~1;
    return a;
}
constexpr uint32_t kCtaTileBytes = 5 * kCH * kCW * sizeof(double);
constexpr int kNoCell = -(1 << 28);

template <int NW>
struct CtaShared {
    alignas(128) double tile[5][kCH][kCW];   // Psi, Ez, Bx, By, Bz patch (TMA destination)
    double comb[NW][12][32];                 // per warp: {jx, jy, chi, rhomjz} x 3 rows of the centre column
    int ref_cc[32], ref_j0[32];              // warp 0's deposit cells: the cells the CTA combines on
    int wbox[NW][4];                         // per warp: min i0, min j0 (gather stencil origins)
    alignas(8) uint64_t mbar;
};

// Neighbour-lane aggregation of deposit_aggregated, but the centre column is RETURNED (v[comp * 3 +
// row], comp = jx, jy, chi, rhomjz) instead of reduced; side columns without an aligned neighbour
// lane are reduced here by the particle itself.  Must be called by all 32 lanes.
__device__ __forceinline__ void
deposit_centre_column(const SliceView &a, int c_jx, int c_jy, int c_chi, int c_rhomjz, bool active, int lane,
                      int i0, int j0, const double sx[3], const double sy[3], double q_invvol, double vx_c,
                      double vy_c, double chi_fac, double v[12])
{
    const int cc = active ? i0 + 1 : kNoCell - 3 * lane;
    const int cL = __shfl_up_sync(kFull, cc, 1), jL = __shfl_up_sync(kFull, j0, 1);
    const int cR = __shfl_down_sync(kFull, cc, 1), jR = __shfl_down_sync(kFull, j0, 1);
    const bool L_ok = active && lane > 0 && cL == cc - 1 && jL == j0;
    const bool R_ok = active && lane < 31 && cR == cc + 1 && jR == j0;
    const double vxL = __shfl_up_sync(kFull, vx_c, 1), vxR = __shfl_down_sync(kFull, vx_c, 1);
    const double vyL = __shfl_up_sync(kFull, vy_c, 1), vyR = __shfl_down_sync(kFull, vy_c, 1);
    const double cfL = __shfl_up_sync(kFull, chi_fac, 1), cfR = __shfl_down_sync(kFull, chi_fac, 1);
    double P0[3], P2[3];
#pragma unroll
    for (int iy = 0; iy < 3; ++iy) {
        P0[iy] = q_invvol * sx[0] * sy[iy];
        const double P1 = q_invvol * sx[1] * sy[iy];
        P2[iy] = q_invvol * sx[2] * sy[iy];
        double PL = __shfl_up_sync(kFull, P2[iy], 1);        // right column of the left neighbour
        double PR = __shfl_down_sync(kFull, P0[iy], 1);      // left column of the right neighbour
        if (!L_ok) PL = 0.;
        if (!R_ok) PR = 0.;
        v[0 + iy] = P1 * vx_c + PL * vxL + PR * vxR;
        v[3 + iy] = P1 * vy_c + PL * vyL + PR * vyR;
        v[6 + iy] = P1 * chi_fac + PL * cfL + PR * cfR;
        v[9 + iy] = P1 + PL + PR;
    }
    if (!active || (L_ok && R_ok)) return;
    double *jx = a.comp(c_jx), *jy = a.comp(c_jy), *chi = a.comp(c_chi), *rmj = a.comp(c_rhomjz);
#pragma unroll
    for (int iy = 0; iy < 3; ++iy) {
        const long o = a.idx(cc, j0 + iy);
        if (!L_ok) {
            red_add(jx + o - 1, P0[iy] * vx_c); red_add(jy + o - 1, P0[iy] * vy_c);
            red_add(chi + o - 1, P0[iy] * chi_fac); red_add(rmj + o - 1, P0[iy]);
        }
        if (!R_ok) {
            red_add(jx + o + 1, P2[iy] * vx_c); red_add(jy + o + 1, P2[iy] * vy_c);
            red_add(chi + o + 1, P2[iy] * chi_fac); red_add(rmj + o + 1, P2[iy]);
        }
    }
}

template <int NW, int MINB, bool DEPOSIT, bool TMA>
__global__ void __launch_bounds__(NW * 32, MINB)
k_advance_plasma_cta(PlasmaPtrs pl, SliceView a, const __grid_constant__ CUtensorMap tmap, int nx_tot,
                     int ny_tot, int lat_nx, int c_psi, int c_ez, int c_bx, int c_by, int c_bz, double x_off,
                     double y_off, double dx_inv, double dy_inv, double clight, double qmc, double dz,
                     int n_subcycles, int temp_slice, int bc, double lox, double loy, double hix, double hiy,
                     DepositArgs dep)
{
    __shared__ CtaShared<NW> sh;
    hpb_pdl_prologue();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // the 32 lattice cells of this CTA: a run inside one lattice row when the row length is known
    long cell;
    bool in_range;
    if (lat_nx > 0) {
        const unsigned gpr = (unsigned)(lat_nx + 31) >> 5, g = blockIdx.x;
        const unsigned row = g / gpr, col = (g - row * gpr) * 32 + lane;
        cell = (long)row * lat_nx + col;
        in_range = col < (unsigned)lat_nx && cell < pl.lat_n;
    } else {
        cell = (long)blockIdx.x * 32 + lane;
        in_range = cell < pl.lat_n;
    }
    const long ip = (long)w * pl.lat_n + cell;
    uint64_t idcpu = 0;
    double xp0 = 0., yp0 = 0., ux0 = 0., uy0 = 0., psi0 = 1., wq = 0.;
    if (in_range) {
        idcpu = ld_stream(&pl.idcpu[ip]);
        xp0 = ld_stream(&pl.r[HPB_X_PREV][ip]);
        yp0 = ld_stream(&pl.r[HPB_Y_PREV][ip]);
        ux0 = ld_stream(&pl.r[HPB_UX_HALF][ip]);
        uy0 = ld_stream(&pl.r[HPB_UY_HALF][ip]);
        psi0 = ld_stream(&pl.r[HPB_PSI_HALF][ip]);
        if (DEPOSIT) wq = ld_stream(&pl.r[HPB_W][ip]);
    }
    // Prefetch: the lattice says where the CTA's particles started, and most of the plasma has moved
    // less than a cell or two from there -- the patch around the home cells is requested NOW, in the
    // shadow of the particle loads, instead of after them (one DRAM/L2 round trip instead of two).
    int bi = 0, bj = 0;
    const bool prefetched = TMA && lat_nx > 0;
    if (TMA && tid == 0) {
        hpb_tma_prefetch_desc(&tmap);
        hpb_mbar_init(&sh.mbar, 1);
    }
    if (prefetched) {
        const unsigned gpr = (unsigned)(lat_nx + 31) >> 5, g = blockIdx.x;
        const unsigned row = g / gpr;
        // stencil origins col - 2 .. col - 1, 4 wide: cols col - 2 .. col + 2, rows row - 2 .. row + 2
        const int ax = patch_origin((int)((g - row * gpr) * 32) - 4, a.lo_x, nx_tot, kCW, true);
        const int ay = patch_origin((int)row - 3, a.lo_y, ny_tot, kCH, false);
        bi = ax + a.lo_x; bj = ay + a.lo_y;
        if (tid == 0) {
            hpb_mbar_arrive_expect_tx(&sh.mbar, kCtaTileBytes);
            hpb_tma_load_3d(&sh.tile[0][0][0], &tmap, &sh.mbar, ax, ay, c_psi);
            hpb_tma_load_3d(&sh.tile[1][0][0], &tmap, &sh.mbar, ax, ay, c_ez);
            hpb_tma_load_3d(&sh.tile[2][0][0], &tmap, &sh.mbar, ax, ay, c_bx);
            hpb_tma_load_3d(&sh.tile[3][0][0], &tmap, &sh.mbar, ax, ay, c_by);
            hpb_tma_load_3d(&sh.tile[4][0][0], &tmap, &sh.mbar, ax, ay, c_bz);
        }
    }
    bool valid = in_range && hpb_is_valid(idcpu);
    const double clight_inv = 1.0 / clight;
    const double *F0 = a.comp(c_psi), *F1 = a.comp(c_ez), *F2 = a.comp(c_bx);
    const double *F3 = a.comp(c_by), *F4 = a.comp(c_bz);
    double xp = xp0, yp = yp0, ux = ux0, uy = uy0, psi = psi0;
    uint32_t phase = 0;

    for (int isc = 0; isc < n_subcycles; ++isc) {
        xp = xp0; yp = yp0;
        double sx[4], dsx[4], sy[4], dsy[4];
        const int i0 = dshape2_nodal((xp - x_off) * dx_inv, sx, dsx);
        const int j0 = dshape2_nodal((yp - y_off) * dy_inv, sy, dsy);
        // bounding box of the CTA's stencil origins
        const int big = 1 << 30;
        const int imin = __reduce_min_sync(kFull, valid ? i0 : big), imax = __reduce_max_sync(kFull, valid ? i0 : -big);
        const int jmin = __reduce_min_sync(kFull, valid ? j0 : big), jmax = __reduce_max_sync(kFull, valid ? j0 : -big);
        if (lane == 0) { sh.wbox[w][0] = imin; sh.wbox[w][1] = jmin; sh.wbox[w][2] = imax; sh.wbox[w][3] = jmax; }
        __syncthreads();        // (also: the old tile has been consumed)
        int ci = big, cj = big, ei = -big, ej = -big;
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            ci = min(ci, sh.wbox[q][0]); cj = min(cj, sh.wbox[q][1]);
            ei = max(ei, sh.wbox[q][2]); ej = max(ej, sh.wbox[q][3]);
        }
        const bool any = ci != big;
        // the prefetched patch serves if every stencil of the CTA lies inside it; otherwise the patch
        // is (re)staged at the bounding box origin, and particles that do not fit even that (a stirred
        // group) gather with direct loads
        const bool pre_ok = prefetched && isc == 0 && any && ci >= bi && ei + 4 <= bi + kCW && cj >= bj
                            && ej + 4 <= bj + kCH;
        if (prefetched && isc == 0 && !pre_ok) {   // drain the prefetch before its buffer is rewritten
            hpb_mbar_wait(&sh.mbar, phase); phase ^= 1u;
            __syncthreads();
        }
        if (any && !pre_ok) {
            const int ax = patch_origin(ci, a.lo_x, nx_tot, kCW, TMA);      // array coordinates of the patch origin
            const int ay = patch_origin(cj, a.lo_y, ny_tot, kCH, false);
            bi = ax + a.lo_x; bj = ay + a.lo_y;
            if (TMA) {
                if (tid == 0) {
                    hpb_fence_proxy_async();
                    hpb_mbar_arrive_expect_tx(&sh.mbar, kCtaTileBytes);
                    hpb_tma_load_3d(&sh.tile[0][0][0], &tmap, &sh.mbar, ax, ay, c_psi);
                    hpb_tma_load_3d(&sh.tile[1][0][0], &tmap, &sh.mbar, ax, ay, c_ez);
                    hpb_tma_load_3d(&sh.tile[2][0][0], &tmap, &sh.mbar, ax, ay, c_bx);
                    hpb_tma_load_3d(&sh.tile[3][0][0], &tmap, &sh.mbar, ax, ay, c_by);
                    hpb_tma_load_3d(&sh.tile[4][0][0], &tmap, &sh.mbar, ax, ay, c_bz);
                }
            } else {
                // the same patch with asynchronous 8-byte copies by all threads (arrays the TMA unit
                // cannot address: odd row length); cells outside the array are never read
#pragma unroll 1
                for (int e = tid; e < 5 * kCH * kCW; e += NW * 32) {
                    const int f = e / (kCH * kCW), rem = e - f * (kCH * kCW), r = rem / kCW, c = rem - r * kCW;
                    const double *F = f == 0 ? F0 : f == 1 ? F1 : f == 2 ? F2 : f == 3 ? F3 : F4;
                    if (ax + c >= 0 && ay + r >= 0 && ax + c < nx_tot && ay + r < ny_tot)
                        cp_async_8(&sh.tile[f][r][c], F + (long)(ay + r) * a.jstride + (ax + c));
                }
                cp_async_wait_all();
                __syncthreads();
            }
        }
        const bool staged = any && i0 >= bi && i0 + 4 <= bi + kCW && j0 >= bj && j0 + 4 <= bj + kCH;
        if (TMA && any) { hpb_mbar_wait(&sh.mbar, phase); phase ^= 1u; }
        if (valid) {
            PushFields f;
            if (staged) {
                const double *t0 = &sh.tile[0][j0 - bj][i0 - bi];
                f = gather_rows([&](int fi, int ix, int iy) { return t0[fi * (kCH * kCW) + iy * kCW + ix]; },
                                sx, dsx, sy, dsy, dx_inv, dy_inv);
            } else {
                const long o = a.idx(i0, j0);
                const long js = a.jstride;
                f = gather_rows([&](int fi, int ix, int iy) {
                        const double *F = fi == 0 ? F0 : fi == 1 ? F1 : fi == 2 ? F2 : fi == 3 ? F3 : F4;
                        return F[o + iy * js + ix];
                    }, sx, dsx, sy, dsy, dx_inv, dy_inv);
            }
            f.Bx_c *= clight;
            f.By_c *= clight;
            const PushLaser las = {0., 0., 0.};
            constexpr int nsub = 4;
            const double sdz = dz / nsub;
            ux = ux0; uy = uy0; psi = psi0;
#pragma unroll 1
            for (int isub = 0; isub < nsub; ++isub) push_substep<false>(ux, uy, psi, f, clight_inv, qmc, sdz, las);

            xp += dz * clight_inv * (ux * (1.0 / psi));
            yp += dz * clight_inv * (uy * (1.0 / psi));
            if (enforce_bc(xp, yp, ux, uy, bc, lox, loy, hix, hiy)) {
                st_stream(&pl.r[HPB_W][ip], 0.0);
                st_stream(&pl.idcpu[ip], hpb_make_invalid(idcpu));
                valid = false;
            } else {
                st_stream(&pl.r[HPB_X][ip], xp);
                st_stream(&pl.r[HPB_Y][ip], yp);
                if (!temp_slice) {
                    st_stream(&pl.r[HPB_UX_HALF][ip], ux);
                    st_stream(&pl.r[HPB_UY_HALF][ip], uy);
                    st_stream(&pl.r[HPB_PSI_HALF][ip], psi);
                    st_stream(&pl.r[HPB_X_PREV][ip], xp);
                    st_stream(&pl.r[HPB_Y_PREV][ip], yp);
                    xp0 = xp; yp0 = yp; ux0 = ux; uy0 = uy; psi0 = psi;
                }
#pragma unroll 1
                for (int isub = 0; isub < nsub / 2; ++isub) push_substep<false>(ux, uy, psi, f, clight_inv, qmc, sdz, las);
                st_stream(&pl.r[HPB_UX][ip], ux);
                st_stream(&pl.r[HPB_UY][ip], uy);
                st_stream(&pl.r[HPB_PSI][ip], psi);
            }
        }
    }
    if (!DEPOSIT) return;

    // ::DepositCurrent of the pushed particle (same expressions as k_deposit_current)
    bool active = valid;
    const double psi_inv = 1.0 / psi;
    const double vx_c = ux * psi_inv, vy_c = uy * psi_inv;
    double q_invvol = dep.charge_invvol * wq;
    const double gamma_psi = 0.5 * (psi_inv * psi_inv + vx_c * vx_c * dep.clightinv * dep.clightinv
                                    + vy_c * vy_c * dep.clightinv * dep.clightinv + 1.0);
    if (active && (gamma_psi < 0.0 || gamma_psi > dep.max_qsa || psi_inv < 0.0)) {
        if (dep.n_qsa_violation) atomicAdd(dep.n_qsa_violation, 1);
        st_stream(&pl.r[HPB_W][ip], 0.0);
        st_stream(&pl.idcpu[ip], hpb_make_invalid(idcpu));
        active = false;
    }
    double dsx3[3] = {0., 0., 0.}, dsy3[3] = {0., 0., 0.};
    int di0 = 0, dj0 = 0;
    if (active) {
        di0 = shape2((xp - x_off) * dx_inv, dsx3);
        dj0 = shape2((yp - y_off) * dy_inv, dsy3);
    } else {
        q_invvol = 0.;
    }
    double v[12];
    deposit_centre_column(a, dep.c_jx, dep.c_jy, dep.c_chi, dep.c_rhomjz, active, lane, di0, dj0, dsx3, dsy3,
                          q_invvol, vx_c, vy_c, dep.charge_mu0_mass_ratio * psi_inv, v);
    // combine the centre columns of the NW passes on warp 0's cells
    const int cc = active ? di0 + 1 : kNoCell - 3 * lane;
    if (w == 0) { sh.ref_cc[lane] = cc; sh.ref_j0[lane] = dj0; }
    __syncthreads();
    const int rcc = sh.ref_cc[lane], rj0 = sh.ref_j0[lane];
    const bool match = active && cc == rcc && dj0 == rj0;
#pragma unroll
    for (int k = 0; k < 12; ++k) sh.comb[w][k][lane] = match ? v[k] : 0.;
    if (active && !match) {     // this pass has left warp 0's cell: its own reductions
#pragma unroll
        for (int iy = 0; iy < 3; ++iy) {
            const long o = a.idx(cc, dj0 + iy);
            red_add(a.comp(dep.c_jx) + o, v[0 + iy]);
            red_add(a.comp(dep.c_jy) + o, v[3 + iy]);
            red_add(a.comp(dep.c_chi) + o, v[6 + iy]);
            red_add(a.comp(dep.c_rhomjz) + o, v[9 + iy]);
        }
    }
    __syncthreads();
    if (rcc <= kNoCell) return;     // warp 0's lane holds no particle: nobody matched it
#pragma unroll
    for (int kk = 0; kk < (12 + NW - 1) / NW; ++kk) {
        const int k = w + kk * NW;
        if (k >= 12) break;
        double sum = 0.;
#pragma unroll
        for (int q = 0; q < NW; ++q) sum += sh.comb[q][k][lane];
        const int comp = k < 3 ? dep.c_jx : k < 6 ? dep.c_jy : k < 9 ? dep.c_chi : dep.c_rhomjz;
        const int row = k - (k / 3) * 3;
        red_add(a.comp(comp) + a.idx(rcc, rj0 + row), sum);
    }
}

// -------------------------------------------------------------------------------------------
// Row-tile gather + push + deposit: the round-1 thread map (128 consecutive particles per CTA, in
// lattice order a run of 128 cells of one row and one ppc pass) with the per-warp LDGSTS staging
// loop -- 27 % of all executed instructions of that kernel -- replaced by ONE 136 x 6 TMA patch per
// field and CTA, requested before the particle loads return.  No shared-memory combine is needed: the
// warps of a CTA deposit into different cells, the aggregation of neighbouring lanes is the warp-shuffle
// scheme of deposit_aggregated.  Works for any particle order (lanes outside the patch use direct
// loads); the lattice hint only supplies the prefetch position.
constexpr int kRowThreads = 128;
constexpr int kRW = 136;
constexpr uint32_t kRowTileBytes = 5 * kCH * kRW * sizeof(double);
struct RowShared {
    alignas(128) double tile[5][kCH][kRW];
    int wbox[kRowThreads / 32][4];
    alignas(8) uint64_t mbar;
};

template <int MINB, bool DEPOSIT, bool TMA, bool NOBOX = false>
__global__ void __launch_bounds__(kRowThreads, MINB)
k_advance_plasma_row(PlasmaPtrs pl, SliceView a, const __grid_constant__ CUtensorMap tmap, int nx_tot,
                     int ny_tot, int lat_nx, int c_psi, int c_ez, int c_bx, int c_by, int c_bz, double x_off,
                     double y_off, double dx_inv, double dy_inv, double clight, double qmc, double dz,
                     int n_subcycles, int temp_slice, int bc, double lox, double loy, double hix, double hiy,
                     DepositArgs dep)
{
    constexpr int NW = kRowThreads / 32;
    __shared__ RowShared sh;
    hpb_pdl_prologue();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // kRowThreads consecutive particles: in lattice order a run of cells of one row and one pass
    const long ip = (long)blockIdx.x * kRowThreads + tid;
    const bool in_range = ip < pl.np;
    uint64_t idcpu = 0;
    double xp0 = 0., yp0 = 0., ux0 = 0., uy0 = 0., psi0 = 1., wq = 0.;
    if (in_range) {
        idcpu = ld_stream(&pl.idcpu[ip]);
        xp0 = ld_stream(&pl.r[HPB_X_PREV][ip]);
        yp0 = ld_stream(&pl.r[HPB_Y_PREV][ip]);
        ux0 = ld_stream(&pl.r[HPB_UX_HALF][ip]);
        uy0 = ld_stream(&pl.r[HPB_UY_HALF][ip]);
        psi0 = ld_stream(&pl.r[HPB_PSI_HALF][ip]);
        if (DEPOSIT) wq = ld_stream(&pl.r[HPB_W][ip]);
    }
    // Prefetch: the lattice says where the CTA's particles started, and most of the plasma has moved
    // less than a cell or two from there -- the patch around the home cells is requested NOW, in the
    // shadow of the particle loads, instead of after them (one DRAM/L2 round trip instead of two).
    int bi = 0, bj = 0;
    // home cell of the CTA's first particle (lattice order: particle = pass * lat_n + cell, x fastest)
    const long cell0 = pl.lat_n > 0 ? ((long)blockIdx.x * kRowThreads) % pl.lat_n : 0;
    const int hrow = lat_nx > 0 ? (int)(cell0 / lat_nx) : 0, hcol = lat_nx > 0 ? (int)(cell0 - (long)hrow * lat_nx) : 0;
    // (a run that crosses the end of a lattice row does not fit one patch: no prefetch for it)
    const bool prefetched = TMA && lat_nx > 0 && hcol + kRowThreads <= lat_nx;
    if (TMA && tid == 0) {
        hpb_tma_prefetch_desc(&tmap);
        hpb_mbar_init(&sh.mbar, 1);
    }
    if (prefetched) {
        const int ax = patch_origin(hcol - 4, a.lo_x, nx_tot, kRW, true);
        const int ay = patch_origin(hrow - 3, a.lo_y, ny_tot, kCH, false);
        bi = ax + a.lo_x; bj = ay + a.lo_y;
        if (tid == 0) {
            hpb_mbar_arrive_expect_tx(&sh.mbar, kRowTileBytes);
            hpb_tma_load_3d(&sh.tile[0][0][0], &tmap, &sh.mbar, ax, ay, c_psi);
            hpb_tma_load_3d(&sh.tile[1][0][0], &tmap, &sh.mbar, ax, ay, c_ez);
            hpb_tma_load_3d(&sh.tile[2][0][0], &tmap, &sh.mbar, ax, ay, c_bx);
            hpb_tma_load_3d(&sh.tile[3][0][0], &tmap, &sh.mbar, ax, ay, c_by);
            hpb_tma_load_3d(&sh.tile[4][0][0], &tmap, &sh.mbar, ax, ay, c_bz);
        }
    }
    const bool nobox = NOBOX && prefetched && n_subcycles == 1;
    if (NOBOX) __syncthreads();      // the initialised barrier is visible to every warp (they are still together)
    bool valid = in_range && hpb_is_valid(idcpu);
    const double clight_inv = 1.0 / clight;
    const double *F0 = a.comp(c_psi), *F1 = a.comp(c_ez), *F2 = a.comp(c_bx);
    const double *F3 = a.comp(c_by), *F4 = a.comp(c_bz);
    double xp = xp0, yp = yp0, ux = ux0, uy = uy0, psi = psi0;
    uint32_t phase = 0;

    for (int isc = 0; isc < n_subcycles; ++isc) {
        xp = xp0; yp = yp0;
        double sx[4], dsx[4], sy[4], dsy[4];
        const int i0 = dshape2_nodal((xp - x_off) * dx_inv, sx, dsx);
        const int j0 = dshape2_nodal((yp - y_off) * dy_inv, sy, dsy);
        // NOBOX: no bounding box, no restaging, no block barrier: every lane tests its own stencil against
        // the prefetched patch and gathers with direct loads if it lies outside
        const int big = 1 << 30;
        int ci = big, cj = big, ei = -big, ej = -big;
        if (!nobox) {
            // bounding box of the CTA's stencil origins
            const int imin = __reduce_min_sync(kFull, valid ? i0 : big), imax = __reduce_max_sync(kFull, valid ? i0 : -big);
            const int jmin = __reduce_min_sync(kFull, valid ? j0 : big), jmax = __reduce_max_sync(kFull, valid ? j0 : -big);
            if (lane == 0) { sh.wbox[w][0] = imin; sh.wbox[w][1] = jmin; sh.wbox[w][2] = imax; sh.wbox[w][3] = jmax; }
            __syncthreads();        // (also: the old tile has been consumed)
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                ci = min(ci, sh.wbox[q][0]); cj = min(cj, sh.wbox[q][1]);
                ei = max(ei, sh.wbox[q][2]); ej = max(ej, sh.wbox[q][3]);
            }
        }
        const bool any = nobox || ci != big;
        // the prefetched patch serves if every stencil of the CTA lies inside it; otherwise the patch
        // is (re)staged at the bounding box origin, and particles that do not fit even that (a stirred
        // group) gather with direct loads
        const bool pre_ok = nobox || (prefetched && isc == 0 && any && ci >= bi && ei + 4 <= bi + kRW && cj >= bj
                                      && ej + 4 <= bj + kCH);
        if (prefetched && isc == 0 && !pre_ok) {   // drain the prefetch before its buffer is rewritten
            hpb_mbar_wait(&sh.mbar, phase); phase ^= 1u;
            __syncthreads();
        }
        if (any && !pre_ok) {
            const int ax = patch_origin(ci, a.lo_x, nx_tot, kRW, TMA);      // array coordinates of the patch origin
            const int ay = patch_origin(cj, a.lo_y, ny_tot, kCH, false);
            bi = ax + a.lo_x; bj = ay + a.lo_y;
            if (TMA) {
                if (tid == 0) {
                    hpb_fence_proxy_async();
                    hpb_mbar_arrive_expect_tx(&sh.mbar, kRowTileBytes);
                    hpb_tma_load_3d(&sh.tile[0][0][0], &tmap, &sh.mbar, ax, ay, c_psi);
                    hpb_tma_load_3d(&sh.tile[1][0][0], &tmap, &sh.mbar, ax, ay, c_ez);
                    hpb_tma_load_3d(&sh.tile[2][0][0], &tmap, &sh.mbar, ax, ay, c_bx);
                    hpb_tma_load_3d(&sh.tile[3][0][0], &tmap, &sh.mbar, ax, ay, c_by);
                    hpb_tma_load_3d(&sh.tile[4][0][0], &tmap, &sh.mbar, ax, ay, c_bz);
                }
            } else {
                // the same patch with asynchronous 8-byte copies by all threads (arrays the TMA unit
                // cannot address: odd row length); cells outside the array are never read
#pragma unroll 1
                for (int e = tid; e < 5 * kCH * kRW; e += kRowThreads) {
                    const int f = e / (kCH * kRW), rem = e - f * (kCH * kRW), r = rem / kRW, c = rem - r * kRW;
                    const double *F = f == 0 ? F0 : f == 1 ? F1 : f == 2 ? F2 : f == 3 ? F3 : F4;
                    if (ax + c >= 0 && ay + r >= 0 && ax + c < nx_tot && ay + r < ny_tot)
                        cp_async_8(&sh.tile[f][r][c], F + (long)(ay + r) * a.jstride + (ax + c));
                }
                cp_async_wait_all();
                __syncthreads();
            }
        }
        const bool staged = any && i0 >= bi && i0 + 4 <= bi + kRW && j0 >= bj && j0 + 4 <= bj + kCH;
        if (TMA && any) { hpb_mbar_wait(&sh.mbar, phase); phase ^= 1u; }
        if (valid) {
            PushFields f;
            if (staged) {
                const double *t0 = &sh.tile[0][j0 - bj][i0 - bi];
                f = gather_rows([&](int fi, int ix, int iy) { return t0[fi * (kCH * kRW) + iy * kRW + ix]; },
                                sx, dsx, sy, dsy, dx_inv, dy_inv);
            } else {
                const long o = a.idx(i0, j0);
                const long js = a.jstride;
                f = gather_rows([&](int fi, int ix, int iy) {
                        const double *F = fi == 0 ? F0 : fi == 1 ? F1 : fi == 2 ? F2 : fi == 3 ? F3 : F4;
                        return F[o + iy * js + ix];
                    }, sx, dsx, sy, dsy, dx_inv, dy_inv);
            }
            f.Bx_c *= clight;
            f.By_c *= clight;
            const PushLaser las = {0., 0., 0.};
            constexpr int nsub = 4;
            const double sdz = dz / nsub;
            ux = ux0; uy = uy0; psi = psi0;
#pragma unroll 1
            for (int isub = 0; isub < nsub; ++isub) push_substep<false>(ux, uy, psi, f, clight_inv, qmc, sdz, las);

            xp += dz * clight_inv * (ux * (1.0 / psi));
            yp += dz * clight_inv * (uy * (1.0 / psi));
            if (enforce_bc(xp, yp, ux, uy, bc, lox, loy, hix, hiy)) {
                st_stream(&pl.r[HPB_W][ip], 0.0);
                st_stream(&pl.idcpu[ip], hpb_make_invalid(idcpu));
                valid = false;
            } else {
                st_stream(&pl.r[HPB_X][ip], xp);
                st_stream(&pl.r[HPB_Y][ip], yp);
                if (!temp_slice) {
                    st_stream(&pl.r[HPB_UX_HALF][ip], ux);
                    st_stream(&pl.r[HPB_UY_HALF][ip], uy);
                    st_stream(&pl.r[HPB_PSI_HALF][ip], psi);
                    st_stream(&pl.r[HPB_X_PREV][ip], xp);
                    st_stream(&pl.r[HPB_Y_PREV][ip], yp);
                    xp0 = xp; yp0 = yp; ux0 = ux; uy0 = uy; psi0 = psi;
                }
#pragma unroll 1
                for (int isub = 0; isub < nsub / 2; ++isub) push_substep<false>(ux, uy, psi, f, clight_inv, qmc, sdz, las);
                st_stream(&pl.r[HPB_UX][ip], ux);
                st_stream(&pl.r[HPB_UY][ip], uy);
                st_stream(&pl.r[HPB_PSI][ip], psi);
            }
        }
    }
    if (!DEPOSIT) return;

    // ::DepositCurrent of the pushed particle (same expressions as k_deposit_current)
    bool active = valid;
    const double psi_inv = 1.0 / psi;
    const double vx_c = ux * psi_inv, vy_c = uy * psi_inv;
    double q_invvol = dep.charge_invvol * wq;
    const double gamma_psi = 0.5 * (psi_inv * psi_inv + vx_c * vx_c * dep.clightinv * dep.clightinv
                                    + vy_c * vy_c * dep.clightinv * dep.clightinv + 1.0);
    if (active && (gamma_psi < 0.0 || gamma_psi > dep.max_qsa || psi_inv < 0.0)) {
        if (dep.n_qsa_violation) atomicAdd(dep.n_qsa_violation, 1);
        st_stream(&pl.r[HPB_W][ip], 0.0);
        st_stream(&pl.idcpu[ip], hpb_make_invalid(idcpu));
        active = false;
    }
    double dsx3[3] = {0., 0., 0.}, dsy3[3] = {0., 0., 0.};
    int di0 = 0, dj0 = 0;
    if (active) {
        di0 = shape2((xp - x_off) * dx_inv, dsx3);
        dj0 = shape2((yp - y_off) * dy_inv, dsy3);
    } else {
        q_invvol = 0.;
    }
    deposit_aggregated<true, false, true, true>(a, dep.c_jx, dep.c_jy, -1, dep.c_chi, dep.c_rhomjz,
                                                active, true, lane, di0, dj0, dsx3, dsy3, q_invvol,
                                                vx_c, vy_c, gamma_psi,
                                                dep.charge_mu0_mass_ratio * psi_inv);
}

// -------------------------------------------------------------------------------------------
// CTA-tile explicit deposition (lattice-ordered plasmas with ppc 4 or 9)
// -------------------------------------------------------------------------------------------
// Same idea as k_advance_plasma_cta: the NW warps of a CTA hold the NW passes of the same 28
// lattice cells (+ 2 feed-only lanes on each side, see k_explicit_deposition).  The four field
// planes the owned centre columns read (Bz, Ez, ExmBy, EypBx: 3 rows each) arrive as ONE TMA patch
// per CTA, requested before the particle loads return (the round-1 kernel was latency bound on the
// second, dependent round trip: long-scoreboard stalls 3.4 per issue at 26 % occupancy), and the 10
// accumulated values of every centre column are combined across the passes in shared memory before
// the fp64 reductions: 10 / NW REDs per lane instead of 10.
constexpr int kEW = 40, kEH = 6;
constexpr uint32_t kExplTileBytes = 4 * kEH * kEW * sizeof(double);

template <int NW>
struct ExplShared {
    alignas(128) double tile[4][kEH][kEW];   // Bz, Ez, ExmBy, EypBx
    double comb[NW][10][32];                 // per warp: Sy rows 0..4, Sx rows 0..4 of the centre column
    int ref_cc[32], ref_j0[32];
    int wbox[NW][4];
    alignas(8) uint64_t mbar;
};

template <int NW, int MINB, bool TMA, bool COMBINE, bool NOBOX = false>
__global__ void __launch_bounds__(NW * 32, MINB)
k_explicit_deposition_cta(PlasmaPtrs pl, SliceView a, const __grid_constant__ CUtensorMap tmap, int nx_tot,
                          int ny_tot, int lat_nx, int c_sy, int c_sx, int c_bz, int c_ez, int c_exmby,
                          int c_eypbx, double x_off, double y_off, double dx_inv, double dy_inv,
                          double a_clight, double clight_inv, double charge_invvol_mu0, double q_mass_ratio)
{
    __shared__ ExplShared<NW> sh;
    hpb_pdl_prologue();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // 28 owned lattice cells per CTA, lanes 0, 1, 30, 31 only feed their neighbours
    long cell;
    bool active;
    int col0 = 0, row = 0;
    if (lat_nx > 0) {
        const unsigned gpr = (unsigned)(lat_nx + kExplOwn - 1) / kExplOwn, g = blockIdx.x;
        row = (int)(g / gpr);
        col0 = (int)(g - (unsigned)row * gpr) * kExplOwn;
        const int col = col0 - 2 + lane;
        cell = (long)row * lat_nx + col;
        active = col >= 0 && col < lat_nx && cell < pl.lat_n;
    } else {
        cell = (long)blockIdx.x * kExplOwn - 2 + lane;
        active = cell >= 0 && cell < pl.lat_n;
    }
    const long ip = (long)w * pl.lat_n + cell;
    const bool owner = lane >= 2 && lane < 2 + kExplOwn;

    int bi = 0, bj = 0;
    const bool prefetched = TMA && lat_nx > 0;
    if (TMA && tid == 0) {
        hpb_tma_prefetch_desc(&tmap);
        hpb_mbar_init(&sh.mbar, 1);
    }
    if (prefetched) {
        // owned centre columns col0 .. col0 + 27 with 6 cells of slack on each side; rows row - 1 ..
        // row + 1 with two / one row(s) of slack
        const int ax = patch_origin(col0 - 6, a.lo_x, nx_tot, kEW, true);
        const int ay = patch_origin(row - 3, a.lo_y, ny_tot, kEH, false);
        bi = ax + a.lo_x; bj = ay + a.lo_y;
        if (tid == 0) {
            hpb_mbar_arrive_expect_tx(&sh.mbar, kExplTileBytes);
            hpb_tma_load_3d(&sh.tile[0][0][0], &tmap, &sh.mbar, ax, ay, c_bz);
            hpb_tma_load_3d(&sh.tile[1][0][0], &tmap, &sh.mbar, ax, ay, c_ez);
            hpb_tma_load_3d(&sh.tile[2][0][0], &tmap, &sh.mbar, ax, ay, c_exmby);
            hpb_tma_load_3d(&sh.tile[3][0][0], &tmap, &sh.mbar, ax, ay, c_eypbx);
        }
    }
    if (NOBOX) __syncthreads();      // the initialised barrier is visible to every warp (they are still together)

    double vx = 0., vy = 0., gamma_psi = 1., yint = 0.;
    double sx[5] = {0., 0., 0., 0., 0.}, dsx[5] = {0., 0., 0., 0., 0.};
    ExplPart e = {};
    int i0 = 0, j0 = 0;
    if (active) {
        const uint64_t idcpu = ld_stream(&pl.idcpu[ip]);
        const double psi = ld_stream(&pl.r[HPB_PSI][ip]);
        const double xp = ld_stream(&pl.r[HPB_X][ip]);
        const double yp = ld_stream(&pl.r[HPB_Y][ip]);
        const double ux = ld_stream(&pl.r[HPB_UX][ip]);
        const double uy = ld_stream(&pl.r[HPB_UY][ip]);
        const double wt = ld_stream(&pl.r[HPB_W][ip]);
        active = hpb_is_valid(idcpu);
        if (active) {
            const double psi_inv = 1.0 / psi;
            vx = ux * psi_inv * clight_inv;
            vy = uy * psi_inv * clight_inv;
            const double cdm = charge_invvol_mu0 * wt;
            gamma_psi = 0.5 * (psi_inv * psi_inv + vx * vx + vy * vy + 1.0);
            i0 = dshape2_centered((xp - x_off) * dx_inv, sx, dsx);
            const double ym = (yp - y_off) * dy_inv + 0.5;
            const double yfl = floor(ym);
            yint = ym - yfl;
            j0 = (int)yfl - 2;
            e = expl_part(vx, vy, gamma_psi, cdm, psi_inv, q_mass_ratio, a_clight);
        }
    }
    const int cc = active ? i0 + 2 : kNoCell - 7 * lane;      // centre column / sentinel
    const double *Bz = a.comp(c_bz), *Ez = a.comp(c_ez);
    const double *ExmBy = a.comp(c_exmby), *EypBx = a.comp(c_eypbx);
    double *Sy = a.comp(c_sy), *Sx = a.comp(c_sx);

    // NOBOX: no bounding box, no restaging, no block barrier -- every lane tests its own cells against the
    // prefetched patch and falls back to direct loads; the warps of the CTA never wait for each other
    const bool nobox = NOBOX && prefetched;
    bool any = true;
    if (nobox) {
        hpb_mbar_wait(&sh.mbar, 0);
    } else {
    // bounding box of the owned centre cells (rows j0 + 1 .. j0 + 3) of the whole CTA
    {
        const int big = 1 << 30;
        const bool own = active && owner;
        const int imin = __reduce_min_sync(kFull, own ? cc : big), imax = __reduce_max_sync(kFull, own ? cc : -big);
        const int jmin = __reduce_min_sync(kFull, own ? j0 + 1 : big), jmax = __reduce_max_sync(kFull, own ? j0 + 3 : -big);
        if (lane == 0) { sh.wbox[w][0] = imin; sh.wbox[w][1] = jmin; sh.wbox[w][2] = imax; sh.wbox[w][3] = jmax; }
    }
    __syncthreads();
    {
        const int big = 1 << 30;
        int ci = big, cj = big, ei = -big, ej = -big;
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            ci = min(ci, sh.wbox[q][0]); cj = min(cj, sh.wbox[q][1]);
            ei = max(ei, sh.wbox[q][2]); ej = max(ej, sh.wbox[q][3]);
        }
        any = ci != big;
        const bool pre_ok = prefetched && any && ci >= bi && ei < bi + kEW && cj >= bj && ej < bj + kEH;
        if (prefetched && !pre_ok) {       // drain the prefetch before its buffer is rewritten
            hpb_mbar_wait(&sh.mbar, 0);
            __syncthreads();
        }
        if (any && !pre_ok) {
            const int ax = patch_origin(ci - 2, a.lo_x, nx_tot, kEW, TMA);
            const int ay = patch_origin(cj, a.lo_y, ny_tot, kEH, false);
            bi = ax + a.lo_x; bj = ay + a.lo_y;
            if (TMA) {
                if (tid == 0) {
                    hpb_fence_proxy_async();
                    hpb_mbar_arrive_expect_tx(&sh.mbar, kExplTileBytes);
                    hpb_tma_load_3d(&sh.tile[0][0][0], &tmap, &sh.mbar, ax, ay, c_bz);
                    hpb_tma_load_3d(&sh.tile[1][0][0], &tmap, &sh.mbar, ax, ay, c_ez);
                    hpb_tma_load_3d(&sh.tile[2][0][0], &tmap, &sh.mbar, ax, ay, c_exmby);
                    hpb_tma_load_3d(&sh.tile[3][0][0], &tmap, &sh.mbar, ax, ay, c_eypbx);
                }
            } else {
#pragma unroll 1
                for (int q = tid; q < 4 * kEH * kEW; q += NW * 32) {
                    const int f = q / (kEH * kEW), rem = q - f * (kEH * kEW), r = rem / kEW, c = rem - r * kEW;
                    const double *F = f == 0 ? Bz : f == 1 ? Ez : f == 2 ? ExmBy : EypBx;
                    if (ax + c >= 0 && ay + r >= 0 && ax + c < nx_tot && ay + r < ny_tot)
                        cp_async_8(&sh.tile[f][r][c], F + (long)(ay + r) * a.jstride + (ax + c));
                }
                cp_async_wait_all();
                __syncthreads();
            }
        }
        if (TMA && any) hpb_mbar_wait(&sh.mbar, (prefetched && !pre_ok) ? 1u : 0u);
    }
    }

    // fields at the three inner rows of the owned column: from the patch, or straight from the slice
    // if the particle has left it
    double fBz[3], fEz[3], fEx[3], fEy[3];
    const bool staged = any && active && owner && cc >= bi && cc < bi + kEW && j0 + 1 >= bj && j0 + 3 < bj + kEH;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        fBz[r] = fEz[r] = fEx[r] = fEy[r] = 0.;
        if (staged) {
            const double *t = &sh.tile[0][j0 + 1 + r - bj][cc - bi];
            fBz[r] = t[0]; fEz[r] = t[kEH * kEW]; fEx[r] = t[2 * kEH * kEW]; fEy[r] = t[3 * kEH * kEW];
        } else if (active && owner) {
            const long o = a.idx(cc, j0 + 1 + r);
            fBz[r] = Bz[o]; fEz[r] = Ez[o]; fEx[r] = ExmBy[o]; fEy[r] = EypBx[o];
        }
    }
    double accy[5] = {0., 0., 0., 0., 0.}, accx[5] = {0., 0., 0., 0., 0.};
    unsigned absorbed = 0;      // bit ix set: my column ix is deposited by the lane that owns it
#pragma unroll
    for (int d = -2; d <= 2; ++d) {
        const int src = lane + d;
        const int srcl = src < 0 ? 0 : (src > 31 ? 31 : src);
        const ExplCol mine = expl_col(e, sx[2 - d], dsx[2 - d], dx_inv, dy_inv);
        if (d == -2 || d == 2) {
            const int c_s = __shfl_sync(kFull, cc, srcl), j_s = __shfl_sync(kFull, j0, srcl);
            const double B = __shfl_sync(kFull, mine.B, srcl), Bp = __shfl_sync(kFull, mine.Bp, srcl);
            const double yint_s = __shfl_sync(kFull, yint, srcl);
            const bool ok = active && owner && src >= 0 && src <= 31 && c_s == cc + d && j_s == j0;
            if (ok) {
                absorbed |= 1u << (2 + d);
                double shy[5], shdy[5];
                dshape2_centered_frac(yint_s, shy, shdy);
#pragma unroll
                for (int r = 1; r <= 3; ++r) { accy[r] += shy[r] * B; accx[r] += shy[r] * Bp; }
            }
        } else {
            ExplCol k;
            double vx_s, vy_s, gp_s, yint_s;
            bool ok;
            if (d == 0) {
                k = mine; vx_s = vx; vy_s = vy; gp_s = gamma_psi; yint_s = yint; ok = active && owner;
            } else {
                const int c_s = __shfl_sync(kFull, cc, srcl), j_s = __shfl_sync(kFull, j0, srcl);
                k.A = __shfl_sync(kFull, mine.A, srcl);   k.B = __shfl_sync(kFull, mine.B, srcl);
                k.C = __shfl_sync(kFull, mine.C, srcl);   k.Ap = -k.A;
                k.Bp = __shfl_sync(kFull, mine.Bp, srcl); k.Cp = __shfl_sync(kFull, mine.Cp, srcl);
                vx_s = __shfl_sync(kFull, vx, srcl);      vy_s = __shfl_sync(kFull, vy, srcl);
                gp_s = __shfl_sync(kFull, gamma_psi, srcl);
                yint_s = __shfl_sync(kFull, yint, srcl);
                ok = active && owner && src >= 0 && src <= 31 && c_s == cc + d && j_s == j0;
                if (ok) absorbed |= 1u << (2 + d);
            }
            if (ok) {
                const double a3 = -vx_s * vy_s, a4 = gp_s - vy_s * vy_s, a5 = gp_s - vx_s * vx_s;
                double shy[5], shdy[5];
                dshape2_centered_frac(yint_s, shy, shdy);
                accy[0] += shdy[0] * k.C; accx[0] += shdy[0] * k.Cp;      // shy[0] = shy[4] = 0
                accy[4] += shdy[4] * k.C; accx[4] += shdy[4] * k.Cp;
#pragma unroll
                for (int r = 1; r <= 3; ++r)
                    expl_cell(k, vx_s, vy_s, a3, a4, a5, shy[r], shdy[r], fBz[r - 1], fEz[r - 1],
                              fEx[r - 1], fEy[r - 1], clight_inv, accy[r], accx[r]);
            }
        }
    }
    const bool own = active && owner;
    if (!COMBINE) {
        // every pass reduces its own centre column (the round-1 scheme; the patch prefetch is the gain)
        if (own) {
#pragma unroll
            for (int iy = 0; iy < 5; ++iy) {
                const long o = a.idx(cc, j0 + iy);
                red_add(Sy + o, accy[iy]);
                red_add(Sx + o, accx[iy]);
            }
        }
    }
    // combine the centre columns of the NW passes on warp 0's cells
    if (COMBINE && w == 0) { sh.ref_cc[lane] = own ? cc : kNoCell - 7 * lane; sh.ref_j0[lane] = j0; }
    if (COMBINE) __syncthreads();
    const int rcc = COMBINE ? sh.ref_cc[lane] : kNoCell, rj0 = COMBINE ? sh.ref_j0[lane] : 0;
    const bool match = own && cc == rcc && j0 == rj0;
    if (COMBINE) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            sh.comb[w][k][lane] = match ? accy[k] : 0.;
            sh.comb[w][5 + k][lane] = match ? accx[k] : 0.;
        }
        if (own && !match) {
#pragma unroll
            for (int iy = 0; iy < 5; ++iy) {
                const long o = a.idx(cc, j0 + iy);
                red_add(Sy + o, accy[iy]);
                red_add(Sx + o, accx[iy]);
            }
        }
        __syncthreads();
    }
    if (COMBINE && rcc > kNoCell) {
#pragma unroll
        for (int kk = 0; kk < (10 + NW - 1) / NW; ++kk) {
            const int k = w + kk * NW;
            if (k < 10) {
                double sum = 0.;
#pragma unroll
                for (int q = 0; q < NW; ++q) sum += sh.comb[q][k][lane];
                const int iy = k < 5 ? k : k - 5;
                red_add((k < 5 ? Sy : Sx) + a.idx(rcc, rj0 + iy), sum);
            }
        }
    }
    if (!own || absorbed == 0x1bu) return;       // columns 0, 1, 3, 4 all owned by aligned neighbours
    // scatter the columns nobody owns (same expressions, fields read at the target cells)
    double shy[5], shdy[5];
    dshape2_centered_frac(yint, shy, shdy);
#pragma unroll 1
    for (int ix = 0; ix < 5; ++ix) {
        if (ix == 2 || ((absorbed >> ix) & 1u)) continue;
        const double wx = ix == 0 ? sx[0] : ix == 1 ? sx[1] : ix == 3 ? sx[3] : sx[4];
        const double wdx = ix == 0 ? dsx[0] : ix == 1 ? dsx[1] : ix == 3 ? dsx[3] : dsx[4];
        const ExplCol k = expl_col(e, wx, wdx, dx_inv, dy_inv);
#pragma unroll
        for (int iy = 0; iy < 5; ++iy) {
            if ((ix == 0 || ix == 4) && (iy == 0 || iy == 4)) continue;
            const long o = a.idx(i0 + ix, j0 + iy);
            double vy_ = 0., vx_ = 0.;
            expl_cell(k, e.vx, e.vy, e.a3, e.a4, e.a5, shy[iy], shdy[iy], Bz[o], Ez[o], ExmBy[o],
                      EypBx[o], clight_inv, vy_, vx_);
            red_add(Sy + o, vy_);
            red_add(Sx + o, vx_);
        }
    }
}

inline unsigned nblocks(long n) { return (unsigned)((n + kThreads - 1) / kThreads); }

}  // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
static int deposit_current(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge, double mass,
                           int c_jx, int c_jy, int c_rho, int c_chi, int c_rhomjz, int c_aabs,
                           double max_qsa, int *d_n_qsa_violation)
{
    if (!ctx) return HPB_ERR_ARG;
    if (pl.np == 0) return HPB_OK;
    const hpb_geom &g = ctx->g;
    // invvol: 1 in normalised units at lev 0, 1/(dx dy dz) in SI (PlasmaDepositCurrent.cpp:71-73)
    const double invvol = g.normalized ? 1.0 : (1.0 / g.dx) * (1.0 / g.dy) * (1.0 / g.dz);
    if ((c_jx >= 0) != (c_jy >= 0)) { hpb_set_error("deposit: jx and jy go together"); return HPB_ERR_ARG; }
    if (hpb_use_generic_order(ctx))
        return hpb_gen_deposit_current(ctx, pl, sl, charge, mass, c_jx, c_jy, -1, c_rho, c_chi, c_rhomjz,
                                       c_aabs, max_qsa, d_n_qsa_violation);
    const int mask = (c_jx >= 0 ? 8 : 0) | (c_rho >= 0 ? 4 : 0) | (c_chi >= 0 ? 2 : 0) | (c_rhomjz >= 0 ? 1 : 0);
    // (charge/q_e)^2 (m_e/mass)^2, PlasmaDepositCurrent.cpp:80-81
    const double laser_norm = (charge / g.q_e) * (g.m_e / mass) * (charge / g.q_e) * (g.m_e / mass);
#define HPB_DEP_(M, LAS)                                                                          \
        hpb_launch(k_deposit_current<((M) & 8) != 0, ((M) & 4) != 0, ((M) & 2) != 0, ((M) & 1) != 0, LAS>, \
               (unsigned)((pl.np + kDepOwn * (kThreads / 32) - 1) / (kDepOwn * (kThreads / 32))), \
               kThreads, 0, ctx->stream,                                                        \
                to_ptrs(pl), make_view(sl), c_jx, c_jy, c_rho, c_chi, c_rhomjz, g.x_off, g.y_off, \
                1.0 / g.dx, 1.0 / g.dy, 1.0 / g.c, charge * invvol, charge * g.mu0 / mass,        \
                max_qsa, d_n_qsa_violation, c_aabs, laser_norm)
#define HPB_DEP(M)                                                                                \
    case M:                                                                                       \
        if (c_aabs >= 0) HPB_DEP_(M, true); else HPB_DEP_(M, false);                              \
        break;
    switch (mask) {
        HPB_DEP(1) HPB_DEP(2) HPB_DEP(3) HPB_DEP(4) HPB_DEP(5) HPB_DEP(6) HPB_DEP(7) HPB_DEP(8)
        HPB_DEP(9) HPB_DEP(10) HPB_DEP(11) HPB_DEP(12) HPB_DEP(13) HPB_DEP(14) HPB_DEP(15)
    default: return HPB_OK;      // nothing to deposit
    }
#undef HPB_DEP
#undef HPB_DEP_
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_deposit_current(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge,
                                   double mass, int c_jx, int c_jy, int c_rho, int c_chi,
                                   int c_rhomjz, double max_qsa, int *d_n_qsa_violation)
{
    return deposit_current(ctx, pl, sl, charge, mass, c_jx, c_jy, c_rho, c_chi, c_rhomjz, -1, max_qsa,
                           d_n_qsa_violation);
}

extern "C" int hpb_deposit_current_laser(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge,
                                         double mass, int c_jx, int c_jy, int c_rho, int c_chi,
                                         int c_rhomjz, int c_aabs, double max_qsa,
                                         int *d_n_qsa_violation)
{
    return deposit_current(ctx, pl, sl, charge, mass, c_jx, c_jy, c_rho, c_chi, c_rhomjz, c_aabs, max_qsa,
                           d_n_qsa_violation);
}

// ::DepositCurrent with every destination of the reference (PlasmaDepositCurrent.cpp:53-58), jz
// included: what the predictor-corrector solver deposits.  Always the one-thread-per-particle
// kernels of generic_order.cu (the warp-aggregated kernels above have no jz variant).
extern "C" int hpb_deposit_current_jz(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge,
                                      double mass, int c_jx, int c_jy, int c_jz, int c_rho, int c_chi,
                                      int c_rhomjz, int c_aabs, double max_qsa, int *d_n_qsa_violation)
{
    if (!ctx) return HPB_ERR_ARG;
    if (pl.np == 0) return HPB_OK;
    if ((c_jx >= 0) != (c_jy >= 0)) { hpb_set_error("deposit: jx and jy go together"); return HPB_ERR_ARG; }
    return hpb_gen_deposit_current(ctx, pl, sl, charge, mass, c_jx, c_jy, c_jz, c_rho, c_chi, c_rhomjz, c_aabs,
                                   max_qsa, d_n_qsa_violation);
}

// TEST HOOK (host only): the thread -> particle map of the push kernel for a lattice of
// cells_per_pass x ppc particles, mode 0 (linear), 1 (passes interleaved warp by warp) or 2 (CTA by
// CTA).  out[w * 32 + lane] = particle index or -1; returns the number of warps (tests check that
// every mode is a permutation of the particles).
extern "C" long hpb_debug_push_thread_map(long cells_per_pass, int ppc, int mode, long *out, long out_len)
{
    PlasmaPtrs p;
    for (int i = 0; i < HPB_PLASMA_NREAL; ++i) p.r[i] = nullptr;
    p.idcpu = nullptr; p.np = cells_per_pass * ppc;
    p.lat_n = mode ? cells_per_pass : 0; p.lat_ppc = ppc; p.lat_mode = mode;
    const long nwarps = lattice_warps(p, 32, kPushThreads / 32);
    if (!out) return nwarps;
    if (out_len < nwarps * 32) return -1;
    for (long w = 0; w < nwarps; ++w)
        for (int lane = 0; lane < 32; ++lane) {
            bool in_range = false;
            long ip;
            if (mode == 2) ip = lattice_particle<2, kPushThreads / 32>(p, w, lane, 32, 0, in_range);
            else if (mode == 1) ip = lattice_particle<1>(p, w, lane, 32, 0, in_range);
            else ip = lattice_particle<0>(p, w, lane, 32, 0, in_range);
            out[w * 32 + lane] = in_range ? ip : -1;
        }
    return nwarps;
}

extern "C" int hpb_set_plasma_lattice_hint(hpb_ctx *ctx, long cells_per_pass, int ppc)
{
    if (!ctx || cells_per_pass < 0 || ppc < 0) return HPB_ERR_ARG;
    ctx->order_n = cells_per_pass;
    ctx->order_ppc = ppc;
    return HPB_OK;
}

extern "C" int hpb_beam_deposit(hpb_ctx *ctx, hpb_beam_slice bm, hpb_slice sl, double charge,
                                int c_jx, int c_jy, int c_jz)
{
    if (!ctx) return HPB_ERR_ARG;
    if (bm.np == 0 || (c_jx < 0 && c_jz < 0)) return HPB_OK;
    if (hpb_use_generic_order(ctx)) return hpb_gen_beam_deposit(ctx, bm, sl, charge, c_jx, c_jy, c_jz);
    const hpb_geom &g = ctx->g;
    // BeamDepositCurrent.cpp:72-82: invvol = 1 in normalised units at lev 0
    const double invvol = g.normalized ? 1.0 : (1.0 / g.dx) * (1.0 / g.dy) * (1.0 / g.dz);
    BeamPtrs b{bm.x, bm.y, bm.z, bm.w, bm.ux, bm.uy, bm.uz, bm.idcpu, bm.np, bm.d_np};
    hpb_launch(k_beam_deposit, nblocks(bm.np), kThreads, 0, ctx->stream, 
        b, make_view(sl), c_jx, c_jy, c_jz, g.x_off, g.y_off, 1.0 / g.dx, 1.0 / g.dy,
        1.0 / (g.c * g.c), charge * invvol);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_explicit_deposition(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge,
                                       double mass, const int *comps)
{
    if (!ctx || !comps) return HPB_ERR_ARG;
    if (pl.np == 0) return HPB_OK;
    if (hpb_use_generic_order(ctx)) return hpb_gen_explicit_deposition(ctx, pl, sl, charge, mass, comps);
    const hpb_geom &g = ctx->g;
    const double invvol = g.normalized ? 1.0 : (1.0 / g.dx) * (1.0 / g.dy) * (1.0 / g.dz);
    const int variant = ctx->tune_expl_variant;
    // expl_variant 4 (default) / 7: the CTA-tile kernel (TMA / cp.async staging) whenever the particles
    // carry the lattice order with ppc 4 or 9 and there is no laser; 0: the round-1 warp kernel
    const bool lattice = ctx->order_n > 0 && ctx->order_ppc > 1 && ctx->order_n * ctx->order_ppc == pl.np;
    if ((variant == 4 || (variant >= 7 && variant <= 12)) && comps[HPB_C_AABS] < 0 && lattice
        && (ctx->order_ppc == 4 || ctx->order_ppc == 9)) {
        PlasmaPtrs pp = to_ptrs(pl);
        pp.lat_n = ctx->order_n; pp.lat_ppc = ctx->order_ppc; pp.lat_mode = 1;
        const int lat_nx = pp.lat_n == (long)g.nx * g.ny ? g.nx : 0;
        const long groups = lat_nx > 0 ? (long)g.ny * ((lat_nx + kExplOwn - 1) / kExplOwn)
                                       : (pp.lat_n + kExplOwn - 1) / kExplOwn + 1;
        const CUtensorMap *tm = variant != 7 ? (const CUtensorMap *)hpb_slice_tmap(ctx, 1, sl, kEW, kEH) : nullptr;
        CUtensorMap none;
        memset(&none, 0, sizeof(none));
#define HPB_LAUNCH_ECTA(NW, MB, TMA_, CMB, ...)                                                    \
        hpb_launch(k_explicit_deposition_cta<NW, MB, TMA_, CMB, ##__VA_ARGS__>, (unsigned)groups, NW * 32, 0, ctx->stream, \
                   pp, make_view(sl), TMA_ ? *tm : none, sl.nx_tot, sl.ny_tot, lat_nx, comps[HPB_C_SY],  \
                   comps[HPB_C_SX], comps[HPB_C_BZ], comps[HPB_C_EZ], comps[HPB_C_EXMBY], comps[HPB_C_EYPBX], \
                   g.x_off, g.y_off, 1.0 / g.dx, 1.0 / g.dy, g.c, 1.0 / g.c, charge * invvol * g.mu0, charge / mass)
        if (ctx->order_ppc == 4 && variant == 8 && tm) HPB_LAUNCH_ECTA(4, 4, true, true);      // 128 registers
        else if (ctx->order_ppc == 4 && variant == 9 && tm) HPB_LAUNCH_ECTA(4, 5, true, false);  // no combine
        else if (ctx->order_ppc == 4 && variant == 10 && tm) HPB_LAUNCH_ECTA(4, 4, true, false); // no combine, 128 regs
        else if (ctx->order_ppc == 4 && variant == 11 && tm) HPB_LAUNCH_ECTA(4, 5, true, false, true);   // + no bounding box
        else if (ctx->order_ppc == 4 && variant == 12 && tm) HPB_LAUNCH_ECTA(4, 4, true, false, true);   // + 128 regs
        else if (ctx->order_ppc == 4) { if (tm) HPB_LAUNCH_ECTA(4, 5, true, true); else HPB_LAUNCH_ECTA(4, 5, false, true); }
        else { if (tm) HPB_LAUNCH_ECTA(9, 2, true, true); else HPB_LAUNCH_ECTA(9, 2, false, true); }
#undef HPB_LAUNCH_ECTA
        hpb_count_launch(ctx);
        HPB_CUDA_CHECK(cudaGetLastError());
        return HPB_OK;
    }
    // (the laser variant uses the plain particle order)
    const PlasmaPtrs pp = comps[HPB_C_AABS] >= 0 ? to_ptrs(pl) : to_ptrs(ctx, pl, 1);
    const long nwarps = lattice_warps(pp, kExplOwn);
    // laser: ExplicitDeposition.cpp:55 laser_fac = (m_e / q_e)^2, a0 is always normalised
    ExplLaser las = {comps[HPB_C_AABS], 0., 0.};
    if (las.c_aabs >= 0) {
        const double laser_fac = (g.m_e / g.q_e) * (g.m_e / g.q_e);
        las.fac_c = laser_fac * g.c;
        las.a_norm = laser_fac * (charge / mass) * (charge / mass);
    }
#define HPB_LAUNCH_EXPL(NT, MB)                                                                    \
    do { if (las.c_aabs >= 0) HPB_LAUNCH_EXPL_(128, 3, false, true);                               \
         else if (pp.lat_n > 0) HPB_LAUNCH_EXPL_(NT, MB, true, false);                             \
         else HPB_LAUNCH_EXPL_(NT, MB, false, false); } while (0)
#define HPB_LAUNCH_EXPL_(NT, MB, LAT, LAS)                                                         \
    hpb_launch(k_explicit_deposition<NT, MB, LAT, LAS>, (unsigned)((nwarps + (NT / 32) - 1) / (NT / 32)), \
                                    NT, 0, ctx->stream,                                         \
        pp, make_view(sl), comps[HPB_C_SY], comps[HPB_C_SX], comps[HPB_C_BZ],             \
        comps[HPB_C_EZ], comps[HPB_C_EXMBY], comps[HPB_C_EYPBX], g.x_off, g.y_off, 1.0 / g.dx,     \
        1.0 / g.dy, g.c, 1.0 / g.c, charge * invvol * g.mu0, charge / mass, las)
    if ((variant == 13 || variant == 14 || variant == 15) && las.c_aabs < 0) {
        // persistent, software-pipelined kernel: 4 (variant 13, 128 registers), 5 (14) or 3 (15) CTAs per SM
        int dev = 0, nsm = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        const int per_sm = variant == 13 ? 4 : variant == 14 ? 5 : 3;
        long ncta = (nwarps + 3) / 4;
        if (ncta > (long)nsm * per_sm) ncta = (long)nsm * per_sm;
#define HPB_LAUNCH_PIPE(MB, LAT)                                                                   \
        hpb_launch(k_explicit_deposition_pipe<128, MB, LAT>, (unsigned)ncta, 128, 0, ctx->stream,  \
            pp, make_view(sl), comps[HPB_C_SY], comps[HPB_C_SX], comps[HPB_C_BZ], comps[HPB_C_EZ],   \
            comps[HPB_C_EXMBY], comps[HPB_C_EYPBX], g.x_off, g.y_off, 1.0 / g.dx, 1.0 / g.dy, g.c,   \
            1.0 / g.c, charge * invvol * g.mu0, charge / mass, nwarps)
        if (variant == 13) { if (pp.lat_n > 0) HPB_LAUNCH_PIPE(4, true); else HPB_LAUNCH_PIPE(4, false); }
        else if (variant == 14) { if (pp.lat_n > 0) HPB_LAUNCH_PIPE(5, true); else HPB_LAUNCH_PIPE(5, false); }
        else { if (pp.lat_n > 0) HPB_LAUNCH_PIPE(3, true); else HPB_LAUNCH_PIPE(3, false); }
#undef HPB_LAUNCH_PIPE
    } else
    if (variant == 2) HPB_LAUNCH_EXPL(256, 2);
    else if (variant == 3) HPB_LAUNCH_EXPL(128, 3);
    else if (variant == 5) HPB_LAUNCH_EXPL(128, 6);
    else if (variant == 6) HPB_LAUNCH_EXPL(128, 8);
    else if (variant == 1) HPB_LAUNCH_EXPL(256, 1);
    else if (variant == 16) HPB_LAUNCH_EXPL(128, 5);    // the default until r02H: 96 registers, 200 bytes of spills
    // 4 CTAs / SM, 124 registers, no spills: the kernel's L1TEX pipe (72 % of peak, r02F) also carried the
    // spill traffic -- 0.244 -> 0.228 ms, 929 -> 951 slices/s (profiles/r02/r02H_tune.txt)
    else HPB_LAUNCH_EXPL(128, 4);
#undef HPB_LAUNCH_EXPL
#undef HPB_LAUNCH_EXPL_
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

static int advance_plasma(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge, double mass,
                          int n_subcycles, int temp_slice, int particle_bc, const double bc_lo[2],
                          const double bc_hi[2], const int *comps, bool deposit, double max_qsa,
                          int *d_n_qsa_violation)
{
    if (!ctx || !comps || !bc_lo || !bc_hi || n_subcycles < 1) return HPB_ERR_ARG;
    if (pl.np == 0) return HPB_OK;
    if (hpb_use_generic_order(ctx)) {
        if (deposit) {
            hpb_set_error("advance+deposit: only for depos_order_xy = 2 with the centred derivative");
            return HPB_ERR_UNSUPPORTED;
        }
        return hpb_gen_advance_plasma(ctx, pl, sl, charge, mass, n_subcycles, temp_slice, particle_bc, bc_lo,
                                      bc_hi, comps);
    }
    const hpb_geom &g = ctx->g;
    const int variant = ctx->tune_push_variant;
    DepositArgs dep = {};
    if (deposit) {
        const double invvol = g.normalized ? 1.0 : (1.0 / g.dx) * (1.0 / g.dy) * (1.0 / g.dz);
        dep.c_jx = comps[HPB_C_JX]; dep.c_jy = comps[HPB_C_JY]; dep.c_chi = comps[HPB_C_CHI];
        dep.c_rhomjz = comps[HPB_C_RHOMJZ];
        dep.clightinv = 1.0 / g.c; dep.charge_invvol = charge * invvol;
        dep.charge_mu0_mass_ratio = charge * g.mu0 / mass;
        dep.max_qsa = max_qsa; dep.n_qsa_violation = d_n_qsa_violation;
    }
    PushLaserArgs lasa = {comps[HPB_C_AABS], 0.};
    if (lasa.c_aabs >= 0) {
        if (deposit) { hpb_set_error("advance+deposit: not available with a laser"); return HPB_ERR_UNSUPPORTED; }
        lasa.norm = (charge / g.q_e) * (g.m_e / mass) * (charge / g.q_e) * (g.m_e / mass);
    }
    // push_variant 0 (default) / 4: the CTA-tile kernel (TMA / cp.async staging) whenever the particles
    // carry the lattice order with ppc 4 or 9; 2: the round-1 warp-staged kernel, 1: direct loads,
    // 3: warp-staged with 128 registers
    const bool lattice = ctx->order_n > 0 && ctx->order_ppc > 1 && ctx->order_n * ctx->order_ppc == pl.np;
    // push_variant 6 / 7 / 8: the row-tile kernel (round-1 thread map, one TMA patch per CTA; 96 / 128
    // registers / cp.async staging); any particle order
    // (an array the TMA unit cannot address -- odd row length, e.g. the 1023 + 4 cells of a 2^n - 1 grid --
    // runs the round-1 warp-staged kernel below: the cooperative cp.async staging of variant 8 is the
    // slowest of the three, 0.55 vs 0.34 vs 0.29 ms at 1024^2 ppc 4)
    const bool row_tma_ok = variant == 8 || hpb_slice_tmap(ctx, 2, sl, kRW, kCH) != nullptr;
    if ((variant >= 6 && variant <= 10) && lasa.c_aabs < 0 && row_tma_ok) {
        PlasmaPtrs pp = to_ptrs(pl);
        const bool lat = ctx->order_n > 0 && ctx->order_n * ctx->order_ppc == pl.np;
        pp.lat_n = lat ? ctx->order_n : 0; pp.lat_ppc = lat ? ctx->order_ppc : 1;
        const int lat_nx = lat && pp.lat_n == (long)g.nx * g.ny ? g.nx : 0;
        const long nblk = (pl.np + kRowThreads - 1) / kRowThreads;
        const CUtensorMap *tm = variant != 8 ? (const CUtensorMap *)hpb_slice_tmap(ctx, 2, sl, kRW, kCH) : nullptr;
        CUtensorMap none;
        memset(&none, 0, sizeof(none));
#define HPB_LAUNCH_ROW(MB, DEP, TMA_, ...)                                                         \
        hpb_launch(k_advance_plasma_row<MB, DEP, TMA_, ##__VA_ARGS__>, (unsigned)nblk, kRowThreads, 0, ctx->stream, \
                   pp, make_view(sl), TMA_ ? *tm : none, sl.nx_tot, sl.ny_tot, lat_nx, comps[HPB_C_PSI], \
                   comps[HPB_C_EZ], comps[HPB_C_BX], comps[HPB_C_BY], comps[HPB_C_BZ], g.x_off, g.y_off,  \
                   1.0 / g.dx, 1.0 / g.dy, g.c, charge / (mass * g.c), g.dz / n_subcycles, n_subcycles,  \
                   temp_slice, particle_bc, bc_lo[0], bc_lo[1], bc_hi[0], bc_hi[1], dep)
#define HPB_LAUNCH_ROW_MB(MB)                                                                     \
        do { if (deposit) { if (tm) HPB_LAUNCH_ROW(MB, true, true); else HPB_LAUNCH_ROW(MB, true, false); } \
             else { if (tm) HPB_LAUNCH_ROW(MB, false, true); else HPB_LAUNCH_ROW(MB, false, false); } } while (0)
        if (variant == 9 && tm) { if (deposit) HPB_LAUNCH_ROW(5, true, true, true); else HPB_LAUNCH_ROW(5, false, true, true); }
        else if (variant == 7) HPB_LAUNCH_ROW_MB(4); else if (variant == 10) HPB_LAUNCH_ROW_MB(6); else HPB_LAUNCH_ROW_MB(5);
#undef HPB_LAUNCH_ROW_MB
#undef HPB_LAUNCH_ROW
        hpb_count_launch(ctx);
        HPB_CUDA_CHECK(cudaGetLastError());
        return HPB_OK;
    }
    if ((variant == 0 || variant == 4 || variant == 5) && lasa.c_aabs < 0 && lattice && (ctx->order_ppc == 4 || ctx->order_ppc == 9)) {
        PlasmaPtrs pp = to_ptrs(pl);
        pp.lat_n = ctx->order_n; pp.lat_ppc = ctx->order_ppc; pp.lat_mode = 1;
        // lattice row length: known when the lattice is the whole box (x fastest)
        const int lat_nx = pp.lat_n == (long)g.nx * g.ny ? g.nx : 0;
        const long groups = lat_nx > 0 ? (long)g.ny * ((lat_nx + 31) / 32) : (pp.lat_n + 31) / 32;
        const CUtensorMap *tm = variant != 4 ? (const CUtensorMap *)hpb_slice_tmap(ctx, 0, sl, kCW, kCH) : nullptr;
        CUtensorMap none;
        memset(&none, 0, sizeof(none));
#define HPB_LAUNCH_CTA(NW, MB, DEP, TMA_)                                                          \
        hpb_launch(k_advance_plasma_cta<NW, MB, DEP, TMA_>, (unsigned)groups, NW * 32, 0, ctx->stream, \
                   pp, make_view(sl), TMA_ ? *tm : none, sl.nx_tot, sl.ny_tot, lat_nx, comps[HPB_C_PSI], \
                   comps[HPB_C_EZ], comps[HPB_C_BX], comps[HPB_C_BY], comps[HPB_C_BZ], g.x_off, g.y_off,  \
                   1.0 / g.dx, 1.0 / g.dy, g.c, charge / (mass * g.c), g.dz / n_subcycles, n_subcycles,  \
                   temp_slice, particle_bc, bc_lo[0], bc_lo[1], bc_hi[0], bc_hi[1], dep)
#define HPB_LAUNCH_CTA_NW(NW, MB)                                                                 \
        do { if (deposit) { if (tm) HPB_LAUNCH_CTA(NW, MB, true, true); else HPB_LAUNCH_CTA(NW, MB, true, false); } \
             else { if (tm) HPB_LAUNCH_CTA(NW, MB, false, true); else HPB_LAUNCH_CTA(NW, MB, false, false); } } while (0)
        if (ctx->order_ppc == 4 && variant == 5) HPB_LAUNCH_CTA_NW(4, 4);       // 128 registers
        else if (ctx->order_ppc == 4) HPB_LAUNCH_CTA_NW(4, 5);
        else HPB_LAUNCH_CTA_NW(9, 2);
#undef HPB_LAUNCH_CTA_NW
#undef HPB_LAUNCH_CTA
        hpb_count_launch(ctx);
        HPB_CUDA_CHECK(cudaGetLastError());
        return HPB_OK;
    }
    const PlasmaPtrs pp = comps[HPB_C_AABS] >= 0 ? to_ptrs(pl) : to_ptrs(ctx, pl, 2);
    const long nwarps = lattice_warps(pp, 32, kPushThreads / 32);
#define HPB_LAUNCH_PUSH(MB, DEP, STG)                                                             \
    do { if (pp.lat_n > 0 && pp.lat_mode == 2) HPB_LAUNCH_PUSH_(MB, DEP, STG, 2, false);           \
         else if (pp.lat_n > 0) HPB_LAUNCH_PUSH_(MB, DEP, STG, 1, false);                          \
         else HPB_LAUNCH_PUSH_(MB, DEP, STG, 0, false); } while (0)
#define HPB_LAUNCH_PUSH_(MB, DEP, STG, LAT, LAS)                                                  \
    hpb_launch(k_advance_plasma<MB, DEP, STG, LAT, LAS>,                                          \
        (unsigned)((nwarps + kPushThreads / 32 - 1) / (kPushThreads / 32)),                       \
        kPushThreads, 0, ctx->stream,                                                             \
        pp, make_view(sl), comps[HPB_C_PSI], comps[HPB_C_EZ], comps[HPB_C_BX],           \
        comps[HPB_C_BY], comps[HPB_C_BZ], g.x_off, g.y_off, 1.0 / g.dx, 1.0 / g.dy, g.c,          \
        charge / (mass * g.c), g.dz / n_subcycles, n_subcycles, temp_slice, particle_bc,         \
        bc_lo[0], bc_lo[1], bc_hi[0], bc_hi[1], dep, lasa)
    // round-1 kernels: staged gather, 5 blocks / SM (the shared-memory limit; variants 0 / 2 / 4),
    // 1 = direct loads only, 3 = staged with 128 registers
    if (lasa.c_aabs >= 0) {
        HPB_LAUNCH_PUSH_(4, false, true, 0, true);          // laser: staged gather, plain order, 128 registers
    } else if (deposit) {
        if (variant == 1) HPB_LAUNCH_PUSH(6, true, false);
        else if (variant == 3) HPB_LAUNCH_PUSH(4, true, true);
        else HPB_LAUNCH_PUSH(5, true, true);
    } else {
        if (variant == 1) HPB_LAUNCH_PUSH(6, false, false);
        else if (variant == 3) HPB_LAUNCH_PUSH(4, false, true);
        else HPB_LAUNCH_PUSH(5, false, true);
    }
#undef HPB_LAUNCH_PUSH
#undef HPB_LAUNCH_PUSH_
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_advance_plasma_particles(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl,
                                            double charge, double mass, int n_subcycles,
                                            int temp_slice, int particle_bc, const double bc_lo[2],
                                            const double bc_hi[2], const int *comps)
{
    return advance_plasma(ctx, pl, sl, charge, mass, n_subcycles, temp_slice, particle_bc, bc_lo, bc_hi,
                          comps, false, 0., nullptr);
}

extern "C" int hpb_advance_plasma_particles_and_deposit(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl,
                                                        double charge, double mass, int n_subcycles,
                                                        int particle_bc, const double bc_lo[2],
                                                        const double bc_hi[2], const int *comps,
                                                        double max_qsa_weighting_factor,
                                                        int *d_n_qsa_violation)
{
    if (comps && (comps[HPB_C_JX] < 0 || comps[HPB_C_JY] < 0 || comps[HPB_C_CHI] < 0
                  || comps[HPB_C_RHOMJZ] < 0 || comps[HPB_C_RHO] >= 0)) {
        hpb_set_error("advance+deposit: needs jx, jy, chi, rhomjz and no rho component");
        return HPB_ERR_UNSUPPORTED;
    }
    return advance_plasma(ctx, pl, sl, charge, mass, n_subcycles, 0, particle_bc, bc_lo, bc_hi, comps,
                          true, max_qsa_weighting_factor, d_n_qsa_violation);
}
