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

namespace {

constexpr int kThreads = 256;

struct PlasmaPtrs {
    double *r[HPB_PLASMA_NREAL];
    uint64_t *idcpu;
    long np;
};
PlasmaPtrs to_ptrs(const hpb_plasma &pl)
{
    PlasmaPtrs p;
    for (int i = 0; i < HPB_PLASMA_NREAL; ++i) p.r[i] = pl.r[i];
    p.idcpu = pl.idcpu; p.np = pl.np;
    return p;
}

// -------------------------------------------------------------------------------------------
// DepositCurrent
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_deposit_current(PlasmaPtrs pl, SliceView a, int c_jx, int c_jy, int c_rho, int c_chi,
                  int c_rhomjz, double x_off, double y_off, double dx_inv, double dy_inv,
                  double clightinv, double charge_invvol, double charge_mu0_mass_ratio,
                  double max_qsa, int *n_qsa_violation)
{
    const long ip = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= pl.np) return;
    const uint64_t idcpu = pl.idcpu[ip];
    if (!hpb_is_valid(idcpu)) return;

    const double psi_inv = 1.0 / pl.r[HPB_PSI][ip];
    const double xp = pl.r[HPB_X][ip];
    const double yp = pl.r[HPB_Y][ip];
    const double vx_c = pl.r[HPB_UX][ip] * psi_inv;
    const double vy_c = pl.r[HPB_UY][ip] * psi_inv;
    const double q_invvol = charge_invvol * pl.r[HPB_W][ip];

    const double gamma_psi = 0.5 * (psi_inv * psi_inv
                                    + vx_c * vx_c * clightinv * clightinv
                                    + vy_c * vy_c * clightinv * clightinv + 1.0);
    if (gamma_psi < 0.0 || gamma_psi > max_qsa || psi_inv < 0.0) {
        // QSA violation: discard the particle (PlasmaDepositCurrent.cpp:197-204)
        if (n_qsa_violation) atomicAdd(n_qsa_violation, 1);
        pl.r[HPB_W][ip] = 0.0;
        pl.idcpu[ip] = hpb_make_invalid(idcpu);
        return;
    }
    double sx[3], sy[3];
    const int i0 = shape2((xp - x_off) * dx_inv, sx);
    const int j0 = shape2((yp - y_off) * dy_inv, sy);
    const double chi_fac = charge_mu0_mass_ratio * psi_inv;
#pragma unroll
    for (int iy = 0; iy < 3; ++iy) {
#pragma unroll
        for (int ix = 0; ix < 3; ++ix) {
            const double cd = q_invvol * sx[ix] * sy[iy];
            const long o = a.idx(i0 + ix, j0 + iy);
            if (c_jx >= 0) {
                red_add(a.comp(c_jx) + o, cd * vx_c);
                red_add(a.comp(c_jy) + o, cd * vy_c);
            }
            if (c_rho >= 0) red_add(a.comp(c_rho) + o, cd * gamma_psi);
            if (c_chi >= 0) red_add(a.comp(c_chi) + o, cd * chi_fac);
            if (c_rhomjz >= 0) red_add(a.comp(c_rhomjz) + o, cd);
        }
    }
}

// -------------------------------------------------------------------------------------------
// beam DepositCurrentSlice
// -------------------------------------------------------------------------------------------
struct BeamPtrs { double *x, *y, *z, *w, *ux, *uy, *uz; uint64_t *idcpu; long np; };

__global__ void __launch_bounds__(kThreads)
k_beam_deposit(BeamPtrs b, SliceView a, int c_jx, int c_jy, int c_jz, double x_off, double y_off,
               double dx_inv, double dy_inv, double clightsq, double q_invvol)
{
    const long ip = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= b.np) return;
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
__global__ void __launch_bounds__(kThreads)
k_explicit_deposition(PlasmaPtrs pl, SliceView a, int c_sy, int c_sx, int c_bz, int c_ez,
                      int c_exmby, int c_eypbx, double x_off, double y_off, double dx_inv,
                      double dy_inv, double a_clight, double clight_inv,
                      double charge_invvol_mu0, double q_mass_ratio)
{
    const long ip = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= pl.np) return;
    if (!hpb_is_valid(pl.idcpu[ip])) return;

    const double psi_inv = 1.0 / pl.r[HPB_PSI][ip];
    const double xp = pl.r[HPB_X][ip];
    const double yp = pl.r[HPB_Y][ip];
    const double vx = pl.r[HPB_UX][ip] * psi_inv * clight_inv;
    const double vy = pl.r[HPB_UY][ip] * psi_inv * clight_inv;
    const double cdm = charge_invvol_mu0 * pl.r[HPB_W][ip];
    const double gamma_psi = 0.5 * (psi_inv * psi_inv + vx * vx + vy * vy + 1.0);

    double sx[5], dsx[5], sy[5], dsy[5];
    const int i0 = dshape2_centered((xp - x_off) * dx_inv, sx, dsx);
    const int j0 = dshape2_centered((yp - y_off) * dy_inv, sy, dsy);
    const double *Bz = a.comp(c_bz), *Ez = a.comp(c_ez);
    const double *ExmBy = a.comp(c_exmby), *EypBx = a.comp(c_eypbx);
    double *Sy = a.comp(c_sy), *Sx = a.comp(c_sx);
    const double qp = q_mass_ratio * psi_inv;
#pragma unroll
    for (int iy = 0; iy < 5; ++iy) {
#pragma unroll
        for (int ix = 0; ix < 5; ++ix) {
            if ((ix == 0 || ix == 4) && (iy == 0 || iy == 4)) continue;
            const long o = a.idx(i0 + ix, j0 + iy);
            const double shx = sx[ix], shdx = dsx[ix], shy = sy[iy], shdy = dsy[iy];
            const double Bz_v = Bz[o], Ez_v = Ez[o], ExmBy_v = ExmBy[o], EypBx_v = EypBx[o];
            red_add(Sy + o, cdm * (
                - shx * shy * (
                    - Bz_v * vx
                    + ( Ez_v * vy
                    + ExmBy_v * (          - vx * vy)
                    + EypBx_v * (gamma_psi - vy * vy) ) * clight_inv
                ) * qp
                + ( - shdx * shy * dx_inv * ( - vx * vy )
                    - shx * shdy * dy_inv * ( gamma_psi - vy * vy - 1.0 )) * a_clight));
            red_add(Sx + o, cdm * (
                + shx * shy * (
                    + Bz_v * vy
                    + ( Ez_v * vx
                    + ExmBy_v * (gamma_psi - vx * vx)
                    + EypBx_v * (          - vx * vy) ) * clight_inv
                ) * qp
                + ( + shdx * shy * dx_inv * ( gamma_psi - vx * vx - 1.0 )
                    + shx * shdy * dy_inv * ( - vx * vy )) * a_clight));
        }
    }
}

// -------------------------------------------------------------------------------------------
// gather + push
// -------------------------------------------------------------------------------------------
struct Dual { double v, e; };
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.e + b.e}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.e - b.e}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.e * b.v + a.v * b.e}; }
__device__ __forceinline__ Dual operator*(Dual a, double b) { return {a.v * b, a.e * b}; }
__device__ __forceinline__ Dual operator*(double a, Dual b) { return {a * b.v, a * b.e}; }
__device__ __forceinline__ Dual operator+(Dual a, double b) { return {a.v + b, a.e}; }
__device__ __forceinline__ Dual operator+(double a, Dual b) { return {a + b.v, b.e}; }
__device__ __forceinline__ Dual operator-(Dual a, double b) { return {a.v - b, a.e}; }

struct PushFields { double ExmBy, EypBx, Ez, Bx_c, By_c, Bz; };

// PlasmaMomentumPush<T>, PushPlasmaParticles.H:39-75 (laser terms are zero)
template <class T>
__device__ __forceinline__ void momentum_push(const T &ux, const T &uy, const T &psi_inv,
                                              const PushFields &f, double clight_inv, double qmc,
                                              T &dz_ux, T &dz_uy, T &dz_psi)
{
    const double c2 = clight_inv * clight_inv;
    const T gamma_psi = 0.5 * psi_inv * psi_inv * (1.0 + ux * ux * c2 + uy * uy * c2) + 0.5;
    dz_ux = qmc * (gamma_psi * f.ExmBy + f.By_c + (uy * f.Bz) * psi_inv);
    dz_uy = qmc * (gamma_psi * f.EypBx - f.Bx_c - (ux * f.Bz) * psi_inv);
    dz_psi = (qmc * clight_inv) * ((ux * f.ExmBy + uy * f.EypBx) * clight_inv * psi_inv - f.Ez);
}

__device__ __forceinline__ void push_substep(double &ux, double &uy, double &psi,
                                             const PushFields &f, double clight_inv, double qmc,
                                             double sdz)
{
    const double psi_inv = 1.0 / psi;
    double dz_ux, dz_uy, dz_psi;
    momentum_push<double>(ux, uy, psi_inv, f, clight_inv, qmc, dz_ux, dz_uy, dz_psi);
    const Dual ux_d{ux, dz_ux}, uy_d{uy, dz_uy}, pi_d{psi_inv, -psi_inv * psi_inv * dz_psi};
    Dual d_ux, d_uy, d_psi;
    momentum_push<Dual>(ux_d, uy_d, pi_d, f, clight_inv, qmc, d_ux, d_uy, d_psi);
    ux += sdz * dz_ux + 0.5 * sdz * sdz * d_ux.e;
    uy += sdz * dz_uy + 0.5 * sdz * sdz * d_uy.e;
    psi += sdz * dz_psi + 0.5 * sdz * sdz * d_psi.e;
}

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

__global__ void __launch_bounds__(kThreads)
k_advance_plasma(PlasmaPtrs pl, SliceView a, int c_psi, int c_ez, int c_bx, int c_by, int c_bz,
                 double x_off, double y_off, double dx_inv, double dy_inv, double clight,
                 double qmc, double dz, int n_subcycles, int temp_slice, int bc, double lox,
                 double loy, double hix, double hiy)
{
    const long ip = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= pl.np) return;
    const uint64_t idcpu = pl.idcpu[ip];
    if (!hpb_is_valid(idcpu)) return;
    const double clight_inv = 1.0 / clight;
    const double *Psi = a.comp(c_psi), *Ez = a.comp(c_ez), *Bx = a.comp(c_bx);
    const double *By = a.comp(c_by), *Bz = a.comp(c_bz);

    for (int isc = 0; isc < n_subcycles; ++isc) {
        double xp = pl.r[HPB_X_PREV][ip];
        double yp = pl.r[HPB_Y_PREV][ip];
        // doGatherShapeN<2>, FieldGather.H:45-96
        double sx[4], dsx[4], sy[4], dsy[4];
        const int i0 = dshape2_nodal((xp - x_off) * dx_inv, sx, dsx);
        const int j0 = dshape2_nodal((yp - y_off) * dy_inv, sy, dsy);
        PushFields f = {0., 0., 0., 0., 0., 0.};
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
                f.Bx_c += w * Bx[o];
                f.By_c += w * By[o];
                f.Bz += w * Bz[o];
            }
        }
        f.Bx_c *= clight;
        f.By_c *= clight;

        constexpr int nsub = 4;
        const double sdz = dz / nsub;
        double ux = pl.r[HPB_UX_HALF][ip];
        double uy = pl.r[HPB_UY_HALF][ip];
        double psi = pl.r[HPB_PSI_HALF][ip];
#pragma unroll 1
        for (int isub = 0; isub < nsub; ++isub) push_substep(ux, uy, psi, f, clight_inv, qmc, sdz);

        xp += dz * clight_inv * (ux * (1.0 / psi));
        yp += dz * clight_inv * (uy * (1.0 / psi));
        if (enforce_bc(xp, yp, ux, uy, bc, lox, loy, hix, hiy)) {
            pl.r[HPB_W][ip] = 0.0;
            pl.idcpu[ip] = hpb_make_invalid(idcpu);
            return;
        }
        pl.r[HPB_X][ip] = xp;
        pl.r[HPB_Y][ip] = yp;
        if (!temp_slice) {
            pl.r[HPB_UX_HALF][ip] = ux;
            pl.r[HPB_UY_HALF][ip] = uy;
            pl.r[HPB_PSI_HALF][ip] = psi;
            pl.r[HPB_X_PREV][ip] = xp;
            pl.r[HPB_Y_PREV][ip] = yp;
        }
#pragma unroll 1
        for (int isub = 0; isub < nsub / 2; ++isub) push_substep(ux, uy, psi, f, clight_inv, qmc, sdz);
        pl.r[HPB_UX][ip] = ux;
        pl.r[HPB_UY][ip] = uy;
        pl.r[HPB_PSI][ip] = psi;
    }
}

inline unsigned nblocks(long n) { return (unsigned)((n + kThreads - 1) / kThreads); }

}  // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" int hpb_deposit_current(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge,
                                   double mass, int c_jx, int c_jy, int c_rho, int c_chi,
                                   int c_rhomjz, double max_qsa, int *d_n_qsa_violation)
{
    if (!ctx) return HPB_ERR_ARG;
    if (pl.np == 0) return HPB_OK;
    const hpb_geom &g = ctx->g;
    // invvol: 1 in normalised units at lev 0, 1/(dx dy dz) in SI (PlasmaDepositCurrent.cpp:71-73)
    const double invvol = g.normalized ? 1.0 : (1.0 / g.dx) * (1.0 / g.dy) * (1.0 / g.dz);
    k_deposit_current<<<nblocks(pl.np), kThreads, 0, ctx->stream>>>(
        to_ptrs(pl), make_view(sl), c_jx, c_jy, c_rho, c_chi, c_rhomjz, g.x_off, g.y_off,
        1.0 / g.dx, 1.0 / g.dy, 1.0 / g.c, charge * invvol, charge * g.mu0 / mass, max_qsa,
        d_n_qsa_violation);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_beam_deposit(hpb_ctx *ctx, hpb_beam_slice bm, hpb_slice sl, double charge,
                                int c_jx, int c_jy, int c_jz)
{
    if (!ctx) return HPB_ERR_ARG;
    if (bm.np == 0 || (c_jx < 0 && c_jz < 0)) return HPB_OK;
    const hpb_geom &g = ctx->g;
    // BeamDepositCurrent.cpp:72-82: invvol = 1 in normalised units at lev 0
    const double invvol = g.normalized ? 1.0 : (1.0 / g.dx) * (1.0 / g.dy) * (1.0 / g.dz);
    BeamPtrs b{bm.x, bm.y, bm.z, bm.w, bm.ux, bm.uy, bm.uz, bm.idcpu, bm.np};
    k_beam_deposit<<<nblocks(bm.np), kThreads, 0, ctx->stream>>>(
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
    const hpb_geom &g = ctx->g;
    const double invvol = g.normalized ? 1.0 : (1.0 / g.dx) * (1.0 / g.dy) * (1.0 / g.dz);
    k_explicit_deposition<<<nblocks(pl.np), kThreads, 0, ctx->stream>>>(
        to_ptrs(pl), make_view(sl), comps[HPB_C_SY], comps[HPB_C_SX], comps[HPB_C_BZ],
        comps[HPB_C_EZ], comps[HPB_C_EXMBY], comps[HPB_C_EYPBX], g.x_off, g.y_off, 1.0 / g.dx,
        1.0 / g.dy, g.c, 1.0 / g.c, charge * invvol * g.mu0, charge / mass);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_advance_plasma_particles(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl,
                                            double charge, double mass, int n_subcycles,
                                            int temp_slice, int particle_bc, const double bc_lo[2],
                                            const double bc_hi[2], const int *comps)
{
    if (!ctx || !comps || !bc_lo || !bc_hi || n_subcycles < 1) return HPB_ERR_ARG;
    if (pl.np == 0) return HPB_OK;
    const hpb_geom &g = ctx->g;
    k_advance_plasma<<<nblocks(pl.np), kThreads, 0, ctx->stream>>>(
        to_ptrs(pl), make_view(sl), comps[HPB_C_PSI], comps[HPB_C_EZ], comps[HPB_C_BX],
        comps[HPB_C_BY], comps[HPB_C_BZ], g.x_off, g.y_off, 1.0 / g.dx, 1.0 / g.dy, g.c,
        charge / (mass * g.c), g.dz / n_subcycles, n_subcycles, temp_slice, particle_bc,
        bc_lo[0], bc_lo[1], bc_hi[0], bc_hi[1]);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}
