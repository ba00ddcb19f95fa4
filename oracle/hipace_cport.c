/* C/OpenMP restatement of the particle, stencil and multigrid kernels of the HiPACE++ slice loop.
 *
 * THIS IS TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.  It is the multi-threaded companion
 * of oracle/hipace_oracle.py (same arithmetic, same cites) and exists so that bench.py can time
 * "the reference's CPU algorithm" on all host cores in seconds instead of minutes.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * Parity status: PINNED -- tests/test_cport.py checks it against the NumPy oracle per cell and,
 * through oracle/cport.py, against the reference's golden checksums.
 *
 * Threading follows the reference's CPU strategy: particles are binned to 32x32-cell tiles and
 * the tiles are processed in 4 colours so that concurrently processed tiles never overlap
 * (src/particles/deposition/DepositionUtil.H:204-254); gather/push is a parallel loop over
 * particles (src/utils/OMPUtil.H:28-35); hpmg sweeps are parallel over rows
 * (src/mg_solver/HpMultiGrid.cpp:594-740 computes the same values as whole-array sweeps).
 *
 * Array convention: slice components are a[(j+G)*nxt + (i+G)], nxt = nx + 2G, G = 2.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define G 2
#define TILE 32

int hpc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* the caller decides the thread count (bench.py: the host's physical cores, whatever
 * OMP_NUM_THREADS a launcher such as torchrun exported); returns the count in effect */
int hpc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ---- shape factors (src/particles/particles_utils/ShapeFactors.H) ----------------------- */

/* compute_single_shape_factor<false,2>, :165-174 */
static inline int shape2(double xmid, double s[3])
{
    const double xfloor = floor(xmid + 0.5);
    const double xint = xmid - xfloor;
    s[0] = 0.5 * (0.5 - xint) * (0.5 - xint);
    s[1] = 0.75 - xint * xint;
    s[2] = 0.5 * (0.5 + xint) * (0.5 + xint);
    return (int)xfloor - 1;
}

/* single_derivative_shape_factor<2,2>, :405-430 (ds already "-sdx") */
static inline int dshape2_centered(double xmid, double s[5], double ds[5])
{
    xmid += 0.5;
    const double xfloor = floor(xmid);
    const double xint = xmid - xfloor;
    const double x2 = xint * xint;
    s[0] = 0.; s[1] = 0.5 * x2 - xint + 0.5; s[2] = -x2 + xint + 0.5; s[3] = 0.5 * x2; s[4] = 0.;
    ds[0] = -(-0.25 * x2 + 0.5 * xint - 0.25);
    ds[1] = -(0.5 * x2 - 0.5 * xint - 0.25);
    ds[2] = -(0.25 - 0.5 * xint);
    ds[3] = -(-0.5 * x2 + 0.5 * xint + 0.25);
    ds[4] = -(0.25 * x2);
    return (int)xfloor - 2;
}

/* single_derivative_shape_factor<1,2>, :305-329 */
static inline int dshape2_nodal(double xmid, double s[4], double ds[4])
{
    const double xfloor = floor(xmid);
    const double xint = xmid - xfloor;
    const double x2 = xint * xint;
    const int lo = xint < 0.5;
    s[0] = lo ? 0.5 * x2 - 0.5 * xint + 0.125 : 0.;
    s[1] = lo ? 0.75 - x2 : 0.5 * x2 - 1.5 * xint + 1.125;
    s[2] = lo ? 0.5 * x2 + 0.5 * xint + 0.125 : -x2 + 2. * xint - 0.25;
    s[3] = lo ? 0. : 0.5 * x2 - 0.5 * xint + 0.125;
    ds[0] = -(-0.5 * x2 + xint - 0.5);
    ds[1] = -(1.5 * x2 - 2. * xint);
    ds[2] = -(-1.5 * x2 + xint + 0.5);
    ds[3] = -(0.5 * x2);
    return (int)xfloor - 1;
}

/* ---- tile binning -------------------------------------------------------------------------- */
/* stable counting sort of the valid particles by the tile of their lowest deposit cell;
 * returns perm (length nvalid) and tile offsets (ntile+1).  shift = stencil half width. */
static long bin_particles(long np, const double *x, const double *y, const uint8_t *valid,
                          double x_off, double y_off, double dx_inv, double dy_inv, int nx, int ny,
                          int centered, int ntx, int nty, long **perm_out, long **off_out)
{
    const long ntile = (long)ntx * nty;
    long *cnt = (long *)calloc(ntile + 1, sizeof(long));
    int *tile = (int *)malloc(sizeof(int) * (np > 0 ? np : 1));
#pragma omp parallel for schedule(static)
    for (long p = 0; p < np; ++p) {
        if (!valid[p]) { tile[p] = -1; continue; }
        const double xm = (x[p] - x_off) * dx_inv, ym = (y[p] - y_off) * dy_inv;
        int i0, j0;
        if (centered) { i0 = (int)floor(xm + 0.5) - 2; j0 = (int)floor(ym + 0.5) - 2; }
        else { i0 = (int)floor(xm + 0.5) - 1; j0 = (int)floor(ym + 0.5) - 1; }
        int ti = (i0 + G) / TILE, tj = (j0 + G) / TILE;
        if (ti < 0) ti = 0; if (ti >= ntx) ti = ntx - 1;
        if (tj < 0) tj = 0; if (tj >= nty) tj = nty - 1;
        tile[p] = tj * ntx + ti;
    }
    for (long p = 0; p < np; ++p) if (tile[p] >= 0) cnt[tile[p] + 1]++;
    for (long t = 0; t < ntile; ++t) cnt[t + 1] += cnt[t];
    const long nvalid = cnt[ntile];
    long *perm = (long *)malloc(sizeof(long) * (nvalid > 0 ? nvalid : 1));
    long *pos = (long *)malloc(sizeof(long) * (ntile + 1));
    memcpy(pos, cnt, sizeof(long) * (ntile + 1));
    for (long p = 0; p < np; ++p) if (tile[p] >= 0) perm[pos[tile[p]]++] = p;
    free(pos); free(tile);
    (void)nx; (void)ny;
    *perm_out = perm; *off_out = cnt;
    return nvalid;
}

/* ---- ::DepositCurrent (src/particles/deposition/PlasmaDepositCurrent.cpp:155-246) ---------- */
long hpc_deposit_current(long np, const double *x, const double *y, double *w, const double *ux,
                         const double *uy, const double *psi, uint8_t *valid, double *jx,
                         double *jy, double *rho, double *chi, double *rhomjz, int nx, int ny,
                         double x_off, double y_off, double dx_inv, double dy_inv, double clightinv,
                         double charge_invvol, double charge_mu0_mass_ratio, double max_qsa)
{
    const int nxt = nx + 2 * G;
    long n_bad = 0;
    /* QSA check first (:197-204) */
#pragma omp parallel for schedule(static) reduction(+ : n_bad)
    for (long p = 0; p < np; ++p) {
        if (!valid[p]) continue;
        const double psi_inv = 1.0 / psi[p];
        const double vx_c = ux[p] * psi_inv, vy_c = uy[p] * psi_inv;
        const double gamma_psi = 0.5 * (psi_inv * psi_inv + vx_c * vx_c * clightinv * clightinv
                                        + vy_c * vy_c * clightinv * clightinv + 1.0);
        if (gamma_psi < 0.0 || gamma_psi > max_qsa || psi_inv < 0.0) {
            w[p] = 0.0; valid[p] = 0; ++n_bad;
        }
    }
    const int ntx = (nxt + TILE - 1) / TILE, nty = (ny + 2 * G + TILE - 1) / TILE;
    long *perm, *off;
    bin_particles(np, x, y, valid, x_off, y_off, dx_inv, dy_inv, nx, ny, 0, ntx, nty, &perm, &off);
    for (int colour = 0; colour < 4; ++colour) {
        const int cx = colour & 1, cy = colour >> 1;
        const int ncx = (ntx - cx + 1) / 2, ncy = (nty - cy + 1) / 2;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
        for (int tyh = 0; tyh < ncy; ++tyh) {
            for (int txh = 0; txh < ncx; ++txh) {
                const long t = (long)(2 * tyh + cy) * ntx + (2 * txh + cx);
                for (long k = off[t]; k < off[t + 1]; ++k) {
                    const long p = perm[k];
                    const double psi_inv = 1.0 / psi[p];
                    const double vx_c = ux[p] * psi_inv, vy_c = uy[p] * psi_inv;
                    const double q_invvol = charge_invvol * w[p];
                    const double gamma_psi = 0.5 * (psi_inv * psi_inv + vx_c * vx_c * clightinv * clightinv
                                                    + vy_c * vy_c * clightinv * clightinv + 1.0);
                    double sx[3], sy[3];
                    const int i0 = shape2((x[p] - x_off) * dx_inv, sx);
                    const int j0 = shape2((y[p] - y_off) * dy_inv, sy);
                    for (int iy = 0; iy < 3; ++iy)
                        for (int ix = 0; ix < 3; ++ix) {
                            const double cd = q_invvol * sx[ix] * sy[iy];
                            const long o = (long)(j0 + iy + G) * nxt + (i0 + ix + G);
                            if (jx) { jx[o] += cd * vx_c; jy[o] += cd * vy_c; }
                            if (rho) rho[o] += cd * gamma_psi;
                            if (chi) chi[o] += cd * charge_mu0_mass_ratio * psi_inv;
                            if (rhomjz) rhomjz[o] += cd;
                        }
                }
            }
        }
    }
    free(perm); free(off);
    return n_bad;
}

/* ---- ::ExplicitDeposition (src/particles/deposition/ExplicitDeposition.cpp:140-261) -------- */
void hpc_explicit_deposition(long np, const double *x, const double *y, const double *w,
                             const double *ux, const double *uy, const double *psi,
                             const uint8_t *valid, double *Sy, double *Sx, const double *Bz,
                             const double *Ez, const double *ExmBy, const double *EypBx, int nx,
                             int ny, double x_off, double y_off, double dx_inv, double dy_inv,
                             double a_clight, double clight_inv, double charge_invvol_mu0,
                             double q_mass_ratio)
{
    const int nxt = nx + 2 * G;
    const int ntx = (nxt + TILE - 1) / TILE, nty = (ny + 2 * G + TILE - 1) / TILE;
    long *perm, *off;
    bin_particles(np, x, y, valid, x_off, y_off, dx_inv, dy_inv, nx, ny, 1, ntx, nty, &perm, &off);
    for (int colour = 0; colour < 4; ++colour) {
        const int cx = colour & 1, cy = colour >> 1;
        const int ncx = (ntx - cx + 1) / 2, ncy = (nty - cy + 1) / 2;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
        for (int tyh = 0; tyh < ncy; ++tyh) {
            for (int txh = 0; txh < ncx; ++txh) {
                const long t = (long)(2 * tyh + cy) * ntx + (2 * txh + cx);
                for (long k = off[t]; k < off[t + 1]; ++k) {
                    const long p = perm[k];
                    const double psi_inv = 1.0 / psi[p];
                    const double vx = ux[p] * psi_inv * clight_inv, vy = uy[p] * psi_inv * clight_inv;
                    const double cdm = charge_invvol_mu0 * w[p];
                    const double gamma_psi = 0.5 * (psi_inv * psi_inv + vx * vx + vy * vy + 1.0);
                    double sx[5], dsx[5], sy[5], dsy[5];
                    const int i0 = dshape2_centered((x[p] - x_off) * dx_inv, sx, dsx);
                    const int j0 = dshape2_centered((y[p] - y_off) * dy_inv, sy, dsy);
                    for (int iy = 0; iy < 5; ++iy)
                        for (int ix = 0; ix < 5; ++ix) {
                            if ((ix == 0 || ix == 4) && (iy == 0 || iy == 4)) continue;
                            const long o = (long)(j0 + iy + G) * nxt + (i0 + ix + G);
                            const double shx = sx[ix], shdx = dsx[ix], shy = sy[iy], shdy = dsy[iy];
                            const double Bz_v = Bz[o], Ez_v = Ez[o], ExmBy_v = ExmBy[o], EypBx_v = EypBx[o];
                            Sy[o] += cdm * (
                                - shx * shy * (
                                    - Bz_v * vx
                                    + ( Ez_v * vy
                                    + ExmBy_v * (          - vx * vy)
                                    + EypBx_v * (gamma_psi - vy * vy) ) * clight_inv
                                ) * q_mass_ratio * psi_inv
                                + ( - shdx * shy * dx_inv * ( - vx * vy )
                                    - shx * shdy * dy_inv * ( gamma_psi - vy * vy - 1.0 )) * a_clight);
                            Sx[o] += cdm * (
                                + shx * shy * (
                                    + Bz_v * vy
                                    + ( Ez_v * vx
                                    + ExmBy_v * (gamma_psi - vx * vx)
                                    + EypBx_v * (          - vx * vy) ) * clight_inv
                                ) * q_mass_ratio * psi_inv
                                + ( + shdx * shy * dx_inv * ( gamma_psi - vx * vx - 1.0 )
                                    + shx * shdy * dy_inv * ( - vx * vy )) * a_clight);
                        }
                }
            }
        }
    }
    free(perm); free(off);
}

/* ---- gather + push (src/particles/pusher/PlasmaParticleAdvance.cpp:92-217) ----------------- */
typedef struct { double v, e; } dual;
static inline dual dmul(dual a, dual b) { dual r = {a.v * b.v, a.e * b.v + a.v * b.e}; return r; }
static inline dual dadd(dual a, dual b) { dual r = {a.v + b.v, a.e + b.e}; return r; }
static inline dual dsub(dual a, dual b) { dual r = {a.v - b.v, a.e - b.e}; return r; }
static inline dual dr(double a) { dual r = {a, 0.0}; return r; }

/* one sub-step: PlasmaMomentumPush<Real> + <DualNumber> (PushPlasmaParticles.H:39-75,
 * DualNumbers.H:13-43), update as PlasmaParticleAdvance.cpp:152-166 */
static inline void substep(double *ux, double *uy, double *psi, double ExmBy, double EypBx,
                           double Ez, double Bx_c, double By_c, double Bz, double clight_inv,
                           double qmc, double sdz)
{
    const double c2 = clight_inv * clight_inv;
    const double psi_inv = 1.0 / *psi;
    const double gp = 0.5 * psi_inv * psi_inv * (1.0 + (*ux) * (*ux) * c2 + (*uy) * (*uy) * c2) + 0.5;
    const double dux = qmc * (gp * ExmBy + By_c + ((*uy) * Bz) * psi_inv);
    const double duy = qmc * (gp * EypBx - Bx_c - ((*ux) * Bz) * psi_inv);
    const double dps = qmc * clight_inv * (((*ux) * ExmBy + (*uy) * EypBx) * clight_inv * psi_inv - Ez);
    const dual uxd = {*ux, dux}, uyd = {*uy, duy}, pid = {psi_inv, -psi_inv * psi_inv * dps};
    dual t = dmul(dmul(dr(0.5), pid), pid);
    dual s = dadd(dadd(dr(1.0 + 0.0), dmul(dmul(uxd, uxd), dr(c2))), dmul(dmul(uyd, uyd), dr(c2)));
    dual gpd = dadd(dmul(t, s), dr(0.5));
    dual a = dadd(dadd(dmul(gpd, dr(ExmBy)), dr(By_c)), dmul(dmul(uyd, dr(Bz)), pid));
    const double duxe = dmul(dr(qmc), a).e;
    a = dsub(dsub(dmul(gpd, dr(EypBx)), dr(Bx_c)), dmul(dmul(uxd, dr(Bz)), pid));
    const double duye = dmul(dr(qmc), a).e;
    a = dsub(dmul(dmul(dadd(dmul(uxd, dr(ExmBy)), dmul(uyd, dr(EypBx))), dr(clight_inv)), pid), dr(Ez));
    const double dpse = dmul(dr(qmc * clight_inv), a).e;
    *ux = *ux + (sdz * dux + 0.5 * sdz * sdz * duxe);
    *uy = *uy + (sdz * duy + 0.5 * sdz * sdz * duye);
    *psi = *psi + (sdz * dps + 0.5 * sdz * sdz * dpse);
}

static inline double wrap(double v, double lo, double len)
{
    v = fmod(v - lo, len);
    if (v < 0) v += len;
    return v + lo;
}

/* bc: 0 reflecting, 1 periodic, 2 absorbing (GetAndSetPosition.H:29-99) */
void hpc_advance_plasma(long np, double *x, double *y, double *w, double *ux, double *uy,
                        double *psi, double *x_prev, double *y_prev, double *ux_half,
                        double *uy_half, double *psi_half, uint8_t *valid, const double *Psi,
                        const double *Ez, const double *Bx, const double *By, const double *Bz,
                        int nx, int ny, double x_off, double y_off, double dx_inv, double dy_inv,
                        double clight, double qmc, double dz, int n_subcycles, int temp_slice,
                        int bc, double lox, double loy, double hix, double hiy)
{
    const int nxt = nx + 2 * G;
    const double clight_inv = 1.0 / clight;
    (void)ny;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < np; ++p) {
        if (!valid[p]) continue;
        for (int isc = 0; isc < n_subcycles; ++isc) {
            double xp = x_prev[p], yp = y_prev[p];
            double sx[4], dsx[4], sy[4], dsy[4];
            const int i0 = dshape2_nodal((xp - x_off) * dx_inv, sx, dsx);
            const int j0 = dshape2_nodal((yp - y_off) * dy_inv, sy, dsy);
            double fExmBy = 0., fEypBx = 0., fEz = 0., fBx = 0., fBy = 0., fBz = 0.;
            for (int iy = 0; iy < 4; ++iy)
                for (int ix = 0; ix < 4; ++ix) {
                    const long o = (long)(j0 + iy + G) * nxt + (i0 + ix + G);
                    const double psi_v = Psi[o];
                    fExmBy += (dsx[ix] * sy[iy]) * psi_v * dx_inv;
                    fEypBx += (sx[ix] * dsy[iy]) * psi_v * dy_inv;
                    const double ww = sx[ix] * sy[iy];
                    fEz += ww * Ez[o]; fBx += ww * Bx[o]; fBy += ww * By[o]; fBz += ww * Bz[o];
                }
            fBx *= clight; fBy *= clight;
            const double sdz = dz / 4;
            double u = ux_half[p], v = uy_half[p], ps = psi_half[p];
            for (int k = 0; k < 4; ++k) substep(&u, &v, &ps, fExmBy, fEypBx, fEz, fBx, fBy, fBz, clight_inv, qmc, sdz);
            xp = xp + dz * clight_inv * (u * (1.0 / ps));
            yp = yp + dz * clight_inv * (v * (1.0 / ps));
            if (xp < lox || yp < loy || xp > hix || yp > hiy) {
                const double lenx = hix - lox, leny = hiy - loy;
                if (bc == 1) { xp = wrap(xp, lox, lenx); yp = wrap(yp, loy, leny); }
                else if (bc == 0) {
                    xp = wrap(xp, lox, 2 * lenx); if (xp > hix) { xp = 2 * hix - xp; u = -u; }
                    yp = wrap(yp, loy, 2 * leny); if (yp > hiy) { yp = 2 * hiy - yp; v = -v; }
                } else { w[p] = 0.0; valid[p] = 0; break; }
            }
            x[p] = xp; y[p] = yp;
            if (!temp_slice) { ux_half[p] = u; uy_half[p] = v; psi_half[p] = ps; x_prev[p] = xp; y_prev[p] = yp; }
            for (int k = 0; k < 2; ++k) substep(&u, &v, &ps, fExmBy, fEypBx, fEz, fBx, fBy, fBz, clight_inv, qmc, sdz);
            ux[p] = u; uy[p] = v; psi[p] = ps;
        }
    }
}

/* ---- field stencils (src/fields/Fields.cpp:880-956, src/Hipace.cpp:775-788) ----------------- */
void hpc_poisson_rhs(const double *rhomjz, const double *jx, const double *jy, int nx, int ny,
                     double *rhs /* 3*ny*nx */, double f_psi, double f_ez, double mu0, double dx,
                     double dy)
{
    const int nxt = nx + 2 * G;
    const long n = (long)nx * ny;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const long o = (long)(j + G) * nxt + (i + G), s = (long)j * nx + i;
            const double dx_jx = (jx[o + 1] - jx[o - 1]) * (0.5 / dx);
            const double dy_jy = (jy[o + nxt] - jy[o - nxt]) * (0.5 / dy);
            const double dy_jx = (jx[o + nxt] - jx[o - nxt]) * (0.5 / dy);
            const double dx_jy = (jy[o + 1] - jy[o - 1]) * (0.5 / dx);
            rhs[s] = f_psi * rhomjz[o];
            rhs[s + n] = f_ez * dx_jx + f_ez * dy_jy;
            rhs[s + 2 * n] = mu0 * dy_jx + (-mu0) * dx_jy;
        }
}

void hpc_store_valid(double *dst, const double *src /* ny*nx */, int nx, int ny)
{
    const int nxt = nx + 2 * G;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
        memcpy(dst + (long)(j + G) * nxt + G, src + (long)j * nx, sizeof(double) * nx);
}

void hpc_exmby_eypbx(const double *psi, double *exmby, double *eypbx, int nx, int ny, double dx,
                     double dy)
{
    const int nxt = nx + 2 * G;
#pragma omp parallel for schedule(static)
    for (int j = -1; j < ny + 1; ++j)
        for (int i = -1; i < nx + 1; ++i) {
            const long o = (long)(j + G) * nxt + (i + G);
            exmby[o] = -(psi[o + 1] - psi[o - 1]) * (0.5 / dx);
            eypbx[o] = -(psi[o + nxt] - psi[o - nxt]) * (0.5 / dy);
        }
}

void hpc_sxsy_from_beam(double *Sy, double *Sx, const double *jzb, const double *prev_jxb,
                        const double *prev_jyb, const double *next_jxb, const double *next_jyb,
                        int nx, int ny, double mu0, double dx, double dy, double dz)
{
    const int nxt = nx + 2 * G;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const long o = (long)(j + G) * nxt + (i + G);
            const double dx_jzb = (jzb[o + 1] - jzb[o - 1]) / (2.0 * dx);
            const double dy_jzb = (jzb[o + nxt] - jzb[o - nxt]) / (2.0 * dy);
            const double dz_jxb = (prev_jxb[o] - next_jxb[o]) / (2.0 * dz);
            const double dz_jyb = (prev_jyb[o] - next_jyb[o]) / (2.0 * dz);
            Sy[o] = mu0 * (-dy_jzb + dz_jyb);
            Sx[o] = -mu0 * (-dx_jzb + dz_jxb);
        }
}

/* ---- hpmg system type 1 (src/mg_solver/HpMultiGrid.cpp) ------------------------------------- */
typedef struct {
    int nx, ny;         /* points of the level box */
    double *acf, *res, *cor, *rescor;   /* res/cor/rescor: 2 comps */
} mglev;

typedef struct {
    int cc, nlev;
    double dx, dy;
    mglev lev[32];
} hpc_mg;

hpc_mg *hpc_mg_create(int nx, int ny, double dx, double dy)
{
    if ((nx % 2) != (ny % 2)) return NULL;                       /* :1051-1052 */
    hpc_mg *m = (hpc_mg *)calloc(1, sizeof(hpc_mg));
    m->cc = (nx % 2 == 0); m->dx = dx; m->dy = dy;
    int w = m->cc ? nx : nx + 2, h = m->cc ? ny : ny + 2, nl = 0;
    for (;;) {                                                   /* :1054-1072 */
        m->lev[nl].nx = w; m->lev[nl].ny = h; ++nl;
        int ok = m->cc ? (w % 2 == 0 && h % 2 == 0 && w >= 4 && h >= 4)
                       : ((w - 1) % 2 == 0 && (h - 1) % 2 == 0 && w >= 8 && h >= 8);
        if (!ok || nl >= 30) break;
        if (m->cc) { w /= 2; h /= 2; } else { w = (w - 1) / 2 + 1; h = (h - 1) / 2 + 1; }
    }
    m->nlev = nl;
    for (int l = 0; l < nl; ++l) {
        const size_t n = (size_t)m->lev[l].nx * m->lev[l].ny;
        m->lev[l].acf = (double *)calloc(n, sizeof(double));
        m->lev[l].res = (double *)calloc(2 * n, sizeof(double));
        m->lev[l].cor = (double *)calloc(2 * n, sizeof(double));
        m->lev[l].rescor = (double *)calloc(2 * n, sizeof(double));
    }
    return m;
}

void hpc_mg_destroy(hpc_mg *m)
{
    if (!m) return;
    for (int l = 0; l < m->nlev; ++l) { free(m->lev[l].acf); free(m->lev[l].res); free(m->lev[l].cor); free(m->lev[l].rescor); }
    free(m);
}

typedef struct { int nx, ny, vlo, vhix, vhiy, cc; double facx, facy; } lgeom;

static lgeom level_geom(const hpc_mg *m, int l)
{
    lgeom g;
    g.nx = m->lev[l].nx; g.ny = m->lev[l].ny; g.cc = m->cc;
    g.vlo = g.cc ? 0 : 1; g.vhix = g.cc ? g.nx - 1 : g.nx - 2; g.vhiy = g.cc ? g.ny - 1 : g.ny - 2;
    const double dx = m->dx * (double)(1 << l), dy = m->dy * (double)(1 << l);
    g.facx = 1.0 / (dx * dx); g.facy = 1.0 / (dy * dy);
    return g;
}

/* gs1 (:265-292) for both components at (i, j) */
static inline void gs_point(const lgeom *g, double *p0, double *p1, const double *r0,
                            const double *r1, const double *acf, int i, int j)
{
    const long o = (long)j * g->nx + i;
    double c0 = -(acf[o] + 2.0 * (g->facx + g->facy));
    double l0, l1;
    if (g->cc && i == g->vlo) { l0 = g->facx * (4. / 3.) * p0[o + 1]; l1 = g->facx * (4. / 3.) * p1[o + 1]; c0 -= 2.0 * g->facx; }
    else if (g->cc && i == g->vhix) { l0 = g->facx * (4. / 3.) * p0[o - 1]; l1 = g->facx * (4. / 3.) * p1[o - 1]; c0 -= 2.0 * g->facx; }
    else { l0 = g->facx * (p0[o - 1] + p0[o + 1]); l1 = g->facx * (p1[o - 1] + p1[o + 1]); }
    if (g->cc && j == g->vlo) { l0 += g->facy * (4. / 3.) * p0[o + g->nx]; l1 += g->facy * (4. / 3.) * p1[o + g->nx]; c0 -= 2.0 * g->facy; }
    else if (g->cc && j == g->vhiy) { l0 += g->facy * (4. / 3.) * p0[o - g->nx]; l1 += g->facy * (4. / 3.) * p1[o - g->nx]; c0 -= 2.0 * g->facy; }
    else { l0 += g->facy * (p0[o - g->nx] + p0[o + g->nx]); l1 += g->facy * (p1[o - g->nx] + p1[o + g->nx]); }
    const double c0_inv = 1.0 / c0;
    p0[o] = (r0[o] - l0) * c0_inv;
    p1[o] = (r1[o] - l1) * c0_inv;
}

/* nsweeps red-black half-sweeps, colour (i + j + icolor) % 2 == 0 (:367-404, :521-548) */
static void gsrb(const lgeom *g, double *phi, const double *rhs, const double *acf, int nsweeps)
{
    const long n = (long)g->nx * g->ny;
    for (int ic = 0; ic < nsweeps; ++ic) {
#pragma omp parallel for schedule(static)
        for (int j = g->vlo; j <= g->vhiy; ++j) {
            int i = g->vlo + ((g->vlo + j + ic) & 1);
            for (; i <= g->vhix; i += 2) gs_point(g, phi, phi + n, rhs, rhs + n, acf, i, j);
        }
    }
}

/* residual1 (:184-190) with the laplacian of :163-182 */
static void residual(const lgeom *g, double *res, const double *phi, const double *rhs, const double *acf)
{
    const long n = (long)g->nx * g->ny;
#pragma omp parallel for schedule(static)
    for (int j = g->vlo; j <= g->vhiy; ++j)
        for (int i = g->vlo; i <= g->vhix; ++i) {
            const long o = (long)j * g->nx + i;
            for (int c = 0; c < 2; ++c) {
                const double *p = phi + c * n;
                double lap = -2.0 * (g->facx + g->facy) * p[o];
                if (g->cc && i == g->vlo) lap += g->facx * ((4. / 3.) * p[o + 1] - 2.0 * p[o]);
                else if (g->cc && i == g->vhix) lap += g->facx * ((4. / 3.) * p[o - 1] - 2.0 * p[o]);
                else lap += g->facx * (p[o - 1] + p[o + 1]);
                if (g->cc && j == g->vlo) lap += g->facy * ((4. / 3.) * p[o + g->nx] - 2.0 * p[o]);
                else if (g->cc && j == g->vhiy) lap += g->facy * ((4. / 3.) * p[o - g->nx] - 2.0 * p[o]);
                else lap += g->facy * (p[o - g->nx] + p[o + g->nx]);
                res[c * n + o] = rhs[c * n + o] + acf[o] * p[o] - lap;
            }
        }
}

/* restrict_cc / restrict_nd (:29-52) */
static void restrict_lvl(const lgeom *gc, double *crse, const double *fine, int fnx, long fcs, int ncomp)
{
    const long ccs = (long)gc->nx * gc->ny;
#pragma omp parallel for schedule(static)
    for (int j = gc->vlo; j <= gc->vhiy; ++j)
        for (int i = gc->vlo; i <= gc->vhix; ++i)
            for (int c = 0; c < ncomp; ++c) {
                const double *f = fine + c * fcs;
#define F(ii, jj) f[(long)(jj) * fnx + (ii)]
                double v;
                if (gc->cc) v = 0.25 * (F(2 * i, 2 * j) + F(2 * i + 1, 2 * j) + F(2 * i, 2 * j + 1) + F(2 * i + 1, 2 * j + 1));
                else v = (1. / 16.) * (F(2 * i - 1, 2 * j - 1) + 2. * F(2 * i, 2 * j - 1) + F(2 * i + 1, 2 * j - 1)
                                       + 2. * F(2 * i - 1, 2 * j) + 4. * F(2 * i, 2 * j) + 2. * F(2 * i + 1, 2 * j)
                                       + F(2 * i - 1, 2 * j + 1) + 2. * F(2 * i, 2 * j + 1) + F(2 * i + 1, 2 * j + 1));
#undef F
                crse[c * ccs + (long)j * gc->nx + i] = v;
            }
}

/* interpcpy_cc / interpcpy_nd (:88-121): out = fine + I(crse) */
static void interp_add(const lgeom *gf, double *out, const double *fine, const double *crse, int cnx, long ccs)
{
    const long fcs = (long)gf->nx * gf->ny;
#pragma omp parallel for schedule(static)
    for (int j = gf->vlo; j <= gf->vhiy; ++j)
        for (int i = gf->vlo; i <= gf->vhix; ++i) {
            const int ic = i >> 1, jc = j >> 1;
            for (int c = 0; c < 2; ++c) {
                const double *cr = crse + c * ccs;
#define Cc(ii, jj) cr[(long)(jj) * cnx + (ii)]
                double add;
                if (gf->cc) add = Cc(ic, jc);
                else {
                    const int io = (ic * 2 != i), jo = (jc * 2 != j);
                    if (io && jo) add = (Cc(ic, jc) + Cc(ic + 1, jc) + Cc(ic, jc + 1) + Cc(ic + 1, jc + 1)) * 0.25;
                    else if (io) add = (Cc(ic, jc) + Cc(ic + 1, jc)) * 0.5;
                    else if (jo) add = (Cc(ic, jc) + Cc(ic, jc + 1)) * 0.5;
                    else add = Cc(ic, jc);
                }
#undef Cc
                const long o = c * fcs + (long)j * gf->nx + i;
                out[o] = fine[o] + add;
            }
        }
}

static double maxabs2(const lgeom *g, const double *a)
{
    const long n = (long)g->nx * g->ny;
    double m = 0.;
#pragma omp parallel for schedule(static) reduction(max : m)
    for (int j = g->vlo; j <= g->vhiy; ++j)
        for (int i = g->vlo; i <= g->vhix; ++i) {
            const long o = (long)j * g->nx + i;
            const double v = fmax(fabs(a[o]), fabs(a[n + o]));
            if (v > m) m = v;
        }
    return m;
}

static void vcycle(hpc_mg *m, double *sol0, const double *rhs0)
{
    const int nl = m->nlev;
    for (int l = 0; l < nl - 1; ++l) {                                      /* :1442-1460 */
        const lgeom g = level_geom(m, l);
        const long n = (long)g.nx * g.ny;
        if (l > 0) {
            memset(m->lev[l].cor, 0, sizeof(double) * 2 * n);
            gsrb(&g, m->lev[l].cor, m->lev[l].res, m->lev[l].acf, 4);
            residual(&g, m->lev[l].rescor, m->lev[l].cor, m->lev[l].res, m->lev[l].acf);
        }
        const lgeom gc = level_geom(m, l + 1);
        restrict_lvl(&gc, m->lev[l + 1].res, m->lev[l].rescor, g.nx, n, 2);
    }
    {                                                                       /* bottom :1514-1594 */
        const int l = nl - 1;
        const lgeom g = level_geom(m, l);
        int nsw = 16;
        const int mx = g.nx > g.ny ? g.nx : g.ny;
        if ((mx + 1) / 2 * 2 > nsw) nsw = (mx + 1) / 2 * 2;
        memset(m->lev[l].cor, 0, sizeof(double) * 2 * (size_t)g.nx * g.ny);
        gsrb(&g, m->lev[l].cor, m->lev[l].res, m->lev[l].acf, nsw);
    }
    for (int l = nl - 2; l >= 0; --l) {                                     /* :1476-1499 */
        const lgeom g = level_geom(m, l);
        const long n = (long)g.nx * g.ny;
        interp_add(&g, m->lev[l].rescor, m->lev[l].cor, m->lev[l + 1].cor, m->lev[l + 1].nx,
                   (long)m->lev[l + 1].nx * m->lev[l + 1].ny);
        if (l == 0) {
            memcpy(sol0, m->lev[0].rescor, sizeof(double) * 2 * n);
            gsrb(&g, sol0, rhs0, m->lev[0].acf, 4);
        } else {
            memcpy(m->lev[l].cor, m->lev[l].rescor, sizeof(double) * 2 * n);
            gsrb(&g, m->lev[l].cor, m->lev[l].res, m->lev[l].acf, 4);
        }
    }
    const lgeom g0 = level_geom(m, 0);                                      /* :1501-1503 */
    memcpy(m->lev[0].cor, sol0, sizeof(double) * 2 * (size_t)g0.nx * g0.ny);
    gsrb(&g0, m->lev[0].cor, rhs0, m->lev[0].acf, 4);
    residual(&g0, m->lev[0].rescor, m->lev[0].cor, rhs0, m->lev[0].acf);
}

/* solve1 (:1169-1190) + solve_doit (:1307-1427).  sol (2 comps), rhs (2 comps): slice
 * components with guard cells (component stride cs = nxt*nyt); acf: chi.  Returns the number of
 * V-cycles, or -1 on failure. */
int hpc_mg_solve1(hpc_mg *m, double *sol, const double *rhs, const double *acf, int nx, int ny,
                  double tol_rel, double tol_abs, int max_iters)
{
    const int nxt = nx + 2 * G;
    const long cs = (long)nxt * (ny + 2 * G);
    const lgeom g0 = level_geom(m, 0);
    const long n0 = (long)g0.nx * g0.ny;
    const int sh = m->cc ? 0 : 1;        /* level index = cell index + sh */
    double *sol0 = (double *)calloc(2 * n0, sizeof(double));
    double *rhs0 = (double *)calloc(2 * n0, sizeof(double));
    memset(m->lev[0].acf, 0, sizeof(double) * n0);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const long o = (long)(j + G) * nxt + (i + G), d = (long)(j + sh) * g0.nx + (i + sh);
            m->lev[0].acf[d] = acf[o];
            sol0[d] = sol[o]; sol0[n0 + d] = sol[cs + o];
            rhs0[d] = rhs[o]; rhs0[n0 + d] = rhs[cs + o];
        }
    for (int l = 1; l < m->nlev; ++l) {                                     /* average_down_acoef */
        const lgeom gc = level_geom(m, l);
        memset(m->lev[l].acf, 0, sizeof(double) * (size_t)gc.nx * gc.ny);
        restrict_lvl(&gc, m->lev[l].acf, m->lev[l - 1].acf, m->lev[l - 1].nx, 0, 1);
    }
    memcpy(m->lev[0].cor, sol0, sizeof(double) * 2 * n0);                   /* :1326-1327 */
    gsrb(&g0, m->lev[0].cor, rhs0, m->lev[0].acf, 4);
    residual(&g0, m->lev[0].rescor, m->lev[0].cor, rhs0, m->lev[0].acf);
    const double resnorm0 = maxabs2(&g0, m->lev[0].rescor), rhsnorm0 = maxabs2(&g0, rhs0);
    const double max_norm = rhsnorm0 > resnorm0 ? rhsnorm0 : resnorm0;
    const double res_target = fmax(tol_abs, fmax(tol_rel, 1.e-16) * max_norm);   /* :1361 */
    int iters = 0, ok = 1;
    if (resnorm0 > res_target) {
        ok = 0;
        for (int it = 0; it < max_iters; ++it) {
            vcycle(m, sol0, rhs0);
            iters = it + 1;
            const double norminf = maxabs2(&g0, m->lev[0].rescor);
            if (norminf <= res_target) { ok = 1; break; }
            if (norminf > 1.e20 * max_norm) break;
        }
    }
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)                                            /* :1419-1426 */
        for (int i = 0; i < nx; ++i) {
            const long o = (long)(j + G) * nxt + (i + G), d = (long)(j + sh) * g0.nx + (i + sh);
            sol[o] = m->lev[0].cor[d]; sol[cs + o] = m->lev[0].cor[n0 + d];
        }
    free(sol0); free(rhs0);
    return ok ? iters : -1;
}
