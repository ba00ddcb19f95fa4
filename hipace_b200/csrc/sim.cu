// Slice-loop driver: Hipace::Hipace / InitData / Evolve / SolveOneSlice restated over the
// hpb_* kernel seams (src/Hipace.cpp:74-295 constructor + InitData, :393-554 Evolve,
// :556-728 SolveOneSlice, explicit-solver branch, level 0).
//
// Everything lives on the device between hpb_sim_create() and the host-buffer queries; the host
// only enqueues kernels.  Plasma is re-created every time step (Hipace.cpp:450) by a device
// kernel that evaluates the deck's density expression; fixed_ppc beams are created for all
// slices at once (the reference creates them slice by slice on the head rank at step 0,
// src/particles/beam/BeamParticleContainerInit.cpp:198-346 -- same particles, same order, same
// ids because our storage order is the head-first processing order).
#include "sim.hpp"
#include <stdlib.h>
#include <cub/cub.cuh>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <algorithm>
#include <memory>

int hpb_advance_beam_impl(hpb_ctx *ctx, hpb_beam_slice bm, int *d_nsub, hpb_slice sl, double charge,
                          double mass, int n_subcycles, double dt, double time, double min_z,
                          int do_z_push, int particle_bc, const double bc_lo[2], const double bc_hi[2],
                          const int *comps, const hpb_extfields *ext, int *d_class_counts,
                          double *d_checksum, unsigned long long *d_n_pushed);   // beam.cu

using hpb::Deck;
using hpb::RpnInstr;

namespace {

// ---- plasma: PlasmaParticleContainer::InitParticles (PlasmaParticleContainerInit.cpp:17-316) ----
struct PlasmaInitArgs {
    int ilo, jlo, ncx, ncy;       // candidate cell box
    int ppcx, ppcy;
    double plo_x, plo_y, dx, dy;
    double blo_x, blo_y, bhi_x, bhi_y;
    double radius_sq, hollow_sq, min_density, c_t, scale;
    double ux0, uy0, psi0;
};

__device__ __forceinline__ bool plasma_candidate(const PlasmaInitArgs &a, const DevRpn &dens,
                                                 long idx, double &x, double &y, double &d)
{
    const long ncell = (long)a.ncx * a.ncy;
    const int i_part = (int)(idx / ncell);
    const long cell = idx - (long)i_part * ncell;
    const int j = (int)(cell / a.ncx) + a.jlo, i = (int)(cell % a.ncx) + a.ilo;
    const double rx = (0.5 + (i_part % a.ppcx)) / a.ppcx;       // ParticleUtil.H:72-80
    const double ry = (0.5 + (i_part / a.ppcx)) / a.ppcy;
    // separate multiply and add (no FMA contraction): positions are bit-identical to a CPU build
    x = __dadd_rn(a.plo_x, __dmul_rn(i + rx, a.dx));
    y = __dadd_rn(a.plo_y, __dmul_rn(j + ry, a.dy));
    const double rsq = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
    if (x >= a.bhi_x || x < a.blo_x || y >= a.bhi_y || y < a.blo_y || rsq > a.radius_sq ||
        rsq < a.hollow_sq)
        return false;
    d = rpn_eval(dens, x, y, a.c_t);
    return !(d <= a.min_density);
}

__global__ void k_plasma_flag(PlasmaInitArgs a, DevRpn dens, long ncand, unsigned *flag)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncand) return;
    double x, y, d;
    flag[idx] = plasma_candidate(a, dens, idx, x, y, d) ? 1u : 0u;
}

__global__ void k_plasma_fill(PlasmaInitArgs a, DevRpn dens, long ncand, const unsigned *flag,
                              const unsigned *offs, hpb_plasma pl)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncand || !flag[idx]) return;
    double x, y, d;
    plasma_candidate(a, dens, idx, x, y, d);
    const long p = offs[idx];
    pl.r[HPB_X][p] = x;
    pl.r[HPB_Y][p] = y;
    pl.r[HPB_W][p] = d * a.scale;
    pl.r[HPB_UX][p] = a.ux0;
    pl.r[HPB_UY][p] = a.uy0;
    pl.r[HPB_PSI][p] = a.psi0;
    pl.r[HPB_X_PREV][p] = x;
    pl.r[HPB_Y_PREV][p] = y;
    pl.r[HPB_UX_HALF][p] = a.ux0;
    pl.r[HPB_UY_HALF][p] = a.uy0;
    pl.r[HPB_PSI_HALF][p] = a.psi0;
    pl.idcpu[p] = hpb_make_idcpu(1, 0);       // id 1, cpu (= MR level) 0, :283-284
}

// ---- beam: InitBeamFixedPPCSlice for all slices (BeamParticleContainerInit.cpp:198-346) ----
struct BeamInitArgs {
    int ilo, jlo, ncx, ncy, nz;
    int ppcx, ppcy, ppcz;
    double plo_x, plo_y, plo_z, dx, dy, dz;
    double x_mean, y_mean, z_mean, sx, sy, sz;
    double zmin, zmax, radius_sq, density, min_density, scale;
    int profile;                  // 0 gaussian, 1 flattop
    double ux0, uy0, uz0;
};

__device__ __forceinline__ bool beam_candidate(const BeamInitArgs &a, long idx, double &x,
                                               double &y, double &z, double &d)
{
    const int nppc = a.ppcx * a.ppcy * a.ppcz;
    const long per_slice = (long)a.ncx * a.ncy * nppc;
    const int s = (int)(idx / per_slice);            // storage slot: head slice first
    const int islice = a.nz - 1 - s;
    long rem = idx - (long)s * per_slice;
    const long cell = rem / nppc;
    const int i_part = (int)(rem - cell * nppc);
    const int j = (int)(cell / a.ncx) + a.jlo, i = (int)(cell % a.ncx) + a.ilo;
    const int pyz = a.ppcy * a.ppcz;
    const int ix_p = i_part / pyz, iy_p = (i_part % pyz) % a.ppcy, iz_p = (i_part % pyz) / a.ppcy;
    x = __dadd_rn(a.plo_x, __dmul_rn(i + (0.5 + ix_p) / a.ppcx, a.dx));
    y = __dadd_rn(a.plo_y, __dmul_rn(j + (0.5 + iy_p) / a.ppcy, a.dy));
    z = __dadd_rn(a.plo_z, __dmul_rn(islice + (0.5 + iz_p) / a.ppcz, a.dz));
    const double xr = x - a.x_mean, yr = y - a.y_mean;
    if (z >= a.zmax || z < a.zmin || __dadd_rn(__dmul_rn(xr, xr), __dmul_rn(yr, yr)) > a.radius_sq)
        return false;
    if (a.profile == 0) {                            // GetInitialDensity.H:33-51
        const double dxn = (x - a.x_mean) / a.sx, dyn = (y - a.y_mean) / a.sy,
                     dzn = (z - a.z_mean) / a.sz;
        d = a.density * exp(-0.5 * dxn * dxn) * exp(-0.5 * dyn * dyn) * exp(-0.5 * dzn * dzn);
    } else {
        d = a.density;
    }
    return !(d <= a.min_density);
}

__global__ void k_beam_flag(BeamInitArgs a, long ncand, unsigned *flag)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncand) return;
    double x, y, z, d;
    flag[idx] = beam_candidate(a, idx, x, y, z, d) ? 1u : 0u;
}

__global__ void k_beam_fill(BeamInitArgs a, long ncand, const unsigned *flag, const unsigned *offs,
                            const long *slot_off, BeamRing ring, uint64_t first_id)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncand || !flag[idx]) return;
    double x, y, z, d;
    beam_candidate(a, idx, x, y, z, d);
    const long per_slice = (long)a.ncx * a.ncy * (a.ppcx * a.ppcy * a.ppcz);
    const int slot = (int)(idx / per_slice);
    const long gp = offs[idx];                 // global particle index = id - first_id
    const long p = gp - slot_off[slot];
    if (p >= ring.cap) return;
    const hpb_beam_slice b = ring.view(slot);
    b.x[p] = x; b.y[p] = y; b.z[p] = z;
    b.w[p] = fabs(d * a.scale);
    b.ux[p] = a.ux0; b.uy[p] = a.uy0; b.uz[p] = a.uz0;
    b.idcpu[p] = hpb_make_idcpu(first_id + (uint64_t)gp, 0);
}

__global__ void k_ring_set_counts(BeamRing ring, const long *slot_off)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ring.nslots) return;
    int64_t *h = ring.hdr(s);
    const long n = slot_off[s + 1] - slot_off[s];
    h[0] = n; h[1] = n;
    for (int k = 2; k < 8; ++k) h[k] = 0;
}

// offsets of the slice slots: slot_off[s] = offs[s * per_slice], slot_off[nslots] = total
__global__ void k_slot_offsets(const unsigned *flag, const unsigned *offs, long per_slice,
                               int nslots, long ncand, long *slot_off)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > nslots) return;
    if (s == nslots) slot_off[s] = ncand > 0 ? (long)offs[ncand - 1] + flag[ncand - 1] : 0;
    else slot_off[s] = offs[(long)s * per_slice];
}

__global__ void k_count_valid(const uint64_t *idcpu, long np, unsigned long long *out)
{
    const long ip = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool v = ip < np && hpb_is_valid(idcpu[ip]);
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

__global__ void k_valid_bytes(const uint64_t *idcpu, long np, uint8_t *out)
{
    const long ip = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ip < np) out[ip] = hpb_is_valid(idcpu[ip]) ? 1 : 0;
}

void species_charge_mass(const Deck &dk, const hpb_geom &g, const std::string &pre,
                         const std::string &def_el, double &charge, double &mass)
{
    const double m_p = g.normalized ? 1836.15267343 : 1.67262192369e-27;
    const std::string el = dk.str(pre + ".element", def_el);
    charge = mass = 0.;
    if (el == "electron") { charge = -g.q_e; mass = g.m_e; }
    else if (el == "positron") { charge = g.q_e; mass = g.m_e; }
    else if (el == "proton") { charge = g.q_e; mass = m_p; }
    else if (!el.empty()) throw std::runtime_error("unsupported element '" + el + "'");
    charge = dk.num(pre + ".charge", charge);
    mass = dk.num(pre + ".mass", mass);
}

inline unsigned nb256(long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

namespace {

int ensure_scan(hpb_sim *s, long n)
{
    if (n > 0x7fffffffL) { hpb_set_error("init: too many candidate particles"); return HPB_ERR_UNSUPPORTED; }
    if (n <= s->scan_cap) return HPB_OK;
    cudaFree(s->d_flag); cudaFree(s->d_offs); cudaFree(s->d_cub);
    s->d_flag = s->d_offs = nullptr; s->d_cub = nullptr;
    SIM_CUDA(cudaMalloc(&s->d_flag, sizeof(unsigned) * n));
    SIM_CUDA(cudaMalloc(&s->d_offs, sizeof(unsigned) * n));
    s->cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, s->cub_bytes, s->d_flag, s->d_offs, (int)n, s->stream);
    SIM_CUDA(cudaMalloc(&s->d_cub, s->cub_bytes));
    s->scan_cap = n;
    return HPB_OK;
}

int read_deck(hpb_sim *s)
{
    const Deck &d = s->deck;
    hpb_geom &g = s->g;
    g.normalized = (int)d.num("hipace.normalized_units", 0);
    if (g.normalized) { g.c = g.ep0 = g.mu0 = g.q_e = g.m_e = 1.0; }       // Constants.H:54-81
    else { g.c = 299792458.; g.ep0 = 8.8541878128e-12; g.mu0 = 1.25663706212e-06;
           g.q_e = 1.602176634e-19; g.m_e = 9.1093837015e-31; }
    const auto nc = d.nums("amr.n_cell", {});
    const auto lo = d.nums("geometry.prob_lo", {}), hi = d.nums("geometry.prob_hi", {});
    if (nc.size() != 3 || lo.size() != 3 || hi.size() != 3)
        throw std::runtime_error("amr.n_cell / geometry.prob_lo / geometry.prob_hi need 3 values");
    g.nx = (int)nc[0]; g.ny = (int)nc[1]; s->nz = (int)nc[2];
    for (int k = 0; k < 3; ++k) { s->prob_lo[k] = lo[k]; s->prob_hi[k] = hi[k]; }
    g.dx = (hi[0] - lo[0]) / g.nx; g.dy = (hi[1] - lo[1]) / g.ny; g.dz = (hi[2] - lo[2]) / s->nz;
    // GetPosOffset with the grown fab box [-g, n-1+g] (Fields.H:71-77)
    s->depos_order = (int)d.num("hipace.depos_order_xy", 2);
    s->depos_dtype = (int)d.num("hipace.depos_derivative_type", 2);
    if (s->depos_order < 0 || s->depos_order > 3 || s->depos_dtype < 0 || s->depos_dtype > 2)
        throw std::runtime_error("hipace.depos_order_xy must be 0..3 and hipace.depos_derivative_type 0..2");
    if (s->depos_order == 0 && s->depos_dtype == 0)                        // Hipace.cpp:52-53
        throw std::runtime_error("Analytic derivative with depos_order=0 would vanish");
    s->ng = HPB_NGUARD_OF(s->depos_order);
    g.x_off = 0.5 * (lo[0] + hi[0] - g.dx * ((-s->ng) + (g.nx - 1 + s->ng)));
    g.y_off = 0.5 * (lo[1] + hi[1] - g.dy * ((-s->ng) + (g.ny - 1 + s->ng)));
    if ((int)d.num("grid_current.use_grid_current", 0)) {                  // GridCurrent.cpp:14-23
        const auto m = d.nums("grid_current.position_mean", {}), sd = d.nums("grid_current.position_std", {});
        if (m.size() != 3 || sd.size() != 3 || !d.has("grid_current.peak_current_density"))
            throw std::runtime_error("grid_current needs peak_current_density, position_mean and position_std");
        s->use_grid_current = true;
        s->gc_peak = d.num("grid_current.peak_current_density", 0.);
        for (int k = 0; k < 3; ++k) { s->gc_mean[k] = m[k]; s->gc_std[k] = sd[k]; }
    }
    const std::string solver = d.str("hipace.bxby_solver", "explicit");
    if (solver != "explicit" && solver != "predictor-corrector")
        throw std::runtime_error("hipace.bxby_solver must be explicit or predictor-corrector");
    s->explicit_solver = solver == "explicit";
    s->predcorr_tol = d.num("hipace.predcorr_B_error_tolerance", 4e-2);
    s->predcorr_max_iter = (int)d.num("hipace.predcorr_max_iterations", 30);
    s->predcorr_mix = d.num("hipace.predcorr_B_mixing_factor", 0.05);
    const std::string bfield = d.str("boundary.field", "");
    if (bfield != "Dirichlet" && bfield != "Open" && bfield != "Periodic")
        throw std::runtime_error("boundary.field must be Dirichlet, Periodic or Open");        // Hipace.cpp:190-200
    s->open_bc = bfield == "Open";
    s->field_periodic = bfield == "Periodic";
    // Fields.cpp:34-40, :179-208: the GPU default is FFTDirichletFast; FFTDirichletDirect / Expanded are the
    // same equation through other transforms (one solver here); MGDirichlet is not available
    const std::string psolver = d.str("fields.poisson_solver", "FFTDirichletFast");
    if (psolver == "FFTPeriodic") s->poisson_periodic = true;
    else if (psolver != "FFTDirichletFast" && psolver != "FFTDirichletDirect" && psolver != "FFTDirichletExpanded")
        throw std::runtime_error("fields.poisson_solver must be FFTDirichletFast, FFTDirichletDirect, "
                                 "FFTDirichletExpanded or FFTPeriodic (MGDirichlet is not supported)");
    if ((s->field_periodic || s->poisson_periodic) && !s->explicit_solver)
        throw std::runtime_error("boundary.field = Periodic / fields.poisson_solver = FFTPeriodic are only "
                                 "available with hipace.bxby_solver = explicit");
    if (s->open_bc && s->poisson_periodic)
        throw std::runtime_error("boundary.field = Open needs a Dirichlet Poisson solver");
    if (s->open_bc && s->explicit_solver)
        throw std::runtime_error("boundary.field = Open is only available with hipace.bxby_solver = predictor-corrector");
    const std::string pbc = d.str("boundary.particle", "");
    if (pbc == "Periodic") s->particle_bc = HPB_BC_PERIODIC;
    else if (pbc == "Reflecting") s->particle_bc = HPB_BC_REFLECTING;
    else if (pbc == "Absorbing") s->particle_bc = HPB_BC_ABSORBING;
    else throw std::runtime_error("boundary.particle must be Periodic, Reflecting or Absorbing");
    const auto blo = d.nums("boundary.particle_lo", {lo[0], lo[1]});
    const auto bhi = d.nums("boundary.particle_hi", {hi[0], hi[1]});
    s->bc_lo[0] = blo[0]; s->bc_lo[1] = blo[1]; s->bc_hi[0] = bhi[0]; s->bc_hi[1] = bhi[1];
    s->max_step = (int)d.num("max_step", 0);
    s->adaptive_dt = d.str("hipace.dt", "") == "adaptive";            // AdaptiveTimeStep.cpp:17-35
    s->dt = s->adaptive_dt ? 0. : d.num("hipace.dt", 0.);
    if (s->adaptive_dt) {
        s->adp.nt_per_betatron = d.num("hipace.nt_per_betatron", 20.);
        s->adp.dt_max = d.num("hipace.dt_max", INFINITY);
        s->adp.threshold_uz = d.num("hipace.adaptive_threshold_uz", 2.);
        s->adp.phase_tolerance = d.num("hipace.adaptive_phase_tolerance", 4e-4);
        s->adp.phase_substeps = (int)d.num("hipace.adaptive_phase_substeps", 2000);
        s->adp.control_phase = (int)d.num("hipace.adaptive_control_phase_advance", 1);
        s->adp.predict_step = (int)d.num("hipace.adaptive_predict_step", 1);
        s->adp.c = g.c; s->adp.ep0 = g.ep0;
        s->adaptive_density = d.num("plasmas.adaptive_density", 0.);
        if ((int)d.num("hipace.adaptive_gather_ez", 0))
            throw std::runtime_error("hipace.adaptive_gather_ez = 1 is not supported (the reference calls it buggy)");
    }
    s->mg_tol_rel = d.num("hipace.MG_tolerance_rel", 1e-4);
    s->mg_tol_abs = d.num("hipace.MG_tolerance_abs", DBL_MIN);
    s->deposit_rho = (int)d.num("hipace.deposit_rho", 0) != 0;
    for (auto &t : d.strs("diagnostic.field_data")) if (t == "rho") s->deposit_rho = true;
    {
        const std::string dt_ = d.str("diagnostic.diag_type", "xyz");
        if (dt_ != "xyz" && dt_ != "xz")
            throw std::runtime_error("diagnostic.diag_type must be xyz or xz (the checksums follow it)");
        s->diag_xz = dt_ == "xz";
    }
    s->field_insitu_period = (int)d.num("fields.insitu_period", 0);
    s->field_insitu_prefix = d.str("fields.insitu_file_prefix", "diags/field_insitu");
    if (s->field_insitu_period > 0 && !s->explicit_solver)                     // Fields.cpp:1311-1312
        throw std::runtime_error("Must use explicit solver for field insitu diagnostic");
    s->do_beam_jx_jy = (int)d.num("hipace.do_beam_jx_jy_deposition", 1) != 0;

    auto pn = d.strs("plasmas.names");
    if (!pn.empty() && pn[0] != "no_plasma") {
        for (auto &nm : pn) {
            Species sp;
            sp.name = nm;
            species_charge_mass(d, g, nm, "", sp.charge, sp.mass);
            std::string expr = d.str(nm + ".density(x,y,z)", "", "plasmas.density(x,y,z)");
            if (expr.empty()) expr = "0.";
            else {      // the value may contain blanks: re-join the tokens
                auto v = d.find(nm + ".density(x,y,z)", "plasmas.density(x,y,z)");
                expr.clear();
                for (auto &t : *v) expr += t;
            }
            std::vector<RpnInstr> code;
            d.compile(expr, {"x", "y", "z"}, code);
            { const double origin[3] = {0., 0., 0.}; sp.density0 = d.run(code, origin); }
            sp.density_host = code;
            sp.density.n = (int)code.size();
            for (size_t k = 0; k < code.size(); ++k) sp.density.code[k] = code[k];
            const auto ppc = d.nums(nm + ".ppc", {}, "plasmas.ppc");
            if (ppc.size() != 2) throw std::runtime_error(nm + ".ppc needs 2 values");
            sp.ppc[0] = (int)ppc[0]; sp.ppc[1] = (int)ppc[1];
            sp.neutralize = (int)d.num(nm + ".neutralize_background", 1, "plasmas.neutralize_background") != 0;
            sp.max_qsa = d.num(nm + ".max_qsa_weighting_factor", 35., "plasmas.max_qsa_weighting_factor");
            sp.n_subcycles = (int)d.num(nm + ".n_subcycles", 1, "plasmas.n_subcycles");
            sp.reorder_period = (int)d.num(nm + ".reorder_period", 0, "plasmas.reorder_period");
            {
                const auto it = d.nums(nm + ".reorder_idx_type", {0., 0.}, "plasmas.reorder_idx_type");
                if (it.size() != 2) throw std::runtime_error("reorder_idx_type needs 2 values");
                // 0: cell, 1: node, 2: both (two sorts in the reference; binned by cell here)
                sp.reorder_idx[0] = (int)it[0] == 1; sp.reorder_idx[1] = (int)it[1] == 1;
            }
            sp.insitu_period = (int)d.num(nm + ".insitu_period", 0, "plasmas.insitu_period");
            sp.insitu_radius = d.num(nm + ".insitu_radius", INFINITY, "plasmas.insitu_radius");
            sp.insitu_file_prefix = d.str(nm + ".insitu_file_prefix", "diags/plasma_insitu", "plasmas.insitu_file_prefix");
            sp.radius = d.num(nm + ".radius", INFINITY, "plasmas.radius");
            sp.hollow = d.num(nm + ".hollow_core_radius", 0., "plasmas.hollow_core_radius");
            sp.min_density = d.num(nm + ".min_density", 0., "plasmas.min_density");
            const auto um = d.nums(nm + ".u_mean", {0., 0., 0.});
            for (int k = 0; k < 3 && k < (int)um.size(); ++k) sp.u_mean[k] = um[k];
            const auto us = d.nums(nm + ".u_std", {0., 0., 0.});
            for (double v : us)
                if (v != 0.) throw std::runtime_error("plasma u_std != 0 needs AMReX's RNG stream: unsupported");
            s->plasmas.push_back(sp);
            s->any_neutral = s->any_neutral || sp.neutralize;
        }
    }
    // lasers (laser/MultiLaser.cpp:26-56, laser/Laser.cpp:18-47): gaussian envelopes on the field grid
    auto ln = d.strs("lasers.names");
    if (!ln.empty() && ln[0] != "no_laser") {
        s->use_laser = true;
        if (s->field_periodic || s->poisson_periodic)
            throw std::runtime_error("lasers with boundary.field = Periodic / fields.poisson_solver = FFTPeriodic are not supported");
        s->laser_lambda0 = d.num("lasers.lambda0", 0.);
        s->laser_interp_order = (int)d.num("lasers.interp_order", 1);
        if (s->laser_lambda0 <= 0.) throw std::runtime_error("lasers.lambda0 must be given");
        for (const char *k : {"lasers.n_cell", "lasers.patch_lo", "lasers.patch_hi"})
            if (d.find(k)) throw std::runtime_error(std::string(k) + ": only the default laser grid (= field grid) is supported");
        {
            const std::string st = d.str("lasers.solver_type", "multigrid");
            if (st != "fft" && st != "multigrid") throw std::runtime_error("lasers.solver_type must be fft or multigrid");
            s->laser_use_mg = st == "multigrid";
            s->laser_mg_tol_rel = d.num("lasers.MG_tolerance_rel", 1e-4);
            s->laser_mg_tol_abs = d.num("lasers.MG_tolerance_abs", 0.);
            s->laser_mg_avg_rhs = (int)d.num("lasers.MG_average_rhs", 1) != 0;
        }
        if (s->max_step > 0 && s->adaptive_dt)
            throw std::runtime_error("lasers cannot be combined with an adaptive time step");       // Hipace.cpp:407-409
        s->laser_use_phase = (int)d.num("lasers.use_phase", 1) != 0;
        s->laser_insitu_period = (int)d.num("lasers.insitu_period", 0);
        s->laser_insitu_prefix = d.str("lasers.insitu_file_prefix", "diags/laser_insitu");
        if (s->laser_insitu_period > 0 && !(s->max_step > 0 && s->dt != 0.))
            throw std::runtime_error("lasers.insitu_period needs the stored envelope (max_step > 0, dt != 0)");
        if (!s->explicit_solver)
            throw std::runtime_error("lasers are only supported with hipace.bxby_solver = explicit");
        if (s->diag_xz && !(s->max_step > 0 && s->dt != 0.))
            throw std::runtime_error("lasers: diagnostic.diag_type = xz needs the stored envelope (max_step > 0, dt != 0)");
        if ((int)ln.size() > HPB_MAX_LASERS) throw std::runtime_error("too many lasers");
        for (auto &nm : ln) {
            if (d.str(nm + ".init_type", "gaussian") != "gaussian")
                throw std::runtime_error("laser init_type must be gaussian");
            hpb_laser L = {};
            L.a0 = d.num(nm + ".a0", 0.); L.w0 = d.num(nm + ".w0", 0.); L.cep = d.num(nm + ".CEP", 0.);
            L.propagation_angle_yz = d.num(nm + ".propagation_angle_yz", 0.);
            L.pft_yz = d.num(nm + ".PFT_yz", 1.5707963267948966);
            const bool has_L0 = d.find(nm + ".L0") != nullptr, has_tau = d.find(nm + ".tau") != nullptr;
            if (has_L0 == has_tau)                                  // Laser.cpp:38-41
                throw std::runtime_error("specify exclusively either L0 or tau of laser " + nm);
            L.L0 = has_L0 ? d.num(nm + ".L0", 0.) : d.num(nm + ".tau", 0.) * g.c;
            L.focal_distance = d.num(nm + ".focal_distance", 0.);
            const auto pm = d.nums(nm + ".position_mean", {0., 0., 0.});
            for (int k = 0; k < 3 && k < (int)pm.size(); ++k) L.position_mean[k] = pm[k];
            s->lasers.push_back(L);
        }
    }
    auto bn = d.strs("beams.names");
    if (!bn.empty() && bn[0] != "no_beam") {
        for (auto &nm : bn) {
            BeamSp b;
            b.name = nm;
            if (d.str(nm + ".injection_type", "") != "fixed_ppc")
                throw std::runtime_error("only beam injection_type = fixed_ppc is supported (others need AMReX's RNG)");
            species_charge_mass(d, g, nm, "electron", b.charge, b.mass);
            const auto ppc = d.nums(nm + ".ppc", {1., 1., 1.});
            for (int k = 0; k < 3 && k < (int)ppc.size(); ++k) b.ppc[k] = (int)ppc[k];
            const std::string prof = d.str(nm + ".profile", "");
            if (prof == "gaussian") b.profile = 0;
            else if (prof == "flattop") b.profile = 1;
            else throw std::runtime_error("beam profile must be gaussian or flattop");
            b.density = fabs(d.num(nm + ".density", 0.));
            b.zmin = d.num(nm + ".zmin", -INFINITY); b.zmax = d.num(nm + ".zmax", INFINITY);
            b.radius = d.num(nm + ".radius", INFINITY);
            b.min_density = fabs(d.num(nm + ".min_density", 0.));
            const auto pm = d.nums(nm + ".position_mean", {0., 0., 0.});
            const auto ps = d.nums(nm + ".position_std", {0., 0., 0.});
            const auto um = d.nums(nm + ".u_mean", {0., 0., 0.});
            for (int k = 0; k < 3; ++k) {
                if (k < (int)pm.size()) b.pos_mean[k] = pm[k];
                if (k < (int)ps.size()) b.pos_std[k] = ps[k];
                if (k < (int)um.size()) b.u_mean[k] = um[k];
            }
            const auto us = d.nums(nm + ".u_std", {0., 0., 0.});
            for (double v : us)
                if (v != 0.) throw std::runtime_error("beam u_std != 0 needs AMReX's RNG stream: unsupported");
            b.u_mean_z = um.size() > 2 ? um[2] : 0.;
            b.n_subcycles = (int)d.num(nm + ".n_subcycles", 10);
            if (b.n_subcycles < 1) throw std::runtime_error(nm + ".n_subcycles must be >= 1");
            b.do_z_push = (int)d.num(nm + ".do_z_push", 1) != 0;
            (void)d.has(nm + ".slice_capacity");        // our addition, read when the rings are allocated (init_beam)
            // "<beam name> or beams" (queryWithParserAlt)
            auto alt_num = [&](const char *key, double dflt) {
                return d.has(nm + "." + key) ? d.num(nm + "." + key, dflt) : d.num(std::string("beams.") + key, dflt);
            };
            b.insitu_period = (int)alt_num("insitu_period", 0);
            b.insitu_radius = alt_num("insitu_radius", INFINITY);
            b.insitu_file_prefix = d.has(nm + ".insitu_file_prefix") ? d.str(nm + ".insitu_file_prefix", "")
                                 : d.str("beams.insitu_file_prefix", "diags/insitu");
            if ((int)d.num(nm + ".do_radiation_reaction", 0) != 0 || (int)d.num(nm + ".do_spin_tracking", 0) != 0)
                throw std::runtime_error("beam radiation reaction / spin tracking are not supported");
            // external_E(x,y,z,t) / external_B(x,y,z,t): beam name first, then "beams"
            const char *keys[2] = {"external_E(x,y,z,t)", "external_B(x,y,z,t)"};
            for (int k = 0; k < 2; ++k) {
                auto v = d.find(nm + "." + keys[k], std::string("beams.") + keys[k]);
                for (int c = 0; c < 3; ++c) b.ext_expr[3 * k + c] = "0.";
                if (!v) continue;
                if (v->size() != 3) throw std::runtime_error(std::string(keys[k]) + " needs 3 expressions");
                for (int c = 0; c < 3; ++c) b.ext_expr[3 * k + c] = (*v)[c];
                b.use_ext = true;
            }
            s->beams.push_back(b);
        }
    }
    // options whose default is what this implementation does: accepted at that value only
    if ((int)d.num("amr.max_level", 0) != 0) throw std::runtime_error("amr.max_level > 0 (mesh refinement) is not supported");
    if ((int)d.num("hipace.do_beam_jz_minus_rho", 0) != 0) throw std::runtime_error("hipace.do_beam_jz_minus_rho is not supported");
    if (d.has("hipace.max_time")) throw std::runtime_error("hipace.max_time is not supported (use max_step)");
    // Anything the deck sets that was not read above is an option of the reference this
    // implementation does not have (ionisation, collisions, mesh refinement, SALAME, ...): refuse
    // instead of running different physics.  Output-only and runtime-tuning groups are harmless.
    const auto left = d.unused({"my_constants.", "amrex.", "diagnostic.", "hipace.verbose", "hipace.tile_size",
                                "hipace.do_tiling", "hipace.output_period", "hipace.file_prefix",
                                "hipace.openpmd_backend", "hipace.comms_buffer_", "hipace.do_device_synchronize",
                                "hipace.m_numprocs", "hipace.numprocs_", "hipace.MG_verbose",
                                "hipace.use_small_dst", "hipace.do_shared_depos"});
    // (blocks of names the deck does not activate -- "beam.*" next to beams.names = no_beam --
    // are inert in the reference as well)
    std::set<std::string> groups = {"hipace", "amr", "geometry", "boundary", "plasmas", "beams", "lasers",
                                    "grid_current", "max_step", "fields"};
    for (auto &sp : s->plasmas) groups.insert(sp.name);
    for (auto &b : s->beams) groups.insert(b.name);
    for (auto &nm : d.strs("lasers.names")) groups.insert(nm);
    std::string msg;
    for (auto &k : left)
        if (groups.count(k.substr(0, k.find('.')))) msg += " " + k;
    if (!msg.empty()) throw std::runtime_error("unsupported or unknown input parameter(s):" + msg);
    return HPB_OK;
}

// component table, allocation order of Fields::AllocData (Fields.cpp:70-122), explicit solver
void build_components(hpb_sim *s)
{
    for (int k = 0; k < HPB_C_COUNT; ++k) s->comps[k] = -1;
    int n = 0;
    auto add = [&](int id, const char *which, const char *name) {
        s->comps[id] = n++;
        s->comp_names.emplace_back(which, name);
        s->comp_id.push_back(id);
    };
    if (!s->explicit_solver) {          // Fields.cpp:124-163
        add(HPB_C_NEXT_JX, "Next", "jx"); add(HPB_C_NEXT_JY, "Next", "jy");
        add(HPB_C_EXMBY, "This", "ExmBy"); add(HPB_C_EYPBX, "This", "EypBx"); add(HPB_C_EZ, "This", "Ez");
        add(HPB_C_BX, "This", "Bx"); add(HPB_C_BY, "This", "By"); add(HPB_C_BZ, "This", "Bz");
        add(HPB_C_PSI, "This", "Psi"); add(HPB_C_JX, "This", "jx"); add(HPB_C_JY, "This", "jy");
        add(HPB_C_JZ, "This", "jz"); add(HPB_C_RHOMJZ, "This", "rhomjz");
        if (s->deposit_rho) add(HPB_C_RHO, "This", "rho");
        add(HPB_C_PREV_BX, "Previous", "Bx"); add(HPB_C_PREV_BY, "Previous", "By");
        add(HPB_C_PREV_JX, "Previous", "jx"); add(HPB_C_PREV_JY, "Previous", "jy");
        if (s->any_neutral) add(HPB_C_IONS_RHOMJZ, "RhomJzIons", "rhomjz");
        add(HPB_C_PCITER_BX, "PCIter", "Bx"); add(HPB_C_PCITER_BY, "PCIter", "By");
        add(HPB_C_PCPREV_BX, "PCPrevIter", "Bx"); add(HPB_C_PCPREV_BY, "PCPrevIter", "By");
        s->sl.ncomp = n;
        for (int k = 0; k < HPB_C_COUNT; ++k) s->comps0[k] = s->comps[k];
        return;
    }
    add(HPB_C_NEXT_JX_BEAM, "Next", "jx_beam"); add(HPB_C_NEXT_JY_BEAM, "Next", "jy_beam");
    add(HPB_C_CHI, "This", "chi"); add(HPB_C_SY, "This", "Sy"); add(HPB_C_SX, "This", "Sx");
    add(HPB_C_EXMBY, "This", "ExmBy"); add(HPB_C_EYPBX, "This", "EypBx"); add(HPB_C_EZ, "This", "Ez");
    add(HPB_C_BX, "This", "Bx"); add(HPB_C_BY, "This", "By"); add(HPB_C_BZ, "This", "Bz");
    add(HPB_C_PSI, "This", "Psi"); add(HPB_C_JX_BEAM, "This", "jx_beam");
    add(HPB_C_JY_BEAM, "This", "jy_beam"); add(HPB_C_JZ_BEAM, "This", "jz_beam");
    add(HPB_C_JX, "This", "jx"); add(HPB_C_JY, "This", "jy"); add(HPB_C_RHOMJZ, "This", "rhomjz");
    if (s->deposit_rho) add(HPB_C_RHO, "This", "rho");
    if (s->use_laser) add(HPB_C_AABS, "This", "aabs");
    add(HPB_C_PREV_JX_BEAM, "Previous", "jx_beam"); add(HPB_C_PREV_JY_BEAM, "Previous", "jy_beam");
    if (s->any_neutral) add(HPB_C_IONS_RHOMJZ, "RhomJzIons", "rhomjz");
    s->sl.ncomp = n;
    for (int k = 0; k < HPB_C_COUNT; ++k) s->comps0[k] = s->comps[k];
}

// physical plane of a logical component (they differ once the beam-current planes have rotated)
inline int phys_comp(const hpb_sim *s, int logical) { return s->comps[s->comp_id[logical]]; }

// sum|Q| of every This-slice component of the finished slice into the step's checksums
// (tests/checksum/backend/openpmd_backend.py:40-45, one slice at a time): one launch for all components
int slice_checksums(hpb_sim *s)
{
    int comps[64], slots[64], n = 0;
    for (int c = 0; c < s->sl.ncomp && n < 64; ++c)
        if (s->comp_names[c].first == "This") { comps[n] = phys_comp(s, c); slots[n] = c; ++n; }
    if (!s->diag_xz) return hpb_abs_sum_multi(s->ctx, s->sl, comps, slots, n, s->d_checksum);
    for (int k = 0; k < n; ++k)
        if (int rc = hpb_abs_sum_xz(s->ctx, s->sl, comps[k], s->d_checksum + slots[k])) return rc;
    return HPB_OK;
}

int init_plasma(hpb_sim *s, Species &sp, double c_t)
{
    const hpb_geom &g = s->g;
    PlasmaInitArgs a;
    int ilo = 0, ihi = g.nx - 1, jlo = 0, jhi = g.ny - 1;
    if (std::isfinite(sp.radius)) {      // :70-82
        ilo = std::max(ilo, (int)lround((-sp.radius - s->prob_lo[0]) / g.dx - 2));
        jlo = std::max(jlo, (int)lround((-sp.radius - s->prob_lo[1]) / g.dy - 2));
        ihi = std::min(ihi, (int)lround((sp.radius - s->prob_lo[0]) / g.dx + 2));
        jhi = std::min(jhi, (int)lround((sp.radius - s->prob_lo[1]) / g.dy + 2));
    }
    a.ilo = ilo; a.jlo = jlo; a.ncx = std::max(0, ihi - ilo + 1); a.ncy = std::max(0, jhi - jlo + 1);
    a.ppcx = sp.ppc[0]; a.ppcy = sp.ppc[1];
    const int nppc = a.ppcx * a.ppcy;
    a.plo_x = s->prob_lo[0]; a.plo_y = s->prob_lo[1]; a.dx = g.dx; a.dy = g.dy;
    a.blo_x = s->bc_lo[0]; a.blo_y = s->bc_lo[1]; a.bhi_x = s->bc_hi[0]; a.bhi_y = s->bc_hi[1];
    a.radius_sq = sp.radius * sp.radius; a.hollow_sq = sp.hollow * sp.hollow;
    a.min_density = sp.min_density; a.c_t = c_t;
    a.scale = nppc <= 0 ? 0. : (g.normalized ? 1. / nppc : g.dx * g.dy * g.dz / nppc);   // :40-41
    a.ux0 = sp.u_mean[0] * g.c; a.uy0 = sp.u_mean[1] * g.c;
    a.psi0 = sqrt(1. + sp.u_mean[0] * sp.u_mean[0] + sp.u_mean[1] * sp.u_mean[1]
                  + sp.u_mean[2] * sp.u_mean[2]) - sp.u_mean[2];
    const long ncand = (long)a.ncx * a.ncy * nppc;
    sp.d.np = 0;
    if (ncand == 0) return HPB_OK;
    int rc = ensure_scan(s, ncand);
    if (rc) return rc;
    k_plasma_flag<<<nb256(ncand), 256, 0, s->stream>>>(a, sp.density, ncand, s->d_flag);
    size_t tb = s->cub_bytes;
    SIM_CUDA(cub::DeviceScan::ExclusiveSum(s->d_cub, tb, s->d_flag, s->d_offs, (int)ncand, s->stream));
    unsigned last[2];
    SIM_CUDA(cudaMemcpyAsync(&last[0], s->d_flag + ncand - 1, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
    SIM_CUDA(cudaMemcpyAsync(&last[1], s->d_offs + ncand - 1, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
    SIM_CUDA(cudaStreamSynchronize(s->stream));
    const long np = (long)last[0] + last[1];
    if (np > sp.capacity) {
        for (int k = 0; k < HPB_PLASMA_NREAL; ++k) { cudaFree(sp.d.r[k]); sp.d.r[k] = nullptr; }
        cudaFree(sp.d.idcpu); sp.d.idcpu = nullptr;
        for (int k = 0; k < HPB_PLASMA_NREAL; ++k) SIM_CUDA(cudaMalloc(&sp.d.r[k], sizeof(double) * np));
        SIM_CUDA(cudaMalloc(&sp.d.idcpu, sizeof(uint64_t) * np));
        sp.capacity = np;
    }
    sp.d.np = np;
    sp.lattice_n = (np == ncand && nppc > 0) ? ncand / nppc : 0;
    sp.lattice_ppc = nppc;
    if (np > 0)
        k_plasma_fill<<<nb256(ncand), 256, 0, s->stream>>>(a, sp.density, ncand, s->d_flag, s->d_offs, sp.d);
    SIM_CUDA(cudaGetLastError());
    return HPB_OK;
}

int init_beam(hpb_sim *s, BeamSp &b)
{
    const hpb_geom &g = s->g;
    BeamInitArgs a;
    int ilo = 0, ihi = g.nx - 1, jlo = 0, jhi = g.ny - 1;
    if (std::isfinite(b.radius)) {   // bounding box of the radius cut: a superset, order preserved
        ilo = std::max(ilo, (int)floor((b.pos_mean[0] - b.radius - s->prob_lo[0]) / g.dx) - 2);
        jlo = std::max(jlo, (int)floor((b.pos_mean[1] - b.radius - s->prob_lo[1]) / g.dy) - 2);
        ihi = std::min(ihi, (int)ceil((b.pos_mean[0] + b.radius - s->prob_lo[0]) / g.dx) + 2);
        jhi = std::min(jhi, (int)ceil((b.pos_mean[1] + b.radius - s->prob_lo[1]) / g.dy) + 2);
    }
    a.ilo = ilo; a.jlo = jlo; a.ncx = std::max(0, ihi - ilo + 1); a.ncy = std::max(0, jhi - jlo + 1);
    a.nz = s->nz;
    a.ppcx = b.ppc[0]; a.ppcy = b.ppc[1]; a.ppcz = b.ppc[2];
    const int nppc = a.ppcx * a.ppcy * a.ppcz;
    a.plo_x = s->prob_lo[0]; a.plo_y = s->prob_lo[1]; a.plo_z = s->prob_lo[2];
    a.dx = g.dx; a.dy = g.dy; a.dz = g.dz;
    a.x_mean = b.pos_mean[0]; a.y_mean = b.pos_mean[1]; a.z_mean = b.pos_mean[2];
    a.sx = b.pos_std[0]; a.sy = b.pos_std[1]; a.sz = b.pos_std[2];
    a.zmin = b.zmin; a.zmax = b.zmax; a.radius_sq = b.radius * b.radius;
    a.density = b.density; a.min_density = b.min_density;
    a.scale = g.normalized ? 1. / nppc : g.dx * g.dy * g.dz / nppc;
    a.profile = b.profile;
    a.ux0 = b.u_mean[0] * g.c; a.uy0 = b.u_mean[1] * g.c; a.uz0 = b.u_mean[2] * g.c;
    const long per_slice = (long)a.ncx * a.ncy * nppc;
    const long ncand = per_slice * s->nz;
    std::vector<long> slot_off(s->nz + 1, 0);
    if (ncand > 0) {
        int rc = ensure_scan(s, ncand);
        if (rc) return rc;
        k_beam_flag<<<nb256(ncand), 256, 0, s->stream>>>(a, ncand, s->d_flag);
        size_t tb = s->cub_bytes;
        SIM_CUDA(cub::DeviceScan::ExclusiveSum(s->d_cub, tb, s->d_flag, s->d_offs, (int)ncand, s->stream));
        k_slot_offsets<<<nb256(s->nz + 1), 256, 0, s->stream>>>(s->d_flag, s->d_offs, per_slice, s->nz, ncand,
                                                                s->d_slot_off);
        SIM_CUDA(cudaMemcpyAsync(slot_off.data(), s->d_slot_off, sizeof(long) * (s->nz + 1),
                                 cudaMemcpyDeviceToHost, s->stream));
        SIM_CUDA(cudaStreamSynchronize(s->stream));
    } else {
        SIM_CUDA(cudaMemsetAsync(s->d_slot_off, 0, sizeof(long) * (s->nz + 1), s->stream));
    }
    // slice capacity: the largest initial slice + 25 % + 1024 for slipped particles, unless the
    // deck fixes it (our addition: <beam>.slice_capacity); identical on every rank of a pipeline
    long mx = 0;
    for (int k = 0; k < s->nz; ++k) mx = std::max(mx, slot_off[k + 1] - slot_off[k]);
    long cap = (long)s->deck.num(b.name + ".slice_capacity", (double)(mx + mx / 4 + 1024));
    if (cap < mx) { hpb_set_error("beam %s: slice_capacity %ld < largest slice %ld", b.name.c_str(), cap, mx); return HPB_ERR_CAPACITY; }
    int rc = hpb_beam_rings_alloc(s, b, cap);
    if (rc) return rc;
    const BeamRing &r = b.ring[b.cur];
    k_ring_set_counts<<<nb256(s->nz), 256, 0, s->stream>>>(r, s->d_slot_off);
    if (slot_off[s->nz] > 0)
        k_beam_fill<<<nb256(ncand), 256, 0, s->stream>>>(a, ncand, s->d_flag, s->d_offs, s->d_slot_off, r, 1);
    SIM_CUDA(cudaGetLastError());
    b.initialised = true;
    b.from_host = false;
    b.cs_valid = false;
    return HPB_OK;
}

// view of the beam particles of slice islice in the current ring (empty for islice < 0)
hpb_beam_slice beam_slice_view(const hpb_sim *s, const BeamSp &b, int islice)
{
    hpb_beam_slice v = {};
    if (islice < 0 || islice >= s->nz || !b.ring[b.cur].base) return v;
    return b.ring[b.cur].view(s->nz - 1 - islice);
}

enum { ST_DEPOSIT = 0, ST_POISSON, ST_EXPLICIT, ST_MG, ST_PUSH, ST_OTHER, ST_N };

struct StageTimer {
    hpb_sim *s; double *acc;
    cudaEvent_t a, b;
    StageTimer(hpb_sim *s_, int st) : s(s_), acc(nullptr)
    {
        if (!s->opt_profile) return;
        static double *slots[ST_N];
        slots[ST_DEPOSIT] = &s->stats.ms_deposit; slots[ST_POISSON] = &s->stats.ms_poisson;
        slots[ST_EXPLICIT] = &s->stats.ms_explicit; slots[ST_MG] = &s->stats.ms_mg;
        slots[ST_PUSH] = &s->stats.ms_push; slots[ST_OTHER] = &s->stats.ms_other;
        acc = slots[st];
        a = s->pev[0]; b = s->pev[1];
        cudaEventRecord(a, s->stream);
    }
    ~StageTimer()
    {
        if (!acc) return;
        cudaEventRecord(b, s->stream);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        *acc += ms;
    }
};

// MultiPlasma::maxChargeDensity (particles/plasma/MultiPlasma.cpp:63-73) as a function of z = c t
std::function<double(double)> max_charge_density(hpb_sim *s)
{
    return [s](double z) {
        double m = fabs(s->adaptive_density * s->g.q_e);
        for (auto &sp : s->plasmas) {
            const double xyz[3] = {0., 0., z};
            m = fmax(m, fabs(sp.charge * s->deck.run(sp.density_host, xyz)));
        }
        return m;
    };
}

// CalculateFromMinUz (AdaptiveTimeStep.cpp:143-233) from the host copies BeamSp::ts
int adaptive_from_min_uz(hpb_sim *s, double t_now, const std::function<double(double)> &rho_at)
{
    const int nb = (int)s->beams.size();
    std::vector<double> ts(4 * nb), q(nb), m(nb);
    for (int ib = 0; ib < nb; ++ib) {
        for (int k = 0; k < 4; ++k) ts[4 * ib + k] = s->beams[ib].ts[k];
        q[ib] = s->beams[ib].charge; m[ib] = s->beams[ib].mass;
    }
    double dt_out = s->dt, mq = s->min_uz_mq;
    if (!adaptive_dt_from_min_uz(s->adp, nb, ts.data(), q.data(), m.data(), rho_at, t_now, s->dt, dt_out, mq)) {
        hpb_set_error("adaptive time step: the sum of the beam weights is 0 or no plasma density > 0 is given "
                      "(plasmas.adaptive_density)");
        return HPB_ERR_ARG;
    }
    s->dt = dt_out; s->min_uz_mq = mq;
    return HPB_OK;
}

int begin_step(hpb_sim *s, int step)
{
    // ResetAllQuantities (Hipace.cpp:730-742)
    SIM_CUDA(cudaMemsetAsync(s->sl.p, 0, sizeof(double) * s->sl.nstride * s->sl.ncomp, s->stream));
    SIM_CUDA(cudaMemsetAsync(s->d_checksum, 0, sizeof(double) * (s->sl.ncomp + 1), s->stream));
    for (int k = 0; k < HPB_C_COUNT; ++k) s->comps[k] = s->comps0[k];
    s->prepared = false;
    // the physical time of a step after the first comes from the rank that owns the step before it
    // (MultiBuffer::get_time, Hipace.cpp:411); without a pipeline it is the value this rank computed itself
    double t_recv = s->next_time;
    if (int rc = hpb_pipeline_get_time(s, step, &t_recv)) return rc;
    if (s->adaptive_dt) {                                            // Hipace.cpp:275-281, 411, 420, 434
        const auto rho_at = max_charge_density(s);
        s->adp.numprocs = hpb_pipeline_world(s);
        if (step == 0 || !s->adaptive_initialised) {
            // Hipace.cpp:275-281: the head rank's estimate from the beam as specified in the deck is
            // broadcast to every rank -- it depends on the deck only, so every rank computes it
            if (s->beams.empty()) { hpb_set_error("adaptive time step: needs a beam"); return HPB_ERR_ARG; }
            for (auto &b : s->beams) {                               // GatherMinUzSlice(initial)
                b.ts[0] = b.u_mean_z; b.ts[1] = 1.; b.ts[2] = b.u_mean_z; b.ts[3] = b.u_mean_z * b.u_mean_z;
            }
            s->min_uz_mq = DBL_MAX;
            s->dt = 0.;
            int rc = adaptive_from_min_uz(s, 0., rho_at);
            if (rc) return rc;
            s->dt = adaptive_dt_from_density(s->adp, s->min_uz_mq, 0., s->dt, rho_at);
            s->adaptive_initialised = true;
        }
        s->time = step == 0 ? 0. : t_recv;
        s->dt = adaptive_dt_from_density(s->adp, s->min_uz_mq, s->time, s->dt, rho_at);
        s->next_time = s->time + s->dt;
        for (auto &b : s->beams) {                                   // CalculateFromDensity resets the data
            b.ts[0] = 1e30; b.ts[1] = b.ts[2] = b.ts[3] = 0.;
            if (!b.d_ts) SIM_CUDA(cudaMalloc(&b.d_ts, 4 * sizeof(double)));
            SIM_CUDA(cudaMemcpyAsync(b.d_ts, b.ts, 4 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        }
    } else if (hpb_pipeline_receives(s, step)) {
        s->time = t_recv;
        s->next_time = s->time + s->dt;
    } else {
        // the reference accumulates next_time = time + dt from step to step (Hipace.cpp:411-434);
        // replaying the sum gives the same rounding whichever step a run (or a test) starts from
        s->time = 0.;
        for (int k = 0; k < step; ++k) s->time += s->dt;
        s->next_time = s->time + s->dt;
    }
    s->dt_step = s->dt;
    if (int rc = hpb_pipeline_put_time(s, step, s->next_time)) return rc;             // Hipace.cpp:445-447
    const double c_t = s->g.c * s->time;
    if (s->use_laser && s->max_step > 0 && s->dt != 0.) {            // the envelope is stored and advanced
        if (!s->laser_state) {
            int rc = hpb_laser_state_create(&s->laser_state, s->ctx, s->nz, s->lasers.data(), (int)s->lasers.size(),
                                            s->laser_lambda0, s->laser_interp_order, s->laser_use_phase);
            if (rc) return rc;
            hpb_laser_set_solver(s->laser_state, s->laser_use_mg, s->laser_mg_tol_rel, s->laser_mg_tol_abs,
                                 s->laser_mg_avg_rhs);
        }
        // MultiLaser::SetInitialChi (laser/MultiLaser.cpp:293-332): sum over species of density * q^2 mu0 / m
        std::vector<double> chi0((size_t)s->g.nx * s->g.ny, 0.);
        for (auto &sp : s->plasmas) {
            const double f = sp.charge * sp.charge * s->g.mu0 / sp.mass;
            for (int j = 0; j < s->g.ny; ++j)
                for (int i = 0; i < s->g.nx; ++i) {
                    const double xyz[3] = {i * s->g.dx + s->g.x_off, j * s->g.dy + s->g.y_off, c_t};
                    chi0[(size_t)j * s->g.nx + i] += s->deck.run(sp.density_host, xyz) * f;
                }
        }
        SIM_CUDA(cudaStreamSynchronize(s->stream));          // chi0 is a temporary
        int rc = hpb_laser_begin_step(s->laser_state, s->ctx, chi0.data());
        if (rc) return rc;
        SIM_CUDA(cudaStreamSynchronize(s->stream));
    }
    for (auto &sp : s->plasmas) {                                    // Hipace.cpp:450
        int rc = init_plasma(s, sp, c_t);
        if (rc) return rc;
    }
    // beams: created at step 0 on the head rank (BeamParticleContainerInit.cpp), afterwards they
    // arrive slice by slice from the previous time step (MultiBuffer::get_data) -- from the
    // upstream rank in a pipeline, otherwise from the ring the previous step on this GPU filled
    const bool recv = hpb_pipeline_receives(s, step);
    // end_step left BeamSp::cur on the ring this rank is still sending from: receive into the
    // other one (the previous step's input) and reuse the sending ring, slot by slot, as output
    const bool flip = recv && hpb_pipeline_out_ring_busy(s);
    for (auto &b : s->beams) {
        if (flip && b.ring[0].base) b.cur ^= 1;
        if (!b.ring[0].base || (!recv && !b.initialised) || (!recv && step == 0 && !b.from_host)) {
            int rc = init_beam(s, b);
            if (rc) return rc;
        }
        SIM_CUDA(cudaMemsetAsync(b.d_cs, 0, 9 * sizeof(double), s->stream));
        b.cs_valid = false;
    }
    if (hpb_pipeline_active(s)) { int rc = hpb_pipeline_begin_step(s, step); if (rc) return rc; }
    // DepositNeutralizingBackground (Hipace.cpp:468-470, MultiPlasma.cpp:106-118)
    for (auto &sp : s->plasmas) {
        if (!sp.neutralize) continue;
        int rc = hpb_deposit_current(s->ctx, sp.d, s->sl, -sp.charge, sp.mass, -1, -1, -1, -1,
                                     s->comps[HPB_C_IONS_RHOMJZ], sp.max_qsa, s->d_nqsa);
        if (rc) return rc;
    }
    s->cur_step = step;
    return HPB_OK;
}

// after the last slice of a step: the pushed beam becomes the current one
int write_beam_insitu(hpb_sim *s);
int end_step(hpb_sim *s)
{
    if (s->adaptive_dt) {                                            // Hipace.cpp:482-483
        for (auto &b : s->beams) {
            SIM_CUDA(cudaStreamSynchronize(s->stream));
            if (s->stream2) SIM_CUDA(cudaStreamSynchronize(s->stream2));
            SIM_CUDA(cudaMemcpy(b.ts, b.d_ts, 4 * sizeof(double), cudaMemcpyDeviceToHost));
        }
        if (int rc = adaptive_from_min_uz(s, s->time, max_charge_density(s))) return rc;
    }
    if (int rc = write_beam_insitu(s)) return rc;
    if (s->laser_state) hpb_laser_end_step(s->laser_state);
    for (auto &b : s->beams) { b.cur ^= 1; b.cs_valid = s->opt_checksums; b.initialised = true; }
    if (hpb_pipeline_active(s)) return hpb_pipeline_end_step(s, s->cur_step);
    return HPB_OK;
}

// the kernel seams take their stream from the context: route a group of calls to another stream
struct StreamScope {
    hpb_sim *s; cudaStream_t old_ctx, old_beam;
    StreamScope(hpb_sim *s_, cudaStream_t st) : s(s_), old_ctx(s_->ctx->stream), old_beam(s_->beam_stream)
    {
        s->ctx->stream = st; s->beam_stream = st;
    }
    ~StreamScope() { s->ctx->stream = old_ctx; s->beam_stream = old_beam; }
};

// beam jz of slice isl (Hipace.cpp:582-585, 613-614)
int beam_deposit_jz(hpb_sim *s, int isl)
{
    int rc;
    if (s->beams.empty() || isl < 0) return HPB_OK;
    if ((rc = hpb_pipeline_wait_slice(s, isl))) return rc;
    for (auto &b : s->beams)
        if ((rc = hpb_beam_deposit(s->ctx, beam_slice_view(s, b, isl), s->sl, b.charge, -1, -1,
                                   s->comps[HPB_C_JZ_BEAM]))) return rc;
    return HPB_OK;
}

// beam jx, jy of slice isl - 1 into Next, then the Sx / Sy seed of slice isl (Hipace.cpp:639-660)
int beam_next_and_sxsy(hpb_sim *s, int isl)
{
    int rc;
    if (!s->beams.empty() && isl > 0 && (rc = hpb_pipeline_wait_slice(s, isl - 1))) return rc;
    if (s->do_beam_jx_jy)
        for (auto &b : s->beams)
            if ((rc = hpb_beam_deposit(s->ctx, beam_slice_view(s, b, isl - 1), s->sl, b.charge,
                                       s->comps[HPB_C_NEXT_JX_BEAM], s->comps[HPB_C_NEXT_JY_BEAM], -1))) return rc;
    return hpb_fields_sxsy_from_beam(s->ctx, s->sl, s->comps);
}

// utils::doDiagnostics (utils/IOUtil.cpp:75-82; no max_time in our decks)
bool do_diagnostics(int period, int step, int max_step)
{
    return period > 0 && (step == max_step || step % period == 0);
}

// mkdir -p <prefix>; <prefix>/reduced_<name>.<rank, 4 digits>.txt
std::string insitu_path(hpb_sim *s, const std::string &prefix, const std::string &name)
{
    std::string dir;
    for (char ch : prefix + "/") {
        if (ch == '/' && !dir.empty()) mkdir(dir.c_str(), 0777);
        dir += ch;
    }
    char rank[16];
    snprintf(rank, sizeof(rank), "%04d", hpb_pipeline_active(s) ? hpb_pipeline_rank(s) : 0);
    return prefix + "/reduced_" + name + "." + rank + ".txt";
}

// Beam- and PlasmaParticleContainer::InSituWriteToFile at the end of a time step (Hipace.cpp:488-489)
int write_beam_insitu(hpb_sim *s)
{
    if (s->d_field_insitu && do_diagnostics(s->field_insitu_period, s->cur_step, s->max_step)) {   // :487
        std::vector<double> h(10 * (size_t)s->nz);
        SIM_CUDA(cudaStreamSynchronize(s->stream));
        SIM_CUDA(cudaMemcpy(h.data(), s->d_field_insitu, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
        SIM_CUDA(cudaMemset(s->d_field_insitu, 0, sizeof(double) * h.size()));
        int rc = hpb_insitu_write_fields(insitu_path(s, s->field_insitu_prefix, "fields").c_str(), s->time,
                                         s->cur_step, s->nz, s->prob_lo[2], s->prob_hi[2], s->g.normalized,
                                         s->g.dx * s->g.dy * s->g.dz, h.data());
        if (rc) return rc;
    }
    if (s->d_laser_insitu && do_diagnostics(s->laser_insitu_period, s->cur_step, s->max_step)) {   // :490
        std::vector<double> h(8 * (size_t)s->nz);
        SIM_CUDA(cudaStreamSynchronize(s->stream));
        SIM_CUDA(cudaMemcpy(h.data(), s->d_laser_insitu, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
        SIM_CUDA(cudaMemset(s->d_laser_insitu, 0, sizeof(double) * h.size()));
        int rc = hpb_insitu_write_laser(insitu_path(s, s->laser_insitu_prefix, "laser").c_str(), s->time,
                                        s->cur_step, s->nz, s->prob_lo[2], s->prob_hi[2], s->g.normalized,
                                        s->g.dx * s->g.dy * s->g.dz, s->g.nx, s->g.ny, h.data());
        if (rc) return rc;
    }
    for (auto &sp : s->plasmas) {
        if (!sp.d_insitu || !do_diagnostics(sp.insitu_period, s->cur_step, s->max_step)) continue;
        std::vector<double> h(15 * (size_t)s->nz);
        SIM_CUDA(cudaStreamSynchronize(s->stream));
        SIM_CUDA(cudaMemcpy(h.data(), sp.d_insitu, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
        SIM_CUDA(cudaMemset(sp.d_insitu, 0, sizeof(double) * h.size()));
        const double ndf = s->g.normalized ? s->g.dx * s->g.dy * s->g.dz : 1.0;
        int rc = hpb_insitu_write_plasma(insitu_path(s, sp.insitu_file_prefix, sp.name).c_str(), s->time,
                                         s->cur_step, s->nz, sp.charge, sp.mass, s->prob_lo[2], s->prob_hi[2],
                                         ndf, s->g.normalized, h.data());
        if (rc) return rc;
    }
    for (auto &b : s->beams) {
        if (!b.d_insitu || !do_diagnostics(b.insitu_period, s->cur_step, s->max_step)) continue;
        std::vector<double> h(23 * (size_t)s->nz);
        SIM_CUDA(cudaStreamSynchronize(s->stream));
        if (s->stream2) SIM_CUDA(cudaStreamSynchronize(s->stream2));
        SIM_CUDA(cudaMemcpy(h.data(), b.d_insitu, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
        SIM_CUDA(cudaMemset(b.d_insitu, 0, sizeof(double) * h.size()));
        const std::string path = insitu_path(s, b.insitu_file_prefix, b.name);
        const double ndf = s->g.normalized ? s->g.dx * s->g.dy * s->g.dz : 1.0;
        int rc = hpb_insitu_write_beam(path.c_str(), s->time, s->cur_step, s->nz, b.charge, b.mass,
                                       s->prob_lo[2], s->prob_hi[2], ndf, s->g.normalized, h.data());
        if (rc) return rc;
    }
    return HPB_OK;
}

// AdvanceBeamParticlesSlice + shiftSlippedParticles + MultiBuffer::put_data (Hipace.cpp:707-716)
int beam_push_and_send(hpb_sim *s, int islice)
{
    int rc;
    // (a laser-only deck still hands its envelope slices to the next time step)
    if (s->beams.empty()) return s->laser_state ? hpb_pipeline_send_slice(s, islice, s->cur_step) : HPB_OK;
    const int slot = s->nz - 1 - islice;
    const double min_z = s->prob_lo[2] + islice * s->g.dz;
    const double time = s->time;
    if ((rc = hpb_pipeline_wait_out_slot(s, islice))) return rc;
    // the slipped particles of this slice are appended to the NEXT slice's packet: its receive must
    // have landed whichever solver branch called us (the predictor-corrector branch only waits for
    // it when it deposits the next slice's beam currents)
    if (islice > 0 && (rc = hpb_pipeline_wait_slice(s, islice - 1))) return rc;
    for (auto &b : s->beams) {
        const BeamRing &in = b.ring[b.cur], &out = b.ring[b.cur ^ 1];
        const hpb_beam_slice bm = in.view(slot);
        // in-situ diagnostics of the slice BEFORE the push (Hipace.cpp:680-681)
        if (do_diagnostics(b.insitu_period, s->cur_step, s->max_step)) {
            if (!b.d_insitu) {
                SIM_CUDA(cudaMalloc(&b.d_insitu, sizeof(double) * 23 * (size_t)s->nz));
                SIM_CUDA(cudaMemsetAsync(b.d_insitu, 0, sizeof(double) * 23 * (size_t)s->nz, s->ctx->stream));
            }
            if ((rc = hpb_beam_insitu_slice(s->ctx, bm, b.insitu_radius, b.d_insitu + islice, s->nz))) return rc;
        }
        if ((rc = hpb_advance_beam_impl(s->ctx, bm, in.nsub(slot), s->sl, b.charge, b.mass, b.n_subcycles,
                                        s->dt, time, min_z, b.do_z_push, s->particle_bc, s->bc_lo,
                                        s->bc_hi, s->comps, b.ext, b.d_class,
                                        s->opt_checksums ? b.d_cs : nullptr, s->d_count))) return rc;
        hpb_beam_slice next = {};
        if (islice > 0) next = in.view(slot + 1);
        if ((rc = hpb_beam_shift_slipped(s->ctx, bm, in.nsub(slot), min_z, b.d_class, out.view(slot),
                                         out.hdr(slot), next, islice > 0 ? in.hdr(slot + 1) : nullptr,
                                         islice > 0 ? in.nsub(slot + 1) : nullptr, s->d_overflow))) return rc;
        // GatherMinUzSlice of the pushed slice without its slipped particles (Hipace.cpp:715)
        if (s->adaptive_dt && (rc = hpb_beam_min_uz_slice(s->ctx, out.view(slot), b.d_ts))) return rc;
    }
    return hpb_pipeline_send_slice(s, islice, s->cur_step);
}

// Hipace::SolveOneSlice, explicit branch (Hipace.cpp:556-728).
// Reference order (opt_fuse = 0): InitializeSlices, DepositCurrent, beam jz, Poisson, beam jx jy
// (Next), Sx Sy seed, ExplicitDeposition, multigrid, AdvancePlasmaParticles, beam push, ShiftSlices.
// Fused order (default), after the multigrid solve of slice k:
//   main stream : shift + initialise the planes for slice k-1 (one pass), then the plasma push
//                 of slice k, which deposits jx jy chi rhomjz of slice k-1 from registers;
//   side stream : beam push / re-binning / hand-off of slice k, then the beam deposits and the
//                 Sx, Sy seed of slice k-1 -- small latency-bound kernels that now run beside the
//                 plasma push instead of in front of it.
// The next call then starts at the Poisson solve.  (Without a side stream the same calls are
// issued on the main stream in the same order.)
// hipace.bxby_solver = predictor-corrector: Hipace::SolveOneSlice (the non-explicit branches of
// Hipace.cpp:556-728) and PredictorCorrectorLoopToSolveBxBy (Hipace.cpp:935-1031), with
// boundary.field = Dirichlet or Open.  Built from the validated kernels of the explicit path (push,
// deposit, DST Poisson solves) plus the field kernels of pc_fields.cu.
int solve_one_slice_pc(hpb_sim *s, int islice)
{
    hpb_ctx *ctx = s->ctx;
    const int *C = s->comps;
    const hpb_slice &sl = s->sl;
    int rc;
    auto open_bc = [&](double *plane, int monopole) -> int {
        return s->open_bc ? hpb_fields_open_boundary(ctx, plane, monopole, s->prob_lo[0], s->prob_hi[0],
                                                     s->prob_lo[1], s->prob_hi[1], s->d_pc_scal) : HPB_OK;
    };
    auto rel_error = [&](int ca, int cb, double &err) -> int {                  // Fields.cpp:1227-1286
        int r = hpb_fields_rel_b_error(ctx, sl, ca, cb, s->d_pc_scal + 38);
        if (r) return r;
        double h[2];
        SIM_CUDA(cudaMemcpyAsync(h, s->d_pc_scal + 38, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        SIM_CUDA(cudaStreamSynchronize(ctx->stream));
        err = h[0] > 0. ? h[1] / h[0] : 0.;
        return HPB_OK;
    };
    const long plane = (long)s->g.nx * s->g.ny;
    {
        StageTimer t(s, ST_OTHER);                                              // Fields.cpp:565-566
        const int z[7] = {C[HPB_C_EXMBY], C[HPB_C_EYPBX], C[HPB_C_JX], C[HPB_C_JY], C[HPB_C_JZ],
                          C[HPB_C_RHOMJZ], C[HPB_C_RHO]};
        if ((rc = hpb_fields_zero(ctx, sl, z, 7))) return rc;
    }
    {
        StageTimer t(s, ST_DEPOSIT);
        for (auto &sp : s->plasmas)                                             // :617-618
            if ((rc = hpb_deposit_current_jz(ctx, sp.d, sl, sp.charge, sp.mass, C[HPB_C_JX], C[HPB_C_JY],
                                             C[HPB_C_JZ], C[HPB_C_RHO], -1, C[HPB_C_RHOMJZ], -1, sp.max_qsa,
                                             s->d_nqsa))) return rc;
    }
    {
        StageTimer t(s, ST_OTHER);
        if (!s->beams.empty()) {                                                // :621-623
            if ((rc = hpb_pipeline_wait_slice(s, islice))) return rc;
            for (auto &b : s->beams)
                if ((rc = hpb_beam_deposit(ctx, beam_slice_view(s, b, islice), sl, b.charge,
                                           s->do_beam_jx_jy ? C[HPB_C_JX] : -1,
                                           s->do_beam_jx_jy ? C[HPB_C_JY] : -1, C[HPB_C_JZ]))) return rc;
        }
        if ((rc = hpb_fields_add_rho_ions(ctx, sl, C))) return rc;              // :626
        if (s->use_grid_current)                                                // :629
            if ((rc = hpb_fields_grid_current(ctx, sl, C[HPB_C_JZ], s->gc_peak, s->gc_mean, s->gc_std,
                                              s->prob_lo[0], s->prob_lo[1],
                                              s->prob_lo[2] + islice * s->g.dz))) return rc;
    }
    {
        StageTimer t(s, ST_POISSON);                                            // :633
        if (!s->open_bc) {
            if ((rc = hpb_fields_solve_psi_ez_bz(ctx, sl, C))) return rc;
        } else {
            if ((rc = hpb_fields_psi_ez_bz_rhs(ctx, sl, C, s->d_pc_rhs))) return rc;
            for (int b = 0; b < 3; ++b)                 // Ez and Bz have no monopole (Fields.cpp:729-733)
                if ((rc = open_bc(s->d_pc_rhs + b * plane, b == 0))) return rc;
            const int lhs[3] = {C[HPB_C_PSI], C[HPB_C_EZ], C[HPB_C_BZ]};
            if ((rc = hpb_poisson_solve(ctx, s->d_pc_rhs, sl, lhs, 3))) return rc;
            if ((rc = hpb_launch_exmby_eypbx(ctx, sl, C))) return rc;
        }
    }
    {
        StageTimer t(s, ST_MG);         // (the stage slot of the Bx/By solve)
        double err = 0.;
        if ((rc = rel_error(C[HPB_C_PREV_BX], C[HPB_C_PCPREV_BX], err))) return rc;      // :941-943
        const double q = err / (2.5 * s->predcorr_tol);
        const double mix0 = exp(-0.5 * q * q);                                           // Fields.cpp:1149-1170
        if ((rc = hpb_fields_lincomb2(ctx, sl, C[HPB_C_BX], 1.0 + mix0, C[HPB_C_PREV_BX], -mix0,
                                      C[HPB_C_PCPREV_BX]))) return rc;
        if ((rc = hpb_fields_lincomb2(ctx, sl, C[HPB_C_PCITER_BX], 0., C[HPB_C_BX], 0., C[HPB_C_BX]))) return rc;
        if ((rc = hpb_fields_lincomb2(ctx, sl, C[HPB_C_PCPREV_BX], 1., C[HPB_C_BX], 0., C[HPB_C_BX]))) return rc;
        int i_iter = 0;
        double err_prev = 1.0;
        err = 1.0;
        while (err > s->predcorr_tol && i_iter < s->predcorr_max_iter) {                 // :961-1010
            ++i_iter;
            for (auto &sp : s->plasmas) {                   // push to the temporary slice
                hpb_set_plasma_lattice_hint(ctx, sp.lattice_n, sp.lattice_ppc);
                rc = hpb_advance_plasma_particles(ctx, sp.d, sl, sp.charge, sp.mass, sp.n_subcycles, 1,
                                                  s->particle_bc, s->bc_lo, s->bc_hi, C);
                hpb_set_plasma_lattice_hint(ctx, 0, 0);
                if (rc) return rc;
            }
            for (auto &sp : s->plasmas)                     // jx, jy of the next slice
                if ((rc = hpb_deposit_current(ctx, sp.d, sl, sp.charge, sp.mass, C[HPB_C_NEXT_JX],
                                              C[HPB_C_NEXT_JY], -1, -1, -1, sp.max_qsa, s->d_nqsa))) return rc;
            if (s->do_beam_jx_jy && !s->beams.empty() && islice > 0) {
                if ((rc = hpb_pipeline_wait_slice(s, islice - 1))) return rc;
                for (auto &b : s->beams)
                    if ((rc = hpb_beam_deposit(ctx, beam_slice_view(s, b, islice - 1), sl, b.charge,
                                               C[HPB_C_NEXT_JX], C[HPB_C_NEXT_JY], -1))) return rc;
            }
            // SolvePoissonBxBy -> PCIter (Fields.cpp:1008-1078)
            if ((rc = hpb_fields_bxby_rhs(ctx, sl, C, s->d_pc_rhs))) return rc;
            if ((rc = open_bc(s->d_pc_rhs, 1))) return rc;
            if ((rc = open_bc(s->d_pc_rhs + plane, 1))) return rc;
            const int lhs[2] = {C[HPB_C_PCITER_BX], C[HPB_C_PCITER_BY]};
            if ((rc = hpb_poisson_solve(ctx, s->d_pc_rhs, sl, lhs, 2))) return rc;
            if ((rc = rel_error(C[HPB_C_BX], C[HPB_C_PCITER_BX], err))) return rc;
            if (i_iter == 1) err_prev = err;
            // MixAndShiftBfields (Fields.cpp:1172-1225)
            double w_it = 0.5, w_prev = 0.5;
            if (err != 0. || err_prev != 0.) { w_it = err_prev / (err + err_prev); w_prev = err / (err + err_prev); }
            if ((rc = hpb_fields_lincomb2(ctx, sl, C[HPB_C_PCPREV_BX], w_it, C[HPB_C_PCITER_BX], w_prev,
                                          C[HPB_C_PCPREV_BX]))) return rc;
            if ((rc = hpb_fields_lincomb2(ctx, sl, C[HPB_C_BX], 1.0 - s->predcorr_mix, C[HPB_C_BX],
                                          s->predcorr_mix, C[HPB_C_PCPREV_BX]))) return rc;
            if ((rc = hpb_fields_lincomb2(ctx, sl, C[HPB_C_PCPREV_BX], 1., C[HPB_C_PCITER_BX], 0.,
                                          C[HPB_C_PCITER_BX]))) return rc;
            const int z[2] = {C[HPB_C_NEXT_JX], C[HPB_C_NEXT_JY]};                      // :996-999
            if ((rc = hpb_fields_zero(ctx, sl, z, 2))) return rc;
            err_prev = err;
        }
        s->n_predcorr_iters += i_iter;
        s->stats.n_mg_vcycles += i_iter;        // reported in the iteration slot of the statistics
    }
    if (s->opt_checksums) {
        StageTimer t(s, ST_OTHER);
        if ((rc = slice_checksums(s))) return rc;
    }
    {
        StageTimer t(s, ST_PUSH);
        for (auto &sp : s->plasmas) {
            s->stats.n_plasma_pushed += (double)sp.d.np;
            hpb_set_plasma_lattice_hint(ctx, sp.lattice_n, sp.lattice_ppc);
            rc = hpb_advance_plasma_particles(ctx, sp.d, sl, sp.charge, sp.mass, sp.n_subcycles, 0,
                                              s->particle_bc, s->bc_lo, s->bc_hi, C);
            hpb_set_plasma_lattice_hint(ctx, 0, 0);
            if (rc) return rc;
        }
    }
    {
        StageTimer t(s, ST_OTHER);
        if ((rc = beam_push_and_send(s, islice))) return rc;
        // ShiftSlices (Fields.cpp:600-603)
        if ((rc = hpb_fields_lincomb2(ctx, sl, C[HPB_C_PCPREV_BX], 1., C[HPB_C_PREV_BX], 0., C[HPB_C_PREV_BX]))) return rc;
        if ((rc = hpb_fields_lincomb2(ctx, sl, C[HPB_C_PREV_BX], 1., C[HPB_C_BX], 0., C[HPB_C_BX]))) return rc;
        if ((rc = hpb_fields_lincomb2(ctx, sl, C[HPB_C_PREV_JX], 1., C[HPB_C_JX], 0., C[HPB_C_JX]))) return rc;
    }
    s->stats.n_cells_updated += (double)s->g.nx * s->g.ny;
    s->stats.n_slices += 1;
    return HPB_OK;
}

// PlasmaParticleContainer::InSituComputeDiags at the start of a slice (Hipace.cpp:587)
int plasma_insitu(hpb_sim *s, int islice)
{
    for (auto &sp : s->plasmas) {
        if (!do_diagnostics(sp.insitu_period, s->cur_step, s->max_step)) continue;
        if (!sp.d_insitu) {
            SIM_CUDA(cudaMalloc(&sp.d_insitu, sizeof(double) * 15 * (size_t)s->nz));
            SIM_CUDA(cudaMemsetAsync(sp.d_insitu, 0, sizeof(double) * 15 * (size_t)s->nz, s->ctx->stream));
        }
        int rc = hpb_plasma_insitu_slice(s->ctx, sp.d, sp.insitu_radius, sp.d_insitu + islice, s->nz);
        if (rc) return rc;
    }
    return HPB_OK;
}

// MultiPlasma::ReorderParticles (Hipace.cpp:595): every reorder_period slices the species is sorted by
// cell into its second SoA and the two are swapped; the InitParticles lattice order is gone afterwards
int reorder_plasma(hpb_sim *s, int islice)
{
    for (auto &sp : s->plasmas) {
        if (sp.reorder_period <= 0 || islice % sp.reorder_period != 0 || sp.d.np == 0) continue;
        if (sp.capacity2 < sp.d.np) {
            for (int k = 0; k < HPB_PLASMA_NREAL; ++k) { cudaFree(sp.d2.r[k]); sp.d2.r[k] = nullptr; }
            cudaFree(sp.d2.idcpu); sp.d2.idcpu = nullptr;
            for (int k = 0; k < HPB_PLASMA_NREAL; ++k) SIM_CUDA(cudaMalloc(&sp.d2.r[k], sizeof(double) * sp.capacity));
            SIM_CUDA(cudaMalloc(&sp.d2.idcpu, sizeof(uint64_t) * sp.capacity));
            sp.capacity2 = sp.capacity;
        }
        sp.d2.np = sp.d.np;
        int rc = hpb_plasma_reorder(s->ctx, sp.d, sp.d2, s->prob_lo[0], s->prob_lo[1], sp.reorder_idx[0],
                                    sp.reorder_idx[1]);
        if (rc) return rc;
        std::swap(sp.d, sp.d2);
        std::swap(sp.capacity, sp.capacity2);
        sp.lattice_n = 0;
        s->stats.n_reorders += 1;
    }
    return HPB_OK;
}

int solve_one_slice(hpb_sim *s, int islice)
{
    if (int rcr = reorder_plasma(s, islice)) return rcr;
    if (int rci = plasma_insitu(s, islice)) return rci;
    if (!s->explicit_solver) return solve_one_slice_pc(s, islice);
    hpb_ctx *ctx = s->ctx;
    const int *C = s->comps;
    int rc;
    // (the fused push + deposit exists for the default order only; the grid current is added
    // between the beam deposit and the Poisson solve of the reference order)
    bool fuse = s->opt_fuse && C[HPB_C_RHO] < 0 && !s->use_laser && !hpb_use_generic_order(ctx)
                && !s->use_grid_current;
    // (plasma in-situ diagnostics see the particles before this slice's deposit may discard QSA
    // violators: the fused order has deposited already)
    for (auto &sp : s->plasmas) if (sp.insitu_period > 0) fuse = false;
    for (auto &sp : s->plasmas) if (sp.n_subcycles < 1) fuse = false;
    // (stage timers synchronise the main stream per stage: keep one stream when profiling)
    const bool side = fuse && s->opt_side_stream && s->stream2 && !s->beams.empty() && !s->opt_profile;
    const bool was_prepared = s->prepared;
    if (!was_prepared) {
        {
            StageTimer t(s, ST_OTHER);
            if ((rc = hpb_fields_initialize_slices(ctx, s->sl, C))) return rc;          // :598-600
        }
        if (s->use_laser) {                                                             // :583, :603
            StageTimer t(s, ST_OTHER);
            const double z = islice * s->g.dz + (s->prob_lo[2] + 0.5 * s->g.dz);        // GetPosOffset(2)
            if (s->laser_state) {
                // the stored A^n, A^{n-1} of this slice come from the upstream rank in a pipeline
                if ((rc = hpb_pipeline_wait_slice_on(s, islice, ctx->stream))) return rc;
                if ((rc = hpb_laser_get_slice(s->laser_state, ctx, s->sl, C[HPB_C_AABS], islice, s->cur_step, z,
                                              s->opt_checksums ? s->d_checksum + s->sl.ncomp : nullptr,
                                              s->diag_xz))) return rc;
            } else if ((rc = hpb_laser_update_aabs(ctx, s->sl, C[HPB_C_AABS], s->lasers.data(),
                                                   (int)s->lasers.size(), s->laser_lambda0, s->laser_interp_order, z,
                                                   s->opt_checksums ? s->d_checksum + s->sl.ncomp : nullptr))) return rc;
        }
        {
            StageTimer t(s, ST_DEPOSIT);
            for (auto &sp : s->plasmas)                                                 // :609-610
                if ((rc = hpb_deposit_current_laser(ctx, sp.d, s->sl, sp.charge, sp.mass, C[HPB_C_JX],
                                                    C[HPB_C_JY], C[HPB_C_RHO], C[HPB_C_CHI],
                                                    C[HPB_C_RHOMJZ], C[HPB_C_AABS], sp.max_qsa,
                                                    s->d_nqsa))) return rc;
        }
        StageTimer t(s, ST_OTHER);
        if ((rc = beam_deposit_jz(s, islice))) return rc;
        if ((rc = hpb_fields_add_rho_ions(ctx, s->sl, C))) return rc;                   // :626
        if (s->use_grid_current)                                                        // :629
            if ((rc = hpb_fields_grid_current(ctx, s->sl, C[HPB_C_JZ_BEAM], s->gc_peak, s->gc_mean, s->gc_std,
                                              s->prob_lo[0], s->prob_lo[1],
                                              s->prob_lo[2] + islice * s->g.dz))) return rc;
    }
    s->prepared = false;
    // chi of this slice is final (the deposits are done; with periodic fields it is summed later, a laser's
    // multigrid reuses the level arrays): its coarse-level averages can be formed beside the next stages
    const bool early_acf = s->opt_mg_early && s->stream3 && !s->field_periodic && !s->use_laser && !s->opt_profile;
    if (early_acf) {
        SIM_CUDA(cudaEventRecord(s->ev_chi, s->stream));
        SIM_CUDA(cudaStreamWaitEvent(s->stream3, s->ev_chi, 0));
        {
            StreamScope sc(s, s->stream3);
            rc = hpb_mg_prepare_acf(ctx, s->sl, C[HPB_C_CHI]);
        }
        if (rc) return rc;
        SIM_CUDA(cudaEventRecord(s->ev_acf, s->stream3));
    }
    {
        StageTimer t(s, ST_POISSON);
        if (!s->field_periodic && !s->poisson_periodic) {
            if ((rc = hpb_fields_solve_psi_ez_bz(ctx, s->sl, C))) return rc;            // :633
        } else {
            // Fields.cpp:859-861, :920-922: the periodic images of the deposits are summed into the valid
            // box before, and the guard cells of the potentials filled after, the three solves
            const int src[3] = {C[HPB_C_JX], C[HPB_C_JY], C[HPB_C_RHOMJZ]};
            const int lhs[3] = {C[HPB_C_PSI], C[HPB_C_EZ], C[HPB_C_BZ]};
            if (s->field_periodic && (rc = hpb_fields_enforce_periodic(ctx, s->sl, 1, src, 3))) return rc;
            if ((rc = hpb_fields_psi_ez_bz_rhs(ctx, s->sl, C, s->d_pc_rhs))) return rc;
            if ((rc = (s->poisson_periodic ? hpb_poisson_solve_periodic : hpb_poisson_solve)(
                     ctx, s->d_pc_rhs, s->sl, lhs, 3))) return rc;
            if (s->field_periodic && (rc = hpb_fields_enforce_periodic(ctx, s->sl, 0, lhs, 3))) return rc;
            if ((rc = hpb_launch_exmby_eypbx(ctx, s->sl, C))) return rc;
        }
    }
    if (s->laser_state) {                                                               // :637
        StageTimer t(s, ST_OTHER);
        // (the slots A^{n+1}, A^n of this slice are written into may still be leaving for the downstream rank)
        if ((rc = hpb_pipeline_wait_out_slot_on(s, islice, ctx->stream))) return rc;
        if ((rc = hpb_laser_advance_slice(s->laser_state, ctx, s->sl, C[HPB_C_CHI], islice, s->dt, s->cur_step,
                                          s->prob_hi[0] - s->prob_lo[0], s->prob_hi[1] - s->prob_lo[1]))) return rc;
    }
    if (!was_prepared) {
        StageTimer t(s, ST_OTHER);
        if ((rc = beam_next_and_sxsy(s, islice))) return rc;
    }
    {
        StageTimer t(s, ST_EXPLICIT);
        for (auto &sp : s->plasmas) {                                                   // :663
            hpb_set_plasma_lattice_hint(ctx, sp.lattice_n, sp.lattice_ppc);
            if ((rc = hpb_explicit_deposition(ctx, sp.d, s->sl, sp.charge, sp.mass, C))) return rc;
        }
        hpb_set_plasma_lattice_hint(ctx, 0, 0);
    }
    {
        StageTimer t(s, ST_MG);
        int iters = 0;                                                                  // :666
        const int srcs[3] = {C[HPB_C_SY], C[HPB_C_SX], C[HPB_C_CHI]};                   // Hipace.cpp:817-821
        if (s->field_periodic && (rc = hpb_fields_enforce_periodic(ctx, s->sl, 1, srcs, 3))) return rc;
        if (early_acf) SIM_CUDA(cudaStreamWaitEvent(s->stream, s->ev_acf, 0));
        if ((rc = hpb_mg_solve1(ctx, s->sl, C[HPB_C_BX], C[HPB_C_SY], C[HPB_C_CHI], s->mg_tol_rel,
                                s->mg_tol_abs, 200, &iters))) return rc;
        const int bxy[2] = {C[HPB_C_BX], C[HPB_C_BY]};                                  // Hipace.cpp:924-927
        if (s->field_periodic && (rc = hpb_fields_enforce_periodic(ctx, s->sl, 0, bxy, 2))) return rc;
        s->stats.n_mg_vcycles += iters;
        s->mg_iters.push_back(iters);
    }
    if (do_diagnostics(s->field_insitu_period, s->cur_step, s->max_step)) {             // :685
        StageTimer t(s, ST_OTHER);
        if (!s->d_field_insitu) {
            SIM_CUDA(cudaMalloc(&s->d_field_insitu, sizeof(double) * 10 * (size_t)s->nz));
            SIM_CUDA(cudaMemsetAsync(s->d_field_insitu, 0, sizeof(double) * 10 * (size_t)s->nz, ctx->stream));
        }
        if ((rc = hpb_fields_insitu_slice(ctx, s->sl, C, s->d_field_insitu + islice, s->nz))) return rc;
    }
    if (s->laser_state && do_diagnostics(s->laser_insitu_period, s->cur_step, s->max_step)) {   // :688
        StageTimer t(s, ST_OTHER);
        if (!s->d_laser_insitu) {
            SIM_CUDA(cudaMalloc(&s->d_laser_insitu, sizeof(double) * 8 * (size_t)s->nz));
            SIM_CUDA(cudaMemsetAsync(s->d_laser_insitu, 0, sizeof(double) * 8 * (size_t)s->nz, ctx->stream));
        }
        if ((rc = hpb_laser_insitu_slice(s->laser_state, ctx, s->d_laser_insitu + islice, s->nz))) return rc;
    }
    if (s->opt_checksums) {                                                             // :681-691
        StageTimer t(s, ST_OTHER);
        if ((rc = slice_checksums(s))) return rc;
    }
    auto push_plasma = [&](bool with_deposit) -> int {
        StageTimer t(s, ST_PUSH);
        for (auto &sp : s->plasmas) {                                                   // :699-701
            s->stats.n_plasma_pushed += (double)sp.d.np;
            hpb_set_plasma_lattice_hint(ctx, sp.lattice_n, sp.lattice_ppc);
            int r;
            if (with_deposit)
                r = hpb_advance_plasma_particles_and_deposit(ctx, sp.d, s->sl, sp.charge, sp.mass,
                                                             sp.n_subcycles, s->particle_bc, s->bc_lo,
                                                             s->bc_hi, s->comps, sp.max_qsa, s->d_nqsa);
            else
                r = hpb_advance_plasma_particles(ctx, sp.d, s->sl, sp.charge, sp.mass, sp.n_subcycles,
                                                 0, s->particle_bc, s->bc_lo, s->bc_hi, s->comps);
            hpb_set_plasma_lattice_hint(ctx, 0, 0);
            if (r) return r;
        }
        return HPB_OK;
    };
    if (!fuse) {
        if ((rc = push_plasma(false))) return rc;
        StageTimer t(s, ST_OTHER);
        if ((rc = beam_push_and_send(s, islice))) return rc;
        if ((rc = hpb_fields_shift_slices(ctx, s->sl, C))) return rc;                   // :721
        if (s->laser_state) hpb_laser_shift_slices(s->laser_state);                     // :727
    } else if (side && s->opt_side_late) {
        // Host issue order matters as much as the stream graph: the plasma push (the longest kernel of the
        // slice, on the main stream) is enqueued BEFORE the side stream's dozen small beam launches and
        // the NCCL send / receive bookkeeping, so that the GPU is busy while the host works through them
        // (with the beam work enqueued first the main stream sat idle for the host time of those calls:
        // ~80 us per slice as soon as the pipeline hand-off was on, SCALE_r01).  The side-stream kernels
        // still read this slice's field planes: they get the component map from before the rotation.
        cudaStream_t bs = s->stream2;
        SIM_CUDA(cudaEventRecord(s->ev_fields, s->stream));
        SIM_CUDA(cudaStreamWaitEvent(bs, s->ev_fields, 0));
        int before[HPB_C_COUNT], after[HPB_C_COUNT];
        memcpy(before, s->comps, sizeof(before));
        if ((rc = hpb_fields_shift_and_initialize(ctx, s->sl, s->comps))) return rc;
        if (islice > 0) SIM_CUDA(cudaEventRecord(s->ev_shift, s->stream));
        if ((rc = push_plasma(true))) return rc;
        memcpy(after, s->comps, sizeof(after));
        memcpy(s->comps, before, sizeof(before));
        {
            StreamScope sc(s, bs);
            rc = beam_push_and_send(s, islice);
        }
        memcpy(s->comps, after, sizeof(after));
        if (rc) return rc;
        if (islice > 0) {       // the beam deposits of the next slice go into freshly zeroed planes
            SIM_CUDA(cudaStreamWaitEvent(bs, s->ev_shift, 0));
            StreamScope sc(s, bs);
            if ((rc = beam_deposit_jz(s, islice - 1))) return rc;
            if ((rc = beam_next_and_sxsy(s, islice - 1))) return rc;
        }
        SIM_CUDA(cudaEventRecord(s->ev_side, bs));
        SIM_CUDA(cudaStreamWaitEvent(s->stream, s->ev_side, 0));
        s->prepared = islice > 0;
    } else {
        cudaStream_t bs = side ? s->stream2 : s->stream;
        {
            StageTimer t(s, ST_OTHER);
            if (side) {     // fork: the beam push gathers this slice's fields
                SIM_CUDA(cudaEventRecord(s->ev_fields, s->stream));
                SIM_CUDA(cudaStreamWaitEvent(bs, s->ev_fields, 0));
            }
            {
                StreamScope sc(s, bs);
                if ((rc = beam_push_and_send(s, islice))) return rc;
            }
            if ((rc = hpb_fields_shift_and_initialize(ctx, s->sl, s->comps))) return rc;
            if (islice > 0) {
                if (side) {     // the beam deposits of the next slice go into freshly zeroed planes
                    SIM_CUDA(cudaEventRecord(s->ev_shift, s->stream));
                    SIM_CUDA(cudaStreamWaitEvent(bs, s->ev_shift, 0));
                }
                StreamScope sc(s, bs);
                if ((rc = beam_deposit_jz(s, islice - 1))) return rc;
                if ((rc = beam_next_and_sxsy(s, islice - 1))) return rc;
            }
        }
        if ((rc = push_plasma(true))) return rc;
        if (side) {         // join: whatever follows on the main stream sees the side stream's work
            SIM_CUDA(cudaEventRecord(s->ev_side, bs));
            SIM_CUDA(cudaStreamWaitEvent(s->stream, s->ev_side, 0));
        }
        s->prepared = islice > 0;
    }
    s->stats.n_cells_updated += (double)s->g.nx * s->g.ny;
    s->stats.n_slices += 1;
    if (fuse) s->stats.n_fused_slices += 1;
    return HPB_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
// Host-only: the adaptive time step arithmetic (adaptive_dt.hpp) for a uniform plasma density,
// so that the CPU tests can hold it to the oracle
extern "C" int hpb_adaptive_dt_next(const hpb_adaptive_par *par, int nbeams, const double *ts,
                                    const double *charge, const double *mass, double rho, double t_next,
                                    double dt_in, double *dt_out, double *min_uz_mq)
{
    if (!par || !ts || !charge || !mass || !dt_out || !min_uz_mq) return HPB_ERR_ARG;
    double dt = dt_in, mq = *min_uz_mq;
    if (!adaptive_dt_from_min_uz(*par, nbeams, ts, charge, mass, [rho](double) { return rho; }, 0., dt_in, dt, mq))
        return HPB_ERR_ARG;
    *dt_out = adaptive_dt_from_density(*par, mq, t_next, dt, [rho](double) { return rho; });
    *min_uz_mq = mq;
    return HPB_OK;
}

// Host-only: parse a deck exactly as hpb_sim_create does (same parser, same "unsupported"
// errors) and describe what it would run -- needs no GPU, so the input-deck surface is testable
// on any machine.  summary (may be NULL) receives "key=value;" pairs.
extern "C" int hpb_deck_check(const char *deck, const char *overrides, char *summary, size_t n)
{
    if (!deck) return HPB_ERR_ARG;
    hpb_sim s;
    try {
        s.deck.parse(deck);
        if (overrides) s.deck.parse(overrides);
        read_deck(&s);
    } catch (const std::exception &e) {
        hpb_set_error("deck: %s", e.what());
        return HPB_ERR_PARSE;
    }
    if (summary && n) {
        std::string o;
        char b[512];
        snprintf(b, sizeof b, "nx=%d;ny=%d;nz=%d;dx=%.17g;dy=%.17g;dz=%.17g;x_off=%.17g;y_off=%.17g;normalized=%d;"
                 "particle_bc=%d;max_step=%d;dt=%.17g;mg_tol_rel=%.17g;deposit_rho=%d;",
                 s.g.nx, s.g.ny, s.nz, s.g.dx, s.g.dy, s.g.dz, s.g.x_off, s.g.y_off, s.g.normalized,
                 s.particle_bc, s.max_step, s.dt, s.mg_tol_rel, (int)s.deposit_rho);
        o += b;
        snprintf(b, sizeof b, "n_plasmas=%zu;n_beams=%zu;n_lasers=%zu;", s.plasmas.size(), s.beams.size(),
                 s.lasers.size());
        o += b;
        for (size_t k = 0; k < s.plasmas.size(); ++k) {
            const Species &sp = s.plasmas[k];
            snprintf(b, sizeof b, "plasma%zu.name=%s;plasma%zu.charge=%.17g;plasma%zu.mass=%.17g;plasma%zu.ppc=%dx%d;"
                     "plasma%zu.neutralize=%d;plasma%zu.n_subcycles=%d;plasma%zu.density0=%.17g;", k,
                     sp.name.c_str(), k, sp.charge, k, sp.mass, k, sp.ppc[0], sp.ppc[1], k, (int)sp.neutralize,
                     k, sp.n_subcycles, k, sp.density0);
            o += b;
        }
        for (size_t k = 0; k < s.beams.size(); ++k) {
            const BeamSp &bm = s.beams[k];
            snprintf(b, sizeof b, "beam%zu.name=%s;beam%zu.charge=%.17g;beam%zu.mass=%.17g;beam%zu.ppc=%dx%dx%d;"
                     "beam%zu.profile=%d;beam%zu.n_subcycles=%d;beam%zu.external_fields=%d;", k, bm.name.c_str(),
                     k, bm.charge, k, bm.mass, k, bm.ppc[0], bm.ppc[1], bm.ppc[2], k, bm.profile, k,
                     bm.n_subcycles, k, (int)bm.use_ext);
            o += b;
        }
        snprintf(summary, n, "%s", o.c_str());
    }
    return HPB_OK;
}

extern "C" int hpb_sim_create(hpb_sim **out, const char *deck, const char *overrides, int device)
{
    if (!out || !deck) return HPB_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        hpb_set_error("hpb_sim_create: no CUDA device (this library has no CPU fallback)");
        return HPB_ERR_CUDA;
    }
    std::unique_ptr<hpb_sim> s(new hpb_sim());
    try {
        s->deck.parse(deck);
        if (overrides) s->deck.parse(overrides);
        read_deck(s.get());
    } catch (const std::exception &e) {
        hpb_set_error("deck: %s", e.what());
        return HPB_ERR_PARSE;
    }
    s->device = device;
    SIM_CUDA(cudaSetDevice(device));
    SIM_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    build_components(s.get());
    int rc = hpb_create(&s->ctx, &s->g, (void *)s->stream);
    if (rc) return rc;
    hpb_slice &sl = s->sl;
    if ((rc = hpb_set_deposition_order(s->ctx, s->depos_order, s->depos_dtype))) return rc;
    sl.lo_x = sl.lo_y = -s->ng;
    sl.nx_tot = s->g.nx + 2 * s->ng; sl.ny_tot = s->g.ny + 2 * s->ng;
    // rows padded to an even number of cells: every row and plane then starts on a 16-byte boundary, which
    // the TMA tensor maps of the particle kernels need (an odd grid -- the reference's recommended 2^n - 1
    // cells -- otherwise ran the round-1 push without the TMA patch)
    sl.jstride = (sl.nx_tot + 1) & ~1; sl.nstride = (long)sl.jstride * sl.ny_tot;
    SIM_CUDA(cudaMalloc(&sl.p, sizeof(double) * sl.nstride * sl.ncomp));
    SIM_CUDA(cudaMemset(sl.p, 0, sizeof(double) * sl.nstride * sl.ncomp));
    SIM_CUDA(cudaMalloc(&s->d_checksum, sizeof(double) * (sl.ncomp + 1)));     // + laserEnvelope
    SIM_CUDA(cudaMemset(s->d_checksum, 0, sizeof(double) * (sl.ncomp + 1)));
    if (!s->explicit_solver || s->field_periodic || s->poisson_periodic)
        SIM_CUDA(cudaMalloc(&s->d_pc_rhs, sizeof(double) * 3 * (size_t)s->g.nx * s->g.ny));
    if (!s->explicit_solver) {
        SIM_CUDA(cudaMalloc(&s->d_pc_scal, sizeof(double) * 40));
    }
    SIM_CUDA(cudaMalloc(&s->d_nqsa, sizeof(int)));
    SIM_CUDA(cudaMemset(s->d_nqsa, 0, sizeof(int)));
    SIM_CUDA(cudaMalloc(&s->d_count, sizeof(unsigned long long)));
    SIM_CUDA(cudaMemset(s->d_count, 0, sizeof(unsigned long long)));
    SIM_CUDA(cudaMalloc(&s->d_overflow, sizeof(int)));
    SIM_CUDA(cudaMemset(s->d_overflow, 0, sizeof(int)));
    SIM_CUDA(cudaMalloc(&s->d_slot_off, sizeof(long) * (size_t)(s->nz + 1)));
    for (auto &b : s->beams) {
        SIM_CUDA(cudaMalloc(&b.d_cs, 9 * sizeof(double)));
        SIM_CUDA(cudaMemset(b.d_cs, 0, 9 * sizeof(double)));
        if (b.use_ext) {
            const char *ex[6];
            for (int k = 0; k < 6; ++k) ex[k] = b.ext_expr[k].c_str();
            if ((rc = hpb_extfields_create_deck(&b.ext, ex, &s->deck))) return rc;     // my_constants.* visible
        }
    }
    SIM_CUDA(cudaEventCreate(&s->ev0));
    SIM_CUDA(cudaEventCreate(&s->ev1));
    {
        // highest priority: the side stream's small kernels must get CTA slots while the plasma
        // push (tens of thousands of CTAs, launched at about the same time) is being dispatched
        int least = 0, greatest = 0;
        SIM_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        SIM_CUDA(cudaStreamCreateWithPriority(&s->stream2, cudaStreamNonBlocking, greatest));
        SIM_CUDA(cudaStreamCreateWithPriority(&s->stream3, cudaStreamNonBlocking, greatest));
    }
    SIM_CUDA(cudaEventCreateWithFlags(&s->ev_fields, cudaEventDisableTiming));
    SIM_CUDA(cudaEventCreateWithFlags(&s->ev_shift, cudaEventDisableTiming));
    SIM_CUDA(cudaEventCreateWithFlags(&s->ev_side, cudaEventDisableTiming));
    SIM_CUDA(cudaEventCreateWithFlags(&s->ev_chi, cudaEventDisableTiming));
    SIM_CUDA(cudaEventCreateWithFlags(&s->ev_acf, cudaEventDisableTiming));
    s->pev.resize(4);
    for (auto &e : s->pev) SIM_CUDA(cudaEventCreate(&e));
    *out = s.release();
    return HPB_OK;
}

extern "C" void hpb_sim_destroy(hpb_sim *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    for (auto &sp : s->plasmas) {
        for (int k = 0; k < HPB_PLASMA_NREAL; ++k) cudaFree(sp.d.r[k]);
        cudaFree(sp.d.idcpu);
        for (int k = 0; k < HPB_PLASMA_NREAL; ++k) cudaFree(sp.d2.r[k]);
        cudaFree(sp.d2.idcpu);
        cudaFree(sp.d_insitu);
    }
    hpb_pipeline_destroy(s);
    for (auto &b : s->beams) {
        hpb_beam_rings_free(b);
        cudaFree(b.d_cs); cudaFree(b.d_stage); cudaFree(b.d_stage_off);
        cudaFree(b.d_insitu); cudaFree(b.d_ts);
        hpb_extfields_destroy(b.ext);
    }
    cudaFree(s->d_overflow); cudaFree(s->d_slot_off);
    cudaFree(s->d_pc_rhs); cudaFree(s->d_pc_scal); cudaFree(s->d_field_insitu);
    hpb_laser_state_destroy(s->laser_state);
    cudaFree(s->d_laser_insitu);
    cudaFree(s->sl.p); cudaFree(s->d_checksum); cudaFree(s->d_nqsa); cudaFree(s->d_count);
    cudaFree(s->d_flag); cudaFree(s->d_offs); cudaFree(s->d_cub);
    hpb_destroy(s->ctx);
    cudaEventDestroy(s->ev0); cudaEventDestroy(s->ev1);
    if (s->stream2) { cudaStreamSynchronize(s->stream2); cudaStreamDestroy(s->stream2); }
    if (s->ev_fields) cudaEventDestroy(s->ev_fields);
    if (s->ev_shift) cudaEventDestroy(s->ev_shift);
    if (s->ev_side) cudaEventDestroy(s->ev_side);
    if (s->ev_chi) cudaEventDestroy(s->ev_chi);
    if (s->ev_acf) cudaEventDestroy(s->ev_acf);
    if (s->stream3) { cudaStreamSynchronize(s->stream3); cudaStreamDestroy(s->stream3); }
    for (auto &e : s->pev) cudaEventDestroy(e);
    cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" int hpb_sim_begin_step(hpb_sim *s, int step)
{
    if (!s) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    return begin_step(s, step);
}

extern "C" int hpb_sim_solve_one_slice(hpb_sim *s, int islice)
{
    if (!s || islice < 0 || islice >= s->nz) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    int rc = solve_one_slice(s, islice);
    if (rc == HPB_OK && islice == 0) rc = end_step(s);
    return rc;
}

extern "C" int hpb_sim_evolve(hpb_sim *s, int step_begin, int step_end, int n_slices)
{
    if (!s) return HPB_ERR_ARG;
    if (step_end < 0) step_end = s->max_step;        // the whole run of the deck (Hipace.cpp:401)
    if (step_end < step_begin) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    s->stats = hpb_sim_stats();
    s->mg_iters.clear();
    SIM_CUDA(cudaMemsetAsync(s->d_nqsa, 0, sizeof(int), s->stream));
    SIM_CUDA(cudaMemsetAsync(s->d_count, 0, sizeof(unsigned long long), s->stream));
    const long launches0 = s->ctx->n_launch;
    double loop_ms = 0.;
    for (int step = step_begin; step <= step_end; ++step) {
        int rc = begin_step(s, step);
        if (rc) return rc;
        const int stop = (n_slices > 0) ? std::max(-1, s->nz - 1 - n_slices) : -1;
        SIM_CUDA(cudaEventRecord(s->ev0, s->stream));
        for (int isl = s->nz - 1; isl > stop; --isl) {               // Hipace.cpp:478-480
            rc = solve_one_slice(s, isl);
            if (rc) return rc;
        }
        SIM_CUDA(cudaEventRecord(s->ev1, s->stream));
        if (stop == -1 && (rc = end_step(s))) return rc;
        SIM_CUDA(cudaEventSynchronize(s->ev1));
        float ms = 0.f;
        SIM_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        loop_ms += ms;
    }
    s->stats.slice_loop_ms = loop_ms;
    s->stats.n_kernel_launches = s->ctx->n_launch - launches0;
    int nq = 0, ovf = 0;
    unsigned long long nbp = 0;
    SIM_CUDA(cudaMemcpyAsync(&nq, s->d_nqsa, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    SIM_CUDA(cudaMemcpyAsync(&ovf, s->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    SIM_CUDA(cudaMemcpyAsync(&nbp, s->d_count, sizeof(nbp), cudaMemcpyDeviceToHost, s->stream));
    SIM_CUDA(cudaStreamSynchronize(s->stream));
    s->stats.n_qsa_violation = nq;
    s->stats.n_beam_pushed = (double)nbp;
    if (ovf) {
        hpb_set_error("a beam slice outgrew its packet capacity (raise <beam>.slice_capacity)");
        return HPB_ERR_CAPACITY;
    }
    return HPB_OK;
}

extern "C" int hpb_sim_geometry(hpb_sim *s, int n_cell[3], double prob_lo[3], double prob_hi[3])
{
    if (!s) return HPB_ERR_ARG;
    if (n_cell) { n_cell[0] = s->g.nx; n_cell[1] = s->g.ny; n_cell[2] = s->nz; }
    for (int k = 0; k < 3; ++k) {
        if (prob_lo) prob_lo[k] = s->prob_lo[k];
        if (prob_hi) prob_hi[k] = s->prob_hi[k];
    }
    return HPB_OK;
}

extern "C" int hpb_sim_ncomp(hpb_sim *s) { return s ? s->sl.ncomp : -1; }
extern "C" int hpb_sim_nguard(hpb_sim *s) { return s ? s->ng : -1; }

extern "C" int hpb_sim_comp_index(hpb_sim *s, const char *which_slice, const char *name)
{
    if (!s || !which_slice || !name) return -1;
    for (size_t c = 0; c < s->comp_names.size(); ++c)
        if (s->comp_names[c].first == which_slice && s->comp_names[c].second == name) return (int)c;
    return -1;
}

extern "C" int hpb_sim_get_field(hpb_sim *s, int comp, double *h_out)
{
    if (!s || comp < 0 || comp >= s->sl.ncomp || !h_out) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    SIM_CUDA(cudaStreamSynchronize(s->stream));
    SIM_CUDA(cudaMemcpy2D(h_out, sizeof(double) * s->sl.nx_tot, s->sl.p + phys_comp(s, comp) * s->sl.nstride,
                          sizeof(double) * s->sl.jstride, sizeof(double) * s->sl.nx_tot, s->sl.ny_tot,
                          cudaMemcpyDeviceToHost));
    return HPB_OK;
}

extern "C" int hpb_sim_set_field(hpb_sim *s, int comp, const double *h_in)
{
    if (!s || comp < 0 || comp >= s->sl.ncomp || !h_in) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    SIM_CUDA(cudaStreamSynchronize(s->stream));
    SIM_CUDA(cudaMemcpy2D(s->sl.p + phys_comp(s, comp) * s->sl.nstride, sizeof(double) * s->sl.jstride, h_in,
                          sizeof(double) * s->sl.nx_tot, sizeof(double) * s->sl.nx_tot, s->sl.ny_tot,
                          cudaMemcpyHostToDevice));
    return HPB_OK;
}

extern "C" long hpb_sim_plasma_np(hpb_sim *s, int species)
{
    if (!s || species < 0 || species >= (int)s->plasmas.size()) return -1;
    return s->plasmas[species].d.np;
}

extern "C" int hpb_sim_get_plasma_real(hpb_sim *s, int species, int idx, double *h_out)
{
    if (!s || species < 0 || species >= (int)s->plasmas.size() || idx < 0 ||
        idx >= HPB_PLASMA_NREAL || !h_out) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    SIM_CUDA(cudaStreamSynchronize(s->stream));
    const Species &sp = s->plasmas[species];
    if (sp.d.np > 0)
        SIM_CUDA(cudaMemcpy(h_out, sp.d.r[idx], sizeof(double) * sp.d.np, cudaMemcpyDeviceToHost));
    return HPB_OK;
}

extern "C" int hpb_sim_get_plasma_valid(hpb_sim *s, int species, uint8_t *h_out)
{
    if (!s || species < 0 || species >= (int)s->plasmas.size() || !h_out) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    const Species &sp = s->plasmas[species];
    if (sp.d.np == 0) return HPB_OK;
    uint8_t *d = nullptr;
    SIM_CUDA(cudaMalloc(&d, sp.d.np));
    k_valid_bytes<<<nb256(sp.d.np), 256, 0, s->stream>>>(sp.d.idcpu, sp.d.np, d);
    SIM_CUDA(cudaStreamSynchronize(s->stream));
    SIM_CUDA(cudaMemcpy(h_out, d, sp.d.np, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return HPB_OK;
}

extern "C" int hpb_sim_checksum_count(hpb_sim *s)
{
    if (!s) return -1;
    int n = 0;
    for (auto &c : s->comp_names) n += (c.first == "This");
    return n + (s->use_laser ? 1 : 0);
}

extern "C" const char *hpb_sim_checksum_name(hpb_sim *s, int k)
{
    if (!s) return nullptr;
    int n = 0;
    for (auto &c : s->comp_names)
        if (c.first == "This") { if (n == k) return c.second.c_str(); ++n; }
    if (s->use_laser && n == k) return "laserEnvelope";
    return nullptr;
}

extern "C" int hpb_sim_get_checksums(hpb_sim *s, double *h_out)
{
    if (!s || !h_out) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    SIM_CUDA(cudaStreamSynchronize(s->stream));
    std::vector<double> all(s->sl.ncomp + 1);
    SIM_CUDA(cudaMemcpy(all.data(), s->d_checksum, sizeof(double) * (s->sl.ncomp + 1), cudaMemcpyDeviceToHost));
    int n = 0;
    for (size_t c = 0; c < s->comp_names.size(); ++c)
        if (s->comp_names[c].first == "This") h_out[n++] = all[c];
    if (s->use_laser) h_out[n++] = all[s->sl.ncomp];
    return HPB_OK;
}

extern "C" int hpb_sim_get_beam_checksums(hpb_sim *s, int beam, double h_out[9])
{
    if (!s || beam < 0 || beam >= (int)s->beams.size() || !h_out) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    BeamSp &b = s->beams[beam];
    if (!b.ring[0].base) { int rc = init_beam(s, b); if (rc) return rc; }
    // after a complete step with checksums on: the state the step's beam diagnostic saw (before
    // the push, Hipace.cpp:682-683); otherwise the beam as it currently sits in the ring
    if (!b.cs_valid) { int rc = hpb_beam_ring_checksum(s, b.ring[b.cur], b.d_cs); if (rc) return rc; }
    SIM_CUDA(cudaMemcpyAsync(h_out, b.d_cs, 9 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    SIM_CUDA(cudaStreamSynchronize(s->stream));
    // momenta are stored as proper velocities u c (BeamParticleContainerInit.cpp:52-61); the
    // reference's checksum reads them normalised to m c (openPMD-viewer): 1 in normalised units
    for (int k = 3; k <= 5; ++k) h_out[k] *= 1.0 / s->g.c;
    return HPB_OK;
}

extern "C" long hpb_sim_beam_np(hpb_sim *s, int beam)
{
    if (!s || beam < 0 || beam >= (int)s->beams.size()) return -1;
    if (cudaSetDevice(s->device) != cudaSuccess) return -1;
    BeamSp &b = s->beams[beam];
    if (!b.ring[0].base && init_beam(s, b) != HPB_OK) return -1;
    std::vector<long> off;
    if (hpb_beam_ring_counts(s, b.ring[b.cur], off) != HPB_OK) return -1;
    return off[s->nz];
}

extern "C" long hpb_sim_beam_slice_capacity(hpb_sim *s, int beam)
{
    if (!s || beam < 0 || beam >= (int)s->beams.size()) return -1;
    if (cudaSetDevice(s->device) != cudaSuccess) return -1;
    BeamSp &b = s->beams[beam];
    if (!b.ring[0].base && init_beam(s, b) != HPB_OK) return -1;
    return b.ring[0].cap;
}

// Host <-> device transfer of a whole beam (the role of MultiBuffer::get_data / put_data with
// host staging buffers, src/utils/MultiBuffer.cpp:444-609, and of beam.injection_type =
// from_file).  h_real: 7 arrays x y z w ux uy uz of np doubles; h_idcpu: np; h_slot_off: nz+1
// offsets, slot s = slice nz-1-s.  Asynchronous on the simulation stream when the host memory
// is pinned; hpb_sim_evolve orders itself after an upload.
extern "C" int hpb_sim_get_beam(hpb_sim *s, int beam, double *const h_real[7], uint64_t *h_idcpu,
                                long *h_slot_off)
{
    if (!s || beam < 0 || beam >= (int)s->beams.size() || !h_real) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    BeamSp &b = s->beams[beam];
    if (!b.ring[0].base) { int rc = init_beam(s, b); if (rc) return rc; }
    std::vector<long> off;
    int rc = hpb_beam_ring_counts(s, b.ring[b.cur], off);
    if (rc) return rc;
    if (h_slot_off) memcpy(h_slot_off, off.data(), sizeof(long) * (s->nz + 1));
    if ((rc = hpb_beam_ring_gather(s, b, b.ring[b.cur], off, h_real, h_idcpu))) return rc;
    SIM_CUDA(cudaStreamSynchronize(s->stream));
    return HPB_OK;
}

extern "C" int hpb_sim_set_beam(hpb_sim *s, int beam, const double *const h_real[7],
                                const uint64_t *h_idcpu, const long *h_slot_off)
{
    if (!s || beam < 0 || beam >= (int)s->beams.size() || !h_real || !h_idcpu || !h_slot_off)
        return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    BeamSp &b = s->beams[beam];
    long mx = 0;
    for (int k = 0; k < s->nz; ++k) {
        if (h_slot_off[k + 1] < h_slot_off[k]) { hpb_set_error("set_beam: slot offsets must not decrease"); return HPB_ERR_ARG; }
        mx = std::max(mx, h_slot_off[k + 1] - h_slot_off[k]);
    }
    if (!b.ring[0].base || b.ring[0].cap < mx) {
        SIM_CUDA(cudaStreamSynchronize(s->stream));
        const long cap = (long)s->deck.num(b.name + ".slice_capacity", (double)(mx + mx / 4 + 1024));
        if (cap < mx) { hpb_set_error("beam %s: slice_capacity %ld < largest slice %ld", b.name.c_str(), cap, mx); return HPB_ERR_CAPACITY; }
        int rc = hpb_beam_rings_alloc(s, b, cap);
        if (rc) return rc;
    }
    int rc = hpb_beam_ring_scatter(s, b, b.ring[b.cur], h_slot_off, h_real, h_idcpu);
    if (rc) return rc;
    b.initialised = true;
    b.from_host = true;
    b.cs_valid = false;
    return HPB_OK;
}

extern "C" int hpb_sim_get_time(hpb_sim *s, double *time, double *dt)
{
    if (!s) return HPB_ERR_ARG;
    if (time) *time = s->time;
    if (dt) *dt = s->dt_step;
    return HPB_OK;
}

extern "C" int hpb_sim_get_stats(hpb_sim *s, hpb_sim_stats *out)
{
    if (!s || !out) return HPB_ERR_ARG;
    *out = s->stats;
    return HPB_OK;
}

extern "C" long hpb_sim_get_mg_iters(hpb_sim *s, int *h_out, long n)
{
    if (!s) return -1;
    const long cnt = (long)s->mg_iters.size();
    if (h_out) for (long k = 0; k < cnt && k < n; ++k) h_out[k] = s->mg_iters[k];
    return cnt;
}

extern "C" int hpb_sim_timer_start(hpb_sim *s)
{
    if (!s) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    SIM_CUDA(cudaEventRecord(s->pev[2], s->stream));
    return HPB_OK;
}

extern "C" int hpb_sim_timer_stop(hpb_sim *s, double *ms)
{
    if (!s || !ms) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    SIM_CUDA(cudaEventRecord(s->pev[3], s->stream));
    SIM_CUDA(cudaEventSynchronize(s->pev[3]));
    float f = 0.f;
    SIM_CUDA(cudaEventElapsedTime(&f, s->pev[2], s->pev[3]));
    *ms = f;
    return HPB_OK;
}

extern "C" int hpb_sim_get_beam_packet(hpb_sim *s, int beam, int slot, void *h_out)
{
    if (!s || beam < 0 || beam >= (int)s->beams.size() || slot < 0 || slot >= s->nz || !h_out) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    BeamSp &b = s->beams[beam];
    if (!b.ring[0].base) { int rc = init_beam(s, b); if (rc) return rc; }
    const BeamRing &r = b.ring[b.cur];
    SIM_CUDA(cudaMemcpyAsync(h_out, r.packet(slot), r.msg_bytes(), cudaMemcpyDeviceToHost, s->stream));
    SIM_CUDA(cudaStreamSynchronize(s->stream));
    return HPB_OK;
}

extern "C" int hpb_sim_set_option(hpb_sim *s, const char *key, double value)
{
    if (!s || !key) return HPB_ERR_ARG;
    const std::string k(key);
    if (k == "checksums") s->opt_checksums = value != 0.;
    else if (k == "fuse") s->opt_fuse = value != 0.;
    else if (k == "generic_order_kernels") s->ctx->force_generic = value != 0.;
    else if (k == "side_stream") s->opt_side_stream = value != 0.;
    else if (k == "side_late") s->opt_side_late = value != 0.;
    else if (k == "mg_early") s->opt_mg_early = value != 0.;
    else if (k == "beam_from_host") { for (auto &b : s->beams) b.from_host = value != 0.; }   // 0: step 0 re-creates the deck's beam
    else if (k == "profile") s->opt_profile = value != 0.;
    else if (k == "max_step") s->max_step = (int)value;
    else return hpb_set_option(s->ctx, key, value);      // kernel variants: the context's switches
    return HPB_OK;
}
