/* hpb200.h -- C-ABI of the B200-native per-zeta-slice quasi-static PIC hot path.
 *
 * Two layers:
 *
 *  (1) KERNEL SEAMS.  One entry point per reference call seam of Hipace::SolveOneSlice
 *      (src/Hipace.cpp:556-728).  All pointers are DEVICE pointers, every call only enqueues work
 *      on the stream given to hpb_create(), returns 0 on success / non-zero on error (the
 *      reference aborts instead, amrex::Abort), and never synchronises unless stated.
 *      The caller owns field and particle memory (as Fields::m_slices / the AMReX particle
 *      tiles do in the reference); the context owns solver scratch.
 *
 *  (2) SLICE-LOOP DRIVER.  hpb_sim_* runs a HiPACE++ input deck through the same call order
 *      as Hipace::Evolve (src/Hipace.cpp:393-554) with HOST buffers at the boundary; this is
 *      what the command-line driver and bench.py's end-to-end leg call.
 *
 * Memory layouts (identical to what an AMReX build of the reference holds):
 *   slice array : component-major, p[(i - lo_x) + (j - lo_y)*jstride + n*nstride], box
 *                 [-g, n-1+g]^2 with g = (depos_order_xy+1)/2 + 1 guard cells, 2 by default (src/utils/GPUUtil.H:98-182,
 *                 src/fields/Fields.cpp:63-64,169-174)
 *   plasma      : pure SoA in PlasmaIdx order (src/particles/plasma/PlasmaParticleContainer.H:
 *                 21-46): x y w ux uy psi x_prev y_prev ux_half_step uy_half_step psi_half_step,
 *                 plus a 64-bit idcpu whose top bit is the validity flag (AMReX >= 24 packing)
 *   beam slice  : SoA x y z w ux uy uz + idcpu (src/particles/beam/BeamParticleContainer.H)
 */
#ifndef HPB200_H_
#define HPB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPB_OK 0
#define HPB_ERR_ARG 1
#define HPB_ERR_CUDA 2
#define HPB_ERR_UNSUPPORTED 3
#define HPB_ERR_MG_DIVERGED 4      /* hpmg "failing so lets stop here", HpMultiGrid.cpp:1399-1416 */
#define HPB_ERR_PARSE 5
#define HPB_ERR_CAPACITY 6         /* a beam slice outgrew its packet (slipped particles) */
#define HPB_ERR_NCCL 7

#define HPB_NGUARD 2               /* (depos_order_xy+1)/2 + 1 for the default order 2, Fields.cpp:63-64 */
#define HPB_NGUARD_OF(order) (((order) + 1) / 2 + 1)
#define HPB_PLASMA_NREAL 11

/* PlasmaIdx, PlasmaParticleContainer.H:21-46 */
enum { HPB_X = 0, HPB_Y, HPB_W, HPB_UX, HPB_UY, HPB_PSI, HPB_X_PREV, HPB_Y_PREV,
       HPB_UX_HALF, HPB_UY_HALF, HPB_PSI_HALF };

/* ParticleBoundary, src/Hipace.H */
enum { HPB_BC_REFLECTING = 0, HPB_BC_PERIODIC = 1, HPB_BC_ABSORBING = 2 };

typedef struct hpb_ctx hpb_ctx;     /* solver scratch, plans, stream; one per (nx, ny) */
typedef struct hpb_sim hpb_sim;     /* slice-loop driver state */

/* Array3 view of the slice MultiFab (src/utils/GPUUtil.H:98-147) */
typedef struct {
    double *p;                      /* device */
    int lo_x, lo_y;                 /* -g, -g */
    int nx_tot, ny_tot;             /* nx + 2g, ny + 2g */
    long jstride, nstride;          /* nx_tot, nx_tot*ny_tot (or padded) */
    int ncomp;
} hpb_slice;

/* plasma particle tile: ParticleTileData of PlasmaParticleContainer */
typedef struct {
    double *r[HPB_PLASMA_NREAL];    /* device, PlasmaIdx order */
    uint64_t *idcpu;                /* device */
    long np;
} hpb_plasma;

/* per-slice beam tile (BeamIdx: x y z w ux uy uz).  The particle counts may live on the device
 * so that a slice can be pushed, re-binned and handed to the next rank without a host round
 * trip: d_np (device, may be NULL) points to {number of particles without slipped ones, number
 * including slipped ones} (BeamParticleContainer.H:175-181: getNumParticles /
 * getNumParticlesIncludingSlipped); np is then only the capacity the kernels are launched for.
 * With d_np == NULL both counts are np. */
typedef struct {
    double *x, *y, *z, *w, *ux, *uy, *uz;   /* device */
    uint64_t *idcpu;                         /* device */
    long np;
    const int64_t *d_np;                     /* device int64[2] or NULL */
} hpb_beam_slice;

/* beams.external_E(x,y,z,t) / external_B(x,y,z,t) (BeamParticleContainer.cpp:71-88,
 * ExternalFields.H:29-58): six expressions Ex Ey Ez Bx By Bz compiled to device byte-code */
typedef struct hpb_extfields hpb_extfields;

/* geometry + PhysConst (src/fields/Fields.H:71-77 GetPosOffset, src/utils/Constants.H:54-81) */
typedef struct {
    int nx, ny;
    double dx, dy, dz;
    double x_off, y_off;            /* x = i*dx + x_off, GetPosOffset with the grown fab box */
    double c, ep0, mu0, q_e, m_e;
    int normalized;                 /* hipace.normalized_units */
} hpb_geom;

/* ---------------------------------------------------------------------------------------- */
/* (1) kernel seams                                                                         */
/* ---------------------------------------------------------------------------------------- */

/* stream: a cudaStream_t passed as void* (NULL = legacy default stream).  Replaces the solver
 * objects built in Fields::AllocData (Fields.cpp:179-208) and Hipace::ExplicitMGSolveBxBy
 * (Hipace.cpp:914-918). */
int hpb_create(hpb_ctx **out, const hpb_geom *geom, void *stream);
void hpb_destroy(hpb_ctx *ctx);
const char *hpb_last_error(void);
const char *hpb_version(void);

/* ::DepositCurrent (src/particles/deposition/PlasmaDepositCurrent.cpp:22-257).  Component
 * indices < 0 mean "do not deposit" exactly as the reference's -1 (:53-58).  d_n_qsa_violation:
 * device int incremented once per particle killed by the QSA check (:197-204); may be NULL. */
int hpb_deposit_current(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge, double mass,
                        int c_jx, int c_jy, int c_rho, int c_chi, int c_rhomjz,
                        double max_qsa_weighting_factor, int *d_n_qsa_violation);

/* the same with a laser: c_aabs is the |a|^2 plane gathered for gamma/psi (:182-195); < 0 = none */
int hpb_deposit_current_laser(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge, double mass,
                              int c_jx, int c_jy, int c_rho, int c_chi, int c_rhomjz, int c_aabs,
                              double max_qsa_weighting_factor, int *d_n_qsa_violation);

/* Laser envelope at time step 0 (first part): MultiLaser::InitLaserSlice, gaussian
 * branch (src/laser/MultiLaser.cpp:881-917) evaluated on the field grid (the default laser
 * geometry, :58-118) + MultiLaser::UpdateLaserAabs (:214-291): component c_aabs of the slice
 * receives |a|^2 interpolated with order interp_order, guard cells included.  lasers: HOST array.
 * d_envelope_abs_sum (device, may be NULL) accumulates sum |a| over the slice (the laserEnvelope
 * checksum).  With comps[HPB_C_AABS] >= 0, hpb_explicit_deposition and
 * hpb_advance_plasma_particles apply the ponderomotive terms of ExplicitDeposition.cpp:167-250
 * and PushPlasmaParticles.H:59-69.  The envelope ADVANCE: hpb_laser_state_* below (fft solver). */
typedef struct {
    double a0, w0, cep, propagation_angle_yz, pft_yz, L0, focal_distance, position_mean[3];
} hpb_laser;
#define HPB_MAX_LASERS 4
int hpb_laser_update_aabs(hpb_ctx *ctx, hpb_slice sl, int c_aabs, const hpb_laser *lasers, int nlasers,
                          double lambda0, int interp_order, double z_slice, double *d_envelope_abs_sum);

/* beam ::DepositCurrentSlice (src/particles/deposition/BeamDepositCurrent.cpp:21-195) */
int hpb_beam_deposit(hpb_ctx *ctx, hpb_beam_slice bm, hpb_slice sl, double charge,
                     int c_jx, int c_jy, int c_jz);

/* AdvanceBeamParticlesSlice (src/particles/pusher/BeamParticleAdvance.cpp:19-336; level 0, no
 * radiation reaction, no spin).  Pushes the d_np[1] particles of the slice (slipped ones
 * included) through n_subcycles sub-steps of dt/n_subcycles, stopping a particle once
 * z < min_z (:147-152).  The first d_np[0] particles start at sub-cycle 0 (BeamIdx::nsubcycles
 * is zeroed when a slice is unpacked, MultiBuffer.cpp:897-910), the slipped ones continue from
 * their own counter d_nsubcycles[ip]; the counter reached is stored back for every particle.
 * ext may be NULL.  d_class_counts (device, 2 ints per 256 particles of capacity, may be NULL)
 * receives per block the number of valid particles that stay (z >= min_z) / slipped -- the
 * input of hpb_beam_shift_slipped.  d_checksum (device, 9 doubles, may be NULL) accumulates
 * sum|x| |y| |z| |ux| |uy| |uz| |w|, sum of ids and the count of the d_np[0] particles BEFORE
 * the push (what the reference's beam diagnostic writes, Hipace.cpp:682-683). */
int hpb_extfields_create(hpb_extfields **out, const char *const expr[6]);
void hpb_extfields_destroy(hpb_extfields *ext);
int hpb_advance_beam_particles(hpb_ctx *ctx, hpb_beam_slice bm, int *d_nsubcycles, hpb_slice sl,
                               double charge, double mass, int n_subcycles, double dt,
                               double time, double min_z, int do_z_push, int particle_bc,
                               const double bc_lo[2], const double bc_hi[2], const int *comps,
                               const hpb_extfields *ext, int *d_class_counts,
                               double *d_checksum);

/* In-situ beam diagnostics.  hpb_beam_insitu_slice = the reduction of
 * BeamParticleContainer::InSituComputeDiags (src/particles/beam/BeamParticleContainer.cpp:476-557)
 * for one beam slice: d_record[k * stride] += the k-th of the 23 raw sums (sum w, sum w x, sum w x^2,
 * ..., sum w gamma^2, Np; order of that file), i.e. with d_record = base + islice and stride =
 * n_slices the array [23][n_slices] that hpb_insitu_write_beam takes.
 * hpb_insitu_write_beam (HOST only, no GPU) = InSituWriteToFile (:596-732): normalises by sum(w),
 * forms the averages / totals over slices and appends one record in the NumPy-structured format of
 * src/utils/InsituUtil.H, which the reference's tools/read_insitu_diagnostics.py reads unchanged. */
int hpb_beam_insitu_slice(hpb_ctx *ctx, hpb_beam_slice bm, double insitu_radius, double *d_record,
                          long stride);
int hpb_insitu_write_beam(const char *path, double time, int step, int n_slices, double charge,
                          double mass, double z_lo, double z_hi, double normalized_density_factor,
                          int is_normalized_units, const double *h_sums);
/* the same for a plasma species at the START of a slice (Hipace.cpp:587):
 * PlasmaParticleContainer::InSituComputeDiags / InSituWriteToFile
 * (src/particles/plasma/PlasmaParticleContainer.cpp:443-526, 530-618); 15 raw sums per slice */
int hpb_plasma_insitu_slice(hpb_ctx *ctx, hpb_plasma pl, double insitu_radius, double *d_record,
                            long stride);
int hpb_insitu_write_plasma(const char *path, double time, int step, int n_slices, double charge,
                            double mass, double z_lo, double z_hi, double normalized_density_factor,
                            int is_normalized_units, const double *h_sums);
/* and for the fields of a slice once they are all computed (Hipace.cpp:685): Fields::InSituComputeDiags
 * / InSituWriteToFile (src/fields/Fields.cpp:1289-1428); 10 raw sums per slice, explicit solver only */
int hpb_fields_insitu_slice(hpb_ctx *ctx, hpb_slice sl, const int *comps, double *d_record, long stride);
int hpb_insitu_write_fields(const char *path, double time, int step, int n_slices, double z_lo,
                            double z_hi, int is_normalized_units, double dxdydz, const double *h_sums);
/* hipace.dt = adaptive (src/utils/AdaptiveTimeStep.cpp).  hpb_beam_min_uz_slice = GatherMinUzSlice
 * (:108-141) of one pushed beam slice: d_acc[4] = {min uz/c, sum w, sum w uz/c, sum w uz^2/c^2}
 * accumulated over the slices of a step.  hpb_adaptive_dt_next (HOST only) = CalculateFromMinUz
 * (:143-233) followed by CalculateFromDensity (:315-369) for a uniform plasma charge density rho. */
typedef struct {
    double nt_per_betatron, dt_max, threshold_uz, phase_tolerance;
    int phase_substeps, control_phase;
    double c, ep0;
    int numprocs, predict_step;     /* ranks of the time-step pipeline; hipace.adaptive_predict_step */
} hpb_adaptive_par;
int hpb_beam_min_uz_slice(hpb_ctx *ctx, hpb_beam_slice bm, double *d_acc);
int hpb_adaptive_dt_next(const hpb_adaptive_par *par, int nbeams, const double *ts, const double *charge,
                         const double *mass, double rho, double t_next, double dt_in, double *dt_out,
                         double *min_uz_mq);
/* Laser envelope ADVANCE over time steps with the fft solver (second part):
 * MultiLaser::AdvanceSliceFFT (src/laser/MultiLaser.cpp:609-801), InterpolateChi (:334-407),
 * UpdateLaserAabs (:214-291), ShiftLaserSlices (:180-212) and the hand-over of A^{n+1}, A^n to the next
 * time step (src/utils/MultiBuffer.cpp:840-851, 913-923).  The state owns the nine complex work
 * slices, 4 x nz stored slice planes and the solver scratch; laser grid = field grid.
 *   begin_step: empty work slices; h_chi_initial = nx*ny host values of the unperturbed chi (:293-332)
 *   get_slice:  the envelope of slice islice (step 0: the analytic pulse; later: from the store), |a|^2
 *               into component c_aabs, sum |A| of the diagnostic into d_env_abs_sum (may be NULL)
 *   advance_slice: after the Poisson solves of the slice (Hipace.cpp:637); no-op for dt = 0
 *   shift_slices: end of the slice (Hipace.cpp:727); end_step: after the last slice */
typedef struct hpb_laser_state hpb_laser_state;
int hpb_laser_state_create(hpb_laser_state **out, hpb_ctx *ctx, int nz, const hpb_laser *lasers, int nlasers,
                           double lambda0, int interp_order, int use_phase);
void hpb_laser_state_destroy(hpb_laser_state *st);
int hpb_laser_begin_step(hpb_laser_state *st, hpb_ctx *ctx, const double *h_chi_initial);
int hpb_laser_get_slice(hpb_laser_state *st, hpb_ctx *ctx, hpb_slice sl, int c_aabs, int islice, int step,
                        double z_slice, double *d_env_abs_sum, int diag_xz);
int hpb_laser_advance_slice(hpb_laser_state *st, hpb_ctx *ctx, hpb_slice sl, int c_chi, int islice, double dt,
                            int step, double prob_len_x, double prob_len_y);
int hpb_laser_shift_slices(hpb_laser_state *st);
/* lasers.solver_type (0: fft, MultiLaser::AdvanceSliceFFT; 1: multigrid, AdvanceSliceMG on hpmg type 2,
 * the reference's default), lasers.MG_tolerance_rel / MG_tolerance_abs / MG_average_rhs
 * (src/laser/MultiLaser.cpp:40-56, 429-607) */
int hpb_laser_set_solver(hpb_laser_state *st, int use_multigrid, double tol_rel, double tol_abs, int average_rhs);
long hpb_laser_mg_vcycles(hpb_laser_state *st);
/* hpmg::MultiGrid::solve2 (src/mg_solver/HpMultiGrid.cpp:1192-1296): lap(A) - (a_r + i a_i) A = rhs with an
 * array real coefficient and a scalar imaginary one; planar [2][ny][nx] arrays over the valid box,
 * d_sol2 holds the initial guess on entry */
int hpb_mg_solve2(hpb_ctx *ctx, double *d_sol2, const double *d_rhs2, const double *d_acf_r, double acf_i,
                  double tol_rel, double tol_abs, int max_iters, int *h_iters);
/* MultiLaser::InSituComputeDiags / InSituWriteToFile (src/laser/MultiLaser.cpp:923-1075): 8 raw values per
 * slice (max |a|^2, [|a|^2], its x, x^2, y, y^2 moments, the on-axis sum re / im) and the host writer */
int hpb_laser_insitu_slice(hpb_laser_state *st, hpb_ctx *ctx, double *d_record, long stride);
int hpb_insitu_write_laser(const char *path, double time, int step, int n_slices, double z_lo, double z_hi,
                           int is_normalized_units, double dxdydz, int nx, int ny, const double *h_sums);
int hpb_laser_end_step(hpb_laser_state *st);
/* shiftSlippedParticles (src/particles/sorting/SliceSort.cpp:13-67) fused with the packing of
 * MultiBuffer::put_data (src/utils/MultiBuffer.cpp:730-905): invalid particles are dropped, the
 * particles with z >= min_z go (stable order) to `stay` whose counts d_stay_np[0..1] are set,
 * the others are appended to `next` behind its d_np[0] particles and d_next_np[1] is updated
 * (next.idcpu == NULL: dropped, the reference's behaviour below the last slice).
 * d_class_counts is the array filled by hpb_advance_beam_particles.  *d_overflow (device) is set
 * to 1 if a destination capacity (stay.np / next.np) is exceeded. */
int hpb_beam_shift_slipped(hpb_ctx *ctx, hpb_beam_slice bm, const int *d_nsubcycles, double min_z,
                           const int *d_class_counts, hpb_beam_slice stay, int64_t *d_stay_np,
                           hpb_beam_slice next, int64_t *d_next_np, int *d_next_nsubcycles,
                           int *d_overflow);

/* Fields::InitializeSlices / AddRhoIons / ShiftSlices (src/fields/Fields.cpp:535-615):
 * comps[] is the component table in the order of hpb_comp below. */
int hpb_fields_initialize_slices(hpb_ctx *ctx, hpb_slice sl, const int *comps);
int hpb_fields_add_rho_ions(hpb_ctx *ctx, hpb_slice sl, const int *comps);
/* GridCurrent::DepositCurrentSlice (src/utils/GridCurrent.cpp:25-70): jz_beam += peak *
 * exp(-((x-mean_x)/std_x)^2/2 - ((y-mean_y)/std_y)^2/2 - ((z-mean_z)/std_z)^2/2) on the valid box,
 * x = plo_x + (i + 1/2) dx, y likewise, z = prob_lo_z + islice dz (passed in) */
int hpb_fields_grid_current(hpb_ctx *ctx, hpb_slice sl, int c_jz_beam, double peak_current_density,
                            const double position_mean[3], const double position_std[3],
                            double plo_x, double plo_y, double z);
int hpb_fields_shift_slices(hpb_ctx *ctx, hpb_slice sl, const int *comps);
/* setVal(0., ...) on up to 12 components (entries < 0 are skipped), grown box */
int hpb_fields_zero(hpb_ctx *ctx, hpb_slice sl, const int *comp_list, int n);

/* ---- predictor-corrector Bx/By solver and open field boundaries ----------------
 * hpb_deposit_current_jz: ::DepositCurrent with every destination incl. jz
 *     (src/particles/deposition/PlasmaDepositCurrent.cpp:53-58, :223).
 * hpb_fields_bxby_rhs: the two right-hand sides of Fields::SolvePoissonBxBy
 *     (src/fields/Fields.cpp:1008-1078) into d_stage[2][ny][nx]; comps needs JZ, PREV_JX/JY, NEXT_JX/JY.
 * hpb_fields_psi_ez_bz_rhs: the three right-hand sides of Fields.cpp:886-912 into d_stage[3][ny][nx]
 *     (the Dirichlet path assembles them inside hpb_fields_solve_psi_ez_bz; the Open path needs them
 *     in memory to fold the boundary values in).
 * hpb_fields_open_boundary: Fields::SetBoundaryCondition for boundary.field = Open
 *     (src/fields/Fields.cpp:685-738, SetDirichletBoundaries :628-673, src/fields/OpenBoundary.H): the
 *     multipole moments of one staging plane and the free-space potential one cell outside every
 *     edge, folded into the plane in place; monopole = 0 for Ez and Bz.  d_moments: 38 doubles.
 * hpb_fields_rel_b_error: Fields::ComputeRelBFieldError (:1227-1286): d_out2 = {sum |B_a|,
 *     sum |B_a - B_b|} over the valid box; c_* = the Bx component of an adjacent (Bx, By) pair.
 * hpb_fields_lincomb2: MultiFab::LinComb on an adjacent component pair over the grown box,
 *     dst = fa * A + fb * B (InitialBfieldGuess :1149-1170, MixAndShiftBfields :1172-1225, shifts). */
int hpb_deposit_current_jz(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge, double mass,
                           int c_jx, int c_jy, int c_jz, int c_rho, int c_chi, int c_rhomjz, int c_aabs,
                           double max_qsa_weighting_factor, int *d_n_qsa_violation);
int hpb_fields_bxby_rhs(hpb_ctx *ctx, hpb_slice sl, const int *comps, double *d_stage);
int hpb_fields_psi_ez_bz_rhs(hpb_ctx *ctx, hpb_slice sl, const int *comps, double *d_stage);
int hpb_fields_open_boundary(hpb_ctx *ctx, double *d_rhs, int monopole, double prob_lo_x,
                             double prob_hi_x, double prob_lo_y, double prob_hi_y, double *d_moments);
int hpb_fields_rel_b_error(hpb_ctx *ctx, hpb_slice sl, int c_bx_a, int c_bx_b, double *d_out2);
int hpb_fields_lincomb2(hpb_ctx *ctx, hpb_slice sl, int c_dst, double fa, int c_a, double fb, int c_b);

/* FFTPoissonSolver::SolvePoissonEquation (src/fields/fft_poisson_solver/FFTPoissonSolver.H:26-57,
 * ...DirichletFast.cpp:286-328): laplace(lhs) = rhs with lhs = 0 at the first guard cell.
 * d_rhs: nbatch contiguous nx*ny staging areas; c_lhs[b]: destination component of solve b. */
int hpb_poisson_solve(hpb_ctx *ctx, const double *d_rhs, hpb_slice sl, const int *c_lhs,
                      int nbatch);

/* fields.poisson_solver = FFTPeriodic: FFTPoissonSolverPeriodic::SolvePoissonEquation
 * (src/fields/fft_poisson_solver/FFTPoissonSolverPeriodic.cpp:111-149; inv_k2 of :69-92, which is zero
 * on the whole kx = 0 row and ky = 0 column).  Same arguments as hpb_poisson_solve. */
int hpb_poisson_solve_periodic(hpb_ctx *ctx, const double *d_rhs, hpb_slice sl, const int *c_lhs,
                               int nbatch);

/* boundary.field = Periodic: Fields::EnforcePeriodic (src/fields/Fields.cpp:1117-1145).  do_sum != 0:
 * AMReX SumBoundary -- the guard cells are added to their periodic images in the valid box (and keep
 * their own values); do_sum == 0: FillBoundary -- the guard cells are overwritten with their periodic
 * images.  comp_list: up to 12 component indices (entries < 0 are skipped). */
int hpb_fields_enforce_periodic(hpb_ctx *ctx, hpb_slice sl, int do_sum, const int *comp_list, int n);

/* Fields::SolvePoissonPsiExmByEypBxEzBz (src/fields/Fields.cpp:840-957): RHS assembly + three
 * Poisson solves + ExmBy/EypBx stencil, fused. */
int hpb_fields_solve_psi_ez_bz(hpb_ctx *ctx, hpb_slice sl, const int *comps);

/* hpmg average_down_acoef (src/mg_solver/HpMultiGrid.cpp:1640-1700), which the reference runs inside
 * solve1 (:1177-1187), as a call of its own: the coefficient hierarchy depends on the coefficient plane
 * only, so a driver may enqueue it early (ctx stream at the time of the call; the caller orders it against
 * hpb_mg_solve1).  The next hpb_mg_solve1 with the same c_acf skips its own pass. */
int hpb_mg_prepare_acf(hpb_ctx *ctx, hpb_slice sl, int c_acf);

/* Hipace::InitializeSxSyWithBeam (src/Hipace.cpp:744-790) */
int hpb_fields_sxsy_from_beam(hpb_ctx *ctx, hpb_slice sl, const int *comps);

/* ::ExplicitDeposition (src/particles/deposition/ExplicitDeposition.cpp:20-263) */
int hpb_explicit_deposition(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge,
                            double mass, const int *comps);

/* hpmg::MultiGrid::solve1 (src/mg_solver/HpMultiGrid.H:64-66, .cpp:1169-1190): solves
 * laplace(sol) - acf*sol = rhs for the two adjacent components (c_sol, c_sol+1) with rhs
 * (c_rhs, c_rhs+1); sol holds the initial guess.  h_iters (host, may be NULL) receives the
 * number of V-cycles; this call synchronises the stream when h_iters != NULL. */
int hpb_mg_solve1(hpb_ctx *ctx, hpb_slice sl, int c_sol, int c_rhs, int c_acf,
                  double tol_rel, double tol_abs, int max_iters, int *h_iters);

/* AdvancePlasmaParticles (src/particles/pusher/PlasmaParticleAdvance.cpp:29-305) */
int hpb_advance_plasma_particles(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl, double charge,
                                 double mass, int n_subcycles, int temp_slice, int particle_bc,
                                 const double bc_lo[2], const double bc_hi[2], const int *comps);

/* OUR ADDITIONS (no single reference call; used by the slice-loop driver, the seams above stay
 * available and equivalent):
 * hpb_fields_shift_and_initialize = Fields::ShiftSlices (Fields.cpp:596-599) of this slice +
 *   Fields::InitializeSlices (:551-560) + AddRhoIons (:606-615) of the next one in ONE pass; the
 *   Previous / This / Next planes of jx_beam, jy_beam are rotated through `comps` (in/out)
 *   instead of copied.
 * hpb_advance_plasma_particles_and_deposit = AdvancePlasmaParticles of this slice fused with
 *   ::DepositCurrent (jx, jy, chi, rhomjz) of the next slice: the pushed particle is deposited
 *   from registers.  Call after hpb_fields_shift_and_initialize. */
int hpb_fields_shift_and_initialize(hpb_ctx *ctx, hpb_slice sl, int *comps);
/* Optional hint for the particle kernels that follow: the species' SoA holds `ppc` passes of
 * `cells_per_pass` particles each in InitParticles order (pass outermost, cells x-fastest,
 * PlasmaParticleContainerInit.cpp:189-316) with no particle filtered out.  The kernels then map
 * consecutive warps to the SAME 32 cells of successive passes, so the slice planes a warp gathers
 * from / reduces into are fetched from HBM once per slice instead of once per pass.  Results do
 * not depend on it; cells_per_pass = 0 clears the hint. */
int hpb_set_plasma_lattice_hint(hpb_ctx *ctx, long cells_per_pass, int ppc);
/* PlasmaParticleContainer::ReorderParticles (src/particles/plasma/PlasmaParticleContainer.cpp:196-208,
 * called every <plasma>.reorder_period slices, src/Hipace.cpp:595): counting sort of the particle
 * SoA by transverse cell (idx_type 0) or node (1) per direction (<plasma>.reorder_idx_type), x
 * fastest, invalid particles last.  `out` is a second SoA of the same capacity (the caller swaps
 * them afterwards); the particle set and every value are unchanged, the order inside a cell is
 * unspecified.  After a reorder the lattice hint above no longer holds. */
int hpb_plasma_reorder(hpb_ctx *ctx, hpb_plasma in, hpb_plasma out, double prob_lo_x, double prob_lo_y,
                       int idx_type_x, int idx_type_y);
/* test hook, host only: the thread -> particle map the push kernel uses with that hint (mode 0 linear,
 * 1 passes interleaved warp by warp, 2 CTA by CTA); out[warp * 32 + lane] = particle or -1 */
long hpb_debug_push_thread_map(long cells_per_pass, int ppc, int mode, long *out, long out_len);
/* Host-only: the coefficient table of the odd-prime FFT stage p in mma.m8n8k4 A-fragment order
 * (csrc/fft_smem.cuh: fft_prime_frag_table), for the CPU tests.  Copies min(out_len, count) doubles,
 * returns the count = ceil((h+1)/8) * ceil(h/4) * 64 with h = (p-1)/2. */
long hpb_debug_fft_prime_table(int p, double *out, long out_len);
/* hipace.depos_order_xy (0..3) and hipace.depos_derivative_type (0 analytic, 1 nodal, 2 centred) for
 * every particle kernel called with this context afterwards (Hipace.cpp:49-53; the reference selects
 * them at compile time through CompileTimeOptions, e.g. ExplicitDeposition.cpp:62-67).  Default 2 / 2.
 * The caller's slice must then carry (order_xy + 1) / 2 + 1 guard cells (Fields.cpp:63-64):
 * hpb_slice.lo_x = lo_y = -guards.  The default has warp-aggregated, staged kernels; every other
 * combination runs one-thread-per-particle kernels (csrc/generic_order.cu). */
int hpb_set_deposition_order(hpb_ctx *ctx, int order_xy, int derivative_type);
/* measurement infrastructure (bench.py): the fp64 FMA peak of `device` in TFLOP/s (2 flops per DFMA),
 * best of `reps` launches of a register-resident DFMA kernel -- the compute roof the particle kernels
 * are reported against beside the HBM roof */
int hpb_measure_fp64_peak(int device, int reps, double *tflops);
/* behaviour switches of the kernels (A/B measurements and cross-checks; the library never reads the
 * host's environment): "pdl" (programmatic dependent launch, process-wide, ctx may be NULL),
 * "generic" (generic-order kernels for the default order too), "order" (bit mask: which particle
 * kernels use the pass-interleaved thread map), "expl_variant", "push_variant" (launch geometry /
 * staging variants of the particle kernels), "fft_variant" (0: 5 CTAs per SM + the tensor-core prime
 * stage, 1: 6 CTAs per SM, 2: 6 CTAs per SM + the scalar prime stage), "mg_wide", "mg_fuse", "mg_rotate"
 * (level-0 buffer rotation, no final copy), "mg_lean" (lean interior-tile path of the tile smoother),
 * "mg_persist" (1 / 2: the mid levels of a V-cycle in one persistent launch, cooperative / plain --
 * measured slower, off), "poisson_impl" (1: the measurement arm that runs the reference's
 * FFTPoissonSolverDirichletFast sequence on cuFFT, csrc/ref_gpu_arm.cu).  All variants give the same
 * results (the mg_* ones bit-identical).  Unknown keys return HPB_ERR_ARG. */
int hpb_set_option(hpb_ctx *ctx, const char *key, double value);
int hpb_advance_plasma_particles_and_deposit(hpb_ctx *ctx, hpb_plasma pl, hpb_slice sl,
                                             double charge, double mass, int n_subcycles,
                                             int particle_bc, const double bc_lo[2],
                                             const double bc_hi[2], const int *comps,
                                             double max_qsa_weighting_factor,
                                             int *d_n_qsa_violation);

/* sum|Q| over the valid box of component c -> d_out[0] += ...   (checksum of
 * tests/checksum/backend/openpmd_backend.py:40-45, one slice at a time) */
int hpb_abs_sum(hpb_ctx *ctx, hpb_slice sl, int c, double *d_out);
/* n components in one launch: d_out[slots[k]] += sum|Q| of component comp_list[k] */
int hpb_abs_sum_multi(hpb_ctx *ctx, hpb_slice sl, const int *comp_list, const int *slots, int n,
                      double *d_out);
/* the same for diagnostic.diag_type = xz: the slice contributes its y = mid-domain line (order-1
 * interpolation: mean of the two central rows for even ny; src/diagnostics/Diagnostic.cpp:393-407) */
int hpb_abs_sum_xz(hpb_ctx *ctx, hpb_slice sl, int c, double *d_out);

/* component table passed as "comps" above (names of src/fields/Fields.cpp:70-122) */
enum hpb_comp {
    HPB_C_NEXT_JX_BEAM = 0, HPB_C_NEXT_JY_BEAM,
    HPB_C_CHI, HPB_C_SY, HPB_C_SX, HPB_C_EXMBY, HPB_C_EYPBX, HPB_C_EZ, HPB_C_BX, HPB_C_BY,
    HPB_C_BZ, HPB_C_PSI, HPB_C_JX_BEAM, HPB_C_JY_BEAM, HPB_C_JZ_BEAM, HPB_C_JX, HPB_C_JY,
    HPB_C_RHOMJZ, HPB_C_RHO /* -1 if not allocated */,
    HPB_C_PREV_JX_BEAM, HPB_C_PREV_JY_BEAM,
    HPB_C_IONS_RHOMJZ /* -1 if no neutralising background */,
    HPB_C_AABS /* |a|^2 of the laser envelope on the field grid, -1 without a laser (Fields.cpp:98-101) */,
    /* predictor-corrector solver only (Fields.cpp:124-163); -1 with the explicit solver */
    HPB_C_JZ, HPB_C_NEXT_JX, HPB_C_NEXT_JY, HPB_C_PREV_BX, HPB_C_PREV_BY, HPB_C_PREV_JX, HPB_C_PREV_JY,
    HPB_C_PCITER_BX, HPB_C_PCITER_BY, HPB_C_PCPREV_BX, HPB_C_PCPREV_BY,
    HPB_C_COUNT
};

/* ---------------------------------------------------------------------------------------- */
/* (2) slice-loop driver: Hipace::Hipace / InitData / Evolve / SolveOneSlice                */
/* ---------------------------------------------------------------------------------------- */

/* deck: HiPACE++ input-deck text (ParmParse syntax); overrides: extra "key = value" lines
 * appended after it (the reference takes them on the command line). device: CUDA ordinal. */
int hpb_sim_create(hpb_sim **out, const char *deck, const char *overrides, int device);
/* host-only dry run of the deck parser (same code path and the same HPB_ERR_PARSE messages as
 * hpb_sim_create; no GPU needed): summary receives "key=value;" pairs describing the run */
int hpb_deck_check(const char *deck, const char *overrides, char *summary, size_t n);
void hpb_sim_destroy(hpb_sim *sim);

/* Hipace::Evolve for time steps [step_begin, step_end] on this rank (all slices, or only the
 * first n_slices from the head if n_slices > 0); step_end < 0: up to the deck's max_step.
 * Blocks until the device is idle. */
int hpb_sim_evolve(hpb_sim *sim, int step_begin, int step_end, int n_slices);

/* finer control for tests: begin a time step (plasma re-init + neutralising background,
 * Hipace.cpp:403-475), then one Hipace::SolveOneSlice per call. */
int hpb_sim_begin_step(hpb_sim *sim, int step);
int hpb_sim_solve_one_slice(hpb_sim *sim, int islice);

/* queries (host buffers) */
int hpb_sim_geometry(hpb_sim *sim, int n_cell[3], double prob_lo[3], double prob_hi[3]);
/* guard cells of the slice array: (hipace.depos_order_xy + 1) / 2 + 1 (Fields.cpp:63-64) */
int hpb_sim_nguard(hpb_sim *sim);
int hpb_sim_ncomp(hpb_sim *sim);
int hpb_sim_comp_index(hpb_sim *sim, const char *which_slice, const char *name); /* -1 if absent */
/* copy one component of the current slice array (with guard cells, (ny+2g)*(nx+2g) doubles) */
int hpb_sim_get_field(hpb_sim *sim, int comp, double *h_out);
int hpb_sim_set_field(hpb_sim *sim, int comp, const double *h_in);
long hpb_sim_plasma_np(hpb_sim *sim, int species);
/* copy one PlasmaIdx real component / the validity flags of a species to the host */
int hpb_sim_get_plasma_real(hpb_sim *sim, int species, int idx, double *h_out);
int hpb_sim_get_plasma_valid(hpb_sim *sim, int species, uint8_t *h_out);
/* checksums accumulated over the slices of the last step: sum|Q| per This-slice component,
 * in component order; names via hpb_sim_checksum_name. */
int hpb_sim_checksum_count(hpb_sim *sim);
const char *hpb_sim_checksum_name(hpb_sim *sim, int k);
int hpb_sim_get_checksums(hpb_sim *sim, double *h_out);
/* beam checksums of the last step: x y z ux uy uz w (sum|.|), id sum, count */
int hpb_sim_get_beam_checksums(hpb_sim *sim, int beam, double h_out[9]);
/* Whole-beam host <-> device transfer: the role of MultiBuffer::get_data / put_data with host
 * staging buffers (src/utils/MultiBuffer.cpp:444-609) and of beam.injection_type = from_file.
 * h_real: 7 host arrays x y z w ux uy uz (np doubles each), h_idcpu: np, h_slot_off: nz+1
 * offsets with slot s = slice nz-1-s (head slice first).  Copies run on the simulation stream
 * (asynchronously if the host memory is pinned); get synchronises before returning. */
long hpb_sim_beam_np(hpb_sim *sim, int beam);
int hpb_sim_get_beam(hpb_sim *sim, int beam, double *const h_real[7], uint64_t *h_idcpu,
                     long *h_slot_off);
int hpb_sim_set_beam(hpb_sim *sim, int beam, const double *const h_real[7],
                     const uint64_t *h_idcpu, const long *h_slot_off);
/* performance counters of Hipace.cpp:509-553 and solver statistics */
typedef struct {
    double n_plasma_pushed, n_beam_pushed /* without slipped, BeamParticleAdvance.cpp:115-116 */,
           n_cells_updated;
    double slice_loop_ms;           /* device time of the slice loop(s) of the last evolve */
    long n_slices, n_mg_vcycles, n_qsa_violation, n_kernel_launches;
    double ms_deposit, ms_poisson, ms_explicit, ms_mg, ms_push, ms_other; /* if profiling on */
    long n_reorders;                /* plasma sorts (<plasma>.reorder_period) */
    long n_fused_slices;            /* slices that ran the fused driver order (push + next deposit in one kernel);
                                     * 0 means the deck switched it off: laser, rho diagnostic, grid current,
                                     * non-default deposition order, plasma in-situ diagnostics, n_subcycles < 1 */
} hpb_sim_stats;
int hpb_sim_get_stats(hpb_sim *sim, hpb_sim_stats *out);
/* m_physical_time and m_dt of the step that ran last (src/Hipace.H; Hipace.cpp:411-434).  With a pipeline
 * the time came from the upstream rank (MultiBuffer::get_time), with hipace.dt = adaptive dt is this
 * rank's own value. */
int hpb_sim_get_time(hpb_sim *sim, double *time, double *dt);
/* multigrid V-cycles (or predictor-corrector iterations) of every slice of the last evolve, in
 * slice-loop order (head first): what hpmg::MultiGrid::solve1 reports per call
 * (src/mg_solver/HpMultiGrid.cpp:1307-1427).  Copies min(n, count) entries; returns the count. */
long hpb_sim_get_mg_iters(hpb_sim *sim, int *h_out, long n);
/* device-side stopwatch on the simulation stream: start records an event, stop records a second
 * one, waits for it and returns the elapsed device time (bench.py's timed region) */
int hpb_sim_timer_start(hpb_sim *sim);
int hpb_sim_timer_stop(hpb_sim *sim, double *ms);
/* raw wire message (hpb_sim_pipeline_message_bytes / n_beams bytes) of one slice packet of the
 * current beam ring: what the downstream rank receives for slice nz-1-slot (tests) */
int hpb_sim_get_beam_packet(hpb_sim *sim, int beam, int slot, void *h_out);
/* options: "checksums" (0/1), "profile" (0/1), "max_step" (last time step of the run: the beam
 * is handed on only while step + 1 <= max_step, Hipace.cpp:441-443) */
int hpb_sim_set_option(hpb_sim *sim, const char *key, double value);

/* ---------------------------------------------------------------------------------------- */
/* (3) time-step pipeline over the GPUs of one box (the role of MultiBuffer, MPI replaced by  */
/*     NCCL point-to-point over NVLink; src/utils/MultiBuffer.cpp, Hipace.cpp:401)            */
/* ---------------------------------------------------------------------------------------- */
/* Rank r of `world` owns the time steps r, r + world, ...; for every slice it receives the beam
 * slice from rank r-1 (ring) and sends the pushed slice to rank r+1: one message per slice,
 * header {np per beam, step, time} + idcpu + x y z w ux uy uz of each beam, MultiBuffer's layout
 * (MultiBuffer.cpp:611-728) with a fixed per-slice capacity so that no count has to visit the
 * host.  Each directed edge r -> r+1 is its own 2-rank NCCL communicator with its own stream.
 * id_recv / id_send: ncclUniqueId bytes of the edges (r-1 -> r) and (r -> r+1), created with
 * hpb_nccl_unique_id on the edge's sending rank and exchanged by the caller (any transport). */
#define HPB_NCCL_ID_BYTES 128
int hpb_nccl_unique_id(char out[HPB_NCCL_ID_BYTES]);
int hpb_sim_pipeline_init(hpb_sim *sim, int rank, int world, const char *id_recv,
                          const char *id_send);
/* bytes of one slice message (all beams) and the slice capacity in particles of beam `beam` */
long hpb_sim_pipeline_message_bytes(hpb_sim *sim);
long hpb_sim_beam_slice_capacity(hpb_sim *sim, int beam);

#ifdef __cplusplus
}
#endif
#endif /* HPB200_H_ */
