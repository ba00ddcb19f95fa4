// State of the slice-loop driver (hpb_sim), shared by sim.cu (Hipace::Evolve / SolveOneSlice),
// beam.cu (beam rings) and pipeline.cu (NCCL time-step pipeline).
#pragma once
#include "common.cuh"
#include "rpn.cuh"
#include "adaptive_dt.hpp"
#include <float.h>
#include <math.h>
#include <string>
#include <vector>

// ---- beam slice packets -----------------------------------------------------------------------
// A beam lives in two rings of nz fixed-capacity packets (slot s = slice nz-1-s, head first):
// ring[cur] holds the beam of the time step being computed (locally initialised, left by the
// previous step, or received from the upstream rank), ring[cur^1] receives the pushed slices
// (next step's input / message to the downstream rank).  One packet:
//   int64 hdr[8]   {np without slipped, np with slipped, step, time bits, 0...}   64 B
//   uint64 idcpu[cap]; double x[cap] y z w ux uy uz                               64 B * cap
//   int nsub[cap]   (BeamIdx::nsubcycles: never communicated, BeamParticleContainer.H:35-37)
// The first 64 + 64*cap bytes are the wire message (MultiBuffer.cpp:611-728 layout: metadata,
// idcpu, then the real components, each padded to the slice capacity).
struct BeamRing {
    char *base = nullptr;
    long cap = 0;            // particles per packet
    size_t stride = 0;       // bytes per packet
    int nslots = 0;
    __host__ __device__ size_t msg_bytes() const { return 64 + 64 * (size_t)cap; }
    __host__ __device__ char *packet(int slot) const { return base + (size_t)slot * stride; }
    __host__ __device__ int64_t *hdr(int slot) const { return (int64_t *)packet(slot); }
    __host__ __device__ int *nsub(int slot) const { return (int *)(packet(slot) + msg_bytes()); }
    __host__ __device__ hpb_beam_slice view(int slot) const
    {
        hpb_beam_slice v;
        char *p = packet(slot) + 64;
        v.idcpu = (uint64_t *)p;
        double *r = (double *)(p + 8 * (size_t)cap);
        v.x = r; v.y = r + cap; v.z = r + 2 * cap; v.w = r + 3 * cap;
        v.ux = r + 4 * cap; v.uy = r + 5 * cap; v.uz = r + 6 * cap;
        v.np = cap;
        v.d_np = hdr(slot);
        return v;
    }
};

struct Species {
    double density0 = 0.;                        // density(0, 0, 0): reported by hpb_deck_check
    std::vector<hpb::RpnInstr> density_host;     // the density expression for host-side evaluation
    // in-situ diagnostics (PlasmaParticleContainer.H:164, 198, 215)
    int insitu_period = 0;
    std::string insitu_file_prefix = "diags/plasma_insitu";
    double insitu_radius = INFINITY;
    double *d_insitu = nullptr;                  // [15][nz] raw per-slice sums of the current step
    long lattice_n = 0; int lattice_ppc = 1;     // regular InitParticles lattice (nothing filtered): cells per pass, passes
    std::string name;
    double charge = 0, mass = 0;
    int ppc[2] = {1, 1};
    DevRpn density;
    bool neutralize = true;
    double max_qsa = 35.;
    int n_subcycles = 1;
    double radius = INFINITY, hollow = 0., min_density = 0.;
    double u_mean[3] = {0, 0, 0};
    hpb_plasma d = {};
    long capacity = 0;
    // <plasma>.reorder_period / reorder_idx_type (PlasmaParticleContainer.cpp:147-151): the sort target
    int reorder_period = 0;
    int reorder_idx[2] = {0, 0};
    hpb_plasma d2 = {};
    long capacity2 = 0;
};

struct BeamSp {
    std::string name;
    double charge = 0, mass = 0;
    int ppc[3] = {1, 1, 1};
    int profile = 0;
    double density = 0, zmin = 0, zmax = 0, radius = 0, min_density = 0;
    double pos_mean[3] = {0, 0, 0}, pos_std[3] = {0, 0, 0}, u_mean[3] = {0, 0, 0};
    int n_subcycles = 10;
    bool do_z_push = true;
    // in-situ diagnostics (BeamParticleContainer.cpp:61-63, 296-311)
    double u_mean_z = 0.;                         // as in the deck (adaptive dt, AdaptiveTimeStep.cpp:96-106)
    double ts[4] = {1e30, 0., 0., 0.};            // adaptive dt: min uz, sum w, sum w uz, sum w uz^2
    double *d_ts = nullptr;
    int insitu_period = 0;
    std::string insitu_file_prefix = "diags/insitu";
    double insitu_radius = INFINITY;
    double *d_insitu = nullptr;   // [23][nz] raw per-slice sums of the current step
    std::string ext_expr[6];
    bool use_ext = false;
    hpb_extfields *ext = nullptr;
    // device state
    BeamRing ring[2];
    int cur = 0;                  // ring[cur]: this step's beam
    int *d_class = nullptr;       // per-256-particle-block {stay, slipped} counts of the last push
    double *d_cs = nullptr;       // 9 checksum accumulators of the current step (pre-push state)
    bool initialised = false;     // ring[cur] holds a beam
    bool from_host = false;       // ... that came from hpb_sim_set_beam (not re-created at step 0)
    bool cs_valid = false;        // d_cs holds a complete step
    // staging for whole-beam host transfers
    double *d_stage = nullptr; long stage_cap = 0; long *d_stage_off = nullptr;
};

struct hpb_pipeline;              // pipeline.cu

struct hpb_sim {
    hpb::Deck deck;
    int device = 0;
    cudaStream_t stream = nullptr;
    hpb_geom g = {};
    int nz = 0;
    double prob_lo[3], prob_hi[3];
    double bc_lo[2], bc_hi[2];
    int particle_bc = HPB_BC_PERIODIC;
    int depos_order = 2, depos_dtype = 2;         // hipace.depos_order_xy / depos_derivative_type
    int ng = HPB_NGUARD;                          // guard cells of the slice, Fields.cpp:63-64
    // hipace.bxby_solver = predictor-corrector (Hipace.cpp:935-1031, Hipace.H:210-222) and
    // boundary.field = Open (Fields.cpp:685-738)
    bool explicit_solver = true, open_bc = false;
    // boundary.field = Periodic (Fields::EnforcePeriodic) / fields.poisson_solver = FFTPeriodic
    bool field_periodic = false, poisson_periodic = false;
    double predcorr_tol = 4e-2, predcorr_mix = 0.05;
    int predcorr_max_iter = 30;
    double *d_pc_rhs = nullptr;                   // 3 staging planes nx * ny
    double *d_pc_scal = nullptr;                  // 38 multipole moments + 2 norms
    long n_predcorr_iters = 0;
    // hipace.dt = adaptive (utils/AdaptiveTimeStep.cpp)
    bool diag_xz = false;                         // diagnostic.diag_type = xz
    int field_insitu_period = 0;                  // fields.insitu_period (Fields.cpp:41-42)
    std::string field_insitu_prefix = "diags/field_insitu";
    double *d_field_insitu = nullptr;             // [10][nz]
    bool adaptive_dt = false;
    hpb_adaptive_par adp = {20., INFINITY, 2., 4e-4, 2000, 1, 1., 1., 1, 1};
    double adaptive_density = 0., min_uz_mq = DBL_MAX, time = 0., next_time = 0.;
    double dt_step = 0.;                          // the dt the current / last step runs with (end_step may already hold the next)
    bool adaptive_initialised = false;            // this rank's dt / min_uz_mq hold the initial estimate (Hipace.cpp:275-281)
    bool use_grid_current = false;                // utils/GridCurrent.cpp
    double gc_peak = 0., gc_mean[3] = {0., 0., 0.}, gc_std[3] = {1., 1., 1.};
    int max_step = 0;
    double dt = 0.;
    double mg_tol_rel = 1e-4, mg_tol_abs = DBL_MIN;
    bool deposit_rho = false, do_beam_jx_jy = true, any_neutral = false;
    hpb_ctx *ctx = nullptr;
    hpb_slice sl = {};
    int comps[HPB_C_COUNT];             // enum hpb_comp -> physical plane (rotates in fused mode)
    int comps0[HPB_C_COUNT];            // the table as built (identity order of comp_names)
    std::vector<std::pair<std::string, std::string>> comp_names;   // (which_slice, name) by LOGICAL index
    std::vector<int> comp_id;           // logical index -> enum hpb_comp
    // fused slice transition (our addition, see hpb_fields_shift_and_initialize): the end of
    // slice k shifts / initialises the planes for slice k-1 and the plasma push of slice k
    // deposits jx jy chi rhomjz of slice k-1 from registers
    // lasers (time step 0 only: analytic envelope + ponderomotive terms)
    std::vector<hpb_laser> lasers;
    double laser_lambda0 = 0.;
    int laser_interp_order = 1;
    bool use_laser = false;
    // the envelope advance over time steps (laser.cu, second half); null for dt = 0 / max_step = 0 decks,
    // which evaluate the analytic envelope per cell instead of storing it
    hpb_laser_state *laser_state = nullptr;
    bool laser_use_phase = true;
    bool laser_use_mg = true, laser_mg_avg_rhs = true;      // lasers.solver_type = multigrid (the default)
    double laser_mg_tol_rel = 1e-4, laser_mg_tol_abs = 0.;
    int laser_insitu_period = 0;                  // lasers.insitu_period (MultiLaser.cpp)
    std::string laser_insitu_prefix = "diags/laser_insitu";
    double *d_laser_insitu = nullptr;             // [8][nz]
    bool opt_fuse = true;
    // beam-side work of the fused order (beam push / re-binning / hand-off of this slice, beam
    // deposits and the Sx, Sy seed of the next one) runs on a second stream beside the plasma push
    bool opt_side_stream = true;
    bool opt_side_late = true;       // enqueue the side stream's beam work behind the plasma push (host order)
    // the multigrid's coefficient hierarchy (hpb_mg_prepare_acf: five small dependent launches) depends on chi
    // only: option "mg_early" runs it on a third stream beside the Poisson solve and the explicit deposition
    // (bit-identical results; measured: no gain over the in-stream PDL chain, 880 vs 880 slices/s, so off)
    bool opt_mg_early = false;
    cudaStream_t stream3 = nullptr;
    cudaEvent_t ev_chi = nullptr, ev_acf = nullptr;
    cudaStream_t stream2 = nullptr;
    cudaStream_t beam_stream = nullptr;     // where the pipeline's per-slice waits / records go
    cudaEvent_t ev_fields = nullptr, ev_shift = nullptr, ev_side = nullptr;
    bool prepared = false;              // the current slice was initialised + deposited by its predecessor
    std::vector<Species> plasmas;
    std::vector<BeamSp> beams;
    // scratch for init
    unsigned *d_flag = nullptr, *d_offs = nullptr;
    long scan_cap = 0;
    void *d_cub = nullptr;
    long *d_slot_off = nullptr;         // nz + 1
    size_t cub_bytes = 0;
    // diagnostics
    double *d_checksum = nullptr;       // ncomp
    int *d_nqsa = nullptr;
    int *d_overflow = nullptr;          // beam packet overflow flag
    unsigned long long *d_count = nullptr;   // [0]: beam particles pushed (without slipped)
    bool opt_checksums = true, opt_profile = false;
    hpb_sim_stats stats = {};
    std::vector<int> mg_iters;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> pev;
    int cur_step = -1;
    hpb_pipeline *pipe = nullptr;
};

#define SIM_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            hpb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            return HPB_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

// beam.cu
int hpb_beam_rings_alloc(hpb_sim *s, BeamSp &b, long cap);
void hpb_beam_rings_free(BeamSp &b);
int hpb_beam_ring_clear(hpb_sim *s, const BeamRing &r);
int hpb_beam_ring_counts(hpb_sim *s, const BeamRing &r, std::vector<long> &slot_off);
int hpb_beam_ring_checksum(hpb_sim *s, const BeamRing &r, double *d_out9);
int hpb_beam_ring_gather(hpb_sim *s, BeamSp &b, const BeamRing &r, const std::vector<long> &slot_off,
                         double *const h_real[7], uint64_t *h_idcpu);
int hpb_beam_ring_scatter(hpb_sim *s, BeamSp &b, const BeamRing &r, const long *h_slot_off,
                          const double *const h_real[7], const uint64_t *h_idcpu);

// pipeline.cu
void hpb_pipeline_destroy(hpb_sim *s);
int hpb_pipeline_begin_step(hpb_sim *s, int step);          // post receives / order ring reuse
int hpb_pipeline_wait_slice(hpb_sim *s, int islice);        // compute stream waits for slice islice
int hpb_pipeline_wait_slice_on(hpb_sim *s, int islice, cudaStream_t st);       // ... a given stream
int hpb_pipeline_wait_out_slot_on(hpb_sim *s, int islice, cudaStream_t st);    // out slot free, on a given stream
int hpb_extfields_create_deck(hpb_extfields **out, const char *const expr[6], const hpb::Deck *deck);
struct hpb_laser_state;
void hpb_laser_packet(hpb_laser_state *st, int islice, void *recv[2], void *send[2], size_t *bytes);
int hpb_pipeline_send_slice(hpb_sim *s, int islice, int step);
int hpb_pipeline_wait_out_slot(hpb_sim *s, int islice);     // out-ring slot free (its last send left)
bool hpb_pipeline_out_ring_busy(const hpb_sim *s);          // sends of the previous owned step pending
int hpb_pipeline_end_step(hpb_sim *s, int step);
// MultiBuffer::get_time / put_time (utils/MultiBuffer.cpp:611-651): the physical time of a step travels
// from the rank that owns the step before it; get blocks the host until the value is there
int hpb_pipeline_get_time(hpb_sim *s, int step, double *t);
int hpb_pipeline_put_time(hpb_sim *s, int step, double t_next);
int hpb_pipeline_world(const hpb_sim *s);
bool hpb_pipeline_receives(const hpb_sim *s, int step);     // this step's beam comes from upstream
bool hpb_pipeline_active(const hpb_sim *s);
int hpb_pipeline_rank(const hpb_sim *s);
