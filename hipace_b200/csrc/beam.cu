// Beam particles of one slice: push, re-binning of slipped particles, and the ring of
// fixed-capacity slice packets that carries a beam from one time step to the next (on this GPU
// or, through pipeline.cu, on the next one).  Restates
//   src/particles/pusher/BeamParticleAdvance.cpp:19-336   (AdvanceBeamParticlesSlice, level 0)
//   src/particles/pusher/ExternalFields.H:29-58           (ApplyExternalField)
//   src/particles/sorting/SliceSort.cpp:13-67             (shiftSlippedParticles)
//   src/utils/MultiBuffer.cpp:611-728, 730-905            (slice message layout, pack / unpack)
// B200 formulation: all particle counts stay on the device (packet header), so a slice is pushed,
// partitioned into {stay, slipped, dropped} with a stable two-level prefix (block counts come out
// of the push kernel itself) and packed into the wire layout by two launches and no host sync.
#include "sim.hpp"
#include "generic_order.cuh"
#include <string.h>
#include <memory>

struct hpb_extfields { DevRpn *d_prog; };      // 6 programs: Ex Ey Ez Bx By Bz

namespace {

constexpr int kBT = 256;      // threads per block == particles per class-count block

__device__ __forceinline__ long np_of(const hpb_beam_slice &b, int which)
{
    if (!b.d_np) return b.np;
    const long n = (long)b.d_np[which];
    return n < b.np ? n : b.np;
}

// ---- push -------------------------------------------------------------------------------------
// ORDER = -2: the specialised order-2 gather of common.cuh; 0..3: the generic one (generic_order.cuh)
template <bool EXT, int ORDER>
__global__ void __launch_bounds__(kBT)
k_advance_beam(hpb_beam_slice b, int *__restrict__ nsub, SliceView a, int c_psi, int c_ez, int c_bx,
               int c_by, int c_bz, double x_off, double y_off, double dx_inv, double dy_inv,
               double clight, double charge_mass_ratio, int n_subcycles, double dt, double time,
               double min_z, int do_z_push, int bc, double lox, double loy, double hix, double hiy,
               const DevRpn *__restrict__ ext, int *__restrict__ class_counts,
               double *__restrict__ checksum, unsigned long long *__restrict__ n_pushed)
{
    hpb_pdl_prologue();
    const long ip = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long np = np_of(b, 0), np_tot = np_of(b, 1);
    int cls = 0;                 // 0 dropped / beyond the slice, 1 stays, 2 slipped
    double cs[7] = {0., 0., 0., 0., 0., 0., 0.};
    unsigned long long cs_id = 0;
    bool counted = false;
    if (ip < np_tot) {
        const uint64_t idcpu = b.idcpu[ip];
        double xp = b.x[ip], yp = b.y[ip], zp = b.z[ip];
        double ux = b.ux[ip], uy = b.uy[ip], uz = b.uz[ip];
        if (checksum && ip < np) {       // beam diagnostic: state before the push
            cs[0] = fabs(xp); cs[1] = fabs(yp); cs[2] = fabs(zp);
            cs[3] = fabs(ux); cs[4] = fabs(uy); cs[5] = fabs(uz); cs[6] = fabs(b.w[ip]);
            cs_id = (idcpu & ~HPB_ID_VALID_BIT) >> 24;
            counted = true;
        }
        if (hpb_is_valid(idcpu)) {
            const double inv_c2 = 1.0 / (clight * clight);
            bool dead = false;
            int i = ip < np ? 0 : nsub[ip];     // counters restart at 0 on arrival
            for (; i < n_subcycles; ++i) {
                if (zp < min_z) break;        // not on this slice any more (:147-152)
                const double gammap_inv = 1.0 / sqrt(1.0 + (ux * ux + uy * uy + uz * uz) * inv_c2);
                xp += dt * 0.5 * ux * gammap_inv;
                yp += dt * 0.5 * uy * gammap_inv;
                if (enforce_particle_bc(xp, yp, ux, uy, bc, lox, loy, hix, hiy)) { dead = true; break; }
                GatheredFields f;
                if (ORDER < 0) {
                    f = gather_order2(a, c_psi, c_ez, c_bx, c_by, c_bz, x_off, y_off, dx_inv, dy_inv, xp, yp);
                } else {
                    const GenGrid gr = {x_off, y_off, dx_inv, dy_inv};
                    f = gen_gather<(ORDER < 0 ? 2 : ORDER)>(a, c_psi, c_ez, c_bx, c_by, c_bz, gr, xp, yp);
                }
                if (EXT) {
                    const double Ex = rpn_eval(ext[0], xp, yp, zp, time), Ey = rpn_eval(ext[1], xp, yp, zp, time);
                    const double Ez = rpn_eval(ext[2], xp, yp, zp, time), Bx = rpn_eval(ext[3], xp, yp, zp, time);
                    const double By = rpn_eval(ext[4], xp, yp, zp, time), Bz = rpn_eval(ext[5], xp, yp, zp, time);
                    f.ExmBy += Ex - clight * By;
                    f.EypBx += Ey + clight * Bx;
                    f.Ez += Ez; f.Bx += Bx; f.By += By; f.Bz += Bz;
                }
                const double ux_next = ux + dt * charge_mass_ratio
                    * (f.ExmBy + (clight - uz * gammap_inv) * f.By + uy * gammap_inv * f.Bz);
                const double uy_next = uy + dt * charge_mass_ratio
                    * (f.EypBx + (uz * gammap_inv - clight) * f.Bx - ux * gammap_inv * f.Bz);
                const double ux_i = (ux_next + ux) * 0.5, uy_i = (uy_next + uy) * 0.5;
                const double uz_i = uz + dt * 0.5 * charge_mass_ratio * f.Ez;
                const double gi_inv = 1.0 / sqrt(1.0 + (ux_i * ux_i + uy_i * uy_i + uz_i * uz_i) * inv_c2);
                const double uz_next = uz + dt * charge_mass_ratio
                    * (f.Ez + (ux_i * f.By - uy_i * f.Bx) * gi_inv);
                const double gn_inv = 1.0 / sqrt(1.0 + (ux_next * ux_next + uy_next * uy_next
                                                        + uz_next * uz_next) * inv_c2);
                xp += dt * 0.5 * ux_next * gn_inv;
                yp += dt * 0.5 * uy_next * gn_inv;
                if (do_z_push) zp += dt * (uz_next * gn_inv - clight);
                ux = ux_next; uy = uy_next; uz = uz_next;
            }
            if (!dead && enforce_particle_bc(xp, yp, ux, uy, bc, lox, loy, hix, hiy)) dead = true;
            if (dead) {                       // EnforceBC: w = 0, id invalid, nothing else stored
                b.w[ip] = 0.0;
                b.idcpu[ip] = hpb_make_invalid(idcpu);
            } else {
                b.x[ip] = xp; b.y[ip] = yp; b.z[ip] = zp;
                b.ux[ip] = ux; b.uy[ip] = uy; b.uz[ip] = uz;
                nsub[ip] = i;
                cls = (zp >= min_z) ? 1 : 2;
            }
        }
    }
    if (class_counts) {
        const int n1 = __syncthreads_count(cls == 1), n2 = __syncthreads_count(cls == 2);
        if (threadIdx.x == 0) { class_counts[2 * blockIdx.x] = n1; class_counts[2 * blockIdx.x + 1] = n2; }
    }
    if (n_pushed && ip == 0) atomicAdd(n_pushed, (unsigned long long)np);     // :115-116
    if (checksum) {
        // warp tree, then one atomic per warp and quantity
        const unsigned any = __ballot_sync(0xffffffffu, counted);
        if (any) {
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                double v = cs[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if ((threadIdx.x & 31) == 0) atomicAdd(&checksum[k], v);
            }
            unsigned long long idsum = cs_id;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) idsum += __shfl_xor_sync(0xffffffffu, idsum, o);
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(&checksum[7], (double)idsum);
                atomicAdd(&checksum[8], (double)__popc(any));
            }
        }
    }
}

// ---- shiftSlippedParticles + pack ---------------------------------------------------------------
__global__ void __launch_bounds__(kBT)
k_beam_partition(hpb_beam_slice b, const int *__restrict__ nsub, double min_z,
                 const int *__restrict__ class_counts, hpb_beam_slice stay, int64_t *stay_np,
                 hpb_beam_slice next, int64_t *next_np,
                 int *__restrict__ next_nsub, int *overflow)
{
    hpb_pdl_prologue();
    __shared__ long s_base[2];
    __shared__ int s_warp[2][kBT / 32];
    const long np_tot = np_of(b, 1);
    const int nblk_used = (int)((np_tot + kBT - 1) / kBT);
    if ((int)blockIdx.x >= nblk_used && blockIdx.x != 0) return;
    // offsets of this block = sum of the class counts of the preceding blocks; block 0 also
    // needs the grand totals for the headers
    const int upto = blockIdx.x == 0 ? nblk_used : (int)blockIdx.x;
    long a1 = 0, a2 = 0;
    for (int k = threadIdx.x; k < upto; k += kBT) { a1 += class_counts[2 * k]; a2 += class_counts[2 * k + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ long s_part[2][kBT / 32];
    if (lane == 0) { s_part[0][warp] = a1; s_part[1][warp] = a2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long t1 = 0, t2 = 0;
        for (int w = 0; w < kBT / 32; ++w) { t1 += s_part[0][w]; t2 += s_part[1][w]; }
        s_base[0] = t1; s_base[1] = t2;
    }
    __syncthreads();
    const long next_np0 = next_np ? (long)next_np[0] : 0;
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            long n_stay = s_base[0], n_slip = s_base[1];
            if (n_stay > stay.np) { n_stay = stay.np; *overflow = 1; }
            stay_np[0] = n_stay; stay_np[1] = n_stay;
            if (next.idcpu) {
                if (next_np0 + n_slip > next.np) { n_slip = next.np - next_np0; *overflow = 1; }
                next_np[1] = next_np0 + n_slip;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_base[0] = 0; s_base[1] = 0; }      // block 0 starts at offset 0
        __syncthreads();
    }
    const long ip = (long)blockIdx.x * kBT + threadIdx.x;
    int cls = 0;
    uint64_t idcpu = 0;
    double z = 0.;
    if (ip < np_tot) {
        idcpu = b.idcpu[ip];
        z = b.z[ip];
        if (hpb_is_valid(idcpu)) cls = (z >= min_z) ? 1 : 2;
    }
    // stable rank inside the block
    const unsigned m1 = __ballot_sync(0xffffffffu, cls == 1), m2 = __ballot_sync(0xffffffffu, cls == 2);
    if (lane == 0) { s_warp[0][warp] = __popc(m1); s_warp[1][warp] = __popc(m2); }
    __syncthreads();
    if (cls == 0) return;
    const int c = cls - 1;
    long pos = s_base[c];
    for (int w = 0; w < warp; ++w) pos += s_warp[c][w];
    pos += __popc((c == 0 ? m1 : m2) & ((1u << lane) - 1u));
    if (cls == 1) {
        if (pos >= stay.np) return;
        stay.idcpu[pos] = idcpu;
        stay.x[pos] = b.x[ip]; stay.y[pos] = b.y[ip]; stay.z[pos] = z; stay.w[pos] = b.w[ip];
        stay.ux[pos] = b.ux[ip]; stay.uy[pos] = b.uy[ip]; stay.uz[pos] = b.uz[ip];
    } else if (next.idcpu) {
        pos += next_np0;
        if (pos >= next.np) return;
        next.idcpu[pos] = idcpu;
        next.x[pos] = b.x[ip]; next.y[pos] = b.y[ip]; next.z[pos] = z; next.w[pos] = b.w[ip];
        next.ux[pos] = b.ux[ip]; next.uy[pos] = b.uy[ip]; next.uz[pos] = b.uz[ip];
        if (next_nsub) next_nsub[pos] = nsub[ip];
    }
}

// ---- ring maintenance ---------------------------------------------------------------------------
__global__ void k_ring_clear(BeamRing r)
{
    hpb_pdl_prologue();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= r.nslots) return;
    int64_t *h = r.hdr(s);
    for (int k = 0; k < 8; ++k) h[k] = 0;
}

__global__ void __launch_bounds__(kBT) k_ring_checksum(BeamRing r, double *out)
{
    hpb_pdl_prologue();
    const int s = blockIdx.y;
    const hpb_beam_slice b = r.view(s);
    const long n = np_of(b, 0);
    double cs[7] = {0., 0., 0., 0., 0., 0., 0.};
    unsigned long long ids = 0, cnt = 0;
    for (long ip = (long)blockIdx.x * blockDim.x + threadIdx.x; ip < n; ip += (long)gridDim.x * blockDim.x) {
        cs[0] += fabs(b.x[ip]); cs[1] += fabs(b.y[ip]); cs[2] += fabs(b.z[ip]);
        cs[3] += fabs(b.ux[ip]); cs[4] += fabs(b.uy[ip]); cs[5] += fabs(b.uz[ip]); cs[6] += fabs(b.w[ip]);
        ids += (b.idcpu[ip] & ~HPB_ID_VALID_BIT) >> 24;
        cnt += 1;
    }
    if (__syncthreads_count(cnt != 0) == 0) return;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        double v = cs[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v != 0.) atomicAdd(&out[k], v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ids += __shfl_xor_sync(0xffffffffu, ids, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt) { atomicAdd(&out[7], (double)ids); atomicAdd(&out[8], (double)cnt); }
}

// staging: 8 contiguous arrays of ntot entries: x y z w ux uy uz idcpu(bits)
template <bool GATHER>
__global__ void __launch_bounds__(kBT) k_ring_copy(BeamRing r, const long *__restrict__ off, double *stage, long ntot)
{
    hpb_pdl_prologue();
    const int s = blockIdx.y;
    const hpb_beam_slice b = r.view(s);
    const long o = off[s], n = off[s + 1] - o;
    double *arr[7] = {b.x, b.y, b.z, b.w, b.ux, b.uy, b.uz};
    uint64_t *sid = (uint64_t *)(stage + 7 * ntot);
    for (long ip = (long)blockIdx.x * blockDim.x + threadIdx.x; ip < n; ip += (long)gridDim.x * blockDim.x) {
        if (GATHER) {
#pragma unroll
            for (int k = 0; k < 7; ++k) stage[k * ntot + o + ip] = arr[k][ip];
            sid[o + ip] = b.idcpu[ip];
        } else {
#pragma unroll
            for (int k = 0; k < 7; ++k) arr[k][ip] = stage[k * ntot + o + ip];
            b.idcpu[ip] = sid[o + ip];
        }
    }
    if (!GATHER && blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t *h = r.hdr(s);
        h[0] = n; h[1] = n;
        for (int k = 2; k < 8; ++k) h[k] = 0;
    }
}

inline unsigned nblk(long n) { return (unsigned)((n + kBT - 1) / kBT); }

}  // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI seams
// ---------------------------------------------------------------------------------------------
// with the constants of a deck (my_constants.* of the input file) visible to the expressions
int hpb_extfields_create_deck(hpb_extfields **out, const char *const expr[6], const hpb::Deck *deck);

extern "C" int hpb_extfields_create(hpb_extfields **out, const char *const expr[6])
{
    return hpb_extfields_create_deck(out, expr, nullptr);
}

int hpb_extfields_create_deck(hpb_extfields **out, const char *const expr[6], const hpb::Deck *deck)
{
    if (!out || !expr) return HPB_ERR_ARG;
    std::vector<DevRpn> prog(6);
    const hpb::Deck empty;
    const hpb::Deck &d = deck ? *deck : empty;
    try {
        for (int k = 0; k < 6; ++k) {
            std::vector<hpb::RpnInstr> code;
            d.compile(expr[k] ? expr[k] : "0.", {"x", "y", "z", "t"}, code);
            rpn_from_code(prog[k], code);
        }
    } catch (const std::exception &e) {
        hpb_set_error("external field: %s", e.what());
        return HPB_ERR_PARSE;
    }
    std::unique_ptr<hpb_extfields> x(new hpb_extfields());
    HPB_CUDA_CHECK(cudaMalloc(&x->d_prog, 6 * sizeof(DevRpn)));
    HPB_CUDA_CHECK(cudaMemcpy(x->d_prog, prog.data(), 6 * sizeof(DevRpn), cudaMemcpyHostToDevice));
    *out = x.release();
    return HPB_OK;
}

extern "C" void hpb_extfields_destroy(hpb_extfields *ext)
{
    if (!ext) return;
    cudaFree(ext->d_prog);
    delete ext;
}

int hpb_advance_beam_impl(hpb_ctx *ctx, hpb_beam_slice bm, int *d_nsub, hpb_slice sl, double charge,
                          double mass, int n_subcycles, double dt, double time, double min_z,
                          int do_z_push, int particle_bc, const double bc_lo[2], const double bc_hi[2],
                          const int *comps, const hpb_extfields *ext, int *d_class_counts,
                          double *d_checksum, unsigned long long *d_n_pushed)
{
    if (!ctx || !comps || !bc_lo || !bc_hi || !d_nsub || n_subcycles < 1) return HPB_ERR_ARG;
    if (bm.np == 0) return HPB_OK;
    const hpb_geom &g = ctx->g;
#define HPB_ADV_BEAM_(E, O)                                                                       \
    hpb_launch(k_advance_beam<E, O>, nblk(bm.np), kBT, 0, ctx->stream,                                    \
        bm, d_nsub, make_view(sl), comps[HPB_C_PSI], comps[HPB_C_EZ], comps[HPB_C_BX],            \
        comps[HPB_C_BY], comps[HPB_C_BZ], g.x_off, g.y_off, 1.0 / g.dx, 1.0 / g.dy, g.c,          \
        charge / mass, n_subcycles, dt / n_subcycles, time, min_z, do_z_push, particle_bc,        \
        bc_lo[0], bc_lo[1], bc_hi[0], bc_hi[1], ext ? ext->d_prog : nullptr, d_class_counts,      \
        d_checksum, d_n_pushed)
#define HPB_ADV_BEAM(O) do { if (ext) HPB_ADV_BEAM_(true, O); else HPB_ADV_BEAM_(false, O); } while (0)
    if (!hpb_use_generic_order(ctx)) HPB_ADV_BEAM(-2);
    else if (ctx->depos_order == 0) HPB_ADV_BEAM(0);
    else if (ctx->depos_order == 1) HPB_ADV_BEAM(1);
    else if (ctx->depos_order == 2) HPB_ADV_BEAM(2);
    else HPB_ADV_BEAM(3);
#undef HPB_ADV_BEAM
#undef HPB_ADV_BEAM_
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_advance_beam_particles(hpb_ctx *ctx, hpb_beam_slice bm, int *d_nsubcycles,
                                          hpb_slice sl, double charge, double mass, int n_subcycles,
                                          double dt, double time, double min_z, int do_z_push,
                                          int particle_bc, const double bc_lo[2], const double bc_hi[2],
                                          const int *comps, const hpb_extfields *ext,
                                          int *d_class_counts, double *d_checksum)
{
    return hpb_advance_beam_impl(ctx, bm, d_nsubcycles, sl, charge, mass, n_subcycles, dt, time, min_z,
                                 do_z_push, particle_bc, bc_lo, bc_hi, comps, ext, d_class_counts,
                                 d_checksum, nullptr);
}

extern "C" int hpb_beam_shift_slipped(hpb_ctx *ctx, hpb_beam_slice bm, const int *d_nsubcycles,
                                      double min_z, const int *d_class_counts, hpb_beam_slice stay,
                                      int64_t *d_stay_np, hpb_beam_slice next,
                                      int64_t *d_next_np, int *d_next_nsubcycles, int *d_overflow)
{
    if (!ctx || !d_class_counts || !d_stay_np || !d_overflow || !d_nsubcycles) return HPB_ERR_ARG;
    if (next.idcpu && !d_next_np) return HPB_ERR_ARG;
    const unsigned nb = bm.np > 0 ? nblk(bm.np) : 1;
    hpb_launch(k_beam_partition, nb, kBT, 0, ctx->stream, bm, d_nsubcycles, min_z, d_class_counts, stay, d_stay_np,
                                                  next, d_next_np, d_next_nsubcycles, d_overflow);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

// ---------------------------------------------------------------------------------------------
// rings (driver side)
// ---------------------------------------------------------------------------------------------
int hpb_beam_rings_alloc(hpb_sim *s, BeamSp &b, long cap)
{
    cap = (cap + 31) / 32 * 32;
    if (b.ring[0].base && b.ring[0].cap >= cap) return HPB_OK;
    hpb_beam_rings_free(b);
    for (int k = 0; k < 2; ++k) {
        BeamRing &r = b.ring[k];
        r.cap = cap; r.nslots = s->nz;
        r.stride = (64 + 64 * (size_t)cap + 4 * (size_t)cap + 255) / 256 * 256;
        SIM_CUDA(cudaMalloc(&r.base, r.stride * (size_t)r.nslots));
        int rc = hpb_beam_ring_clear(s, r);
        if (rc) return rc;
    }
    SIM_CUDA(cudaMalloc(&b.d_class, 2 * sizeof(int) * (size_t)(nblk(cap) + 1)));
    SIM_CUDA(cudaMemset(b.d_class, 0, 2 * sizeof(int) * (size_t)(nblk(cap) + 1)));
    if (!b.d_cs) SIM_CUDA(cudaMalloc(&b.d_cs, 9 * sizeof(double)));
    if (!b.d_stage_off) SIM_CUDA(cudaMalloc(&b.d_stage_off, sizeof(long) * (size_t)(s->nz + 1)));
    b.cur = 0;
    return HPB_OK;
}

void hpb_beam_rings_free(BeamSp &b)
{
    for (int k = 0; k < 2; ++k) { cudaFree(b.ring[k].base); b.ring[k] = BeamRing(); }
    cudaFree(b.d_class); b.d_class = nullptr;
}

int hpb_beam_ring_clear(hpb_sim *s, const BeamRing &r)
{
    hpb_launch(k_ring_clear, (r.nslots + 255) / 256, 256, 0, s->stream, r);
    SIM_CUDA(cudaGetLastError());
    return HPB_OK;
}

// slot_off[s] = exclusive prefix of the per-slot counts (without slipped); synchronises
int hpb_beam_ring_counts(hpb_sim *s, const BeamRing &r, std::vector<long> &slot_off)
{
    std::vector<int64_t> h((size_t)r.nslots);
    SIM_CUDA(cudaMemcpy2DAsync(h.data(), sizeof(int64_t), r.base, r.stride, sizeof(int64_t), r.nslots,
                               cudaMemcpyDeviceToHost, s->stream));
    SIM_CUDA(cudaStreamSynchronize(s->stream));
    slot_off.assign(r.nslots + 1, 0);
    for (int k = 0; k < r.nslots; ++k) slot_off[k + 1] = slot_off[k] + (h[k] < r.cap ? h[k] : r.cap);
    return HPB_OK;
}

int hpb_beam_ring_checksum(hpb_sim *s, const BeamRing &r, double *d_out9)
{
    SIM_CUDA(cudaMemsetAsync(d_out9, 0, 9 * sizeof(double), s->stream));
    dim3 grid(nblk(r.cap) < 64 ? nblk(r.cap) : 64, r.nslots);
    hpb_launch(k_ring_checksum, grid, kBT, 0, s->stream, r, d_out9);
    SIM_CUDA(cudaGetLastError());
    return HPB_OK;
}

static int ensure_stage(BeamSp &b, long np)
{
    if (np <= b.stage_cap) return HPB_OK;
    cudaFree(b.d_stage); b.d_stage = nullptr;
    SIM_CUDA(cudaMalloc(&b.d_stage, 8 * sizeof(double) * (size_t)np));
    b.stage_cap = np;
    return HPB_OK;
}

int hpb_beam_ring_gather(hpb_sim *s, BeamSp &b, const BeamRing &r, const std::vector<long> &slot_off,
                         double *const h_real[7], uint64_t *h_idcpu)
{
    const long np = slot_off[r.nslots];
    if (np == 0) return HPB_OK;
    int rc = ensure_stage(b, np);
    if (rc) return rc;
    SIM_CUDA(cudaMemcpyAsync(b.d_stage_off, slot_off.data(), sizeof(long) * (r.nslots + 1),
                             cudaMemcpyHostToDevice, s->stream));
    dim3 grid(nblk(r.cap) < 64 ? nblk(r.cap) : 64, r.nslots);
    hpb_launch(k_ring_copy<true>, grid, kBT, 0, s->stream, r, b.d_stage_off, b.d_stage, b.stage_cap);
    SIM_CUDA(cudaGetLastError());
    for (int k = 0; k < 7; ++k)
        SIM_CUDA(cudaMemcpyAsync(h_real[k], b.d_stage + (size_t)k * b.stage_cap, sizeof(double) * np,
                                 cudaMemcpyDeviceToHost, s->stream));
    if (h_idcpu)
        SIM_CUDA(cudaMemcpyAsync(h_idcpu, b.d_stage + 7 * (size_t)b.stage_cap, sizeof(uint64_t) * np,
                                 cudaMemcpyDeviceToHost, s->stream));
    return HPB_OK;
}

int hpb_beam_ring_scatter(hpb_sim *s, BeamSp &b, const BeamRing &r, const long *h_slot_off,
                          const double *const h_real[7], const uint64_t *h_idcpu)
{
    const long np = h_slot_off[r.nslots];
    int rc = ensure_stage(b, np > 0 ? np : 1);
    if (rc) return rc;
    // the offsets are consumed by the kernel after this call returns: stage them on the device
    // through a pageable copy (synchronous w.r.t. the host buffer)
    SIM_CUDA(cudaMemcpyAsync(b.d_stage_off, h_slot_off, sizeof(long) * (r.nslots + 1),
                             cudaMemcpyHostToDevice, s->stream));
    for (int k = 0; k < 7 && np > 0; ++k)
        SIM_CUDA(cudaMemcpyAsync(b.d_stage + (size_t)k * b.stage_cap, h_real[k], sizeof(double) * np,
                                 cudaMemcpyHostToDevice, s->stream));
    if (np > 0)
        SIM_CUDA(cudaMemcpyAsync(b.d_stage + 7 * (size_t)b.stage_cap, h_idcpu, sizeof(uint64_t) * np,
                                 cudaMemcpyHostToDevice, s->stream));
    dim3 grid(nblk(r.cap) < 64 ? nblk(r.cap) : 64, r.nslots);
    hpb_launch(k_ring_copy<false>, grid, kBT, 0, s->stream, r, b.d_stage_off, b.d_stage, b.stage_cap);
    SIM_CUDA(cudaGetLastError());
    return HPB_OK;
}
