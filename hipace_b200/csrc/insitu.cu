// In-situ beam diagnostics: per-slice weighted moments of a beam slice reduced on the
// device, and the host-side writer of the reference's NumPy-structured file format, so that
// tools/read_insitu_diagnostics.py of the reference reads our files unchanged.
//   BeamParticleContainer::InSituComputeDiags / InSituWriteToFile
//       src/particles/beam/BeamParticleContainer.cpp:476-557, 596-732
//   insitu_utils::DataNode / write_header / write_data      src/utils/InsituUtil.H:22-90
#include "common.cuh"
#include "insitu.cuh"
#include <stdio.h>
#include <string>
#include <vector>

namespace {

constexpr int kT = 256;

// one launch per beam slice: block tree reduction, then 23 atomics per block into the slice's record
__global__ void __launch_bounds__(kT)
k_beam_insitu(hpb_beam_slice b, double clight_inv, double radius_sq, double *__restrict__ rec, long stride)
{
    hpb_pdl_prologue();
    __shared__ double red[kT / 32][23];
    long np = b.np;
    if (b.d_np) { const long n = (long)b.d_np[0]; if (n < np) np = n; }      // getNumParticles(This)
    double acc[23];
#pragma unroll
    for (int k = 0; k < 23; ++k) acc[k] = 0.;
    for (long ip = (long)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += (long)gridDim.x * blockDim.x) {
        double t[23];
        if (insitu_beam_terms(hpb_is_valid(b.idcpu[ip]), b.x[ip], b.y[ip], b.z[ip], b.ux[ip], b.uy[ip],
                              b.uz[ip], b.w[ip], clight_inv, radius_sq, t)) {
#pragma unroll
            for (int k = 0; k < 23; ++k) acc[k] += t[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 23; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 23) {
        double v = 0.;
        for (int wq = 0; wq < kT / 32; ++wq) v += red[wq][threadIdx.x];
        if (v != 0.) atomicAdd(rec + threadIdx.x * stride, v);
    }
}

// the plasma at the start of a slice: grid-stride over the SoA, 15 atomics per block
struct PlasmaSoA5 { const double *x, *y, *w, *ux, *uy, *psi; const uint64_t *idcpu; long np; };
__global__ void __launch_bounds__(kT)
k_plasma_insitu(PlasmaSoA5 p, double clight_inv, double radius_sq, double *__restrict__ rec, long stride)
{
    hpb_pdl_prologue();
    __shared__ double red[kT / 32][15];
    double acc[15];
#pragma unroll
    for (int k = 0; k < 15; ++k) acc[k] = 0.;
    for (long ip = (long)blockIdx.x * blockDim.x + threadIdx.x; ip < p.np; ip += (long)gridDim.x * blockDim.x) {
        double t[15];
        if (insitu_plasma_terms(hpb_is_valid(p.idcpu[ip]), p.x[ip], p.y[ip], p.ux[ip], p.uy[ip], p.psi[ip],
                                p.w[ip], clight_inv, radius_sq, t)) {
#pragma unroll
            for (int k = 0; k < 15; ++k) acc[k] += t[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 15; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 15) {
        double v = 0.;
        for (int wq = 0; wq < kT / 32; ++wq) v += red[wq][threadIdx.x];
        if (v != 0.) atomicAdd(rec + threadIdx.x * stride, v);
    }
}

// the fields of a slice: 10 sums over the valid box
struct FieldInsituComps { int exmby, eypbx, ez, bx, by, bz, jzb; };
__global__ void __launch_bounds__(kT)
k_field_insitu(SliceView a, FieldInsituComps c, int nx, int ny, double clight, double *__restrict__ rec,
               long stride)
{
    hpb_pdl_prologue();
    __shared__ double red[kT / 32][10];
    double acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.;
    const long n = (long)nx * ny;
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long)gridDim.x * blockDim.x) {
        const int j = (int)(q / nx), i = (int)(q - (long)j * nx);
        const long o = a.idx(i, j);
        double t[10];
        insitu_field_terms(a.comp(c.exmby)[o], a.comp(c.eypbx)[o], a.comp(c.ez)[o], a.comp(c.bx)[o],
                           a.comp(c.by)[o], a.comp(c.bz)[o], a.comp(c.jzb)[o], clight, t);
#pragma unroll
        for (int k = 0; k < 10; ++k) acc[k] += t[k];
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        double v = 0.;
        for (int wq = 0; wq < kT / 32; ++wq) v += red[wq][threadIdx.x];
        if (v != 0.) atomicAdd(rec + threadIdx.x * stride, v);
    }
}

// GatherMinUzSlice of one beam slice after its push: acc[0] = min(acc[0], min uz / c), acc[1..3] += sums
__device__ __forceinline__ void atomic_min_double(double *addr, double v)
{
    unsigned long long *a = (unsigned long long *)addr, old = *a;
    while (__longlong_as_double((long long)old) > v) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__global__ void __launch_bounds__(kT)
k_beam_min_uz(hpb_beam_slice b, double clight_inv, double *__restrict__ acc)
{
    hpb_pdl_prologue();
    __shared__ double red[kT / 32][4];
    long np = b.np;
    if (b.d_np) { const long n = (long)b.d_np[0]; if (n < np) np = n; }      // getNumParticles(This)
    double mn = 1e300, s1 = 0., s2 = 0., s3 = 0.;
    for (long ip = (long)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += (long)gridDim.x * blockDim.x) {
        double t[4];
        if (adaptive_uz_terms(hpb_is_valid(b.idcpu[ip]), b.uz[ip], b.w[ip], clight_inv, t)) {
            mn = fmin(mn, t[0]); s1 += t[1]; s2 += t[2]; s3 += t[3];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5][0] = mn; red[threadIdx.x >> 5][1] = s1;
        red[threadIdx.x >> 5][2] = s2; red[threadIdx.x >> 5][3] = s3;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kT / 32; ++w) {
            red[0][0] = fmin(red[0][0], red[w][0]);
            red[0][1] += red[w][1]; red[0][2] += red[w][2]; red[0][3] += red[w][3];
        }
        atomic_min_double(acc, red[0][0]);
        if (red[0][1] != 0.) { atomicAdd(acc + 1, red[0][1]); atomicAdd(acc + 2, red[0][2]); atomicAdd(acc + 3, red[0][3]); }
    }
}

// ---- the file format -------------------------------------------------------------------------
struct Node {
    std::string name, format;        // format empty: a nested structured datatype
    const void *data = nullptr;
    size_t bytes = 0;
    std::vector<Node> sub;
};
Node leaf(const char *name, const char *fmt, const void *data, size_t bytes, size_t count = 1)
{
    Node n;
    n.name = name;
    n.format = (count > 1 ? "(" + std::to_string(count) + ",)" : std::string()) + fmt;
    n.data = data; n.bytes = bytes * count;
    return n;
}
Node f8(const char *name, const double *p, size_t count = 1) { return leaf(name, "<f8", p, 8, count); }
Node i4(const char *name, const int *p, size_t count = 1) { return leaf(name, "<i4", p, 4, count); }

void header(const std::vector<Node> &nodes, std::string &out, const std::string &indent)
{
    out += indent + "{\n" + indent + "    \"names\": [\n";
    for (size_t i = 0; i < nodes.size(); ++i)
        out += indent + "        \"" + nodes[i].name + "\"" + (i + 1 == nodes.size() ? "\n" : ",\n");
    out += indent + "    ],\n" + indent + "    \"formats\": [\n";
    for (size_t i = 0; i < nodes.size(); ++i) {
        if (!nodes[i].format.empty()) out += indent + "        \"" + nodes[i].format + "\"";
        else header(nodes[i].sub, out, "        ");
        out += (i + 1 == nodes.size() ? "\n" : ",\n");
    }
    out += indent + "    ]\n" + indent + "}";
}
void payload(const std::vector<Node> &nodes, std::string &out)
{
    for (const Node &n : nodes) {
        if (!n.format.empty()) out.append((const char *)n.data, n.bytes);
        else payload(n.sub, out);
    }
}

}  // namespace

extern "C" int hpb_beam_insitu_slice(hpb_ctx *ctx, hpb_beam_slice bm, double insitu_radius,
                                     double *d_record, long stride)
{
    if (!ctx || !d_record || stride < 1) return HPB_ERR_ARG;
    if (bm.np == 0) return HPB_OK;
    unsigned nb = (unsigned)((bm.np + kT - 1) / kT);
    if (nb > 64) nb = 64;
    hpb_launch(k_beam_insitu, nb, kT, 0, ctx->stream, bm, 1.0 / ctx->g.c, insitu_radius * insitu_radius,
               d_record, stride);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

extern "C" int hpb_plasma_insitu_slice(hpb_ctx *ctx, hpb_plasma pl, double insitu_radius, double *d_record,
                                       long stride)
{
    if (!ctx || !d_record || stride < 1) return HPB_ERR_ARG;
    if (pl.np == 0) return HPB_OK;
    unsigned nb = (unsigned)((pl.np + kT - 1) / kT);
    if (nb > 1184) nb = 1184;         // 8 blocks per SM
    const PlasmaSoA5 p = {pl.r[HPB_X], pl.r[HPB_Y], pl.r[HPB_W], pl.r[HPB_UX], pl.r[HPB_UY], pl.r[HPB_PSI],
                          pl.idcpu, pl.np};
    hpb_launch(k_plasma_insitu, nb, kT, 0, ctx->stream, p, 1.0 / ctx->g.c, insitu_radius * insitu_radius,
               d_record, stride);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

// Host only.  sums: [15][n_slices] RAW per-slice sums in the order of insitu_plasma_terms
// (PlasmaParticleContainer::InSituWriteToFile, PlasmaParticleContainer.cpp:530-618)
extern "C" int hpb_insitu_write_plasma(const char *path, double time, int step, int n_slices, double charge,
                                       double mass, double z_lo, double z_hi,
                                       double normalized_density_factor, int is_normalized_units,
                                       const double *sums)
{
    if (!path || !sums || n_slices < 1) return HPB_ERR_ARG;
    constexpr int NR = kPlasmaInsituNReal;
    const size_t ns = (size_t)n_slices;
    std::vector<double> r(NR * ns), tot(NR, 0.);
    std::vector<int> np(ns);
    int np_tot = 0;
    for (size_t s = ns; s-- > 0;) {          // running sums from the head, like the beam's
        const double sw = sums[s];
        const double sum_w_inv = sw <= 0. ? 0. : 1. / sw;
        for (int i = 0; i < NR; ++i) {
            const double v = sums[i * ns + s];
            r[i * ns + s] = v * ((i == 0 || i == NR - 1) ? 1. : sum_w_inv);      // :513-517
            tot[i] += v;
        }
        np[s] = (int)sums[NR * ns + s];
        np_tot += np[s];
    }
    const double sum_w0 = tot[0];
    for (int i = 1; i < NR - 1; ++i) tot[i] /= sum_w0;
    static const char *names[NR] = {"sum(w)", "[x]", "[x^2]", "[y]", "[y^2]", "[ux]", "[ux^2]", "[uy]", "[uy^2]",
                                    "[uz]", "[uz^2]", "[ga]", "[ga^2]", "[(ga-1)*(1-vz)]"};
    std::vector<Node> all = {f8("time", &time), i4("step", &step), i4("n_slices", &n_slices),
                             f8("charge", &charge), f8("mass", &mass), f8("z_lo", &z_lo), f8("z_hi", &z_hi),
                             f8("normalized_density_factor", &normalized_density_factor),
                             i4("is_normalized_units", &is_normalized_units)};
    for (int i = 1; i < NR; ++i) all.push_back(f8(names[i], &r[i * ns], ns));
    all.push_back(f8("sum(w)", &r[0], ns));
    all.push_back(i4("Np", np.data(), ns));
    Node avg; avg.name = "average";
    for (int i = 1; i < NR - 1; ++i) avg.sub.push_back(f8(names[i], &tot[i]));
    Node total; total.name = "total";
    total.sub.push_back(f8("sum(w)", &tot[0]));
    total.sub.push_back(f8(names[NR - 1], &tot[NR - 1]));
    total.sub.push_back(i4("Np", &np_tot));
    all.push_back(avg);
    all.push_back(total);
    FILE *f = fopen(path, "ab");
    if (!f) { hpb_set_error("in-situ diagnostics: cannot open %s (does the directory exist?)", path); return HPB_ERR_ARG; }
    std::string out;
    fseek(f, 0, SEEK_END);
    if (ftell(f) == 0) header(all, out, "");
    payload(all, out);
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    if (fclose(f) != 0 || !ok) { hpb_set_error("in-situ diagnostics: error while writing %s", path); return HPB_ERR_ARG; }
    return HPB_OK;
}

extern "C" int hpb_fields_insitu_slice(hpb_ctx *ctx, hpb_slice sl, const int *comps, double *d_record,
                                       long stride)
{
    if (!ctx || !comps || !d_record || stride < 1) return HPB_ERR_ARG;
    if (comps[HPB_C_JZ_BEAM] < 0) {
        hpb_set_error("Must use explicit solver for field insitu diagnostic");       // Fields.cpp:1311-1312
        return HPB_ERR_UNSUPPORTED;
    }
    const hpb_geom &g = ctx->g;
    const FieldInsituComps c = {comps[HPB_C_EXMBY], comps[HPB_C_EYPBX], comps[HPB_C_EZ], comps[HPB_C_BX],
                                comps[HPB_C_BY], comps[HPB_C_BZ], comps[HPB_C_JZ_BEAM]};
    unsigned nb = (unsigned)(((long)g.nx * g.ny + kT - 1) / kT);
    if (nb > 592) nb = 592;
    hpb_launch(k_field_insitu, nb, kT, 0, ctx->stream, make_view(sl), c, g.nx, g.ny, g.c, d_record, stride);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

// Host only.  sums: [10][n_slices] RAW per-slice sums; every value is multiplied by dx dy dz
// (Fields::InSituWriteToFile, src/fields/Fields.cpp:1349-1428)
extern "C" int hpb_insitu_write_fields(const char *path, double time, int step, int n_slices, double z_lo,
                                       double z_hi, int is_normalized_units, double dxdydz,
                                       const double *sums)
{
    if (!path || !sums || n_slices < 1) return HPB_ERR_ARG;
    const size_t ns = (size_t)n_slices;
    std::vector<double> r(10 * ns), tot(10, 0.);
    for (size_t s = ns; s-- > 0;)
        for (int i = 0; i < 10; ++i) {
            r[i * ns + s] = sums[i * ns + s] * dxdydz;                               // :1343-1346
            tot[i] += r[i * ns + s];
        }
    static const char *names[10] = {"[Ex^2]", "[Ey^2]", "[Ez^2]", "[Bx^2]", "[By^2]", "[Bz^2]", "[ExmBy^2]",
                                    "[EypBx^2]", "[jz_beam]", "[Ez*jz_beam]"};
    std::vector<Node> all = {f8("time", &time), i4("step", &step), i4("n_slices", &n_slices), f8("z_lo", &z_lo),
                             f8("z_hi", &z_hi), i4("is_normalized_units", &is_normalized_units)};
    for (int i = 0; i < 10; ++i) all.push_back(f8(names[i], &r[i * ns], ns));
    Node integ; integ.name = "integrated";
    for (int i = 0; i < 10; ++i) integ.sub.push_back(f8(names[i], &tot[i]));
    all.push_back(integ);
    FILE *f = fopen(path, "ab");
    if (!f) { hpb_set_error("in-situ diagnostics: cannot open %s (does the directory exist?)", path); return HPB_ERR_ARG; }
    std::string out;
    fseek(f, 0, SEEK_END);
    if (ftell(f) == 0) header(all, out, "");
    payload(all, out);
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    if (fclose(f) != 0 || !ok) { hpb_set_error("in-situ diagnostics: error while writing %s", path); return HPB_ERR_ARG; }
    return HPB_OK;
}

// Host only.  sums: [8][n_slices] RAW per-slice values of hpb_laser_insitu_slice (max |a|^2, five sums,
// the on-axis sum re / im); nx, ny decide the on-axis averaging factor
// (MultiLaser::InSituWriteToFile, src/laser/MultiLaser.cpp:1003-1075)
extern "C" int hpb_insitu_write_laser(const char *path, double time, int step, int n_slices, double z_lo,
                                      double z_hi, int is_normalized_units, double dxdydz, int nx, int ny,
                                      const double *sums)
{
    if (!path || !sums || n_slices < 1) return HPB_ERR_ARG;
    const size_t ns = (size_t)n_slices;
    const double mid_factor = ((nx - 1) / 2 == nx / 2 ? 1. : 0.5) * ((ny - 1) / 2 == ny / 2 ? 1. : 0.5);   // :946-947
    std::vector<double> r(6 * ns), tot(6, 0.), axis(2 * ns);
    for (size_t s = ns; s-- > 0;) {
        r[s] = sums[s];
        tot[0] = tot[0] > sums[s] ? tot[0] : sums[s];                                   // :984-986
        for (int i = 1; i < 6; ++i) { r[i * ns + s] = sums[i * ns + s] * dxdydz; tot[i] += r[i * ns + s]; }
        axis[2 * s] = sums[6 * ns + s] * mid_factor;
        axis[2 * s + 1] = sums[7 * ns + s] * mid_factor;
    }
    static const char *names[6] = {"max(|a|^2)", "[|a|^2]", "[|a|^2*x]", "[|a|^2*x*x]", "[|a|^2*y]", "[|a|^2*y*y]"};
    std::vector<Node> all = {f8("time", &time), i4("step", &step), i4("n_slices", &n_slices), f8("z_lo", &z_lo),
                             f8("z_hi", &z_hi), i4("is_normalized_units", &is_normalized_units)};
    for (int i = 0; i < 6; ++i) all.push_back(f8(names[i], &r[i * ns], ns));
    all.push_back(leaf("axis(a)", "<c16", axis.data(), 16, ns));
    Node integ; integ.name = "integrated";
    for (int i = 0; i < 6; ++i) integ.sub.push_back(f8(names[i], &tot[i]));
    all.push_back(integ);
    FILE *f = fopen(path, "ab");
    if (!f) { hpb_set_error("in-situ diagnostics: cannot open %s (does the directory exist?)", path); return HPB_ERR_ARG; }
    std::string out;
    fseek(f, 0, SEEK_END);
    if (ftell(f) == 0) header(all, out, "");
    payload(all, out);
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    if (fclose(f) != 0 || !ok) { hpb_set_error("in-situ diagnostics: error while writing %s", path); return HPB_ERR_ARG; }
    return HPB_OK;
}

// d_acc[4] = {min uz / c, sum w, sum w uz / c, sum w uz^2 / c^2}, accumulated over the slices of a step
extern "C" int hpb_beam_min_uz_slice(hpb_ctx *ctx, hpb_beam_slice bm, double *d_acc)
{
    if (!ctx || !d_acc) return HPB_ERR_ARG;
    if (bm.np == 0) return HPB_OK;
    unsigned nb = (unsigned)((bm.np + kT - 1) / kT);
    if (nb > 64) nb = 64;
    hpb_launch(k_beam_min_uz, nb, kT, 0, ctx->stream, bm, 1.0 / ctx->g.c, d_acc);
    hpb_count_launch(ctx);
    HPB_CUDA_CHECK(cudaGetLastError());
    return HPB_OK;
}

// Host only (no GPU needed).  sums: [23][n_slices] RAW per-slice sums in the order of
// insitu_beam_terms (sum w, sum w x, ..., Np).  Appends one record to `path`; writes the JSON
// datatype header first when the file is empty (InSituWriteToFile :596-721).
extern "C" int hpb_insitu_write_beam(const char *path, double time, int step, int n_slices, double charge,
                                     double mass, double z_lo, double z_hi,
                                     double normalized_density_factor, int is_normalized_units,
                                     const double *sums)
{
    if (!path || !sums || n_slices < 1) return HPB_ERR_ARG;
    const size_t ns = (size_t)n_slices;
    std::vector<double> r(kInsituNReal * ns), tot(kInsituNReal, 0.);
    std::vector<int> np(ns);
    int np_tot = 0;
    for (size_t s = ns; s-- > 0;) {          // slice by slice from the head, like the running sums (:548)
        const double sw = sums[s];
        const double sum_w_inv = sw <= 0. ? 0. : 1. / sw;                   // :542
        for (int i = 0; i < kInsituNReal; ++i) {
            const double v = sums[i * ns + s];
            r[i * ns + s] = v * (i == 0 ? 1. : sum_w_inv);                  // :544-549
            tot[i] += v;
        }
        np[s] = (int)sums[kInsituNReal * ns + s];
        np_tot += np[s];
    }
    const double sum_w0 = tot[0];
    for (int i = 1; i < kInsituNReal; ++i) tot[i] /= sum_w0;                // :655-675
    static const char *names[kInsituNReal] = {"sum(w)", "[x]", "[x^2]", "[y]", "[y^2]", "[z]", "[z^2]",
        "[ux]", "[ux^2]", "[uy]", "[uy^2]", "[uz]", "[uz^2]", "[x*ux]", "[y*uy]", "[z*uz]", "[x*uy]",
        "[y*ux]", "[ux/uz]", "[uy/uz]", "[ga]", "[ga^2]"};
    std::vector<Node> all = {f8("time", &time), i4("step", &step), i4("n_slices", &n_slices),
                             f8("charge", &charge), f8("mass", &mass), f8("z_lo", &z_lo), f8("z_hi", &z_hi),
                             f8("normalized_density_factor", &normalized_density_factor),
                             i4("is_normalized_units", &is_normalized_units)};
    for (int i = 1; i < kInsituNReal; ++i) all.push_back(f8(names[i], &r[i * ns], ns));
    all.push_back(f8("sum(w)", &r[0], ns));
    all.push_back(i4("Np", np.data(), ns));
    Node avg; avg.name = "average";
    for (int i = 1; i < kInsituNReal; ++i) avg.sub.push_back(f8(names[i], &tot[i]));
    Node total; total.name = "total";
    total.sub.push_back(f8("sum(w)", &tot[0]));
    total.sub.push_back(i4("Np", &np_tot));
    all.push_back(avg);
    all.push_back(total);

    FILE *f = fopen(path, "ab");
    if (!f) { hpb_set_error("in-situ diagnostics: cannot open %s (does the directory exist?)", path); return HPB_ERR_ARG; }
    std::string out;
    fseek(f, 0, SEEK_END);
    if (ftell(f) == 0) header(all, out, "");
    payload(all, out);
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    if (fclose(f) != 0 || !ok) { hpb_set_error("in-situ diagnostics: error while writing %s", path); return HPB_ERR_ARG; }
    return HPB_OK;
}
