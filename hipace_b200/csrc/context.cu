// hpb_ctx life cycle and error reporting.
#include "common.cuh"
#include "tma.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int hpb_poisson_init(hpb_ctx *ctx);
void hpb_poisson_free(hpb_ctx *ctx);
int hpb_mg_init(hpb_ctx *ctx);
void hpb_mg_free(hpb_ctx *ctx);
void hpb_reorder_free(hpb_ctx *ctx);
void hpb_ref_arm_free(hpb_ctx *ctx);
void hpb_periodic_free(hpb_ctx *ctx);

static thread_local char g_err[1024] = "";

void hpb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void hpb_count_launch(hpb_ctx *ctx, int n) { ctx->n_launch += n; }

// programmatic dependent launch of the kernel chain: process-wide (the launch helper has no context)
static int g_pdl = 1;
bool hpb_pdl_enabled() { return g_pdl != 0; }
// FFT plans are built when a context is created: the Bluestein threshold is process-wide as well
static int g_bluestein_min_prime = 64;
int hpb_bluestein_min_prime() { return g_bluestein_min_prime; }

bool hpb_use_generic_order(const hpb_ctx *ctx)
{
    return ctx->force_generic || ctx->depos_order != 2 || ctx->depos_dtype != 2;
}

// Behaviour switches of the library (A/B measurements, cross-checks).  A drop-in library takes
// them through this call, never from the host's environment.
extern "C" int hpb_set_option(hpb_ctx *ctx, const char *key, double value)
{
    if (!key) return HPB_ERR_ARG;
    const int v = (int)value;
    if (!strcmp(key, "pdl")) { g_pdl = v != 0; return HPB_OK; }
    if (!strcmp(key, "bluestein_min_prime")) { g_bluestein_min_prime = v < 5 ? 5 : v; return HPB_OK; }
    if (!ctx) return HPB_ERR_ARG;
    if (!strcmp(key, "generic")) ctx->force_generic = v != 0;
    else if (!strcmp(key, "order")) ctx->tune_order = v;
    else if (!strcmp(key, "expl_variant")) ctx->tune_expl_variant = v;
    else if (!strcmp(key, "push_variant")) ctx->tune_push_variant = v;
    else if (!strcmp(key, "fft_variant")) ctx->tune_fft_variant = v;
    else if (!strcmp(key, "mg_wide")) ctx->tune_mg_wide = v;
    else if (!strcmp(key, "mg_fuse")) ctx->tune_mg_fuse = v;
    else if (!strcmp(key, "mg_rotate")) ctx->tune_mg_rotate = v;
    else if (!strcmp(key, "mg_lean")) ctx->tune_mg_lean = v;
    else if (!strcmp(key, "mg_persist")) ctx->tune_mg_persist = v;
    else if (!strcmp(key, "poisson_impl")) ctx->tune_poisson_impl = v;
    else { hpb_set_error("unknown option %s", key); return HPB_ERR_ARG; }
    return HPB_OK;
}

bool hpb_encode_slice_tmap(const hpb_slice &sl, int box_w, int box_h, CUtensorMap *out)
{
    typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
    static EncodeTiled encode = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess
            && q == cudaDriverEntryPointSuccess)
            encode = (EncodeTiled)fp;
        else
            (void)cudaGetLastError();
    }
    if (!encode || !sl.p || !out) return false;
    // TMA needs a 16-byte aligned base and strides that are multiples of 16 bytes; the inner box
    // extent must be a multiple of 16 bytes as well
    if (((uintptr_t)sl.p & 15) || (sl.jstride & 1) || (sl.nstride & 1) || (box_w & 1) || box_w > 256
        || box_h > 256 || sl.jstride < sl.nx_tot)
        return false;
    const cuuint64_t dims[3] = {(cuuint64_t)sl.nx_tot, (cuuint64_t)sl.ny_tot, (cuuint64_t)sl.ncomp};
    const cuuint64_t strides[2] = {(cuuint64_t)sl.jstride * sizeof(double), (cuuint64_t)sl.nstride * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, sl.p, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

const void *hpb_slice_tmap(hpb_ctx *ctx, int which, const hpb_slice &sl, int box_w, int box_h)
{
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    hpb_slice &k = ctx->tmap_key[which];
    if (!(k.p == sl.p && k.nx_tot == sl.nx_tot && k.ny_tot == sl.ny_tot && k.jstride == sl.jstride
          && k.nstride == sl.nstride && k.ncomp == sl.ncomp)) {
        k = sl;
        ctx->tmap_ok[which] = hpb_encode_slice_tmap(sl, box_w, box_h, (CUtensorMap *)ctx->tmap[which]) ? 1 : 0;
    }
    return ctx->tmap_ok[which] ? ctx->tmap[which] : nullptr;
}

extern "C" int hpb_set_deposition_order(hpb_ctx *ctx, int order_xy, int derivative_type)
{
    // Hipace.cpp:49-53: orders 0..3, derivative types 0..2, "analytic derivative with order 0 would vanish"
    if (!ctx || order_xy < 0 || order_xy > 3 || derivative_type < 0 || derivative_type > 2
        || (order_xy == 0 && derivative_type == 0)) {
        hpb_set_error("hpb_set_deposition_order: order %d / derivative type %d not allowed", order_xy,
                      derivative_type);
        return HPB_ERR_ARG;
    }
    ctx->depos_order = order_xy;
    ctx->depos_dtype = derivative_type;
    return HPB_OK;
}

extern "C" const char *hpb_last_error(void) { return g_err; }
extern "C" const char *hpb_version(void) { return "hpb200 0.1 (sm_100a, fp64)"; }

extern "C" int hpb_create(hpb_ctx **out, const hpb_geom *geom, void *stream)
{
    if (!out || !geom || geom->nx < 2 || geom->ny < 2) {
        hpb_set_error("hpb_create: bad arguments");
        return HPB_ERR_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        hpb_set_error("hpb_create: no CUDA device (this library has no CPU fallback)");
        return HPB_ERR_CUDA;
    }
    hpb_ctx *ctx = new hpb_ctx();
    memset(ctx, 0, sizeof(*ctx));
    ctx->g = *geom;
    ctx->stream = (cudaStream_t)stream;
    ctx->depos_order = 2; ctx->depos_dtype = 2;
    // defaults = the measured best on the 1024^2 ppc 4 deck (profiles/README.md): the row-tile TMA
    // push, the round-1 warp-aggregated explicit deposition
    ctx->tune_order = 1; ctx->tune_expl_variant = 0; ctx->tune_push_variant = 6;
    ctx->tune_mg_lean = 1;      // lean interior-tile path of the tile smoother: 0.365 -> 0.347 ms, bit-identical (r02J)
    ctx->tune_mg_rotate = 1;    // the V-cycle's last smoother writes into the caller's planes, no final copy: 0.373 -> 0.359 ms
    ctx->tune_mg_wide = 1;      // 1024-thread tiles on the multigrid levels that do not fill the GPU: 0.393 -> 0.367 ms
    int rc = hpb_poisson_init(ctx);
    if (rc == HPB_OK) rc = hpb_mg_init(ctx);
    if (rc == HPB_OK && cudaMalloc(&ctx->d_scalar_i, 16 * sizeof(int)) != cudaSuccess) rc = HPB_ERR_CUDA;
    if (rc != HPB_OK) { delete ctx; return rc; }
    *out = ctx;
    return HPB_OK;
}

extern "C" void hpb_destroy(hpb_ctx *ctx)
{
    if (!ctx) return;
    hpb_poisson_free(ctx);
    hpb_mg_free(ctx);
    hpb_reorder_free(ctx);
    hpb_ref_arm_free(ctx);
    hpb_periodic_free(ctx);
    cudaFree(ctx->d_scalar_i);
    delete ctx;
}
