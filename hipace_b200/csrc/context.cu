// hpb_ctx life cycle and error reporting.
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int hpb_poisson_init(hpb_ctx *ctx);
void hpb_poisson_free(hpb_ctx *ctx);
int hpb_mg_init(hpb_ctx *ctx);
void hpb_mg_free(hpb_ctx *ctx);

static thread_local char g_err[1024] = "";

void hpb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void hpb_count_launch(hpb_ctx *ctx, int n) { ctx->n_launch += n; }

bool hpb_pdl_enabled()
{
    static int on = -1;
    if (on < 0) { const char *e = getenv("HPB_PDL"); on = e ? (atoi(e) != 0) : 1; }
    return on != 0;
}

bool hpb_use_generic_order(const hpb_ctx *ctx)
{
    static int force = -1;       // HPB_GENERIC=1: the generic kernels also for order 2 / centred (cross-check)
    if (force < 0) { const char *e = getenv("HPB_GENERIC"); force = e ? (atoi(e) != 0) : 0; }
    return force || ctx->force_generic || ctx->depos_order != 2 || ctx->depos_dtype != 2;
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
    cudaFree(ctx->d_scalar_i);
    delete ctx;
}
