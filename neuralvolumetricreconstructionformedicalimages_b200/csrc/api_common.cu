// api_common.cu -- error string, device info, level-table construction.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void nafb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int nafb_debug_flags() {
    static int cached = -1;
    if (cached < 0) {
        const char *e = getenv("NAFB_DEBUG_SKIP");
        cached = e ? atoi(e) : 0;
    }
    return cached;
}

// Per-DEVICE caches (one process may drive several GPUs: the SM count sizes persistent grids, and function attributes such as
// the dynamic shared-memory limit are per device).
int nafb_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
    return dev < NAFB_MAX_DEVICES ? dev : NAFB_MAX_DEVICES - 1;
}

int nafb_sm_count() {
    static int cached[NAFB_MAX_DEVICES] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev < 0 || dev >= NAFB_MAX_DEVICES) {
        int n = 0;
        return cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0 ? n : 148;
    }
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// Restates the index-mode decision of get_grid_index (hashencoder.cu:55-74) per level:
// the loop multiplies `stride` in uint32, so the comparison `stride <= hashmap_size`
// sees the WRAPPED value (levels 12/13 of the 16x2, 2^19, base-16 config stay linear).
int nafb_make_grid_params(const nafb_grid *g, GridParams *out) {
    if (!g || !g->table || !g->h_offsets) NAFB_FAIL(NAFB_ERR_INVALID, "nafb_grid: null pointer");
    if (g->D != 2 && g->D != 3) NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "GridEncoding: C must be 1, 2, 4, or 8.");  // hashencoder.cu:324 (sic)
    if (g->C != 1 && g->C != 2 && g->C != 4 && g->C != 8)
        NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "GridEncoding: C must be 1, 2, 4, or 8.");                           // hashencoder.cu:310
    if (g->L < 1 || g->L > NAFB_MAX_LEVELS)
        NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "nafb_grid: L=%u outside 1..%d", g->L, NAFB_MAX_LEVELS);
    out->table = g->table;
    out->L = g->L; out->C = g->C; out->D = g->D; out->H = g->H;
    for (uint32_t l = 0; l < g->L; ++l) {
        LevelParams &lp = out->lv[l];
        const int64_t sz = (int64_t)g->h_offsets[l + 1] - (int64_t)g->h_offsets[l];
        if (sz <= 0 || g->h_offsets[l] < 0) NAFB_FAIL(NAFB_ERR_INVALID, "nafb_grid: offsets not increasing at level %u", l);
        lp.offset = (uint32_t)g->h_offsets[l];
        lp.size = (uint32_t)sz;
        lp.mask = ((lp.size & (lp.size - 1)) == 0) ? lp.size - 1 : 0;
        lp.scale = fmaf(exp2f((float)l), (float)g->H, -1.0f);
        const uint32_t res = (uint32_t)ceilf(lp.scale) + 1u;
        uint32_t stride = 1, strides[3] = {0, 0, 0};
        uint32_t d = 0;
        for (; d < g->D && stride <= lp.size; ++d) {
            strides[d] = stride;
            stride *= (res + 1u);  // uint32 wrap on purpose
        }
        lp.hashed = stride > lp.size ? 1u : 0u;
        lp.s1 = strides[1];
        lp.s2 = strides[2];
    }
    for (uint32_t l = g->L; l < NAFB_MAX_LEVELS; ++l) out->lv[l] = LevelParams{0, 1, 0, 0.f, 0, 0, 0, 0};
    return NAFB_OK;
}

extern "C" {

int nafb_abi_version(void) { return NAFB_ABI_VERSION; }

const char *nafb_last_error(void) { return g_err; }

int nafb_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) NAFB_FAIL(NAFB_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    int n = 0, ma = 0, mi = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev);
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = ma;
    if (cc_minor) *cc_minor = mi;
    return NAFB_OK;
}

}  // extern "C"
