// diag_common.cu -- the two host helpers of common.cuh for libnafb200_diag.so (diagnostics live outside the product library).
#include <stdarg.h>
#include <stdio.h>

#include "../common.cuh"
#include "../../../include/nafb200_diag.h"

static thread_local char g_err[512] = "";

void nafb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int nafb_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}

extern "C" const char *nafb_diag_last_error(void) { return g_err; }
