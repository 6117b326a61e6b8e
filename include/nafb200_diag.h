/*
 * nafb200_diag.h -- diagnostics kept OUT of the product library: libnafb200_diag.so (csrc/diag/) is built next to
 * libnafb200.so and loaded only by tests/ and scripts/.
 */
#ifndef NAFB200_DIAG_H
#define NAFB200_DIAG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* Known-answer test of the tcgen05 plumbing (one 128-row tile, bf16x3 split precision):
 *   D1 [128,32] = A[:, :32] . W[:, :32]^T ; D2 [128,64] = A[:, :32] . W ; D3 [128,32] = A^T . X
 * with A [128,128], X [128,32], W [32,64] fp32 row-major. */
int nafb_selftest_umma(const float *A, const float *X, const float *W, float *D1, float *D2, float *D3, void *stream);
/* Random-access microbenchmarks over a device buffer of n_floats floats (measured denominators of the
 * L2 gather / scatter roofline): mode 0/1/2 = ld.f32/.v2/.v4, 3/4/5 = red.add .f32/.v2/.v4,
 * 6 = red.v2 warp-uniform address, 7 = red.v2 lane pairs on one address, 8 = two adjacent ld.v2.
 * Enqueues one kernel; *h_ops receives the number of operations it performs. */
int nafb_microbench(int mode, float *buf, uint32_t n_floats, int iters, float *sink, uint64_t *h_ops, void *stream);
const char *nafb_diag_last_error(void);
#ifdef __cplusplus
}
#endif
#endif
