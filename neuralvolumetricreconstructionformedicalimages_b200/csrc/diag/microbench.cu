// microbench.cu -- measured denominators for the L2-resident gather / scatter roofline
// (first-GPU-call checklist item 6): random vector loads and random
// no-return float reductions over a resident table, plus contended variants.
#include "../common.cuh"

namespace {

__device__ __forceinline__ uint32_t lcg(uint32_t &s) {
    s = s * 1664525u + 1013904223u;
    return s ^ (s >> 15);
}

// mode: 0 ld.f32  1 ld.v2  2 ld.v4  3 red.f32  4 red.v2  5 red.v4  6 red.v2 warp-uniform address
//       7 red.v2 with pairs of lanes on the same address   8 ld.v2, pairs of adjacent entries (x, x+1)
template <int MODE>
__global__ void __launch_bounds__(256) k_micro(float *__restrict__ buf, uint32_t n_entries, int iters, float *__restrict__ sink) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    const unsigned lane = threadIdx.x & 31;
    for (int i = 0; i < iters; i += 4) {
        uint32_t idx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t r = lcg(s);
            if (MODE == 6) r = __shfl_sync(0xffffffffu, r, 0);
            if (MODE == 7) r = __shfl_sync(0xffffffffu, r, lane & ~1u);
            idx[j] = r % n_entries;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (MODE == 0) acc += __ldg(buf + idx[j]);
            if (MODE == 1) { const float2 v = __ldg(reinterpret_cast<const float2 *>(buf) + (idx[j] >> 1)); acc += v.x + v.y; }
            if (MODE == 2) { const float4 v = __ldg(reinterpret_cast<const float4 *>(buf) + (idx[j] >> 2)); acc += v.x + v.w; }
            if (MODE == 3) red_add_f32(buf + idx[j], 1e-9f);
            if (MODE == 4 || MODE == 6 || MODE == 7) red_add_f32x2(buf + (idx[j] & ~1u), 1e-9f, 1e-9f);
            if (MODE == 5) red_add_f32x4(buf + (idx[j] & ~3u), 1e-9f, 1e-9f, 1e-9f, 1e-9f);
            if (MODE == 8) {
                const uint32_t e = (idx[j] >> 1) & ~1u;   // even entry: (e, e+1) share a 16-byte pair
                const float2 a = __ldg(reinterpret_cast<const float2 *>(buf) + e);
                const float2 b = __ldg(reinterpret_cast<const float2 *>(buf) + e + 1);
                acc += a.x + b.y;
            }
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

}  // namespace

// Launches `reps` back-to-back kernels of (sm_count*8 blocks x 256 threads x iters ops); the caller times them.
extern "C" int nafb_microbench(int mode, float *buf, uint32_t n_floats, int iters, float *sink, uint64_t *ops_out, nafb_stream_t stream) {
    if (!buf || !sink) NAFB_FAIL(NAFB_ERR_INVALID, "microbench: null pointer");
    const int blocks = nafb_sm_count() * 8;
    cudaStream_t s = (cudaStream_t)stream;
    switch (mode) {
        case 0: k_micro<0><<<blocks, 256, 0, s>>>(buf, n_floats, iters, sink); break;
        case 1: k_micro<1><<<blocks, 256, 0, s>>>(buf, n_floats, iters, sink); break;
        case 2: k_micro<2><<<blocks, 256, 0, s>>>(buf, n_floats, iters, sink); break;
        case 3: k_micro<3><<<blocks, 256, 0, s>>>(buf, n_floats, iters, sink); break;
        case 4: k_micro<4><<<blocks, 256, 0, s>>>(buf, n_floats, iters, sink); break;
        case 5: k_micro<5><<<blocks, 256, 0, s>>>(buf, n_floats, iters, sink); break;
        case 6: k_micro<6><<<blocks, 256, 0, s>>>(buf, n_floats, iters, sink); break;
        case 7: k_micro<7><<<blocks, 256, 0, s>>>(buf, n_floats, iters, sink); break;
        case 8: k_micro<8><<<blocks, 256, 0, s>>>(buf, n_floats, iters, sink); break;
        default: NAFB_FAIL(NAFB_ERR_INVALID, "microbench: unknown mode %d", mode);
    }
    NAFB_CHECK_LAUNCH("microbench");
    if (ops_out) *ops_out = (uint64_t)blocks * 256 * (uint64_t)iters;
    return NAFB_OK;
}
