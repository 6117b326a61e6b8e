// metrics.cu -- the 3-D evaluation metrics of the reference's eval_step (train.py:253-258) computed where the volume lives:
//   get_psnr_3d  (src/utils/util.py:55-84): mean squared difference in float64 -> 20 log10(PIXEL_MAX / sqrt(mse)) on the host;
//   get_ssim_3d  (src/utils/util.py:87-139): what skimage.metrics.structural_similarity evaluates for a 3-D array without a
//                channel axis -- uniform 7x7x7 window, SAMPLE covariance (NP / (NP - 1)), K1 = 0.01, K2 = 0.03, mean of the local
//                SSIM over the windows that fit (the reference averages three axis permutations of this same number).
// Float64 arithmetic like the reference's numpy code.  The SSIM kernel evaluates every window directly (343 voxels x 5 moments):
// 1.8 M windows for a 128^3 volume, a few milliseconds, no intermediate volumes.
#include "common.cuh"

namespace {

__device__ __forceinline__ double block_sum(double v, double *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];   // fixed order
    return t;   // valid in thread 0
}

// partial[block] = sum over the block's elements of (a - b)^2 (float64)
__global__ void __launch_bounds__(256) k_sqdiff_f64(const float *__restrict__ a, const float *__restrict__ b, uint64_t n, double *__restrict__ partial) {
    __shared__ double s_red[8];
    double s = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double d = (double)a[i] - (double)b[i];
        s += d * d;
    }
    const double t = block_sum(s, s_red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// one thread per window position (i, j, k) of the (n1-w+1) x (n2-w+1) x (n3-w+1) interior; partial[block] = sum of the local SSIM
__global__ void __launch_bounds__(256) k_ssim3d_f64(const float *__restrict__ a, const float *__restrict__ b, uint32_t n1, uint32_t n2, uint32_t n3,
                                                    uint32_t w, double c1, double c2, double *__restrict__ partial) {
    __shared__ double s_red[8];
    const uint32_t m1 = n1 - w + 1, m2 = n2 - w + 1, m3 = n3 - w + 1;
    const uint64_t total = (uint64_t)m1 * m2 * m3;
    const double np = (double)w * w * w, cov_norm = np / (np - 1.0);
    double acc = 0.0;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t k = (uint32_t)(q % m3), j = (uint32_t)((q / m3) % m2), i = (uint32_t)(q / ((uint64_t)m3 * m2));
        double sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
        for (uint32_t di = 0; di < w; ++di)
            for (uint32_t dj = 0; dj < w; ++dj) {
                const uint64_t base = ((uint64_t)(i + di) * n2 + (j + dj)) * n3 + k;
                for (uint32_t dk = 0; dk < w; ++dk) {
                    const double x = (double)__ldg(a + base + dk), y = (double)__ldg(b + base + dk);
                    sa += x; sb += y; saa += x * x; sbb += y * y; sab += x * y;
                }
            }
        const double ux = sa / np, uy = sb / np;
        const double vx = cov_norm * (saa / np - ux * ux), vy = cov_norm * (sbb / np - uy * uy), vxy = cov_norm * (sab / np - ux * uy);
        acc += ((2.0 * ux * uy + c1) * (2.0 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
    }
    const double t = block_sum(acc, s_red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

}  // namespace

extern "C" {

int nafb_sqdiff_f64(const float *a, const float *b, uint64_t n, double *partial, uint32_t n_partial, nafb_stream_t stream) {
    if (!a || !b || !partial || n_partial == 0) NAFB_FAIL(NAFB_ERR_INVALID, "sqdiff_f64: bad argument");
    k_sqdiff_f64<<<n_partial, 256, 0, (cudaStream_t)stream>>>(a, b, n, partial);
    NAFB_CHECK_LAUNCH("sqdiff_f64");
    return NAFB_OK;
}

int nafb_ssim3d_f64(const float *a, const float *b, uint32_t n1, uint32_t n2, uint32_t n3, uint32_t win, double data_range, double *partial,
                    uint32_t n_partial, nafb_stream_t stream) {
    if (!a || !b || !partial || n_partial == 0) NAFB_FAIL(NAFB_ERR_INVALID, "ssim3d_f64: bad argument");
    if (win < 2 || (win & 1u) == 0) NAFB_FAIL(NAFB_ERR_INVALID, "ssim3d_f64: win_size must be odd");
    if (n1 < win || n2 < win || n3 < win) NAFB_FAIL(NAFB_ERR_INVALID, "win_size exceeds image extent");   // skimage's message
    const double c1 = (0.01 * data_range) * (0.01 * data_range), c2 = (0.03 * data_range) * (0.03 * data_range);
    k_ssim3d_f64<<<n_partial, 256, 0, (cudaStream_t)stream>>>(a, b, n1, n2, n3, win, c1, c2, partial);
    NAFB_CHECK_LAUNCH("ssim3d_f64");
    return NAFB_OK;
}

}  // extern "C"
