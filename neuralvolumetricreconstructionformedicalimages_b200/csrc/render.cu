// render.cu -- the unfused pieces of the renderer at the reference's Python-op granularity:
// sampling (src/render/render.py:88-105), raw2outputs (:178-212) and the masked chunk-wise
// MSE of train.py:69-127 / src/loss/loss.py:26-46.  The fused training path does not launch
// these (it samples and integrates inside the density kernels); they back the drop-in
// render()/raw2outputs()/calc_mse_loss() operators.
#include "common.cuh"
#include "sampler.cuh"
#include "loss.cuh"

namespace {

// one block per ray: z_vals [N,S], pts [N,S,3], tv_partial[r] = sum_i |pts[r,i+1]-pts[r,i]|_1
__global__ void __launch_bounds__(128) k_sample_points(const SamplerParams sp, float *__restrict__ z_vals, float *__restrict__ pts,
                                                       float *__restrict__ tv_partial) {
    const uint32_t r = blockIdx.x;
    const uint32_t S = sp.n_samples;
    const RayRegs R = load_ray(sp, r);
    const Jitter tr = jitter_for(sp, r);
    float tv = 0.f;
    for (uint32_t i = threadIdx.x; i < S; i += blockDim.x) {
        const float z = z_sample(R.near, R.far, i, S, sp.lin_step, sp.perturb != 0, tr);
        float x[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) x[d] = ray_point(R.o[d], R.d[d], z, sp.clamp);
        if (z_vals) z_vals[(size_t)r * S + i] = z;
        if (pts) {
            float *q = pts + ((size_t)r * S + i) * 3;
            q[0] = x[0]; q[1] = x[1]; q[2] = x[2];
        }
        if (tv_partial && i + 1 < S) {
            const float zn = z_sample(R.near, R.far, i + 1, S, sp.lin_step, sp.perturb != 0, tr);
#pragma unroll
            for (int d = 0; d < 3; ++d) tv += fabsf(__fsub_rn(ray_point(R.o[d], R.d[d], zn, sp.clamp), x[d]));
        }
    }
    if (tv_partial) {
        __shared__ float red[4];
        tv = warp_sum(tv);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tv;
        __syncthreads();
        if (threadIdx.x == 0) tv_partial[r] = (red[0] + red[1]) + (red[2] + red[3]);
    }
}

__device__ __forceinline__ float ray_norm(const float *__restrict__ rays, uint32_t r) {
    const float dx = __ldg(rays + 8 * (size_t)r + 3), dy = __ldg(rays + 8 * (size_t)r + 4), dz = __ldg(rays + 8 * (size_t)r + 5);
    return sqrtf(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
}

// one warp per ray: acc[r] = sum_i raw[r,i,0] * (z[i+1]-z[i]) * |d|    (render.py:192-201)
__global__ void __launch_bounds__(128) k_ray_integral_fwd(const float *__restrict__ raw, uint32_t out_dim, const float *__restrict__ z_vals,
                                                          const float *__restrict__ rays, float *__restrict__ acc, float *__restrict__ absdiff,
                                                          uint32_t N, uint32_t S) {
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= N) return;
    const unsigned lane = threadIdx.x & 31;
    const float norm = ray_norm(rays, r);
    const float *z = z_vals + (size_t)r * S;
    const float *rw = raw + (size_t)r * S * out_dim;
    float s = 0.f;
    for (uint32_t i = lane; i < S; i += 32) {
        const float dist = i + 1 < S ? __fsub_rn(z[i + 1], z[i]) : 1e-10f;
        s = __fmaf_rn(rw[(size_t)i * out_dim], __fmul_rn(dist, norm), s);
        if (absdiff) absdiff[(size_t)r * S + i] = i == 0 ? 1e-10f : fabsf(__fsub_rn(rw[(size_t)i * out_dim + out_dim - 1], rw[(size_t)(i - 1) * out_dim + out_dim - 1]));
    }
    s = warp_sum(s);
    if (lane == 0) acc[r] = s;
}

__global__ void __launch_bounds__(256) k_ray_integral_bwd(const float *__restrict__ dacc, uint32_t out_dim, const float *__restrict__ z_vals,
                                                          const float *__restrict__ rays, float *__restrict__ draw, uint32_t N, uint32_t S) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (uint64_t)N * S) return;
    const uint32_t r = (uint32_t)(p / S), i = (uint32_t)(p - (uint64_t)r * S);
    const float dist = i + 1 < S ? __fsub_rn(z_vals[p + 1], z_vals[p]) : 1e-10f;
    const float g = __fmul_rn(__ldg(dacc + r), __fmul_rn(dist, ray_norm(rays, r)));
    draw[p * out_dim] = g;
    for (uint32_t c = 1; c < out_dim; ++c) draw[p * out_dim + c] = 0.f;
}

// rays [N,8] of N detector pixels (the in-kernel generator of sampler.cuh written out: dataset initialisation, tests)
__global__ void __launch_bounds__(256) k_generate_rays(const SamplerParams sp, float *__restrict__ rays_out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= sp.n_rays) return;
    const RayRegs R = make_ray(sp, r);
    float4 *o = reinterpret_cast<float4 *>(rays_out) + 2 * (size_t)r;
    o[0] = make_float4(R.o[0], R.o[1], R.o[2], R.d[0]);
    o[1] = make_float4(R.d[1], R.d[2], R.near, R.far);
}

// Single block: see loss.cuh
__global__ void __launch_bounds__(1024) k_mse_loss(float *__restrict__ pred, const float *__restrict__ target,
                                                   const uint8_t *__restrict__ mask, uint32_t n, uint32_t chunk, float gscale,
                                                   float *__restrict__ loss_out, float *__restrict__ dpred, int zero_pred) {
    __shared__ float s_mean[MSE_GROUP];
    __shared__ float s_cnt[MSE_GROUP];
    mse_loss_block(pred, target, mask, n, chunk, gscale, loss_out, dpred, zero_pred, s_mean, s_cnt);
}

// ---- hierarchical ("fine") sampling, reference render.py:113-126 + sample_pdf :215-247, one WARP per ray:
//   bins = midpoints of z_vals (S - 1); pdf over the S - 2 inner weights (+ 1e-5), cdf = [0, cumsum(pdf)] (S - 1 values);
//   sample_k = bins[below] + (u_k - cdf[below]) / denom * (bins[above] - bins[below])  with searchsorted(cdf, u_k, right=True);
//   z_out = sort(z_vals ++ samples);  pts = clamp(o + d * z_out);  tv = sum_i |pts[i+1] - pts[i]|_1.
// Shared memory per ray: cdf [S-1], bins [S-1], merge buffer [pow2 >= S + n_fine] (padded with +inf, bitonic-sorted by the warp).
constexpr int FINE_WARPS = 2;
__global__ void __launch_bounds__(32 * FINE_WARPS) k_sample_fine(const float *__restrict__ rays, const float *__restrict__ z_vals,
                                                                 const float *__restrict__ weights, const float *__restrict__ u, uint32_t u_stride,
                                                                 uint32_t n_rays, uint32_t S, uint32_t n_fine, uint32_t P2, float clampv,
                                                                 float *__restrict__ z_out, float *__restrict__ pts_out, float *__restrict__ tv_partial) {
    extern __shared__ float fine_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t r = blockIdx.x * FINE_WARPS + warp;
    if (r >= n_rays) return;
    const uint32_t per_ray = 2u * (S - 1u) + P2;
    float *cdf = fine_smem + (size_t)warp * per_ray, *bins = cdf + (S - 1u), *buf = bins + (S - 1u);
    const float *z = z_vals + (size_t)r * S, *w = weights + (size_t)r * S;
    const uint32_t nb = S - 2u;   // pdf entries
    // coarse depths into the merge buffer, midpoints into bins
    for (uint32_t i = lane; i < S; i += 32) buf[i] = z[i];
    for (uint32_t i = lane; i < S - 1u; i += 32) bins[i] = __fmul_rn(0.5f, __fadd_rn(z[i + 1], z[i]));
    // total weight
    float part = 0.f;
    for (uint32_t i = lane; i < nb; i += 32) part += __fadd_rn(w[i + 1], 1e-5f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const float total = part;
    // inclusive scan of pdf: contiguous chunk per lane, then a warp scan of the chunk sums
    const uint32_t chunk = (nb + 31u) / 32u, c0 = lane * chunk, c1 = c0 + chunk < nb ? c0 + chunk : nb;
    float run = 0.f;
    for (uint32_t i = c0; i < c1; ++i) {
        run += __fdiv_rn(__fadd_rn(w[i + 1], 1e-5f), total);
        cdf[i + 1] = run;
    }
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += up;
    }
    const float before = incl - run;   // sum of the chunks of the lower lanes
    __syncwarp();
    for (uint32_t i = c0; i < c1; ++i) cdf[i + 1] += before;
    if (lane == 0) cdf[0] = 0.f;
    __syncwarp();
    // inverse-cdf samples
    const uint32_t ncdf = S - 1u;
    for (uint32_t k = lane; k < n_fine; k += 32) {
        const float uk = u[(size_t)r * u_stride + k];
        uint32_t lo = 0, hi = ncdf;               // first index with cdf[idx] > uk (searchsorted right=True)
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (cdf[mid] > uk) hi = mid; else lo = mid + 1;
        }
        const uint32_t below = lo > 0 ? lo - 1u : 0u, above = lo < ncdf - 1u ? lo : ncdf - 1u;
        const float cl = cdf[below], ch = cdf[above], bl = bins[below], bh = bins[above];
        float denom = __fsub_rn(ch, cl);
        if (denom < 1e-5f) denom = 1.0f;
        const float t = __fdiv_rn(__fsub_rn(uk, cl), denom);
        buf[S + k] = __fadd_rn(bl, __fmul_rn(t, __fsub_rn(bh, bl)));
    }
    const uint32_t M = S + n_fine;
    for (uint32_t i = M + lane; i < P2; i += 32) buf[i] = __int_as_float(0x7f800000);
    __syncwarp();
    // bitonic sort of the merge buffer
    for (uint32_t k = 2; k <= P2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t c = lane; c < (P2 >> 1); c += 32) {
                const uint32_t i = 2u * c - (c & (j - 1u)), l = i + j;
                const float a = buf[i], b = buf[l];
                if ((a > b) == ((i & k) == 0u)) { buf[i] = b; buf[l] = a; }
            }
            __syncwarp();
        }
    // outputs
    const float *ray = rays + (size_t)r * 8;
    const float o0 = ray[0], o1 = ray[1], o2 = ray[2], d0 = ray[3], d1 = ray[4], d2 = ray[5];
    float tv = 0.f;
    for (uint32_t i = lane; i < M; i += 32) {
        const float zi = buf[i];
        const float x = fminf(fmaxf(__fadd_rn(o0, __fmul_rn(d0, zi)), -clampv), clampv);
        const float y = fminf(fmaxf(__fadd_rn(o1, __fmul_rn(d1, zi)), -clampv), clampv);
        const float zz = fminf(fmaxf(__fadd_rn(o2, __fmul_rn(d2, zi)), -clampv), clampv);
        if (z_out) z_out[(size_t)r * M + i] = zi;
        if (pts_out) {
            float *q = pts_out + ((size_t)r * M + i) * 3;
            q[0] = x; q[1] = y; q[2] = zz;
        }
        if (tv_partial && i + 1 < M) {
            const float zn = buf[i + 1];
            const float xn = fminf(fmaxf(__fadd_rn(o0, __fmul_rn(d0, zn)), -clampv), clampv);
            const float yn = fminf(fmaxf(__fadd_rn(o1, __fmul_rn(d1, zn)), -clampv), clampv);
            const float wn = fminf(fmaxf(__fadd_rn(o2, __fmul_rn(d2, zn)), -clampv), clampv);
            tv += fabsf(xn - x) + fabsf(yn - y) + fabsf(wn - zz);
        }
    }
    if (tv_partial) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tv += __shfl_xor_sync(0xffffffffu, tv, o);
        if (lane == 0) tv_partial[r] = tv;
    }
}

}  // namespace

extern "C" {

int nafb_sample_points(const nafb_sampler *smp, float *z_vals, float *pts, float *tv_partial, nafb_stream_t stream) {
    SamplerParams sp;
    uint64_t P = 0;
    int rc = nafb_make_sampler_params(smp, NAFB_SRC_RAYS, &sp, &P);
    if (rc) return rc;
    if (P == 0) return NAFB_OK;
    k_sample_points<<<sp.n_rays, 128, 0, (cudaStream_t)stream>>>(sp, z_vals, pts, tv_partial);
    NAFB_CHECK_LAUNCH("sample_points");
    return NAFB_OK;
}

int nafb_sample_fine(const float *rays, const float *z_vals, const float *weights, const float *u, uint32_t u_stride, uint32_t n_rays,
                     uint32_t n_samples, uint32_t n_fine, float clamp, float *z_out, float *pts_out, float *tv_partial, nafb_stream_t stream) {
    if (!rays || !z_vals || !weights || !u) NAFB_FAIL(NAFB_ERR_INVALID, "sample_fine: null pointer");
    if (n_samples < 3 || n_fine < 1 || n_samples + n_fine > 1024)
        NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "sample_fine: needs n_samples >= 3, n_fine >= 1, n_samples + n_fine <= 1024 (got %u + %u)", n_samples, n_fine);
    if (u_stride != 0 && u_stride < n_fine) NAFB_FAIL(NAFB_ERR_INVALID, "sample_fine: u_stride %u < n_fine %u", u_stride, n_fine);
    if (n_rays == 0) return NAFB_OK;
    uint32_t P2 = 1;
    while (P2 < n_samples + n_fine) P2 <<= 1;
    const size_t smem = (size_t)FINE_WARPS * (2u * (n_samples - 1u) + P2) * sizeof(float);   // <= 2 * (2046 + 1024) * 4 = 24.6 KB
    k_sample_fine<<<(n_rays + FINE_WARPS - 1) / FINE_WARPS, 32 * FINE_WARPS, smem, (cudaStream_t)stream>>>(rays, z_vals, weights, u, u_stride, n_rays, n_samples,
                                                                                                     n_fine, P2, clamp, z_out, pts_out, tv_partial);
    NAFB_CHECK_LAUNCH("sample_fine");
    return NAFB_OK;
}

int nafb_generate_rays(const nafb_sampler *smp, float *rays_out, nafb_stream_t stream) {
    if (!smp || !smp->pixels || !smp->poses || !rays_out) NAFB_FAIL(NAFB_ERR_INVALID, "generate_rays: pixels, poses and rays_out are required");
    nafb_sampler tmp = *smp;
    tmp.rays = nullptr;
    if (tmp.n_samples == 0) tmp.n_samples = 1;
    tmp.perturb = 0;
    SamplerParams sp;
    uint64_t P = 0;
    int rc = nafb_make_sampler_params(&tmp, NAFB_SRC_RAYS, &sp, &P);
    if (rc) return rc;
    if (sp.n_rays == 0) return NAFB_OK;
    k_generate_rays<<<(sp.n_rays + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sp, rays_out);
    NAFB_CHECK_LAUNCH("generate_rays");
    return NAFB_OK;
}

int nafb_ray_integral_forward(const float *raw, uint32_t out_dim, const float *z_vals, const float *rays, float *acc, float *absdiff,
                              uint32_t n_rays, uint32_t n_samples, nafb_stream_t stream) {
    if (!raw || !z_vals || !rays || !acc) NAFB_FAIL(NAFB_ERR_INVALID, "ray_integral_forward: null pointer");
    if (out_dim < 1 || out_dim > 2) NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "Wrong raw shape");  // render.py:210
    if (n_rays == 0 || n_samples == 0) return NAFB_OK;
    k_ray_integral_fwd<<<(n_rays + 3) / 4, 128, 0, (cudaStream_t)stream>>>(raw, out_dim, z_vals, rays, acc, absdiff, n_rays, n_samples);
    NAFB_CHECK_LAUNCH("ray_integral_forward");
    return NAFB_OK;
}

int nafb_ray_integral_backward(const float *dacc, uint32_t out_dim, const float *z_vals, const float *rays, float *draw, uint32_t n_rays,
                               uint32_t n_samples, nafb_stream_t stream) {
    if (!dacc || !z_vals || !rays || !draw) NAFB_FAIL(NAFB_ERR_INVALID, "ray_integral_backward: null pointer");
    if (n_rays == 0 || n_samples == 0) return NAFB_OK;
    const uint64_t P = (uint64_t)n_rays * n_samples;
    k_ray_integral_bwd<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dacc, out_dim, z_vals, rays, draw, n_rays, n_samples);
    NAFB_CHECK_LAUNCH("ray_integral_backward");
    return NAFB_OK;
}

int nafb_mse_loss(float *pred, const float *target, const uint8_t *mask, uint32_t n, uint32_t chunk, float gscale, float *loss_out,
                  float *dpred, int zero_pred, nafb_stream_t stream) {
    if (!pred || !target || !loss_out) NAFB_FAIL(NAFB_ERR_INVALID, "mse_loss: null pointer");
    if (chunk == 0 || chunk > n) chunk = n ? n : 1;
    const uint32_t n_chunks = (n + chunk - 1) / chunk;
    const unsigned threads = n_chunks >= 32 ? 1024 : (n_chunks < 4 ? 128 : n_chunks * 32);
    k_mse_loss<<<1, threads, 0, (cudaStream_t)stream>>>(pred, target, mask, n, chunk, gscale, loss_out, dpred, zero_pred);
    NAFB_CHECK_LAUNCH("mse_loss");
    return NAFB_OK;
}

}  // extern "C"
