// loss.cuh -- the masked, chunk-wise MSE of train.py:69-127 / src/loss/loss.py:26-46 as a block-level device function, shared by
// the stand-alone kernel (render.cu, nafb_mse_loss) and the tail of the fused forward kernel (density_tc.cu: the last CTA to
// retire computes the loss, so a training step needs no separate loss launch).
#pragma once
#include "common.cuh"

#ifdef __CUDACC__
// Executed by ONE whole block, deterministic: loss = sum_chunks mean_{valid in chunk} (target - pred)^2.
// One warp per chunk (all chunks of a group in flight at once); the chunk means are then added in chunk order by one
// thread, as the reference's python loop does.  `zero_pred`: pred is cleared after it has been consumed (the fused engine
// accumulates the next step's projections into the same buffer).  s_mean / s_cnt: MSE_GROUP floats of shared memory each.
constexpr int MSE_GROUP = 1024;   // chunks per pass
constexpr int MSE_REG = 8;        // rays per lane kept in registers (chunks of up to 256 rays take the single-pass path)

__device__ __forceinline__ void mse_loss_block(float *__restrict__ pred, const float *__restrict__ target, const uint8_t *__restrict__ mask,
                                               uint32_t n, uint32_t chunk, float gscale, float *__restrict__ loss_out,
                                               float *__restrict__ dpred, int zero_pred, float *s_mean, float *s_cnt) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const uint32_t n_chunks = (n + chunk - 1) / chunk;
    float total = 0.f, total_cnt = 0.f;
    if (n_chunks * 2 <= n_warps && chunk > 32u * MSE_REG) {
        // few large chunks (the one-chunk loss of a 65 536-ray batch): the whole block works on one chunk at a time; warp
        // partial sums are combined in warp order, so the result depends on the block size only
        for (uint32_t c = 0; c < n_chunks; ++c) {
            const uint32_t c0 = c * chunk, c1 = c0 + chunk < n ? c0 + chunk : n;
            float s = 0.f, cnt = 0.f;
            for (uint32_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
                if (!mask || mask[i]) {
                    const float d = __fsub_rn(target[i], __ldcg(pred + i));
                    s = __fmaf_rn(d, d, s);
                    cnt += 1.f;
                }
            }
            s = warp_sum(s);
            cnt = warp_sum(cnt);
            if (lane == 0) { s_mean[warp] = s; s_cnt[warp] = cnt; }
            __syncthreads();
            if (threadIdx.x == 0) {
                float ss = 0.f, cc = 0.f;
                for (unsigned w = 0; w < n_warps; ++w) { ss += s_mean[w]; cc += s_cnt[w]; }
                total += ss / cc;
                total_cnt += cc;
                s_mean[MSE_GROUP - 1] = 1.0f / cc;
            }
            __syncthreads();
            const float inv = s_mean[MSE_GROUP - 1];
            for (uint32_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
                if (dpred) {
                    const bool m = !mask || mask[i];
                    dpred[i] = m ? gscale * 2.0f * __fsub_rn(__ldcg(pred + i), target[i]) * inv : 0.f;
                }
                if (zero_pred) pred[i] = 0.f;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) { loss_out[0] = total; loss_out[1] = total_cnt; }
        return;
    }
    for (uint32_t g0 = 0; g0 < n_chunks; g0 += MSE_GROUP) {
        const uint32_t g1 = g0 + MSE_GROUP < n_chunks ? g0 + MSE_GROUP : n_chunks;
        for (uint32_t c = g0 + warp; c < g1; c += n_warps) {
            const uint32_t c0 = c * chunk, c1 = c0 + chunk < n ? c0 + chunk : n;
            float s = 0.f, cnt = 0.f;
            if (c1 - c0 <= 32u * MSE_REG) {
                // the usual case (train.py's 200-ray chunks): every lane keeps its rays in registers -- ONE round of loads, all in
                // flight together, instead of two dependent passes over L2 (this block runs alone at the tail of the forward kernel)
                float pv[MSE_REG], tv[MSE_REG];
                bool mv[MSE_REG];
#pragma unroll
                for (int k = 0; k < MSE_REG; ++k) {
                    const uint32_t i = c0 + lane + 32u * k;
                    mv[k] = i < c1 && (!mask || mask[i]);
                    pv[k] = i < c1 ? __ldcg(pred + i) : 0.f;
                    tv[k] = i < c1 ? target[i] : 0.f;
                }
#pragma unroll
                for (int k = 0; k < MSE_REG; ++k) {
                    if (mv[k]) {
                        const float d = __fsub_rn(tv[k], pv[k]);
                        s = __fmaf_rn(d, d, s);
                        cnt += 1.f;
                    }
                }
                s = warp_sum(s);
                cnt = warp_sum(cnt);
                const float inv = 1.0f / cnt;   // an empty chunk gives mean(empty) = NaN in torch; keep that behaviour
                if (lane == 0) { s_mean[c - g0] = s / cnt; s_cnt[c - g0] = cnt; }
#pragma unroll
                for (int k = 0; k < MSE_REG; ++k) {
                    const uint32_t i = c0 + lane + 32u * k;
                    if (i < c1) {
                        if (dpred) dpred[i] = mv[k] ? gscale * 2.0f * __fsub_rn(pv[k], tv[k]) * inv : 0.f;
                        if (zero_pred) pred[i] = 0.f;
                    }
                }
                continue;
            }
            for (uint32_t i = c0 + lane; i < c1; i += 32) {
                if (!mask || mask[i]) {
                    const float d = __fsub_rn(target[i], __ldcg(pred + i));
                    s = __fmaf_rn(d, d, s);
                    cnt += 1.f;
                }
            }
            s = warp_sum(s);
            cnt = warp_sum(cnt);
            const float inv = 1.0f / cnt;   // an empty chunk gives mean(empty) = NaN in torch; keep that behaviour
            if (lane == 0) { s_mean[c - g0] = s / cnt; s_cnt[c - g0] = cnt; }
            for (uint32_t i = c0 + lane; i < c1; i += 32) {
                if (dpred) {
                    const bool m = !mask || mask[i];
                    dpred[i] = m ? gscale * 2.0f * __fsub_rn(__ldcg(pred + i), target[i]) * inv : 0.f;
                }
                if (zero_pred) pred[i] = 0.f;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (uint32_t c = 0; c < g1 - g0; ++c) { total += s_mean[c]; total_cnt += s_cnt[c]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { loss_out[0] = total; loss_out[1] = total_cnt; }
}
#endif
