// adam.cu -- fused dense Adam + gradient zeroing over one flat parameter vector
// (reference: torch.optim.Adam as configured in src/trainer.py:54 -- betas (0.9, 0.999),
// eps 1e-8, no weight decay, no amsgrad -- plus optimizer.zero_grad() of trainer.py:138).
//
// HBM-bound: per parameter 16 B read (p, g, m, v) + 12 B written (p, m, v) (+4 B for the
// zeroed gradient).  128-bit loads/stores, grid-stride, grid = SMs x 8.
//
// The arithmetic follows torch's multi-tensor (foreach) Adam op by op, each op with its own
// rounding, so that the fused step can be compared bit for bit with torch.optim.Adam:
//   m  = lerp(m, g, 1-b1)              -> fma(w, g - m, m)        (weight < 0.5 branch)
//   v  = v*b2 ; v = addcmul(v, g, g, 1-b2) -> fma(1-b2, g*g, v)
//   dn = sqrt(v) / sqrt(1-b2^t) + eps
//   p  = addcdiv(p, m, dn, -lr/(1-b1^t)) -> fma(-step_size, m/dn, p)
#include <stdlib.h>

#include "adam.cuh"

namespace {

__global__ void __launch_bounds__(256) k_adam(float *__restrict__ param, float *__restrict__ grad, float *__restrict__ m_, float *__restrict__ v_,
                                              uint64_t n, AdamConst c, int zero_grad) {
    const uint64_t n4 = n >> 2;
    float4 *p4 = reinterpret_cast<float4 *>(param);
    float4 *g4 = reinterpret_cast<float4 *>(grad);
    float4 *m4 = reinterpret_cast<float4 *>(m_);
    float4 *v4 = reinterpret_cast<float4 *>(v_);
    // L2 residency: the parameters are what the next forward pass gathers from (and the zeroed gradient what the next backward
    // pass reduces into) -> evict-last; exp_avg / exp_avg_sq are touched once per step -> evict-first, so that the 228 MB they
    // move do not flush the other 114 MB out of the 126 MB L2.
    const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        float4 p = ld_l2hint(p4 + i, keep), g = g4[i], m = ld_l2hint(m4 + i, stream), v = ld_l2hint(v4 + i, stream);
        adam_one(p.x, g.x, m.x, v.x, c);
        adam_one(p.y, g.y, m.y, v.y, c);
        adam_one(p.z, g.z, m.z, v.z, c);
        adam_one(p.w, g.w, m.w, v.w, c);
        st_l2hint(p4 + i, p, keep);
        st_l2hint(m4 + i, m, stream);
        st_l2hint(v4 + i, v, stream);
        if (zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // tail (n % 4)
    const uint64_t tail0 = n4 << 2;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n - tail0) {
        const uint64_t i = tail0 + gid;
        float p = param[i], g = grad[i], m = m_[i], v = v_[i];
        adam_one(p, g, m, v, c);
        param[i] = p; m_[i] = m; v_[i] = v;
        if (zero_grad) grad[i] = 0.f;
    }
}

__global__ void __launch_bounds__(256) k_adam_dev(float *__restrict__ param, float *__restrict__ grad, float *__restrict__ m_, float *__restrict__ v_,
                                                  uint64_t n, double beta1, double beta2, double eps, float gscale, int zero_grad, uint32_t *state) {
    __shared__ AdamConst sc;
    if (threadIdx.x == 0) sc = adam_const_from_state(state, beta1, beta2, eps, gscale);
    __syncthreads();
    const AdamConst c = sc;
    const uint64_t n4 = n >> 2;
    float4 *p4 = reinterpret_cast<float4 *>(param);
    float4 *g4 = reinterpret_cast<float4 *>(grad);
    float4 *m4 = reinterpret_cast<float4 *>(m_);
    float4 *v4 = reinterpret_cast<float4 *>(v_);
    const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();   // see k_adam
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        float4 p = ld_l2hint(p4 + i, keep), g = g4[i], m = ld_l2hint(m4 + i, stream), v = ld_l2hint(v4 + i, stream);
        adam_one(p.x, g.x, m.x, v.x, c);
        adam_one(p.y, g.y, m.y, v.y, c);
        adam_one(p.z, g.z, m.z, v.z, c);
        adam_one(p.w, g.w, m.w, v.w, c);
        st_l2hint(p4 + i, p, keep);
        st_l2hint(m4 + i, m, stream);
        st_l2hint(v4 + i, v, stream);
        if (zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // last block out increments the step (every block has read it by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(state + NAFB_STATE_TICKET, 1u) == gridDim.x - 1) {
            state[NAFB_STATE_TICKET] = 0u;
            state[NAFB_STATE_STEP] = state[NAFB_STATE_STEP] + 1u;
        }
    }
}

}  // namespace

extern "C" int nafb_adam_step_dev(float *param, float *grad, float *exp_avg, float *exp_avg_sq, uint64_t n, double beta1, double beta2, double eps,
                                  float grad_scale, int zero_grad, uint32_t *state, nafb_stream_t stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || !state) NAFB_FAIL(NAFB_ERR_INVALID, "adam_step_dev: null pointer");
    if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0)
        NAFB_FAIL(NAFB_ERR_INVALID, "adam_step_dev: buffers must be 16-byte aligned");
    if (n == 0 || (n & 3)) NAFB_FAIL(NAFB_ERR_INVALID, "adam_step_dev: n must be a positive multiple of 4");
    uint64_t blocks = ((n >> 2) + 255) / 256;
    const uint64_t cap = (uint64_t)nafb_sm_count() * 8;   // two waves of the 4 blocks per SM the registers allow (4 or 6 measured slower)
    if (blocks > cap) blocks = cap;
    k_adam_dev<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps, grad_scale, zero_grad, state);
    NAFB_CHECK_LAUNCH("adam_step_dev");
    return NAFB_OK;
}

extern "C" int nafb_adam_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq, uint64_t n, double lr, double beta1, double beta2,
                              double eps, uint32_t step, float grad_scale, int zero_grad, nafb_stream_t stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq) NAFB_FAIL(NAFB_ERR_INVALID, "adam_step: null pointer");
    if (step == 0) NAFB_FAIL(NAFB_ERR_INVALID, "adam_step: step is 1-based");
    if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0)
        NAFB_FAIL(NAFB_ERR_INVALID, "adam_step: buffers must be 16-byte aligned");
    if (n == 0) return NAFB_OK;
    const AdamConst c = make_adam_const(lr, beta1, beta2, eps, step, grad_scale);
    const uint64_t n4 = n >> 2;
    uint64_t blocks = (n4 + 255) / 256;
    const uint64_t cap = (uint64_t)nafb_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    k_adam<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, c, zero_grad);
    NAFB_CHECK_LAUNCH("adam_step");
    return NAFB_OK;
}
