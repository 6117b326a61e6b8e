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
//   v  = v*b2 ; v = addcmul(v, g, g, 1-b2) -> fma((1-b2)*g, g, v)
//   dn = sqrt(v) / sqrt(1-b2^t) + eps
//   p  = addcdiv(p, m, dn, -lr/(1-b1^t)) -> fma(-step_size, m/dn, p)
#include <math.h>

#include "common.cuh"

namespace {

struct AdamConst {
    float w1;        // 1 - beta1
    float beta2;
    float w2;        // 1 - beta2
    float bc2_sqrt;  // sqrt(1 - beta2^t)
    float eps;
    float neg_step;  // -(lr / (1 - beta1^t))
    float gscale;
};

__device__ __forceinline__ void adam_one(float &p, float &g, float &m, float &v, const AdamConst &c) {
    const float gg = c.gscale == 1.0f ? g : __fmul_rn(g, c.gscale);
    m = __fmaf_rn(c.w1, __fsub_rn(gg, m), m);
    v = __fmul_rn(v, c.beta2);
    v = __fmaf_rn(__fmul_rn(c.w2, gg), gg, v);
    const float dn = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), c.bc2_sqrt), c.eps);
    p = __fmaf_rn(c.neg_step, __fdiv_rn(m, dn), p);
}

__global__ void __launch_bounds__(256) k_adam(float *__restrict__ param, float *__restrict__ grad, float *__restrict__ m_, float *__restrict__ v_,
                                              uint64_t n, AdamConst c, int zero_grad) {
    const uint64_t n4 = n >> 2;
    float4 *p4 = reinterpret_cast<float4 *>(param);
    float4 *g4 = reinterpret_cast<float4 *>(grad);
    float4 *m4 = reinterpret_cast<float4 *>(m_);
    float4 *v4 = reinterpret_cast<float4 *>(v_);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
        adam_one(p.x, g.x, m.x, v.x, c);
        adam_one(p.y, g.y, m.y, v.y, c);
        adam_one(p.z, g.z, m.z, v.z, c);
        adam_one(p.w, g.w, m.w, v.w, c);
        p4[i] = p; m4[i] = m; v4[i] = v;
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

}  // namespace

extern "C" int nafb_adam_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq, uint64_t n, float lr, float beta1, float beta2,
                              float eps, uint32_t step, float grad_scale, int zero_grad, nafb_stream_t stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq) NAFB_FAIL(NAFB_ERR_INVALID, "adam_step: null pointer");
    if (step == 0) NAFB_FAIL(NAFB_ERR_INVALID, "adam_step: step is 1-based");
    if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0)
        NAFB_FAIL(NAFB_ERR_INVALID, "adam_step: buffers must be 16-byte aligned");
    if (n == 0) return NAFB_OK;
    // python-float (double) scalar maths of torch/optim/adam.py, then one cast to fp32
    const double b1 = (double)beta1, b2 = (double)beta2;
    const double bc1 = 1.0 - pow(b1, (double)step);
    const double bc2 = 1.0 - pow(b2, (double)step);
    AdamConst c;
    c.w1 = (float)(1.0 - b1);
    c.beta2 = beta2;
    c.w2 = (float)(1.0 - b2);
    c.bc2_sqrt = (float)sqrt(bc2);
    c.eps = eps;
    c.neg_step = (float)(-((double)lr / bc1));
    c.gscale = grad_scale;
    const uint64_t n4 = n >> 2;
    uint64_t blocks = (n4 + 255) / 256;
    const uint64_t cap = (uint64_t)nafb_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    k_adam<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, c, zero_grad);
    NAFB_CHECK_LAUNCH("adam_step");
    return NAFB_OK;
}
