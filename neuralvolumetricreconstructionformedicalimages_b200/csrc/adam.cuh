// adam.cuh -- the arithmetic of one Adam update, shared by the dense kernel (adam.cu) and the fused exchange kernel
// (exchange.cu).  Follows torch's multi-tensor (foreach) Adam op by op, each op with its own rounding, so that the step can
// be compared bit for bit with torch.optim.Adam:
//   m  = lerp(m, g, 1-b1)                    -> fma(w, g - m, m)        (weight < 0.5 branch)
//   v  = v*b2 ; v = addcmul(v, g, g, 1-b2)   -> fma(1-b2, g*g, v)        (the foreach functor: input + scalar * (t1 * t2))
//   dn = sqrt(v) / sqrt(1-b2^t) + eps
//   p  = addcdiv(p, m, dn, -lr/(1-b1^t))     -> fma(-step_size, m/dn, p)
#pragma once
#include <math.h>

#include "common.cuh"

struct AdamConst {
    float w1;        // 1 - beta1
    float beta2;
    float w2;        // 1 - beta2
    float bc2_sqrt;  // sqrt(1 - beta2^t)
    float eps;
    float neg_step;  // -(lr / (1 - beta1^t))
    float gscale;    // applied to the incoming gradient (1/world for data parallel)
};

// python-float (double) scalar maths of torch/optim/adam.py, then one cast to fp32; host and device
// (beta1 / beta2 / eps / lr arrive as DOUBLES: python's 0.9 is not the double of 0.9f, and 1 - 0.999 would be off by 5e-5)
__host__ __device__ inline AdamConst make_adam_const(double lr, double b1, double b2, double eps, uint32_t step, float gscale) {
    const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
    AdamConst c;
    c.w1 = (float)(1.0 - b1);
    c.beta2 = (float)b2;
    c.w2 = (float)(1.0 - b2);
    c.bc2_sqrt = (float)pow(bc2, 0.5);   // adam.py: bias_correction2 ** 0.5
    c.eps = (float)eps;
    c.neg_step = (float)(-(lr / bc1));
    c.gscale = gscale;
    return c;
}

#ifdef __CUDACC__
__device__ __forceinline__ void adam_one(float &p, const float g, float &m, float &v, const AdamConst &c) {
    const float gg = c.gscale == 1.0f ? g : __fmul_rn(g, c.gscale);
    m = __fmaf_rn(c.w1, __fsub_rn(gg, m), m);
    v = __fmul_rn(v, c.beta2);
    v = __fmaf_rn(c.w2, __fmul_rn(gg, gg), v);
    const float dn = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), c.bc2_sqrt), c.eps);
    p = __fmaf_rn(c.neg_step, __fdiv_rn(m, dn), p);
}

// constants of the step that FOLLOWS the device state (state[STEP] completed steps, learning rate in state[LR])
__device__ __forceinline__ AdamConst adam_const_from_state(const uint32_t *state, double beta1, double beta2, double eps, float gscale) {
    const double lr = __hiloint2double((int)state[NAFB_STATE_LR + 1], (int)state[NAFB_STATE_LR]);
    return make_adam_const(lr, beta1, beta2, eps, state[NAFB_STATE_STEP] + 1u, gscale);
}
#endif
