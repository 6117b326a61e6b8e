// hashgrid_typed.cu -- the hash-grid op of hashgrid.cu for the other storage types the reference's FFI dispatches on
// (AT_DISPATCH_FLOATING_TYPES_AND_HALF, hashencoder.cu:392,423): fp16 tables / inputs / outputs -- what
// `@custom_fwd(cast_inputs=torch.half)` (hashgrid.py:12) feeds the op under autocast -- and fp64.  No shipped NAF
// config reaches these paths (the trainer never enables autocast); they exist so that the FFI is complete, and they follow
// the reference's mixed arithmetic type by type:
//   * positions are always fp32: pos = fma((float)x, scale, 0.5f)                       (hashencoder.cu:106-111)
//   * forward accumulates in a FLOAT register whatever the storage type (`float results[C]`, :114): fp16 -> fma(w, (float)h,
//     acc); fp64 -> acc = (float)fma((double)w, g, (double)acc), i.e. a double FMA rounded back to float per corner
//   * dy_dx: the neighbour difference is taken IN the storage type (a rounded half subtraction for fp16, :187)
//   * backward: fp16 products are rounded to half and leave as one `red.add.noftz.f16x2` per channel pair (:257-263) or a
//     scalar half reduction when C == 1; fp64 products are double and leave as `red.add.f64`
//   * grad_inputs accumulates in the storage type, one rounding per product and per add for fp16 (:295)
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

template <typename T> struct Storage;
template <> struct Storage<__half> {
    static __device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_float(float v) { return __float2half_rn(v); }
    // acc += w * v in the reference's promotion: float * Half -> float
    static __device__ __forceinline__ float accumulate(float w, __half v, float acc) { return __fmaf_rn(w, __half2float(v), acc); }
    static __device__ __forceinline__ __half sub(__half a, __half b) { return __float2half_rn(__fsub_rn(__half2float(a), __half2float(b))); }
};
template <> struct Storage<double> {
    static __device__ __forceinline__ float to_float(double v) { return (float)v; }
    static __device__ __forceinline__ double from_float(float v) { return (double)v; }
    // float * double -> double, += into a float register
    static __device__ __forceinline__ float accumulate(float w, double v, float acc) { return __double2float_rn(__fma_rn((double)w, v, (double)acc)); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
};

template <int D>
__device__ __forceinline__ float corner_weight(const float (&f)[D], uint32_t idx) {
    float w = 1.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) w = __fmul_rn(w, (idx & (1u << d)) ? f[d] : __fsub_rn(1.0f, f[d]));
    return w;
}

template <int D>
__device__ __forceinline__ uint32_t corner_entry(const LevelParams &lp, const uint32_t (&g)[D], uint32_t idx) {
    if constexpr (D == 3)
        return grid_entry3(lp, g[0] + (idx & 1u), g[1] + ((idx >> 1) & 1u), g[2] + ((idx >> 2) & 1u));
    else
        return grid_entry2(lp, g[0] + (idx & 1u), g[1] + ((idx >> 1) & 1u));
}

// C consecutive values of one table entry with the widest load the entry's size allows
template <typename T, int C>
__device__ __forceinline__ void load_entry_t(const T *__restrict__ tab, uint32_t entry, T (&v)[C]) {
    const T *p = tab + (size_t)entry * C;
    constexpr int BYTES = (int)sizeof(T) * C;
    if constexpr (BYTES == 4) {
        const uint32_t q = __ldg(reinterpret_cast<const uint32_t *>(p));
        memcpy(v, &q, 4);
    } else if constexpr (BYTES == 8) {
        const uint2 q = __ldg(reinterpret_cast<const uint2 *>(p));
        memcpy(v, &q, 8);
    } else if constexpr (BYTES >= 16) {
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p) + i);
            memcpy(reinterpret_cast<char *>(v) + 16 * i, &q, 16);
        }
    } else {
        v[0] = __ldg(p);   // one half
    }
}

// thread = (point, level); out_layout selects [L,B,C] (the FFI) or [B,L*C]
template <typename T, int D, int C>
__global__ void __launch_bounds__(256) k_hash_fwd_t(GridParams gp, const T *__restrict__ table, const T *__restrict__ inputs,
                                                     T *__restrict__ outputs, uint32_t B, int layout, T *__restrict__ dy_dx) {
    using S = Storage<T>;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const LevelParams lp = gp.lv[level];
    const T *__restrict__ tab = table + (size_t)lp.offset * C;
    uint32_t g[D];
    float f[D];
#pragma unroll
    for (int d = 0; d < D; ++d) locate(S::to_float(__ldg(inputs + (size_t)b * D + d)), lp.scale, g[d], f[d]);
    T v[1 << D][C];
#pragma unroll
    for (uint32_t idx = 0; idx < (1u << D); ++idx) load_entry_t<T, C>(tab, corner_entry<D>(lp, g, idx), v[idx]);
    float res[C];
#pragma unroll
    for (int c = 0; c < C; ++c) res[c] = 0.f;
#pragma unroll
    for (uint32_t idx = 0; idx < (1u << D); ++idx) {
        const float w = corner_weight<D>(f, idx);
#pragma unroll
        for (int c = 0; c < C; ++c) res[c] = S::accumulate(w, v[idx][c], res[c]);
    }
    T *o = layout == NAFB_LAYOUT_BLC ? outputs + ((size_t)b * gp.L + level) * C : outputs + ((size_t)level * B + b) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = S::from_float(res[c]);
    if (dy_dx) {
        T *q = dy_dx + ((size_t)b * gp.L + level) * (D * C);   // [B, L, D, C]
#pragma unroll
        for (int gd = 0; gd < D; ++gd) {
            float rg[C];
#pragma unroll
            for (int c = 0; c < C; ++c) rg[c] = 0.f;
#pragma unroll
            for (uint32_t sub = 0; sub < (1u << (D - 1)); ++sub) {
                float w = 1.0f;
                uint32_t idx_l = 0;
#pragma unroll
                for (int nd = 0; nd < D - 1; ++nd) {
                    const int d = nd >= gd ? nd + 1 : nd;   // the intended axis skip, see hashgrid.cu
                    const bool hi = (sub >> nd) & 1u;
                    w = __fmul_rn(w, hi ? f[d] : __fsub_rn(1.0f, f[d]));
                    idx_l |= hi ? (1u << d) : 0u;
                }
                const uint32_t idx_r = idx_l | (1u << gd);
#pragma unroll
                for (int c = 0; c < C; ++c) rg[c] = S::accumulate(w, S::sub(v[idx_r][c], v[idx_l][c]), rg[c]);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) q[gd * C + c] = S::from_float(rg[c]);
        }
    }
}

__device__ __forceinline__ void red_add_h2(__half *addr, __half a, __half b) {
    const __half2 v = __halves2half2(a, b);
    uint32_t bits;
    memcpy(&bits, &v, 4);
    asm volatile("red.relaxed.gpu.global.add.noftz.f16x2 [%0], %1;" ::"l"(addr), "r"(bits) : "memory");
}
__device__ __forceinline__ void red_add_h(__half *addr, __half a) {
    asm volatile("red.relaxed.gpu.global.add.noftz.f16 [%0], %1;" ::"l"(addr), "h"(__half_as_ushort(a)) : "memory");
}
__device__ __forceinline__ void red_add_d(double *addr, double a) {
    asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(addr), "d"(a) : "memory");
}

template <typename T, int D, int C>
__global__ void __launch_bounds__(256) k_hash_bwd_t(GridParams gp, const T *__restrict__ grad, const T *__restrict__ inputs,
                                                     T *__restrict__ grad_table, uint32_t B, int layout) {
    using S = Storage<T>;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const LevelParams lp = gp.lv[level];
    const T *gptr = layout == NAFB_LAYOUT_BLC ? grad + ((size_t)b * gp.L + level) * C : grad + ((size_t)level * B + b) * C;
    T gr[C];
#pragma unroll
    for (int c = 0; c < C; ++c) gr[c] = __ldg(gptr + c);
    uint32_t g[D];
    float f[D];
#pragma unroll
    for (int d = 0; d < D; ++d) locate(S::to_float(__ldg(inputs + (size_t)b * D + d)), lp.scale, g[d], f[d]);
    T *tab = grad_table + (size_t)lp.offset * C;
#pragma unroll
    for (uint32_t idx = 0; idx < (1u << D); ++idx) {
        const float w = corner_weight<D>(f, idx);
        T *p = tab + (size_t)corner_entry<D>(lp, g, idx) * C;
        if constexpr (sizeof(T) == 2) {
            // (__half)(grad * w): Half * float -> float, rounded once (hashencoder.cu:261,267)
            if constexpr (C == 1) {
                red_add_h(p, __float2half_rn(__fmul_rn(__half2float(gr[0]), w)));
            } else {
#pragma unroll
                for (int c = 0; c < C; c += 2)
                    red_add_h2(p + c, __float2half_rn(__fmul_rn(__half2float(gr[c]), w)), __float2half_rn(__fmul_rn(__half2float(gr[c + 1]), w)));
            }
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) red_add_d(p + c, __dmul_rn((double)w, gr[c]));   // float * double -> double (:267)
        }
    }
}

// grad_inputs[b,d] += sum_l sum_c grad[b,l,c] * dy_dx[b,l,d,c], every product and every add in the storage type
template <typename T, int D, int C>
__global__ void __launch_bounds__(256) k_hash_input_bwd_t(const T *__restrict__ grad, const T *__restrict__ dy_dx, T *__restrict__ grad_inputs,
                                                           uint32_t B, uint32_t L, int layout) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    T acc = grad_inputs[t];
    for (uint32_t l = 0; l < L; ++l) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const T gv = layout == NAFB_LAYOUT_BLC ? grad[((size_t)b * L + l) * C + c] : grad[((size_t)l * B + b) * C + c];
            const T dv = dy_dx[(((size_t)b * L + l) * D + d) * C + c];
            if constexpr (sizeof(T) == 2) {
                const __half prod = __float2half_rn(__fmul_rn(__half2float(gv), __half2float(dv)));
                acc = __float2half_rn(__fadd_rn(__half2float(acc), __half2float(prod)));
            } else {
                acc = __fma_rn(gv, dv, acc);
            }
        }
    }
    grad_inputs[t] = acc;
}

template <typename T, int D, int C>
int launch_fwd_t(const GridParams &gp, const void *table, const void *inputs, void *outputs, uint32_t B, int layout, void *dy_dx, cudaStream_t s) {
    dim3 grid((B + 255) / 256, gp.L);
    k_hash_fwd_t<T, D, C><<<grid, 256, 0, s>>>(gp, (const T *)table, (const T *)inputs, (T *)outputs, B, layout, (T *)dy_dx);
    NAFB_CHECK_LAUNCH("hash_encode_forward_dtype");
    return NAFB_OK;
}

template <typename T, int D, int C>
int launch_bwd_t(const GridParams &gp, const void *grad, const void *inputs, void *grad_table, uint32_t B, int layout, const void *dy_dx,
                 void *grad_inputs, cudaStream_t s) {
    dim3 grid((B + 255) / 256, gp.L);
    k_hash_bwd_t<T, D, C><<<grid, 256, 0, s>>>(gp, (const T *)grad, (const T *)inputs, (T *)grad_table, B, layout);
    NAFB_CHECK_LAUNCH("hash_encode_backward_dtype");
    if (grad_inputs) {
        k_hash_input_bwd_t<T, D, C><<<(B * D + 255) / 256, 256, 0, s>>>((const T *)grad, (const T *)dy_dx, (T *)grad_inputs, B, gp.L, layout);
        NAFB_CHECK_LAUNCH("hash_encode_backward_dtype(input)");
    }
    return NAFB_OK;
}

int typed_grid_params(const nafb_grid *grid, const void *table, GridParams *gp) {
    if (!grid) NAFB_FAIL(NAFB_ERR_INVALID, "nafb_grid: null pointer");
    nafb_grid g = *grid;
    g.table = (const float *)table;   // only its non-nullness is checked; the typed pointer travels separately
    return nafb_make_grid_params(&g, gp);
}

}  // namespace

#define DISPATCH_TDC(T_, D_, C_, CALL)                               \
    do {                                                             \
        if (D_ == 2) {                                               \
            switch (C_) {                                            \
                case 1: return CALL(T_, 2, 1);                       \
                case 2: return CALL(T_, 2, 2);                       \
                case 4: return CALL(T_, 2, 4);                       \
                default: return CALL(T_, 2, 8);                      \
            }                                                        \
        } else {                                                     \
            switch (C_) {                                            \
                case 1: return CALL(T_, 3, 1);                       \
                case 2: return CALL(T_, 3, 2);                       \
                case 4: return CALL(T_, 3, 4);                       \
                default: return CALL(T_, 3, 8);                      \
            }                                                        \
        }                                                            \
    } while (0)

extern "C" {

int nafb_hash_encode_forward_dtype(const nafb_grid *grid, int dtype, const void *table, const void *inputs, void *outputs, uint32_t B,
                                   int out_layout, int calc_grad_inputs, void *dy_dx, nafb_stream_t stream) {
    if (dtype == NAFB_DTYPE_F32) {
        if (!grid) NAFB_FAIL(NAFB_ERR_INVALID, "nafb_grid: null pointer");
        nafb_grid g = *grid;
        g.table = (const float *)table;
        return nafb_hash_encode_forward(&g, (const float *)inputs, (float *)outputs, B, out_layout, calc_grad_inputs, (float *)dy_dx, stream);
    }
    if (dtype != NAFB_DTYPE_F16 && dtype != NAFB_DTYPE_F64) NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "hash_encode_forward: inputs must be a floating tensor");
    GridParams gp;
    int rc = typed_grid_params(grid, table, &gp);
    if (rc) return rc;
    if (B == 0) return NAFB_OK;
    if (!inputs || !outputs) NAFB_FAIL(NAFB_ERR_INVALID, "hash_encode_forward: null pointer");
    if (calc_grad_inputs && !dy_dx) NAFB_FAIL(NAFB_ERR_INVALID, "hash_encode_forward: dy_dx required when calc_grad_inputs");
    if (out_layout != NAFB_LAYOUT_LBC && out_layout != NAFB_LAYOUT_BLC) NAFB_FAIL(NAFB_ERR_INVALID, "hash_encode_forward: bad layout");
    cudaStream_t s = (cudaStream_t)stream;
    void *dd = calc_grad_inputs ? dy_dx : nullptr;
#define CALL(T_, D_, C_) launch_fwd_t<T_, D_, C_>(gp, table, inputs, outputs, B, out_layout, dd, s)
    if (dtype == NAFB_DTYPE_F16) DISPATCH_TDC(__half, gp.D, gp.C, CALL);
    DISPATCH_TDC(double, gp.D, gp.C, CALL);
#undef CALL
}

int nafb_hash_encode_backward_dtype(const nafb_grid *grid, int dtype, const void *grad, const void *inputs, void *grad_table, uint32_t B,
                                    int grad_layout, int calc_grad_inputs, const void *dy_dx, void *grad_inputs, nafb_stream_t stream) {
    if (dtype == NAFB_DTYPE_F32) {
        if (!grid) NAFB_FAIL(NAFB_ERR_INVALID, "nafb_grid: null pointer");
        nafb_grid g = *grid;
        g.table = (const float *)grad_table;
        return nafb_hash_encode_backward(&g, (const float *)grad, (const float *)inputs, (float *)grad_table, B, grad_layout, calc_grad_inputs,
                                         (const float *)dy_dx, (float *)grad_inputs, stream);
    }
    if (dtype != NAFB_DTYPE_F16 && dtype != NAFB_DTYPE_F64) NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "hash_encode_backward: grad must be a floating tensor");
    GridParams gp;
    int rc = typed_grid_params(grid, grad_table, &gp);
    if (rc) return rc;
    if (B == 0) return NAFB_OK;
    if (!grad || !inputs || !grad_table) NAFB_FAIL(NAFB_ERR_INVALID, "hash_encode_backward: null pointer");
    if (calc_grad_inputs && (!dy_dx || !grad_inputs)) NAFB_FAIL(NAFB_ERR_INVALID, "hash_encode_backward: dy_dx/grad_inputs required");
    if (grad_layout != NAFB_LAYOUT_LBC && grad_layout != NAFB_LAYOUT_BLC) NAFB_FAIL(NAFB_ERR_INVALID, "hash_encode_backward: bad layout");
    cudaStream_t s = (cudaStream_t)stream;
    void *gi = calc_grad_inputs ? grad_inputs : nullptr;
#define CALL(T_, D_, C_) launch_bwd_t<T_, D_, C_>(gp, grad, inputs, grad_table, B, grad_layout, dy_dx, gi, s)
    if (dtype == NAFB_DTYPE_F16) DISPATCH_TDC(__half, gp.D, gp.C, CALL);
    DISPATCH_TDC(double, gp.D, gp.C, CALL);
#undef CALL
}

}  // extern "C"
