// common.cuh -- shared device helpers of libnafb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nafb200.h"

// ----------------------------------------------------------------------------- host side
void nafb_set_error(const char *fmt, ...);
#define NAFB_FAIL(code, ...)          \
    do {                              \
        nafb_set_error(__VA_ARGS__);  \
        return (code);                \
    } while (0)
#define NAFB_CHECK_LAUNCH(name)                                                        \
    do {                                                                               \
        cudaError_t e_ = cudaGetLastError();                                           \
        if (e_ != cudaSuccess) NAFB_FAIL(NAFB_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e_)); \
    } while (0)

// per-device helpers (api_common.cu): SM count of the CURRENT device; index of the current device clamped to the cache size
constexpr int NAFB_MAX_DEVICES = 64;
int nafb_sm_count();
int nafb_current_device();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device): `done` is a function-local
// static bool[NAFB_MAX_DEVICES] of the launcher
#define NAFB_CONFIGURE_SMEM(done, kernel, bytes, name)                                                         \
    do {                                                                                                       \
        const int dev_ = nafb_current_device();                                                                \
        if (!(done)[dev_]) {                                                                                   \
            cudaError_t e_ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
            if (e_ != cudaSuccess) NAFB_FAIL(NAFB_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e_));          \
            (done)[dev_] = true;                                                                               \
        }                                                                                                      \
    } while (0)

// ----------------------------------------------------------------------------- level table
// Per-level constants of the multi-resolution grid, evaluated once on the host and passed
// BY VALUE in the kernel parameter block (constant bank: no loads, no runtime `%` for the
// power-of-two levels).  Restates hashencoder.cu:55-74 and :98-100.
struct alignas(16) LevelParams {   // 32 bytes: two 128-bit loads when a kernel keeps the table in shared memory
    uint32_t offset;   // entry offset of the level                         (hashencoder.cu:94)
    uint32_t size;     // hashmap_size = offsets[l+1]-offsets[l]            (:98)
    uint32_t mask;     // size-1 if size is a power of two, else 0 (use `% size`)
    float scale;       // fma(exp2f(l), H, -1)                              (:99)
    uint32_t s1, s2;   // uint32-wrapped linear strides (res+1), (res+1)^2  (:61-65)
    uint32_t hashed;   // 1: xor-prime hash, 0: (wrapped) linear index      (:67-71)
    uint32_t pad;
};

struct GridParams {
    const float *table;
    uint32_t L, C, D, H;
    LevelParams lv[NAFB_MAX_LEVELS];
};

// Builds GridParams from the ABI descriptor. Returns NAFB_OK or an error (message set).
int nafb_make_grid_params(const nafb_grid *g, GridParams *out);

// ----------------------------------------------------------------------------- device side
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t wrap_index(uint32_t idx, const LevelParams &lp) {
    // (idx % size): power-of-two sizes are a mask; for the dense levels the linear index of an in-range
    // position is already < size, so the division is only reached by out-of-range inputs
    return lp.mask ? (idx & lp.mask) : (idx < lp.size ? idx : idx % lp.size);
}

// get_grid_index<3,C>(ch=0)/C  (hashencoder.cu:55-74); `linear` already carries the uint32 wrap.
__device__ __forceinline__ uint32_t grid_entry3(const LevelParams &lp, uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t hashed = x ^ (y * 19349663u) ^ (z * 83492791u);   // fast_hash<3>, :36-52
    const uint32_t linear = x + y * lp.s1 + z * lp.s2;
    return wrap_index(lp.hashed ? hashed : linear, lp);
}
__device__ __forceinline__ uint32_t grid_entry2(const LevelParams &lp, uint32_t x, uint32_t y) {
    const uint32_t hashed = x ^ (y * 19349663u);
    const uint32_t linear = x + y * lp.s1;
    return wrap_index(lp.hashed ? hashed : linear, lp);
}

// The 8 corner entries of one cell with the per-axis terms hoisted: hash(x, y, z) = x ^ y*P1 ^ z*P2 and the linear index
// x + y*s1 + z*s2 are both "combine(a[dx], b[dy], c[dz])", and (y+1)*m = y*m + m in uint32 arithmetic -- 2 multiplies
// per level instead of 16+.  Same uint32 wrap-around as get_grid_index (hashencoder.cu:55-74): bit-identical indices.
struct CellTerms {
    uint32_t a[2], b[2], c[2];
};
__device__ __forceinline__ CellTerms cell_terms3(const LevelParams &lp, uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t m1 = lp.hashed ? 19349663u : lp.s1, m2 = lp.hashed ? 83492791u : lp.s2;
    CellTerms t;
    t.a[0] = x; t.a[1] = x + 1u;
    t.b[0] = y * m1; t.b[1] = t.b[0] + m1;
    t.c[0] = z * m2; t.c[1] = t.c[0] + m2;
    return t;
}
// all 8 corner entries of the cell, corner idx = dx | dy << 1 | dz << 2, with ONE warp-uniform decision per level (instead
// of a select per corner): a hashed level always has a power-of-two size (it is hashed because (res+1)^3 exceeds 2^log2T);
// a linear level is masked when its size is a power of two (the wrapped "linear-overflow" levels) and otherwise dense, where
// the index of an in-range position needs no reduction at all.
__device__ __forceinline__ void cell_entries8(const LevelParams &lp, const CellTerms &t, uint32_t (&e)[8]) {
    if (lp.hashed) {
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) e[i] = (t.a[i & 1u] ^ t.b[(i >> 1) & 1u] ^ t.c[i >> 2]) & lp.mask;
    } else if (lp.mask) {
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) e[i] = (t.a[i & 1u] + t.b[(i >> 1) & 1u] + t.c[i >> 2]) & lp.mask;
    } else {
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) {
            const uint32_t l = t.a[i & 1u] + t.b[(i >> 1) & 1u] + t.c[i >> 2];
            e[i] = l < lp.size ? l : l % lp.size;
        }
    }
}
__device__ __forceinline__ uint32_t cell_entry(const LevelParams &lp, const CellTerms &t, uint32_t dx, uint32_t dy, uint32_t dz) {
    const uint32_t h = t.a[dx] ^ t.b[dy] ^ t.c[dz], l = t.a[dx] + t.b[dy] + t.c[dz];
    return wrap_index(lp.hashed ? h : l, lp);
}

// pos = fma(x, scale, 0.5); g = floor(pos); f = pos - g   (hashencoder.cu:106-111).
__device__ __forceinline__ void locate(float x01, float scale, uint32_t &g, float &f) {
    const float pos = __fmaf_rn(x01, scale, 0.5f);
    const float fl = floorf(pos);
    g = (uint32_t)fl;
    f = __fsub_rn(pos, (float)g);
}

// read-only vector loads of C consecutive floats of a table entry
template <int C> struct EntryVec;
template <> struct EntryVec<1> { float v[1]; };
template <> struct EntryVec<2> { float v[2]; };
template <> struct EntryVec<4> { float v[4]; };
template <> struct EntryVec<8> { float v[8]; };

template <int C>
__device__ __forceinline__ void load_entry(const float *__restrict__ base, uint32_t entry, float (&v)[C]) {
    if constexpr (C == 1) {
        v[0] = __ldg(base + entry);
    } else if constexpr (C == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(base) + entry);
        v[0] = t.x; v[1] = t.y;
    } else if constexpr (C == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(base) + entry);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        const float4 t0 = __ldg(reinterpret_cast<const float4 *>(base) + 2 * (size_t)entry);
        const float4 t1 = __ldg(reinterpret_cast<const float4 *>(base) + 2 * (size_t)entry + 1);
        v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w;
        v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
    }
}

// no-return vector reduction into the gradient table (red.global.add.v2.f32 on sm_90+)
__device__ __forceinline__ void red_add_f32x2(float *addr, float a, float b) {
    asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_f32x4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
__device__ __forceinline__ void red_add_f32(float *addr, float a) {
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
template <int C>
__device__ __forceinline__ void red_add_entry(float *base, uint32_t entry, const float (&v)[C]) {
    float *p = base + (size_t)entry * C;
    if constexpr (C == 1) red_add_f32(p, v[0]);
    else if constexpr (C == 2) red_add_f32x2(p, v[0], v[1]);
    else if constexpr (C == 4) red_add_f32x4(p, v[0], v[1], v[2], v[3]);
    else { red_add_f32x4(p, v[0], v[1], v[2], v[3]); red_add_f32x4(p + 4, v[4], v[5], v[6], v[7]); }
}

// ---- x-neighbour pairs.  The two x-neighbours of a (y, z) corner are ADJACENT table entries about half of the time:
// always on an xor-hashed level when x is even ((x ^ h) and ((x+1) ^ h) differ in bit 0 only), and e1 = e0 + 1 on a
// linear level.  With C == 2 an adjacent pair is 16 contiguous bytes; when its lower entry sits on a 16-byte boundary one
// 128-bit access serves both corners.  The gather is bound by address-divergent wavefronts and the scatter by the number
// of reductions (not by bytes), so every merged pair is an operation saved.  `par` = bit 3 of the level's base address
// ((uintptr_t)tab >> 3) & 1: entry e of the level is 16-byte aligned iff (e + par) is even.  Values and the order in
// which the caller combines them are unchanged (bit-exact).
__device__ __forceinline__ uint32_t addr_parity8(const void *p) { return (uint32_t)(reinterpret_cast<uintptr_t>(p) >> 3) & 1u; }

template <int C>
__device__ __forceinline__ void load_entry_pair(const float *__restrict__ tab, uint32_t par, uint32_t e0, uint32_t e1, float (&v0)[C],
                                                float (&v1)[C]) {
    if constexpr (C == 2) {
        // adjacent AND aligned <=> the two indices, shifted by the level's parity, differ in bit 0 only
        if (((e0 + par) ^ (e1 + par)) == 1u) {
            const uint32_t lo = e0 < e1 ? e0 : e1;
            const float4 q = __ldg(reinterpret_cast<const float4 *>(tab + 2 * (size_t)lo));
            const bool fwd = e0 < e1;
            v0[0] = fwd ? q.x : q.z; v0[1] = fwd ? q.y : q.w;
            v1[0] = fwd ? q.z : q.x; v1[1] = fwd ? q.w : q.y;
            return;
        }
    }
    load_entry<C>(tab, e0, v0);
    load_entry<C>(tab, e1, v1);
}

template <int C>
__device__ __forceinline__ void red_add_entry_pair(float *tab, uint32_t par, uint32_t e0, uint32_t e1, const float (&v0)[C],
                                                   const float (&v1)[C]) {
    if constexpr (C == 2) {
        if (((e0 + par) ^ (e1 + par)) == 1u) {
            const uint32_t lo = e0 < e1 ? e0 : e1;
            const bool fwd = e0 < e1;
            red_add_f32x4(tab + 2 * (size_t)lo, fwd ? v0[0] : v1[0], fwd ? v0[1] : v1[1], fwd ? v1[0] : v0[0], fwd ? v1[1] : v0[1]);
            return;
        }
    }
    red_add_entry<C>(tab, e0, v0);
    red_add_entry<C>(tab, e1, v1);
}

// ---- L2 residency hints (createpolicy + ld/st .L2::cache_hint): what the next kernel gathers from stays, streams go first
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld_l2hint(const float4 *a, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_l2hint(float4 *a, const float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// x > 0 ? x : 0.01 x, as max(x, 0.01 x): the same value for every input (0.01 x > x exactly when x < 0), one instruction less
__device__ __forceinline__ float leaky_relu(float x) { return fmaxf(x, __fmul_rn(0.01f, x)); }

__device__ __forceinline__ float head_activation(float x, uint32_t head) {
    switch (head) {
        case NAFB_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
        case NAFB_ACT_LRELU: return leaky_relu(x);
        case NAFB_ACT_TANH: return tanhf(x);
        default: return x;
    }
}
// derivative of the head wrt its pre-activation, given pre-activation x and output y
__device__ __forceinline__ float head_derivative(float x, float y, uint32_t head) {
    switch (head) {
        case NAFB_ACT_SIGMOID: return y * (1.0f - y);
        case NAFB_ACT_LRELU: return x > 0.f ? 1.0f : 0.01f;
        case NAFB_ACT_TANH: return 1.0f - y * y;
        default: return 1.0f;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__
