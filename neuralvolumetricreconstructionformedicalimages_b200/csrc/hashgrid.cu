// hashgrid.cu -- the multi-resolution hash-grid op at the reference's native-FFI boundary
// (replaces src/encoder/hashencoder/src/hashencoder.cu kernel_grid / kernel_grid_backward /
// kernel_input_backward).  fp32 tables.  Level constants travel in the kernel parameter block,
// table entries are fetched with one vector load per corner (float2 for C=2) on the read-only
// path, gradients leave with one vector `red.global.add` per corner.
#include "common.cuh"
#include "density_tc.cuh"   // pair-merged loads / reductions and the warp-aggregated scatter shared with the fused kernels

namespace {

// corner weights in the reference's op order: w = ((1*a0)*a1)*a2 (hashencoder.cu:120-132)
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

// one (point, level): gather + D-linear interpolation, optional dy/dx.
template <int D, int C>
__device__ __forceinline__ void encode_level(const GridParams &gp, uint32_t level, const float (&x01)[D], float (&res)[C],
                                             float *dy_dx /* [D*C] or nullptr */) {
    const LevelParams lp = gp.lv[level];
    const float *__restrict__ tab = gp.table + (size_t)lp.offset * C;
    uint32_t g[D];
    float f[D];
#pragma unroll
    for (int d = 0; d < D; ++d) locate(x01[d], lp.scale, g[d], f[d]);
    float v[1 << D][C];
    if constexpr (D == 3) {   // hoisted index terms; the two x-neighbours of a (y, z) corner share one 128-bit load when adjacent + aligned
        const uint32_t par = addr_parity8(tab);
        const CellTerms ct = cell_terms3(lp, g[0], g[1], g[2]);
        uint32_t e[8];
        cell_entries8(lp, ct, e);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) load_entry_pair<C>(tab, par, e[2 * j], e[2 * j + 1], v[2 * j], v[2 * j + 1]);
    } else {
#pragma unroll
        for (uint32_t idx = 0; idx < (1u << D); ++idx) load_entry<C>(tab, corner_entry<D>(lp, g, idx), v[idx]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) res[c] = 0.f;
#pragma unroll
    for (uint32_t idx = 0; idx < (1u << D); ++idx) {
        const float w = corner_weight<D>(f, idx);
#pragma unroll
        for (int c = 0; c < C; ++c) res[c] = __fmaf_rn(w, v[idx][c], res[c]);  // hashencoder.cu:139
    }
    if (dy_dx) {
        // hashencoder.cu:153-197 with the intended axis skip (nd >= gd); see DESIGN.md / oracle note.
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
                    const int d = nd >= gd ? nd + 1 : nd;
                    const bool hi = (sub >> nd) & 1u;
                    w = __fmul_rn(w, hi ? f[d] : __fsub_rn(1.0f, f[d]));
                    idx_l |= hi ? (1u << d) : 0u;
                }
                const uint32_t idx_r = idx_l | (1u << gd);
#pragma unroll
                for (int c = 0; c < C; ++c) rg[c] = __fmaf_rn(w, __fsub_rn(v[idx_r][c], v[idx_l][c]), rg[c]);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) dy_dx[gd * C + c] = rg[c];
        }
    }
}

// ---- forward, reference layout [L,B,C]: thread = (point, level), level = blockIdx.y
template <int D, int C>
__global__ void __launch_bounds__(256) k_hash_fwd_lbc(GridParams gp, const float *__restrict__ inputs, float *__restrict__ outputs,
                                                       uint32_t B, float *__restrict__ dy_dx) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    float x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = __ldg(inputs + (size_t)b * D + d);
    float res[C];
    float dd[D * C];
    encode_level<D, C>(gp, level, x, res, dy_dx ? dd : nullptr);
    float *o = outputs + ((size_t)level * B + b) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = res[c];
    if (dy_dx) {
        float *q = dy_dx + ((size_t)b * gp.L + level) * (D * C);  // [B, L, D, C]  (hashencoder.cu:155)
#pragma unroll
        for (int i = 0; i < D * C; ++i) q[i] = dd[i];
    }
}

// ---- forward, fused-permute layout [B, L*C]: thread = (point, level) with level fastest so
// that a warp writes one contiguous run of the output.
template <int D, int C>
__global__ void __launch_bounds__(256) k_hash_fwd_blc(GridParams gp, const float *__restrict__ inputs, float *__restrict__ outputs,
                                                       uint32_t B, float *__restrict__ dy_dx) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t L = gp.L;
    if (t >= (uint64_t)B * L) return;
    const uint32_t b = (uint32_t)(t / L), level = (uint32_t)(t - (uint64_t)b * L);
    float x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = __ldg(inputs + (size_t)b * D + d);
    float res[C];
    float dd[D * C];
    encode_level<D, C>(gp, level, x, res, dy_dx ? dd : nullptr);
    float *o = outputs + t * C;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = res[c];
    if (dy_dx) {
        float *q = dy_dx + t * (D * C);
#pragma unroll
        for (int i = 0; i < D * C; ++i) q[i] = dd[i];
    }
}

// ---- forward, [B, L*C] layout, D == 3, no dy_dx: thread = POINT, levels in a loop with the loads of the next level in flight
// while one is interpolated (the gather of the fused forward kernel).  The lanes of a warp are 32 consecutive points working on
// the SAME level, so on the coarse levels they hit the same lines (with thread = (point, level) and the level fastest, every lane of
// a warp reads another level's table); each thread writes its point's whole 128-byte row.  Same loads, same operation order:
// bit-identical to k_hash_fwd_blc.
// The same for the shape every shipped configuration has (16 levels x 2 features: rows of 128 bytes): the outputs are staged in
// shared memory and every warp writes its 32 rows with whole-line stores (56 us against 63 us with one 16-byte store per level
// pair and thread, chest_50 point set).  DEPTH = levels of loads in flight ahead of the one being interpolated: 1, 2 and 3 measure
// the same (56-57 us; 4 and 8 cost occupancy: 70 us) -- the gather runs at 0.76 of the L2's random-sector rate whatever the depth
// (scripts/native_fwd_time.py, scripts/calls/r2_call48.sh).
template <int C, int DEPTH>
__global__ void __launch_bounds__(256) k_hash_fwd_point16(GridParams gp, const float *__restrict__ inputs, float *__restrict__ outputs, uint32_t B) {
    constexpr int LMAX = 16, ROW = LMAX * C + 4;
    __shared__ __align__(16) float stage[256 * ROW];
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    float x[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) x[d] = b < B ? __ldg(inputs + (size_t)b * 3 + d) : 0.5f;
    float v[LMAX][8][C];
    auto issue = [&](const int l) {
        const LevelParams lp = gp.lv[l];
        const float *__restrict__ tab = gp.table + (size_t)lp.offset * C;
        uint32_t g[3];
        float f[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) locate(x[d], lp.scale, g[d], f[d]);
        const uint32_t par = addr_parity8(tab);
        const CellTerms ct = cell_terms3(lp, g[0], g[1], g[2]);
        uint32_t e[8];
        cell_entries8(lp, ct, e);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) load_entry_pair<C>(tab, par, e[2 * j], e[2 * j + 1], v[l][2 * j], v[l][2 * j + 1]);
    };
#pragma unroll
    for (int l = 0; l < DEPTH; ++l) issue(l);
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
        if (l + DEPTH < LMAX) issue(l + DEPTH);
        uint32_t g;
        float f[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) locate(x[d], gp.lv[l].scale, g, f[d]);
        float res[C];
#pragma unroll
        for (int c = 0; c < C; ++c) res[c] = 0.f;
#pragma unroll
        for (uint32_t idx = 0; idx < 8; ++idx) {
            const float w = corner_weight<3>(f, idx);
#pragma unroll
            for (int c = 0; c < C; ++c) res[c] = __fmaf_rn(w, v[l][idx][c], res[c]);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) stage[(threadIdx.x) * ROW + l * C + c] = res[c];
    }
    // each warp writes its 32 rows (128 B each for L*C = 32) with whole-line stores: 8 lanes x 16 B per row, 4 rows per instruction
    __syncwarp();
    const uint32_t lane = threadIdx.x & 31u, w0 = threadIdx.x & ~31u;
    const uint32_t b0 = blockIdx.x * blockDim.x + w0;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const uint32_t row = it * 4 + (lane >> 3), col4 = lane & 7u;
        if (b0 + row < B) {
            const float4 q = *reinterpret_cast<const float4 *>(&stage[(w0 + row) * ROW + col4 * 4]);
            *reinterpret_cast<float4 *>(outputs + (size_t)(b0 + row) * (LMAX * C) + col4 * 4) = q;
        }
    }
}

template <int C>
__global__ void __launch_bounds__(256) k_hash_fwd_point(GridParams gp, const float *__restrict__ inputs, float *__restrict__ outputs, uint32_t B) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float x[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) x[d] = __ldg(inputs + (size_t)b * 3 + d);
    float *o = outputs + (size_t)b * gp.L * C;
    float v[2][8][C];
    auto issue = [&](const uint32_t l, const int slot) {
        const LevelParams lp = gp.lv[l];
        const float *__restrict__ tab = gp.table + (size_t)lp.offset * C;
        uint32_t g[3];
        float f[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) locate(x[d], lp.scale, g[d], f[d]);
        const uint32_t par = addr_parity8(tab);
        const CellTerms ct = cell_terms3(lp, g[0], g[1], g[2]);
        uint32_t e[8];
        cell_entries8(lp, ct, e);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) load_entry_pair<C>(tab, par, e[2 * j], e[2 * j + 1], v[slot][2 * j], v[slot][2 * j + 1]);
    };
    auto consume = [&](const uint32_t l, const int slot, float (&res)[C]) {
        uint32_t g;
        float f[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) locate(x[d], gp.lv[l].scale, g, f[d]);
#pragma unroll
        for (int c = 0; c < C; ++c) res[c] = 0.f;
#pragma unroll
        for (uint32_t idx = 0; idx < 8; ++idx) {
            const float w = corner_weight<3>(f, idx);
#pragma unroll
            for (int c = 0; c < C; ++c) res[c] = __fmaf_rn(w, v[slot][idx][c], res[c]);  // hashencoder.cu:139
        }
    };
    issue(0, 0);
    for (uint32_t l = 0; l < gp.L; l += 2) {
        float r0[C], r1[C];
        if (l + 1 < gp.L) issue(l + 1, 1);
        consume(l, 0, r0);
        if (l + 2 < gp.L) issue(l + 2, 0);
        if (l + 1 < gp.L) {
            consume(l + 1, 1, r1);
            if constexpr (C == 2) {
                if ((gp.L & 1u) == 0u) {   // rows are 16-byte aligned: one 128-bit store per level pair
                    *reinterpret_cast<float4 *>(o + (size_t)l * 2) = make_float4(r0[0], r0[1], r1[0], r1[1]);
                    continue;
                }
            }
#pragma unroll
            for (int c = 0; c < C; ++c) { o[(size_t)l * C + c] = r0[c]; o[(size_t)(l + 1) * C + c] = r1[c]; }
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) o[(size_t)l * C + c] = r0[c];
        }
    }
}

// ---- backward scatter: thread = (point, level); grad is [B, L*C] (BLC) or [L,B,C] (LBC).  D == 3: the scatter of the fused
// backward kernel (density_tc.cuh scatter_level): x-neighbour pairs leave as one red.v4 when adjacent + aligned, and on the coarse
// levels the lanes of a warp (32 consecutive points: consecutive samples of a ray in render()'s order) that fall into one cell are
// summed by shuffles and reduced once.
template <int D, int C>
__global__ void __launch_bounds__(256) k_hash_bwd(GridParams gp, const float *__restrict__ grad, const float *__restrict__ inputs,
                                                   float *__restrict__ grad_table, uint32_t B, int layout) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = b < B;
    if constexpr (D != 3) {
        if (!valid) return;
    }
    const uint32_t level = blockIdx.y;
    const LevelParams lp = gp.lv[level];
    float x[D];
    float gr[C];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = valid ? __ldg(inputs + (size_t)b * D + d) : 0.f;
    if (valid) {
        const float *gptr = layout == NAFB_LAYOUT_BLC ? grad + ((size_t)b * gp.L + level) * C : grad + ((size_t)level * B + b) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) gr[c] = __ldg(gptr + c);
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) gr[c] = 0.f;
    }
    if constexpr (D == 3) {
        tc::scatter_level<C, false>(lp, x[0], x[1], x[2], 0u, gr, valid, (int)level < tc::AGG_LEVELS ? tc::AGG_MAX_RUNS : 0, grad_table);
    } else {
        uint32_t g[D];
        float f[D];
#pragma unroll
        for (int d = 0; d < D; ++d) locate(x[d], lp.scale, g[d], f[d]);
        float *tab = grad_table + (size_t)lp.offset * C;
#pragma unroll
        for (uint32_t idx = 0; idx < (1u << D); ++idx) {
            const float w = corner_weight<D>(f, idx);
            float v[C];
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = __fmul_rn(w, gr[c]);  // hashencoder.cu:268
            red_add_entry<C>(tab, corner_entry<D>(lp, g, idx), v);
        }
    }
}

// ---- grad_inputs[b,d] += sum_l sum_c grad[b,l,c] * dy_dx[b,l,d,c]   (hashencoder.cu:275-298)
template <int D, int C>
__global__ void __launch_bounds__(256) k_hash_input_bwd(const float *__restrict__ grad, const float *__restrict__ dy_dx,
                                                         float *__restrict__ grad_inputs, uint32_t B, uint32_t L, int layout) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    float acc = grad_inputs[t];
    for (uint32_t l = 0; l < L; ++l) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float gv = layout == NAFB_LAYOUT_BLC ? grad[((size_t)b * L + l) * C + c] : grad[((size_t)l * B + b) * C + c];
            acc = __fmaf_rn(gv, dy_dx[(((size_t)b * L + l) * D + d) * C + c], acc);
        }
    }
    grad_inputs[t] = acc;
}

// ---- fused min/max (range check of hashgrid.py:122 in one pass)
__global__ void __launch_bounds__(256) k_minmax(const float *__restrict__ x, uint64_t n, float *out2) {
    float lo = INFINITY, hi = -INFINITY;
    bool nan = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        nan |= (v != v);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    nan = __any_sync(0xffffffffu, nan);
    if ((threadIdx.x & 31) == 0) {
        // float atomic min/max through the ordered-int trick
        auto amin = [](float *a, float v) {
            if (v >= 0) atomicMin(reinterpret_cast<int *>(a), __float_as_int(v));
            else atomicMax(reinterpret_cast<unsigned int *>(a), __float_as_uint(v));
        };
        auto amax = [](float *a, float v) {
            if (v >= 0) atomicMax(reinterpret_cast<int *>(a), __float_as_int(v));
            else atomicMin(reinterpret_cast<unsigned int *>(a), __float_as_uint(v));
        };
        amin(out2 + 0, lo);
        amax(out2 + 1, hi);
        if (nan) { out2[0] = NAN; out2[1] = NAN; }  // torch.min/max propagate NaN
    }
}
__global__ void k_minmax_init(float *out2) { out2[0] = INFINITY; out2[1] = -INFINITY; }

template <int D, int C>
int launch_fwd(const GridParams &gp, const float *inputs, float *outputs, uint32_t B, int layout, float *dy_dx, cudaStream_t s) {
    if (layout == NAFB_LAYOUT_LBC) {
        dim3 grid((B + 255) / 256, gp.L);
        k_hash_fwd_lbc<D, C><<<grid, 256, 0, s>>>(gp, inputs, outputs, B, dy_dx);
    } else if (D == 3 && !dy_dx && ((uintptr_t)outputs & 15) == 0) {
        if constexpr (D == 3) {
            if constexpr (C == 2) {
                if (gp.L == 16) k_hash_fwd_point16<2, 1><<<(B + 255) / 256, 256, 0, s>>>(gp, inputs, outputs, B);
                else k_hash_fwd_point<C><<<(B + 255) / 256, 256, 0, s>>>(gp, inputs, outputs, B);
            } else {
                k_hash_fwd_point<C><<<(B + 255) / 256, 256, 0, s>>>(gp, inputs, outputs, B);
            }
        }
    } else {
        const uint64_t n = (uint64_t)B * gp.L;
        k_hash_fwd_blc<D, C><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(gp, inputs, outputs, B, dy_dx);
    }
    NAFB_CHECK_LAUNCH("hash_encode_forward");
    return NAFB_OK;
}

template <int D, int C>
int launch_bwd(const GridParams &gp, const float *grad, const float *inputs, float *grad_table, uint32_t B, int layout,
               const float *dy_dx, float *grad_inputs, cudaStream_t s) {
    // thread = (point, level), level = blockIdx.y.  (A thread-per-point version with the levels looped, like the forward kernel, is
    // slower here: 104 us against 93 us -- the reductions want the sixteen-fold parallelism.)
    dim3 grid((B + 255) / 256, gp.L);
    k_hash_bwd<D, C><<<grid, 256, 0, s>>>(gp, grad, inputs, grad_table, B, layout);
    NAFB_CHECK_LAUNCH("hash_encode_backward");
    if (grad_inputs) {
        k_hash_input_bwd<D, C><<<(B * D + 255) / 256, 256, 0, s>>>(grad, dy_dx, grad_inputs, B, gp.L, layout);
        NAFB_CHECK_LAUNCH("hash_encode_backward(input)");
    }
    return NAFB_OK;
}

}  // namespace

#define DISPATCH_DC(D_, C_, CALL)                                    \
    do {                                                             \
        if (D_ == 2) {                                               \
            switch (C_) {                                            \
                case 1: return CALL(2, 1);                           \
                case 2: return CALL(2, 2);                           \
                case 4: return CALL(2, 4);                           \
                default: return CALL(2, 8);                          \
            }                                                        \
        } else {                                                     \
            switch (C_) {                                            \
                case 1: return CALL(3, 1);                           \
                case 2: return CALL(3, 2);                           \
                case 4: return CALL(3, 4);                           \
                default: return CALL(3, 8);                          \
            }                                                        \
        }                                                            \
    } while (0)

extern "C" {

int nafb_hash_encode_forward(const nafb_grid *grid, const float *inputs, float *outputs, uint32_t B, int out_layout,
                             int calc_grad_inputs, float *dy_dx, nafb_stream_t stream) {
    GridParams gp;
    int rc = nafb_make_grid_params(grid, &gp);
    if (rc) return rc;
    if (B == 0) return NAFB_OK;
    if (!inputs || !outputs) NAFB_FAIL(NAFB_ERR_INVALID, "hash_encode_forward: null pointer");
    if (calc_grad_inputs && !dy_dx) NAFB_FAIL(NAFB_ERR_INVALID, "hash_encode_forward: dy_dx required when calc_grad_inputs");
    if (out_layout != NAFB_LAYOUT_LBC && out_layout != NAFB_LAYOUT_BLC) NAFB_FAIL(NAFB_ERR_INVALID, "hash_encode_forward: bad layout");
    if (B == 0) return NAFB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    float *dd = calc_grad_inputs ? dy_dx : nullptr;
#define CALL(D_, C_) launch_fwd<D_, C_>(gp, inputs, outputs, B, out_layout, dd, s)
    DISPATCH_DC(gp.D, gp.C, CALL);
#undef CALL
}

int nafb_hash_encode_backward(const nafb_grid *grid, const float *grad, const float *inputs, float *grad_table, uint32_t B,
                              int grad_layout, int calc_grad_inputs, const float *dy_dx, float *grad_inputs, nafb_stream_t stream) {
    GridParams gp;
    int rc = nafb_make_grid_params(grid, &gp);
    if (rc) return rc;
    if (B == 0) return NAFB_OK;
    if (!grad || !inputs || !grad_table) NAFB_FAIL(NAFB_ERR_INVALID, "hash_encode_backward: null pointer");
    if (calc_grad_inputs && (!dy_dx || !grad_inputs)) NAFB_FAIL(NAFB_ERR_INVALID, "hash_encode_backward: dy_dx/grad_inputs required");
    if (grad_layout != NAFB_LAYOUT_LBC && grad_layout != NAFB_LAYOUT_BLC) NAFB_FAIL(NAFB_ERR_INVALID, "hash_encode_backward: bad layout");
    if (B == 0) return NAFB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    float *gi = calc_grad_inputs ? grad_inputs : nullptr;
#define CALL(D_, C_) launch_bwd<D_, C_>(gp, grad, inputs, grad_table, B, grad_layout, dy_dx, gi, s)
    DISPATCH_DC(gp.D, gp.C, CALL);
#undef CALL
}

int nafb_minmax(const float *x, uint64_t n, float *out2, nafb_stream_t stream) {
    if (!x || !out2) NAFB_FAIL(NAFB_ERR_INVALID, "minmax: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    k_minmax_init<<<1, 1, 0, s>>>(out2);
    if (n) {
        const unsigned blocks = (unsigned)((n + 256 * 8 - 1) / (256 * 8));
        const unsigned cap = (unsigned)nafb_sm_count() * 8;
        k_minmax<<<blocks < cap ? blocks : cap, 256, 0, s>>>(x, n, out2);
    }
    NAFB_CHECK_LAUNCH("minmax");
    return NAFB_OK;
}

}  // extern "C"
