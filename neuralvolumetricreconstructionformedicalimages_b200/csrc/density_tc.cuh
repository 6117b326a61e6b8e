// density_tc.cuh -- helpers shared by the tcgen05 editions of the fused density kernels (density_tc.cu: forward + the
// single-role backward; density_bwd_tc.cu: the warp-specialised backward): operand-tile layout, weight images, the pair-merged
// gather, the warp-aggregated scatter, the encoding stash, column sums.
#pragma once
#include "common.cuh"
#include "sampler.cuh"
#include "umma.cuh"

// profiling knobs (env NAFB_DEBUG_SKIP, read once): bit 0 = skip the gradient scatter, bit 1 = skip the table
// gather (synthetic encodings), bit 4 = no warp aggregation, bit 5 = phase time stamps of the backward kernel,
// bit 6 = no gather token in the forward kernel, bits 8-13 / 16-21 = aggregation thresholds.
// Results are wrong with bits 0/1 set; used only to attribute kernel time.
int nafb_debug_flags();

namespace tc {


constexpr int TILE = 128;
constexpr uint32_t LBO = 128;  // bytes between adjacent 8-column chunks of a row group

// ---- weights in shared memory (bf16 hi / lo, rows = output feature, chunks along the input)
constexpr uint32_t W0_OFF = 0, W0_SBO = 512;      // 32 x 32
constexpr uint32_t W1_OFF = 2048, W1_SBO = 512;   // 32 x 32
constexpr uint32_t W2_OFF = 4096, W2_SBO = 1024;  // 32 x 64
constexpr uint32_t W_HALF = 8192;                 // bytes per (hi | lo) weight image

struct SmallParams {   // fp32: biases of the hidden layers, head weights + bias
    float b0[32], b1[32], b2[32], w3[32], b3;
};

__device__ __forceinline__ void load_weight_images(const nafb_mlp &mp, uint8_t *w_hi, uint8_t *w_lo, SmallParams *sp) {
    // one 16-byte chunk (8 consecutive inputs of one output row) per iteration
    auto fill = [&](const float *__restrict__ W, int in_dim, uint32_t off, uint32_t sbo) {
        const int chunks = in_dim / 8;
        for (int i = threadIdx.x; i < 32 * chunks; i += blockDim.x) {
            const int row = i / chunks, c = i - row * chunks;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldg(W + row * in_dim + c * 8 + k);
            umma::store_chunk_split(w_hi, w_lo, off + umma::canon_off(row, c, LBO, sbo), v);
        }
    };
    fill(mp.W[0], 32, W0_OFF, W0_SBO);
    fill(mp.W[1], 32, W1_OFF, W1_SBO);
    fill(mp.W[2], 64, W2_OFF, W2_SBO);
    if (threadIdx.x < 32) {
        sp->b0[threadIdx.x] = __ldg(mp.b[0] + threadIdx.x);
        sp->b1[threadIdx.x] = __ldg(mp.b[1] + threadIdx.x);
        sp->b2[threadIdx.x] = __ldg(mp.b[2] + threadIdx.x);
        sp->w3[threadIdx.x] = __ldg(mp.W[3] + threadIdx.x);
        if (threadIdx.x == 0) sp->b3 = __ldg(mp.b[3]);
    }
}

// Gradient scatter of ONE level of one point: d(encoding) of the level is read from the point's TMEM lane
// (`taddr`: C columns), the 8 corner reductions go to the gradient table.  One copy of this code serves all
// levels and all call sites (__noinline__): the backward kernel calls it from the wait slots of the NEXT tile's
// MMA chain, so the reductions drain through the LSU while the tensor core and the epilogues work.
//
// Coarse levels are WARP-AGGREGATED: the 32 lanes of a warp hold 32 consecutive sample points (for the
// ray source: consecutive samples of one ray), which fall into a few grid cells only.  Lanes of one
// cell form a contiguous run; a segmented suffix sum over the run (5 shuffle steps per value) leaves
// the run totals of the 8 corners x C channels in the run's first lane, which issues the only
// reductions of the run.  The L2 atomic unit serialises per address (and the few hot lines of a coarse
// level live in a handful of L2 slices), so the number of reductions -- not their bytes -- is what the
// backward pass pays for.  The choice is made per warp and level from the number of runs (ballot).
constexpr int AGG_LEVELS = 8;      // levels 0 .. AGG_LEVELS-1 may be aggregated
constexpr int AGG_MAX_RUNS = 20;   // aggregate when the warp has at most this many runs
// (both can be overridden for experiments through NAFB_DEBUG_SKIP: bits 8-13 = levels + 1, bits 16-21 = runs + 1)

// One level of the scatter for the point this lane holds.  The level's gradient `ge` comes either from TMEM (`taddr`, the fused
// backward kernel: the load is issued first and awaited after the index arithmetic) or from registers (`ge_in`, the native-op
// backward of hashgrid.cu).  All 32 lanes of the warp must call it (shuffles); `valid` masks lanes without a point.
template <int C, bool FROM_TMEM>
__device__ __forceinline__ void scatter_level(const LevelParams lp, const float x0, const float x1, const float x2, const uint32_t taddr,
                                              const float *ge_in, const bool valid, const int agg_max_runs, float *__restrict__ grad_table) {
    const unsigned lane = threadIdx.x & 31u;
    float ge[C];
    if constexpr (FROM_TMEM) {
        umma::tmem_ldn<C>(taddr, ge);
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) ge[c] = ge_in[c];
    }
    float *tab = grad_table + (size_t)lp.offset * C;
    const uint32_t par = addr_parity8(tab);
    uint32_t g[3];
    float f[3];
    locate(x0, lp.scale, g[0], f[0]);
    locate(x1, lp.scale, g[1], f[1]);
    locate(x2, lp.scale, g[2], f[2]);
    const CellTerms ct = cell_terms3(lp, g[0], g[1], g[2]);
    uint32_t e[8];
    cell_entries8(lp, ct, e);
    if constexpr (FROM_TMEM) umma::tmem_wait_ld();
    if (agg_max_runs > 0) {
        // runs of equal cells among consecutive lanes (invalid lanes: a run of their own, never issued)
        const uint32_t k0 = valid ? (g[0] | (g[1] << 16)) : 0xffffffffu, k1 = valid ? g[2] : 0xffffffffu;
        const uint32_t p0 = __shfl_up_sync(0xffffffffu, k0, 1), p1 = __shfl_up_sync(0xffffffffu, k1, 1);
        const bool head = lane == 0 || p0 != k0 || p1 != k1;
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        if (__popc(heads) <= agg_max_runs) {
            float v[8][C];
#pragma unroll
            for (uint32_t idx = 0; idx < 8; ++idx) {
                float w = 1.0f;
#pragma unroll
                for (int d = 0; d < 3; ++d) w = __fmul_rn(w, (idx & (1u << d)) ? f[d] : __fsub_rn(1.0f, f[d]));
#pragma unroll
                for (int c = 0; c < C; ++c) v[idx][c] = valid ? __fmul_rn(w, ge[c]) : 0.f;
            }
            const uint32_t above = lane == 31 ? 0xffffffffu : (heads >> (lane + 1));  // bit j: lane+1+j starts a new run
            // only as many doubling steps as the longest run of the warp needs
            const uint32_t my_len = head ? (above ? (uint32_t)__ffs(above) : 32u - lane) : 0u;
            const uint32_t max_len = __reduce_max_sync(0xffffffffu, my_len);
            for (uint32_t o = 1; o < max_len; o <<= 1) {
                const bool same = (lane + o < 32) && ((above & ((1u << o) - 1u)) == 0u);
#pragma unroll
                for (uint32_t idx = 0; idx < 8; ++idx)
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float up = __shfl_down_sync(0xffffffffu, v[idx][c], o);
                        if (same) v[idx][c] += up;
                    }
            }
            if (head && valid) {
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)
                    red_add_entry_pair<C>(tab, par, e[2 * j], e[2 * j + 1], v[2 * j], v[2 * j + 1]);
            }
            return;
        }
    }
    if (valid) {
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            float v0[C], v1[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                // same products as the corner loop of the reference: ((1 * a0) * a1) * a2 evaluated as (a0 * a1) * a2
                v0[c] = __fmul_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.0f, f[0]), (j & 1u) ? f[1] : __fsub_rn(1.0f, f[1])), (j >> 1) ? f[2] : __fsub_rn(1.0f, f[2])), ge[c]);
                v1[c] = __fmul_rn(__fmul_rn(__fmul_rn(f[0], (j & 1u) ? f[1] : __fsub_rn(1.0f, f[1])), (j >> 1) ? f[2] : __fsub_rn(1.0f, f[2])), ge[c]);
            }
            red_add_entry_pair<C>(tab, par, e[2 * j], e[2 * j + 1], v0, v1);
        }
    }
}

// the fused backward kernel's entry: one __noinline__ copy serves all levels (the fully unrolled version spent 17 % of its samples in
// instruction-cache misses)
template <int C>
__device__ __noinline__ void scatter_one(const LevelParams *__restrict__ lvs, const int l, const float x0, const float x1, const float x2,
                                         const uint32_t taddr, const bool valid, const int agg_max_runs, float *__restrict__ grad_table) {
    scatter_level<C, true>(lvs[l], x0, x1, x2, taddr, nullptr, valid, agg_max_runs, grad_table);
}

// Backward pass without a stash: gather the thread's 16 encoding columns level by level (rolled loop, one copy of
// the code) and drop them as bf16 (hi, lo) into the operand tile.
template <int C>
__device__ __noinline__ void gather_half_to_smem(const LevelParams *__restrict__ lvs, const float *__restrict__ table, const float x0,
                                                 const float x1, const float x2, const int half, uint8_t *hi, uint8_t *lo, const uint32_t row,
                                                 const uint32_t chunk0, const uint32_t sbo) {
    constexpr int LH = 16 / C;
#pragma unroll 1
    for (int li = 0; li < LH; ++li) {
        const LevelParams lp = lvs[half * LH + li];
        const float *__restrict__ tab = table + (size_t)lp.offset * C;
        uint32_t g[3];
        float f[3];
        locate(x0, lp.scale, g[0], f[0]);
        locate(x1, lp.scale, g[1], f[1]);
        locate(x2, lp.scale, g[2], f[2]);
        const CellTerms ct = cell_terms3(lp, g[0], g[1], g[2]);
        uint32_t e[8];
        cell_entries8(lp, ct, e);
        float v[8][C];
#pragma unroll
        for (uint32_t idx = 0; idx < 8; ++idx) load_entry<C>(tab, e[idx], v[idx]);
        float res[C];
#pragma unroll
        for (int c = 0; c < C; ++c) res[c] = 0.f;
#pragma unroll
        for (uint32_t idx = 0; idx < 8; ++idx) {
            float w = 1.0f;
#pragma unroll
            for (int d = 0; d < 3; ++d) w = __fmul_rn(w, (idx & (1u << d)) ? f[d] : __fsub_rn(1.0f, f[d]));
#pragma unroll
            for (int c = 0; c < C; ++c) res[c] = __fmaf_rn(w, v[idx][c], res[c]);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const uint32_t col = (uint32_t)(li * C + c);
            const uint32_t off = umma::canon_off(row, chunk0 + 2 * half + (col >> 3), LBO, sbo) + (col & 7u) * 2u;
            __nv_bfloat16 h, lw;
            umma::split_bf16(res[c], h, lw);
            *reinterpret_cast<__nv_bfloat16 *>(hi + off) = h;
            *reinterpret_cast<__nv_bfloat16 *>(lo + off) = lw;
        }
    }
}

// write this thread's 16 values (two chunks) of a 32-wide block starting at chunk `chunk0`
__device__ __forceinline__ void store_half_row(uint8_t *hi, uint8_t *lo, uint32_t row, uint32_t chunk0, int half, uint32_t sbo, const float (&v)[16]) {
    umma::store_chunk_split(hi, lo, umma::canon_off(row, chunk0 + 2 * half, LBO, sbo), v);
    umma::store_chunk_split(hi, lo, umma::canon_off(row, chunk0 + 2 * half + 1, LBO, sbo), v + 8);
}

// ---- encoding stash: what a training forward leaves for the backward pass (16 KB per 128-point tile):
// the bf16 (hi | lo) images of the tile's encodings in the canonical layout with 4 chunks per row
// (SBO 512), i.e. byte-for-byte what the tensor core consumed.  A warp writes / reads whole 128-byte
// lines; the backward pass then needs no table gather at all.
// Behind the two images: a 2 KB tail with what the backward pass would otherwise recompute per point from the ray -- delta_i |d|
// of the ray integral (render.py:192-201) and the normalised position x01 (hashgrid.py:125): [delta | x | y | z] x 128 fp32.
constexpr uint32_t ST_SBO = 512, ST_HALF = 8192, ST_ENC = 16384, ST_TAIL_DELTA = ST_ENC, ST_TAIL_X01 = ST_ENC + 512, ST_TILE = ST_ENC + 2048;

// sign of the stored activations (hi part is enough): slope of LeakyReLU at h
__device__ __forceinline__ void lrelu_slopes(const uint8_t *hi, uint32_t row, uint32_t chunk0, int half, uint32_t sbo, float (&s)[16]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const uint4 q = *reinterpret_cast<const uint4 *>(hi + umma::canon_off(row, chunk0 + 2 * half + c, LBO, sbo));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // bf16 > 0  <=>  sign bit clear and not zero
            const uint32_t a = w[i] & 0xFFFFu, b = w[i] >> 16;
            s[c * 8 + 2 * i] = (a != 0u && !(a & 0x8000u)) ? 1.0f : 0.01f;
            s[c * 8 + 2 * i + 1] = (b != 0u && !(b & 0x8000u)) ? 1.0f : 0.01f;
        }
    }
}

// column sums over the 32 lanes of a warp of 16 per-lane values: afterwards lane (j & 15) and
// lane (j & 15) + 16 hold the sum of column j.  Recursive halving: 8+4+2+1+1 = 16 shuffles.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], unsigned lane) {
    float a[8];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = up ? v[i] : v[i + 8];
            const float keep = up ? v[i + 8] : v[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    float b[4];
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = up ? a[i] : a[i + 4];
            const float keep = up ? a[i + 4] : a[i];
            b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    float c[2];
    {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = up ? b[i] : b[i + 2];
            const float keep = up ? b[i + 2] : b[i];
            c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    float d;
    {
        const bool up = lane & 2;
        const float send = up ? c[0] : c[1];
        const float keep = up ? c[1] : c[0];
        d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;  // column index held by this lane: colsum_index(lane)
}
// which of the 16 columns ends up in `lane` after warp_colsum16
__device__ __forceinline__ int colsum_index(unsigned lane) {
    return ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
}

struct TileCtl {
    uint64_t mbar;
    uint32_t tmem_base;
    uint32_t pad;
};


}  // namespace tc
