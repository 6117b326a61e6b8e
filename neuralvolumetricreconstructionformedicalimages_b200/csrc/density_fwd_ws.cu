// density_fwd_ws.cu -- warp-specialised forward kernel of the fused "sampling + hash-grid encoder + density MLP
// (+ ray integral)" path, tcgen05 edition (same configuration as density_tc.cu: L*C == 32, 4 x 32 MLP, skip at 2).
//
// The table gather is bound by the SM's address-divergent wavefront rate and by L2 latency, the MLP chain by the latency
// of its three dependent MMA round trips; in one instruction stream they simply add up.  Here they run in different
// warps of the CTA and overlap, tile after tile:
//
//   warps 0-7   epilogue ("chain") warps: thread (r, half) = TMEM lane r, 16 of the 32 feature columns.  TMEM -> bias,
//               LeakyReLU -> bf16 (hi, lo) operand of the next layer; head, ray integral.
//   warps 8-23  producers: thread (g, q) generates sample point g of the NEXT tile (ray generation, stratified
//               sampling, clamp, normalisation -- sampler.cuh, bit-exact), gathers quarter q of its levels (one 16-byte
//               operand chunk = 8 columns; loads of two levels in flight), writes it as bf16 (hi, lo) straight into the
//               operand buffer (double buffered) and into the stash for backward.
//   warp 24     one lane issues every tcgen05.mma of the CTA.
// One CTA per SM (800 threads): 512 gathering threads keep the L2 busy while a single chain runs (the chain of one tile is
// ~2.7 us, the gather of one tile ~5 us at the measured L2 sector rate), and the L1 keeps 150+ KB for the coarse levels.
//
// Hand-off: mbarriers full[b] (128 producer arrivals) / empty[b] (256 epilogue arrivals) around the two encoding
// buffers, one mbarrier for MMA completion (tcgen05.commit), named barrier 1 for "operands of the next layer written".
#include "common.cuh"
#include "sampler.cuh"
#include "umma.cuh"

int nafb_debug_flags();

namespace {

constexpr int TILE = 128;
constexpr int NT_WS = 800;          // 8 epilogue warps + 16 producer warps + 1 MMA warp
constexpr int MMA_WARP = 24;
constexpr uint32_t LBO = 128;
constexpr uint32_t E_SBO = 512, E_IMG = 8192;    // one encoding image (hi or lo): 128 rows x 4 chunks
constexpr uint32_t H_SBO = 512, H_IMG = 8192;    // hidden activations, same shape
constexpr uint32_t W0_OFF = 0, W0_SBO = 512, W1_OFF = 2048, W1_SBO = 512, W2_OFF = 4096, W2_SBO = 1024, W_IMG = 8192;
constexpr uint32_t ST_SBO = 512, ST_HALF = 8192, ST_TILE = 16384;   // stash tile: == one encoding buffer (hi | lo)

struct SmallParams {
    float b0[32], b1[32], b2[32], w3[32], b3;
};
struct Ctl {
    uint64_t full[2], empty[2], mma;
    uint32_t tmem_base, pad;
};
struct Meta {            // what the head needs about the points of a tile (written by the producers)
    float dn[TILE];      // delta * |d| of the sample (render.py:192-194)
    uint32_t ray[TILE];  // ray index, 0xffffffff for rows beyond P
};
constexpr uint32_t WS_SMEM = 2 * 2 * E_IMG + 2 * H_IMG + 2 * W_IMG + sizeof(SmallParams) + 16 + sizeof(Ctl) + 2 * sizeof(Meta) + TILE * sizeof(float) +
                             NAFB_MAX_LEVELS * sizeof(LevelParams) + 128;

__device__ __forceinline__ void mbar_arrive(uint64_t *mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void load_weight_images(const nafb_mlp &mp, uint8_t *w_hi, uint8_t *w_lo, SmallParams *sp) {
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

// 8 consecutive encoding columns (operand chunk `chunk`: levels chunk*LPC .. chunk*LPC+LPC-1) of one point; the loads of
// level li+1 are issued before level li is consumed
template <int C>
__device__ __forceinline__ void gather_chunk_ws(const LevelParams *__restrict__ lvs, const float *__restrict__ table, const float (&x01)[3],
                                                const int chunk, float (&enc8)[8]) {
    constexpr int LPC = 8 / C;
    float v[LPC][8][C];
    auto issue = [&](const int li) {
        const LevelParams lp = lvs[chunk * LPC + li];
        const float *__restrict__ tab = table + (size_t)lp.offset * C;
        uint32_t g[3];
        float f[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) locate(x01[d], lp.scale, g[d], f[d]);
        const uint32_t par = addr_parity8(tab);
        const CellTerms ct = cell_terms3(lp, g[0], g[1], g[2]);
        uint32_t e[8];
        cell_entries8(lp, ct, e);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
            load_entry_pair<C>(tab, par, e[2 * j], e[2 * j + 1], v[li][2 * j], v[li][2 * j + 1]);
    };
    auto consume = [&](const int li) {
        const float scale = lvs[chunk * LPC + li].scale;
        uint32_t g;
        float f[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) locate(x01[d], scale, g, f[d]);
        float res[C];
#pragma unroll
        for (int c = 0; c < C; ++c) res[c] = 0.f;
#pragma unroll
        for (uint32_t idx = 0; idx < 8; ++idx) {
            float w = 1.0f;
#pragma unroll
            for (int d = 0; d < 3; ++d) w = __fmul_rn(w, (idx & (1u << d)) ? f[d] : __fsub_rn(1.0f, f[d]));
#pragma unroll
            for (int c = 0; c < C; ++c) res[c] = __fmaf_rn(w, v[li][idx][c], res[c]);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) enc8[li * C + c] = res[c];
    };
    issue(0);
#pragma unroll
    for (int li = 0; li < LPC; ++li) {
        if (li + 1 < LPC) issue(li + 1);
        consume(li);
    }
}

template <int SRC, int C>
__global__ void __launch_bounds__(NT_WS, 1) k_density_fwd_ws(const GridParams gp, const nafb_mlp mp, const SamplerParams sp, const uint64_t P,
                                                             float *__restrict__ sigma, float *__restrict__ acc_out, float *__restrict__ z_out,
                                                             float *__restrict__ pts_out, int32_t *__restrict__ flags, uint8_t *__restrict__ stash) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *E = smem;                               // [2 buffers][hi | lo]
    uint8_t *H_hi = E + 4 * E_IMG, *H_lo = H_hi + H_IMG;
    uint8_t *W_hi = H_lo + H_IMG, *W_lo = W_hi + W_IMG;
    SmallParams *small = reinterpret_cast<SmallParams *>(W_lo + W_IMG);
    Ctl *ctl = reinterpret_cast<Ctl *>(reinterpret_cast<uint8_t *>(small) + ((sizeof(SmallParams) + 15) & ~15u));
    Meta *meta = reinterpret_cast<Meta *>(ctl + 1);  // [2]
    float *xchg = reinterpret_cast<float *>(meta + 2);
    LevelParams *lvs = reinterpret_cast<LevelParams *>(xchg + TILE);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    load_weight_images(mp, W_hi, W_lo, small);
    if (t < NAFB_MAX_LEVELS) lvs[t] = gp.lv[t];
    if (t == 0) {
        umma::mbar_init(&ctl->full[0], 4 * TILE);
        umma::mbar_init(&ctl->full[1], 4 * TILE);
        umma::mbar_init(&ctl->empty[0], 256);
        umma::mbar_init(&ctl->empty[1], 256);
        umma::mbar_init(&ctl->mma, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&ctl->tmem_base, 32);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = ctl->tmem_base;
    const uint64_t n_tiles = (P + TILE - 1) / TILE;

    if (warp >= 8 && warp < MMA_WARP) {
        // ============================================================ producers
        const int g = (t - 256) & 127, q = (t - 256) >> 7;   // row of the tile, which quarter of the levels (operand chunk)
        int bad = 0;
        uint32_t it = 0;
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t b = it & 1u;
            const uint64_t p = tile * TILE + g;
            const bool valid = p < P;
            float x[3] = {0.f, 0.f, 0.f};
            float dn = 0.f;
            uint32_t ray = 0xffffffffu;
            if (valid) {
                fetch_point<SRC>(sp, p, x);
                if (!(x[0] >= -sp.bound && x[0] <= sp.bound && x[1] >= -sp.bound && x[1] <= sp.bound && x[2] >= -sp.bound && x[2] <= sp.bound))
                    bad |= 1;
                if constexpr (SRC == NAFB_SRC_RAYS) if (q == 0) {
                    ray = (uint32_t)(p / sp.n_samples);
                    const uint32_t i = (uint32_t)(p - (uint64_t)ray * sp.n_samples);
                    if (pts_out) { pts_out[3 * p] = x[0]; pts_out[3 * p + 1] = x[1]; pts_out[3 * p + 2] = x[2]; }
                    if (acc_out || z_out) {
                        const RayRegs R = load_ray(sp, ray);
                        dn = ray_delta(sp, R, ray, i);
                        if (z_out) z_out[p] = z_sample(R.near, R.far, i, sp.n_samples, sp.lin_step, sp.perturb != 0, jitter_for(sp, ray));
                    }
                }
            }
            float x01[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) x01[d] = normalise01(x[d], sp.bound, sp.inv_2bound);
            // the buffer must have been released by the epilogue warps (first use of each buffer passes at once)
            umma::mbar_wait_backoff(&ctl->empty[b], ((it >> 1) & 1u) ^ 1u);
            uint8_t *e_hi = E + b * 2 * E_IMG, *e_lo = e_hi + E_IMG;
            uint8_t *st = stash ? stash + tile * ST_TILE : nullptr;
            {
                float enc8[8];
                gather_chunk_ws<C>(lvs, gp.table, x01, q, enc8);
                uint4 h, l;
                umma::split_chunk(enc8, h, l);
                const uint32_t off = umma::canon_off(g, q, LBO, E_SBO);
                *reinterpret_cast<uint4 *>(e_hi + off) = h;
                *reinterpret_cast<uint4 *>(e_lo + off) = l;
                if (st) {
                    *reinterpret_cast<uint4 *>(st + off) = h;
                    *reinterpret_cast<uint4 *>(st + ST_HALF + off) = l;
                }
            }
            if (q == 0) {
                meta[b].dn[g] = dn;
                meta[b].ray[g] = ray;
            }
            umma::fence_proxy_async();
            mbar_arrive(&ctl->full[b]);
        }
        if (flags && bad) atomicOr(flags, bad);
    } else if (warp == MMA_WARP) {
        // ============================================================ MMA issue
        const uint32_t e0 = umma::smem_u32(E), h_hi = umma::smem_u32(H_hi), h_lo = umma::smem_u32(H_lo);
        const uint32_t w_hi = umma::smem_u32(W_hi), w_lo = umma::smem_u32(W_lo);
        constexpr uint32_t IDESC = umma::idesc_bf16(128, 32, 0, 0);
        uint32_t it = 0;
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t b = it & 1u;
            const uint32_t e_hi = e0 + b * 2 * E_IMG, e_lo = e_hi + E_IMG;
            if (lane == 0) umma::mbar_wait_backoff(&ctl->full[b], (it >> 1) & 1u);
            __syncwarp();
            // (the epilogue warps' TMEM reads of the previous tile are ordered before this by their last arrival on barrier 1)
            if (it > 0) umma::named_bar_sync(1, 288);
            if (lane == 0) {
                umma::fence_after_sync();
                umma::mma_bf16x3(tmem, umma::make_desc(e_hi, LBO, E_SBO), umma::make_desc(e_lo, LBO, E_SBO), umma::make_desc(w_hi + W0_OFF, LBO, W0_SBO),
                                 umma::make_desc(w_lo + W0_OFF, LBO, W0_SBO), 256, 256, 2, IDESC, false);
                umma::commit(&ctl->mma);
            }
            __syncwarp();
            umma::named_bar_sync(1, 288);      // h0 written
            if (lane == 0) {
                umma::fence_after_sync();
                umma::mma_bf16x3(tmem, umma::make_desc(h_hi, LBO, H_SBO), umma::make_desc(h_lo, LBO, H_SBO), umma::make_desc(w_hi + W1_OFF, LBO, W1_SBO),
                                 umma::make_desc(w_lo + W1_OFF, LBO, W1_SBO), 256, 256, 2, IDESC, false);
                umma::commit(&ctl->mma);
            }
            __syncwarp();
            umma::named_bar_sync(1, 288);      // h1 written
            if (lane == 0) {
                umma::fence_after_sync();
                // layer 2 (skip): [enc | h1] . W2^T as two K = 32 halves
                umma::mma_bf16x3(tmem, umma::make_desc(e_hi, LBO, E_SBO), umma::make_desc(e_lo, LBO, E_SBO), umma::make_desc(w_hi + W2_OFF, LBO, W2_SBO),
                                 umma::make_desc(w_lo + W2_OFF, LBO, W2_SBO), 256, 256, 2, IDESC, false);
                umma::mma_bf16x3(tmem, umma::make_desc(h_hi, LBO, H_SBO), umma::make_desc(h_lo, LBO, H_SBO),
                                 umma::make_desc(w_hi + W2_OFF + 4 * LBO, LBO, W2_SBO), umma::make_desc(w_lo + W2_OFF + 4 * LBO, LBO, W2_SBO), 256, 256, 2,
                                 IDESC, true);
                umma::commit(&ctl->mma);
            }
            __syncwarp();
        }
    } else {
        // ============================================================ epilogue warps
        const int r = t & 127, half = t >> 7;
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 16u * half;
        uint32_t phase = 0;
        int bad = 0;
        auto wait_mma = [&](bool long_wait) {
            if (long_wait) umma::mbar_wait_backoff(&ctl->mma, phase, 200);
            else umma::mbar_wait(&ctl->mma, phase);
            phase ^= 1;
            umma::fence_after_sync();
        };
        auto operands_ready = [&]() {
            umma::fence_proxy_async();
            umma::fence_before_sync();
            umma::named_bar_arrive(1, 288);
        };
        auto store_h = [&](const float (&v)[16]) {
            umma::store_chunk_split(H_hi, H_lo, umma::canon_off(r, 2 * half, LBO, H_SBO), v);
            umma::store_chunk_split(H_hi, H_lo, umma::canon_off(r, 2 * half + 1, LBO, H_SBO), v + 8);
        };
        uint32_t it = 0;
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t b = it & 1u;
            const uint64_t p = tile * TILE + r;
            const bool valid = p < P;
            float v[16];
            // ---------------- layer 0 (this wait covers the producers' gather of the tile)
            wait_mma(true);
            umma::tmem_ld16(taddr, v);
            umma::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = leaky_relu(v[i] + small->b0[16 * half + i]);
            store_h(v);
            operands_ready();
            // ---------------- layer 1
            wait_mma(false);
            umma::tmem_ld16(taddr, v);
            umma::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = leaky_relu(v[i] + small->b1[16 * half + i]);
            store_h(v);
            operands_ready();
            // ---------------- layer 2 + head
            wait_mma(false);
            umma::tmem_ld16(taddr, v);
            umma::tmem_wait_ld();
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) part = __fmaf_rn(leaky_relu(v[i] + small->b2[16 * half + i]), small->w3[16 * half + i], part);
            if (half == 1) xchg[r] = part;
            umma::fence_before_sync();
            umma::named_bar_sync(2, 256);
            if (half == 0) {
                const float s = (part + xchg[r]) + small->b3;
                const float y = head_activation(s, mp.head);
                if (valid) {
                    if (sigma) sigma[p] = y;
                    if (!(fabsf(y) <= 3.4028234e38f)) bad |= 2;
                }
                if constexpr (SRC == NAFB_SRC_RAYS) {
                    if (acc_out) {  // warp-shuffle segmented reduction keyed by the ray id (render.py:201)
                        const uint32_t ray = meta[b].ray[r];
                        float contrib = valid ? __fmul_rn(y, meta[b].dn[r]) : 0.f;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const float up = __shfl_down_sync(0xffffffffu, contrib, o);
                            const uint32_t ur = __shfl_down_sync(0xffffffffu, ray, o);
                            if (lane + o < 32 && ur == ray) contrib += up;
                        }
                        const uint32_t prev = __shfl_up_sync(0xffffffffu, ray, 1);
                        if (valid && (lane == 0 || prev != ray)) atomicAdd(acc_out + ray, contrib);
                    }
                }
            }
            // this tile's encoding buffer, its meta block and xchg are free again; the TMEM reads above are ordered before the
            // next tile's first MMA by the arrival below (the MMA warp syncs on barrier 1 before issuing it)
            mbar_arrive(&ctl->empty[b]);
            umma::named_bar_sync(2, 256);
            umma::named_bar_arrive(1, 288);
        }
        if (flags && bad) atomicOr(flags, bad);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 32);
}

template <int SRC, int C>
int launch_ws(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, uint64_t P, float *sigma, float *acc, float *z, float *pts, int32_t *flags,
              uint8_t *stash, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_density_fwd_ws<SRC, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM);
        if (e != cudaSuccess) NAFB_FAIL(NAFB_ERR_CUDA, "density_forward(ws): %s", cudaGetErrorString(e));
        configured = true;
    }
    const uint64_t n_tiles = (P + TILE - 1) / TILE;
    const uint64_t cap = (uint64_t)nafb_sm_count();
    const unsigned grid = (unsigned)(n_tiles < cap ? n_tiles : cap);
    k_density_fwd_ws<SRC, C><<<grid, NT_WS, WS_SMEM, s>>>(gp, mp, sp, P, sigma, acc, z, pts, flags, stash);
    NAFB_CHECK_LAUNCH("density_forward(ws)");
    return NAFB_OK;
}

}  // namespace

int nafb_launch_fwd_ws(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, int src, uint64_t P, float *sigma, float *acc, float *z,
                       float *pts, int32_t *flags, void *stash, cudaStream_t s) {
#define CALL(S_, C_) launch_ws<S_, C_>(gp, mp, sp, P, sigma, acc, z, pts, flags, (uint8_t *)stash, s)
    switch (gp.C) {
        case 1: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 1) : src == NAFB_SRC_RAYS ? CALL(NAFB_SRC_RAYS, 1) : CALL(NAFB_SRC_VOXELS, 1);
        case 2: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 2) : src == NAFB_SRC_RAYS ? CALL(NAFB_SRC_RAYS, 2) : CALL(NAFB_SRC_VOXELS, 2);
        case 4: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 4) : src == NAFB_SRC_RAYS ? CALL(NAFB_SRC_RAYS, 4) : CALL(NAFB_SRC_VOXELS, 4);
        default: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 8) : src == NAFB_SRC_RAYS ? CALL(NAFB_SRC_RAYS, 8) : CALL(NAFB_SRC_VOXELS, 8);
    }
#undef CALL
}
