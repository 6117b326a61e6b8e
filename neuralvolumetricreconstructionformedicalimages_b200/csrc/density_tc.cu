// density_tc.cu -- tcgen05 edition of the fused "hash-grid encoder + density MLP" kernels for the
// configuration every shipped YAML uses: L*C == 32 encoding, 4 layers x 32 hidden, skip at
// layer 2, out_dim 1 (reference src/network/network.py:34-58, config/*.yaml).
//
// One CTA = one 128-point tile at a time (persistent over tiles); 256 epilogue threads (+ one MMA-issue warp in backward).
//   thread t:  row r = t & 127 (sample point, == TMEM lane), half = t >> 7 owns feature
//              columns [16*half, 16*half+16) of every 32-wide activation of its point.
// Forward: sampling / ray generation (sampler.cuh) -> gather (pair-merged 128-bit loads) -> 3 MMA phases -> head -> ray
// integral; the encodings are left in the "stash" for backward.  Backward: stash -> forward chain -> head gradient ->
// 3 backward phases; d(encoding) stays in TMEM and is scattered (warp-aggregated, pair-merged) from the wait slots of the
// NEXT tile.  DESIGN.md section 4 has the measurements behind each of these choices.
// Activations are written ONCE, by the thread that owns the point, as bf16 (hi, lo) pairs in the
// canonical no-swizzle UMMA layout (umma.cuh) and consumed in place by the tensor core:
//   forward      h_l   = lrelu(X . W_l^T + b)        A = X  (K-major)      B = W_l  (K-major)
//   input grad   dX    = G . W_l                     A = G  (K-major)      B = W_l  (MN-major view)
//   weight grad  dW_l += G^T . X   (over 128 points) A = G  (MN-major view) B = X   (MN-major view)
// Accumulators live in TMEM (fp32); weight gradients stay in TMEM across all tiles of the CTA and
// are read out once.  Products are bf16x3 (hi*hi + hi*lo + lo*hi): |rel err| ~ 2^-16 per product.
// Bias, LeakyReLU, the 1-wide head, its gradient and the bias / head-weight gradients are fp32
// SIMT work in the epilogues (TMEM -> registers -> shared memory operand of the next MMA).
#include "density_tc.cuh"
#include "loss.cuh"

using namespace tc;

namespace {

// ================================================================================ forward
// smem: X_hi | X_lo : 128 rows x 8 chunks [enc(0-3) | h(4-7)], SBO 1024 -> 16 KB each
constexpr uint32_t FX_SBO = 1024, FX_HALF = 16384;
constexpr uint32_t FWD_SMEM = 2 * FX_HALF + 2 * W_HALF + sizeof(SmallParams) + sizeof(TileCtl) + TILE * sizeof(float) + 128;

template <int SRC, int C>
__global__ void __launch_bounds__(NT, FWD_CTAS) k_density_fwd_tc(const GridParams gp, const nafb_mlp mp, const SamplerParams sp, const uint64_t P,
                                                          float *__restrict__ sigma, float *__restrict__ acc_out, float *__restrict__ z_out,
                                                          float *__restrict__ pts_out, int32_t *__restrict__ flags, uint8_t *__restrict__ stash,
                                                          const int dbg, const nafb_loss_tail tail) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *X_hi = smem, *X_lo = X_hi + FX_HALF;
    uint8_t *W_hi = X_lo + FX_HALF, *W_lo = W_hi + W_HALF;
    SmallParams *small = reinterpret_cast<SmallParams *>(W_lo + W_HALF);
    TileCtl *ctl = reinterpret_cast<TileCtl *>(reinterpret_cast<uint8_t *>(small) + ((sizeof(SmallParams) + 15) & ~15u));
    float *xchg = reinterpret_cast<float *>(ctl + 1);  // [128] partial head dot products of half 1

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int r = t & 127, half = t >> 7;
    load_weight_images(mp, W_hi, W_lo, small);
    if (t == 0) {
        umma::mbar_init(&ctl->mbar, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&ctl->tmem_base, 32);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = ctl->tmem_base;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 16u * half;
    const uint32_t x_hi = umma::smem_u32(X_hi), x_lo = umma::smem_u32(X_lo), w_hi = umma::smem_u32(W_hi), w_lo = umma::smem_u32(W_lo);
    constexpr uint32_t IDESC = umma::idesc_bf16(128, 32, 0, 0);
    uint32_t phase = 0;
    int bad = 0;

    const uint64_t n_tiles = SRC == NAFB_SRC_VOXELS ? voxel_block_tiles(sp) : (P + TILE - 1) / TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t p = tile * TILE + r;      // index of the point in the outputs
        bool valid = p < P;
        float x[3] = {0.f, 0.f, 0.f};
        if constexpr (SRC == NAFB_SRC_VOXELS) valid = voxel_block_point(sp, tile, (uint32_t)r, x, p);   // 4 x 4 x 8 blocks of the lattice
        float z_mine = 0.f, delta_mine = 0.f;   // RAYS source: this sample's depth and its ray-integral weight delta_i |d|
        uint32_t ray_mine = 0xffffffffu;
        if (valid) {
            if constexpr (SRC == NAFB_SRC_RAYS) {
                ray_mine = P <= 0xffffffffull ? (uint32_t)p / sp.n_samples : (uint32_t)(p / sp.n_samples);
                ray_sample_and_delta(sp, ray_mine, (uint32_t)(p - (uint64_t)ray_mine * sp.n_samples), half == 0 && (acc_out || z_out || stash), x,
                                     z_mine, delta_mine);
            } else if constexpr (SRC != NAFB_SRC_VOXELS) {
                fetch_point<SRC>(sp, p, x);
            }
            if (!(x[0] >= -sp.bound && x[0] <= sp.bound && x[1] >= -sp.bound && x[1] <= sp.bound && x[2] >= -sp.bound && x[2] <= sp.bound))
                bad |= 1;
            if (SRC == NAFB_SRC_RAYS && pts_out && half == 0) {
                pts_out[3 * p] = x[0]; pts_out[3 * p + 1] = x[1]; pts_out[3 * p + 2] = x[2];
            }
        }
        float x01[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) x01[d] = normalise01(x[d], sp.bound, sp.inv_2bound);
        {
            float enc[16];
            if (dbg & 2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) enc[i] = x01[i % 3];
            } else {
                gather_half<C>(gp, x01, half, enc);
            }
            store_half_row_and_stash(X_hi, X_lo, r, 0, half, FX_SBO, enc, stash ? stash + tile * ST_TILE : nullptr);
            if (stash && half == 1) {   // tail of the stash tile: the normalised position (the backward scatter needs it again)
                float *tail = reinterpret_cast<float *>(stash + tile * ST_TILE + ST_TAIL_X01);
                tail[r] = x01[0]; tail[128 + r] = x01[1]; tail[256 + r] = x01[2];
            }
        }
        // ---------------- layer 0: enc . W0^T
        umma::fence_proxy_async();
        umma::fence_before_sync();
        __syncthreads();
        if (t == 0) {
            umma::fence_after_sync();
            umma::mma_bf16x3(tmem, umma::make_desc(x_hi, LBO, FX_SBO), umma::make_desc(x_lo, LBO, FX_SBO),
                             umma::make_desc(w_hi + W0_OFF, LBO, W0_SBO), umma::make_desc(w_lo + W0_OFF, LBO, W0_SBO), 256, 256, 2, IDESC, false);
            umma::commit(&ctl->mbar);
        }
        umma::mbar_wait(&ctl->mbar, phase);
        phase ^= 1;
        umma::fence_after_sync();
        float v[16];
        umma::tmem_ld16(taddr, v);
        umma::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = leaky_relu(v[i] + small->b0[16 * half + i]);
        store_half_row(X_hi, X_lo, r, 4, half, FX_SBO, v);
        // ---------------- layer 1: h0 . W1^T
        umma::fence_proxy_async();
        umma::fence_before_sync();
        __syncthreads();
        if (t == 0) {
            umma::fence_after_sync();
            umma::mma_bf16x3(tmem, umma::make_desc(x_hi + 4 * LBO, LBO, FX_SBO), umma::make_desc(x_lo + 4 * LBO, LBO, FX_SBO),
                             umma::make_desc(w_hi + W1_OFF, LBO, W1_SBO), umma::make_desc(w_lo + W1_OFF, LBO, W1_SBO), 256, 256, 2, IDESC, false);
            umma::commit(&ctl->mbar);
        }
        umma::mbar_wait(&ctl->mbar, phase);
        phase ^= 1;
        umma::fence_after_sync();
        umma::tmem_ld16(taddr, v);
        umma::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = leaky_relu(v[i] + small->b1[16 * half + i]);
        store_half_row(X_hi, X_lo, r, 4, half, FX_SBO, v);
        // ---------------- layer 2 (skip): [enc | h1] . W2^T, K = 64
        umma::fence_proxy_async();
        umma::fence_before_sync();
        __syncthreads();
        if (t == 0) {
            umma::fence_after_sync();
            umma::mma_bf16x3(tmem, umma::make_desc(x_hi, LBO, FX_SBO), umma::make_desc(x_lo, LBO, FX_SBO),
                             umma::make_desc(w_hi + W2_OFF, LBO, W2_SBO), umma::make_desc(w_lo + W2_OFF, LBO, W2_SBO), 256, 256, 4, IDESC, false);
            umma::commit(&ctl->mbar);
        }
        umma::mbar_wait(&ctl->mbar, phase);
        phase ^= 1;
        umma::fence_after_sync();
        umma::tmem_ld16(taddr, v);
        umma::tmem_wait_ld();
        // ---------------- head: sigma = act(w3 . lrelu(.) + b3), two half-row partial sums
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) part = __fmaf_rn(leaky_relu(v[i] + small->b2[16 * half + i]), small->w3[16 * half + i], part);
        if (half == 1) xchg[r] = part;
        umma::fence_before_sync();   // orders the TMEM reads above before the next tile's MMAs
        __syncthreads();
        if (half == 0) {
            const float s = (part + xchg[r]) + small->b3;
            const float y = head_activation(s, mp.head);
            if (valid) {
                if (sigma) sigma[p] = y;
                if (!(fabsf(y) <= 3.4028234e38f)) bad |= 2;
            }
            if constexpr (SRC == NAFB_SRC_RAYS) {
                if (acc_out || z_out || stash) {
                    float contrib = 0.f;
                    const uint32_t ray = ray_mine;
                    if (valid) {
                        if (stash) reinterpret_cast<float *>(stash + tile * ST_TILE + ST_TAIL_DELTA)[r] = delta_mine;
                        contrib = __fmul_rn(y, delta_mine);  // render.py:201
                        if (z_out) z_out[p] = z_mine;
                    }
                    if (acc_out) {  // warp-shuffle segmented reduction keyed by the ray id
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
        }
        __syncthreads();  // xchg / X are reused by the next tile
    }
    if (flags && bad) atomicOr(flags, bad);
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 32);
    // ---- loss tail (training): the last CTA to retire sees every ray integral (atomics into acc_out, made visible by the
    // fence + ticket) and evaluates the masked chunk-wise MSE and d loss / d acc -- what a separate nafb_mse_loss launch did.
    if (tail.ticket) {
        __shared__ uint32_t s_last;
        if (t == 0) {
            __threadfence();
            s_last = atomicAdd(tail.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            float *s_mean = reinterpret_cast<float *>(smem), *s_cnt = s_mean + MSE_GROUP;   // the operand tiles are dead by now
            const uint32_t n = sp.n_rays, chunk = (tail.chunk == 0 || tail.chunk > n) ? n : tail.chunk;
            mse_loss_block(acc_out, tail.target, tail.mask, n, chunk, tail.gscale, tail.loss_out, tail.dacc, tail.zero_pred, s_mean, s_cnt);
            if (t == 0) *tail.ticket = 0u;
        }
    }
}

// ================================================================================ backward
// smem: ALL_hi | ALL_lo : 128 rows x 20 chunks  [h0(0-3) | enc(4-7) | h1(8-11) | h2(12-15, unused) | G(16-19)], SBO 2560,
// plus 1536 B of slack so that the 128-feature window starting at the G block stays inside the allocation.
constexpr uint32_t BX_SBO = 2560, BX_HALF = 16 * 2560 + 1536;
constexpr uint32_t CH_H0 = 0, CH_ENC = 4, CH_H1 = 8, CH_G = 16;
constexpr uint32_t BWD_SMEM = 2 * BX_HALF + 2 * W_HALF + sizeof(SmallParams) + sizeof(TileCtl) + 2 * TILE * sizeof(float) + 8 * 80 * sizeof(float) +
                              NAFB_MAX_LEVELS * sizeof(LevelParams) + 128;
// TMEM columns.  d(encoding) is double buffered: the buffer of tile i is scattered while tile i+1 runs.
constexpr uint32_t T_S = 0, T_DENC0 = 32, T_DW0 = 64, T_DW1 = 96, T_DW2 = 128, T_DENC1 = 192, T_COLS = 256;

// offsets inside one CTA's slot of the partials workspace (floats) -- matches density.cu's MlpLayout for this net
constexpr int PW0 = 0, PW1 = 1024, PW2 = 2048, PW3 = 4096, PB0 = 4128, PB1 = 4160, PB2 = 4192, PB3 = 4224, PTOTAL = 4228;

template <int SRC, int C>
__global__ void __launch_bounds__(NT_B, 2) k_density_bwd_tc(const GridParams gp, const nafb_mlp mp, const SamplerParams sp, const uint64_t P,
                                                          const float *__restrict__ dsig_or_dacc, float *__restrict__ grad_table,
                                                          float *__restrict__ partials, const uint8_t *__restrict__ stash, long long *__restrict__ dbg_stamps,
                                                          const int dbg, const nafb_mlp_grads gr, uint32_t *__restrict__ sync) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *A_hi = smem, *A_lo = A_hi + BX_HALF;
    uint8_t *W_hi = A_lo + BX_HALF, *W_lo = W_hi + W_HALF;
    SmallParams *small = reinterpret_cast<SmallParams *>(W_lo + W_HALF);
    TileCtl *ctl = reinterpret_cast<TileCtl *>(reinterpret_cast<uint8_t *>(small) + ((sizeof(SmallParams) + 15) & ~15u));
    float *xchg = reinterpret_cast<float *>(ctl + 1);  // [128] head partial dot products of half 1
    float *xchg2 = xchg + TILE;                         // [128] head pre-activation gradients
    float *wred = xchg2 + TILE;                         // [8 warps][80]: per-warp column sums flushed at the end
    LevelParams *lvs = reinterpret_cast<LevelParams *>(wred + 8 * 80);   // 16-byte aligned (all blocks above are multiples of 16 B)

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int r = t & 127, half = (t >> 7) & 1;
    load_weight_images(mp, W_hi, W_lo, small);
    if (t < NAFB_MAX_LEVELS) lvs[t] = gp.lv[t];
    // the slack / unused blocks are read (as don't-care rows) by the windowed dW MMAs: keep them finite
    for (uint32_t i = t; i < 2 * BX_HALF / 16; i += NT_B) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (t == 0) {
        umma::mbar_init(&ctl->mbar, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&ctl->tmem_base, T_COLS);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = ctl->tmem_base;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t taddr = tmem + lane_base + 16u * half;
    const uint32_t a_hi = umma::smem_u32(A_hi), a_lo = umma::smem_u32(A_lo), w_hi = umma::smem_u32(W_hi), w_lo = umma::smem_u32(W_lo);
    constexpr uint32_t ID_FWD = umma::idesc_bf16(128, 32, 0, 0);   // K-major x K-major
    constexpr uint32_t ID_DX = umma::idesc_bf16(128, 32, 0, 1);    // A K-major (G), B MN-major view of W
    constexpr uint32_t ID_DW32 = umma::idesc_bf16(128, 32, 1, 1);  // both MN-major views, reduction over points
    constexpr uint32_t ID_DW64 = umma::idesc_bf16(128, 64, 1, 1);
    uint32_t phase = 0;
    bool first_tile = true;
    // per-lane accumulators of the SIMT-side gradients: this lane's column of db2/db1/db0 (16 columns of this half),
    // of dW3 (16 columns) and db3
    float acc_db2 = 0.f, acc_db1 = 0.f, acc_db0 = 0.f, acc_dw3 = 0.f, acc_db3 = 0.f;


    const uint64_t n_tiles = (P + TILE - 1) / TILE;
    auto desc = [&](uint32_t base, uint32_t chunk, uint32_t lbo, uint32_t sbo) { return umma::make_desc(base + chunk * LBO, lbo, sbo); };

    if (warp == 8) {
        // ================= MMA warp: one lane issues every tcgen05.mma of the CTA.  The 8 epilogue warps ARRIVE on named
        // barrier 1 when the operands of a phase are in shared memory (and their TMEM reads are done) and go on with other
        // work (the deferred scatter); this warp SYNCs on it, issues the phase and commits to the mbarrier they wait on.
        bool first = true;
        uint32_t it = 0;
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t T_DENC = (it & 1u) ? T_DENC1 : T_DENC0;
#pragma unroll 1
            for (int ph = 0; ph < 6; ++ph) {
                umma::named_bar_sync(1, NT_B);
                if (lane == 0) {
                    umma::fence_after_sync();
                    switch (ph) {
                        case 0:   // forward layer 0: enc . W0^T
                            umma::mma_bf16x3(tmem + T_S, desc(a_hi, CH_ENC, LBO, BX_SBO), desc(a_lo, CH_ENC, LBO, BX_SBO),
                                             umma::make_desc(w_hi + W0_OFF, LBO, W0_SBO), umma::make_desc(w_lo + W0_OFF, LBO, W0_SBO), 256, 256, 2, ID_FWD, false);
                            break;
                        case 1:   // forward layer 1: h0 . W1^T
                            umma::mma_bf16x3(tmem + T_S, desc(a_hi, CH_H0, LBO, BX_SBO), desc(a_lo, CH_H0, LBO, BX_SBO),
                                             umma::make_desc(w_hi + W1_OFF, LBO, W1_SBO), umma::make_desc(w_lo + W1_OFF, LBO, W1_SBO), 256, 256, 2, ID_FWD, false);
                            break;
                        case 2:   // forward layer 2: [enc | h1] (chunks 4..11) . W2^T
                            umma::mma_bf16x3(tmem + T_S, desc(a_hi, CH_ENC, LBO, BX_SBO), desc(a_lo, CH_ENC, LBO, BX_SBO),
                                             umma::make_desc(w_hi + W2_OFF, LBO, W2_SBO), umma::make_desc(w_lo + W2_OFF, LBO, W2_SBO), 256, 256, 4, ID_FWD, false);
                            break;
                        case 3:   // backward layer 2: dW2 += G^T.[enc|h1];  d_enc = G.W2[:, :32];  dh1 = G.W2[:, 32:]
                            umma::mma_bf16x3(tmem + T_DW2, desc(a_hi, CH_G, BX_SBO, LBO), desc(a_lo, CH_G, BX_SBO, LBO), desc(a_hi, CH_ENC, BX_SBO, LBO),
                                             desc(a_lo, CH_ENC, BX_SBO, LBO), 2 * BX_SBO, 2 * BX_SBO, 8, ID_DW64, !first);
                            umma::mma_bf16x3(tmem + T_DENC, desc(a_hi, CH_G, LBO, BX_SBO), desc(a_lo, CH_G, LBO, BX_SBO),
                                             umma::make_desc(w_hi + W2_OFF, W2_SBO, LBO), umma::make_desc(w_lo + W2_OFF, W2_SBO, LBO), 256, 2 * W2_SBO, 2, ID_DX, false);
                            umma::mma_bf16x3(tmem + T_S, desc(a_hi, CH_G, LBO, BX_SBO), desc(a_lo, CH_G, LBO, BX_SBO),
                                             umma::make_desc(w_hi + W2_OFF + 4 * LBO, W2_SBO, LBO), umma::make_desc(w_lo + W2_OFF + 4 * LBO, W2_SBO, LBO), 256,
                                             2 * W2_SBO, 2, ID_DX, false);
                            break;
                        case 4:   // backward layer 1: dW1 += G^T.h0;  dh0 = G.W1
                            umma::mma_bf16x3(tmem + T_DW1, desc(a_hi, CH_G, BX_SBO, LBO), desc(a_lo, CH_G, BX_SBO, LBO), desc(a_hi, CH_H0, BX_SBO, LBO),
                                             desc(a_lo, CH_H0, BX_SBO, LBO), 2 * BX_SBO, 2 * BX_SBO, 8, ID_DW32, !first);
                            umma::mma_bf16x3(tmem + T_S, desc(a_hi, CH_G, LBO, BX_SBO), desc(a_lo, CH_G, LBO, BX_SBO),
                                             umma::make_desc(w_hi + W1_OFF, W1_SBO, LBO), umma::make_desc(w_lo + W1_OFF, W1_SBO, LBO), 256, 2 * W1_SBO, 2, ID_DX, false);
                            break;
                        default:  // backward layer 0: dW0 += G^T.enc;  d_enc += G.W0
                            umma::mma_bf16x3(tmem + T_DW0, desc(a_hi, CH_G, BX_SBO, LBO), desc(a_lo, CH_G, BX_SBO, LBO), desc(a_hi, CH_ENC, BX_SBO, LBO),
                                             desc(a_lo, CH_ENC, BX_SBO, LBO), 2 * BX_SBO, 2 * BX_SBO, 8, ID_DW32, !first);
                            umma::mma_bf16x3(tmem + T_DENC, desc(a_hi, CH_G, LBO, BX_SBO), desc(a_lo, CH_G, LBO, BX_SBO),
                                             umma::make_desc(w_hi + W0_OFF, W0_SBO, LBO), umma::make_desc(w_lo + W0_OFF, W0_SBO, LBO), 256, 2 * W0_SBO, 2, ID_DX, true);
                            break;
                    }
                    umma::commit(&ctl->mbar);
                }
                __syncwarp();
            }
            first = false;
        }
    } else {
    // ================= the 8 epilogue warps
    auto sync_issue = [&]() {   // operands of the next phase are written, TMEM reads of the last one are done: release the MMA warp
        umma::fence_proxy_async();
        umma::fence_before_sync();
        umma::named_bar_arrive(1, NT_B);
    };
    auto wait_mma = [&]() {
        umma::mbar_wait(&ctl->mbar, phase);
        phase ^= 1;
        umma::fence_after_sync();
    };
    // ---- deferred scatter: the gradient of the PREVIOUS tile (still in TMEM) is scattered, one slot at a time, after
    // each MMA issue of the current tile.  Thread (r, half) owns levels l = 2*li + half (even levels on half 0, odd on
    // half 1: the contended coarse levels and the all-miss fine levels are split evenly over the two halves of the CTA).
    constexpr int NLH = 16 / C;   // levels per thread
    float xp[3] = {0.f, 0.f, 0.f};
    bool valid_prev = false, have_prev = false;
    uint32_t tdenc_prev = 0;      // TMEM address (lane base included) of the previous tile's d(encoding)
    const bool do_scatter = grad_table != nullptr && !(dbg & 1);
    const int agg_levels = (dbg & 16) ? 0 : (((dbg >> 8) & 63) ? ((dbg >> 8) & 63) - 1 : AGG_LEVELS);
    const int agg_runs = ((dbg >> 16) & 63) ? ((dbg >> 16) & 63) - 1 : AGG_MAX_RUNS;
    auto scatter_slot = [&](const int slot) {   // 8 slots cover the NLH levels of the thread
        if (!have_prev) return;
        for (int li = slot * NLH / 8; li < (slot + 1) * NLH / 8; ++li) {
            const int l = 2 * li + half;
            scatter_one<C>(lvs, l, xp[0], xp[1], xp[2], tdenc_prev + (uint32_t)(l * C), valid_prev, l < agg_levels ? agg_runs : 0, grad_table);
        }
    };

    // debug (dbg & 32): thread 0 of the first 4 CTAs stamps clock64() at every phase boundary into the 4 KB debug area at the
    // end of the workspace: [cta][128 stamps]
    long long *stamps = (dbg & 32) && t == 0 && blockIdx.x < 4 ? dbg_stamps + blockIdx.x * 128 : nullptr;
    int n_st = 0;
    auto stamp = [&]() { if (stamps && n_st < 128) stamps[n_st++] = clock64(); };

    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t T_DENC = (it & 1u) ? T_DENC1 : T_DENC0;
        stamp();   // 0: tile start
        const uint64_t p = tile * TILE + r;
        const bool valid = p < P;
        float x[3] = {0.f, 0.f, 0.f};
        float dsig = 0.f;
        if (valid) {
            fetch_point<SRC>(sp, p, x);
            if constexpr (SRC == NAFB_SRC_RAYS) {
                const uint32_t ray = (uint32_t)(p / sp.n_samples);
                const uint32_t i = (uint32_t)(p - (uint64_t)ray * sp.n_samples);
                const RayRegs R = load_ray(sp, ray);
                dsig = __fmul_rn(__ldg(dsig_or_dacc + ray), ray_delta(sp, R, ray, i));
            } else {
                dsig = __ldg(dsig_or_dacc + p);
            }
        }
        float x01[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) x01[d] = normalise01(x[d], sp.bound, sp.inv_2bound);
        if (stash) {
            load_stash_half_row(A_hi, A_lo, r, CH_ENC, half, BX_SBO, stash + tile * ST_TILE);
        } else {
            if (dbg & 2) {
                float enc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) enc[i] = x01[i % 3];
                store_half_row(A_hi, A_lo, r, CH_ENC, half, BX_SBO, enc);
            } else {
                gather_half_to_smem<C>(lvs, gp.table, x01[0], x01[1], x01[2], half, A_hi, A_lo, r, CH_ENC, BX_SBO);
            }
        }
        float v[16];
        stamp();   // 1: encodings in shared memory
        // ---------------- forward layer 0
        sync_issue();
        stamp();   // 2: after the barrier
        if (do_scatter) { scatter_slot(0); scatter_slot(1); }
        stamp();
        wait_mma();
        stamp();
        umma::tmem_ld16(taddr + T_S, v);
        umma::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = leaky_relu(v[i] + small->b0[16 * half + i]);
        store_half_row(A_hi, A_lo, r, CH_H0, half, BX_SBO, v);
        // ---------------- forward layer 1
        sync_issue();
        if (do_scatter) { scatter_slot(2); }
        stamp();
        wait_mma();
        stamp();
        umma::tmem_ld16(taddr + T_S, v);
        umma::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = leaky_relu(v[i] + small->b1[16 * half + i]);
        store_half_row(A_hi, A_lo, r, CH_H1, half, BX_SBO, v);
        // ---------------- forward layer 2: [enc | h1] (chunks 4..11) . W2^T
        sync_issue();
        if (do_scatter) { scatter_slot(3); }
        stamp();
        wait_mma();
        stamp();
        umma::tmem_ld16(taddr + T_S, v);
        umma::tmem_wait_ld();
        // ---------------- head forward + backward (fp32 SIMT)
        float h2[16];
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            h2[i] = leaky_relu(v[i] + small->b2[16 * half + i]);
            part = __fmaf_rn(h2[i], small->w3[16 * half + i], part);
        }
        if (half == 1) xchg[r] = part;
        umma::named_bar_sync(2, NT);
        if (half == 0) {
            const float s = (part + xchg[r]) + small->b3;
            const float y = head_activation(s, mp.head);
            xchg2[r] = dsig * head_derivative(s, y, mp.head);
        }
        umma::named_bar_sync(2, NT);
        const float gpre = xchg2[r];
        {
            float gw3[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                gw3[i] = gpre * h2[i];                                                          // dW3 contribution
                v[i] = __fmul_rn(__fmul_rn(small->w3[16 * half + i], gpre), h2[i] > 0.f ? 1.0f : 0.01f);  // dz2
            }
            acc_dw3 += warp_colsum16(gw3, lane);
            acc_db2 += warp_colsum16(v, lane);
            if (half == 0) acc_db3 += warp_sum(gpre);
        }
        stamp();   // head done
        store_half_row(A_hi, A_lo, r, CH_G, half, BX_SBO, v);
        // ---------------- backward layer 2: dW2 += G^T.[enc|h1];  d_enc = G.W2[:, :32];  dh1 = G.W2[:, 32:]
        sync_issue();
        if (do_scatter) { scatter_slot(4); }
        stamp();
        wait_mma();
        stamp();
        umma::tmem_ld16(taddr + T_S, v);
        umma::tmem_wait_ld();
        {
            float s[16];
            lrelu_slopes(A_hi, r, CH_H1, half, BX_SBO, s);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __fmul_rn(v[i], s[i]);  // dz1
            acc_db1 += warp_colsum16(v, lane);
        }
        store_half_row(A_hi, A_lo, r, CH_G, half, BX_SBO, v);
        // ---------------- backward layer 1: dW1 += G^T.h0;  dh0 = G.W1
        sync_issue();
        if (do_scatter) { scatter_slot(5); }
        stamp();
        wait_mma();
        stamp();
        umma::tmem_ld16(taddr + T_S, v);
        umma::tmem_wait_ld();
        {
            float s[16];
            lrelu_slopes(A_hi, r, CH_H0, half, BX_SBO, s);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __fmul_rn(v[i], s[i]);  // dz0
            acc_db0 += warp_colsum16(v, lane);
        }
        store_half_row(A_hi, A_lo, r, CH_G, half, BX_SBO, v);
        // ---------------- backward layer 0: dW0 += G^T.enc;  d_enc += G.W0
        sync_issue();
        if (do_scatter) { scatter_slot(6); scatter_slot(7); }
        stamp();
        wait_mma();
        stamp();
        // d(encoding) of this tile stays in TMEM; it is scattered from the wait slots of the next tile (or below)
        xp[0] = x01[0]; xp[1] = x01[1]; xp[2] = x01[2];
        valid_prev = valid;
        have_prev = true;
        tdenc_prev = tmem + lane_base + T_DENC;
        first_tile = false;
    }
    if (do_scatter) {   // the last tile of this CTA
#pragma unroll 1
        for (int slot = 0; slot < 8; ++slot) scatter_slot(slot);
    }
    // SIMT-side sums: every warp holds partial column sums over its 32 rows
    wred[warp * 80 + 0 * 16 + colsum_index(lane)] = acc_db2;   // lanes l and l^1 hold the same column: benign duplicate store
    wred[warp * 80 + 1 * 16 + colsum_index(lane)] = acc_db1;
    wred[warp * 80 + 2 * 16 + colsum_index(lane)] = acc_db0;
    wred[warp * 80 + 3 * 16 + colsum_index(lane)] = acc_dw3;
    if (lane == 0) wred[warp * 80 + 64] = acc_db3;
    umma::fence_before_sync();
    }   // epilogue warps

    // ================= flush the MLP gradients of this CTA into its slot of the partials workspace
    float *mine = partials + (size_t)blockIdx.x * PTOTAL;
    const bool did_work = blockIdx.x < n_tiles;
    __syncthreads();
    umma::fence_after_sync();
    // combine the per-warp column sums of the 4 warps of each half
    if (t < 32) {   // column j of the 32-wide vectors: half = j / 16 -> warps 4*half .. 4*half+3
        const int hj = t >> 4, cj = t & 15;
        float s2 = 0.f, s1 = 0.f, s0 = 0.f, sw = 0.f;
        for (int w = 0; w < 4; ++w) {
            const float *q = wred + (4 * hj + w) * 80;
            s2 += q[0 * 16 + cj]; s1 += q[1 * 16 + cj]; s0 += q[2 * 16 + cj]; sw += q[3 * 16 + cj];
        }
        mine[PB2 + t] = s2; mine[PB1 + t] = s1; mine[PB0 + t] = s0; mine[PW3 + t] = sw;
        if (t == 0) {
            float s3 = 0.f;
            for (int w = 0; w < 4; ++w) s3 += wred[w * 80 + 64];
            mine[PB3] = s3;
            mine[PB3 + 1] = mine[PB3 + 2] = mine[PB3 + 3] = 0.f;
        }
    }
    // tensor-core side: dW_l[o][k] sits in TMEM lane o (0..31) -> warps 0 and 4 read it (16 columns at a time)
    if ((warp & 3) == 0 && warp < 8) {
        float w16[16];
        auto dump = [&](uint32_t tcol, int dst, int ldw, int ncols) {
            for (int c0 = 16 * half; c0 < ncols; c0 += 32) {
                if (did_work) {
                    umma::tmem_ld16(tmem + tcol + c0, w16);   // lanes 0..31 of quadrant 0
                    umma::tmem_wait_ld();
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) w16[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) mine[dst + lane * ldw + c0 + i] = w16[i];
            }
        };
        dump(T_DW0, PW0, 32, 32);
        dump(T_DW1, PW1, 32, 32);
        dump(T_DW2, PW2, 64, 64);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, T_COLS);

    // ================= gW / gb += sum over the CTAs' rows.  One grid-wide barrier replaces a separate reduction kernel: the grid
    // is at most 2 CTAs per SM (nafb_tc_bwd_grid) and resident at once, so every CTA can wait for all rows and then sum its
    // share of the columns -- units of 16 columns x 18 row slices; a slice adds its rows in order with four loads in flight, the 18
    // slice sums are added in order: the summation tree is a function of the grid size only (deterministic).
    // sync[0] counts arrivals, sync[1] departures; the last CTA to leave clears both (the workspace starts zero-filled).
    if (sync == nullptr) return;   // the caller reduces the rows with a separate launch (debug knob NAFB_BWD_REDUCE=kernel)
    __syncthreads();
    if (t == 0) {
        __threadfence();
        atomicAdd(sync, 1u);
        uint32_t seen, spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(sync) : "memory");
            if (seen < gridDim.x) {
                __nanosleep(64);
                if (++spins > (1u << 24)) __trap();   // seconds: the grid is not resident at once (or the workspace was not zero-filled)
            }
        } while (seen < gridDim.x);
    }
    __syncthreads();
    {
        const int rows = (int)gridDim.x, j = t & 15, k = t >> 4;   // NT_B = 288 -> k in 0..17
        constexpr int SL = NT_B / 16;
        for (int u = blockIdx.x; u * 16 < PTOTAL; u += rows) {
            const int col = u * 16 + j;
            float s = 0.f;
            if (col < PTOTAL) {
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                int b = k;
                for (; b + 3 * SL < rows; b += 4 * SL) {
                    s0 += __ldcg(partials + (size_t)b * PTOTAL + col);
                    s1 += __ldcg(partials + (size_t)(b + SL) * PTOTAL + col);
                    s2 += __ldcg(partials + (size_t)(b + 2 * SL) * PTOTAL + col);
                    s3 += __ldcg(partials + (size_t)(b + 3 * SL) * PTOTAL + col);
                }
                float t0 = 0.f, t1 = 0.f, t2 = 0.f;   // at most three left
                if (b < rows) t0 = __ldcg(partials + (size_t)b * PTOTAL + col);
                if (b + SL < rows) t1 = __ldcg(partials + (size_t)(b + SL) * PTOTAL + col);
                if (b + 2 * SL < rows) t2 = __ldcg(partials + (size_t)(b + 2 * SL) * PTOTAL + col);
                s = ((s0 + s1) + (s2 + s3)) + ((t0 + t1) + t2);
            }
            wred[k * 16 + j] = s;
            __syncthreads();
            if (t < 16 && col < PTOTAL) {
                float tot = 0.f;
#pragma unroll
                for (int q = 0; q < SL; ++q) tot += wred[q * 16 + t];
                float *dst = nullptr;
                if (col < PW1) dst = gr.gW[0] ? gr.gW[0] + col : nullptr;
                else if (col < PW2) dst = gr.gW[1] ? gr.gW[1] + (col - PW1) : nullptr;
                else if (col < PW3) dst = gr.gW[2] ? gr.gW[2] + (col - PW2) : nullptr;
                else if (col < PB0) dst = gr.gW[3] ? gr.gW[3] + (col - PW3) : nullptr;
                else if (col < PB1) dst = gr.gb[0] ? gr.gb[0] + (col - PB0) : nullptr;
                else if (col < PB2) dst = gr.gb[1] ? gr.gb[1] + (col - PB1) : nullptr;
                else if (col < PB3) dst = gr.gb[2] ? gr.gb[2] + (col - PB2) : nullptr;
                else if (col == PB3) dst = gr.gb[3];
                if (dst) *dst += tot;
            }
            __syncthreads();
        }
    }
    if (t == 0 && atomicAdd(sync + 1, 1u) == gridDim.x - 1) { sync[0] = 0u; sync[1] = 0u; }
}

bool tc_config_ok(const nafb_grid *grid, const nafb_mlp *mlp) {
    return grid->D == 3 && grid->L * grid->C == 32 && mlp->in_dim == 32 && mlp->hidden == 32 && mlp->out_dim == 1 && mlp->n_layers == 4 &&
           mlp->skip_mask == (1u << 2);
}

template <int SRC, int C>
int launch_fwd_tc(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, uint64_t P, float *sigma, float *acc, float *z, float *pts,
                  int32_t *flags, uint8_t *stash, const nafb_loss_tail &tail, cudaStream_t s) {
    static bool configured[NAFB_MAX_DEVICES] = {};
    NAFB_CONFIGURE_SMEM(configured, (k_density_fwd_tc<SRC, C>), (int)FWD_SMEM, "density_forward(tc)");
    const uint64_t n_tiles = SRC == NAFB_SRC_VOXELS ? (uint64_t)((sp.i1 - sp.i0 + 3) / 4) * ((sp.n2 + 3) / 4) * ((sp.n3 + 7) / 8) : (P + TILE - 1) / TILE;
    const uint64_t cap = (uint64_t)nafb_sm_count() * FWD_CTAS;
    const unsigned grid = (unsigned)(n_tiles < cap ? n_tiles : cap);
    const int dbg = nafb_debug_flags();
    k_density_fwd_tc<SRC, C><<<grid, NT, FWD_SMEM, s>>>(gp, mp, sp, P, sigma, acc, z, pts, flags, stash, dbg, tail);
    NAFB_CHECK_LAUNCH("density_forward(tc)");
    return NAFB_OK;
}

}  // namespace

// ---- entry points used by density.cu's dispatcher
bool nafb_tc_config_ok(const nafb_grid *grid, const nafb_mlp *mlp) { return tc_config_ok(grid, mlp); }

int nafb_tc_bwd_grid(uint64_t n_tiles) {
    const uint64_t cap = (uint64_t)nafb_sm_count() * 2;
    return (int)(n_tiles < cap ? n_tiles : cap);
}

uint64_t nafb_tc_stash_bytes(uint64_t n_points) { return (n_points + TILE - 1) / TILE * (uint64_t)ST_TILE; }

int nafb_launch_fwd_tc(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, int src, uint64_t P, float *sigma, float *acc, float *z,
                       float *pts, int32_t *flags, void *stash, const nafb_loss_tail *tail_in, cudaStream_t s) {
    nafb_loss_tail tail = {};
    if (tail_in) tail = *tail_in;
#define CALL(S_, C_) launch_fwd_tc<S_, C_>(gp, mp, sp, P, sigma, acc, z, pts, flags, (uint8_t *)stash, tail, s)
    switch (gp.C) {
        case 1: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 1) : src == NAFB_SRC_RAYS ? CALL(NAFB_SRC_RAYS, 1) : CALL(NAFB_SRC_VOXELS, 1);
        case 2: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 2) : src == NAFB_SRC_RAYS ? CALL(NAFB_SRC_RAYS, 2) : CALL(NAFB_SRC_VOXELS, 2);
        case 4: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 4) : src == NAFB_SRC_RAYS ? CALL(NAFB_SRC_RAYS, 4) : CALL(NAFB_SRC_VOXELS, 4);
        default: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 8) : src == NAFB_SRC_RAYS ? CALL(NAFB_SRC_RAYS, 8) : CALL(NAFB_SRC_VOXELS, 8);
    }
#undef CALL
}

template <int SRC, int C>
static int launch_bwd_tc_t(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, uint64_t P, const float *dsig, float *grad_table,
                           float *partials, const uint8_t *stash, long long *stamps, int grid, const nafb_mlp_grads &gr, cudaStream_t s) {
    static bool configured[NAFB_MAX_DEVICES] = {};
    NAFB_CONFIGURE_SMEM(configured, (k_density_bwd_tc<SRC, C>), (int)BWD_SMEM, "density_backward(tc)");
    // behind the partials: 4096 B of phase time stamps (debug), then the two words of the grid barrier
    uint32_t *sync = gr.gW[0] || gr.gb[0] || gr.gW[1] ? reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(stamps) + 4096) : nullptr;
    k_density_bwd_tc<SRC, C><<<grid, NT_B, BWD_SMEM, s>>>(gp, mp, sp, P, dsig, grad_table, partials, stash, stamps, nafb_debug_flags(), gr, sync);
    NAFB_CHECK_LAUNCH("density_backward(tc)");
    return NAFB_OK;
}

int nafb_launch_bwd_tc(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, int src, uint64_t P, const float *dsig,
                       float *grad_table, float *partials, const void *stash, long long *stamps, int grid, const nafb_mlp_grads &gr, cudaStream_t s) {
#define CALL(S_, C_) launch_bwd_tc_t<S_, C_>(gp, mp, sp, P, dsig, grad_table, partials, (const uint8_t *)stash, stamps, grid, gr, s)
    switch (gp.C) {
        case 1: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 1) : CALL(NAFB_SRC_RAYS, 1);
        case 2: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 2) : CALL(NAFB_SRC_RAYS, 2);
        case 4: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 4) : CALL(NAFB_SRC_RAYS, 4);
        default: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 8) : CALL(NAFB_SRC_RAYS, 8);
    }
#undef CALL
}
