// density_bwd_tc.cu -- warp-specialised tcgen05 backward pass of the fused "hash-grid encoder + density MLP" kernel
// (reference: autograd through src/network/network.py:34-58 + hashencoder.cu:201-272 kernel_grid_backward), for the
// configuration every shipped YAML uses (L*C == 32, 4 x 32 MLP, skip at 2, out_dim 1).
//
// One CTA = 17 warps (56 registers per thread: two CTAs per SM) working on 128-point tiles (persistent, static tile schedule), three roles:
//
//   warps 0-7   EPILOGUE  thread (r, half): TMEM lane r = sample point, 16 of the 32 feature columns.  TMEM -> registers ->
//                         bias / LeakyReLU / head / activation gradients (fp32 SIMT) -> bf16 (hi, lo) operand of the next MMA.
//   warp  8     ISSUE     one lane issues every tcgen05.mma of the CTA and the bulk copy (cp.async.bulk, TMA unit) that
//                         prefetches the NEXT tile's encoding stash (16 KB, already in operand layout) into the other
//                         half of a double buffer while the current tile computes.
//   warps 9-16  SCATTER   two per TMEM lane quadrant (even / odd levels): d(encoding) of the tile that has just finished stays in TMEM (double
//                         buffered) and is scattered into the table gradient -- warp-aggregated on the coarse levels,
//                         x-neighbour pairs merged into one red.v4 (density_tc.cuh) -- while the other warps run the chain of
//                         the next tile.  The reductions drain through the LSU without ever stalling the chain.
//
// Per tile the tensor core runs (bf16x3 split products, fp32 accumulation in TMEM):
//   forward   h0 = enc.W0^T          h1 = h0.W1^T          h2 = [enc|h1].W2^T   (K split in two, no operand contiguity needed)
//   dX        d_enc  = G2.W2[:, :32]   dh1 = G2.W2[:, 32:]   dh0 = G1.W1   d_enc += G0.W0
//   dW        dW2 += G2^T.[enc|h1]    dW1 += G1^T.h0        dW0 += G0^T.enc     (reduction over the 128 points)
// The dW products are OFF the critical path: they are issued after the dX products of their phase and signal a second
// mbarrier that the epilogue only consults before it overwrites G.  Their A operand is the transposed view of the G tile with
// the hi and lo images STACKED along M ([G_hi | G_lo] are adjacent feature blocks of one buffer): one MMA yields G_hi^T.B in
// TMEM lanes 0-31 and G_lo^T.B in lanes 32-63, so a k-step costs two MMAs (B_hi, B_lo) instead of three; the two lane blocks
// are added once, when the CTA dumps its weight gradients.
//
// Hand-offs: named barriers 1 / 3 (alternating by phase) epilogue -> issue warp; mbarriers (tcgen05.commit) issue -> epilogue
// (mma), issue -> epilogue (dw), issue -> scatter (denc_full), scatter -> issue (denc_empty), TMA -> issue (stash_full),
// issue -> TMA / epilogue gather (enc_free).
#include <stdio.h>
#include <stdlib.h>

#include "density_tc.cuh"

using namespace tc;

namespace {

// Two CTAs must share the SM's 64 K registers (the chain of one CTA hides the latencies of the other): NSW = 4 scatter warps
// -> 13 warps x 64 registers; NSW = 8 -> 17 warps x 56 registers.  The launch bound below is what makes ptxas stay inside
// that budget.
constexpr int N_EPI = 256;
constexpr int WARP_ISSUE = 8, WARP_SCATTER0 = 9;
__host__ __device__ constexpr int block_threads(int nsw) { return 32 * (9 + nsw); }
__host__ __device__ constexpr int bound_threads(int nsw) { return nsw == 4 ? 512 : 544; }
constexpr uint32_t BAR_THREADS = N_EPI + 32;   // 256 arrive + the issue warp syncs

// ---- shared memory
// ENC[2]: the stash tile byte for byte: hi image (16 row groups x 4 chunks, SBO 512), then the lo image
constexpr uint32_t ENC_SBO = ST_SBO, ENC_HALF = ST_HALF, ENC_BYTES = ST_ENC;
// ACT_hi: 16 chunks per row group: H0 (0-3) | H1 (4-7) | G_hi (8-11) | G_lo (12-15); + 1 KB so that the 128-feature window
// that starts at G_hi (the stacked A operand of the dW products) stays inside the allocation for the last row group
constexpr uint32_t AH_SBO = 2048, AH_BYTES = 16 * AH_SBO + 1024;
// ACT_lo: 8 chunks per row group: H0 (0-3) | H1 (4-7)
constexpr uint32_t AL_SBO = 1024, AL_BYTES = 16 * AL_SBO;
constexpr uint32_t CH_H0 = 0, CH_H1 = 4, CH_GH = 8, CH_GL = 12;

struct alignas(16) Ctl {
    uint64_t mma, dw;
    uint64_t denc_full[2], denc_empty[2];
    uint64_t stash_full[2], enc_free[2];
    uint32_t tmem_base, pad;
};

constexpr uint32_t WRED_FLOATS = 8 * 80;   // >= (NTW / 16) * 16 = 544 for the final reduction
constexpr uint32_t BWS_SMEM = 2 * ENC_BYTES + AH_BYTES + AL_BYTES + 2 * W_HALF + ((sizeof(SmallParams) + 15) & ~15u) + sizeof(Ctl) +
                              2 * TILE * sizeof(float) + WRED_FLOATS * sizeof(float) + NAFB_MAX_LEVELS * sizeof(LevelParams) + 128;

// TMEM columns (fp32): scratch accumulator, d(encoding) double buffer, weight gradients
constexpr uint32_t T_S = 0, T_DENC0 = 32, T_DW0 = 64, T_DW1 = 96, T_DW2 = 128, T_DENC1 = 192, T_COLS = 256;
// offsets inside one CTA's slot of the partials workspace (floats) -- matches density.cu's MlpLayout for this net
constexpr int PW0 = 0, PW1 = 1024, PW2 = 2048, PW3 = 4096, PB0 = 4128, PB1 = 4160, PB2 = 4192, PB3 = 4224, PTOTAL = 4228;

__device__ __forceinline__ void mbar_arrive(uint64_t *mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(mbar)) : "memory");
}
// one bulk copy global -> shared through the TMA unit; completion (bytes) is signalled on `mbar`
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(mbar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(umma::smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(umma::smem_u32(mbar))
                 : "memory");
}

// D (+)= [A_hi | A_lo]^T-stacked x (B_hi + B_lo): two MMAs per k-step (small term first); the hi and lo images of B may live
// in buffers with different row-group strides
__device__ __forceinline__ void mma_stacked(uint32_t d_tmem, uint64_t a, uint64_t b_hi, uint64_t b_lo, uint32_t a_step, uint32_t b_step_hi,
                                            uint32_t b_step_lo, int ksteps, uint32_t idesc, bool accumulate) {
#pragma unroll
    for (int k = 0; k < ksteps; ++k) {
        const uint64_t ak = umma::advance_desc(a, k * a_step);
        umma::mma_bf16(d_tmem, ak, umma::advance_desc(b_lo, k * b_step_lo), idesc, accumulate || k > 0);
        umma::mma_bf16(d_tmem, ak, umma::advance_desc(b_hi, k * b_step_hi), idesc, true);
    }
}

// this thread's 16 values (two chunks) of a 32-wide block: hi image at (hi, chunk_hi, sbo_hi), lo image at (lo, chunk_lo, sbo_lo)
__device__ __forceinline__ void store_half_row2(uint8_t *hi, uint32_t chunk_hi, uint32_t sbo_hi, uint8_t *lo, uint32_t chunk_lo, uint32_t sbo_lo,
                                                uint32_t row, int half, const float (&v)[16]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint4 h, l;
        umma::split_chunk(v + 8 * c, h, l);
        *reinterpret_cast<uint4 *>(hi + umma::canon_off(row, chunk_hi + 2 * half + c, LBO, sbo_hi)) = h;
        *reinterpret_cast<uint4 *>(lo + umma::canon_off(row, chunk_lo + 2 * half + c, LBO, sbo_lo)) = l;
    }
}

template <int SRC, int C, int NSW>
__global__ void __launch_bounds__(bound_threads(NSW), 2) k_density_bwd_ws(const GridParams gp, const nafb_mlp mp, const SamplerParams sp, const uint64_t P,
                                                         const float *__restrict__ dsig_or_dacc, float *__restrict__ grad_table,
                                                         float *__restrict__ partials, const uint8_t *__restrict__ stash,
                                                         long long *__restrict__ dbg_stamps, const int dbg, const nafb_mlp_grads gr,
                                                         uint32_t *__restrict__ sync) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *ENC = smem;                              // [2][ENC_BYTES]
    uint8_t *A_hi = ENC + 2 * ENC_BYTES, *A_lo = A_hi + AH_BYTES;
    uint8_t *W_hi = A_lo + AL_BYTES, *W_lo = W_hi + W_HALF;
    SmallParams *small = reinterpret_cast<SmallParams *>(W_lo + W_HALF);
    Ctl *ctl = reinterpret_cast<Ctl *>(reinterpret_cast<uint8_t *>(small) + ((sizeof(SmallParams) + 15) & ~15u));
    LevelParams *lvs = reinterpret_cast<LevelParams *>(ctl + 1);   // 16-byte aligned: sizeof(Ctl) is a multiple of 16
    float *xchg = reinterpret_cast<float *>(lvs + NAFB_MAX_LEVELS);  // [128] head partial dot products of half 1
    float *xchg2 = xchg + TILE;                         // [128] head pre-activation gradients
    float *wred = xchg2 + TILE;                         // per-warp column sums / slices of the final reduction

    constexpr int NTW = block_threads(NSW);
    constexpr uint32_t N_SCATTER_WARPS = NSW;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    // debug (dbg & 32): thread 0 of CTA 0 records clock64 / %globaltimer at the kernel's milestones in slot 3, entries 112..
    long long *kst = (dbg & 32) && t == 0 && blockIdx.x == 0 ? dbg_stamps + 3 * 128 + 96 : nullptr;
    auto kstamp = [&](int i) {
        if (kst) {
            unsigned long long g, c;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)::"memory");
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)::"memory");   // (memory clobber: not to be scheduled across a barrier)
            kst[2 * i] = (long long)c;
            kst[2 * i + 1] = (long long)g;
        }
    };
    kstamp(0);
    load_weight_images(mp, W_hi, W_lo, small);
    if (t < NAFB_MAX_LEVELS) lvs[t] = gp.lv[t];
    // The stacked 128-feature window of the dW products (A = [G_hi | G_lo | 64 more rows]) reads, beyond G, the H0 / H1 chunks of
    // the NEXT row group -- written by the epilogue before the first dW product -- and, for the last row group, the 1 KB of slack
    // behind the buffer: that slack is the only operand memory no role ever writes, so it is what gets cleared (the rows it feeds,
    // TMEM lanes 64-127 of the dW accumulators, are never read back; they only have to stay finite).
    for (uint32_t i = t; i < 1024 / 16; i += NTW) reinterpret_cast<uint4 *>(A_hi + 16 * AH_SBO)[i] = make_uint4(0, 0, 0, 0);
    if (t == 0) {
        umma::mbar_init(&ctl->mma, 1);
        umma::mbar_init(&ctl->dw, 1);
        for (int b = 0; b < 2; ++b) {
            umma::mbar_init(&ctl->denc_full[b], 1);
            umma::mbar_init(&ctl->denc_empty[b], N_SCATTER_WARPS);
            umma::mbar_init(&ctl->stash_full[b], 1);
            umma::mbar_init(&ctl->enc_free[b], 1);
        }
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&ctl->tmem_base, T_COLS);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = ctl->tmem_base;
    kstamp(1);   // set-up done
    const uint64_t n_tiles = (P + TILE - 1) / TILE;
    const bool did_work = blockIdx.x < n_tiles;
    const bool do_scatter = grad_table != nullptr && !(dbg & 1);

    if (warp == WARP_ISSUE) {
        // ================================================================================ ISSUE warp
        const uint32_t enc0 = umma::smem_u32(ENC), a_hi = umma::smem_u32(A_hi), a_lo = umma::smem_u32(A_lo);
        const uint32_t w_hi = umma::smem_u32(W_hi), w_lo = umma::smem_u32(W_lo);
        constexpr uint32_t ID_FWD = umma::idesc_bf16(128, 32, 0, 0);   // K-major x K-major
        constexpr uint32_t ID_DX = umma::idesc_bf16(128, 32, 0, 1);    // A K-major (G), B MN-major view of W
        constexpr uint32_t ID_DW = umma::idesc_bf16(128, 32, 1, 1);    // both MN-major views, reduction over the points
        // K-major A operands
        const uint64_t dH0_hi = umma::make_desc(a_hi + CH_H0 * LBO, LBO, AH_SBO), dH0_lo = umma::make_desc(a_lo + CH_H0 * LBO, LBO, AL_SBO);
        const uint64_t dH1_hi = umma::make_desc(a_hi + CH_H1 * LBO, LBO, AH_SBO), dH1_lo = umma::make_desc(a_lo + CH_H1 * LBO, LBO, AL_SBO);
        const uint64_t dG_hi = umma::make_desc(a_hi + CH_GH * LBO, LBO, AH_SBO), dG_lo = umma::make_desc(a_hi + CH_GL * LBO, LBO, AH_SBO);
        // MN-major views (LBO / SBO swapped): the stacked [G_hi | G_lo | ...] window as A, the activations as B
        const uint64_t tG = umma::make_desc(a_hi + CH_GH * LBO, AH_SBO, LBO);
        const uint64_t tH0_hi = umma::make_desc(a_hi + CH_H0 * LBO, AH_SBO, LBO), tH0_lo = umma::make_desc(a_lo + CH_H0 * LBO, AL_SBO, LBO);
        const uint64_t tH1_hi = umma::make_desc(a_hi + CH_H1 * LBO, AH_SBO, LBO), tH1_lo = umma::make_desc(a_lo + CH_H1 * LBO, AL_SBO, LBO);
        // weights: K-major for the forward products, MN-major views for dX
        const uint64_t W0_hi = umma::make_desc(w_hi + W0_OFF, LBO, W0_SBO), W0_lo = umma::make_desc(w_lo + W0_OFF, LBO, W0_SBO);
        const uint64_t W1_hi = umma::make_desc(w_hi + W1_OFF, LBO, W1_SBO), W1_lo = umma::make_desc(w_lo + W1_OFF, LBO, W1_SBO);
        const uint64_t W2a_hi = umma::make_desc(w_hi + W2_OFF, LBO, W2_SBO), W2a_lo = umma::make_desc(w_lo + W2_OFF, LBO, W2_SBO);
        const uint64_t W2b_hi = umma::make_desc(w_hi + W2_OFF + 4 * LBO, LBO, W2_SBO), W2b_lo = umma::make_desc(w_lo + W2_OFF + 4 * LBO, LBO, W2_SBO);
        const uint64_t tW0_hi = umma::make_desc(w_hi + W0_OFF, W0_SBO, LBO), tW0_lo = umma::make_desc(w_lo + W0_OFF, W0_SBO, LBO);
        const uint64_t tW1_hi = umma::make_desc(w_hi + W1_OFF, W1_SBO, LBO), tW1_lo = umma::make_desc(w_lo + W1_OFF, W1_SBO, LBO);
        const uint64_t tW2a_hi = umma::make_desc(w_hi + W2_OFF, W2_SBO, LBO), tW2a_lo = umma::make_desc(w_lo + W2_OFF, W2_SBO, LBO);
        const uint64_t tW2b_hi = umma::make_desc(w_hi + W2_OFF + 4 * LBO, W2_SBO, LBO), tW2b_lo = umma::make_desc(w_lo + W2_OFF + 4 * LBO, W2_SBO, LBO);

        if (lane == 0 && did_work && stash) bulk_load(ENC, stash + (uint64_t)blockIdx.x * ST_TILE, ENC_BYTES, &ctl->stash_full[0]);
        uint32_t it = 0;
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t b = it & 1u;
            const uint32_t enc_hi = enc0 + b * ENC_BYTES, enc_lo = enc_hi + ENC_HALF;
            const uint64_t dE_hi = umma::make_desc(enc_hi, LBO, ENC_SBO), dE_lo = umma::make_desc(enc_lo, LBO, ENC_SBO);
            const uint64_t tE_hi = umma::make_desc(enc_hi, ENC_SBO, LBO), tE_lo = umma::make_desc(enc_lo, ENC_SBO, LBO);
            const uint32_t T_DENC = b ? T_DENC1 : T_DENC0;
            const bool first = it == 0;
            // ---------------- phase 0: h0 = enc . W0^T
            umma::named_bar_sync_imm<1>(BAR_THREADS);
            if (lane == 0) {
                if (stash) umma::mbar_wait(&ctl->stash_full[b], (it >> 1) & 1u);
                umma::fence_after_sync();
                umma::mma_bf16x3(tmem + T_S, dE_hi, dE_lo, W0_hi, W0_lo, 256, 256, 2, ID_FWD, false);
                umma::commit(&ctl->mma);
                // prefetch the next tile's stash into the other buffer (free once the dW0 products of the previous tile are done)
                const uint64_t next = tile + gridDim.x;
                if (stash && next < n_tiles) {
                    if (it >= 1) umma::mbar_wait(&ctl->enc_free[b ^ 1u], ((it - 1) >> 1) & 1u);
                    bulk_load(ENC + (b ^ 1u) * ENC_BYTES, stash + next * ST_TILE, ENC_BYTES, &ctl->stash_full[b ^ 1u]);
                }
            }
            __syncwarp();
            // ---------------- phase 1: h1 = h0 . W1^T
            umma::named_bar_sync_imm<3>(BAR_THREADS);
            if (lane == 0) {
                umma::fence_after_sync();
                umma::mma_bf16x3(tmem + T_S, dH0_hi, dH0_lo, W1_hi, W1_lo, 256, 256, 2, ID_FWD, false);
                umma::commit(&ctl->mma);
            }
            __syncwarp();
            // ---------------- phase 2: h2 = enc . W2[:, :32]^T + h1 . W2[:, 32:]^T
            umma::named_bar_sync_imm<1>(BAR_THREADS);
            if (lane == 0) {
                umma::fence_after_sync();
                umma::mma_bf16x3(tmem + T_S, dE_hi, dE_lo, W2a_hi, W2a_lo, 256, 256, 2, ID_FWD, false);
                umma::mma_bf16x3(tmem + T_S, dH1_hi, dH1_lo, W2b_hi, W2b_lo, 256, 256, 2, ID_FWD, true);
                umma::commit(&ctl->mma);
            }
            __syncwarp();
            // ---------------- phase 3: d_enc = G.W2[:, :32]; dh1 = G.W2[:, 32:]   |   dW2 += G^T.[enc | h1]
            umma::named_bar_sync_imm<3>(BAR_THREADS);
            if (lane == 0) {
                if (it >= 2) umma::mbar_wait(&ctl->denc_empty[b], ((it - 2) >> 1) & 1u);   // the scatter warps have drained this buffer
                umma::fence_after_sync();
                umma::mma_bf16x3(tmem + T_S, dG_hi, dG_lo, tW2b_hi, tW2b_lo, 256, 2 * W2_SBO, 2, ID_DX, false);
                umma::mma_bf16x3(tmem + T_DENC, dG_hi, dG_lo, tW2a_hi, tW2a_lo, 256, 2 * W2_SBO, 2, ID_DX, false);
                umma::commit(&ctl->mma);
                mma_stacked(tmem + T_DW2, tG, tE_hi, tE_lo, 2 * AH_SBO, 2 * ENC_SBO, 2 * ENC_SBO, 8, ID_DW, !first);
                mma_stacked(tmem + T_DW2 + 32, tG, tH1_hi, tH1_lo, 2 * AH_SBO, 2 * AH_SBO, 2 * AL_SBO, 8, ID_DW, !first);
                umma::commit(&ctl->dw);
            }
            __syncwarp();
            // ---------------- phase 4: dh0 = G.W1   |   dW1 += G^T.h0
            umma::named_bar_sync_imm<1>(BAR_THREADS);
            if (lane == 0) {
                umma::fence_after_sync();
                umma::mma_bf16x3(tmem + T_S, dG_hi, dG_lo, tW1_hi, tW1_lo, 256, 2 * W1_SBO, 2, ID_DX, false);
                umma::commit(&ctl->mma);
                mma_stacked(tmem + T_DW1, tG, tH0_hi, tH0_lo, 2 * AH_SBO, 2 * AH_SBO, 2 * AL_SBO, 8, ID_DW, !first);
                umma::commit(&ctl->dw);
            }
            __syncwarp();
            // ---------------- phase 5: d_enc += G.W0 (nobody but the scatter warps waits for it)   |   dW0 += G^T.enc
            umma::named_bar_sync_imm<3>(BAR_THREADS);
            if (lane == 0) {
                umma::fence_after_sync();
                umma::mma_bf16x3(tmem + T_DENC, dG_hi, dG_lo, tW0_hi, tW0_lo, 256, 2 * W0_SBO, 2, ID_DX, true);
                umma::commit(&ctl->denc_full[b]);
                mma_stacked(tmem + T_DW0, tG, tE_hi, tE_lo, 2 * AH_SBO, 2 * ENC_SBO, 2 * ENC_SBO, 8, ID_DW, !first);
                umma::commit(&ctl->dw);
                umma::commit(&ctl->enc_free[b]);
            }
            __syncwarp();
        }
    } else if (warp >= WARP_SCATTER0) {
        // ================================================================================ SCATTER warps
        const uint32_t q = (uint32_t)warp & 3u;              // TMEM lane quadrant this warp may read
        constexpr int SETS = NSW / 4;                                 // scatter warps per TMEM lane quadrant
        const uint32_t set = (uint32_t)(warp - WARP_SCATTER0) >> 2;   // this warp takes the levels l with l % SETS == set (coarse and fine levels split evenly)
        const uint32_t r = q * 32u + (uint32_t)lane;         // row of the tile == TMEM lane
        const int agg_levels = (dbg & 16) ? 0 : (((dbg >> 8) & 63) ? ((dbg >> 8) & 63) - 1 : AGG_LEVELS);
        const int agg_runs = ((dbg >> 16) & 63) ? ((dbg >> 16) & 63) - 1 : AGG_MAX_RUNS;
        constexpr int NLH = (32 / C) / SETS;   // levels per scatter warp
        long long *stamps = (dbg & 32) && lane == 0 && warp == WARP_SCATTER0 && blockIdx.x < 2 ? dbg_stamps + (2 + blockIdx.x) * 128 : nullptr;
        int n_st = 0;
        uint32_t it = 0;
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t b = it & 1u;
            const uint64_t p = tile * TILE + r;
            const bool valid = p < P;
            float x01[3] = {0.f, 0.f, 0.f};
            if (stash) {   // the forward pass left the normalised positions in the tail of the stash tile
                const float *tail = reinterpret_cast<const float *>(stash + tile * ST_TILE + ST_TAIL_X01);
                x01[0] = __ldcs(tail + r); x01[1] = __ldcs(tail + 128 + r); x01[2] = __ldcs(tail + 256 + r);
            } else {
                float x[3] = {0.f, 0.f, 0.f};
                if (valid) fetch_point<SRC>(sp, p, x);
#pragma unroll
                for (int d = 0; d < 3; ++d) x01[d] = normalise01(x[d], sp.bound, sp.inv_2bound);
            }
            if (stamps && n_st < 127) stamps[n_st++] = clock64();
            umma::mbar_wait_backoff(&ctl->denc_full[b], (it >> 1) & 1u, 32);
            umma::fence_after_sync();
            if (stamps && n_st < 127) stamps[n_st++] = clock64();
            const uint32_t tdenc = tmem + ((q * 32u) << 16) + (b ? T_DENC1 : T_DENC0);
            if (do_scatter) {
#pragma unroll 1
                for (int li = 0; li < NLH; ++li) {
                    const int l = SETS * li + (int)set;
                    scatter_one<C>(lvs, l, x01[0], x01[1], x01[2], tdenc + (uint32_t)(l * C), valid, l < agg_levels ? agg_runs : 0, grad_table);
                }
            }
            umma::tmem_wait_ld();
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->denc_empty[b]);
        }
        if (stamps && n_st < 128) stamps[n_st++] = clock64();
    } else {
        // ================================================================================ EPILOGUE warps
        const int r = t & 127, half = t >> 7;
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 16u * half;
        uint32_t n_mma = 0, n_dw = 0, n_bar = 0;   // completed waits / arrivals (parities)
        auto sync_issue = [&]() {   // operands of the next phase are written, TMEM reads of the last one are done: release the issue warp
            umma::fence_proxy_async();
            umma::fence_before_sync();
            if (n_bar & 1u) umma::named_bar_arrive_imm<3>(BAR_THREADS); else umma::named_bar_arrive_imm<1>(BAR_THREADS);
            ++n_bar;
        };
        auto wait_mma = [&]() {
            umma::mbar_wait(&ctl->mma, n_mma & 1u);
            ++n_mma;
            umma::fence_after_sync();
        };
        auto wait_dw = [&]() {   // the dW products that read the current G tile (and the activations of this tile) have completed
            umma::mbar_wait(&ctl->dw, n_dw & 1u);
            ++n_dw;
        };
        // per-lane accumulators of the SIMT-side gradients: this lane's column of db2/db1/db0 (16 columns of this half),
        // of dW3 (16 columns) and db3
        float acc_db2 = 0.f, acc_db1 = 0.f, acc_db0 = 0.f, acc_dw3 = 0.f, acc_db3 = 0.f;
        long long *stamps = (dbg & 32) && t == 0 && blockIdx.x < 2 ? dbg_stamps + blockIdx.x * 128 : nullptr;
        int n_st = 0;
        auto stamp = [&]() { if (stamps && n_st < 128) stamps[n_st++] = clock64(); };

        uint32_t it = 0;
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t b = it & 1u;
            stamp();   // 0: tile start
            const uint64_t p = tile * TILE + r;
            const bool valid = p < P;
            float dsig = 0.f;
            float x[3] = {0.f, 0.f, 0.f};
            if (valid) {
                if (!stash) fetch_point<SRC>(sp, p, x);
                if constexpr (SRC == NAFB_SRC_RAYS) {
                    // dsigma = dacc[ray] * delta_i |d| (render.py:192-201); the forward pass left delta in the tail of the stash tile
                    const uint32_t ray = P <= 0xffffffffull ? (uint32_t)p / sp.n_samples : (uint32_t)(p / sp.n_samples);
                    float delta;
                    if (stash) {
                        delta = __ldcs(reinterpret_cast<const float *>(stash + tile * ST_TILE + ST_TAIL_DELTA) + r);
                    } else {
                        const uint32_t i = (uint32_t)(p - (uint64_t)ray * sp.n_samples);
                        delta = ray_delta(sp, load_ray(sp, ray), ray, i);
                    }
                    dsig = __fmul_rn(__ldg(dsig_or_dacc + ray), delta);
                } else {
                    dsig = __ldg(dsig_or_dacc + p);
                }
            }
            if (!stash) {   // no stash from the forward pass: gather the encodings again (into the buffer the TMA would have filled)
                if (it >= 2) umma::mbar_wait(&ctl->enc_free[b], ((it - 2) >> 1) & 1u);
                uint8_t *e_hi = ENC + b * ENC_BYTES, *e_lo = e_hi + ENC_HALF;
                float x01[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) x01[d] = normalise01(x[d], sp.bound, sp.inv_2bound);
                if (dbg & 2) {
                    float enc[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) enc[i] = x01[i % 3];
                    store_half_row(e_hi, e_lo, r, 0, half, ENC_SBO, enc);
                } else {
                    gather_half_to_smem<C>(lvs, gp.table, x01[0], x01[1], x01[2], half, e_hi, e_lo, r, 0, ENC_SBO);
                }
            }
            float v[16];
            // ---------------- forward layer 0
            sync_issue();
            stamp();   // 1
            wait_mma();
            stamp();   // 2
            umma::tmem_ld16(taddr + T_S, v);
            umma::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = leaky_relu(v[i] + small->b0[16 * half + i]);
            store_half_row2(A_hi, CH_H0, AH_SBO, A_lo, CH_H0, AL_SBO, r, half, v);
            // ---------------- forward layer 1
            sync_issue();
            stamp();   // 3
            wait_mma();
            stamp();   // 4
            umma::tmem_ld16(taddr + T_S, v);
            umma::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = leaky_relu(v[i] + small->b1[16 * half + i]);
            store_half_row2(A_hi, CH_H1, AH_SBO, A_lo, CH_H1, AL_SBO, r, half, v);
            // ---------------- forward layer 2
            sync_issue();
            stamp();   // 5
            wait_mma();
            stamp();   // 6
            umma::tmem_ld16(taddr + T_S, v);
            umma::tmem_wait_ld();
            // ---------------- head forward + backward (fp32 SIMT)
            {
                float h2[16];
                float part = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    h2[i] = leaky_relu(v[i] + small->b2[16 * half + i]);
                    part = __fmaf_rn(h2[i], small->w3[16 * half + i], part);
                }
                // the two halves of a row live in warps w and w + 4: each pair of warps has its own 64-thread barrier
                // (hardware barriers are a per-SM resource shared by the resident CTAs: this kernel uses ids 0-3 only -- with 8 of
                // them only one CTA per SM was resident)
                auto pair_sync = [&]() { umma::named_bar_sync_imm<2>(N_EPI); };
                if (half == 1) xchg[r] = part;
                pair_sync();
                if (half == 0) {
                    const float s = (part + xchg[r]) + small->b3;
                    const float y = head_activation(s, mp.head);
                    xchg2[r] = dsig * head_derivative(s, y, mp.head);
                }
                pair_sync();
                const float gpre = xchg2[r];
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
            stamp();   // 7: head done
            if (it > 0) wait_dw();   // dW0 of the previous tile has read the old G
            store_half_row2(A_hi, CH_GH, AH_SBO, A_hi, CH_GL, AH_SBO, r, half, v);
            // ---------------- backward layer 2
            sync_issue();
            stamp();   // 8
            wait_mma();
            stamp();   // 9
            umma::tmem_ld16(taddr + T_S, v);
            umma::tmem_wait_ld();
            {
                float s[16];
                lrelu_slopes(A_hi, r, CH_H1, half, AH_SBO, s);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __fmul_rn(v[i], s[i]);  // dz1
                acc_db1 += warp_colsum16(v, lane);
            }
            wait_dw();   // dW2 has read G (dz2)
            store_half_row2(A_hi, CH_GH, AH_SBO, A_hi, CH_GL, AH_SBO, r, half, v);
            // ---------------- backward layer 1
            sync_issue();
            stamp();   // 10
            wait_mma();
            stamp();   // 11
            umma::tmem_ld16(taddr + T_S, v);
            umma::tmem_wait_ld();
            {
                float s[16];
                lrelu_slopes(A_hi, r, CH_H0, half, AH_SBO, s);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __fmul_rn(v[i], s[i]);  // dz0
                acc_db0 += warp_colsum16(v, lane);
            }
            wait_dw();   // dW1 has read G (dz1) and h0
            store_half_row2(A_hi, CH_GH, AH_SBO, A_hi, CH_GL, AH_SBO, r, half, v);
            // ---------------- backward layer 0: issued, not waited for (d_enc goes to the scatter warps)
            sync_issue();
            stamp();   // 12
        }
        if (did_work) wait_dw();   // the last dW0: the weight gradients in TMEM are complete
        umma::fence_after_sync();
        // SIMT-side sums: every warp holds partial column sums over its 32 rows
        wred[warp * 80 + 0 * 16 + colsum_index(lane)] = acc_db2;   // lanes l and l^1 hold the same column: benign duplicate store
        wred[warp * 80 + 1 * 16 + colsum_index(lane)] = acc_db1;
        wred[warp * 80 + 2 * 16 + colsum_index(lane)] = acc_db0;
        wred[warp * 80 + 3 * 16 + colsum_index(lane)] = acc_dw3;
        if (lane == 0) wred[warp * 80 + 64] = acc_db3;
        umma::fence_before_sync();
    }

    // ================= the MLP gradients of this CTA -> its slot of the partials workspace -> grid-wide reduction.
    // Only the epilogue and issue warps (288 threads, named barrier 1) take part: the scatter warps are still busy with the LAST
    // tile's d(encoding) -- a tail with nothing to overlap, were it not for this work -- and join at the final __syncthreads.
    constexpr int NTF = N_EPI + 32;   // threads of the tail
    if (warp <= WARP_ISSUE) {
        float *mine = partials + (size_t)blockIdx.x * PTOTAL;
        kstamp(2);   // epilogue warp 0 has left its tile loop
        umma::named_bar_sync_imm<1>(NTF);
        umma::fence_after_sync();
        // Weight gradients: dW_l[o][k] = TMEM lane o (the G_hi rows) + TMEM lane 32 + o (the G_lo rows).  Warps 0 / 4 read the
        // first block, warps 1 / 5 (lane quadrant 1) the second, 16 columns at a time, into two images in shared memory (the
        // operand buffers are dead: every MMA has completed); rows padded by one float against bank conflicts.
        float *stg = reinterpret_cast<float *>(smem);                   // [2][SW_TOTAL]
        constexpr int SW0 = 0, SW1 = 32 * 33, SW2 = 2 * 32 * 33, SW_TOTAL = 2 * 32 * 33 + 32 * 65;
        static_assert(2 * SW_TOTAL * sizeof(float) <= 2 * ENC_BYTES + AH_BYTES, "staging area of the weight-gradient dump");
        if ((warp & 3) < 2 && warp < 8) {
            const int pass = warp & 3, half = warp >> 2;
            float w16[16];
            auto dump = [&](uint32_t tcol, int dst, int ld, int ncols) {
                for (int c0 = 16 * half; c0 < ncols; c0 += 32) {
                    if (did_work) {
                        umma::tmem_ld16(tmem + ((uint32_t)(pass * 32) << 16) + tcol + c0, w16);
                        umma::tmem_wait_ld();
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) w16[i] = 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) stg[pass * SW_TOTAL + dst + lane * ld + c0 + i] = w16[i];
                }
            };
            dump(T_DW0, SW0, 33, 32);
            dump(T_DW1, SW1, 33, 32);
            dump(T_DW2, SW2, 65, 64);
        }
        umma::fence_before_sync();
        umma::named_bar_sync_imm<1>(NTF);
        // coalesced copy-out: hi block + lo block
        for (int i = t; i < PW3; i += NTF) {
            int si;
            if (i < PW1) si = SW0 + (i >> 5) * 33 + (i & 31);
            else if (i < PW2) si = SW1 + ((i - PW1) >> 5) * 33 + ((i - PW1) & 31);
            else si = SW2 + ((i - PW2) >> 6) * 65 + ((i - PW2) & 63);
            mine[i] = stg[si] + stg[SW_TOTAL + si];
        }
        // SIMT-side column sums of the 4 warps of each half
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
        kstamp(4);   // partials written
        // ---- gW / gb += sum over the CTAs' rows.  One grid-wide barrier replaces a separate reduction kernel: the grid is sized
        // so that every CTA is resident (launch_bwd_ws_n), so every CTA can wait for all rows and then sum its share of the
        // columns -- units of 16 columns x 18 row slices; a slice adds its rows in order with four loads in flight, the slice
        // sums are added in order: the summation tree is a function of the grid size only (deterministic).
        // sync[0] counts arrivals, sync[1] departures; the last CTA to leave clears both (the workspace starts zero-filled).
        // A CTA that has waited 10 s raises sync[2] (the host reads it: NAFEngine.check_health) and leaves without reducing.
        if (sync != nullptr) {
            __shared__ uint32_t s_ok;
            __threadfence();
            umma::named_bar_sync_imm<1>(NTF);
            if (t == 0) {
                atomicAdd(sync, 1u);
                uint32_t seen;
                unsigned long long t0 = 0;
                s_ok = 1u;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(sync) : "memory");
                    if (seen >= gridDim.x) break;
                    __nanosleep(64);
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t0 == 0) t0 = now;
                    if (now - t0 > 10000000000ull) {   // 10 s: a CTA of this grid never became resident
                        atomicExch(sync + 2, 1u);
                        s_ok = 0u;
                        break;
                    }
                }
            }
            umma::named_bar_sync_imm<1>(NTF);
            kstamp(5);   // grid barrier passed
            if (s_ok) {
                const int rows = (int)gridDim.x, j = t & 15, k = t >> 4;   // NTF = 288 -> k in 0..17
                constexpr int SL = NTF / 16;
                static_assert(SL * 16 <= (int)WRED_FLOATS, "wred too small for the final reduction");
                for (int u = blockIdx.x; u * 16 < PTOTAL; u += rows) {
                    const int col = u * 16 + j;
                    float s = 0.f;
                    if (col < PTOTAL) {
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                        int bb = k;
                        for (; bb + 3 * SL < rows; bb += 4 * SL) {
                            s0 += __ldcg(partials + (size_t)bb * PTOTAL + col);
                            s1 += __ldcg(partials + (size_t)(bb + SL) * PTOTAL + col);
                            s2 += __ldcg(partials + (size_t)(bb + 2 * SL) * PTOTAL + col);
                            s3 += __ldcg(partials + (size_t)(bb + 3 * SL) * PTOTAL + col);
                        }
                        float t0 = 0.f, t1 = 0.f, t2 = 0.f;   // at most three left
                        if (bb < rows) t0 = __ldcg(partials + (size_t)bb * PTOTAL + col);
                        if (bb + SL < rows) t1 = __ldcg(partials + (size_t)(bb + SL) * PTOTAL + col);
                        if (bb + 2 * SL < rows) t2 = __ldcg(partials + (size_t)(bb + 2 * SL) * PTOTAL + col);
                        s = ((s0 + s1) + (s2 + s3)) + ((t0 + t1) + t2);
                    }
                    wred[k * 16 + j] = s;
                    umma::named_bar_sync_imm<1>(NTF);
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
                    umma::named_bar_sync_imm<1>(NTF);
                }
            }
            kstamp(6);   // reduction done
            if (t == 0 && atomicAdd(sync + 1, 1u) == gridDim.x - 1) { sync[0] = 0u; sync[1] = 0u; }
        }
    }
    // every role is done with tensor memory (the scatter warps have read the last d(encoding))
    umma::fence_before_sync();
    __syncthreads();
    kstamp(3);
    if (warp == 0) umma::tmem_dealloc(tmem, T_COLS);
}

template <int SRC, int C, int NSW>
int launch_bwd_ws_n(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, uint64_t P, const float *dsig, float *grad_table,
                    float *partials, const uint8_t *stash, long long *stamps, int grid, const nafb_mlp_grads &gr, cudaStream_t s) {
    constexpr int NTW = block_threads(NSW);
    static bool configured[NAFB_MAX_DEVICES] = {};
    NAFB_CONFIGURE_SMEM(configured, (k_density_bwd_ws<SRC, C, NSW>), (int)BWS_SMEM, "density_backward(tc)");
    // The grid barrier needs every CTA resident.  cudaOccupancyMaxActiveBlocksPerMultiprocessor (and with it cooperative launch)
    // answers 1 for ANY kernel that allocates tensor memory, whatever its size (scripts/probe/occ_probe.cu, measured on this
    // pool's B200), although the hardware co-schedules such CTAs; so the residency is computed here from the kernel's own
    // resources: two CTAs per SM when shared memory, registers (allocated per warp) and the 512 TMEM columns allow it.
    static int per_sm[NAFB_MAX_DEVICES] = {};
    const int dev = nafb_current_device();
    if (per_sm[dev] == 0) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, k_density_bwd_ws<SRC, C, NSW>);
        if (e != cudaSuccess) NAFB_FAIL(NAFB_ERR_CUDA, "density_backward(tc): %s", cudaGetErrorString(e));
        int smem_sm = 0, regs_sm = 0, real_dev = 0;
        cudaGetDevice(&real_dev);
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, real_dev);
        cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, real_dev);
        const int by_smem = smem_sm / (int)(BWS_SMEM + fa.sharedSizeBytes + 1024);
        const int by_regs = regs_sm / (((fa.numRegs * 32 + 255) / 256 * 256) * (NTW / 32));
        int n = by_smem < by_regs ? by_smem : by_regs;
        if (n > (int)(512 / T_COLS)) n = 512 / T_COLS;
        if (n < 1) NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "density_backward(tc): the kernel does not fit on this device (%d B shared memory, %d registers per SM)", smem_sm, regs_sm);
        per_sm[dev] = n;
        if (nafb_debug_flags() & 32)
            fprintf(stderr, "[nafb] k_density_bwd_ws<%d,%d,%d>: %d CTA(s) per SM (shared memory %d, registers %d: %d regs x %d threads), %u B shared memory\n", SRC, C,
                    NSW, n, by_smem, by_regs, fa.numRegs, NTW, BWS_SMEM);
    }
    const int cap = nafb_sm_count() * per_sm[dev];
    if (grid > cap) grid = cap;
    // behind the partials: 4096 B of phase time stamps (debug), then the words of the grid barrier
    uint32_t *sync = gr.gW[0] || gr.gb[0] || gr.gW[1] ? reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(stamps) + 4096) : nullptr;
    k_density_bwd_ws<SRC, C, NSW><<<grid, NTW, BWS_SMEM, s>>>(gp, mp, sp, P, dsig, grad_table, partials, stash, stamps, nafb_debug_flags(), gr, sync);
    NAFB_CHECK_LAUNCH("density_backward(tc)");
    return NAFB_OK;
}

// Scatter warps per CTA.  Measured at chest_50 (1024 x 192, two CTAs per SM): 4 warps (13 warps x 64 registers) 138 us, 8 warps
// (17 warps x 56 registers) 124 us; the single-role kernel this one replaced (scatter from the MMA wait slots of the epilogue
// warps) took 146 us.
constexpr int SCATTER_WARPS = 8;

template <int SRC, int C>
int launch_bwd_ws_t(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, uint64_t P, const float *dsig, float *grad_table,
                    float *partials, const uint8_t *stash, long long *stamps, int grid, const nafb_mlp_grads &gr, cudaStream_t s) {
    return launch_bwd_ws_n<SRC, C, SCATTER_WARPS>(gp, mp, sp, P, dsig, grad_table, partials, stash, stamps, grid, gr, s);
}

}  // namespace

int nafb_launch_bwd_ws(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, int src, uint64_t P, const float *dsig,
                       float *grad_table, float *partials, const void *stash, long long *stamps, int grid, const nafb_mlp_grads &gr, cudaStream_t s) {
#define CALL(S_, C_) launch_bwd_ws_t<S_, C_>(gp, mp, sp, P, dsig, grad_table, partials, (const uint8_t *)stash, stamps, grid, gr, s)
    switch (gp.C) {
        case 1: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 1) : CALL(NAFB_SRC_RAYS, 1);
        case 2: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 2) : CALL(NAFB_SRC_RAYS, 2);
        case 4: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 4) : CALL(NAFB_SRC_RAYS, 4);
        default: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 8) : CALL(NAFB_SRC_RAYS, 8);
    }
#undef CALL
}
