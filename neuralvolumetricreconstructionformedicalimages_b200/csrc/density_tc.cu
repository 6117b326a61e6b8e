// density_tc.cu -- tcgen05 edition of the fused FORWARD kernel "ray generation + sampling + hash-grid encoder + density MLP + ray
// integral (+ loss)" for the configuration every shipped YAML uses: L*C == 32 encoding, 4 layers x 32 hidden, skip at layer 2,
// out_dim 1 (reference src/render/render.py:88-131, src/encoder/hashencoder/src/hashencoder.cu:77-198,
// src/network/network.py:34-58, train.py:69-127).  The backward kernel lives in density_bwd_tc.cu.
//
// One CTA = one 128-point tile at a time (persistent over tiles), 256 threads:
//   thread t:  row r = t & 127 (sample point, == TMEM lane), half = t >> 7 owns feature
//              columns [16*half, 16*half+16) of every 32-wide activation of its point.
// Per tile: sampling / ray generation (sampler.cuh) -> gather (pair-merged 128-bit loads) -> 3 MMA phases -> head -> ray
// integral; the encodings (bf16 hi | lo, operand layout), the ray-integral weights and the normalised positions are left in
// the "stash" for the backward pass.  The LAST CTA to retire evaluates the masked chunk-wise MSE (loss.cuh).
// Activations are written ONCE, by the thread that owns the point, as bf16 (hi, lo) pairs in the canonical no-swizzle UMMA
// layout (umma.cuh) and consumed in place by the tensor core:  h_l = lrelu(X . W_l^T + b), A = X (K-major), B = W_l (K-major),
// accumulators in TMEM (fp32), products bf16x3 (hi*hi + hi*lo + lo*hi): |rel err| ~ 2^-16 per product.
// Bias, LeakyReLU and the 1-wide head are fp32 SIMT work in the epilogues (TMEM -> registers -> shared-memory operand).
#include <stdio.h>
#include <stdlib.h>

#include "density_tc.cuh"
#include "loss.cuh"

using namespace tc;

namespace {

// ================================================================================ forward
// CTA = 128 * NQ threads: thread (r, q) = (t & 127, t >> 7) owns the 32 / NQ feature columns [q * 32 / NQ, ...) of point r.
//   NQ = 2 (256 threads, 3 CTAs per SM): every thread gathers HALF of the levels of its point;
//   NQ = 4 (512 threads, 2 CTAs per SM, 64 registers): a QUARTER -- the latency chain of a tile's gather is half as long.
// Measured: both process a tile in ~10 k cycles per SM -- the gather is bound by the SM's outstanding
// L2 misses, not by how the CTAs are cut -- NQ = 2 is ~8 % faster on the voxel lattice and is what the launcher instantiates.
// smem: X_hi | X_lo : 128 rows x 8 chunks [enc(0-3) | h(4-7)], SBO 1024 -> 16 KB each
constexpr uint32_t FX_SBO = 1024, FX_HALF = 16384;
constexpr int TILE_RAYS = 4;   // rays of one tile kept in shared memory (a tile of 128 points spans at most 128 / S + 2 rays)
struct alignas(16) RaySlot { float v[12]; };   // RayRegs: o[3] d[3] near far norm
constexpr uint32_t FWD_SMEM = 2 * FX_HALF + 2 * W_HALF + ((sizeof(SmallParams) + 15) & ~15u) + sizeof(TileCtl) + 3 * TILE * sizeof(float) +
                              TILE_RAYS * sizeof(RaySlot) + 128;

// gather the COLS encoding features [COLS*q, COLS*(q+1)) of one point (COLS / C levels), loads of level li + 1 in flight while
// level li is consumed
template <int C, int COLS>
__device__ __forceinline__ void gather_part(const GridParams &gp, const float (&x01)[3], int q, float (&enc)[COLS]) {
    constexpr int LQ = COLS / C;   // levels per thread
    float v[LQ][8][C];
    auto issue = [&](const int li) {
        const LevelParams lp = gp.lv[q * LQ + li];
        const float *__restrict__ tab = gp.table + (size_t)lp.offset * C;
        uint32_t g[3];
        float f[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) locate(x01[d], lp.scale, g[d], f[d]);
        const uint32_t par = addr_parity8(tab);
        const CellTerms ct = cell_terms3(lp, g[0], g[1], g[2]);
        uint32_t e[8];
        cell_entries8(lp, ct, e);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)   // (y, z) corner; the two x-neighbours share one access when adjacent + aligned
            load_entry_pair<C>(tab, par, e[2 * j], e[2 * j + 1], v[li][2 * j], v[li][2 * j + 1]);
    };
    auto consume = [&](const int li) {
        const float scale = gp.lv[q * LQ + li].scale;
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
        for (int c = 0; c < C; ++c) enc[li * C + c] = res[c];
    };
    issue(0);
#pragma unroll
    for (int li = 0; li < LQ; ++li) {
        if (li + 1 < LQ) issue(li + 1);
        consume(li);
    }
}

template <int SRC, int C, int NQ>
__global__ void __launch_bounds__(128 * NQ, NQ == 2 ? 3 : 2) k_density_fwd_tc(const GridParams gp, const nafb_mlp mp, const SamplerParams sp, const uint64_t P,
                                                         float *__restrict__ sigma, float *__restrict__ acc_out, float *__restrict__ z_out,
                                                         float *__restrict__ pts_out, int32_t *__restrict__ flags, uint8_t *__restrict__ stash,
                                                         const int dbg, const nafb_loss_tail tail, long long *__restrict__ dbg_stamps) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *X_hi = smem, *X_lo = X_hi + FX_HALF;
    uint8_t *W_hi = X_lo + FX_HALF, *W_lo = W_hi + W_HALF;
    SmallParams *small = reinterpret_cast<SmallParams *>(W_lo + W_HALF);
    TileCtl *ctl = reinterpret_cast<TileCtl *>(reinterpret_cast<uint8_t *>(small) + ((sizeof(SmallParams) + 15) & ~15u));
    RaySlot *rayslot = reinterpret_cast<RaySlot *>(ctl + 1);           // [TILE_RAYS]  (sizeof(TileCtl) == 16)
    float *xchg = reinterpret_cast<float *>(rayslot + TILE_RAYS);     // [NQ - 1][128] partial head dot products of parts 1..
    constexpr int COLS = 32 / NQ;                                      // feature columns per thread

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int r = t & 127, q = t >> 7;
    // debug (NAFB_DEBUG_SKIP bit 5): thread 0 of CTA 0 records (clock64, %globaltimer) pairs; the launcher prints them
    long long *kst = dbg_stamps && t == 0 && blockIdx.x == 0 ? dbg_stamps : nullptr;
    int n_kst = 0;
    auto kstamp = [&]() {
        if (kst && n_kst < 118) {
            unsigned long long g, c;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)::"memory");
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)::"memory");
            kst[2 * n_kst] = (long long)c;
            kst[2 * n_kst + 1] = (long long)g;
            ++n_kst;
        }
    };
    kstamp();   // 0: start
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
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(COLS * q);
    const uint32_t x_hi = umma::smem_u32(X_hi), x_lo = umma::smem_u32(X_lo), w_hi = umma::smem_u32(W_hi), w_lo = umma::smem_u32(W_lo);
    constexpr uint32_t IDESC = umma::idesc_bf16(128, 32, 0, 0);
    uint32_t phase = 0;
    int bad = 0;
    kstamp();   // 1: set-up done
    const bool want_delta = q == 0 && (acc_out || z_out || stash);
    // this thread's chunks (8 columns each) of a 32-wide block that starts at chunk `chunk0`
    auto store_part = [&](uint32_t chunk0, const float (&vv)[COLS]) {
#pragma unroll
        for (int c = 0; c < COLS / 8; ++c)
            umma::store_chunk_split(X_hi, X_lo, umma::canon_off(r, chunk0 + (COLS / 8) * q + c, LBO, FX_SBO), vv + 8 * c);
    };
    auto load_acc = [&](float (&vv)[COLS]) {
        if constexpr (COLS == 16) umma::tmem_ld16(taddr, vv);
        else umma::tmem_ldn<COLS>(taddr, vv);
        umma::tmem_wait_ld();
    };

    const uint64_t n_tiles = SRC == NAFB_SRC_VOXELS ? voxel_block_tiles(sp) : (P + TILE - 1) / TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t p = tile * TILE + r;      // index of the point in the outputs
        bool valid = p < P;
        float x[3] = {0.f, 0.f, 0.f};
        if constexpr (SRC == NAFB_SRC_VOXELS) valid = voxel_block_point(sp, tile, (uint32_t)r, x, p);   // 8 x 4 x 4 blocks of the lattice
        float z_mine = 0.f, delta_mine = 0.f;   // RAYS source: this sample's depth and its ray-integral weight delta_i |d|
        uint32_t ray_mine = 0xffffffffu;
        if constexpr (SRC == NAFB_SRC_RAYS) {
            // the rays of this tile (two when n_samples >= 128) are generated / loaded ONCE into shared memory; every thread then
            // places its own sample from there (no global loads, no dependent chain pixels -> poses per thread)
            const uint64_t p_first = tile * TILE, p_last = (p_first + TILE - 1 < P ? p_first + TILE - 1 : P - 1);
            uint32_t ray_first, ray_last;
            if (P <= 0xffffffffull) {   // 32-bit divisions (a 64-bit one costs ~80 instructions, twice per thread and tile)
                ray_first = (uint32_t)p_first / sp.n_samples; ray_last = (uint32_t)p_last / sp.n_samples;
            } else {
                ray_first = (uint32_t)(p_first / sp.n_samples); ray_last = (uint32_t)(p_last / sp.n_samples);
            }
            const bool in_smem = ray_last - ray_first < (uint32_t)TILE_RAYS;
            if (in_smem) {
                if (t <= (int)(ray_last - ray_first)) {
                    const RayRegs R = load_ray(sp, ray_first + (uint32_t)t);
                    RaySlot &sl = rayslot[t];
                    sl.v[0] = R.o[0]; sl.v[1] = R.o[1]; sl.v[2] = R.o[2]; sl.v[3] = R.d[0]; sl.v[4] = R.d[1]; sl.v[5] = R.d[2];
                    sl.v[6] = R.near; sl.v[7] = R.far; sl.v[8] = R.norm;
                }
                __syncthreads();
            }
            if (valid) {
                ray_mine = P <= 0xffffffffull ? (uint32_t)p / sp.n_samples : (uint32_t)(p / sp.n_samples);
                RayRegs R;
                if (in_smem) {
                    const RaySlot &sl = rayslot[ray_mine - ray_first];
                    R.o[0] = sl.v[0]; R.o[1] = sl.v[1]; R.o[2] = sl.v[2]; R.d[0] = sl.v[3]; R.d[1] = sl.v[4]; R.d[2] = sl.v[5];
                    R.near = sl.v[6]; R.far = sl.v[7]; R.norm = sl.v[8];
                } else {
                    R = load_ray(sp, ray_mine);
                }
                ray_sample_and_delta(sp, R, ray_mine, (uint32_t)(p - (uint64_t)ray_mine * sp.n_samples), want_delta, x, z_mine, delta_mine);
            }
        } else if constexpr (SRC == NAFB_SRC_POINTS) {
            if (valid) fetch_point<SRC>(sp, p, x);
        }
        if (valid) {
            if (!(x[0] >= -sp.bound && x[0] <= sp.bound && x[1] >= -sp.bound && x[1] <= sp.bound && x[2] >= -sp.bound && x[2] <= sp.bound))
                bad |= 1;
            if (SRC == NAFB_SRC_RAYS && pts_out && q == 0) {
                pts_out[3 * p] = x[0]; pts_out[3 * p + 1] = x[1]; pts_out[3 * p + 2] = x[2];
            }
        }
        float x01[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) x01[d] = normalise01(x[d], sp.bound, sp.inv_2bound);
        kstamp();   // tile: sampled
        {
            float enc[COLS];
            if (dbg & 2) {
#pragma unroll
                for (int i = 0; i < COLS; ++i) enc[i] = x01[i % 3];
            } else {
                gather_part<C, COLS>(gp, x01, q, enc);
            }
            uint8_t *st = stash ? stash + tile * ST_TILE : nullptr;
#pragma unroll
            for (int c = 0; c < COLS / 8; ++c) {
                uint4 h, l;
                umma::split_chunk(enc + 8 * c, h, l);
                const uint32_t chunk = (COLS / 8) * q + c;
                const uint32_t off = umma::canon_off(r, chunk, LBO, FX_SBO);
                *reinterpret_cast<uint4 *>(X_hi + off) = h;
                *reinterpret_cast<uint4 *>(X_lo + off) = l;
                if (st) {   // what the backward pass reads back: the operand images (default policy: it finds them in L2) ...
                    const uint32_t so = umma::canon_off(r, chunk, LBO, ST_SBO);
                    *reinterpret_cast<uint4 *>(st + so) = h;
                    *reinterpret_cast<uint4 *>(st + ST_HALF + so) = l;
                }
            }
            if (st && q == NQ - 1) {   // ... and the normalised position (the backward scatter needs it again)
                float *tl = reinterpret_cast<float *>(st + ST_TAIL_X01);
                tl[r] = x01[0]; tl[128 + r] = x01[1]; tl[256 + r] = x01[2];
            }
        }
        kstamp();   // tile: gathered
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
        float v[COLS];
        load_acc(v);
#pragma unroll
        for (int i = 0; i < COLS; ++i) v[i] = leaky_relu(v[i] + small->b0[COLS * q + i]);
        store_part(4, v);
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
        load_acc(v);
#pragma unroll
        for (int i = 0; i < COLS; ++i) v[i] = leaky_relu(v[i] + small->b1[COLS * q + i]);
        store_part(4, v);
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
        load_acc(v);
        kstamp();   // tile: three layers done
        // ---------------- head: sigma = act(w3 . lrelu(.) + b3), NQ partial sums per row
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < COLS; ++i) part = __fmaf_rn(leaky_relu(v[i] + small->b2[COLS * q + i]), small->w3[COLS * q + i], part);
        if (q > 0) xchg[(q - 1) * TILE + r] = part;
        umma::fence_before_sync();   // orders the TMEM reads above before the next tile's MMAs
        __syncthreads();
        if (q == 0) {
            float s = part + xchg[r];
            if constexpr (NQ == 4) s = s + (xchg[TILE + r] + xchg[2 * TILE + r]);
            s += small->b3;
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
        __syncthreads();  // xchg / X / the ray slots are reused by the next tile
        kstamp();   // tile: done
    }
    if (flags && bad) atomicOr(flags, bad);
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 32);
    kstamp();   // loop done, TMEM released
    if (dbg_stamps && t == 0) {   // debug: when did every CTA leave its tile loop (ns, %globaltimer)
        unsigned long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)::"memory");
        dbg_stamps[240 + blockIdx.x] = (long long)g;
    }
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
            if (dbg_stamps && t == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)::"memory"); dbg_stamps[236] = (long long)g; }
            float *s_mean = reinterpret_cast<float *>(smem), *s_cnt = s_mean + MSE_GROUP;   // the operand tiles are dead by now
            const uint32_t n = sp.n_rays, chunk = (tail.chunk == 0 || tail.chunk > n) ? n : tail.chunk;
            mse_loss_block(acc_out, tail.target, tail.mask, n, chunk, tail.gscale, tail.loss_out, tail.dacc, tail.zero_pred, s_mean, s_cnt);
            if (t == 0) {
                *tail.ticket = 0u;
                if (tail.done_flag) {   // the loss is out: tell a polling host thread (the stores above must be visible first)
                    __threadfence_system();
                    *reinterpret_cast<volatile uint32_t *>(tail.done_flag) = tail.step_state ? tail.step_state[NAFB_STATE_STEP] + 1u : 1u;
                }
            }
            if (dbg_stamps && t == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)::"memory"); dbg_stamps[237] = (long long)g; }
        }
    }
    kstamp();   // end
    if (kst) kst[239] = n_kst;
}

bool tc_config_ok(const nafb_grid *grid, const nafb_mlp *mlp) {
    return grid->D == 3 && grid->L * grid->C == 32 && mlp->in_dim == 32 && mlp->hidden == 32 && mlp->out_dim == 1 && mlp->n_layers == 4 &&
           mlp->skip_mask == (1u << 2);
}

template <int SRC, int C, int NQ>
int launch_fwd_tc_n(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, uint64_t P, float *sigma, float *acc, float *z, float *pts,
                    int32_t *flags, uint8_t *stash, const nafb_loss_tail &tail, cudaStream_t s) {
    static bool configured[NAFB_MAX_DEVICES] = {};
    NAFB_CONFIGURE_SMEM(configured, (k_density_fwd_tc<SRC, C, NQ>), (int)FWD_SMEM, "density_forward(tc)");
    const uint64_t n_tiles = SRC == NAFB_SRC_VOXELS ? voxel_block_tiles(sp) : (P + TILE - 1) / TILE;
    const uint64_t cap = (uint64_t)nafb_sm_count() * (NQ == 2 ? 3 : 2);
    const unsigned grid = (unsigned)(n_tiles < cap ? n_tiles : cap);
    const int dbg = nafb_debug_flags();
    static long long *d_stamps = nullptr;   // debug only (bit 5): never allocated otherwise
    if ((dbg & 32) && !d_stamps) { cudaMalloc(&d_stamps, (240 + 1024) * sizeof(long long)); }
    if (d_stamps) cudaMemsetAsync(d_stamps, 0, (240 + 1024) * sizeof(long long), s);
    k_density_fwd_tc<SRC, C, NQ><<<grid, 128 * NQ, FWD_SMEM, s>>>(gp, mp, sp, P, sigma, acc, z, pts, flags, stash, dbg, tail, (dbg & 32) ? d_stamps : nullptr);
    NAFB_CHECK_LAUNCH("density_forward(tc)");
    if (d_stamps && getenv("NAFB_FWD_STAMPS")) {
        static long long h[240 + 1024];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, d_stamps, sizeof(h), cudaMemcpyDeviceToHost);
        const int n = (int)h[239];
        {   // distribution of the CTAs' loop-exit times relative to CTA 0's start
            long long mn = 1ll << 62, mx = 0; double sum = 0; int cnt = 0;
            for (unsigned b = 0; b < grid && b < 1024; ++b) { const long long v = h[240 + b] - h[1]; if (v < mn) mn = v; if (v > mx) mx = v; sum += (double)v; ++cnt; }
            fprintf(stderr, "[nafb] fwd<NQ=%d> CTA loop-exit times since CTA 0 started (ns): min %lld mean %.0f max %lld; CTA 0 %lld\n", NQ, mn, sum / cnt, mx, h[240] - h[1]);
            fprintf(stderr, "[nafb]   loss tail (last CTA): starts %lld ns, ends %lld ns after CTA 0 started\n", h[236] - h[1], h[237] - h[1]);
        }
        fprintf(stderr, "[nafb] fwd stamps of CTA 0 (grid %u): idx  cycles  ns  (since start)\n", grid);
        for (int i = 1; i < n; ++i) fprintf(stderr, "[nafb]   %3d %9lld %9lld   (+%lld cyc)\n", i, h[2 * i] - h[0], h[2 * i + 1] - h[1], h[2 * i] - h[2 * i - 2]);
    }
    return NAFB_OK;
}

// Column groups per point.  Measured at chest_50 / 512^3 voxel query / 65536 x 384 forward: NQ = 2 (3 CTAs of 256 threads per SM)
// 86.5 us / 32.3 ms / 7.4 ms, NQ = 4 (2 CTAs of 512) 85.2 us / 34.5 ms / 8.0 ms -- the gather costs ~10 k cycles per tile and SM
// either way (bound by the SM's outstanding L2 misses), so the cheaper-per-tile organisation wins where tiles are many.
constexpr int FWD_NQ = 2;

template <int SRC, int C>
int launch_fwd_tc(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, uint64_t P, float *sigma, float *acc, float *z, float *pts,
                  int32_t *flags, uint8_t *stash, const nafb_loss_tail &tail, cudaStream_t s) {
    return launch_fwd_tc_n<SRC, C, FWD_NQ>(gp, mp, sp, P, sigma, acc, z, pts, flags, stash, tail, s);
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
