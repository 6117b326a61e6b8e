// density.cu -- fused "encoder + density MLP" kernels (reference src/network/network.py:34-58
// on top of src/encoder/hashencoder), fp32 SIMT edition.
//
//   k_density_fwd : points (tensor | rays+sampling | voxel lattice) -> hash-grid gather ->
//                   MLP -> sigma [-> Beer-Lambert partial sums per ray]
//   k_density_bwd : recompute the forward for a 128-point tile, back-propagate through the
//                   MLP (dW accumulated in shared memory per CTA, reduced deterministically by
//                   k_reduce_partials) and scatter d(encoding) into the gradient table.
//
// Data layout inside a CTA: activations live in shared memory as feature-major tiles
// T[32 features][TS = 132] (128 points + 4 floats of padding, so that 8 consecutive rows start
// in 8 different bank quads and every LDS.128 below is conflict free).  All matrix products are
// register-blocked fp32 FMAs: 4 points x 8 outputs per thread, k/o stepped by 4 with LDS.128 on
// both operands (10.7 FMA per LDS).  The order of every dot product is "bias, then k ascending".
//
// Scope of this edition: in_dim (L*C) == 32, hidden == 32, out_dim == 1, any number of layers
// up to NAFB_MAX_LAYERS, skips anywhere in [1, n_layers-2] -- i.e. every shipped config
// (config/*.yaml: 16x2 hash grid, 4x32 MLP, skips [2]).
#include <stdlib.h>

#include "common.cuh"
#include "sampler.cuh"

// tcgen05 edition (density_tc.cu)
bool nafb_tc_config_ok(const nafb_grid *grid, const nafb_mlp *mlp);
int nafb_tc_bwd_grid(uint64_t n_tiles);
uint64_t nafb_tc_stash_bytes(uint64_t n_points);
int nafb_launch_fwd_tc(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, int src, uint64_t P, float *sigma, float *acc, float *z,
                       float *pts, int32_t *flags, void *stash, const nafb_loss_tail *tail, cudaStream_t s);
// the tensor-core backward kernel (density_bwd_tc.cu)
int nafb_launch_bwd_ws(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, int src, uint64_t P, const float *dsig,
                       float *grad_table, float *partials, const void *stash, long long *stamps, int grid, const nafb_mlp_grads &gr, cudaStream_t s);

namespace {

constexpr int TILE = 128;  // points per CTA tile == threads per CTA
constexpr int F = 32;      // feature width (in_dim == hidden == 32)
constexpr int TS = 132;    // tile row stride in floats
constexpr int TILE_FLOATS = F * TS;

struct MlpLayout {
    int n_layers;
    int in[NAFB_MAX_LAYERS], out[NAFB_MAX_LAYERS], woff[NAFB_MAX_LAYERS], boff[NAFB_MAX_LAYERS];
    int total;  // floats, multiple of 4
};

__host__ __device__ inline MlpLayout make_layout(const nafb_mlp &m) {
    MlpLayout lo;
    lo.n_layers = (int)m.n_layers;
    int off = 0;
    for (int l = 0; l < lo.n_layers; ++l) {
        lo.in[l] = (l == 0) ? (int)m.in_dim : (int)m.hidden + (((m.skip_mask >> l) & 1u) ? (int)m.in_dim : 0);
        lo.out[l] = (l == lo.n_layers - 1) ? (int)m.out_dim : (int)m.hidden;
        lo.woff[l] = off;
        off += lo.in[l] * lo.out[l];
    }
    for (int l = 0; l < lo.n_layers; ++l) {
        lo.boff[l] = off;
        off += lo.out[l];
    }
    lo.total = (off + 3) & ~3;
    return lo;
}

__device__ __forceinline__ void load_weights(const nafb_mlp &mp, const MlpLayout &lo, float *Ws) {
    for (int l = 0; l < lo.n_layers; ++l) {
        const int nw = lo.in[l] * lo.out[l];
        for (int i = threadIdx.x; i < nw; i += blockDim.x) Ws[lo.woff[l] + i] = __ldg(mp.W[l] + i);
        for (int i = threadIdx.x; i < lo.out[l]; i += blockDim.x) Ws[lo.boff[l] + i] = __ldg(mp.b[l] + i);
    }
}

// acc[i][j] += sum_{k<32} AT[k][4pg+i] * Wrow[(8og+j)*ldw + k]          (forward: W is [out][in])
__device__ __forceinline__ void gemm_fwd32(const float *__restrict__ AT, const float *__restrict__ Wrow, int ldw, int pg, int og,
                                           float (&acc)[4][8]) {
#pragma unroll 2
    for (int k0 = 0; k0 < F; k0 += 4) {
        float4 a[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) a[kk] = *reinterpret_cast<const float4 *>(AT + (k0 + kk) * TS + 4 * pg);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 w = *reinterpret_cast<const float4 *>(Wrow + (8 * og + j) * ldw + k0);
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                acc[0][j] = __fmaf_rn(a[kk].x, wv[kk], acc[0][j]);
                acc[1][j] = __fmaf_rn(a[kk].y, wv[kk], acc[1][j]);
                acc[2][j] = __fmaf_rn(a[kk].z, wv[kk], acc[2][j]);
                acc[3][j] = __fmaf_rn(a[kk].w, wv[kk], acc[3][j]);
            }
        }
    }
}

// acc[i][j] += sum_{o<32} GT[o][4pg+i] * Wcol[o*ldw + 8og + j]          (input gradient: G . W)
__device__ __forceinline__ void gemm_dx32(const float *__restrict__ GT, const float *__restrict__ Wcol, int ldw, int pg, int og,
                                          float (&acc)[4][8]) {
#pragma unroll 2
    for (int o0 = 0; o0 < F; o0 += 4) {
        float4 a[4];
#pragma unroll
        for (int oo = 0; oo < 4; ++oo) a[oo] = *reinterpret_cast<const float4 *>(GT + (o0 + oo) * TS + 4 * pg);
#pragma unroll
        for (int oo = 0; oo < 4; ++oo) {
            const float4 w0 = *reinterpret_cast<const float4 *>(Wcol + (o0 + oo) * ldw + 8 * og);
            const float4 w1 = *reinterpret_cast<const float4 *>(Wcol + (o0 + oo) * ldw + 8 * og + 4);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            const float av[4] = {a[oo].x, a[oo].y, a[oo].z, a[oo].w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(av[i], wv[j], acc[i][j]);
        }
    }
}

// dWs[(i)*ldw + koff + j] += sum_p GT[i][p] * XT[j][p]   for the 32x32 block (i, j), 8 outputs / thread
__device__ __forceinline__ void gemm_dw32(const float *__restrict__ GT, const float *__restrict__ XT, float *__restrict__ dW, int ldw) {
    const int ig = threadIdx.x >> 3, jg = threadIdx.x & 7;  // i = ig + 16 ii, j = jg + 8 jj
    float acc[2][4];
#pragma unroll
    for (int ii = 0; ii < 2; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[ii][jj] = 0.f;
#pragma unroll 2
    for (int p0 = 0; p0 < TILE; p0 += 4) {
        float4 g[2], x[4];
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) g[ii] = *reinterpret_cast<const float4 *>(GT + (ig + 16 * ii) * TS + p0);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) x[jj] = *reinterpret_cast<const float4 *>(XT + (jg + 8 * jj) * TS + p0);
#pragma unroll
        for (int ii = 0; ii < 2; ++ii)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float s = acc[ii][jj];
                s = __fmaf_rn(g[ii].x, x[jj].x, s);
                s = __fmaf_rn(g[ii].y, x[jj].y, s);
                s = __fmaf_rn(g[ii].z, x[jj].z, s);
                s = __fmaf_rn(g[ii].w, x[jj].w, s);
                acc[ii][jj] = s;
            }
    }
#pragma unroll
    for (int ii = 0; ii < 2; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) dW[(ig + 16 * ii) * ldw + jg + 8 * jj] += acc[ii][jj];
}

// gather the 32-wide encoding of this thread's point into column t of tile TE
template <int C>
__device__ __forceinline__ void gather_to_tile(const GridParams &gp, const float (&x01)[3], float *TE) {
    constexpr int L = F / C;
#pragma unroll 2
    for (int l = 0; l < L; ++l) {
        const LevelParams lp = gp.lv[l];
        const float *__restrict__ tab = gp.table + (size_t)lp.offset * C;
        uint32_t g[3];
        float f[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) locate(x01[d], lp.scale, g[d], f[d]);
        float v[8][C];
#pragma unroll
        for (uint32_t idx = 0; idx < 8; ++idx)
            load_entry<C>(tab, grid_entry3(lp, g[0] + (idx & 1u), g[1] + ((idx >> 1) & 1u), g[2] + (idx >> 2)), v[idx]);
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
        for (int c = 0; c < C; ++c) TE[(l * C + c) * TS + threadIdx.x] = res[c];
    }
}

// scatter column t of tile TG (d loss / d encoding) into the gradient table
template <int C>
__device__ __forceinline__ void scatter_from_tile(const GridParams &gp, const float (&x01)[3], const float *TG, float *grad_table) {
    constexpr int L = F / C;
#pragma unroll 2
    for (int l = 0; l < L; ++l) {
        const LevelParams lp = gp.lv[l];
        float *tab = grad_table + (size_t)lp.offset * C;
        uint32_t g[3];
        float f[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) locate(x01[d], lp.scale, g[d], f[d]);
        float gr[C];
#pragma unroll
        for (int c = 0; c < C; ++c) gr[c] = TG[(l * C + c) * TS + threadIdx.x];
#pragma unroll
        for (uint32_t idx = 0; idx < 8; ++idx) {
            float w = 1.0f;
#pragma unroll
            for (int d = 0; d < 3; ++d) w = __fmul_rn(w, (idx & (1u << d)) ? f[d] : __fsub_rn(1.0f, f[d]));
            float v[C];
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = __fmul_rn(w, gr[c]);
            red_add_entry<C>(tab, grid_entry3(lp, g[0] + (idx & 1u), g[1] + ((idx >> 1) & 1u), g[2] + (idx >> 2)), v);
        }
    }
}

// hidden layer l of the forward pass: Tout = lrelu(W_l . [TE?, Tin] + b_l)
__device__ __forceinline__ void hidden_layer(const MlpLayout &lo, const float *Ws, int l, bool skip, const float *TE, const float *Tin,
                                             float *Tout) {
    const int pg = threadIdx.x & 31, og = threadIdx.x >> 5;
    float acc[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float b = Ws[lo.boff[l] + 8 * og + j];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = b;
    }
    const float *W = Ws + lo.woff[l];
    const int ldw = lo.in[l];
    if (l == 0) {
        gemm_fwd32(TE, W, ldw, pg, og, acc);
    } else if (skip) {  // cat([encoding, h])  (network.py:45-46)
        gemm_fwd32(TE, W, ldw, pg, og, acc);
        gemm_fwd32(Tin, W + F, ldw, pg, og, acc);
    } else {
        gemm_fwd32(Tin, W, ldw, pg, og, acc);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float4 v;
        v.x = leaky_relu(acc[0][j]); v.y = leaky_relu(acc[1][j]); v.z = leaky_relu(acc[2][j]); v.w = leaky_relu(acc[3][j]);
        *reinterpret_cast<float4 *>(Tout + (8 * og + j) * TS + 4 * pg) = v;
    }
}

// ------------------------------------------------------------------------------------- forward
template <int SRC, int C>
__global__ void __launch_bounds__(TILE, 3) k_density_fwd(const GridParams gp, const nafb_mlp mp, const SamplerParams sp, const uint64_t P,
                                                      float *__restrict__ sigma, float *__restrict__ acc_out, float *__restrict__ z_out,
                                                      float *__restrict__ pts_out, int32_t *__restrict__ flags) {
    extern __shared__ __align__(16) float smem[];
    const MlpLayout lo = make_layout(mp);
    float *Ws = smem;
    float *TE = Ws + lo.total;
    float *TA = TE + TILE_FLOATS;
    float *TB = TA + TILE_FLOATS;
    load_weights(mp, lo, Ws);
    __syncthreads();
    const int t = threadIdx.x;
    const uint64_t n_tiles = (P + TILE - 1) / TILE;
    int bad = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t p = tile * TILE + t;
        const bool valid = p < P;
        float x[3] = {0.f, 0.f, 0.f};
        if (valid) {
            fetch_point<SRC>(sp, p, x);
            if (!(x[0] >= -sp.bound && x[0] <= sp.bound && x[1] >= -sp.bound && x[1] <= sp.bound && x[2] >= -sp.bound && x[2] <= sp.bound))
                bad |= 1;
            if (SRC == NAFB_SRC_RAYS && pts_out) {
                pts_out[3 * p] = x[0]; pts_out[3 * p + 1] = x[1]; pts_out[3 * p + 2] = x[2];
            }
        }
        float x01[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) x01[d] = normalise01(x[d], sp.bound, sp.inv_2bound);
        gather_to_tile<C>(gp, x01, TE);
        __syncthreads();
        const float *cur = TE;
        for (int l = 0; l < lo.n_layers - 1; ++l) {
            float *dst = (l & 1) ? TB : TA;
            hidden_layer(lo, Ws, l, (mp.skip_mask >> l) & 1u, TE, cur, dst);
            __syncthreads();
            cur = dst;
        }
        // head (out_dim == 1): thread per point
        const int ll = lo.n_layers - 1;
        float s = Ws[lo.boff[ll]];
#pragma unroll 8
        for (int k = 0; k < F; ++k) s = __fmaf_rn(cur[k * TS + t], Ws[lo.woff[ll] + k], s);
        const float y = head_activation(s, mp.head);
        if (valid) {
            if (sigma) sigma[p] = y;
            if (!(fabsf(y) <= 3.4028234e38f)) bad |= 2;
        }
        if constexpr (SRC == NAFB_SRC_RAYS) {
            if (acc_out || z_out) {
                float contrib = 0.f;
                uint32_t r = 0xffffffffu;
                if (valid) {
                    r = (uint32_t)(p / sp.n_samples);
                    const uint32_t i = (uint32_t)(p - (uint64_t)r * sp.n_samples);
                    const RayRegs R = load_ray(sp, r);
                    contrib = __fmul_rn(y, ray_delta(sp, R, r, i));   // render.py:201
                    if (z_out)
                        z_out[p] = z_sample(R.near, R.far, i, sp.n_samples, sp.lin_step, sp.perturb != 0,
                                            jitter_for(sp, r));
                }
                if (acc_out) {
                    // warp-shuffle segmented reduction over the ray id, one atomic per (warp, ray) segment
                    const unsigned lane = t & 31;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const float up = __shfl_down_sync(0xffffffffu, contrib, o);
                        const uint32_t ur = __shfl_down_sync(0xffffffffu, r, o);
                        if (lane + o < 32 && ur == r) contrib += up;
                    }
                    const uint32_t prev = __shfl_up_sync(0xffffffffu, r, 1);
                    if (valid && (lane == 0 || prev != r)) atomicAdd(acc_out + r, contrib);
                }
            }
        }
        __syncthreads();  // tiles are reused by the next iteration
    }
    if (flags && bad) atomicOr(flags, bad);
}

// ------------------------------------------------------------------------------------- backward
template <int SRC, int C>
__global__ void __launch_bounds__(TILE, 2) k_density_bwd(const GridParams gp, const nafb_mlp mp, const SamplerParams sp, const uint64_t P,
                                                      const float *__restrict__ dsig_or_dacc, float *__restrict__ grad_table,
                                                      float *__restrict__ partials) {
    extern __shared__ __align__(16) float smem[];
    const MlpLayout lo = make_layout(mp);
    const int nh = lo.n_layers - 1;  // hidden layers
    float *Ws = smem;
    float *dWs = Ws + lo.total;
    float *TE = dWs + lo.total;
    float *TH = TE + TILE_FLOATS;          // TH[l] = TH + l*TILE_FLOATS, l < nh
    float *gs = TH + nh * TILE_FLOATS;     // [TILE] head pre-activation gradients
    load_weights(mp, lo, Ws);
    for (int i = threadIdx.x; i < lo.total; i += blockDim.x) dWs[i] = 0.f;
    __syncthreads();
    const int t = threadIdx.x;
    const int pg = t & 31, og = t >> 5;
    const uint64_t n_tiles = (P + TILE - 1) / TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t p = tile * TILE + t;
        const bool valid = p < P;
        float x[3] = {0.f, 0.f, 0.f};
        float dsig = 0.f;
        if (valid) {
            fetch_point<SRC>(sp, p, x);
            if constexpr (SRC == NAFB_SRC_RAYS) {
                const uint32_t r = (uint32_t)(p / sp.n_samples);
                const uint32_t i = (uint32_t)(p - (uint64_t)r * sp.n_samples);
                const RayRegs R = load_ray(sp, r);
                dsig = __fmul_rn(__ldg(dsig_or_dacc + r), ray_delta(sp, R, r, i));
            } else {
                dsig = __ldg(dsig_or_dacc + p);
            }
        }
        float x01[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) x01[d] = normalise01(x[d], sp.bound, sp.inv_2bound);
        // ---- recompute forward, keeping every activation tile
        gather_to_tile<C>(gp, x01, TE);
        __syncthreads();
        for (int l = 0; l < nh; ++l) {
            hidden_layer(lo, Ws, l, (mp.skip_mask >> l) & 1u, TE, l ? TH + (l - 1) * TILE_FLOATS : TE, TH + l * TILE_FLOATS);
            __syncthreads();
        }
        // ---- head
        float *Tlast = TH + (nh - 1) * TILE_FLOATS;
        {
            float s = Ws[lo.boff[nh]];
#pragma unroll 8
            for (int k = 0; k < F; ++k) s = __fmaf_rn(Tlast[k * TS + t], Ws[lo.woff[nh] + k], s);
            const float y = head_activation(s, mp.head);
            gs[t] = dsig * head_derivative(s, y, mp.head);
        }
        __syncthreads();
        if (t < F) {  // dW_head[k] += sum_p gs[p] * h_last[k][p]
            float s = 0.f;
            for (int p0 = 0; p0 < TILE; p0 += 4) {
                const float4 h = *reinterpret_cast<const float4 *>(Tlast + t * TS + p0);
                const float4 g = *reinterpret_cast<const float4 *>(gs + p0);
                s = __fmaf_rn(g.x, h.x, s); s = __fmaf_rn(g.y, h.y, s); s = __fmaf_rn(g.z, h.z, s); s = __fmaf_rn(g.w, h.w, s);
            }
            dWs[lo.woff[nh] + t] += s;
        } else if (t == F) {
            float s = 0.f;
            for (int p0 = 0; p0 < TILE; ++p0) s += gs[p0];
            dWs[lo.boff[nh]] += s;
        }
        __syncthreads();
        {  // dz_last = W_head * gs * lrelu'(h_last), in place (thread t owns column t)
            const float g = gs[t];
#pragma unroll 8
            for (int k = 0; k < F; ++k) {
                const float h = Tlast[k * TS + t];
                Tlast[k * TS + t] = __fmul_rn(__fmul_rn(Ws[lo.woff[nh] + k], g), h > 0.f ? 1.0f : 0.01f);
            }
        }
        __syncthreads();
        // ---- hidden layers, last to first.  G = TH[l] holds dz_l.
        float denc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) denc[i][j] = 0.f;
        for (int l = nh - 1; l >= 0; --l) {
            float *G = TH + l * TILE_FLOATS;
            const bool skip = (mp.skip_mask >> l) & 1u;
            const int ldw = lo.in[l];
            float *dW = dWs + lo.woff[l];
            const float *W = Ws + lo.woff[l];
            float *Tprev = l ? TH + (l - 1) * TILE_FLOATS : nullptr;
            // (d) weight / bias gradients
            if (l == 0) {
                gemm_dw32(G, TE, dW, ldw);
            } else if (skip) {
                gemm_dw32(G, TE, dW, ldw);
                gemm_dw32(G, Tprev, dW + F, ldw);
            } else {
                gemm_dw32(G, Tprev, dW, ldw);
            }
            if (t < F) {
                float s = 0.f;
                for (int p0 = 0; p0 < TILE; p0 += 4) {
                    const float4 g = *reinterpret_cast<const float4 *>(G + t * TS + p0);
                    s += g.x; s += g.y; s += g.z; s += g.w;
                }
                dWs[lo.boff[l] + t] += s;
            }
            __syncthreads();  // dW reads of Tprev / TE are done before anything is overwritten
            // (e) input gradients
            if (l == 0 || skip) {
                gemm_dx32(G, W, ldw, pg, og, denc);  // columns [0,32) of W_l multiply the encoding
            }
            if (l > 0) {
                float acc[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
                gemm_dx32(G, W + (skip ? F : 0), ldw, pg, og, acc);
                // dz_{l-1} = dh * lrelu'(h_{l-1}), in place over h_{l-1}; this thread owns (4pg.., 8og..)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 *cell = reinterpret_cast<float4 *>(Tprev + (8 * og + j) * TS + 4 * pg);
                    const float4 h = *cell;
                    float4 v;
                    v.x = __fmul_rn(acc[0][j], h.x > 0.f ? 1.0f : 0.01f);
                    v.y = __fmul_rn(acc[1][j], h.y > 0.f ? 1.0f : 0.01f);
                    v.z = __fmul_rn(acc[2][j], h.z > 0.f ? 1.0f : 0.01f);
                    v.w = __fmul_rn(acc[3][j], h.w > 0.f ? 1.0f : 0.01f);
                    *cell = v;
                }
            }
            __syncthreads();
        }
        // ---- d(encoding) -> TE (in place; all dW reads of TE happened before the last barrier)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 v = {denc[0][j], denc[1][j], denc[2][j], denc[3][j]};
            *reinterpret_cast<float4 *>(TE + (8 * og + j) * TS + 4 * pg) = v;
        }
        __syncthreads();
        if (valid && grad_table) scatter_from_tile<C>(gp, x01, TE, grad_table);
        __syncthreads();
    }
    // ---- per-CTA partial sums of the MLP gradients
    float *mine = partials + (size_t)blockIdx.x * lo.total;
    for (int i = threadIdx.x; i < lo.total; i += blockDim.x) mine[i] = dWs[i];
}

// gW/gb += sum over CTAs of the partials.  Block = 32 outputs x 32 slices of the CTA range (1024 threads): a slice sums its
// CTAs in order with four independent loads in flight (about three L2 round trips for 296 CTAs instead of nine), the 32 slice
// sums are combined by a fixed tree: the summation order is a function of n_blocks only (deterministic), loads are coalesced.
constexpr int RP_SLICES = 32;
__global__ void __launch_bounds__(32 * RP_SLICES) k_reduce_partials(const nafb_mlp mp, const nafb_mlp_grads gr, const float *__restrict__ partials,
                                                                    int n_blocks) {
    __shared__ float part[RP_SLICES][33];
    const MlpLayout lo = make_layout(mp);
    const int j = threadIdx.x & 31, k = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + j;
    float s = 0.f;
    if (i < lo.total) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int b = k;
        for (; b + 3 * RP_SLICES < n_blocks; b += 4 * RP_SLICES) {
            s0 += __ldcg(partials + (size_t)b * lo.total + i);
            s1 += __ldcg(partials + (size_t)(b + RP_SLICES) * lo.total + i);
            s2 += __ldcg(partials + (size_t)(b + 2 * RP_SLICES) * lo.total + i);
            s3 += __ldcg(partials + (size_t)(b + 3 * RP_SLICES) * lo.total + i);
        }
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;   // at most three left
        if (b < n_blocks) t0 = __ldcg(partials + (size_t)b * lo.total + i);
        if (b + RP_SLICES < n_blocks) t1 = __ldcg(partials + (size_t)(b + RP_SLICES) * lo.total + i);
        if (b + 2 * RP_SLICES < n_blocks) t2 = __ldcg(partials + (size_t)(b + 2 * RP_SLICES) * lo.total + i);
        s = ((s0 + s1) + (s2 + s3)) + ((t0 + t1) + t2);
    }
    part[k][j] = s;
    __syncthreads();
#pragma unroll
    for (int w = RP_SLICES / 2; w >= 1; w >>= 1) {
        if (k < w) part[k][j] += part[k + w][j];
        __syncthreads();
    }
    if (k != 0 || i >= lo.total) return;
    s = part[0][j];
    for (int l = 0; l < lo.n_layers; ++l) {
        const int nw = lo.in[l] * lo.out[l];
        if (i >= lo.woff[l] && i < lo.woff[l] + nw) { if (gr.gW[l]) gr.gW[l][i - lo.woff[l]] += s; return; }
        if (i >= lo.boff[l] && i < lo.boff[l] + lo.out[l]) { if (gr.gb[l]) gr.gb[l][i - lo.boff[l]] += s; return; }
    }
}

int check_mlp(const nafb_grid *grid, const nafb_mlp *mlp) {
    if (!mlp) NAFB_FAIL(NAFB_ERR_INVALID, "nafb_mlp: null");
    if (grid->D != 3) NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "fused density kernels need input_dim == 3 (got %u)", grid->D);
    if (mlp->in_dim != F || mlp->hidden != F || grid->L * grid->C != F)
        NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "fused density kernels need L*C == hidden_dim == 32 (got L*C=%u, in_dim=%u, hidden=%u)",
                  grid->L * grid->C, mlp->in_dim, mlp->hidden);
    if (mlp->out_dim != 1) NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "fused density kernels need out_dim == 1 (got %u)", mlp->out_dim);
    if (mlp->n_layers < 2 || mlp->n_layers > NAFB_MAX_LAYERS)
        NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "fused density kernels need 2 <= num_layers <= %d (got %u)", NAFB_MAX_LAYERS, mlp->n_layers);
    if (mlp->skip_mask & (1u | (1u << (mlp->n_layers - 1)) | ~((1u << mlp->n_layers) - 1u)))
        NAFB_FAIL(NAFB_ERR_INVALID, "skips must lie in [1, num_layers-2] (network.py:17-19)");
    if (mlp->head > NAFB_ACT_NONE) NAFB_FAIL(NAFB_ERR_INVALID, "unknown head activation %u", mlp->head);
    if (mlp->arith > NAFB_ARITH_SIMT) NAFB_FAIL(NAFB_ERR_INVALID, "unknown arithmetic mode %u (nafb_mlp.arith)", mlp->arith);
    for (uint32_t l = 0; l < mlp->n_layers; ++l)
        if (!mlp->W[l] || !mlp->b[l]) NAFB_FAIL(NAFB_ERR_INVALID, "nafb_mlp: layer %u has a null pointer", l);
    return NAFB_OK;
}

size_t fwd_smem_bytes(const MlpLayout &lo) { return sizeof(float) * (size_t)(lo.total + 3 * TILE_FLOATS); }
size_t bwd_smem_bytes(const MlpLayout &lo) { return sizeof(float) * (size_t)(2 * lo.total + lo.n_layers * TILE_FLOATS + TILE); }

int bwd_grid(const MlpLayout &lo, uint64_t n_tiles) {
    const size_t smem = bwd_smem_bytes(lo) + 1024;
    int per_sm = (int)((228 * 1024) / smem);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    const uint64_t cap = (uint64_t)nafb_sm_count() * per_sm;
    return (int)(n_tiles < cap ? n_tiles : cap);
}
// upper bound used for the workspace size (independent of the problem size)
int bwd_grid_max(const MlpLayout &lo) { return bwd_grid(lo, ~0ull); }
uint64_t partials_bytes(const MlpLayout &lo) {
    const int g_simt = bwd_grid_max(lo), g_tc = nafb_tc_bwd_grid(~0ull);
    return (uint64_t)(g_simt > g_tc ? g_simt : g_tc) * lo.total * sizeof(float);
}

template <int SRC, int C>
int launch_fwd(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, uint64_t P, float *sigma, float *acc, float *z,
               float *pts, int32_t *flags, cudaStream_t s) {
    const MlpLayout lo = make_layout(mp);
    const size_t smem = fwd_smem_bytes(lo);
    static bool configured[NAFB_MAX_DEVICES] = {};
    NAFB_CONFIGURE_SMEM(configured, (k_density_fwd<SRC, C>), 200 * 1024, "density_forward");
    const uint64_t n_tiles = (P + TILE - 1) / TILE;
    int per_sm = (int)((228 * 1024) / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    const uint64_t cap = (uint64_t)nafb_sm_count() * per_sm;
    const unsigned grid = (unsigned)(n_tiles < cap ? n_tiles : cap);
    k_density_fwd<SRC, C><<<grid, TILE, smem, s>>>(gp, mp, sp, P, sigma, acc, z, pts, flags);
    NAFB_CHECK_LAUNCH("density_forward");
    return NAFB_OK;
}

template <int SRC, int C>
int launch_bwd(const GridParams &gp, const nafb_mlp &mp, const SamplerParams &sp, uint64_t P, const float *dsig, float *grad_table,
               const nafb_mlp_grads &gr, float *partials, cudaStream_t s) {
    const MlpLayout lo = make_layout(mp);
    const size_t smem = bwd_smem_bytes(lo);
    if (smem > 227 * 1024) NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "density_backward: %u layers need %zu B of shared memory", mp.n_layers, smem);
    static bool configured[NAFB_MAX_DEVICES] = {};
    NAFB_CONFIGURE_SMEM(configured, (k_density_bwd<SRC, C>), 227 * 1024, "density_backward");
    const uint64_t n_tiles = (P + TILE - 1) / TILE;
    const int grid = bwd_grid(lo, n_tiles);
    k_density_bwd<SRC, C><<<grid, TILE, smem, s>>>(gp, mp, sp, P, dsig, grad_table, partials);
    NAFB_CHECK_LAUNCH("density_backward");
    k_reduce_partials<<<(lo.total + 31) / 32, 32 * RP_SLICES, 0, s>>>(mp, gr, partials, grid);
    NAFB_CHECK_LAUNCH("density_backward(reduce)");
    return NAFB_OK;
}

}  // namespace

#define DISPATCH_SRC_C(SRC_, C_, CALL)                                                    \
    do {                                                                                  \
        switch (C_) {                                                                     \
            case 1: if (SRC_ == NAFB_SRC_POINTS) return CALL(NAFB_SRC_POINTS, 1); if (SRC_ == NAFB_SRC_RAYS) return CALL(NAFB_SRC_RAYS, 1); return CALL(NAFB_SRC_VOXELS, 1); \
            case 2: if (SRC_ == NAFB_SRC_POINTS) return CALL(NAFB_SRC_POINTS, 2); if (SRC_ == NAFB_SRC_RAYS) return CALL(NAFB_SRC_RAYS, 2); return CALL(NAFB_SRC_VOXELS, 2); \
            case 4: if (SRC_ == NAFB_SRC_POINTS) return CALL(NAFB_SRC_POINTS, 4); if (SRC_ == NAFB_SRC_RAYS) return CALL(NAFB_SRC_RAYS, 4); return CALL(NAFB_SRC_VOXELS, 4); \
            default: if (SRC_ == NAFB_SRC_POINTS) return CALL(NAFB_SRC_POINTS, 8); if (SRC_ == NAFB_SRC_RAYS) return CALL(NAFB_SRC_RAYS, 8); return CALL(NAFB_SRC_VOXELS, 8); \
        }                                                                                 \
    } while (0)

extern "C" {

static inline bool use_tc(const nafb_grid *grid, const nafb_mlp *mlp) { return mlp->arith == NAFB_ARITH_TC && nafb_tc_config_ok(grid, mlp); }

uint64_t nafb_density_stash_bytes(const nafb_grid *grid, const nafb_mlp *mlp, uint64_t n_points) {
    if (!grid || !mlp || !use_tc(grid, mlp)) return 0;
    return nafb_tc_stash_bytes(n_points);
}

static int density_forward_impl(const nafb_grid *grid, const nafb_mlp *mlp, const nafb_sampler *smp, int src, float *sigma, float *acc,
                                float *z_vals, float *pts_out, int32_t *flags, void *stash, const nafb_loss_tail *tail, nafb_stream_t stream) {
    GridParams gp;
    int rc = nafb_make_grid_params(grid, &gp);
    if (rc) return rc;
    if ((rc = check_mlp(grid, mlp))) return rc;
    SamplerParams sp;
    uint64_t P = 0;
    if ((rc = nafb_make_sampler_params(smp, src, &sp, &P))) return rc;
    if (tail) {
        if (!use_tc(grid, mlp)) NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "density_forward_loss: tensor-core configurations only (launch nafb_density_forward + nafb_mse_loss)");
        if (!acc || !tail->target || !tail->loss_out || !tail->ticket) NAFB_FAIL(NAFB_ERR_INVALID, "density_forward_loss: null pointer");
        if (P == 0) NAFB_FAIL(NAFB_ERR_INVALID, "density_forward_loss: empty batch");
    }
    if (P == 0) return NAFB_OK;
    if (src != NAFB_SRC_RAYS && (acc || z_vals || pts_out)) NAFB_FAIL(NAFB_ERR_INVALID, "density_forward: acc/z_vals/pts_out need the RAYS source");
    cudaStream_t s = (cudaStream_t)stream;
    if (src == NAFB_SRC_VOXELS) stash = nullptr;            // forward-only source (its tiles are lattice blocks, not point ranges)
    if (use_tc(grid, mlp)) return nafb_launch_fwd_tc(gp, *mlp, sp, src, P, sigma, acc, z_vals, pts_out, flags, stash, tail, s);
#define CALL(S_, C_) launch_fwd<S_, C_>(gp, *mlp, sp, P, sigma, acc, z_vals, pts_out, flags, s)
    DISPATCH_SRC_C(src, gp.C, CALL);
#undef CALL
}

int nafb_density_forward(const nafb_grid *grid, const nafb_mlp *mlp, const nafb_sampler *smp, int src, float *sigma, float *acc,
                         float *z_vals, float *pts_out, int32_t *flags, void *stash, nafb_stream_t stream) {
    return density_forward_impl(grid, mlp, smp, src, sigma, acc, z_vals, pts_out, flags, stash, nullptr, stream);
}

int nafb_density_forward_loss(const nafb_grid *grid, const nafb_mlp *mlp, const nafb_sampler *smp, float *acc, int32_t *flags, void *stash,
                              const nafb_loss_tail *loss, nafb_stream_t stream) {
    if (!loss) NAFB_FAIL(NAFB_ERR_INVALID, "density_forward_loss: null loss descriptor");
    return density_forward_impl(grid, mlp, smp, NAFB_SRC_RAYS, nullptr, acc, nullptr, nullptr, flags, stash, loss, stream);
}

uint64_t nafb_density_backward_workspace_bytes(const nafb_mlp *mlp) {
    if (!mlp) return 0;
    const MlpLayout lo = make_layout(*mlp);
    return partials_bytes(lo) + 4096 + 64;   // + debug area (phase time stamps of the tensor-core kernel) + grid-barrier words
}

int nafb_density_backward(const nafb_grid *grid, const nafb_mlp *mlp, const nafb_sampler *smp, int src, const float *dsigma_or_dacc,
                          float *grad_table, const nafb_mlp_grads *grads, void *workspace, const void *stash, nafb_stream_t stream) {
    GridParams gp;
    int rc = nafb_make_grid_params(grid, &gp);
    if (rc) return rc;
    if ((rc = check_mlp(grid, mlp))) return rc;
    if (src == NAFB_SRC_VOXELS) NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "density_backward: the voxel source is forward only");
    SamplerParams sp;
    uint64_t P = 0;
    if ((rc = nafb_make_sampler_params(smp, src, &sp, &P))) return rc;
    if (P == 0) return NAFB_OK;
    if (!dsigma_or_dacc || !grads || !workspace) NAFB_FAIL(NAFB_ERR_INVALID, "density_backward: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (use_tc(grid, mlp)) {
        const int grid_tc = nafb_tc_bwd_grid((P + TILE - 1) / TILE);
        const MlpLayout lo = make_layout(*mlp);
        long long *stamps = reinterpret_cast<long long *>((char *)workspace + partials_bytes(lo));
        // (the tensor-core kernels reduce the per-CTA MLP gradients themselves, behind a grid-wide barrier: no second launch)
        return nafb_launch_bwd_ws(gp, *mlp, sp, src, P, dsigma_or_dacc, grad_table, (float *)workspace, stash, stamps, grid_tc, *grads, s);
    }
#define CALL(S_, C_) launch_bwd<S_, C_>(gp, *mlp, sp, P, dsigma_or_dacc, grad_table, *grads, (float *)workspace, s)
    switch (gp.C) {
        case 1: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 1) : CALL(NAFB_SRC_RAYS, 1);
        case 2: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 2) : CALL(NAFB_SRC_RAYS, 2);
        case 4: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 4) : CALL(NAFB_SRC_RAYS, 4);
        default: return src == NAFB_SRC_POINTS ? CALL(NAFB_SRC_POINTS, 8) : CALL(NAFB_SRC_RAYS, 8);
    }
#undef CALL
}

}  // extern "C"
