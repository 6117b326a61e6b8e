// exchange.cu -- the multi-GPU exchange step of NAF training as ONE kernel over NVLink peer memory:
//
//      reduce-scatter of the flat gradient  +  dense Adam on the owned slice  +  all-gather of the new parameters
//
// Every rank owns a contiguous slice of the flat parameter vector (and keeps Adam's exp_avg / exp_avg_sq for that slice
// only).  For its slice a rank LOADS the gradient of all W ranks straight from their HBM (P2P over NVLink / NVSwitch),
// adds them in rank order, applies Adam and STORES the new parameters into all W replicas.  Per rank that is
// (W-1)/W * n * 4 bytes in and the same out -- the minimum an all-reduce moves -- with the optimizer's HBM traffic divided
// by W and no separate collective launches: the transfer overlaps the arithmetic element by element.
//
// (What it replaces: NCCL all-reduce of 57 MB, then the dense Adam pass of adam.cu over 456 MB of HBM traffic.)
//
// Synchronisation is by epoch flags in peer memory (epoch = optimizer step, strictly increasing):
//   arrive[r]  rank r's gradient of this epoch is complete (its backward kernels have finished: stream order) and it no
//              longer reads its parameters -> peers may read r's gradient and overwrite r's parameters;
//   done[r]    rank r has finished writing everybody's parameters (and reading everybody's gradient).
// The kernel ends only after it has seen done[*] of every peer, so the next forward pass (stream order) sees complete
// parameters.  Gradients are double buffered by step parity: the buffer the peers read in step s is cleared by its owner
// in step s+1 (after the arrive barrier of s+1 nobody can still be reading it), the backward pass of s+1 accumulates into
// the other one.  Replicas stay bit-identical by construction: each parameter is computed once, by its owner.
//
// Spins are bounded (NAFB_EXCHANGE_TIMEOUT_NS of %globaltimer): on expiry the kernel raises the error word of the local
// flag block and carries on, so a lost peer cannot hang the GPU.
#include <math.h>
#include <string.h>

#include "adam.cuh"

namespace {

constexpr int MAXR = NAFB_MAX_RANKS;
constexpr unsigned long long TIMEOUT_NS = 4000000000ull;   // 4 s

struct ExchangeParams {
    uint32_t world, rank, epoch;
    float *param[MAXR];
    const float *grad[MAXR];
    uint32_t *flags[MAXR];
    float *grad_zero;
    float *m, *v;          // slice-local
    uint64_t n4, s0, s1;   // float4 units
    AdamConst c;           // host-evaluated constants (state == nullptr)
    uint32_t *state;       // device step state, or nullptr
    double beta1, beta2, eps;
    float gscale;
    float *mc_param;       // NVLS multicast addresses of the parameter vector / this parity's gradient, or nullptr
    const float *mc_grad;
    float *stage[MAXR];    // push mode: every rank's staging area [W][slot4 float4] (slot r of rank w: written by rank r)
    float *grad_local;     // push mode: my (single) gradient buffer, cleared as it is consumed
    uint64_t slot4;        // float4 per staging slot (>= the largest slice)
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// peer loads / stores: bypass L1 (the data lives in another GPU's memory and is never reused by this SM)
__device__ __forceinline__ float4 ld_peer(const float4 *p) { return __ldcg(p); }
__device__ __forceinline__ void st_peer(float4 *p, const float4 v) { __stcg(p, v); }

// NVLS (NVLink SHARP): one load returns the SUM over all replicas, reduced inside the NVSwitch; one store is multicast to all
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4 *mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float4 *mc, const float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// wait until flag word `idx0 + w` of the LOCAL block is >= epoch for every rank w
__device__ __forceinline__ void wait_all(const ExchangeParams &P, uint32_t idx0, uint32_t epoch) {
    if (threadIdx.x < P.world) {
        const uint32_t *f = P.flags[P.rank] + idx0 + threadIdx.x;
        const unsigned long long t0 = globaltimer_ns();
        while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
            if (globaltimer_ns() - t0 > TIMEOUT_NS) {
                atomicExch(P.flags[P.rank] + NAFB_XFLAG_ERROR, 1u + idx0 + threadIdx.x);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

template <int W>
__global__ void __launch_bounds__(512) k_adam_exchange(const ExchangeParams P) {
    const uint32_t world = W > 0 ? (uint32_t)W : P.world;
    __shared__ AdamConst s_c;
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) {
        s_epoch = P.state ? P.state[NAFB_STATE_STEP] + 1u : P.epoch;
        s_c = P.state ? adam_const_from_state(P.state, P.beta1, P.beta2, P.eps, P.gscale) : P.c;
    }
    __syncthreads();
    const uint32_t epoch = s_epoch;
    const AdamConst c = s_c;
    // debug time stamps (ns, %globaltimer) of block 0 in flag words 20..27: start, arrived, slice done, all done
    unsigned long long *stamps = reinterpret_cast<unsigned long long *>(P.flags[P.rank] + 20);
    if (blockIdx.x == 0 && threadIdx.x == 0) stamps[0] = globaltimer_ns();
    // ---- 1. tell every rank (myself included) that my gradient is complete, then wait for everybody's
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(P.flags[threadIdx.x] + NAFB_XFLAG_ARRIVE + P.rank, epoch);
    }
    wait_all(P, NAFB_XFLAG_ARRIVE, epoch);
    if (blockIdx.x == 0 && threadIdx.x == 0) stamps[1] = globaltimer_ns();

    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    // ---- 2. clear my gradient buffer of the other parity (its readers finished before they arrived at this epoch)
    if (P.grad_zero) {
        float4 *z = reinterpret_cast<float4 *>(P.grad_zero);
        for (uint64_t i = gid; i < P.n4; i += stride) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (P.mc_grad) {
        // ---- 3 (NVLS). my slice: the switch adds the W gradients (multimem.ld_reduce), Adam, one multicast store updates all
        // W replicas.  Per GPU and direction ~ n*4 bytes instead of 2*(W-1)/W * n*4.
        float4 *m4 = reinterpret_cast<float4 *>(P.m), *v4 = reinterpret_cast<float4 *>(P.v);
        const float4 *p_in = reinterpret_cast<const float4 *>(P.param[P.rank]);
        const float4 *gmc = reinterpret_cast<const float4 *>(P.mc_grad);
        float4 *pmc = reinterpret_cast<float4 *>(P.mc_param);
        constexpr int UN = 4;
        for (uint64_t i0 = P.s0 + gid; i0 < P.s1; i0 += (uint64_t)UN * stride) {
            float4 s[UN], p[UN], m[UN], v[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const uint64_t i = i0 + (uint64_t)u * stride;
                if (i < P.s1) { s[u] = multimem_ld_reduce_add(gmc + i); p[u] = p_in[i]; m[u] = m4[i - P.s0]; v[u] = v4[i - P.s0]; }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const uint64_t i = i0 + (uint64_t)u * stride;
                if (i < P.s1) {
                    adam_one(p[u].x, s[u].x, m[u].x, v[u].x, c);
                    adam_one(p[u].y, s[u].y, m[u].y, v[u].y, c);
                    adam_one(p[u].z, s[u].z, m[u].z, v[u].z, c);
                    adam_one(p[u].w, s[u].w, m[u].w, v[u].w, c);
                    m4[i - P.s0] = m[u];
                    v4[i - P.s0] = v[u];
                    multimem_st(pmc + i, p[u]);
                }
            }
        }
    } else {
    // ---- 3. my slice: sum of the W gradients (rank order), Adam, new parameters to all W replicas.
        // U elements per thread and iteration keep U*W peer loads in flight (NVLink latency is a few microseconds).
        constexpr int U = W == 2 ? 4 : (W == 4 ? 2 : 1);
        float4 *m4 = reinterpret_cast<float4 *>(P.m), *v4 = reinterpret_cast<float4 *>(P.v);
        const float4 *p_in = reinterpret_cast<const float4 *>(P.param[P.rank]);
        for (uint64_t i0 = P.s0 + gid; i0 < P.s1; i0 += (uint64_t)U * stride) {
            float4 g[U][MAXR], p[U], m[U], v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t i = i0 + (uint64_t)u * stride;
                if (i < P.s1) {
#pragma unroll
                    for (int w = 0; w < MAXR; ++w)
                        if (w < (int)world) g[u][w] = ld_peer(reinterpret_cast<const float4 *>(P.grad[w]) + i);
                    p[u] = p_in[i]; m[u] = m4[i - P.s0]; v[u] = v4[i - P.s0];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t i = i0 + (uint64_t)u * stride;
                if (i < P.s1) {
                    float4 s = g[u][0];
#pragma unroll
                    for (int w = 1; w < MAXR; ++w)
                        if (w < (int)world) { s.x += g[u][w].x; s.y += g[u][w].y; s.z += g[u][w].z; s.w += g[u][w].w; }
                    adam_one(p[u].x, s.x, m[u].x, v[u].x, c);
                    adam_one(p[u].y, s.y, m[u].y, v[u].y, c);
                    adam_one(p[u].z, s.z, m[u].z, v[u].z, c);
                    adam_one(p[u].w, s.w, m[u].w, v[u].w, c);
                    m4[i - P.s0] = m[u];
                    v4[i - P.s0] = v[u];
#pragma unroll
                    for (int w = 0; w < MAXR; ++w)
                        if (w < (int)world) st_peer(reinterpret_cast<float4 *>(P.param[w]) + i, p[u]);
                }
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) stamps[2] = globaltimer_ns();
    // ---- 4. last block out: publish "done" to every rank, then wait until every rank is done with MY replica
    __threadfence_system();
    __syncthreads();
    __shared__ uint32_t s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(P.flags[P.rank] + NAFB_XFLAG_TICKET, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (threadIdx.x == 0) P.flags[P.rank][NAFB_XFLAG_TICKET] = 0u;
    if (threadIdx.x < world) st_release_sys(P.flags[threadIdx.x] + NAFB_XFLAG_DONE + P.rank, epoch);
    wait_all(P, NAFB_XFLAG_DONE, epoch);
    if (threadIdx.x == 0) stamps[3] = globaltimer_ns();
    if (P.state && threadIdx.x == 0) P.state[NAFB_STATE_STEP] = epoch;   // every block has read the old value long ago
}

// ---------------------------------------------------------------------------------------------------- push edition
// NVLink carries posted WRITES at a higher rate than it serves READS (measured here: the pull kernel's remote loads reach
// ~450 GB/s, its remote stores drain at ~570 GB/s, and the two phases hardly overlap), so this edition moves every byte with
// a store:
//   phase 1  rank r copies slice w of ITS gradient into slot r of rank w's staging area (remote stores), for every w != r,
//            clearing its gradient as it goes; when all of that is globally visible it raises pushed[r] on every rank;
//   phase 2  once pushed[*] have arrived, rank w adds the W contributions of its slice in rank order (its own from the
//            gradient buffer, the others from its local staging slots), applies Adam and stores the new parameters into
//            all W replicas; done[w] as before.
// pushed[r] doubles as "r no longer reads its parameters".  Peers never read a gradient buffer, so one gradient buffer per
// rank suffices (no parity) and it leaves the kernel zeroed.  Every block waits for flags that other GPUs raise only after
// ALL their blocks have run phase 1, so the grid must be fully resident (the launcher sizes it from the occupancy).
__device__ __forceinline__ void slice_of(uint64_t n4, uint32_t w, uint32_t world, uint64_t &a, uint64_t &b) {
    const uint64_t base = n4 / world, extra = n4 % world;
    a = w * base + (w < extra ? w : extra);
    b = a + base + (w < extra ? 1 : 0);
}

template <int W>
__global__ void __launch_bounds__(512) k_adam_exchange_push(const ExchangeParams P) {
    const uint32_t world = W > 0 ? (uint32_t)W : P.world;
    __shared__ AdamConst s_c;
    __shared__ uint32_t s_epoch, s_last;
    if (threadIdx.x == 0) {
        s_epoch = P.state ? P.state[NAFB_STATE_STEP] + 1u : P.epoch;
        s_c = P.state ? adam_const_from_state(P.state, P.beta1, P.beta2, P.eps, P.gscale) : P.c;
    }
    __syncthreads();
    const uint32_t epoch = s_epoch;
    const AdamConst c = s_c;
    unsigned long long *stamps = reinterpret_cast<unsigned long long *>(P.flags[P.rank] + 20);   // debug: ns of block 0
    if (blockIdx.x == 0 && threadIdx.x == 0) stamps[0] = globaltimer_ns();
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    float4 *g4 = reinterpret_cast<float4 *>(P.grad_local);

    // ---- phase 1: push my contribution to every other owner, clearing what has been sent.  Work items = (piece of blockDim
    // float4, owner) with the owner varying fastest, so at every moment the grid's stores are spread evenly over all W - 1
    // destinations whatever the skew between the ranks (a schedule "everybody sends to rank + k" is a permutation only while the
    // ranks stay in lockstep).  Measured at 8 GPUs: the same 0.404 ms/step either way -- the all-to-all push of 50 MB per GPU
    // takes ~125 us (400 GB/s per GPU; a single pair reaches 665 GB/s) however it is ordered.
    if (world > 1) {
        const uint64_t max_len = (P.n4 + world - 1) / world, pieces = (max_len + blockDim.x - 1) / blockDim.x;
        const uint64_t items = pieces * (world - 1u);
#pragma unroll 1
        for (uint64_t id0 = blockIdx.x; id0 < items; id0 += 2ull * gridDim.x) {
            float4 v[2];
            float4 *src[2], *dst[2];
            bool ok[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const uint64_t id = id0 + (uint64_t)u * gridDim.x;
                const uint32_t k = 1u + (uint32_t)(id % (world - 1u));
                const uint32_t w = (P.rank + k) % world;
                uint64_t a, b;
                slice_of(P.n4, w, world, a, b);
                const uint64_t off = (id / (world - 1u)) * blockDim.x + threadIdx.x;
                ok[u] = id < items && a + off < b;
                src[u] = g4 + a + off;
                dst[u] = reinterpret_cast<float4 *>(P.stage[w]) + (uint64_t)P.rank * P.slot4 + off;
                if (ok[u]) v[u] = *src[u];
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (ok[u]) { st_peer(dst[u], v[u]); *src[u] = make_float4(0.f, 0.f, 0.f, 0.f); }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(P.flags[P.rank] + NAFB_XFLAG_TICKET, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x == 0) P.flags[P.rank][NAFB_XFLAG_TICKET] = 0u;
        if (threadIdx.x < world) st_release_sys(P.flags[threadIdx.x] + NAFB_XFLAG_ARRIVE + P.rank, epoch);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) stamps[1] = globaltimer_ns();
    wait_all(P, NAFB_XFLAG_ARRIVE, epoch);
    if (blockIdx.x == 0 && threadIdx.x == 0) stamps[2] = globaltimer_ns();

    // ---- phase 2: my slice
    float4 *m4 = reinterpret_cast<float4 *>(P.m), *v4 = reinterpret_cast<float4 *>(P.v);
    const float4 *p_in = reinterpret_cast<const float4 *>(P.param[P.rank]);
    const float4 *st = reinterpret_cast<const float4 *>(P.stage[P.rank]);
    for (uint64_t i = P.s0 + gid; i < P.s1; i += stride) {
        float4 g[MAXR];
#pragma unroll
        for (int w = 0; w < MAXR; ++w)
            if (w < (int)world) g[w] = (uint32_t)w == P.rank ? g4[i] : __ldcs(st + (uint64_t)w * P.slot4 + (i - P.s0));
        float4 p = p_in[i], m = m4[i - P.s0], v = v4[i - P.s0];
        float4 s = g[0];
#pragma unroll
        for (int w = 1; w < MAXR; ++w)
            if (w < (int)world) { s.x += g[w].x; s.y += g[w].y; s.z += g[w].z; s.w += g[w].w; }
        adam_one(p.x, s.x, m.x, v.x, c);
        adam_one(p.y, s.y, m.y, v.y, c);
        adam_one(p.z, s.z, m.z, v.z, c);
        adam_one(p.w, s.w, m.w, v.w, c);
        m4[i - P.s0] = m;
        v4[i - P.s0] = v;
        g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < MAXR; ++w)
            if (w < (int)world) st_peer(reinterpret_cast<float4 *>(P.param[w]) + i, p);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) stamps[3] = globaltimer_ns();
    // ---- last block out: "done" to every rank, then wait until every rank is done with MY replica and MY staging area
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(P.flags[P.rank] + NAFB_XFLAG_TICKET2, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (threadIdx.x == 0) P.flags[P.rank][NAFB_XFLAG_TICKET2] = 0u;
    if (threadIdx.x < world) st_release_sys(P.flags[threadIdx.x] + NAFB_XFLAG_DONE + P.rank, epoch);
    wait_all(P, NAFB_XFLAG_DONE, epoch);
    if (P.state && threadIdx.x == 0) P.state[NAFB_STATE_STEP] = epoch;
}

}  // namespace

extern "C" {

int nafb_peer_alloc(uint64_t bytes, void **ptr, unsigned char *handle64) {
    if (!ptr || !handle64 || bytes == 0) NAFB_FAIL(NAFB_ERR_INVALID, "peer_alloc: bad argument");
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) NAFB_FAIL(NAFB_ERR_CUDA, "peer_alloc: cudaMalloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
    e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) { cudaFree(p); NAFB_FAIL(NAFB_ERR_CUDA, "peer_alloc: cudaMemset: %s", cudaGetErrorString(e)); }
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); NAFB_FAIL(NAFB_ERR_CUDA, "peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return NAFB_OK;
}

int nafb_peer_open(const unsigned char *handle64, void **ptr) {
    if (!ptr || !handle64) NAFB_FAIL(NAFB_ERR_INVALID, "peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) NAFB_FAIL(NAFB_ERR_CUDA, "peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    *ptr = p;
    return NAFB_OK;
}

int nafb_peer_close(void *ptr) {
    if (!ptr) return NAFB_OK;
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) NAFB_FAIL(NAFB_ERR_CUDA, "peer_close: %s", cudaGetErrorString(e));
    return NAFB_OK;
}

int nafb_peer_free(void *ptr) {
    if (!ptr) return NAFB_OK;
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) NAFB_FAIL(NAFB_ERR_CUDA, "peer_free: %s", cudaGetErrorString(e));
    return NAFB_OK;
}

int nafb_exchange_slice(uint64_t n, uint32_t rank, uint32_t world, uint64_t *i0, uint64_t *i1) {
    if (!i0 || !i1 || world == 0 || rank >= world || (n & 3)) NAFB_FAIL(NAFB_ERR_INVALID, "exchange_slice: bad argument (n must be a multiple of 4)");
    const uint64_t n4 = n >> 2, base = n4 / world, extra = n4 % world;
    const uint64_t a = rank * base + (rank < extra ? rank : extra);
    *i0 = a << 2;
    *i1 = (a + base + (rank < extra ? 1 : 0)) << 2;
    return NAFB_OK;
}

int nafb_adam_exchange_step(const nafb_exchange *x, double lr, double beta1, double beta2, double eps, uint32_t step, float grad_scale,
                            nafb_stream_t stream) {
    if (!x) NAFB_FAIL(NAFB_ERR_INVALID, "adam_exchange_step: null descriptor");
    if (x->world < 1 || x->world > NAFB_MAX_RANKS || x->rank >= x->world) NAFB_FAIL(NAFB_ERR_INVALID, "adam_exchange_step: world %u / rank %u", x->world, x->rank);
    if (step == 0 && !x->state) NAFB_FAIL(NAFB_ERR_INVALID, "adam_exchange_step: step is 1-based");
    if (x->n == 0 || (x->n & 3)) NAFB_FAIL(NAFB_ERR_INVALID, "adam_exchange_step: n must be a positive multiple of 4");
    if (!x->exp_avg || !x->exp_avg_sq) NAFB_FAIL(NAFB_ERR_INVALID, "adam_exchange_step: null optimizer state");
    ExchangeParams P{};
    P.world = x->world; P.rank = x->rank; P.epoch = step;
    uintptr_t align = (uintptr_t)x->exp_avg | (uintptr_t)x->exp_avg_sq | (uintptr_t)x->grad_zero;
    for (uint32_t w = 0; w < x->world; ++w) {
        if (!x->param[w] || !x->grad[w] || !x->flags[w]) NAFB_FAIL(NAFB_ERR_INVALID, "adam_exchange_step: null pointer for rank %u", w);
        P.param[w] = x->param[w]; P.grad[w] = x->grad[w]; P.flags[w] = x->flags[w];
        align |= (uintptr_t)x->param[w] | (uintptr_t)x->grad[w];
    }
    if (align & 15) NAFB_FAIL(NAFB_ERR_INVALID, "adam_exchange_step: buffers must be 16-byte aligned");
    P.grad_zero = x->grad_zero; P.m = x->exp_avg; P.v = x->exp_avg_sq;
    uint64_t i0 = 0, i1 = 0;
    int rc = nafb_exchange_slice(x->n, x->rank, x->world, &i0, &i1);
    if (rc) return rc;
    P.n4 = x->n >> 2; P.s0 = i0 >> 2; P.s1 = i1 >> 2;
    P.c = make_adam_const(lr, beta1, beta2, eps, step ? step : 1u, grad_scale);
    P.state = x->state; P.beta1 = beta1; P.beta2 = beta2; P.eps = eps; P.gscale = grad_scale;
    if ((x->mc_param != nullptr) != (x->mc_grad != nullptr)) NAFB_FAIL(NAFB_ERR_INVALID, "adam_exchange_step: mc_param and mc_grad go together");
    if (((uintptr_t)x->mc_param | (uintptr_t)x->mc_grad) & 15) NAFB_FAIL(NAFB_ERR_INVALID, "adam_exchange_step: multicast addresses must be 16-byte aligned");
    P.mc_param = x->mc_param; P.mc_grad = x->mc_grad;
    const bool push = x->stage[x->rank] != nullptr;
    if (push) {
        for (uint32_t w = 0; w < x->world; ++w) {
            if (!x->stage[w]) NAFB_FAIL(NAFB_ERR_INVALID, "adam_exchange_step: staging area of rank %u is null", w);
            if ((uintptr_t)x->stage[w] & 15) NAFB_FAIL(NAFB_ERR_INVALID, "adam_exchange_step: staging areas must be 16-byte aligned");
            P.stage[w] = x->stage[w];
        }
        P.grad_local = x->grad[x->rank];
        P.slot4 = x->stage_slot >> 2;
        if ((x->stage_slot & 3) || P.slot4 < (P.n4 + x->world - 1) / x->world) NAFB_FAIL(NAFB_ERR_INVALID, "adam_exchange_step: stage_slot too small");
    }
    // persistent grid: as many blocks of 512 threads as are resident at once (blocks spin on the arrival flags)
    cudaStream_t s = (cudaStream_t)stream;
    auto launch = [&](auto kernel) {
        int per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 512, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        kernel<<<(unsigned)(nafb_sm_count() * per_sm), 512, 0, s>>>(P);
    };
    if (push) {
        switch (x->world) {
            case 2: launch(k_adam_exchange_push<2>); break;
            case 4: launch(k_adam_exchange_push<4>); break;
            case 8: launch(k_adam_exchange_push<8>); break;
            default: launch(k_adam_exchange_push<0>); break;
        }
    } else {
        switch (x->world) {
            case 2: launch(k_adam_exchange<2>); break;
            case 4: launch(k_adam_exchange<4>); break;
            case 8: launch(k_adam_exchange<8>); break;
            default: launch(k_adam_exchange<0>); break;
        }
    }
    NAFB_CHECK_LAUNCH("adam_exchange_step");
    return NAFB_OK;
}

}  // extern "C"
