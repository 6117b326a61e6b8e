// nvlink_store_probe.cu -- how fast can SM-issued traffic cross NVLink from GPU 0 to GPU 1 (and both ways at once)?
// One process, two devices with peer access.  Variants of a 57 MB copy local -> peer:
//   st16     ld.global + st.global.cg.v4.f32 (what the exchange kernel issues), one element per thread and iteration
//   st16x4   the same, four elements in flight per thread
//   st32     256-bit stores (st.global.v8.f32)
//   bulk     cp.async.bulk global -> shared (local), cp.async.bulk shared -> global (peer): the TMA unit moves 16 KB pieces
//   memcpy   cudaMemcpyPeerAsync (copy engines), for reference
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o nvlink_store_probe nvlink_store_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(512) k_st16(const float4 *__restrict__ src, float4 *__restrict__ dst, uint64_t n4) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = gid; i < n4; i += stride) __stcg(dst + i, src[i]);
}
__global__ void __launch_bounds__(512) k_st16x4(const float4 *__restrict__ src, float4 *__restrict__ dst, uint64_t n4) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = gid; i < n4; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i + u * stride < n4) v[u] = src[i + u * stride];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i + u * stride < n4) __stcg(dst + i + u * stride, v[u]);
    }
}
__global__ void __launch_bounds__(512) k_st32(const float4 *__restrict__ src, float4 *__restrict__ dst, uint64_t n4) {
    const uint64_t n8 = n4 / 2;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = gid; i < n8; i += stride) {
        const float4 a = src[2 * i], b = src[2 * i + 1];
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 2 * i), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x),
                     "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr uint32_t PIECE = 16384;
// one CTA = 2 x 16 KB staging buffers; thread 0 drives the TMA unit
__global__ void __launch_bounds__(32) k_bulk(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint64_t bytes) {
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ __align__(8) uint64_t full[2];
    if (threadIdx.x != 0) return;
    for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint64_t pieces = bytes / PIECE;
    uint32_t phase[2] = {0, 0};
    uint64_t k = 0;
    for (uint64_t p = blockIdx.x; p < pieces; p += gridDim.x, ++k) {
        const int b = (int)(k & 1);
        if (k >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // the store that last read buffer b has read it
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[b])), "r"(PIECE) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf + b * PIECE)),
                     "l"(src + p * PIECE), "r"(PIECE), "r"(smem_u32(&full[b])) : "memory");
        uint32_t done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&full[b])), "r"(phase[b]) : "memory");
        }
        phase[b] ^= 1;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + p * PIECE), "r"(smem_u32(buf + b * PIECE)), "r"(PIECE) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct Dev { int id; cudaStream_t s; uint8_t *src, *dst_on_peer; cudaEvent_t e0, e1; };

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    const uint64_t bytes = 57u << 20;
    Dev d[2];
    uint8_t *recv[2];
    for (int i = 0; i < 2; ++i) {
        CK(cudaSetDevice(i));
        CK(cudaDeviceEnablePeerAccess(1 - i, 0));
        CK(cudaStreamCreate(&d[i].s));
        CK(cudaMalloc(&d[i].src, bytes));
        CK(cudaMalloc(&recv[i], bytes));
        CK(cudaMemset(d[i].src, i + 1, bytes));
        CK(cudaEventCreate(&d[i].e0));
        CK(cudaEventCreate(&d[i].e1));
        d[i].id = i;
    }
    for (int i = 0; i < 2; ++i) d[i].dst_on_peer = recv[1 - i];
    CK(cudaSetDevice(0));
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * PIECE));
    CK(cudaSetDevice(1));
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * PIECE));
    const char *names[] = {"st16", "st16x4", "st32", "bulk x1/SM", "bulk x4/SM", "bulk x6/SM", "memcpyPeer"};
    for (int variant = 0; variant < 7; ++variant) {
        for (int both = 0; both < 2; ++both) {
            const int reps = 20;
            float ms[2] = {0, 0};
            for (int pass = 0; pass < 2; ++pass) {   // pass 0 = warm-up
                for (int i = 0; i <= both; ++i) { CK(cudaSetDevice(i)); CK(cudaEventRecord(d[i].e0, d[i].s)); }
                for (int r = 0; r < reps; ++r)
                    for (int i = 0; i <= both; ++i) {
                        CK(cudaSetDevice(i));
                        const uint64_t n4 = bytes / 16;
                        const float4 *s4 = (const float4 *)d[i].src;
                        float4 *t4 = (float4 *)d[i].dst_on_peer;
                        switch (variant) {
                            case 0: k_st16<<<148 * 4, 512, 0, d[i].s>>>(s4, t4, n4); break;
                            case 1: k_st16x4<<<148 * 4, 512, 0, d[i].s>>>(s4, t4, n4); break;
                            case 2: k_st32<<<148 * 4, 512, 0, d[i].s>>>(s4, t4, n4); break;
                            case 3: k_bulk<<<148, 32, 2 * PIECE, d[i].s>>>(d[i].src, d[i].dst_on_peer, bytes); break;
                            case 4: k_bulk<<<148 * 4, 32, 2 * PIECE, d[i].s>>>(d[i].src, d[i].dst_on_peer, bytes); break;
                            case 5: k_bulk<<<148 * 6, 32, 2 * PIECE, d[i].s>>>(d[i].src, d[i].dst_on_peer, bytes); break;
                            case 6: CK(cudaMemcpyPeerAsync(d[i].dst_on_peer, 1 - i, d[i].src, i, bytes, d[i].s)); break;
                        }
                    }
                for (int i = 0; i <= both; ++i) { CK(cudaSetDevice(i)); CK(cudaEventRecord(d[i].e1, d[i].s)); }
                for (int i = 0; i <= both; ++i) { CK(cudaSetDevice(i)); CK(cudaStreamSynchronize(d[i].s)); CK(cudaGetLastError()); CK(cudaEventElapsedTime(&ms[i], d[i].e0, d[i].e1)); }
            }
            printf("%-12s %s: GPU0 -> GPU1 %.0f GB/s", names[variant], both ? "both directions" : "one direction  ", bytes * reps / (ms[0] * 1e-3) / 1e9);
            if (both) printf(", GPU1 -> GPU0 %.0f GB/s", bytes * reps / (ms[1] * 1e-3) / 1e9);
            printf("\n");
        }
    }
    // verify the last variant's data landed
    CK(cudaSetDevice(1));
    uint8_t h[4];
    CK(cudaMemcpy(h, recv[1] + bytes - 4, 4, cudaMemcpyDeviceToHost));
    printf("check: last bytes on GPU1 = %d %d (expect 1 1)\n", h[0], h[3]);
    return 0;
}
