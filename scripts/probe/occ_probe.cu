// occ_probe.cu -- which kernel property caps cudaOccupancyMaxActiveBlocksPerMultiprocessor at 1 on this B200?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *m, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(m)), "r"(c) : "memory"); }
__global__ void k_plain(float *o) { o[threadIdx.x] = 1.f; }
template <int N> __global__ void k_mbar(float *o) {
    __shared__ uint64_t mb[N];
    if (threadIdx.x == 0) for (int i = 0; i < N; ++i) mbar_init(&mb[i], 1);
    __syncthreads();
    o[threadIdx.x] = (float)mb[0];
}
__global__ void k_bulk(float *o, const float *src) {
    __shared__ __align__(128) float buf[256];
    __shared__ uint64_t mb;
    if (threadIdx.x == 0) {
        mbar_init(&mb, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mb)), "r"(1024) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf)), "l"(src), "r"(1024), "r"(smem_u32(&mb)) : "memory");
    }
    __syncthreads();
    o[threadIdx.x] = buf[threadIdx.x];
}
template <int COLS> __global__ void k_tmem(float *o) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    o[threadIdx.x] = (float)slot;
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(COLS) : "memory");
}
template <int COLS, int NMB> __global__ void k_tmem_mbar(float *o) {
    __shared__ uint32_t slot;
    __shared__ uint64_t mb[NMB];
    if (threadIdx.x == 0) for (int i = 0; i < NMB; ++i) mbar_init(&mb[i], 1);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    o[threadIdx.x] = (float)slot + (float)mb[0];
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(COLS) : "memory");
}
__global__ void k_bar4(float *o) {
    asm volatile("bar.sync 1, 256;" ::: "memory");
    asm volatile("bar.sync 2, 256;" ::: "memory");
    asm volatile("bar.sync 3, 256;" ::: "memory");
    o[threadIdx.x] = 1.f;
}
__global__ void __launch_bounds__(512, 2) k_lb(float *o) { o[threadIdx.x] = 1.f; }
#define Q(k) do { int n = -1; cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, 256, 0); printf("%-28s occupancy(256 thr, 0 smem) = %d (%s)\n", #k, n, cudaGetErrorString(e)); } while (0)
int main() {
    Q(k_plain); Q(k_mbar<1>); Q(k_mbar<2>); Q(k_mbar<4>); Q(k_mbar<8>); Q(k_mbar<10>); Q(k_mbar<16>); Q(k_bulk);
    Q(k_tmem<32>); Q(k_tmem<128>); Q(k_tmem<256>); Q(k_tmem<512>); Q((k_tmem_mbar<256, 1>)); Q((k_tmem_mbar<256, 10>)); Q(k_bar4); Q(k_lb);
    return 0;
}
