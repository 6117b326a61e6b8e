// tmem_probe2.cu -- does tcgen05.dealloc wait for the tensor-core work of the OTHER CTA on the SM (or for anything else)?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../neuralvolumetricreconstructionformedicalimages_b200/csrc/umma.cuh"
__global__ void __launch_bounds__(256) k(long long *out, int iters_even, int iters_odd, int mmas, int relinquish_late) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 16384 / 16; i += 256) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { umma::mbar_init(&mbar, 1); umma::fence_mbar_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma::smem_u32(&slot)), "r"(256) : "memory");
        if (!relinquish_late) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = slot;
    const int iters = (blockIdx.x / 148) & 1 ? iters_odd : iters_even;      // the two CTAs of an SM get different lengths
    uint32_t phase = 0;
    constexpr uint32_t IDESC = umma::idesc_bf16(128, 32, 0, 0);
    for (int it = 0; it < iters; ++it) {
        if (threadIdx.x == 0) {
            const uint64_t a = umma::make_desc(umma::smem_u32(smem), 128, 512), b = umma::make_desc(umma::smem_u32(smem) + 8192, 128, 512);
            for (int m = 0; m < mmas; ++m) umma::mma_bf16(tmem, a, b, IDESC, m > 0);
            umma::commit(&mbar);
        }
        umma::mbar_wait(&mbar, phase);
        phase ^= 1;
        umma::fence_after_sync();
        float v[16];
        umma::tmem_ld16(tmem + ((uint32_t)((threadIdx.x >> 5 & 3) * 32) << 16), v);
        umma::tmem_wait_ld();
        if (v[0] == 123.f) out[0] = 1;
        umma::fence_before_sync();
        __syncthreads();
    }
    unsigned long long t2, t3;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t2)::"memory");
    if (threadIdx.x < 32) {
        if (relinquish_late) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    }
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t3)::"memory");
    if (threadIdx.x == 0) out[1 + blockIdx.x] = (long long)(t3 - t2);
}
void run(const char *name, int grid, int ie, int io, int mmas, int late) {
    long long *d; cudaMalloc(&d, (grid + 1) * sizeof(long long));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    for (int rep = 0; rep < 2; ++rep) k<<<grid, 256, 100000>>>(d, ie, io, mmas, late);
    cudaError_t e = cudaDeviceSynchronize();
    long long *h = new long long[grid + 1];
    cudaMemcpy(h, d, (grid + 1) * sizeof(long long), cudaMemcpyDeviceToHost);
    double se = 0, so = 0; long long me = 0, mo = 0; int ne = 0, no = 0;
    for (int i = 0; i < grid; ++i) { long long v = h[1 + i]; if ((i / 148) & 1) { so += v; no++; if (v > mo) mo = v; } else { se += v; ne++; if (v > me) me = v; } }
    printf("%-50s dealloc cycles: first CTA of the SM mean %8.0f max %8lld | second mean %8.0f max %8lld (%s)\n", name, se / (ne ? ne : 1), me, so / (no ? no : 1), mo, cudaGetErrorString(e));
    cudaFree(d); delete[] h;
}
int main() {
    run("1 CTA/SM, 100 x 8 MMAs", 148, 100, 100, 8, 0);
    run("2 CTAs/SM, equal 100 x 8 MMAs", 296, 100, 100, 8, 0);
    run("2 CTAs/SM, 20 vs 400 x 8 MMAs", 296, 20, 400, 8, 0);
    run("2 CTAs/SM, 20 vs 400 x 32 MMAs", 296, 20, 400, 32, 0);
    run("2 CTAs/SM, 20 vs 400 x 8, relinquish at the end", 296, 20, 400, 8, 1);
    return 0;
}
