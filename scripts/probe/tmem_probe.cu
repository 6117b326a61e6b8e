// tmem_probe.cu -- how long do tcgen05.alloc / tcgen05.dealloc take, alone and with a second CTA on the SM?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int COLS, bool RELINQUISH>
__global__ void __launch_bounds__(256) k(long long *out, int spin_cycles) {
    __shared__ uint32_t slot;
    extern __shared__ uint8_t pad[];
    long long t0 = clock64();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(COLS) : "memory");
        if (RELINQUISH) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t1 = clock64();
    while (clock64() - t1 < spin_cycles) {}
    __syncthreads();
    long long t2 = clock64();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(COLS) : "memory");
    long long t3 = clock64();
    if (threadIdx.x == 0) { out[3 * blockIdx.x] = t1 - t0; out[3 * blockIdx.x + 1] = t3 - t2; out[3 * blockIdx.x + 2] = pad[0]; }
}
template <int COLS, bool REL> void run(const char *name, int grid, int smem, int spin) {
    long long *d; cudaMalloc(&d, grid * 3 * sizeof(long long));
    cudaFuncSetAttribute(k<COLS, REL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep) k<COLS, REL><<<grid, 256, smem>>>(d, spin);
    cudaError_t e = cudaDeviceSynchronize();
    long long *h = new long long[grid * 3];
    cudaMemcpy(h, d, grid * 3 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long amax = 0, dmax = 0; double as = 0, ds = 0;
    for (int i = 0; i < grid; ++i) { as += h[3 * i]; ds += h[3 * i + 1]; if (h[3 * i] > amax) amax = h[3 * i]; if (h[3 * i + 1] > dmax) dmax = h[3 * i + 1]; }
    printf("%-44s grid %4d smem %6d: alloc mean %8.0f max %8lld cyc | dealloc mean %8.0f max %8lld cyc  (%s)\n", name, grid, smem, as / grid, amax, ds / grid, dmax, cudaGetErrorString(e));
    cudaFree(d); delete[] h;
}
int main() {
    run<256, true>("256 cols, 1 CTA/SM", 148, 120000, 20000);
    run<256, true>("256 cols, 2 CTAs/SM", 296, 100000, 20000);
    run<256, true>("256 cols, 2 CTAs/SM, long body", 296, 100000, 200000);
    run<32, true>("32 cols, 3 CTAs/SM", 444, 50000, 20000);
    run<32, true>("32 cols, 1 CTA/SM", 148, 120000, 20000);
    run<512, true>("512 cols, 1 CTA/SM", 148, 120000, 20000);
    run<256, false>("256 cols, 2 CTAs/SM, no relinquish", 296, 100000, 20000);
    run<256, true>("256 cols, 4 waves of 2 CTAs/SM", 296 * 4, 100000, 20000);
    return 0;
}
