// stream_probe.cu -- is the optimizer kernel's distance from the copy bandwidth a matter of HOW MANY streams it touches?
// In-place read-modify-write of 228 MB as (a) four 57 MB arrays walked together (the Adam pattern: p, g, m, v), (b) two arrays of
// 114 MB, (c) one array of 228 MB; (d) out-of-place copy of 228 MB (the "copy peak" pattern).  Trivial arithmetic.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

template <int NA>
__global__ void __launch_bounds__(256) k_rmw(float4 *a0, float4 *a1, float4 *a2, float4 *a3, uint64_t n4) {
    float4 *a[4] = {a0, a1, a2, a3};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        float4 v[NA];
#pragma unroll
        for (int k = 0; k < NA; ++k) v[k] = a[k][i];
#pragma unroll
        for (int k = 0; k < NA; ++k) { v[k].x += 1.f; v[k].y += 1.f; v[k].z += 1.f; v[k].w += 1.f; a[k][i] = v[k]; }
    }
}
__global__ void __launch_bounds__(256) k_copy(const float4 *__restrict__ src, float4 *__restrict__ dst, uint64_t n4) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main() {
    const uint64_t n = 14266668ull, n4 = n / 4;            // floats per Adam vector
    float4 *buf, *dst;
    CK(cudaMalloc(&buf, 4 * n4 * sizeof(float4)));
    CK(cudaMalloc(&dst, 4 * n4 * sizeof(float4)));
    CK(cudaMemset(buf, 0, 4 * n4 * sizeof(float4)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = 148 * 8, reps = 50;
    for (int variant = 0; variant < 4; ++variant) {
        float ms = 0;
        for (int pass = 0; pass < 2; ++pass) {
            CK(cudaEventRecord(e0));
            for (int r = 0; r < reps; ++r) {
                if (variant == 0) k_rmw<4><<<grid, 256>>>(buf, buf + n4, buf + 2 * n4, buf + 3 * n4, n4);
                else if (variant == 1) k_rmw<2><<<grid, 256>>>(buf, buf + 2 * n4, nullptr, nullptr, 2 * n4);
                else if (variant == 2) k_rmw<1><<<grid, 256>>>(buf, nullptr, nullptr, nullptr, 4 * n4);
                else k_copy<<<grid, 256>>>(buf, dst, 4 * n4);
            }
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
            CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        const char *names[] = {"in-place, 4 arrays of 57 MB", "in-place, 2 arrays of 114 MB", "in-place, 1 array of 228 MB", "copy 228 MB -> 228 MB"};
        const double bytes = 2.0 * 4 * n4 * sizeof(float4);
        printf("%-30s %.1f us per launch, %.0f GB/s (read + write)\n", names[variant], ms / reps * 1e3, bytes * reps / (ms * 1e-3) / 1e9);
    }
    return 0;
}
