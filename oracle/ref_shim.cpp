// ref_shim.cpp -- host-side harness that compiles the REFERENCE's own kernel text
// (src/encoder/hashencoder/src/hashencoder.cu lines 30-298: div_round_up, fast_hash,
// get_grid_index, kernel_grid, kernel_grid_backward, kernel_input_backward) as plain
// C++ so that the restatement in nafb_oracle.c can be pinned against it without a GPU.
//
// TEST INFRASTRUCTURE ONLY.  The reference source is NOT copied into this repo: the
// build recipe (oracle/build_ref.sh) extracts the line range into a temporary file
// outside the repo and passes its path as -DREF_KERNEL_TEXT="...".  The resulting
// binary lives in oracle/_ref/ (git-ignored).
//
// The shim supplies what the CUDA dialect needs on the host: the __global__ /
// __device__ / __host__ markers, blockIdx / threadIdx / blockDim, at::Half, __half,
// __half2 (only so the dead half branch at :257-263 parses) and a host atomicAdd.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <type_traits>

#define __global__
#define __device__
#define __host__
#ifndef __restrict__
#define __restrict__ __restrict
#endif

struct shim_dim3 { uint32_t x = 1, y = 1, z = 1; };
static thread_local shim_dim3 blockIdx, threadIdx, blockDim;

struct __half {
    float v;
    __half() : v(0) {}
    __half(float f) : v(f) {}
    operator float() const { return v; }
};
struct __half2 { __half x, y; };
namespace at { using Half = __half; }

template <typename T>
static inline T atomicAdd(T *addr, T val) {
    T old;
#pragma omp atomic capture
    { old = *addr; *addr += val; }
    return old;
}
static inline __half2 atomicAdd(__half2 *addr, __half2 val) {
    __half2 old = *addr;
    addr->x = __half(float(addr->x) + float(val.x));
    addr->y = __half(float(addr->y) + float(val.y));
    return old;
}

#include REF_KERNEL_TEXT

// ---- C entry points: emulate the <<<grid, block>>> launches of hashencoder.cu:301-353 ----
template <uint32_t D, uint32_t C>
static void run_fwd(const float *inputs, const float *grid, const int *offsets, float *outputs,
                    uint32_t B, uint32_t L, uint32_t H, bool cgi, float *dy_dx) {
    const uint32_t NT = 512;                       // hashencoder.cu:303
    const uint32_t nbx = div_round_up(B, NT);
#pragma omp parallel for collapse(2) schedule(static)
    for (uint32_t by = 0; by < L; ++by)
        for (uint32_t bx = 0; bx < nbx; ++bx) {
            blockDim.x = NT; blockIdx.x = bx; blockIdx.y = by;
            for (uint32_t t = 0; t < NT; ++t) {
                threadIdx.x = t;
                kernel_grid<float, D, C>(inputs, grid, offsets, outputs, B, L, H, cgi, dy_dx);
            }
        }
}

template <uint32_t D, uint32_t C, uint32_t N_C>
static void run_bwd(const float *grad, const float *inputs, const float *grid, const int *offsets,
                    float *grad_grid, uint32_t B, uint32_t L, uint32_t H, int ordered) {
    const uint32_t NT = 256;                       // hashencoder.cu:331
    const uint32_t nbx = div_round_up(B * C / N_C, NT);
    if (ordered) {
        // deterministic: levels in parallel (disjoint table ranges), points in order
#pragma omp parallel for schedule(dynamic, 1)
        for (uint32_t by = 0; by < L; ++by)
            for (uint32_t bx = 0; bx < nbx; ++bx) {
                blockDim.x = NT; blockIdx.x = bx; blockIdx.y = by;
                for (uint32_t t = 0; t < NT; ++t) {
                    threadIdx.x = t;
                    kernel_grid_backward<float, D, C, N_C>(grad, inputs, grid, offsets, grad_grid, B, L, H);
                }
            }
    } else {
#pragma omp parallel for collapse(2) schedule(static)
        for (uint32_t by = 0; by < L; ++by)
            for (uint32_t bx = 0; bx < nbx; ++bx) {
                blockDim.x = NT; blockIdx.x = bx; blockIdx.y = by;
                for (uint32_t t = 0; t < NT; ++t) {
                    threadIdx.x = t;
                    kernel_grid_backward<float, D, C, N_C>(grad, inputs, grid, offsets, grad_grid, B, L, H);
                }
            }
    }
}

extern "C" {

// returns 0 on success, 1 for an unsupported (D, C) (hashencoder.cu:310,324 throw)
int ref_hash_forward(const float *inputs, const float *grid, const int *offsets, float *outputs,
                     uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t H,
                     int calc_grad_inputs, float *dy_dx) {
    const bool cgi = calc_grad_inputs != 0;
#define FWD(DD, CC) if (D == DD && C == CC) { run_fwd<DD, CC>(inputs, grid, offsets, outputs, B, L, H, cgi, dy_dx); return 0; }
    FWD(2, 1) FWD(2, 2) FWD(2, 4) FWD(2, 8) FWD(3, 1) FWD(3, 2) FWD(3, 4) FWD(3, 8)
#undef FWD
    return 1;
}

int ref_hash_backward(const float *grad, const float *inputs, const float *grid, const int *offsets,
                      float *grad_grid, uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t H,
                      int ordered) {
#define BWD(DD, CC, NC) if (D == DD && C == CC) { run_bwd<DD, CC, NC>(grad, inputs, grid, offsets, grad_grid, B, L, H, ordered); return 0; }
    BWD(2, 1, 1) BWD(2, 2, 2) BWD(2, 4, 2) BWD(2, 8, 2) BWD(3, 1, 1) BWD(3, 2, 2) BWD(3, 4, 2) BWD(3, 8, 2)
#undef BWD
    return 1;
}

int ref_input_backward(const float *grad, const float *dy_dx, float *grad_inputs,
                       uint32_t B, uint32_t D, uint32_t C, uint32_t L) {
    const uint32_t NT = 256;
    const uint32_t nb = div_round_up(B * D, NT);
    for (uint32_t bx = 0; bx < nb; ++bx) {
        blockDim.x = NT; blockIdx.x = bx; blockIdx.y = 0;
        for (uint32_t t = 0; t < NT; ++t) {
            threadIdx.x = t;
#define IB(DD, CC) if (D == DD && C == CC) kernel_input_backward<float, DD, CC>(grad, dy_dx, grad_inputs, B, L);
            IB(2, 1) IB(2, 2) IB(2, 4) IB(2, 8) IB(3, 1) IB(3, 2) IB(3, 4) IB(3, 8)
#undef IB
        }
    }
    return 0;
}

// get_grid_index<3,C>(ch=0)/C for one lattice point (KAT table)
uint32_t ref_grid_index_3(uint32_t C, uint32_t hashmap_size, uint32_t resolution,
                          uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t p[3] = {x, y, z};
    switch (C) {
        case 1: return get_grid_index<3, 1>(0, hashmap_size, resolution, p) / 1;
        case 2: return get_grid_index<3, 2>(0, hashmap_size, resolution, p) / 2;
        case 4: return get_grid_index<3, 4>(0, hashmap_size, resolution, p) / 4;
        default: return get_grid_index<3, 8>(0, hashmap_size, resolution, p) / 8;
    }
}

}  // extern "C"
