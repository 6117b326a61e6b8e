// umma.cuh -- hand-written tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX).
//
// Operand convention used throughout this library
// -----------------------------------------------
// Every matrix operand lives in shared memory in the canonical NO-SWIZZLE ("interleave") UMMA
// layout of 16-bit elements: 8x8-element core matrices of 128 contiguous bytes
// (8 rows x 16 B), addressed as
//
//      byte(row, col) = (row / 8) * SBO + (col / 8) * LBO + (row % 8) * 16 + (col % 8) * 2
//
// Read as a K-MAJOR operand (rows = M or N, cols = K) the descriptor carries (LBO, SBO) as is.
// The very same bytes are a valid MN-MAJOR operand of the transposed matrix (rows = K,
// cols = M or N) with the two offsets swapped -- which is how the backward pass reuses the
// activation tiles (written once, row = sample point) for  dX = G.W  and  dW = G^T.X  without
// any transposed copy.
//
// fp32 values are carried as a bf16 pair (hi = bf16(x), lo = bf16(x - hi)); a product is
// evaluated as  hi*hi + hi*lo + lo*hi  with fp32 accumulation in TMEM ("bf16x3": ~2^-16
// relative error per product instead of bf16's 2^-8).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (64 bit), SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
__device__ __forceinline__ uint64_t advance_desc(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// ---- instruction descriptor, kind::f16 with bf16 operands and fp32 accumulation
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
    return (1u << 4)            // c_format  = F32
           | (1u << 7)          // a_format  = BF16
           | (1u << 10)         // b_format  = BF16
           | (a_mn_major << 15) // a_major   (0 = K, 1 = MN)
           | (b_mn_major << 16) // b_major
           | ((N >> 3) << 17)   // n_dim
           | ((M >> 4) << 24);  // m_dim
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void commit(uint64_t *mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *mbar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(mbar)),
        "r"(parity)
        : "memory");
}

// long waits (a consumer role waiting for a producer role): back off between polls so that the polling warps do not
// compete with the working warps for issue slots and the shared-memory pipe
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *mbar, uint32_t parity, unsigned ns = 64) {
    const uint32_t a = smem_u32(mbar);
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) return;
        __nanosleep(ns);
    }
}

// ---- named barriers (producer / consumer hand-off between warp roles): `threads` = arriving + syncing threads
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// the same with the barrier id as an immediate: ptxas then reserves only the ids that are really used
template <int ID>
__device__ __forceinline__ void named_bar_sync_imm(uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(threads) : "memory");
}
template <int ID>
__device__ __forceinline__ void named_bar_arrive_imm(uint32_t threads) {
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(threads) : "memory");
}

// ---- tensor memory
// one full warp: allocate `ncols` (power of two >= 32) columns, base address written to *slot (smem)
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}

// warp-collective: this thread's TMEM lane (32*(warp%4) + lane), 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// N consecutive 32-bit columns (N = 1, 2, 4, 8) of this thread's TMEM lane
template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[N]) {
    static_assert(N == 1 || N == 2 || N == 4 || N == 8, "tmem_ldn: N must be 1, 2, 4 or 8");
    uint32_t r[8];
    if constexpr (N == 1)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r[0]) : "r"(taddr));
    else if constexpr (N == 2)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
    else if constexpr (N == 4)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    else
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- canonical layout + bf16x2 split stores
__device__ __forceinline__ uint32_t canon_off(uint32_t row, uint32_t chunk, uint32_t lbo, uint32_t sbo) {
    return (row >> 3) * sbo + chunk * lbo + (row & 7u) * 16u;
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// 8 floats -> one 16-byte chunk of the hi matrix and one of the lo matrix (packed conversions: two values per cvt)
__device__ __forceinline__ void split_chunk(const float *v8, uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v8[2 * i], b = v8[2 * i + 1];
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);                       // .x = a (low half), .y = b
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __low2float(h2), b - __high2float(h2));
        h[i] = *reinterpret_cast<const uint32_t *>(&h2);
        l[i] = *reinterpret_cast<const uint32_t *>(&l2);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// pack 8 floats into one 16-byte chunk of the hi matrix and one of the lo matrix
__device__ __forceinline__ void store_chunk_split(uint8_t *hi_base, uint8_t *lo_base, uint32_t off, const float *v8) {
    uint4 h, l;
    split_chunk(v8, h, l);
    *reinterpret_cast<uint4 *>(hi_base + off) = h;
    *reinterpret_cast<uint4 *>(lo_base + off) = l;
}

// D (+)= A*B^T over `ksteps` K-steps of 16, bf16x3: (hi,hi) + (hi,lo) + (lo,hi).
// a_step / b_step: descriptor advance (bytes) per K-step.
__device__ __forceinline__ void mma_bf16x3(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t a_step,
                                           uint32_t b_step, int ksteps, uint32_t idesc, bool accumulate) {
    for (int k = 0; k < ksteps; ++k) {
        const uint64_t ah = advance_desc(a_hi, k * a_step), al = advance_desc(a_lo, k * a_step);
        const uint64_t bh = advance_desc(b_hi, k * b_step), bl = advance_desc(b_lo, k * b_step);
        mma_bf16(d_tmem, al, bh, idesc, accumulate || k > 0);  // small terms first
        mma_bf16(d_tmem, ah, bl, idesc, true);
        mma_bf16(d_tmem, ah, bh, idesc, true);
    }
}

}  // namespace umma
