// selftest.cu -- known-answer kernel for the tcgen05 plumbing in umma.cuh: the three operand
// configurations the density kernels rely on, on one 128-row tile, bf16x3.
//   D1[p][n] = sum_{k<32}  A[p][k] * W[n][k]        A K-major,  B K-major   (forward layer)
//   D2[p][k] = sum_{o<32}  A[p][o] * W[o][k], k<64  A K-major,  B MN-major  (input gradient)
//   D3[m][n] = sum_{p<128} A[p][m] * X[p][n], m<128 A MN-major, B MN-major  (weight gradient)
// A [128,128], X [128,32], W [32,64] fp32 row-major in global memory.
#include "../common.cuh"
#include "../umma.cuh"

namespace {

__global__ void __launch_bounds__(128) k_umma_selftest(const float *__restrict__ A, const float *__restrict__ X, const float *__restrict__ W,
                                                       float *__restrict__ D1, float *__restrict__ D2, float *__restrict__ D3) {
    extern __shared__ __align__(128) uint8_t smem[];
    // A: 128 rows x 16 chunks   LBO 128, SBO 2048  (32 KB per half)
    // X: 128 rows x  4 chunks   LBO 128, SBO  512  ( 8 KB per half)
    // W:  32 rows x  8 chunks   LBO 128, SBO 1024  ( 4 KB per half)
    uint8_t *A_hi = smem, *A_lo = A_hi + 32768;
    uint8_t *X_hi = A_lo + 32768, *X_lo = X_hi + 8192;
    uint8_t *W_hi = X_lo + 8192, *W_lo = W_hi + 4096;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(W_lo + 4096);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
    const int t = threadIdx.x, warp = t >> 5;

    for (int c = 0; c < 16; ++c) {
        float v[8];
        for (int i = 0; i < 8; ++i) v[i] = A[t * 128 + c * 8 + i];
        umma::store_chunk_split(A_hi, A_lo, umma::canon_off(t, c, 128, 2048), v);
    }
    for (int c = 0; c < 4; ++c) {
        float v[8];
        for (int i = 0; i < 8; ++i) v[i] = X[t * 32 + c * 8 + i];
        umma::store_chunk_split(X_hi, X_lo, umma::canon_off(t, c, 128, 512), v);
    }
    if (t < 32) {
        for (int c = 0; c < 8; ++c) {
            float v[8];
            for (int i = 0; i < 8; ++i) v[i] = W[t * 64 + c * 8 + i];
            umma::store_chunk_split(W_hi, W_lo, umma::canon_off(t, c, 128, 1024), v);
        }
    }
    if (t == 0) {
        umma::mbar_init(mbar, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, 128);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t d1 = tmem, d2 = tmem + 32, d3 = tmem + 96;  // 32 + 64 + 32 columns

    if (t == 0) {
        const uint32_t a_hi = umma::smem_u32(A_hi), a_lo = umma::smem_u32(A_lo);
        const uint32_t x_hi = umma::smem_u32(X_hi), x_lo = umma::smem_u32(X_lo);
        const uint32_t w_hi = umma::smem_u32(W_hi), w_lo = umma::smem_u32(W_lo);
        // D1: A K-major (LBO 128, SBO 2048), 2 K-steps of 16 = 2 chunks = 256 B;  B = W rows n, K-major (LBO 128, SBO 1024)
        umma::mma_bf16x3(d1, umma::make_desc(a_hi, 128, 2048), umma::make_desc(a_lo, 128, 2048), umma::make_desc(w_hi, 128, 1024),
                         umma::make_desc(w_lo, 128, 1024), 256, 256, 2, umma::idesc_bf16(128, 32, 0, 0), false);
        // D2: A K-major over o (first 32 columns of A); B = W as MN-major: N' = k (chunks, stride 128 -> SBO'), K' = o (groups of 8, stride 1024 -> LBO')
        umma::mma_bf16x3(d2, umma::make_desc(a_hi, 128, 2048), umma::make_desc(a_lo, 128, 2048), umma::make_desc(w_hi, 1024, 128),
                         umma::make_desc(w_lo, 1024, 128), 256, 2048, 2, umma::idesc_bf16(128, 64, 0, 1), false);
        // D3: A MN-major: M' = feature m (16 chunks, stride 128 -> SBO'), K' = point p (groups of 8, stride 2048 -> LBO');
        //     B = X MN-major: N' = n (4 chunks, stride 128), K' = p (groups, stride 512); 8 K-steps of 16 points = 2 groups
        umma::mma_bf16x3(d3, umma::make_desc(a_hi, 2048, 128), umma::make_desc(a_lo, 2048, 128), umma::make_desc(x_hi, 512, 128),
                         umma::make_desc(x_lo, 512, 128), 4096, 1024, 8, umma::idesc_bf16(128, 32, 1, 1), false);
        umma::commit(mbar);
    }
    umma::mbar_wait(mbar, 0);
    umma::fence_after_sync();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float v[32];
    umma::tmem_ld32(d1 + lane_base, v);
    umma::tmem_wait_ld();
    for (int i = 0; i < 32; ++i) D1[t * 32 + i] = v[i];
    umma::tmem_ld32(d2 + lane_base, v);
    umma::tmem_wait_ld();
    for (int i = 0; i < 32; ++i) D2[t * 64 + i] = v[i];
    umma::tmem_ld32(d2 + 32 + lane_base, v);
    umma::tmem_wait_ld();
    for (int i = 0; i < 32; ++i) D2[t * 64 + 32 + i] = v[i];
    umma::tmem_ld32(d3 + lane_base, v);
    umma::tmem_wait_ld();
    for (int i = 0; i < 32; ++i) D3[t * 32 + i] = v[i];
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 128);
}

}  // namespace

extern "C" int nafb_selftest_umma(const float *A, const float *X, const float *W, float *D1, float *D2, float *D3, nafb_stream_t stream) {
    if (!A || !X || !W || !D1 || !D2 || !D3) NAFB_FAIL(NAFB_ERR_INVALID, "selftest_umma: null pointer");
    const size_t smem = 2 * (32768 + 8192 + 4096) + 64;
    cudaError_t e = cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) NAFB_FAIL(NAFB_ERR_CUDA, "selftest_umma: %s", cudaGetErrorString(e));
    k_umma_selftest<<<1, 128, smem, (cudaStream_t)stream>>>(A, X, W, D1, D2, D3);
    NAFB_CHECK_LAUNCH("selftest_umma");
    return NAFB_OK;
}
