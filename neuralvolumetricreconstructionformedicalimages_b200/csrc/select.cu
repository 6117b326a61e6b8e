// select.cu -- the per-iteration work of the reference's dataset on the device:
//   * get_ptycho_mask (src/utils/util.py:196-205, called every iteration at train.py:59-60), once per scan;
//   * TIGREDataset.__getitem__ (src/dataset/tigre.py:354-382): among the pixels of one projection whose value is non-zero, draw
//     n_rays WITHOUT replacement (np.random.choice(replace=False), :358), gather their projection values (:364) -- and here
//     also their mask bits (train.py:93-95) -- straight into the buffers the fused training step reads.
// Nothing changes between launches but a device-resident draw counter, so the draw sits in the step's CUDA graph.
//
// Uniform sampling without replacement = the n smallest of i.i.d. random keys, in key order (a uniformly random ordered
// subset, the distribution of np.random.choice(replace=False)).  Keys are counter-based: hash(seed, draw, candidate) -- no
// state per candidate, the same draw for the same (seed, draw counter).  One CTA, one pass over the candidates: those whose key
// lies below a threshold just above the expected n-th smallest key are collected in shared memory and bitonic-sorted; the first n
// are the draw, so neither the set nor the order depends on the order the atomics were served in.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// 32-bit key of candidate j in the stream (seed, draw): two rounds of an integer finaliser (full avalanche), a dozen instructions --
// the draw kernel is ONE CTA whose time is the hashing of every candidate
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t draw_key(uint64_t stream_key, uint32_t j) {
    return mix32(mix32(j ^ (uint32_t)stream_key) + (uint32_t)(stream_key >> 32));
}

// m = |hr| < thr; m[1:, :] &= (m[1:, :] == m[:-1, :]); m[:, 1:] &= (m[:, 1:] == m[:, :-1]); keep = ~m
// (each right-hand side is evaluated on the mask as it was BEFORE that statement, as torch does).
__global__ void __launch_bounds__(256) k_ptycho_mask(const float2 *__restrict__ hr, uint32_t P, uint32_t H, uint32_t W, float thr,
                                                     uint8_t *__restrict__ keep) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n = (uint64_t)P * H * W;
    if (i >= n) return;
    const uint32_t c = (uint32_t)(i % W), r = (uint32_t)((i / W) % H);
    const float2 *img = hr + (i - (uint64_t)r * W - c);
    auto m0 = [&](uint32_t rr, uint32_t cc) {
        const float2 v = __ldg(img + (uint64_t)rr * W + cc);
        return hypotf(v.x, v.y) < thr;   // torch.abs of a complex tensor
    };
    auto m1 = [&](uint32_t rr, uint32_t cc) {
        const bool a = m0(rr, cc);
        return rr == 0 ? a : (a && (a == m0(rr - 1, cc)));
    };
    const bool a = m1(r, c);
    const bool m2 = c == 0 ? a : (a && (a == m1(r, c - 1)));
    keep[i] = m2 ? 0 : 1;
}

constexpr int DRAW_THREADS = 1024;
constexpr uint32_t DRAW_MAX = 8192;    // rays per draw (up to 2 x 8192 candidate pairs in shared memory: 128 KB)

struct DrawParams {
    const float *projs;        // [P, H*W]
    const uint8_t *mask;       // [P, H*W] or nullptr
    const int32_t *valid;      // [P, H*W]: the flat indices of the non-zero pixels of every projection, front-packed
    const int32_t *n_valid;    // [P]
    const int32_t *order;      // [n_order] projection of draw k (k modulo n_order), or nullptr: k modulo P
    uint32_t P, HW, W, n_order;
    uint32_t n;                // rays to draw
    int32_t *pixels_out;       // [n, 3] (projection, row, col)
    float *projs_out;          // [n]
    uint8_t *mask_out;         // [n] or nullptr
    uint32_t *state;           // [4]: draw counter, seed lo, seed hi, error flag
};

__global__ void __launch_bounds__(DRAW_THREADS) k_draw_pixels(const DrawParams D) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t *pairs = reinterpret_cast<uint64_t *>(smem_raw);             // [n_pow2] (key << 32 | candidate)
    __shared__ uint32_t s_count;
    const uint32_t t = threadIdx.x;
    const uint32_t draw = D.state[0];
    const uint64_t seed = (uint64_t)D.state[1] | ((uint64_t)D.state[2] << 32);
    const uint32_t p = D.order ? (uint32_t)__ldg(D.order + draw % D.n_order) % D.P : draw % D.P;
    const uint32_t M = (uint32_t)__ldg(D.n_valid + p);
    const int32_t *cand = D.valid + (uint64_t)p * D.HW;
    const uint64_t skey = mix64(seed + (uint64_t)draw * 0xD1342543DE82EF95ull);
    const uint32_t n = D.n;
    uint32_t n_pow2 = 1;
    while (n_pow2 < n) n_pow2 <<= 1;
    if (M < n) {   // fewer valid pixels than rays: the reference's np.random.choice raises; flag it and emit the first pixels
        if (t == 0) D.state[3] = 1u + p;
    }
    // ---- ONE pass: the keys are uniform, so the n-th smallest sits near n / M * 2^32; every candidate whose key is below a threshold
    // six standard deviations above that is collected (n + 6 sqrt(n) + 8 of them on average) and the n smallest are kept by the
    // sort below.  Should the list come out short (probability ~ 1e-9) or overflow, the threshold is adjusted and the pass repeated.
    const uint32_t n_out = n < M ? n : M;
    const uint32_t cap = 2u * n_pow2;                       // pairs that fit in shared memory
    const double want_cnt = (double)n_out + 6.0 * sqrt((double)n_out) + 8.0;
    uint32_t T = want_cnt >= (double)M ? 0xffffffffu : (uint32_t)(want_cnt / (double)M * 4294967296.0);
    for (int attempt = 0; attempt < 40; ++attempt) {
        if (t == 0) s_count = 0;
        for (uint32_t i = t; i < cap; i += DRAW_THREADS) pairs[i] = ~0ull;
        __syncthreads();
        for (uint32_t jj = t; jj < M; jj += DRAW_THREADS) {
            const uint32_t k = draw_key(skey, jj);
            if (k <= T) {
                const uint32_t q = atomicAdd(&s_count, 1u);
                if (q < cap) pairs[q] = ((uint64_t)k << 32) | jj;
            }
        }
        __syncthreads();
        const uint32_t got = s_count;
        __syncthreads();
        if (got >= n_out && got <= cap) break;
        T = got < n_out ? (T >= 0x7fffffffu ? 0xffffffffu : 2u * T + 1u) : T - T / 4u;   // too few: widen; overflow: narrow
    }
    // ---- bitonic sort of the collected pairs (padding = ~0 sorts last): position i of the batch gets the i-th smallest key -- a
    // uniformly random ORDER as well, independent of the order the atomics above were served in
    uint32_t sort_n = 1;
    {
        const uint32_t got = s_count < cap ? s_count : cap;
        while (sort_n < got) sort_n <<= 1;
    }
    for (uint32_t k = 2; k <= sort_n; k <<= 1) {
        for (uint32_t jj = k >> 1; jj > 0; jj >>= 1) {
            for (uint32_t c = t; c < (sort_n >> 1); c += DRAW_THREADS) {     // comparator c works on (i, i + jj), bit jj of i clear
                const uint32_t i = 2u * c - (c & (jj - 1u)), l = i + jj;
                const uint64_t a = pairs[i], b = pairs[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { pairs[i] = b; pairs[l] = a; }
            }
            __syncthreads();
        }
    }
    // ---- gather
    for (uint32_t i = t; i < n; i += DRAW_THREADS) {
        const uint32_t j = i < n_out ? (uint32_t)(pairs[i] & 0xffffffffu) : (M ? i % M : 0u);
        const uint32_t pix = M ? (uint32_t)__ldg(cand + j) : 0u;
        D.pixels_out[3 * i] = (int32_t)p;
        D.pixels_out[3 * i + 1] = (int32_t)(pix / D.W);
        D.pixels_out[3 * i + 2] = (int32_t)(pix % D.W);
        D.projs_out[i] = __ldg(D.projs + (uint64_t)p * D.HW + pix);
        if (D.mask_out) D.mask_out[i] = D.mask ? __ldg(D.mask + (uint64_t)p * D.HW + pix) : (uint8_t)1;
    }
    __syncthreads();
    if (t == 0) D.state[0] = draw + 1u;
}

}  // namespace

extern "C" {

int nafb_ptycho_mask(const float *full_proj, uint32_t n_proj, uint32_t H, uint32_t W, float threshold, uint8_t *keep, nafb_stream_t stream) {
    if (!full_proj || !keep) NAFB_FAIL(NAFB_ERR_INVALID, "ptycho_mask: null pointer");
    const uint64_t n = (uint64_t)n_proj * H * W;
    if (n == 0) return NAFB_OK;
    k_ptycho_mask<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(full_proj), n_proj, H, W, threshold, keep);
    NAFB_CHECK_LAUNCH("ptycho_mask");
    return NAFB_OK;
}

int nafb_draw_pixels(const nafb_pixel_source *src, uint32_t n_rays, int32_t *pixels_out, float *projs_out, uint8_t *mask_out,
                     uint32_t *draw_state, nafb_stream_t stream) {
    if (!src || !src->projs || !src->valid || !src->n_valid || !pixels_out || !projs_out || !draw_state)
        NAFB_FAIL(NAFB_ERR_INVALID, "draw_pixels: null pointer");
    if (n_rays == 0) return NAFB_OK;
    if (n_rays > DRAW_MAX) NAFB_FAIL(NAFB_ERR_UNSUPPORTED, "draw_pixels: at most %u rays per draw (got %u)", DRAW_MAX, n_rays);
    if (src->n_proj == 0 || src->H == 0 || src->W == 0 || (uint64_t)src->H * src->W > 0x7fffffffull) NAFB_FAIL(NAFB_ERR_INVALID, "draw_pixels: bad projection shape");
    if (src->order && src->n_order == 0) NAFB_FAIL(NAFB_ERR_INVALID, "draw_pixels: empty projection order");
    DrawParams D;
    D.projs = src->projs; D.mask = src->mask; D.valid = src->valid; D.n_valid = src->n_valid; D.order = src->order;
    D.P = src->n_proj; D.HW = src->H * src->W; D.W = src->W; D.n_order = src->n_order; D.n = n_rays;
    D.pixels_out = pixels_out; D.projs_out = projs_out; D.mask_out = mask_out; D.state = draw_state;
    uint32_t n_pow2 = 1;
    while (n_pow2 < n_rays) n_pow2 <<= 1;
    const size_t smem = (size_t)2 * n_pow2 * sizeof(uint64_t);
    static bool configured[NAFB_MAX_DEVICES] = {};
    NAFB_CONFIGURE_SMEM(configured, k_draw_pixels, 2 * DRAW_MAX * sizeof(uint64_t), "draw_pixels");
    k_draw_pixels<<<1, DRAW_THREADS, smem, (cudaStream_t)stream>>>(D);
    NAFB_CHECK_LAUNCH("draw_pixels");
    return NAFB_OK;
}

}  // extern "C"
