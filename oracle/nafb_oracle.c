/*
 * nafb_oracle.c -- CPU restatement of the reference's multi-resolution hash-grid op.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker or
 * as the CPU baseline being reported.
 *
 * Every function cites the reference lines it restates
 * (paths relative to /root/reference):
 *   src/encoder/hashencoder/src/hashencoder.cu  -> "hashencoder.cu"
 *
 * Floating-point contract.  The reference is compiled by nvcc with the default
 * -fmad=true, so (SASS probe):
 *     scale      = fma(exp2f(level), H, -1)         hashencoder.cu:99
 *     pos        = fma(x, scale, 0.5)               hashencoder.cu:108
 *     w          = ((1*a0)*a1)*a2 , a_d = pos_d or (1 - pos_d)   :120-132 (no fusion possible)
 *     result[c]  = fma(w, grid[idx+c], result[c])   hashencoder.cu:139, corner order 0..7
 * We spell those with explicit fmaf() and compile with -ffp-contract=off so the
 * host compiler cannot add or remove a fusion.  Pinned against the reference's own
 * kernel text compiled for the host (oracle/_ref, see oracle/build_ref.sh).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define NAFB_MAX_D 3
#define NAFB_MAX_C 8

/* hashencoder.cu:36-52  fast_hash<D>: xor of pos*prime, uint32 wrap-around. */
static inline uint32_t oracle_fast_hash(const uint32_t *pos_grid, uint32_t D) {
    static const uint32_t primes[7] = {1u, 19349663u, 83492791u, 25165843u,
                                       6291469u, 12582917u, 3145739u};
    uint32_t r = 0;
    for (uint32_t i = 0; i < D; ++i) r ^= pos_grid[i] * primes[i];
    return r;
}

/* hashencoder.cu:55-74  get_grid_index<D,C>.
 * NOTE the uint32 stride: `stride *= (resolution + 1)` wraps modulo 2^32, so at
 * 16 levels x base 16 the test `stride > hashmap_size` is false at levels 12, 13
 * and those levels use the (wrapped) linear index instead of the hash. */
uint32_t oracle_grid_index(uint32_t D, uint32_t C, uint32_t ch, uint32_t hashmap_size,
                           uint32_t resolution, const uint32_t *pos_grid) {
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; ++d) {
        index += pos_grid[d] * stride;
        stride *= (resolution + 1u);
    }
    if (stride > hashmap_size) index = oracle_fast_hash(pos_grid, D);
    return (index % hashmap_size) * C + ch;
}

/* hashencoder.cu:98-100: per-level constants. */
static inline void oracle_level_consts(const int32_t *offsets, uint32_t level, uint32_t H,
                                       uint32_t *hashmap_size, float *scale, uint32_t *resolution) {
    *hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    *scale = fmaf(exp2f((float)level), (float)H, -1.0f);
    *resolution = (uint32_t)ceilf(*scale) + 1u;
}

/* hashencoder.cu:103-111: position inside the level's lattice. */
static inline void oracle_locate(const float *x, uint32_t D, float scale, float *pos, uint32_t *pos_grid) {
    for (uint32_t d = 0; d < D; ++d) {
        pos[d] = fmaf(x[d], scale, 0.5f);
        pos_grid[d] = (uint32_t)floorf(pos[d]);
        pos[d] -= (float)pos_grid[d];
    }
}

/* Forward gather + D-linear interpolation.      hashencoder.cu:77-198 (kernel_grid)
 *   inputs   [B, D]   in [0,1]
 *   grid     [sum T_l, C]
 *   offsets  [L+1]    entry offsets
 *   outputs  [L, B, C]                      (hashencoder.cu:96)
 *   dy_dx    [B, L, D, C] or NULL           (hashencoder.cu:153-197)
 */
void oracle_hash_forward(const float *inputs, const float *grid, const int32_t *offsets,
                         float *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                         uint32_t H, float *dy_dx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (uint32_t level = 0; level < L; ++level) {
        for (uint32_t b0 = 0; b0 < B; b0 += 256) {
            uint32_t hashmap_size, resolution;
            float scale;
            oracle_level_consts(offsets, level, H, &hashmap_size, &scale, &resolution);
            const float *g = grid + (size_t)(uint32_t)offsets[level] * C;
            uint32_t b1 = b0 + 256 < B ? b0 + 256 : B;
            for (uint32_t b = b0; b < b1; ++b) {
                float pos[NAFB_MAX_D];
                uint32_t pg[NAFB_MAX_D], pl[NAFB_MAX_D];
                oracle_locate(inputs + (size_t)b * D, D, scale, pos, pg);
                float res[NAFB_MAX_C] = {0};
                for (uint32_t idx = 0; idx < (1u << D); ++idx) { /* :119-143 */
                    float w = 1.0f;
                    for (uint32_t d = 0; d < D; ++d) {
                        if ((idx & (1u << d)) == 0) { w *= 1.0f - pos[d]; pl[d] = pg[d]; }
                        else                        { w *= pos[d];        pl[d] = pg[d] + 1u; }
                    }
                    uint32_t index = oracle_grid_index(D, C, 0, hashmap_size, resolution, pl);
                    for (uint32_t ch = 0; ch < C; ++ch) res[ch] = fmaf(w, g[index + ch], res[ch]);
                }
                float *out = outputs + (size_t)level * B * C + (size_t)b * C;
                for (uint32_t ch = 0; ch < C; ++ch) out[ch] = res[ch];

                if (dy_dx) { /* :153-197 */
                    float *dd = dy_dx + (size_t)b * D * L * C + (size_t)level * D * C;
                    for (uint32_t gd = 0; gd < D; ++gd) {
                        float rg[NAFB_MAX_C] = {0};
                        for (uint32_t idx = 0; idx < (1u << (D - 1)); ++idx) {
                            float w = 1.0f;
                            for (uint32_t nd = 0; nd < D - 1; ++nd) {
                                uint32_t d = nd >= gd ? nd + 1 : nd; /* see note below */
                                if ((idx & (1u << nd)) == 0) { w *= 1.0f - pos[d]; pl[d] = pg[d]; }
                                else                         { w *= pos[d];        pl[d] = pg[d] + 1u; }
                            }
                            pl[gd] = pg[gd];
                            uint32_t il = oracle_grid_index(D, C, 0, hashmap_size, resolution, pl);
                            pl[gd] = pg[gd] + 1u;
                            uint32_t ir = oracle_grid_index(D, C, 0, hashmap_size, resolution, pl);
                            for (uint32_t ch = 0; ch < C; ++ch)
                                rg[ch] = fmaf(w, g[ir + ch] - g[il + ch], rg[ch]);
                        }
                        for (uint32_t ch = 0; ch < C; ++ch) dd[gd * C + ch] = rg[ch];
                    }
                }
            }
        }
    }
}

/* NOTE on dy_dx (hashencoder.cu:170): the reference writes `d = nd > gd ? nd+1 : nd`,
 * which for nd == gd selects d == gd again (the differentiated axis) instead of
 * skipping it, and leaves pos_grid_local[] of the skipped axis UNINITIALISED
 * (undefined behaviour; only gd == D-1 is well defined).  That is a reference
 * defect on a path NAF never runs (calc_grad_inputs is False, hashgrid.py:132).
 * oracle_hash_forward implements the mathematically intended skip (nd >= gd); the
 * pin against oracle/_ref therefore compares dy_dx only for gd == D-1. */

/* Backward scatter.                        hashencoder.cu:201-272 (kernel_grid_backward)
 *   grad       [B, L*C]                     (hashencoder.cu:219)
 *   grad_grid  [sum T_l, C]  accumulated into (caller pre-zeroes, hashgrid.py:59)
 * Levels own disjoint table ranges, so parallelising over levels is race free
 * and the per-entry accumulation order is the point order 0..B-1: deterministic.
 * (The GPU reference uses unordered float atomics; tests use a tolerance.)
 * If grad_grid64 != NULL the same sums are also accumulated in double there.
 */
void oracle_hash_backward(const float *grad, const float *inputs, const int32_t *offsets,
                          float *grad_grid, double *grad_grid64, uint32_t B, uint32_t D,
                          uint32_t C, uint32_t L, uint32_t H) {
#pragma omp parallel for schedule(dynamic, 1)
    for (uint32_t level = 0; level < L; ++level) {
        uint32_t hashmap_size, resolution;
        float scale;
        oracle_level_consts(offsets, level, H, &hashmap_size, &scale, &resolution);
        float *gg = grad_grid + (size_t)(uint32_t)offsets[level] * C;
        double *gg64 = grad_grid64 ? grad_grid64 + (size_t)(uint32_t)offsets[level] * C : NULL;
        for (uint32_t b = 0; b < B; ++b) {
            float pos[NAFB_MAX_D];
            uint32_t pg[NAFB_MAX_D], pl[NAFB_MAX_D];
            oracle_locate(inputs + (size_t)b * D, D, scale, pos, pg);
            const float *gr = grad + (size_t)b * L * C + (size_t)level * C;
            for (uint32_t idx = 0; idx < (1u << D); ++idx) { /* :238-271 */
                float w = 1.0f;
                for (uint32_t d = 0; d < D; ++d) {
                    if ((idx & (1u << d)) == 0) { w *= 1.0f - pos[d]; pl[d] = pg[d]; }
                    else                        { w *= pos[d];        pl[d] = pg[d] + 1u; }
                }
                uint32_t index = oracle_grid_index(D, C, 0, hashmap_size, resolution, pl);
                for (uint32_t c = 0; c < C; ++c) {
                    float v = w * gr[c]; /* :268 */
                    gg[index + c] += v;
                    if (gg64) gg64[index + c] += (double)w * (double)gr[c];
                }
            }
        }
    }
}

/* hashencoder.cu:275-298 (kernel_input_backward): grad_inputs[b,d] += sum_l sum_c grad*dy_dx. */
void oracle_input_backward(const float *grad, const float *dy_dx, float *grad_inputs,
                           uint32_t B, uint32_t D, uint32_t C, uint32_t L) {
    for (uint32_t t = 0; t < B * D; ++t) {
        uint32_t b = t / D, d = t - b * D;
        const float *g = grad + (size_t)b * L * C;
        const float *dd = dy_dx + (size_t)b * L * D * C;
        for (uint32_t l = 0; l < L; ++l)
            for (uint32_t ch = 0; ch < C; ++ch)
                grad_inputs[t] = fmaf(g[l * C + ch], dd[l * D * C + d * C + ch], grad_inputs[t]);
    }
}

/* Corner indices + weights for one (point, level): used by the index-parity tests.
 * entry[8] = get_grid_index(ch=0)/C ; weight[8] in corner order. */
void oracle_corners(const float *x, const int32_t *offsets, uint32_t level, uint32_t D,
                    uint32_t C, uint32_t H, uint32_t *entry, float *weight, uint32_t *pos_grid_out,
                    float *frac_out) {
    uint32_t hashmap_size, resolution;
    float scale;
    oracle_level_consts(offsets, level, H, &hashmap_size, &scale, &resolution);
    float pos[NAFB_MAX_D];
    uint32_t pg[NAFB_MAX_D], pl[NAFB_MAX_D];
    oracle_locate(x, D, scale, pos, pg);
    for (uint32_t d = 0; d < D; ++d) { if (pos_grid_out) pos_grid_out[d] = pg[d]; if (frac_out) frac_out[d] = pos[d]; }
    for (uint32_t idx = 0; idx < (1u << D); ++idx) {
        float w = 1.0f;
        for (uint32_t d = 0; d < D; ++d) {
            if ((idx & (1u << d)) == 0) { w *= 1.0f - pos[d]; pl[d] = pg[d]; }
            else                        { w *= pos[d];        pl[d] = pg[d] + 1u; }
        }
        entry[idx] = oracle_grid_index(D, C, 0, hashmap_size, resolution, pl) / C;
        weight[idx] = w;
    }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
