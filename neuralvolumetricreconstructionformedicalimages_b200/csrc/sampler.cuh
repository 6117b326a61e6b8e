// sampler.cuh -- where sample points come from: a points tensor, rays (+ stratified /
// perturbed sampling, reference src/render/render.py:88-105) or the voxel lattice
// (reference src/dataset/tigre.py:388-400).
//
// The reference is eager PyTorch: every elementary op rounds to fp32 on its own, so every
// product and sum below is spelled with __fmul_rn/__fadd_rn/__fsub_rn (never contracted);
// torch.linspace is the one place where ATen itself fuses (see linspace01).
#pragma once
#include "common.cuh"

struct SamplerParams {
    const float *pts;
    const float *rays;
    const float *t_rand;
    const uint32_t *rng_state;   // device nafb_step_state (step, seed): in-kernel uniforms when t_rand is NULL
    // in-kernel ray generation (RAYS source with `pixels` instead of `rays`): tigre.py:402-456 / :463-528
    const float *poses;          // [n_proj, 12] fp32: R (3x3 row-major) | t (3)
    const int32_t *pixels;       // [N, 3]: projection, detector row, detector column
    float det_w2, det_h2;        // fl(W / 2), fl(H / 2)
    float det_du, det_dv, det_u0, det_v0, det_dsd, det_near, det_far;
    int32_t det_parallel;
    uint32_t n_rays, n_samples;
    int32_t perturb;
    uint32_t n1, n2, n3, i0, i1;
    double s1, s2, s3;
    double vstep1, vstep2, vstep3;   // np.linspace step (s - (-s)) / (n - 1), evaluated once on the host in float64
    float bound, inv_2bound, clamp;
    float lin_step;  // fl(1 / (S-1))
};

#ifdef __CUDACC__

// torch.linspace(0, 1, S)[i]: i < S/2 ? fma(step, i, 0) : fma(-step, S-1-i, 1)   (render.py:91)
__device__ __forceinline__ float linspace01(uint32_t i, uint32_t S, float step) {
    if (S <= 1) return 0.f;
    return i < S / 2 ? __fmul_rn(step, (float)i) : __fmaf_rn(-step, (float)(S - 1 - i), 1.0f);
}

// z = near*(1-t) + far*t            (render.py:92)
__device__ __forceinline__ float z_uniform(float near, float far, uint32_t i, uint32_t S, float step) {
    const float t = linspace01(i, S, step);
    return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(far, t));
}

// Where the uniforms of render.py:99 (torch.rand([n_rays, n_samples])) come from: a tensor the caller filled
// (bit-exact parity with a given draw), or a counter-based generator evaluated in-kernel -- same value for the same
// (seed, step, ray, sample) in the forward and the backward kernel, no [N,S] buffer, no generator launch.
// splitmix64 of the 64-bit counter, top 24 bits -> multiples of 2^-24 in [0,1) like torch.rand for float32.
struct Jitter {
    const float *row;   // t_rand + ray * S, or nullptr
    uint64_t key;       // generator key of this step (seed and step folded together)
    uint64_t base;      // ray * S
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ float jitter_uniform(const Jitter &j, uint32_t i) {
    if (j.row) return __ldg(j.row + i);
    const uint64_t z = splitmix64(j.key + (j.base + i) * 0x9E3779B97F4A7C15ull);
    return (float)(uint32_t)(z >> 40) * 5.9604644775390625e-08f;   // 2^-24
}

// perturbed sample i of a ray (render.py:95-100): lower + (upper-lower)*t_rand
__device__ __forceinline__ float z_sample(float near, float far, uint32_t i, uint32_t S, float step, bool perturb, const Jitter &jit) {
    const float zi = z_uniform(near, far, i, S, step);
    if (!perturb) return zi;
    float lower = zi, upper = zi;
    if (i > 0) lower = __fmul_rn(0.5f, __fadd_rn(zi, z_uniform(near, far, i - 1, S, step)));
    if (i + 1 < S) upper = __fmul_rn(0.5f, __fadd_rn(z_uniform(near, far, i + 1, S, step), zi));
    return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), jitter_uniform(jit, i)));
}

struct RayRegs {
    float o[3], d[3], near, far, norm;
};

__device__ __forceinline__ RayRegs load_ray(const float *__restrict__ rays, uint32_t r) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(rays) + 2 * (size_t)r);
    const float4 b = __ldg(reinterpret_cast<const float4 *>(rays) + 2 * (size_t)r + 1);
    RayRegs R;
    R.o[0] = a.x; R.o[1] = a.y; R.o[2] = a.z;
    R.d[0] = a.w; R.d[1] = b.x; R.d[2] = b.y;
    R.near = b.z; R.far = b.w;
    R.norm = sqrtf(__fmaf_rn(R.d[2], R.d[2], __fmaf_rn(R.d[1], R.d[1], __fmul_rn(R.d[0], R.d[0]))));  // render.py:194
    return R;
}

// pts = clamp(o + d*z, -c, c)       (render.py:103-105)
__device__ __forceinline__ float ray_point(float o, float d, float z, float c) {
    return fminf(fmaxf(__fadd_rn(o, __fmul_rn(d, z)), -c), c);
}

// np.linspace(-s, s, n)[i] in float64, then the fp32 cast of tigre.py:277
__device__ __forceinline__ float voxel_coord(uint32_t i, uint32_t n, double s, double step) {
    if (n == 1) return (float)(-s);
    if (i == n - 1) return (float)s;
    return (float)__dadd_rn(__dmul_rn((double)i, step), -s);
}

// Ray of detector pixel (projection, row, col), generated in-kernel with the arithmetic of the reference's dataset code
// as torch evaluates it on fp32 tensors, one rounding per elementary op (tigre.py:428-447 / :486-501, poses :530-572):
//   uu = ((col + 0.5) - W/2) * dDet[0] + offDet[0]      vv likewise with rows
//   cone:      d = R . [uu/DSD, vv/DSD, 1],  o = t       parallel:  d = R . [0, 0, 1],  o = R . [uu, vv, 0] + t
// and the 3x3 product as torch.matmul rounds it: ((r0*x + r1*y) + r2*z), no fused multiply-add.  near / far: the global
// pair of tigre.py:575-586.  tests/test_gpu_parity.py compares with the reference's own rays bit for bit.
__device__ __forceinline__ float dot3_lr(float r0, float r1, float r2, float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(r0, x), __fmul_rn(r1, y)), __fmul_rn(r2, z));
}
__device__ __forceinline__ RayRegs make_ray(const SamplerParams &sp, uint32_t r) {
    const int32_t pj = __ldg(sp.pixels + 3 * (size_t)r), row = __ldg(sp.pixels + 3 * (size_t)r + 1), col = __ldg(sp.pixels + 3 * (size_t)r + 2);
    const float *__restrict__ T = sp.poses + 12 * (size_t)pj;
    float M[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) M[i] = __ldg(T + i);
    const float uu = __fadd_rn(__fmul_rn(__fsub_rn(__fadd_rn((float)col, 0.5f), sp.det_w2), sp.det_du), sp.det_u0);
    const float vv = __fadd_rn(__fmul_rn(__fsub_rn(__fadd_rn((float)row, 0.5f), sp.det_h2), sp.det_dv), sp.det_v0);
    RayRegs R;
    if (sp.det_parallel) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            R.d[k] = dot3_lr(M[3 * k], M[3 * k + 1], M[3 * k + 2], 0.f, 0.f, 1.f);
            R.o[k] = __fadd_rn(dot3_lr(M[3 * k], M[3 * k + 1], M[3 * k + 2], uu, vv, 0.f), M[9 + k]);
        }
    } else {
        const float x = __fdiv_rn(uu, sp.det_dsd), y = __fdiv_rn(vv, sp.det_dsd);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            R.d[k] = dot3_lr(M[3 * k], M[3 * k + 1], M[3 * k + 2], x, y, 1.f);
            R.o[k] = M[9 + k];
        }
    }
    R.near = sp.det_near; R.far = sp.det_far;
    R.norm = sqrtf(__fmaf_rn(R.d[2], R.d[2], __fmaf_rn(R.d[1], R.d[1], __fmul_rn(R.d[0], R.d[0]))));
    return R;
}
// the ray of index r of the launch: from the rays tensor, or generated from its detector pixel
__device__ __forceinline__ RayRegs load_ray(const SamplerParams &sp, uint32_t r) {
    return sp.pixels ? make_ray(sp, r) : load_ray(sp.rays, r);
}

__device__ __forceinline__ Jitter jitter_for(const SamplerParams &sp, uint32_t ray) {
    Jitter j;
    j.row = sp.t_rand ? sp.t_rand + (size_t)ray * sp.n_samples : nullptr;
    j.base = (uint64_t)ray * sp.n_samples;
    j.key = 0;
    if (!sp.t_rand && sp.rng_state && sp.perturb) {
        const uint64_t seed = (uint64_t)__ldg(sp.rng_state + 1) | ((uint64_t)__ldg(sp.rng_state + 2) << 32);
        j.key = splitmix64(seed + (uint64_t)__ldg(sp.rng_state) * 0xD1342543DE82EF95ull);
    }
    return j;
}

// Voxel source, blocked traversal: a 128-point tile is an 8 (i) x 4 (j) x 4 (k) block of the lattice instead of 128 consecutive
// voxels of one k-line, so that the points of a tile share grid cells -- and therefore table sectors -- in all three directions on
// every level whose cells are larger than a voxel.  The lanes of a warp run along i, the lattice axis that is x[0], the FASTEST
// axis of the table (linear index x0 + x1 R + x2 R^2; hash x0 ^ x1 P1 ^ x2 P2 keeps aligned runs of x0 inside one 128-byte
// line): 8 consecutive i at a fixed (j, k) touch 1-3 lines per corner where 8 consecutive k touch 8.  Measured at 512^3: 30.0 ms
// against 32.9 ms for a 4 x 4 x 8 block with k across the lanes; 8x2x8, 16x4x2, 16x2x4, 32x2x2, 32x4x1 and 128x1x1 blocks with
// i across the lanes all land within 2 % (scripts/calls/r2_call38.sh).  Same voxel centres (voxel_coord), same output index;
// only the order of evaluation changes (bit-identical results).
constexpr uint32_t VOX_DI = 8, VOX_DJ = 4, VOX_DK = 4;
__host__ __device__ __forceinline__ uint64_t voxel_block_tiles(const SamplerParams &sp) {
    return (uint64_t)((sp.i1 - sp.i0 + VOX_DI - 1) / VOX_DI) * ((sp.n2 + VOX_DJ - 1) / VOX_DJ) * ((sp.n3 + VOX_DK - 1) / VOX_DK);
}
__device__ __forceinline__ bool voxel_block_point(const SamplerParams &sp, uint64_t tile, uint32_t r, float (&x)[3], uint64_t &out_index) {
    const uint32_t nbk = (sp.n3 + VOX_DK - 1) / VOX_DK, nbj = (sp.n2 + VOX_DJ - 1) / VOX_DJ;
    const uint32_t bk = (uint32_t)(tile % nbk);
    const uint64_t t2 = tile / nbk;
    const uint32_t bj = (uint32_t)(t2 % nbj), bi = (uint32_t)(t2 / nbj);
    const uint32_t i = sp.i0 + bi * VOX_DI + (r & 7u), k = bk * VOX_DK + ((r >> 3) & 3u), j = bj * VOX_DJ + (r >> 5);
    if (i >= sp.i1 || j >= sp.n2 || k >= sp.n3) return false;
    x[0] = voxel_coord(i, sp.n1, sp.s1, sp.vstep1);
    x[1] = voxel_coord(j, sp.n2, sp.s2, sp.vstep2);
    x[2] = voxel_coord(k, sp.n3, sp.s3, sp.vstep3);
    out_index = ((uint64_t)(i - sp.i0) * sp.n2 + j) * sp.n3 + k;
    return true;
}

// Fetch point p of the launch (world coordinates). For RAYS also yields ray/sample index.
template <int SRC>
__device__ __forceinline__ void fetch_point(const SamplerParams &sp, uint64_t p, float (&x)[3]) {
    if constexpr (SRC == NAFB_SRC_POINTS) {
        x[0] = __ldg(sp.pts + 3 * p);
        x[1] = __ldg(sp.pts + 3 * p + 1);
        x[2] = __ldg(sp.pts + 3 * p + 2);
    } else if constexpr (SRC == NAFB_SRC_RAYS) {
        const uint32_t r = (uint32_t)(p / sp.n_samples), i = (uint32_t)(p - (uint64_t)r * sp.n_samples);
        const RayRegs R = load_ray(sp, r);
        const float z = z_sample(R.near, R.far, i, sp.n_samples, sp.lin_step, sp.perturb != 0, jitter_for(sp, r));
#pragma unroll
        for (int d = 0; d < 3; ++d) x[d] = ray_point(R.o[d], R.d[d], z, sp.clamp);
    } else {
        const uint64_t plane = (uint64_t)sp.n2 * sp.n3;
        const uint32_t i = sp.i0 + (uint32_t)(p / plane);
        const uint32_t rem = (uint32_t)(p - (uint64_t)(i - sp.i0) * plane);
        const uint32_t j = rem / sp.n3, k = rem - j * sp.n3;
        x[0] = voxel_coord(i, sp.n1, sp.s1, sp.vstep1);
        x[1] = voxel_coord(j, sp.n2, sp.s2, sp.vstep2);
        x[2] = voxel_coord(k, sp.n3, sp.s3, sp.vstep3);
    }
}

// delta_i * |d| of raw2outputs (render.py:192-194): (z[i+1]-z[i]) * norm, last = 1e-10 * norm
__device__ __forceinline__ float ray_delta(const SamplerParams &sp, const RayRegs &R, uint32_t r, uint32_t i) {
    const uint32_t S = sp.n_samples;
    const Jitter tr = jitter_for(sp, r);
    float dist;
    if (i + 1 < S) {
        const float z0 = z_sample(R.near, R.far, i, S, sp.lin_step, sp.perturb != 0, tr);
        const float z1 = z_sample(R.near, R.far, i + 1, S, sp.lin_step, sp.perturb != 0, tr);
        dist = __fsub_rn(z1, z0);
    } else {
        dist = 1e-10f;
    }
    return __fmul_rn(dist, R.norm);
}

// Sample i of ray r AND its ray-integral weight delta_i |d| in one go (what fetch_point + ray_delta evaluate separately): the
// uniform positions i-1 .. i+2 once, two jitter draws, one ray.  Same operations, same roundings: bit-identical results.
__device__ __forceinline__ void ray_sample_and_delta(const SamplerParams &sp, const RayRegs &R, uint32_t r, uint32_t i, bool want_delta,
                                                     float (&x)[3], float &z0, float &delta) {
    const uint32_t S = sp.n_samples;
    const bool perturb = sp.perturb != 0;
    const Jitter jit = jitter_for(sp, r);
    const float u_i = z_uniform(R.near, R.far, i, S, sp.lin_step);
    const float u_n = i + 1 < S ? z_uniform(R.near, R.far, i + 1, S, sp.lin_step) : u_i;
    z0 = u_i;
    if (perturb) {
        float lower = u_i, upper = u_i;
        if (i > 0) lower = __fmul_rn(0.5f, __fadd_rn(u_i, z_uniform(R.near, R.far, i - 1, S, sp.lin_step)));
        if (i + 1 < S) upper = __fmul_rn(0.5f, __fadd_rn(u_n, u_i));
        z0 = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), jitter_uniform(jit, i)));
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) x[d] = ray_point(R.o[d], R.d[d], z0, sp.clamp);
    delta = 0.f;
    if (want_delta) {
        float dist = 1e-10f;
        if (i + 1 < S) {
            float z1 = u_n;
            if (perturb) {
                const float lower = __fmul_rn(0.5f, __fadd_rn(u_n, u_i));
                float upper = u_n;
                if (i + 2 < S) upper = __fmul_rn(0.5f, __fadd_rn(z_uniform(R.near, R.far, i + 2, S, sp.lin_step), u_n));
                z1 = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), jitter_uniform(jit, i + 1)));
            }
            dist = __fsub_rn(z1, z0);
        }
        delta = __fmul_rn(dist, R.norm);
    }
}

// normalisation of HashEncoder.forward (hashgrid.py:125) as ATen's CUDA kernels evaluate it:
// (x + size) is one add; "/ (2*size)" with a python-scalar divisor is a multiply by fl(1/fl(2*size)).
__device__ __forceinline__ float normalise01(float x, float bound, float inv_2bound) {
    return __fmul_rn(__fadd_rn(x, bound), inv_2bound);
}

#endif  // __CUDACC__

// host: ABI struct -> kernel parameter block
static inline int nafb_make_sampler_params(const nafb_sampler *s, int src, SamplerParams *out, uint64_t *n_points) {
    if (!s) NAFB_FAIL(NAFB_ERR_INVALID, "nafb_sampler: null");
    SamplerParams p{};
    p.pts = s->pts; p.rays = s->rays; p.t_rand = s->t_rand; p.rng_state = s->rng_state;
    p.poses = s->poses; p.pixels = s->pixels;
    p.det_w2 = (float)((double)s->det_w / 2.0); p.det_h2 = (float)((double)s->det_h / 2.0);
    p.det_du = s->det_du; p.det_dv = s->det_dv; p.det_u0 = s->det_u0; p.det_v0 = s->det_v0; p.det_dsd = s->det_dsd;
    p.det_near = s->det_near; p.det_far = s->det_far; p.det_parallel = s->det_parallel;
    p.n_rays = s->n_rays; p.n_samples = s->n_samples; p.perturb = s->perturb;
    p.n1 = s->n1; p.n2 = s->n2; p.n3 = s->n3; p.i0 = s->i0; p.i1 = s->i1;
    p.s1 = s->s1; p.s2 = s->s2; p.s3 = s->s3;
    p.vstep1 = s->n1 > 1 ? (s->s1 - (-s->s1)) / (double)(s->n1 - 1) : 0.0;
    p.vstep2 = s->n2 > 1 ? (s->s2 - (-s->s2)) / (double)(s->n2 - 1) : 0.0;
    p.vstep3 = s->n3 > 1 ? (s->s3 - (-s->s3)) / (double)(s->n3 - 1) : 0.0;
    p.bound = s->bound;
    p.inv_2bound = 1.0f / (2.0f * s->bound);   // see normalise01: fl(1 / fl(2*size)); the doubling is exact
    p.clamp = s->clamp;
    p.lin_step = s->n_samples > 1 ? 1.0f / (float)(s->n_samples - 1) : 0.f;
    uint64_t n = 0;
    switch (src) {
        case NAFB_SRC_POINTS:
            n = s->n_points;
            if (n && !s->pts) NAFB_FAIL(NAFB_ERR_INVALID, "nafb_sampler: pts is null");
            break;
        case NAFB_SRC_RAYS:
            n = (uint64_t)s->n_rays * s->n_samples;
            if (n && !s->rays && !(s->pixels && s->poses)) NAFB_FAIL(NAFB_ERR_INVALID, "nafb_sampler: rays is null (and no pixels + poses to generate them from)");
            if (s->rays) p.pixels = nullptr;   // an explicit rays tensor wins
            if (n && s->perturb && !s->t_rand && !s->rng_state)
                NAFB_FAIL(NAFB_ERR_INVALID, "nafb_sampler: perturb needs t_rand (or rng_state for in-kernel uniforms)");
            break;
        case NAFB_SRC_VOXELS:
            if (s->i1 > s->n1 || s->i0 > s->i1) NAFB_FAIL(NAFB_ERR_INVALID, "nafb_sampler: bad voxel slab");
            n = (uint64_t)(s->i1 - s->i0) * s->n2 * s->n3;
            break;
        default:
            NAFB_FAIL(NAFB_ERR_INVALID, "nafb_sampler: unknown point source %d", src);
    }
    *out = p;
    *n_points = n;
    return NAFB_OK;
}
