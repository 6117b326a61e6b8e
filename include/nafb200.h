/*
 * nafb200.h -- C ABI of libnafb200.so, the B200 (sm_100a) hot path of NAF
 * (Neural Attenuation Fields: ray sampling -> multi-resolution hash-grid encoding ->
 * density MLP -> Beer-Lambert line integral -> L2 projection loss -> backward -> Adam,
 * plus the forward-only full-volume voxel query).
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  Every `const float*` etc. is a
 *     DEVICE pointer unless the parameter name starts with `h_` (host pointer).
 *   - the caller owns every buffer (outputs and workspaces included); nothing is allocated
 *     here, nothing is retained after the call returns (kernels are enqueued on `stream`).
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *   - return value: NAFB_OK (0) or an error code; nafb_last_error() returns the message of
 *     the last failure on the calling thread (thread-local).  There is no CPU fallback.
 *   - "reference" citations are file:line under the reference repository
 *     (holuca/NeuralVolumetricReconstructionForMedicalImages).
 */
#ifndef NAFB200_H
#define NAFB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NAFB_ABI_VERSION 10

enum nafb_status { NAFB_OK = 0, NAFB_ERR_INVALID = 1, NAFB_ERR_UNSUPPORTED = 2, NAFB_ERR_CUDA = 3 };
/* head activation of the density MLP: reference src/network/network.py:23-32 */
enum nafb_activation { NAFB_ACT_SIGMOID = 0, NAFB_ACT_LRELU = 1, NAFB_ACT_TANH = 2, NAFB_ACT_NONE = 3 };
/* layout of the encoder output / incoming gradient */
enum nafb_layout { NAFB_LAYOUT_LBC = 0 /* [L,B,C], reference FFI (hashencoder.cu:96) */,
                   NAFB_LAYOUT_BLC = 1 /* [B,L*C], what hashgrid.py:40 permutes to   */ };

#define NAFB_MAX_LEVELS 32
#define NAFB_MAX_LAYERS 8

typedef void *nafb_stream_t;

int nafb_abi_version(void);
const char *nafb_last_error(void);
/* SM count / compute capability of the current device (used to size persistent grids). */
int nafb_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------ hash-grid descriptor
 * Mirrors HashEncoder's state (reference src/encoder/hashencoder/hashgrid.py:77-113):
 * `h_offsets` are ENTRY offsets per level (L+1 values, host memory), `table` is the
 * embeddings parameter [h_offsets[L], C]. */
typedef struct nafb_grid {
    const float *table;       /* device, [n_entries, C] fp32 */
    const int32_t *h_offsets; /* HOST, [L+1] */
    uint32_t D;               /* 2 or 3 */
    uint32_t C;               /* 1, 2, 4 or 8 */
    uint32_t L;               /* 1..NAFB_MAX_LEVELS */
    uint32_t H;               /* base resolution */
} nafb_grid;

/* Replaces _backend.hash_encode_forward (reference src/encoder/hashencoder/src/bindings.cpp:6,
 * hashencoder.h:13, hashencoder.cu:373-396).
 *   inputs  [B, D] in [0,1]; outputs [L,B,C] (layout LBC) or [B,L*C] (layout BLC), overwritten;
 *   dy_dx   [B, L*D*C] when calc_grad_inputs != 0 (overwritten), else ignored (may be NULL).
 * Errors: D not in {2,3} or C not in {1,2,4,8} -> NAFB_ERR_UNSUPPORTED with the reference's
 * message "GridEncoding: C must be 1, 2, 4, or 8." (hashencoder.cu:310,324). */
int nafb_hash_encode_forward(const nafb_grid *grid, const float *inputs, float *outputs, uint32_t B,
                             int out_layout, int calc_grad_inputs, float *dy_dx, nafb_stream_t stream);

/* Replaces _backend.hash_encode_backward (bindings.cpp:7, hashencoder.h:14, hashencoder.cu:398-428).
 *   grad [B, L*C] (layout BLC, what autograd hands over) or [L,B,C];
 *   grad_table [n_entries, C]: ACCUMULATED into (caller pre-zeroes, hashgrid.py:59);
 *   grad_inputs [B, D]: accumulated into when calc_grad_inputs != 0 (hashencoder.cu:295). */
int nafb_hash_encode_backward(const nafb_grid *grid, const float *grad, const float *inputs,
                              float *grad_table, uint32_t B, int grad_layout, int calc_grad_inputs,
                              const float *dy_dx, float *grad_inputs, nafb_stream_t stream);

/* The same two operations for the other storage types the reference's FFI dispatches on
 * (AT_DISPATCH_FLOATING_TYPES_AND_HALF, hashencoder.cu:392,423): every tensor -- table, inputs, outputs, dy_dx, grad,
 * grad_table, grad_inputs -- has the element type `dtype`; `table` / `grad_table` replace grid->table, which is ignored.
 * fp16 is what `@custom_fwd(cast_inputs=torch.half)` (hashgrid.py:12) hands the op under autocast; its table gradient
 * leaves as half2 reductions (hashencoder.cu:257-263).  NAFB_DTYPE_F32 forwards to the functions above. */
enum nafb_dtype { NAFB_DTYPE_F32 = 0, NAFB_DTYPE_F16 = 1, NAFB_DTYPE_F64 = 2 };
int nafb_hash_encode_forward_dtype(const nafb_grid *grid, int dtype, const void *table, const void *inputs, void *outputs,
                                   uint32_t B, int out_layout, int calc_grad_inputs, void *dy_dx, nafb_stream_t stream);
int nafb_hash_encode_backward_dtype(const nafb_grid *grid, int dtype, const void *grad, const void *inputs, void *grad_table,
                                    uint32_t B, int grad_layout, int calc_grad_inputs, const void *dy_dx, void *grad_inputs,
                                    nafb_stream_t stream);

/* One fused pass min/max over a float buffer -> out2[0] = min, out2[1] = max
 * (the range check of hashgrid.py:122 without two separate reductions). */
int nafb_minmax(const float *x, uint64_t n, float *out2, nafb_stream_t stream);

/* ------------------------------------------------------------------ density network
 * Mirrors DensityNetwork (reference src/network/network.py:5-58): encoder -> n_layers Linear,
 * LeakyReLU(0.01) after every hidden layer, skip layers take cat([encoding, h]), head act. */
typedef struct nafb_mlp {
    uint32_t n_layers;                  /* 2..NAFB_MAX_LAYERS (hidden layers + output layer) */
    uint32_t in_dim;                    /* encoder.output_dim = L*C */
    uint32_t hidden;                    /* hidden_dim */
    uint32_t out_dim;                   /* 1 */
    uint32_t skip_mask;                 /* bit i set <=> layer i takes cat([enc, h]) */
    uint32_t head;                      /* enum nafb_activation */
    uint32_t arith;                     /* enum nafb_arith: how the fused density kernels evaluate the layers (per call; the
                                           library keeps no mode of its own) */
    const float *W[NAFB_MAX_LAYERS];    /* device, nn.Linear weight [out, in] row-major */
    const float *b[NAFB_MAX_LAYERS];    /* device, [out] */
} nafb_mlp;

/* Arithmetic of the fused density kernels (nafb_mlp.arith):
 *   NAFB_ARITH_TC   tcgen05 tensor cores with bf16x3 split operands (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM)
 *                   wherever the configuration allows (L*C == 32, 4 x 32 MLP, skip at 2, out_dim 1); other shapes run SIMT;
 *   NAFB_ARITH_SIMT fp32 FMAs everywhere (bit-reproducible dot-product order; the general-shape path). */
enum nafb_arith { NAFB_ARITH_TC = 0, NAFB_ARITH_SIMT = 1 };

typedef struct nafb_mlp_grads {
    float *gW[NAFB_MAX_LAYERS]; /* device, same shapes as W; ACCUMULATED into */
    float *gb[NAFB_MAX_LAYERS];
} nafb_mlp_grads;

/* Where a kernel takes its sample points from. */
enum nafb_point_source {
    NAFB_SRC_POINTS = 0, /* pts [P,3] world coordinates in [-bound, bound]            */
    NAFB_SRC_RAYS = 1,   /* rays [N,8] + sampling of render.py:88-105, P = N*n_samples */
    NAFB_SRC_VOXELS = 2  /* voxel lattice of tigre.py:388-400, P = (i1-i0)*n2*n3       */
};

/* Device-resident state of the fused training step (8 x uint32): everything that changes from step to step, so that a
 * captured CUDA graph of the whole iteration can be replayed without touching kernel arguments. */
#define NAFB_STATE_STEP 0     /* completed optimizer steps (incremented by nafb_adam_step_dev / nafb_adam_exchange_step) */
#define NAFB_STATE_SEED_LO 1  /* seed of the in-kernel sampler generator */
#define NAFB_STATE_SEED_HI 2
#define NAFB_STATE_LR 6       /* learning rate: the bits of a DOUBLE in words 6 (low) and 7 (high) -- torch.optim.Adam evaluates
                                 lr / (1 - beta1^t) in python floats, i.e. doubles; a float would change the last bit of the step */
#define NAFB_STATE_TICKET 4   /* internal (zero between launches): last-block detection of the optimizer kernels */
#define NAFB_STATE_TICKET_FWD 5 /* internal (zero between launches): a word the engine hands nafb_loss_tail.ticket */
#define NAFB_STATE_WORDS 8

typedef struct nafb_sampler {
    /* NAFB_SRC_POINTS */
    const float *pts;        /* [P,3] */
    uint64_t n_points;       /* P */
    /* NAFB_SRC_RAYS: reference src/render/render.py:88-105 */
    const float *rays;       /* [N,8] = origin(3), direction(3), near, far */
    const float *t_rand;     /* [N,S] uniforms in [0,1) (torch.rand of render.py:99); NULL when !perturb or with rng_state */
    const uint32_t *rng_state; /* device nafb_step_state: with perturb and t_rand == NULL the uniforms are generated
                                  in-kernel from (seed, step, ray, sample) -- identical in forward and backward */
    uint32_t n_rays;
    uint32_t n_samples;
    int32_t perturb;
    /* NAFB_SRC_RAYS with rays == NULL: the rays are GENERATED in-kernel from detector pixels (reference
     * src/dataset/tigre.py:402-456 get_rays, :463-528 get_rays2, poses of :530-572), bit-identical to the reference's
     * fp32 torch arithmetic:  uu = ((col + .5) - W/2) * du + u0,  cone: d = R.[uu/DSD, vv/DSD, 1], o = t;
     * parallel: d = R.[0,0,1], o = R.[uu,vv,0] + t;  near / far = the global pair of tigre.py:575-586. */
    const float *poses;      /* [n_proj, 12] fp32: rotation (3x3 row-major) then translation, fp32 cast of angle2pose */
    const int32_t *pixels;   /* [N, 3]: projection index, detector row, detector column */
    uint32_t det_w, det_h;   /* nDetector[0], nDetector[1] */
    float det_du, det_dv;    /* dDetector (metres) */
    float det_u0, det_v0;    /* offDetector (metres) */
    float det_dsd;           /* source-detector distance (metres); cone beam only */
    float det_near, det_far;
    int32_t det_parallel;    /* 0: cone beam, 1: parallel beam */
    /* NAFB_SRC_VOXELS: reference src/dataset/tigre.py:388-400 */
    uint32_t n1, n2, n3;     /* full lattice */
    uint32_t i0, i1;         /* slab [i0, i1) of the outermost index handled by this call */
    double s1, s2, s3;       /* half extents sVoxel/2 - dVoxel/2 (float64 linspace end points) */
    /* common */
    float bound;             /* net.bound: positions are normalised as (x + bound) * (1/(2 bound)) */
    float clamp;             /* render.py:104: fp32(bound - 1e-6); RAYS source clamps to +-clamp */
} nafb_sampler;

/* Encoding stash (optional, training only).  A forward pass that is going to be followed by
 * nafb_density_backward on the SAME grid / mlp / sampler may leave the encodings of its points in
 * `stash` (nafb_density_stash_bytes() bytes of device memory, 128 B per point rounded up to 128-point
 * tiles); the backward pass then reads them back with full-line loads instead of repeating the
 * 8-corner gather of every level.  Returns 0 when the current arithmetic mode / configuration does
 * not use a stash (pass NULL then; NULL is always valid and means "recompute"). */
uint64_t nafb_density_stash_bytes(const nafb_grid *grid, const nafb_mlp *mlp, uint64_t n_points);

/* sigma[P*out_dim] = DensityNetwork.forward(points)   (network.py:34-58; forward only).
 * With src == NAFB_SRC_RAYS and acc != NULL it also integrates (render.py:192-201):
 *   acc[r] += sum_i sigma[r,i] * (z[r,i+1]-z[r,i]) * |d_r|   (last delta 1e-10; caller pre-zeroes acc)
 * and writes z_vals [N,S] / pts [N,S,3] when those pointers are non-NULL.
 * flags[0] is OR-ed with 1 if a position leaves [-bound, bound] (hashgrid.py:122), 2 on NaN/Inf. */
int nafb_density_forward(const nafb_grid *grid, const nafb_mlp *mlp, const nafb_sampler *smp, int src,
                         float *sigma, float *acc, float *z_vals, float *pts_out, int32_t *flags,
                         void *stash, nafb_stream_t stream);

/* Training forward of one ray batch WITH the loss in the same launch (train.py:69-127 after render.py:31): as
 * nafb_density_forward(src = NAFB_SRC_RAYS, sigma = z_vals = pts_out = NULL), and the LAST CTA of the kernel to retire
 * evaluates nafb_mse_loss(acc, target, mask, n_rays, chunk, gscale, loss_out, dacc, zero_pred) -- the step needs no
 * separate loss launch.  `ticket`: one device word, zero before the first use (the kernel leaves it zero).
 * Tensor-core configurations only (NAFB_ERR_UNSUPPORTED otherwise: launch nafb_density_forward + nafb_mse_loss). */
typedef struct nafb_loss_tail {
    const float *target;   /* [n_rays] measured projections                                   */
    const uint8_t *mask;   /* [n_rays] or NULL                                                */
    uint32_t chunk;        /* 0: one chunk                                                    */
    float gscale;
    float *loss_out;       /* [2]: loss, number of valid rays (device or device-mapped host)  */
    float *dacc;           /* [n_rays] d loss / d acc                                         */
    int32_t zero_pred;     /* as in nafb_mse_loss                                             */
    uint32_t *ticket;
    /* optional completion flag (e.g. in pinned, device-mapped host memory): once loss_out is written and visible system-wide the
     * kernel stores step_state[NAFB_STATE_STEP] + 1 there -- a host thread polling the flag reads the loss of the step without
     * waiting for the backward pass and the optimizer that follow on the stream.  Both NULL: no flag. */
    uint32_t *done_flag;
    const uint32_t *step_state;
} nafb_loss_tail;
int nafb_density_forward_loss(const nafb_grid *grid, const nafb_mlp *mlp, const nafb_sampler *smp, float *acc,
                              int32_t *flags, void *stash, const nafb_loss_tail *loss, nafb_stream_t stream);

/* Backward of the above.  dsigma [P] (src POINTS) or dacc [N] (src RAYS: dsigma is derived as
 * dacc[r] * delta[r,i] in-kernel).  Recomputes the forward activations from the points (and the
 * stash, when given), accumulates MLP gradients and scatters into grad_table.
 *   workspace: nafb_density_backward_workspace_bytes() bytes of device scratch, ZERO-FILLED before its first use and not shared by
 *   launches that may run concurrently (its tail holds the two words of the backward kernel's grid barrier, which the kernel
 *   leaves zero again). */
uint64_t nafb_density_backward_workspace_bytes(const nafb_mlp *mlp);
int nafb_density_backward(const nafb_grid *grid, const nafb_mlp *mlp, const nafb_sampler *smp, int src,
                          const float *dsigma_or_dacc, float *grad_table, const nafb_mlp_grads *grads,
                          void *workspace, const void *stash, nafb_stream_t stream);

/* ------------------------------------------------------------------ sampling / integral (unfused API)
 * nafb_sample_points: render.py:88-105 -> z_vals [N,S], pts [N,S,3]; tv_partial [N] gets
 * sum_i |pts[r,i+1]-pts[r,i]|_1 per ray (render.py:16-28) when non-NULL. */
int nafb_sample_points(const nafb_sampler *smp, float *z_vals, float *pts, float *tv_partial,
                       nafb_stream_t stream);

/* Hierarchical ("fine") sampling of a ray batch in one kernel (render.py:113-126 with sample_pdf :215-247): per ray the
 * piecewise-constant pdf of the coarse weights[1:-1] (+ 1e-5) over the midpoints of z_vals is inverted at n_fine uniforms
 * (searchsorted right=True, denominators below 1e-5 replaced by 1), the new depths are merged with the coarse ones (sorted
 * ascending) and the clamped sample positions of all S + n_fine depths are written.
 *   rays [N,8], z_vals [N,S] ascending, weights [N,S]; u [N,u_stride >= n_fine] uniforms in [0,1) -- or u_stride == 0: one row
 *   [n_fine] shared by all rays (the reference's deterministic linspace(0,1,n_fine) when perturb == 0);
 *   z_out [N,S+n_fine], pts_out [N,S+n_fine,3], tv_partial [N] = sum_i |pts[i+1]-pts[i]|_1 (render.py:16-28); any output may be NULL.
 * n_samples >= 3, n_samples + n_fine <= 1024 (NAFB_ERR_UNSUPPORTED otherwise). */
int nafb_sample_fine(const float *rays, const float *z_vals, const float *weights, const float *u, uint32_t u_stride,
                     uint32_t n_rays, uint32_t n_samples, uint32_t n_fine, float clamp, float *z_out, float *pts_out,
                     float *tv_partial, nafb_stream_t stream);
/* rays_out [N,8] for the detector pixels of the sampler (the in-kernel generator written out). */
int nafb_generate_rays(const nafb_sampler *smp, float *rays_out, nafb_stream_t stream);
/* raw2outputs (render.py:178-212), raw [N,S,out_dim] (channel 0 integrated): acc [N];
 * absdiff [N,S] = (1e-10, |raw_i - raw_{i-1}|) when non-NULL (the un-normalised `weights`). */
int nafb_ray_integral_forward(const float *raw, uint32_t out_dim, const float *z_vals, const float *rays,
                              float *acc, float *absdiff, uint32_t n_rays, uint32_t n_samples,
                              nafb_stream_t stream);
/* d raw[r,i,0] = dacc[r] * delta[r,i] (other channels zero). */
int nafb_ray_integral_backward(const float *dacc, uint32_t out_dim, const float *z_vals, const float *rays,
                               float *draw, uint32_t n_rays, uint32_t n_samples, nafb_stream_t stream);

/* ------------------------------------------------------------------ loss
 * Masked, chunk-wise MSE of train.py:69-127 / loss.py:26-46:
 *   loss = sum_chunks mean_{r in chunk, mask[r]} (target[r]-pred[r])^2 ;  dpred[r] = dloss/dpred[r] * gscale.
 * mask may be NULL (all rays valid); chunk == 0 means one chunk.  loss_out[0] = loss,
 * loss_out[1] = number of valid rays.  Single deterministic block (one warp per chunk, chunk means
 * added in chunk order).  zero_pred != 0: pred is cleared once consumed (the fused engine accumulates
 * the next step's projections into the same buffer with atomics). */
int nafb_mse_loss(float *pred, const float *target, const uint8_t *mask, uint32_t n, uint32_t chunk,
                  float gscale, float *loss_out, float *dpred, int zero_pred, nafb_stream_t stream);

/* ------------------------------------------------------------------ per-iteration dataset work on the device
 * nafb_ptycho_mask: get_ptycho_mask (reference src/utils/util.py:196-205; train.py:59-60 recomputes it every iteration) for a
 * whole scan at once: full_proj [n_proj, H, W] complex64 (interleaved re, im), keep [n_proj, H, W] = 1 where the pixel is kept. */
int nafb_ptycho_mask(const float *full_proj, uint32_t n_proj, uint32_t H, uint32_t W, float threshold, uint8_t *keep,
                     nafb_stream_t stream);

/* nafb_draw_pixels: TIGREDataset.__getitem__ (reference src/dataset/tigre.py:354-382) -- n_rays of the non-zero pixels of ONE
 * projection, uniformly without replacement and in random order (np.random.choice(replace=False), :358), their projection
 * values (:364) and mask bits (train.py:93-95), written where the fused training step reads its batch.
 *   valid [n_proj, H*W]: per projection the flat indices (row * W + col) of its non-zero pixels, front-packed; n_valid [n_proj];
 *   order: projection of draw k = order[k % n_order] (NULL: k % n_proj, the order of the reference's un-shuffled DataLoader);
 *   draw_state [4] (device): draw counter (incremented by the launch), seed lo, seed hi, error word (1 + projection when a
 *   projection has fewer than n_rays valid pixels -- the reference's np.random.choice raises there).
 * No argument changes between launches: the draw can sit in the CUDA graph of the training step.  n_rays <= 8192. */
typedef struct nafb_pixel_source {
    const float *projs;      /* [n_proj, H, W] */
    const uint8_t *mask;     /* [n_proj, H, W] keep-mask, or NULL (all kept) */
    const int32_t *valid;
    const int32_t *n_valid;
    const int32_t *order;
    uint32_t n_proj, H, W, n_order;
} nafb_pixel_source;
int nafb_draw_pixels(const nafb_pixel_source *src, uint32_t n_rays, int32_t *pixels_out, float *projs_out, uint8_t *mask_out,
                     uint32_t *draw_state, nafb_stream_t stream);

/* ------------------------------------------------------------------ evaluation metrics (reference train.py:253-258)
 * Float64 arithmetic like the reference's numpy code; both kernels leave one partial sum per block (n_partial blocks of 256
 * threads, summed in block order by the caller: deterministic), nothing is copied to the host but those few doubles.
 * nafb_sqdiff_f64: sum (a - b)^2  ->  get_psnr_3d = 20 log10(PIXEL_MAX / sqrt(sum / n))           (src/utils/util.py:55-84).
 * nafb_ssim3d_f64: sum over the interior of the local SSIM of two [n1,n2,n3] volumes -- uniform win^3 window, sample covariance,
 * K1 0.01, K2 0.03, C = (K data_range)^2: what skimage.metrics.structural_similarity computes for a 3-D float array without a
 * channel axis; get_ssim_3d = sum / ((n1-win+1)(n2-win+1)(n3-win+1))                                  (src/utils/util.py:87-139). */
int nafb_sqdiff_f64(const float *a, const float *b, uint64_t n, double *partial, uint32_t n_partial, nafb_stream_t stream);
int nafb_ssim3d_f64(const float *a, const float *b, uint32_t n1, uint32_t n2, uint32_t n3, uint32_t win, double data_range,
                    double *partial, uint32_t n_partial, nafb_stream_t stream);

/* ------------------------------------------------------------------ optimiser
 * torch.optim.Adam (trainer.py:54: betas (0.9,0.999), eps 1e-8, no weight decay, no amsgrad),
 * one fused pass over a flat parameter vector; grad is zeroed in the same pass when zero_grad != 0
 * (trainer.py:138 optimizer.zero_grad()).  `step` is the 1-based step count.  The hyper-parameters are DOUBLES, like the python
 * floats torch.optim.Adam computes its scalar constants from: with them the update is bit-identical to torch's (tested). */
int nafb_adam_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq, uint64_t n, double lr,
                   double beta1, double beta2, double eps, uint32_t step, float grad_scale, int zero_grad,
                   nafb_stream_t stream);

/* ------------------------------------------------------------------ multi-GPU exchange step
 * The reference has no distributed code; the engine shards rays over ranks and needs ONE
 * exchange per iteration: sum the flat gradient over ranks, Adam, identical parameters everywhere.
 * nafb_adam_exchange_step does all of it in one kernel over NVLink peer memory: rank r owns the slice
 * nafb_exchange_slice(n, r, W) of the flat vector, loads that slice of every rank's gradient (P2P), adds them in rank
 * order, applies Adam (state for the slice only) and stores the new parameters into every rank's replica.
 *
 * Buffers that peers touch (parameters, gradients, flag blocks) must come from nafb_peer_alloc (cudaMalloc + CUDA IPC
 * handle, zero-filled); a peer maps them with nafb_peer_open(handle).  Flag block: NAFB_XFLAG_WORDS uint32, zero before
 * the first step.  Gradients are double buffered by the caller: `grad[w]` = this step's buffer of rank w,
 * `grad_zero` = the LOCAL buffer of the other parity (cleared by the kernel; NULL to skip).  `step` (1-based, strictly
 * increasing, the same on every rank) doubles as the synchronisation epoch.  A lost peer raises flag word
 * NAFB_XFLAG_ERROR of the local block after a bounded spin instead of hanging the GPU. */
#define NAFB_MAX_RANKS 8
#define NAFB_XFLAG_ARRIVE 0   /* [0..7]  written by rank i: "my gradient of this epoch is complete"            */
#define NAFB_XFLAG_DONE 8     /* [8..15] written by rank i: "I have finished writing your parameters"          */
#define NAFB_XFLAG_ERROR 16   /* != 0: a spin timed out (value - 1 = index of the flag that never arrived)     */
#define NAFB_XFLAG_TICKET 17  /* local block counter                                                           */
#define NAFB_XFLAG_TICKET2 18 /* local block counter of the push edition's second phase                        */
#define NAFB_XFLAG_WORDS 32

typedef struct nafb_exchange {
    uint32_t world, rank;
    float *param[NAFB_MAX_RANKS];      /* every rank's flat parameters [n] (param[rank] is local)               */
    float *grad[NAFB_MAX_RANKS];       /* every rank's flat gradient [n] of this step's parity                  */
    uint32_t *flags[NAFB_MAX_RANKS];   /* every rank's flag block [NAFB_XFLAG_WORDS]                            */
    float *grad_zero;                  /* local gradient buffer of the other parity, or NULL                    */
    float *exp_avg, *exp_avg_sq;       /* LOCAL, slice-sized: [i1 - i0] of nafb_exchange_slice(n, rank, world)  */
    uint64_t n;                        /* floats, multiple of 4                                                 */
    uint32_t *state;                   /* optional device nafb_step_state: when non-NULL the epoch is state[0] + 1 and
                                          the learning rate the double in state[6..7] (the `step` / `lr` arguments are ignored), and
                                          the kernel increments state[0]: the launch can sit in a replayed CUDA graph */
    float *mc_param;                   /* optional NVLS (NVLink SHARP) multicast address of the parameter vector ...    */
    const float *mc_grad;              /* ... and of this step's gradient: the slice sum is then one multimem.ld_reduce
                                          per element (added inside the NVSwitch) and the new parameters one multimem.st
                                          (multicast to every replica): ~ n*4 bytes per GPU and direction instead of
                                          2*(W-1)/W * n*4.  Both NULL: plain P2P loads / stores.                        */
    float *stage[NAFB_MAX_RANKS];      /* optional PUSH edition (all non-NULL): every rank's staging area, [W][stage_slot]
                                          floats in peer memory.  Ranks then push their gradient slices into the owners'
                                          staging areas and the owners push the parameters back: stores only over NVLink
                                          (posted writes run faster than remote reads).  grad[rank] is then the ONE gradient
                                          buffer of this rank (peers never read it; it leaves the kernel zeroed), grad_zero
                                          and grad[w != rank] are ignored.                                              */
    uint64_t stage_slot;               /* floats per staging slot, multiple of 4, >= the largest slice                  */
} nafb_exchange;

int nafb_peer_alloc(uint64_t bytes, void **ptr, unsigned char *handle64);
int nafb_peer_open(const unsigned char *handle64, void **ptr);
int nafb_peer_close(void *ptr);
int nafb_peer_free(void *ptr);
/* [i0, i1) in floats (multiples of 4) of the slice rank `rank` owns */
int nafb_exchange_slice(uint64_t n, uint32_t rank, uint32_t world, uint64_t *i0, uint64_t *i1);
int nafb_adam_exchange_step(const nafb_exchange *x, double lr, double beta1, double beta2, double eps, uint32_t step,
                            float grad_scale, nafb_stream_t stream);

/* Same optimizer step with the step count and learning rate taken from a device nafb_step_state: (the bias corrections
 * are evaluated on the device in double precision, as torch does on the host); state[0] is incremented when the last
 * block retires.  Graph-capturable: no argument changes from step to step. */
int nafb_adam_step_dev(float *param, float *grad, float *exp_avg, float *exp_avg_sq, uint64_t n, double beta1, double beta2,
                       double eps, float grad_scale, int zero_grad, uint32_t *state, nafb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NAFB200_H */
