/*
 * dad_b200.h — C ABI of the B200-native reverse-diffusion sampler.
 *
 * This is the drop-in boundary underneath the reference's Python classes.  The
 * reference has no FFI of its own (it is pure PyTorch); every entry point below
 * names the reference method whose device work it replaces (paths relative to
 * the reference root).  Plain C: opaque handle, raw pointers, sizes, a CUDA
 * stream passed as void*.  No torch types cross this line.
 *
 * Conventions
 *   - every function returns DAD_OK (0) or a negative dad_status; nothing throws.
 *     dad_last_error() returns a human-readable message for the last failure.
 *   - "device pointer" = memory of the handle's CUDA device.  Setup calls
 *     (weights, schedule, projector, conditions) accept host OR device pointers
 *     (copied with cudaMemcpyDefault); the per-step / per-loop data pointers must
 *     be device pointers except in dad_sample_host.
 *   - trajectories are fp32, row-major (B, H, T) exactly like the reference's
 *     (batch, horizon, transition_dim) tensors.
 *   - a handle is bound to one device and is not thread-safe.  All work is
 *     enqueued on the stream the caller passes; calls return without
 *     synchronising unless stated.
 *   - there is no CPU fallback: on a machine without an sm_100 device
 *     dad_create fails with DAD_ERR_DEVICE.
 */
#ifndef DAD_B200_H
#define DAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAD_ABI_VERSION 1
#define DAD_MAX_LEVELS 8

typedef enum {
  DAD_OK = 0,
  DAD_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  DAD_ERR_DEVICE = -2,      /* no usable sm_100 device, or wrong device */
  DAD_ERR_CUDA = -3,        /* a CUDA runtime/driver call failed */
  DAD_ERR_STATE = -4,       /* call sequence error (e.g. sampling before weights are loaded) */
  DAD_ERR_NOMEM = -5
} dad_status;

typedef enum {
  DAD_PRECISION_FP32 = 0,   /* SIMT fp32 everywhere: the 1e-5 parity mode */
  DAD_PRECISION_BF16 = 1    /* tcgen05 implicit-GEMM convs, bf16 operands, fp32 accumulate */
} dad_precision;

/* flags for dad_step / dad_sample */
#define DAD_FLAG_CONDITIONS            1u   /* inpaint the registered conditions (policies.py:109-110) */
#define DAD_FLAG_PROJECT               2u   /* apply the registered projector every step */
#define DAD_FLAG_PROJECT_AFTER_INPAINT 4u   /* order: denoise -> inpaint -> project (default: project -> inpaint) */
#define DAD_FLAG_PHILOX_INIT           8u   /* dad_sample draws x_S itself (ignores the contents of x) */

typedef struct dad_handle dad_handle;

/* Architecture + process configuration.
 * Mirrors TemporalUnet.__init__ (m_diffuser/models/temporal_unet.py:135-140) and
 * GaussianDiffusion.__init__ (m_diffuser/models/diffusion.py:62-71). */
typedef struct {
  int32_t abi_version;          /* DAD_ABI_VERSION */
  int32_t device;               /* CUDA device ordinal */
  int32_t precision;            /* dad_precision */
  int32_t transition_dim;       /* T = observation_dim + action_dim */
  int32_t dim;                  /* base width */
  int32_t n_levels;             /* len(dim_mults) */
  int32_t dim_mults[DAD_MAX_LEVELS];
  int32_t kernel_size;          /* odd, <= 7 (reference default 5) */
  int32_t time_dim;             /* 0 -> dim (temporal_unet.py:154) */
  int32_t horizon;              /* H; must be divisible by 2^(n_levels-1) */
  int32_t n_timesteps;          /* length of the schedule tables (len(betas)) */
  int32_t predict_epsilon;      /* diffusion.py:192-197 */
  int32_t clip_denoised;        /* diffusion.py:199-200 */
  int32_t max_batch;            /* workspace capacity in samples; larger batches are processed in chunks */
} dad_config;

/* One fp32 tensor of the U-Net state_dict; `name` is the reference's key relative to
 * the TemporalUnet module, e.g. "downs.0.0.blocks.0.block.0.weight"
 * (layout: SURVEY.md 8(a10); temporal_unet.py:135-197). */
typedef struct {
  const char *name;
  const float *data;            /* host or device, contiguous, torch layout */
  int64_t numel;
} dad_tensor;

/* Lifetime. Replaces nothing in the reference (module construction). */
int dad_create(const dad_config *cfg, dad_handle **out);
int dad_destroy(dad_handle *h);
/* h may be NULL: returns the message of the last failed dad_create on this thread. */
const char *dad_last_error(const dad_handle *h);
int dad_abi_version(void);

/* nn.Module.load_state_dict for the U-Net (checkpoint layout: utils/training.py:191-224).
 * Re-packs the weights for the selected precision and precomputes the per-step
 * time-embedding tables (temporal_unet.py:26-32,155-160,97-100 evaluated for every
 * step index 0..n_timesteps-1).  All tensors of the architecture must be present. */
int dad_load_weights(dad_handle *h, const dad_tensor *tensors, int32_t n_tensors);

/* The five per-step coefficient buffers GaussianDiffusion registers (diffusion.py:109-128):
 * sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1,
 * posterior_mean_coef2, posterior_log_variance_clipped; each of length n_timesteps. */
int dad_set_schedule(dad_handle *h, const float *sqrt_recip, const float *sqrt_recipm1,
                     const float *coef1, const float *coef2, const float *log_variance,
                     int32_t n_timesteps);

/* DynamicsAwarePolicy.apply_projection (guides/policies.py:409-485) folded into one affine map on
 * the flattened normalised (H*T) trajectory:  y = x + alpha[i] * (Nmat x + q)  with
 * Nmat = M_P - I (row-major D x D, D = H*T), alpha per step index (policies.py:358-383).
 * Nmat == NULL clears the projector. */
int dad_set_projector(dad_handle *h, const float *Nmat, const float *q, const float *alpha,
                      int32_t D, int32_t n_timesteps);

/* GuidedPolicy.apply_conditions (guides/policies.py:48-63): x[:, h_idx[c], :] = vals[c].
 * vals is (n_cond, T) when per_batch == 0 (broadcast over the batch) or (n_cond, B, T).
 * n_cond == 0 clears. */
int dad_set_conditions(dad_handle *h, const int32_t *h_idx, const float *vals, int32_t n_cond,
                       int32_t per_batch, int32_t B);

/* TemporalUnet.forward (temporal_unet.py:199-241).  x, eps: device (B, H, T) fp32.
 * t: device (B,) int64 timesteps, or NULL to use the uniform `step` for every row
 * (what the sampler does: diffusion.py:248, policies.py:146). */
int dad_unet_forward(dad_handle *h, const float *x, const int64_t *t, int32_t step, float *eps,
                     int32_t B, void *stream);

/* Everything in p_sample_with_guidance after the model call (policies.py:84-110 =
 * diffusion.py:159-223 + guidance + noise + inpainting) plus the projection, as ONE fused
 * kernel: x <- step(x, model_out).  noise: device (B,H,T) or NULL for in-kernel Philox
 * (seed, sample_offset = global index of row 0, used as the Philox subsequence).
 * grad: device (B,H,T) guide gradient or NULL; mean += guide_w * exp(logvar) * grad (:97). */
int dad_step(dad_handle *h, float *x, const float *model_out, const float *noise,
             const float *grad, float guide_w, int32_t step, uint32_t flags, uint64_t seed,
             uint64_t sample_offset, int32_t B, void *stream);

/* DynamicsAwarePolicy.apply_projection(x, t) alone (guides/policies.py:409-485), in place on the
 * device tensor x (B,H,T):  x <- x + alpha[step] * (Nmat x + q).  No inpainting. */
int dad_project(dad_handle *h, float *x, int32_t step, int32_t B, void *stream);

/* GaussianDiffusion.p_sample_loop (diffusion.py:225-251) / GuidedPolicy.sample_loop
 * (policies.py:114-149) without a guide function: runs steps i = n_steps-1 .. 0, each step a
 * replay of one captured CUDA graph (U-Net + fused step kernel).
 *   x          device (B,H,T): x_S on entry (unless DAD_FLAG_PHILOX_INIT), x_0 on return;
 *              with DAD_FLAG_CONDITIONS the conditions are applied to x_S first (policies.py:137-138)
 *   noise_seq  device (n_steps, B, H, T): noise_seq[k] is z for step i = n_steps-1-k; NULL -> Philox
 *   trace      device (n_steps, B, H, T) receiving x after every step, or NULL */
int dad_sample(dad_handle *h, float *x, const float *noise_seq, uint64_t seed,
               uint64_t sample_offset, int32_t B, int32_t n_steps, uint32_t flags, float *trace,
               void *stream);

/* Same loop with HOST buffers: copies x_S (unless PHILOX_INIT) to the device, samples, copies
 * x_0 back and synchronises.  This is the call a non-PyTorch host binds, and the `e2e` bench leg. */
int dad_sample_host(dad_handle *h, float *x_host, const float *noise_seq_host, uint64_t seed,
                    uint64_t sample_offset, int32_t B, int32_t n_steps, uint32_t flags);

/* Introspection for tests / benchmarks. */
typedef struct {
  int64_t conv_flops_per_sample;   /* 2 * MACs of all Conv1d/ConvTranspose1d of one forward (zero-padding taps counted) */
  int64_t launches_per_step;       /* kernels of this library launched per diffusion step */
  int64_t workspace_bytes;
  int32_t n_conv_layers;
  int32_t sm_count;
} dad_info;
int dad_get_info(const dad_handle *h, dad_info *out);
/* Count of kernel launches issued by this library since the handle was created (graph replays
 * count their kernel nodes). */
int64_t dad_launch_count(const dad_handle *h);
/* Batches of at most `max_b` samples run the latency kernels (one CTA per 16 output channels of a sample,
 * weights prefetched across the launch boundary): the shape of GuidedPolicy.get_action, which plans ONE
 * trajectory per call (policies.py:193-223).  Larger batches run the throughput kernels.  0 disables the
 * latency kernels; the default is 24, about where the throughput kernels take over on a B200 (or the
 * DAD_SMALL_MAX_B environment variable).  bf16 mode only. */
int dad_set_latency_batch(dad_handle *h, int32_t max_b);
/* Mixed precision across the reverse process.  A bf16 handle evaluates the U-Net of the reverse steps with index >=
 * `min_step` through `companion`, an fp32-precision handle of the same architecture with the same weights loaded (the
 * rest of those steps -- posterior mean, noise, projector, inpainting -- stays in `h`).  With the cosine schedule
 * beta_{S-1} is clipped to 0.9999 (diffusion.py:41): the first reverse step has d(mean)/d(eps) = 99.98, so ANY bf16
 * evaluation of eps (this library 8e-3 relative, stock torch.autocast 1.1e-2) shows up as 1.3-2e-2 on x at that one
 * step; every other step amplifies eps errors by < 1.5.  min_step = S-1 keeps every step within BASELINE.json's 1e-2
 * at the price of one fp32 pass per plan.  Applies to dad_sample / dad_sample_host; the per-step entry points take
 * whatever eps the caller computed.  companion = NULL detaches.  The caller keeps `companion` alive while attached. */
int dad_set_fp32_steps(dad_handle *h, dad_handle *companion, int32_t min_step);
/* Arithmetic of an fp32-precision handle's convolutions: 0 (default) IEEE fp32 on the SIMT pipes -- the 1e-5 parity mode;
 * 1 TF32 operands (round-to-nearest) on the tensor cores with fp32 accumulation and fp32 activations; 2 3xTF32 error
 * compensation (fp32-level accuracy).  1 is what a bf16 model's fp32 companion needs for the ill-conditioned step
 * (eps several times more accurate than bf16) at a fraction of the SIMT cost. */
int dad_set_fp32_math(dad_handle *h, int32_t mode);
/* How the stride-1 convolutions of the U-Net (temporal_unet.py:106-122, 214-237) are grouped into launches, bf16
 * mode: 3 (default) = one persistent conv_chain launch per run of ResidualTemporalBlocks of one level (their convs
 * synchronise through per-sample-tile counters instead of kernel boundaries), 2 = one launch per block, 1 = one
 * launch per convolution with the same kernel, 0 = the per-layer kernels of round 1 (conv_t3).  Results are
 * bit-identical at levels 1-3.  Drops the captured steps (dad_graph_epoch changes). */
int dad_set_fusion(dad_handle *h, int32_t level);

/* ---- guided sampling inside a caller-captured CUDA graph -------------------------------------------
 * ValueGuidedPolicy (policies.py:243-271) differentiates a user value model at x_t every step
 * (policies.py:87-97).  That forward+backward is the caller's (PyTorch's) work; to keep it out of a per-step
 * host loop the caller captures ONE step into its own CUDA graph --
 *     dad_loop_unet;  <its guidance kernels writing `grad`>;  dad_loop_step
 * -- and replays it n_steps times.  dad_loop_begin puts the loop state on the device first (not captured):
 * x (B,H,T) in/out; noise: NULL = in-kernel Philox, else n_steps slots of (B,H,T) consumed in order, or ONE slot
 * the caller refills every step when noise_single != 0; grad: the (B,H,T) buffer the guidance writes (NULL = no
 * guidance); trace: NULL or (n_steps,B,H,T).  B <= max_batch.  The step index lives on the device and is
 * decremented by dad_loop_unet, so the captured step is the same for every index.  dad_loop_unet and
 * dad_loop_step only enqueue kernels (legal under stream capture); flags as dad_step.
 * dad_graph_epoch changes whenever buffers a captured step points to were reallocated (conditions grown,
 * projector replaced, weights reloaded, latency batch changed): re-capture when it differs.
 * dad_loop_replayed adds n replays of the last captured step to dad_launch_count. */
int dad_loop_begin(dad_handle *h, float *x, const float *noise, int32_t noise_single, const float *grad, float guide_w,
                   uint64_t seed, uint64_t sample_offset, int32_t B, int32_t n_steps, uint32_t flags, float *trace,
                   void *stream);
int dad_loop_unet(dad_handle *h, int32_t B, void *stream);
int dad_loop_step(dad_handle *h, int32_t B, uint32_t flags, void *stream);
int64_t dad_graph_epoch(const dad_handle *h);
int dad_loop_replayed(dad_handle *h, int32_t n);

/* ---- projector build on the device ---------------------------------------------------------------------
 * P = F pinv(F), the orthogonal projector onto range(F) (ProjectionMatrixBuilder.get_projection_matrix,
 * dynamics/projection.py:85-120: numpy SVD pinv in fp64, cast to fp32).  F: rows x cols fp64 row-major, HOST
 * memory; P: rows x rows fp32, HOST memory.  fp64 Gram matrix + blocked Cholesky + triangular solve on `device`;
 * F must have full column rank (the reference's F always has: its rows contain the identity), otherwise
 * DAD_ERR_INVALID and dad_last_error(NULL) names the failing pivot.  No handle needed. */
int dad_build_projection_matrix(int32_t device, const double *F, int32_t rows, int32_t cols, float *P);

/* ---- data-driven dynamics on the device -------------------------------------------------------------------
 * fit_linear_dynamics (dynamics/data_driven.py:107-121): Theta = lstsq([X U], X+), A = Theta[:n]^T, B = Theta[n:]^T.
 * X (N x n), U (N x m), Xn (N x n): fp64 row-major HOST arrays of N transitions; A (n x n), B (n x m): fp64
 * row-major HOST outputs.  Normal equations + Cholesky in fp64 on `device`; DAD_ERR_INVALID when [X U] is rank
 * deficient to working precision.  No handle needed. */
int dad_fit_linear_dynamics(int32_t device, const double *X, const double *U, const double *Xn, int64_t N, int32_t n,
                            int32_t m, double *A, double *B);
/* ProjectionLoss.compute (losses/__init__.py:161-186), the dynamics-violation metric: mean((tau - tau P)^2) with
 * tau = [unnormalised states, last one duplicated | unnormalised actions].  x: (B, H, n + m) normalised trajectories,
 * DEVICE fp32; P: ((H+1) n + H m)^2 DEVICE fp32 row-major; the four statistics: HOST fp32 arrays of n / n / m / m
 * entries; *out: HOST double.  One fused kernel on `stream`; synchronises it. */
int dad_dynamics_residual(int32_t device, const float *x, int32_t B, int32_t H, int32_t n, int32_t m, const float *P,
                          const float *obs_mean, const float *obs_std, const float *act_mean, const float *act_std,
                          double *out, void *stream);

/* ---- measurement hooks (bench.py; no reference counterpart) ------------------------------------ */

/* dad_sample with Philox noise that also returns the device time of every diffusion step (CUDA events
 * recorded on `stream` between the graph replays).  step_ms: HOST array of n_steps floats, step_ms[k] is
 * the k-th executed step (i = n_steps-1-k).  Synchronises `stream` before returning.  B <= max_batch. */
int dad_sample_profile(dad_handle *h, float *x, uint64_t seed, uint64_t sample_offset, int32_t B,
                       int32_t n_steps, uint32_t flags, float *step_ms, void *stream);

/* One convolution layer of the U-Net plan (a Conv1d / ConvTranspose1d phase of temporal_unet.py). */
typedef struct {
  char name[96];            /* state_dict stem of the weight, e.g. "mid_block1.blocks.0.block.0" */
  int32_t L_out, C_in, C_out, taps;
  int32_t tile_n, group_width;   /* bf16 path: output channels per work item and GroupNorm width (0 = no GN) */
  int64_t flops_per_sample;      /* 2 * L_out * taps * C_in(real) * C_out */
  char kernel[64];               /* kernel instantiation that executes the layer, e.g. "conv_t3_kernel<64,1,pair,256>" */
} dad_layer_desc;
int dad_layer_count(const dad_handle *h);
int dad_layer_info(const dad_handle *h, int32_t index, dad_layer_desc *out);
/* Times `iters` back-to-back launches of layer `index` at batch B on `stream` with CUDA events (after one
 * untimed launch); *ms_per_launch receives the average.  Operates on the handle's own workspaces (their
 * contents are whatever the last forward left there).  Synchronises. */
int dad_time_layer(dad_handle *h, int32_t index, int32_t B, int32_t iters, float *ms_per_launch, void *stream);
/* Device-side diagnostics of the conv chains: out4[0] = error code left by a dependency wait that timed out (0 = none);
 * out4[1..3] = waits that had to spin / nanoseconds spent spinning / all dependency waits (counted by -DDAD_TUNING
 * builds only, else 0).  Synchronises the device; `reset` clears the counters. */
int dad_debug_counters(dad_handle *h, uint32_t *out4, int32_t reset);
/* Launch units of one U-Net pass at the current fusion level: a conv chain (layers [first_layer, first_layer +
 * n_layers) in ONE launch) or a single layer. */
typedef struct {
  int32_t first_layer, n_layers;
  int32_t is_chain;
  int32_t L_out, C_out;
  int64_t flops_per_sample;      /* sum over the unit's layers */
  char kernel[64];
} dad_unit_desc;
int dad_unit_count(const dad_handle *h);
int dad_unit_info(const dad_handle *h, int32_t index, dad_unit_desc *out);
/* dad_time_layer for a whole launch unit (the chain's counters are reset before the loop and every launch waits
 * for its own epoch, so the in-chain dependencies are exercised exactly as in a U-Net pass). */
int dad_time_unit(dad_handle *h, int32_t index, int32_t B, int32_t iters, float *ms_per_launch, void *stream);
/* Same for the fused step kernel(s) (K7 [+K8]) at step index `step` on scratch trajectories; Philox noise, or a
 * noise buffer when flags has bit 0x100 set (the parity-mode data path: 16 instead of 12 bytes per element). */
int dad_time_step_kernel(dad_handle *h, int32_t B, int32_t step, uint32_t flags, int32_t iters,
                         float *ms_per_launch, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DAD_B200_H */
