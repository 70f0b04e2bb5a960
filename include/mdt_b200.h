/*
 * mdt_b200.h -- C ABI of the B200-native reverse-diffusion sampler.
 *
 * The reference (lamm-mit/MoleculeDiffusionTransformer) is pure Python and has no FFI of its
 * own; its boundary for this path is the Python object protocol
 *     QMDiffusion.sample(sequences, device, cond_scale, timesteps, clamp)   generative.py:834-870
 *     QMDiffusionForward.sample(...)                                        generative.py:146-180
 * plus the state_dict key layout.  This header is what a ctypes binding of that boundary
 * calls.  Plain pointers and sizes only; no torch types.
 *
 * Conventions: every function returns 0 on success or a negative mdt_status; the message is
 * available from mdt_last_error() (thread-local).  Nothing throws across the ABI.
 * Device pointers are marked _dev; everything else is host memory.
 * A plan is bound to one device and is not thread-safe; use one plan per (device, stream).
 */
#ifndef MDT_B200_H
#define MDT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDT_ABI_VERSION 1
#define MDT_MAX_LEVELS 4

typedef enum {
  MDT_OK = 0,
  MDT_ERR_INVALID = -1,   /* bad argument / unsupported configuration (reference: assert) */
  MDT_ERR_MISSING = -2,   /* a state_dict tensor is missing or has the wrong size */
  MDT_ERR_CUDA = -3,      /* CUDA runtime error */
  MDT_ERR_NO_DEVICE = -4, /* no sm_100 device: there is no CPU fallback */
  MDT_ERR_OOM = -5
} mdt_status;

/* Arithmetic mode of the contraction kernels (GEMM / implicit-GEMM conv). */
typedef enum {
  MDT_PREC_FP32 = 0, /* CUDA-core fp32 FFMA: bit-for-bit-class parity mode            */
  MDT_PREC_TF32 = 1, /* tcgen05 kind::tf32, fp32 accumulate in TMEM                   */
  MDT_PREC_BF16 = 2, /* tcgen05 kind::f16 (bf16 operands), fp32 accumulate in TMEM    */
  MDT_PREC_F16 = 3   /* tcgen05 kind::f16 (fp16 operands: tf32's 11-bit significand in half the bytes, saturating at +-65504),
                        fp32 accumulate in TMEM, fp32 residual stream / norms / sampler state */
} mdt_precision;

/* Static model description.  Mirrors the kwargs QMDiffusion / QMDiffusionForward pass to
 * XUNet1d (generative.py:761-776, 69-83; modules.py:935-957) plus the wrapper's own fields. */
typedef struct {
  int32_t abi_version;          /* MDT_ABI_VERSION */
  int32_t in_channels;          /* pred_dim */
  int32_t out_channels;         /* = in_channels */
  int32_t length;               /* max_length (positions of the generated tensor) */
  int32_t channels;             /* base width */
  int32_t patch_size;
  int32_t num_levels;           /* len(multipliers) - 1 */
  int32_t multipliers[MDT_MAX_LEVELS + 1];
  int32_t factors[MDT_MAX_LEVELS];
  int32_t num_blocks[MDT_MAX_LEVELS];
  int32_t attentions[MDT_MAX_LEVELS + 1]; /* [num_levels] is the bottleneck depth (attentions[-1]) */
  int32_t pre_transformer;      /* self-attention-only layers ahead of the resnets */
  int32_t heads, head_features, ff_multiplier;
  int32_t resnet_groups;
  int32_t kernel_multiplier_downsample;
  int32_t use_skip_scale;
  int32_t mapping_features;     /* channels * context_features_multiplier */
  int32_t ctx_features;         /* context_embedding_features (F) */
  int32_t ctx_max_length;       /* context_embedding_max_length */
  int32_t text_embed_dim;       /* fc1 out features */
  int32_t embed_dim_position;   /* Fourier PE channels */
  int32_t pos_emb_fourier;      /* bool */
  int32_t pos_emb_fourier_add;  /* bool */
  float sigma_data;             /* 0.1 */
  int32_t precision;            /* mdt_precision */
  int32_t max_batch;            /* samples processed per internal chunk (workspace is sized for it) */
  int32_t max_timesteps;        /* FiLM tables are sized for 2 * (max_timesteps - 1) denoiser calls; 0 => 256 */
} mdt_config;

/* One named fp32 host tensor of the reference state_dict ("unet.to_in.block.block1.project.weight"). */
typedef struct {
  const char* name;
  const float* data;
  int64_t numel;
} mdt_tensor;

/* Scalars of one ADPM2 iteration (two denoiser calls), computed on the host exactly as the
 * reference does (diffusion.py:495-515, 789-796); see moleculediffusiontransformer_b200/diffusion.py. */
typedef struct {
  float sigma, c_in_a, c_noise_a, c_skip_a, c_out_a;
  float sigma_mid, c_in_b, c_noise_b, c_skip_b, c_out_b;
  float dt_mid;   /* sigma_mid  - sigma */
  float dt_down;  /* sigma_down - sigma */
  float sigma_up;
} mdt_iter_scalars;

typedef struct mdt_plan mdt_plan;

const char* mdt_last_error(void);
int mdt_abi_version(void);

/* Number of visible CUDA devices with compute capability 10.x (0 => nothing can run). */
int mdt_device_count(void);

/* Fill `out[n_iters]` from a sigma schedule `sigmas[n_iters + 1]` (KarrasSchedule output, host fp32),
 * restating ADPM2Sampler.get_sigmas (diffusion.py:495-500) and get_scale_weights (diffusion.py:789-796)
 * in C for non-Python hosts.  rho is the sampler's rho (1.0 for the QM wrappers). */
int mdt_adpm2_scalars(const float* sigmas, int n_iters, double rho, double sigma_data, mdt_iter_scalars* out);

/* Same for AEulerSampler.get_sigmas / step (diffusion.py:467-475): rows with sigma_mid = sigma and dt_mid = 0, which
 * mdt_plan_sample executes as one denoiser call per step. */
int mdt_aeuler_scalars(const float* sigmas, int n_iters, double sigma_data, mdt_iter_scalars* out);

/* KarrasSchedule.forward (diffusion.py:333-342): out[num_steps + 1], last entry 0. */
int mdt_karras_sigmas(int num_steps, double sigma_min, double sigma_max, double rho, float* out);

/* Build a plan on `device`: looks up every parameter by its reference state_dict key, repacks it
 * into kernel-native layouts in HBM and allocates the activation workspace for cfg->max_batch. */
int mdt_plan_create(const mdt_config* cfg, const mdt_tensor* tensors, int64_t n_tensors, int device,
                    mdt_plan** out);
void mdt_plan_destroy(mdt_plan* plan);

/* Bytes of HBM held by the plan (weights + workspace). */
int64_t mdt_plan_device_bytes(const mdt_plan* plan);
/* Kernels launched by this plan since creation (claims for bench.py's gpu_launches). */
int64_t mdt_plan_launch_count(const mdt_plan* plan);

/* The hot path: replaces QMDiffusion.sample / QMDiffusionForward.sample end to end.
 *   cond_dev        [B, n_ctx] fp32      conditioning (`sequences`)
 *   noise0_dev      [B, P, L] fp32       initial noise, or NULL => Philox(seed, sample_offset + b)
 *   step_noise_dev  [n_iters, B, P, L]   ancestral noise per iteration, or NULL => Philox
 *   iters           [n_iters]            host scalars (n_iters = timesteps - 1)
 *   out_dev         [B, P, L] fp32       result (reference layout), may be NULL if tokens_dev given
 *   tokens_dev      [B, L] uint8         argmax over P (generative.py:1212-1213), may be NULL
 *   stream          cudaStream_t as void* (NULL = legacy default stream)
 * B may exceed cfg->max_batch: the plan walks it in chunks.  Asynchronous on `stream`.
 * Rows whose midpoint is the start (sigma_mid == sigma, dt_mid == 0, the `_b` coefficients those of sigma) describe first-order
 * ancestral Euler steps (AEulerSampler, diffusion.py:456-483): when every row is of that kind the first denoiser call of each
 * iteration is skipped (one call per step). */
int mdt_plan_sample(mdt_plan* plan, const float* cond_dev, int32_t n_ctx, const float* noise0_dev,
                    const float* step_noise_dev, const mdt_iter_scalars* iters, int32_t n_iters,
                    uint64_t seed, uint64_t sample_offset, int64_t B, float cond_scale, int32_t clamp,
                    float* out_dev, uint8_t* tokens_dev, void* stream);

/* Context mode: pre_encoded != 0 makes `cond_dev` of the sampling entry points the already encoded conditioning embedding
 * [B, n_ctx, ctx_features] fp32 -- what QMDiffusion.sample hands to XDiffusion_x.sample as `embedding=` (generative.py:838-860,
 * diffusion.py:724-741) -- instead of the raw `sequences` [B, n_ctx]; the device-side encoder is skipped.  Default 0. */
int mdt_plan_set_context_mode(mdt_plan* plan, int pre_encoded);

/* Sampler mode of mdt_plan_sample.  0 (default): the rows describe ADPM2Sampler / AEulerSampler iterations (see above).
 * 1: KarrasSampler (diffusion.py:399-453): rows carry sigma = sigma_hat, sigma_mid = sigma_next, dt_mid = sigma_next - sigma_hat,
 * dt_down = 0.5 (sigma - sigma_hat) and sigma_up = the noise scale sqrt(sigma_hat^2 - sigma^2) s_noise of the NEXT step; the second
 * update is x_hat + dt_down (d + d'), the step noise of iteration i is drawn ahead of its first denoiser call (slot i of
 * step_noise_dev / Philox stream i + 1) `init_noise_scale` is the scale of step 0 and `init_sigma` = sigmas[0] (x = sigmas[0] * noise, diffusion.py:438; row 0 carries sigma_hat). */
int mdt_plan_set_sampler_mode(mdt_plan* plan, int mode, float init_noise_scale, float init_sigma);

/* Inpainting: replaces QMDiffusion.inpaint -> XDiffusion_x.inpaint -> DiffusionInpainter.forward ->
 * ADPM2Sampler.inpaint (generative.py:871-914, diffusion.py:744-767, 612-625, 526-549).
 *   source_dev [B, P, L] fp32   the draft to keep where mask != 0;   mask_dev [B, P, L] uint8
 *   noise_dev  [1 + n_iters * 2 * num_resamples, B, P, L] fp32 in the reference's draw order (initial state; per iteration: source
 *              noise, then per resample: step noise [, re-noise]) or NULL => Philox keyed by (seed, sample, draw index)
 *   sigmas     [n_iters + 1] host fp32 (the schedule; used for the re-noise scale sqrt(sigma_i^2 - sigma_{i+1}^2))
 * No final clamp (the reference's inpainter has none).  Asynchronous on `stream`. */
int mdt_plan_inpaint(mdt_plan* plan, const float* cond_dev, int32_t n_ctx, const float* source_dev, const uint8_t* mask_dev,
                     const float* noise_dev, const mdt_iter_scalars* iters, const float* sigmas, int32_t n_iters,
                     int32_t num_resamples, uint64_t seed, uint64_t sample_offset, int64_t B, float cond_scale, float* out_dev,
                     void* stream);

/* One raw denoiser-network evaluation (UNetCFG1d.forward, modules.py:1228-1255) for kernel-level
 * parity tests: x_dev [B,P,L], scalar `time` (= c_noise) shared by the batch, out_dev [B,P,L]. */
int mdt_plan_unet_forward(mdt_plan* plan, const float* x_dev, float time, const float* cond_dev,
                          int32_t n_ctx, int64_t B, float cond_scale, float* out_dev, void* stream);

/* Debug taps: after mdt_plan_unet_forward with taps enabled, copy the named stage output
 * (token-major [B_eff * L_stage, C_stage] fp32) to host.  Returns element count or <0. */
int mdt_plan_enable_taps(mdt_plan* plan, int enable);
int64_t mdt_plan_read_tap(mdt_plan* plan, const char* name, float* host_dst, int64_t capacity);

/* ---- single-kernel entry points (tests, micro-benchmarks, ncu captures) ---------------- */

/* C[M,N] = act(A[M,K] @ W[N,K]^T + bias) (+ res); all row-major fp32 device pointers.
 * precision selects the CUDA-core or the tcgen05 kernel.  act: 0 none, 1 exact-erf GELU. */
int mdt_op_linear(const float* a_dev, const float* w_dev, const float* bias_dev, const float* res_dev,
                  float* c_dev, int64_t M, int32_t N, int32_t K, int32_t act, int32_t precision, void* stream);

/* Fused sampler update used after denoiser call A / call B; exposed for HBM-roofline measurement. */
int mdt_op_step_update(int which, const float* net_dev, float* x_dev, float* xmid_dev, float* xin_dev,
                       const float* noise_dev, const mdt_iter_scalars* it, float cond_scale, int64_t B,
                       int32_t P, int32_t L, int cfg, void* stream);

/* Token ids -> text bytes on the device: replaces reverse_tokenize (generative.py:1069-1078; Keras sequences_to_texts on the argmax
 * tokens of generative.py:1212-1213, then the spaces stripped).  lut_dev[256] maps a token id to its character (0: id has no
 * vocabulary entry -- the padding id 0 always -- and is dropped).  out_dev[B][L] receives the compacted text zero padded to L,
 * lengths_dev[B] (may be NULL) the number of characters. */
int mdt_op_decode_tokens(const uint8_t* tokens_dev, const uint8_t* lut_dev, uint8_t* out_dev, int32_t* lengths_dev, int64_t B,
                         int32_t L, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MDT_B200_H */
