// kernels.cuh -- parameter blocks and launcher declarations shared by the CUDA translation units.
//
// Activation layout everywhere inside the library: TOKEN-MAJOR fp32, i.e. a tensor the
// reference holds as (b, C, L) lives in HBM as [b * L + l][c] (row = one position of one
// sample, channels contiguous).  All contractions are then row-major GEMMs with M = B_eff * L.
#pragma once
#include <cuda_runtime.h>
#include "launch.cuh"
#include <stdint.h>

namespace mdt {

// ---------------------------------------------------------------------------------------------
// A-operand loader: describes how element (m, k) of the GEMM's left operand is produced from
// activations in HBM.  It covers plain linear layers, k-tap strided Conv1d as implicit GEMM
// (zero padding at SAMPLE boundaries), channel concatenation of two sources (skip connections),
// LayerNorm / GroupNorm normalisation on load, per-channel affine (GroupNorm gamma/beta with the
// FiLM scale/shift folded in, per denoiser call) and SiLU.
//   k = tap * C + c ;  m = b * L_out + lo ;  li = lo * stride + tap - pad
//   v = src(b * L_in + li, c) [* scale1 for the second segment]
//   v = (v - mean) * rstd            if stats   (mode 1: per source row; mode 2: per (sample, group))
//   v = v * aff[c] + aff[C + c]      if aff
//   v = silu(v)                      if silu
//   v = 0 where li is outside [0, L_in)   (padding applies to the transformed tensor, modules.py:105-122)
// ---------------------------------------------------------------------------------------------
struct ALoad {
  const float* src0;
  const float* src1;
  int c0, c1, C;
  float scale1;
  int L_in, L_out, taps, stride, pad;
  const float* stats;
  int stats_mode;  // 0 none, 1 row, 2 (sample, group)
  int groups, cpg;
  const float* aff;
  int aff_call_stride;  // floats per call index; 0 = static table
  const int* call_idx;  // device scalar: current denoiser call (may be null)
  int silu;
};

struct GemmParams {
  ALoad a;
  const float* W;     // [N][K] fp32, K = taps * C contiguous
  const void* Wtc;    // same matrix pre-converted for the tensor-core kernel (tf32 bits or bf16), or null
  const float* bias;  // [N] or null
  int M, N, K;
  int act;            // 0 none, 1 exact GELU
  const float* res;   // residual [M][ldres] or null (may alias C)
  int ldres;
  float* C;
  int ldc;
};

// Element type of q/k/v/o is selected by the launcher's `kind` (0/1: fp32, 2: bf16); strides are in ELEMENTS.
struct AttnParams {
  const void* q; int ldq;             // [B*nq][ldq], head h at columns h*d..
  const void* k; const void* v;       // per-sample blocks: row (b*nk + j), leading dim ldkv
  int ldkv; long long kv_sample_stride;  // elements between consecutive samples' K/V blocks
  const void* k_null; const void* v_null;  // shared K/V for samples >= n_cond (classifier-free null branch)
  int n_cond;
  void* o; int ldo;
  int B, nq, nk, heads, d;
  float scale;
};

struct NormStatsParams {
  const float* src0; const float* src1;
  int c0, c1; float scale1;
  int L;        // positions per sample (GroupNorm) -- unused for row mode
  int groups;   // GroupNorm groups
  float eps;
  float* stats; // out: [rows][2] or [B][groups][2]   (mean, rstd)
  int rows;     // row mode: number of rows; group mode: number of samples
};

// Per-iteration scalars as laid out in include/mdt_b200.h (mdt_iter_scalars).
struct IterScalars {
  float sigma, c_in_a, c_noise_a, c_skip_a, c_out_a;
  float sigma_mid, c_in_b, c_noise_b, c_skip_b, c_out_b;
  float dt_mid, dt_down, sigma_up;
};

// Per-run values the captured step kernels read from device memory, so one CUDA graph serves every seed, shard offset,
// guidance scale and injected-noise buffer (nothing run-specific is baked into the captured kernel arguments).
struct RunParams {
  unsigned long long seed, sample_offset;
  const float* noise;            // injected ancestral noise base for the running chunk, or null => Philox
  long long noise_iter_stride;   // floats between iterations in the injected noise tensor
  float cond_scale;
  int pad;
};

struct StepParams {
  const IterScalars* iters;  // device table [n_iters]
  const int* call_idx;       // device scalar; iteration = call_idx / 2
  const float* net;          // network output, token-major [B_eff*L][P]; rows [0,B*L) cond, [B*L, 2*B*L) null
  float* x;                  // sampler state, token-major [B*L][P]
  float* xmid;               // midpoint state
  float* xin;                // next network input, token-major [B_eff*L][P]  (both halves written)
  const float* noise;        // injected ancestral noise (B,P,L) for this iteration, or null => Philox
  long long noise_iter_stride;  // floats between iterations in the injected noise tensor
  unsigned long long seed, sample_offset;
  const RunParams* run;      // optional device block overriding seed / sample_offset / noise / noise_iter_stride / cond_scale (graph replays)
  int has_noise;             // with `run`: whether run->noise is set (sizes the transposition tile at launch time)
  float cond_scale;
  int cfg;                   // 1 => two branches
  int B, P, L;
  int n_iters;
  int noise_stream;          // Philox stream id override for the ancestral noise (< 0: iteration + 1)
  int karras;                // KarrasSampler rows (diffusion.py:399-453): update A keeps its slope d in `daux`, update B is
                             // x + dt_down * (d + d'), and the noise added afterwards is that of the NEXT step (slot / stream + 1)
  float* daux;               // [B*L][P] slope of call A (karras only)
  float* out;                // final (B,P,L) result written when the last iteration completes (may be null)
  unsigned char* tokens;     // final argmax tokens [B][L] (may be null)
  int clamp;
};

// ---- launchers (kernels.cu) -------------------------------------------------------------------
// per-device one-time function attributes (call after cudaSetDevice, outside stream capture)
cudaError_t init_kernels();
cudaError_t init_gemm_tc();
cudaError_t launch_gemm_fp32(const GemmParams& p, cudaStream_t s);
// kind: 0 fp32 in/out, 1 fp32 in / tf32-rounded fp32 out, 2 bf16 in/out
cudaError_t launch_attention(const AttnParams& p, int kind, cudaStream_t s);
// TMA-bulk-copy streaming variant (attention_bulk.cu); launch_attention dispatches to it when the shape fits
bool attention_bulk_supported(const AttnParams& p, int kind);
cudaError_t init_attention_bulk();
cudaError_t launch_attention_bulk(const AttnParams& p, int kind, cudaStream_t s);
cudaError_t launch_groupnorm_stats(const NormStatsParams& p, cudaStream_t s);
cudaError_t launch_rownorm_stats(const NormStatsParams& p, cudaStream_t s);
// out[b, o, co] = bias[co] + sum_j Y[b, i_j, k_j * Cout + co] (+ add[b, o, co]);  ConvTranspose1d(k=2f, s=f, p=f/2)
cudaError_t launch_upsample_gather(const float* Y, const float* bias, const float* add, float* out, int B, int Lin,
                                   int Cout, int f, cudaStream_t s);
// Patcher / Unpatcher index maps (modules.py:230, 255): to_patched: [B, L*p, C] -> [B, L, C*p]
cudaError_t launch_patch_permute(const float* in, float* out, int B, int L, int C, int p, int to_patched,
                                 cudaStream_t s);
// conditioning encoder (generative.py:838-850): emb[b, i, :] = [gelu(w * s + beta) | PE(i)] (or sum)
cudaError_t launch_encode_cond(const float* seq, const float* w, const float* bias, const float* inv_freq,
                               float* emb, int B, int n, int text_dim, int pos_dim, int add, cudaStream_t s);
// time features (modules.py:554-559): out[r, :] = [t, sin(2 pi w t), cos(2 pi w t)], t = c_noise of call r
cudaError_t launch_time_features(const float* t, const float* w, float* out, int rows, int half, cudaStream_t s);
// FiLM fold: aff[r][c] = gamma[c] * (1 + ss[r][c]);  aff[r][C + c] = beta[c] * (1 + ss[r][c]) + ss[r][C + c]
cudaError_t launch_film_fold(const float* ss, const float* gamma, const float* beta, float* aff, int rows, int C,
                             cudaStream_t s);
cudaError_t launch_step_init(const float* noise0, float* x, float* xin, const IterScalars* iters,
                             unsigned long long seed, unsigned long long sample_offset, int B, int P, int L,
                             int cfg, float sigma0, cudaStream_t s);   // sigma0 >= 0 overrides iters[0].sigma
cudaError_t launch_step_update(int which, const StepParams& p, cudaStream_t s);
// KarrasSampler step 0: x += scale * eps_0 (injected (B,P,L) tensor or Philox stream 1), xin = c_in * x (both branches)
cudaError_t launch_karras_prenoise(float* x, float* xin, const float* noise, float scale, float c_in, unsigned long long seed,
                                   unsigned long long sample_offset, int B, int P, int L, int cfg, cudaStream_t s);
cudaError_t launch_finalize(const float* x, float* out, unsigned char* tokens, int B, int P, int L, int clamp,
                            cudaStream_t s);
cudaError_t launch_inpaint(int mode, float* x, float* xin, const float* source, const unsigned char* mask, const float* noise,
                           float sigma, float c_in, unsigned long long seed, unsigned long long sample_offset, int stream, int B,
                           int P, int L, int cfg, float* out, cudaStream_t s);
cudaError_t launch_set_run_params(RunParams* dst, const RunParams& v, cudaStream_t s);
// kperm = 1: K fragments in the permuted k order of gemm_attn_frag.cu (b0 = K[key][8 ks + 2q], b1 = K[key][8 ks + 2q + 1]);
// kperm = 2: f16 m16n8k16 fragments (8 K + 8 V fragments of f16x2 pairs per block, half the bytes)
cudaError_t launch_kv_fragment_pack(const float* kv, void* out, long long B, int nk, int heads, int d, int kperm, cudaStream_t s);
cudaError_t launch_decode_tokens(const uint8_t* tokens, const uint8_t* lut, uint8_t* out, int* lengths, long long B, int L, cudaStream_t s);
cudaError_t launch_set_int(int* dst, int v, cudaStream_t s);
cudaError_t launch_add_int(int* dst, int v, cudaStream_t s);
// (B,P,L) <-> token-major [B*L][P], with optional duplication into a second half (classifier-free null rows)
cudaError_t launch_to_token_major(const float* in, float* out, int B, int P, int L, float mul, int dup,
                                  cudaStream_t s);
cudaError_t launch_cfg_mix_to_bpl(const float* net, float* out, int B, int P, int L, float cond_scale, int cfg,
                                  cudaStream_t s);

// ---- operand preparation (prep.cu): normalise + activate once, write the MMA operand dtype -----------
// kind: 1 = tf32 (fp32 storage, round-to-nearest tf32), 2 = bf16
struct GnApplyParams {
  const float* src0; const float* src1; int c0, c1; float scale1;   // concat of two fp32 token-major sources
  int L, groups; float eps;
  const float* aff; int aff_call_stride; const int* call_idx;       // per-channel (A, B) table or null
  int silu;
  void* out;   // normalised operand [B*L][C]
  void* raw;   // optional un-normalised operand copy (1x1 skip projection input) or null
  int B;
  int rev;     // blocks walk the samples from the end (serpentine order)
};
cudaError_t launch_gn_apply(const GnApplyParams& p, int kind, cudaStream_t s);
bool gn_apply_supported(int L, int C, int groups);
struct LnApplyParams { const float* src; int C; float eps; void* out; long long rows; int rev; };
cudaError_t launch_ln_apply(const LnApplyParams& p, int kind, cudaStream_t s);

// ---- TMA-fed tcgen05 GEMM (gemm_tma.cu): both operands arrive by cp.async.bulk.tensor ------------------
struct TmaGemmParams {
  int M, N, BN;
  int taps, pad, kchunks, C;  // K loop = taps x kchunks chunks of 128 bytes; weight column = tap * C + kc * KCH
  int L, Lb, Sb;              // positions per sample; TMA box = (KCH, Lb, Sb), Lb * Sb == 128 rows
  const float* bias; int act;
  const float* res; int ldres;
  float* C32; int ldc;        // fp32 output (residual stream) or null
  void* Cop; int ldcop;       // operand-dtype copy of the same values (next GEMM / attention input) or null
  // optional fused GroupNorm (+ per-call affine with FiLM folded in + SiLU) of the biased output, written to Cop only:
  // a 32-row x 32-column epilogue block holds whole (sample, group) sets when gn_L | 32 and gn_cpg | 32 (gn_L = 0: off)
  int gn_L, gn_cpg; float gn_eps;
  const float* gn_aff; int gn_aff_stride; const int* gn_call;
  int rev;                    // walk the row tiles from the end (serpentine order: start with what the producer wrote last)
  int cop_ln; float ln_eps;   // Cop = LayerNorm(C32 row) without affine (N == BN: the row block holds whole rows) instead of the raw copy
};
// 128-byte CUtensorMap blobs (64-byte aligned) built on the host
int make_tmap_act(void* map128, const void* base, int kind, int C, int L, long long samples);
int make_tmap_weight(void* map128, const void* base, int kind, long long K, int N, int BN);
int tma_pick_bn(int N);
bool gemm_tma_shape_ok(int kind, int C, int L, int N);
cudaError_t init_gemm_tma();
cudaError_t launch_gemm_tma(const void* tmA, const void* tmB, const TmaGemmParams& p, int kind, cudaStream_t s);

// ---- fused per-head projection + attention (gemm_attn.cu) ------------------------------------------------
struct GemmAttnParams {
  int M;                  // valid rows = B_eff * L
  int heads, d;
  int kchunks, C;         // K loop over the (LayerNorm-ed) activation channels
  int L, Sb;              // positions per sample; samples per 128-row tile
  int cross;              // 0: self-attention (per-head [q|k|v], BN = 3d); 1: cross-attention (q only, BN = d)
  const float* bias;      // folded q bias [heads * d] (k bias cancels in the softmax; v bias is folded into the out-projection)
  float scale;
  void* att; int ldo;     // head outputs [M][heads * d] in the operand dtype
  const void* kc; const void* kn;  // cross: conditioning K|V cache [B][nk][2 * heads * d] and the shared null-branch block
  int ldkv; long long kv_sample_stride; int n_cond; int nk;
  int kv_fp32;            // the cache pointers hold fp32 (always true in tf32 mode; bf16 mode may pass the fp32 cache)
  int pack_self;          // self, L <= 8: 16 / L samples per m16 tile with a block-diagonal mask
  const void* kvf_c; const void* kvf_n;   // cross, L <= 8: fragment-ordered tf32 K/V cache (kv_fragment_pack_kernel), 8 KB per (sample, head); null = off
  int rev;                // walk the row tiles from the end (serpentine order)
};
bool gemm_attn_supported(int kind, int C, int L, int heads, int d, int cross, int nk_max);
cudaError_t init_gemm_attn();
cudaError_t launch_gemm_attn(const void* tmA, const void* tmB, const GemmAttnParams& p, int kind, cudaStream_t s);
// same contract, attention core on tcgen05 too (gemm_attn_umma.cu): block-diagonal S = Q K^T and O = P V per 128-row tile
bool gemm_attn_umma_supported(int kind, int C, int L, int heads, int d, int cross, int nk_max);
cudaError_t init_gemm_attn_umma();
cudaError_t launch_gemm_attn_umma(const void* tmA, const void* tmB, const GemmAttnParams& p, int kind, cudaStream_t s);

// ---- one whole attention layer (gemm_attn_layer.cu): projection -> attention -> out-projection + bias + residual -----------
struct AttnLayerParams {
  GemmAttnParams a;        // projection + attention core as in gemm_attn.cu (att / ldo unused: head outputs go to the scratch slots)
  int Cout;                // out-projection width (the model width C)
  const float* bias_o;     // [Cout] (to_out bias, v bias folded in for self-attention)
  const float* res; int ldres;   // residual: the fp32 token stream (may alias C32)
  float* C32; int ldc;     // fp32 output
  void* Cop; int ldcop;    // operand-dtype copy of the output (the next GEMM's A operand) or null
  void* scratch;           // [SMs][attn_layer_slots(heads)][128][d] operand dtype: CTA-private head-output slots (L2 resident)
  int f16;                 // gemm_attn_frag.cu: attention core on f16 m16n8k16 MMAs (cross: K / V cache packed with kperm = 2)
  int l2_hint;             // scratch stores carry an L2 evict_last policy (MDT_L2_HINT=1; measured in profiles/README.md)
  int cop_ln; float ln_eps;  // gemm_attn_frag.cu, fused: Cop = LayerNorm(C32 row) without affine (the next attention stage's operand)
  int fused;               // 1: out-projection inside the kernel; 0 (gemm_attn_frag.cu only): head outputs go to a.att, one work item per (row block, head)
  int nslot;               // filled by the launcher from here on
  int nst, stage_bytes, nacc; unsigned tmem_cols;
  int ares_bytes;          // gemm_attn_frag.cu: bytes of the resident activation tile in front of the stage ring (0: streamed per head)
};
bool attn_layer_supported(int kind, int C, int L, int heads, int d, int cross, int Cout);
size_t attn_layer_scratch_bytes(int kind, int heads, int d);
int attn_layer_sms();
int attn_layer_slots(int heads);   // scratch slots per CTA
cudaError_t init_attn_layer();
// tmA: activation map, tmB: per-head projection weights (as launch_gemm_attn); tmS: scratch viewed as [SMs][slots * 128][d];
// tmW: out-projection weight [Cout][heads * d] with a (KCH x Cout) box
cudaError_t launch_attn_layer(const void* tmA, const void* tmB, const void* tmS, const void* tmW, const AttnLayerParams& p, int kind,
                              cudaStream_t s);

// same contract with the q / k fragments read straight from TMEM and sixteen attention warps (gemm_attn_frag.cu)
bool attn_frag_supported(int kind, int C, int L, int heads, int d, int cross, int Cout);
cudaError_t init_attn_frag();
cudaError_t launch_attn_frag(const void* tmA, const void* tmB, const void* tmS, const void* tmW, const AttnLayerParams& p, int kind,
                             cudaStream_t s);

// ---- FeedForward chain (gemm_chain.cu): Linear -> GELU -> Linear + residual (+ LayerNorm of the result) in one kernel ---------
struct FFChainParams {
  int M, C, mid;          // rows, model width, hidden width (mid = C * multiplier)
  int L, Sb;              // positions per sample; samples per 128-row block (TMA box of the activation map)
  const float* b0;        // [mid]
  const float* b2;        // [C]
  const float* res; int ldres;   // residual: the fp32 token stream (may alias C32)
  float* C32; int ldc;    // fp32 output
  void* Cop; int ldcop;   // operand-dtype copy of the output, or LayerNorm(output) when cop_ln (no affine), or null
  int cop_ln; float ln_eps;
  void* scratch;          // [SMs][2][128][mid] operand dtype: CTA-private hidden blocks (L2 resident)
  int rev;                // walk the row blocks from the end (serpentine order)
  int nst, stage_bytes; unsigned tmem_cols;   // filled by the launcher
};
bool ff_chain_supported(int kind, int C, int mid, int L);
size_t ff_chain_scratch_bytes(int kind, int mid);
int ff_chain_sms();
cudaError_t init_ff_chain();
// tmA: activation map; tmB0: W0 [mid][C] with a (KCH x 128) box; tmS: scratch viewed as [SMs][2 * 128][mid]; tmW: W2 [C][mid], (KCH x C) box
cudaError_t launch_ff_chain(const void* tmA, const void* tmB0, const void* tmS, const void* tmW, const FFChainParams& p, int kind,
                            cudaStream_t s);

// ---- whole Patcher / Unpatcher resnet at level 0 for few channels (resnet_small.cu) ---------------------------------------
struct ResnetSmallParams {
  const float* x; int Cin;          // input [B * L][Cin] fp32 token-major
  int L, Cout;
  const float* aff1;                // [2 * Cin] gamma | beta of block1.groupnorm (one group)
  const float* w1; const float* b1; // conv1 packed [Cout][3 * Cin] (k = tap * Cin + ci), [Cout]
  const float* ws; const float* bs; // 1x1 skip projection [Cout][Cin], [Cout]; null = identity (Cin == Cout)
  const float* aff2; int aff2_stride; const int* call_idx;   // block2.groupnorm affine with FiLM folded in, per denoiser call
  const float* w2; const float* b2; // conv2 packed [Cout][3 * Cout], [Cout]    (mode 0)
  float* out;                       // mode 0: block output [B * L][Cout] fp32; mode 1: the skip / residual term, same shape
  void* a2op; int kind;             // mode 1: SiLU(FiLM(GN(h1))) in the operand dtype (1 tf32, 2 bf16) for the conv2 GEMM
  int B; int mode;                  // 0: whole block; 1: up to the input of conv2
  float eps;
  int split1, split2;               // run conv1 + skip / conv2 as 3xTF32 (hi/lo operand split: fp32-grade) instead of one tf32 pass
};
bool resnet_small_supported(int L, int Cin, int Cout, int groups, bool proj, int mode);
cudaError_t init_resnet_small();
cudaError_t launch_resnet_small(const ResnetSmallParams& p, cudaStream_t s);

// ---- tensor-core GEMM (gemm_tc.cu) --------------------------------------------------------------
// kind: 1 = tf32, 2 = bf16.  Wtc must hold the weights pre-converted by convert_weights_tc().
cudaError_t launch_gemm_tc(const GemmParams& p, int kind, cudaStream_t s);
cudaError_t convert_weights_tc(const float* W, void* Wtc, long long n, int kind, cudaStream_t s);
bool gemm_tc_supported(const GemmParams& p);

}  // namespace mdt
