// plan.cu -- host side of the C ABI: weight repacking, the UNet op program, the ADPM2 step driver.
//
// A plan turns the reference state_dict into kernel-native weight layouts once, builds a flat
// list of kernel launches ("program") for one denoiser-network evaluation, and drives the
// (timesteps - 1)-iteration ADPM2 loop entirely on the device: every scalar the loop needs
// lives in a device table indexed by a device-side call counter, so one captured CUDA graph of
// a single iteration is replayed for the whole run with no host synchronisation.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mdt_b200.h"
#include "kernels.cuh"

using namespace mdt;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

struct MdtError {
  int code;
  std::string msg;
};
[[noreturn]] static void raise(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw MdtError{code, buf};
}
#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) raise(MDT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// program representation
// ------------------------------------------------------------------------------------------------
enum OpType { OP_GEMM, OP_GN_STATS, OP_ROW_STATS, OP_ATTN, OP_UPGATHER, OP_PERMUTE, OP_GN_APPLY, OP_LN_APPLY, OP_GEMM_TMA, OP_GEMM_ATTN, OP_DUP_ROWS, OP_GEMM_FF, OP_ATTN_LAYER, OP_RESNET_SMALL };

struct Op {
  OpType type = OP_GEMM;
  int rps = 0;  // output rows per sample (GEMM M = B_eff * rps; row stats rows = B_eff * rps)
  GemmParams g{};
  NormStatsParams ns{};
  AttnParams at{};
  int attn_kind = 0;   // storage type of q/k/v/o (0 fp32, 1 tf32-rounded fp32, 2 bf16)
  GnApplyParams ga{};
  LnApplyParams la{};
  TmaGemmParams tg{};
  GemmAttnParams gat{};
  FFChainParams ffc{};
  AttnLayerParams al{};
  ResnetSmallParams rs{};
  alignas(64) unsigned char tmA[128];
  alignas(64) unsigned char tmB[128];
  alignas(64) unsigned char tmC[128];
  alignas(64) unsigned char tmD[128];
  bool cross = false;  // attention reads the precomputed conditioning K/V
  bool umma_core = false;  // fused attention with the softmax-attention core on tcgen05 (gemm_attn_umma.cu)
  bool frag = false;       // attention layer with q / k fragments straight from TMEM (gemm_attn_frag.cu)
  int cross_layer = -1;
  // upsample gather / permute
  const float* in0 = nullptr; const float* in1 = nullptr; const float* in2 = nullptr; float* out = nullptr;
  int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
  // debug tap after this op
  std::string tap;
  const float* tap_ptr = nullptr;
  int tap_rps = 0, tap_c = 0;
  // classifier-free guidance: ops ahead of the first cross-attention see identical inputs in the conditional and the
  // null branch (modules.py:1250-1251 call the same x, time), so they run on the conditional rows only
  bool half = false;
};

struct Film {   // one ResnetBlock1d's FiLM table
  const float* w_ss; const float* b_ss;  // Linear(mapping -> 2C)
  const float* gamma; const float* beta; // block2.groupnorm affine
  int C;
  float* ss;   // [max_calls][2C] raw scale|shift
  float* aff;  // [max_calls][2C] folded (gamma * (1 + scale), beta * (1 + scale) + shift)
};

struct CrossLayer {
  const float* wkv; const float* bkv;  // folded norm_context -> to_kv  [2*Hd][F], [2*Hd]
  float* kv_cond;                      // [max_batch * n_ctx_max][2*Hd]
  float* kv_null;                      // [n_ctx_max][2*Hd]
  void* kv_cond_op = nullptr;          // operand-dtype copies read by the attention kernel in the tensor-core modes
  void* kv_null_op = nullptr;
  void* kvf_cond = nullptr;            // fragment-ordered tf32 copies for the packed cross-attention path (L <= 8), 8 KB per (sample, head)
  void* kvf_null = nullptr;
  int kperm = 0;                       // K fragments packed in the permuted k order of gemm_attn_frag.cu
};

struct mdt_plan {
  mdt_config cfg{};
  int device = 0;
  int prec = 0;
  int max_calls = 0;
  std::unordered_map<std::string, std::pair<const float*, int64_t>> tensors;
  // device slab for weights
  char* wslab = nullptr; size_t wcap = 0, wused = 0;
  std::vector<void*> allocs;  // activation buffers
  size_t act_bytes = 0;
  long long launches = 0;
  bool serpentine = false;

  // model dims
  int P = 0, L0 = 0, Hd = 0, F = 0, Bmax = 0, Beff_max = 0;
  size_t S_act = 0;

  // programs
  std::vector<Op> unet;
  std::vector<Film> films;
  std::vector<CrossLayer> cross;
  // act pool
  std::vector<float*> pool_free;
  // named buffers
  float *xin = nullptr, *net_out = nullptr, *x = nullptr, *xmid = nullptr;
  float *emb = nullptr, *emb_stats = nullptr, *emb_null = nullptr, *emb_null_stats = nullptr;
  float *qkv = nullptr, *att = nullptr, *qc = nullptr, *ff = nullptr, *upy = nullptr;
  float *gn_stats = nullptr, *row_stats = nullptr;
  void* attn_scratch = nullptr;   // CTA-private head-output slots of the fused attention-layer kernel (gemm_attn_layer.cu)
  void* ff_scratch = nullptr;     // CTA-private hidden blocks of the FeedForward chain kernel (gemm_chain.cu)
  size_t ff_scratch_mid = 0;
  // time path
  float *t_calls = nullptr, *t_feat = nullptr, *t_a = nullptr, *t_b = nullptr, *t_map = nullptr;
  const float *w_time_freq = nullptr, *w_time = nullptr, *b_time = nullptr, *w_map0 = nullptr, *b_map0 = nullptr,
              *w_map2 = nullptr, *b_map2 = nullptr;
  // encoder
  const float *w_fc1 = nullptr, *b_fc1 = nullptr, *inv_freq = nullptr, *w_null_emb = nullptr;
  // step driver
  IterScalars* d_iters = nullptr;
  IterScalars* h_iters = nullptr;  // pinned
  float* h_tcalls = nullptr;       // pinned
  int* d_call = nullptr;
  RunParams* d_run = nullptr;      // seed / first sample index / injected-noise base / guidance scale of the running chunk (captured kernels read it)
  cudaEvent_t staged = nullptr;    // the pinned staging tables of the previous call have been copied to the device
  bool ctx_pre_encoded = false;    // cond_dev holds the encoded embedding [B, n_ctx, F] (XDiffusion_x.sample(embedding=...))
  int sampler_mode = 0;            // 0: ADPM2 / AEuler rows, 1: KarrasSampler rows (mdt_plan_set_sampler_mode)
  float init_noise_scale = 0.f, init_sigma = 0.f;
  float* daux = nullptr;           // slope of denoiser call A (KarrasSampler)
  int n_ctx_cur = 0;
  // graph cache: key (Bc, n_ctx, cfg, has_step_noise, n_iters, single_call); everything else is device resident
  struct GraphEntry { cudaGraphExec_t exec; long long launches; };
  std::map<std::vector<long long>, GraphEntry> graphs;
  bool use_graph = true;
  // taps
  bool taps_on = false;
  std::map<std::string, std::vector<float>> tap_store;
};

// ------------------------------------------------------------------------------------------------
// builder
// ------------------------------------------------------------------------------------------------
struct Src {  // activation source: up to two channel segments
  const float* p0; int c0; const float* p1; int c1; float scale1;
  int C() const { return c0 + c1; }
};

struct Builder {
  mdt_plan& pl;
  explicit Builder(mdt_plan& p) : pl(p) {}

  // ---- state_dict access
  const float* T(const std::string& name, int64_t numel) {
    auto it = pl.tensors.find(name);
    if (it == pl.tensors.end()) raise(MDT_ERR_MISSING, "state_dict tensor '%s' is missing", name.c_str());
    if (it->second.second != numel)
      raise(MDT_ERR_MISSING, "state_dict tensor '%s' has %lld elements, expected %lld", name.c_str(),
            (long long)it->second.second, (long long)numel);
    return it->second.first;
  }
  bool has(const std::string& name) { return pl.tensors.count(name) != 0; }

  // ---- device weight slab
  float* upload(const float* host, size_t n) {
    const size_t bytes = (n * sizeof(float) + 255) & ~(size_t)255;
    if (pl.wused + bytes > pl.wcap) raise(MDT_ERR_OOM, "weight slab exhausted (%zu + %zu > %zu)", pl.wused, bytes, pl.wcap);
    float* d = reinterpret_cast<float*>(pl.wslab + pl.wused);
    pl.wused += bytes;
    CK(cudaMemcpy(d, host, n * sizeof(float), cudaMemcpyHostToDevice));
    return d;
  }
  float* upload(const std::vector<float>& v) { return upload(v.data(), v.size()); }
  void* slab_alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (pl.wused + bytes > pl.wcap) raise(MDT_ERR_OOM, "weight slab exhausted");
    void* d = pl.wslab + pl.wused;
    pl.wused += bytes;
    return d;
  }
  // tensor-core copy of a packed [N][K] matrix (tf32-rounded fp32 bits, bf16 or fp16), when the mode needs it
  const void* tc_copy(const float* dW, size_t n) {
    if (pl.prec == MDT_PREC_FP32) return nullptr;
    void* d = slab_alloc(n * (pl.prec >= MDT_PREC_BF16 ? 2 : 4));
    CK(convert_weights_tc(dW, d, (long long)n, pl.prec, 0));
    return d;
  }

  // ---- activation buffers
  float* dalloc(size_t floats) {
    void* p = nullptr;
    const size_t bytes = floats * sizeof(float);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) raise(MDT_ERR_OOM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    pl.allocs.push_back(p);
    pl.act_bytes += bytes;
    return reinterpret_cast<float*>(p);
  }
  float* acquire() {
    if (!pl.pool_free.empty()) { float* p = pl.pool_free.back(); pl.pool_free.pop_back(); return p; }
    return dalloc(pl.S_act * pl.Beff_max);
  }
  void release(const float* p) { pl.pool_free.push_back(const_cast<float*>(p)); }

  // ---- weight packing helpers (host)
  // Conv1d weight [N][C][taps] -> [N][taps*C]  (k = tap*C + c)
  std::vector<float> pack_conv(const float* w, int N, int C, int taps) {
    std::vector<float> o((size_t)N * C * taps);
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < C; ++c)
        for (int t = 0; t < taps; ++t) o[((size_t)n * taps + t) * C + c] = w[((size_t)n * C + c) * taps + t];
    return o;
  }
  // y = W (x * gamma + beta) + b  ==>  W' = W diag(gamma), b' = b + W beta
  void fold_affine(std::vector<float>& W, std::vector<float>& b, int N, int K, const float* gamma, const float* beta) {
    for (int n = 0; n < N; ++n) {
      double acc = 0.0;
      for (int k = 0; k < K; ++k) {
        acc += (double)W[(size_t)n * K + k] * (double)beta[k];
        W[(size_t)n * K + k] *= gamma[k];
      }
      b[n] = (float)((double)b[n] + acc);
    }
  }

  // ---- op emitters
  ALoad make_aload(const Src& s, int L_in, int L_out, int taps, int stride, int pad) {
    ALoad a{};
    a.src0 = s.p0; a.src1 = s.p1; a.c0 = s.c0; a.c1 = s.c1; a.C = s.c0 + s.c1; a.scale1 = s.scale1;
    a.L_in = L_in; a.L_out = L_out; a.taps = taps; a.stride = stride; a.pad = pad;
    a.stats = nullptr; a.stats_mode = 0; a.groups = 1; a.cpg = a.C; a.aff = nullptr; a.aff_call_stride = 0;
    a.call_idx = nullptr; a.silu = 0;
    return a;
  }
  // ---- shared classifier-free-guidance prefix bookkeeping
  bool prefix = true;                      // ops emitted while true run on the conditional rows only
  struct Keep { float* p; int rps; int C; };
  std::vector<Keep> prefix_keep;           // fp32 tensors produced in the prefix that are read after it (skips)
  Op& emit(std::vector<Op>& prog, const Op& op) { prog.push_back(op); prog.back().half = prefix; return prog.back(); }
  // end of the prefix: replicate every live tensor into the null-branch rows, then continue on all rows
  void end_prefix(std::vector<Op>& prog, float* live, int rps, int C) {
    if (!prefix) return;
    prefix = false;
    auto dup = [&](float* ptr, int r, int c) {
      Op op; op.type = OP_DUP_ROWS; op.out = ptr; op.i0 = r; op.i1 = c;
      prog.push_back(op);
    };
    if (live) dup(live, rps, C);
    for (const Keep& k : prefix_keep) if (k.p != live) dup(k.p, k.rps, k.C);
    prefix_keep.clear();
  }

  Op gemm_op(const ALoad& a, const float* dW, const void* dWtc, const float* dbias, int N, int act, const float* res,
             float* out, int rps) {
    Op op; op.type = OP_GEMM; op.rps = rps;
    op.g.a = a; op.g.W = dW; op.g.Wtc = dWtc; op.g.bias = dbias; op.g.M = 0; op.g.N = N; op.g.K = a.taps * a.C;
    op.g.act = act; op.g.res = res; op.g.ldres = N; op.g.C = out; op.g.ldc = N;
    return op;
  }
  void gn_stats(std::vector<Op>& prog, const Src& s, int L, int groups, float eps) {
    Op op; op.type = OP_GN_STATS;
    op.ns.src0 = s.p0; op.ns.src1 = s.p1; op.ns.c0 = s.c0; op.ns.c1 = s.c1; op.ns.scale1 = s.scale1;
    op.ns.L = L; op.ns.groups = groups; op.ns.eps = eps; op.ns.stats = pl.gn_stats; op.ns.rows = 0;
    if (s.C() % groups) raise(MDT_ERR_INVALID, "GroupNorm: %d channels not divisible by %d groups", s.C(), groups);
    emit(prog, op);
  }
  void row_stats(std::vector<Op>& prog, const float* src, int C, int rps, float* stats) {
    Op op; op.type = OP_ROW_STATS; op.rps = rps;
    op.ns.src0 = src; op.ns.src1 = nullptr; op.ns.c0 = C; op.ns.c1 = 0; op.ns.scale1 = 1.f; op.ns.L = 1; op.ns.groups = 1;
    op.ns.eps = 1e-5f; op.ns.stats = stats; op.ns.rows = 0;
    emit(prog, op);
  }
  void set_tap(std::vector<Op>& prog, const std::string& name, const float* ptr, int rps, int c) {
    Op& op = prog.back(); op.tap = name; op.tap_ptr = ptr; op.tap_rps = rps; op.tap_c = c;
  }

  // ---- TMA path helpers (tensor-core precisions only)
  bool tma() const { return pl.prec != MDT_PREC_FP32; }
  int esz() const { return pl.prec >= MDT_PREC_BF16 ? 2 : 4; }
  void* op_off(void* base, size_t elems) const { return reinterpret_cast<char*>(base) + elems * (size_t)esz(); }
  bool tma_ok(int C, int L, int N) const { return tma() && gemm_tma_shape_ok(pl.prec, C, L, N); }

  void emit_gn_apply(std::vector<Op>& prog, const Src& in, int L, int groups, float eps, const float* aff, int aff_stride,
                     const int* call_idx, int silu, void* out, void* raw) {
    Op op; op.type = OP_GN_APPLY;
    op.ga.src0 = in.p0; op.ga.src1 = in.p1; op.ga.c0 = in.c0; op.ga.c1 = in.c1; op.ga.scale1 = in.scale1;
    op.ga.L = L; op.ga.groups = groups; op.ga.eps = eps; op.ga.aff = aff; op.ga.aff_call_stride = aff_stride;
    op.ga.call_idx = call_idx; op.ga.silu = silu; op.ga.out = out; op.ga.raw = raw; op.ga.B = 0;
    emit(prog, op);
  }
  void emit_ln_apply(std::vector<Op>& prog, const float* src, int C, int L, void* out) {
    Op op; op.type = OP_LN_APPLY; op.rps = L;
    op.la.src = src; op.la.C = C; op.la.eps = 1e-5f; op.la.out = out; op.la.rows = 0;
    emit(prog, op);
  }
  // A: operand-dtype activation [Beff_max * L][C]; dW32: packed fp32 weights [N][taps * C] already on the device
  void emit_gemm_tma(std::vector<Op>& prog, const void* A, int C, int L, int taps, const float* dW32, const float* bias, int N,
                     int act, const float* res, float* C32, void* Cop, int ldcop = 0, bool cop_ln = false) {
    Op op; op.type = OP_GEMM_TMA; op.rps = L;
    const int kch = pl.prec == MDT_PREC_TF32 ? 32 : 64;
    TmaGemmParams& g = op.tg;
    g.M = 0; g.N = N; g.BN = tma_pick_bn(N); g.taps = taps; g.pad = taps / 2; g.kchunks = C / kch; g.C = C;
    g.L = L; g.Lb = L >= 128 ? 128 : L; g.Sb = L >= 128 ? 1 : 128 / L;
    g.bias = bias; g.act = act; g.res = res; g.ldres = N; g.C32 = C32; g.ldc = N; g.Cop = Cop; g.ldcop = ldcop ? ldcop : N;
    g.cop_ln = cop_ln ? 1 : 0; g.ln_eps = 1e-5f;
    const void* wop = tc_copy(dW32, (size_t)N * taps * C);
    if (make_tmap_act(op.tmA, A, pl.prec, C, L, (long long)pl.Beff_max) != 0) raise(MDT_ERR_CUDA, "cuTensorMapEncodeTiled(activation C=%d L=%d) failed", C, L);
    if (make_tmap_weight(op.tmB, wop, pl.prec, (long long)taps * C, N, g.BN) != 0) raise(MDT_ERR_CUDA, "cuTensorMapEncodeTiled(weight K=%d N=%d) failed", taps * C, N);
    emit(prog, op);
  }

  // LayerNorm in the epilogue of the GEMM that produces the row (gemm_tma.cu cop_ln): the row block must hold whole rows
  bool ln_epilogue_ok(int N) const {
    static const bool off = getenv("MDT_NO_LN_EPILOGUE") != nullptr;
    return !off && N >= 128 && tma_pick_bn(N) == N;
  }

  // FeedForward chain (gemm_chain.cu): t = t + b2 + GELU(x_op W0^T + b0) W2^T; the hidden activation goes through CTA-private
  // L2 scratch; cop receives the raw operand copy of the new t (cop_ln = 0) or LayerNorm(t) for the next attention layer
  void emit_ff_chain(std::vector<Op>& prog, const void* xop, int C, int mid, int L, const float* dW0, const float* b0,
                     const float* dW2, const float* b2, float* t, void* cop, int cop_ln) {
    Op op; op.type = OP_GEMM_FF; op.rps = L;
    FFChainParams& g = op.ffc;
    g.M = 0; g.C = C; g.mid = mid; g.L = L; g.Sb = L >= 128 ? 1 : 128 / L; g.b0 = b0; g.b2 = b2;
    g.res = t; g.ldres = C; g.C32 = t; g.ldc = C; g.Cop = cop; g.ldcop = C; g.cop_ln = cop_ln; g.ln_eps = 1e-5f;
    g.scratch = pl.ff_scratch;   // sized in build() for the widest hidden layer
    const void* w0 = tc_copy(dW0, (size_t)mid * C);
    const void* w2 = tc_copy(dW2, (size_t)C * mid);
    if (make_tmap_act(op.tmA, xop, pl.prec, C, L, (long long)pl.Beff_max) != 0) raise(MDT_ERR_CUDA, "cuTensorMapEncodeTiled(ff activation) failed");
    if (make_tmap_weight(op.tmB, w0, pl.prec, (long long)C, mid, 128) != 0) raise(MDT_ERR_CUDA, "cuTensorMapEncodeTiled(ff W0) failed");
    if (make_tmap_act(op.tmC, pl.ff_scratch, pl.prec, mid, 2 * 128, (long long)ff_chain_sms()) != 0) raise(MDT_ERR_CUDA, "cuTensorMapEncodeTiled(ff scratch) failed");
    if (make_tmap_weight(op.tmD, w2, pl.prec, (long long)mid, C, C) != 0) raise(MDT_ERR_CUDA, "cuTensorMapEncodeTiled(ff W2) failed");
    emit(prog, op);
  }

  // fused per-head projection + attention (gemm_attn.cu); dW32 = [heads * BN][C] fp32 on the device, head-major
  void emit_gemm_attn(std::vector<Op>& prog, const void* A, int C, int L, const float* dW32, const float* bias, int cross,
                      int cross_layer, const void* kc, const void* kn, const void* kvf_c = nullptr, const void* kvf_n = nullptr) {
    Op op; op.type = OP_GEMM_ATTN; op.rps = L; op.cross = cross != 0; op.cross_layer = cross_layer;
    const int kch = pl.prec == MDT_PREC_TF32 ? 32 : 64;
    const int heads = pl.cfg.heads, d = pl.cfg.head_features, BN = cross ? d : 3 * d;
    GemmAttnParams& g = op.gat;
    g.M = 0; g.heads = heads; g.d = d; g.kchunks = C / kch; g.C = C; g.L = L; g.Sb = 128 / L; g.cross = cross;
    g.bias = bias; g.scale = 1.0f / sqrtf((float)d); g.att = pl.att; g.ldo = heads * d;
    g.kc = kc; g.kn = kn; g.ldkv = 2 * heads * d; g.kv_sample_stride = 0; g.n_cond = 0; g.nk = L; g.kv_fp32 = 1;
    g.kvf_c = kvf_c; g.kvf_n = kvf_n;
    const void* wop = tc_copy(dW32, (size_t)heads * BN * C);
    if (make_tmap_act(op.tmA, A, pl.prec, C, L, (long long)pl.Beff_max) != 0) raise(MDT_ERR_CUDA, "cuTensorMapEncodeTiled(attn activation) failed");
    if (make_tmap_weight(op.tmB, wop, pl.prec, (long long)C, heads * BN, BN) != 0) raise(MDT_ERR_CUDA, "cuTensorMapEncodeTiled(attn weight) failed");
    // Attention core: per-sample / packed mma.sync (gemm_attn.cu, default) or tcgen05 (block-diagonal S / P V UMMAs over the whole
    // 128-row tile, gemm_attn_umma.cu).  Measured on the README model at B = 4096 (profiles/README.md): the UMMA core's serial
    // stage -> S -> softmax -> P V -> store chain is ~8 k cycles per tile, which loses at L = 16 (277 vs 210 us) and, since the
    // mma.sync kernel packs 16 / L short samples into one block-diagonal m16 tile, at L = 4 too (78 vs 69 us).  It stays available:
    // MDT_UMMA_ATTN = short (L <= 8) | all.
    const char* um = getenv("MDT_UMMA_ATTN");
    const bool um_all = um && um[0] == 'a' && um[1] == 'l', um_short = um && um[0] == 's';
    op.umma_core = (um_all || (um_short && L <= 8)) && gemm_attn_umma_supported(pl.prec, C, L, heads, d, cross, pl.cfg.ctx_max_length);
    const char* ps = getenv("MDT_PACK_SELF");
    g.pack_self = (!cross && L >= 2 && L <= 8 && !(ps && ps[0] == '0')) ? 1 : 0;
    emit(prog, op);
  }

  // whole attention layer in one kernel (gemm_attn_layer.cu): projection + attention + out-projection + bias + residual into t
  // (and the operand copy `cop` of the new token stream when the next op reads it raw)
  bool layer_ok(int C, int L, int cross, bool packed_cross) const {
    if (pl.prec == MDT_PREC_FP32 || getenv("MDT_NO_FUSED_LAYER")) return false;
    // measured on the README model at B = 4096 (profiles/README.md): the fused layer wins where the (rows x heads*d) attention tensor
    // would not fit L2 (level 1: 256 vs 293 us self, 259 vs 272 us cross) and loses at level 2, where only 256 row blocks exist for
    // 148 SMs and the unfused pair keeps its intermediate in L2 anyway (131 vs 113 us); default: width <= 128 only
    const char* mc = getenv("MDT_FUSED_LAYER_MAXC");
    if (C > (mc ? atoi(mc) : 128)) return false;
    if (cross && !packed_cross) return false;
    return attn_layer_supported(pl.prec, C, L, pl.cfg.heads, pl.cfg.head_features, cross, C);
  }
  bool frag_ok(int C, int L, int cross, bool fused = true) const {
    const char* e = getenv("MDT_ATTN_FRAG");
    if (e && e[0] == '0') return false;
    if (!fused && getenv("MDT_NO_FRAG_UNFUSED")) return false;
    return attn_frag_supported(pl.prec, C, L, pl.cfg.heads, pl.cfg.head_features, cross, fused ? C : 0);
  }
  void emit_attn_layer(std::vector<Op>& prog, const void* A, int C, int L, const float* dW32, const float* bias_q, int cross,
                       int cross_layer, const void* kn_flag, const void* kvf_c, const void* kvf_n, const float* dWo32,
                       const float* bias_o, float* t, void* cop, bool fused = true, bool cop_ln = false) {
    // fused == false: gemm_attn_frag.cu writes the head outputs to pl.att and the caller emits the out-projection GEMM itself
    Op op; op.type = OP_ATTN_LAYER; op.rps = L; op.cross = cross != 0; op.cross_layer = cross_layer;
    const int kch = pl.prec == MDT_PREC_TF32 ? 32 : 64;
    const int heads = pl.cfg.heads, d = pl.cfg.head_features, BN = cross ? d : 3 * d, Hd = heads * d;
    AttnLayerParams& y = op.al;
    GemmAttnParams& g = y.a;
    g.M = 0; g.heads = heads; g.d = d; g.kchunks = C / kch; g.C = C; g.L = L; g.Sb = 128 / L; g.cross = cross;
    g.bias = bias_q; g.scale = 1.0f / sqrtf((float)d); g.att = fused ? nullptr : (void*)pl.att; g.ldo = fused ? 0 : Hd;
    g.kc = nullptr; g.kn = kn_flag; g.ldkv = 2 * Hd; g.kv_sample_stride = 0; g.n_cond = 0; g.nk = L; g.kv_fp32 = 1;
    g.kvf_c = kvf_c; g.kvf_n = kvf_n;
    const char* ps = getenv("MDT_PACK_SELF");
    g.pack_self = (!cross && L >= 2 && L <= 8 && !(ps && ps[0] == '0')) ? 1 : 0;
    y.Cout = C; y.bias_o = bias_o; y.res = t; y.ldres = C; y.C32 = t; y.ldc = C; y.Cop = cop; y.ldcop = C; y.fused = fused ? 1 : 0;
    { const char* lh = getenv("MDT_L2_HINT"); y.l2_hint = (lh && lh[0] == '1') ? 1 : 0; }
    if (fused && !pl.attn_scratch) {
      const size_t bytes = attn_layer_scratch_bytes(pl.prec, heads, d);
      pl.attn_scratch = dalloc((bytes + 3) / 4);
      CK(cudaMemset(pl.attn_scratch, 0, bytes));
    }
    y.scratch = pl.attn_scratch;
    const void* wop = tc_copy(dW32, (size_t)heads * BN * C);
    if (make_tmap_act(op.tmA, A, pl.prec, C, L, (long long)pl.Beff_max) != 0) raise(MDT_ERR_CUDA, "cuTensorMapEncodeTiled(layer activation) failed");
    if (make_tmap_weight(op.tmB, wop, pl.prec, (long long)C, heads * BN, BN) != 0) raise(MDT_ERR_CUDA, "cuTensorMapEncodeTiled(layer weight) failed");
    if (fused) {
      const void* woo = tc_copy(dWo32, (size_t)C * Hd);
      if (make_tmap_act(op.tmC, pl.attn_scratch, pl.prec, d, attn_layer_slots(heads) * 128, (long long)attn_layer_sms()) != 0) raise(MDT_ERR_CUDA, "cuTensorMapEncodeTiled(layer scratch) failed");
      if (make_tmap_weight(op.tmD, woo, pl.prec, (long long)Hd, C, C) != 0) raise(MDT_ERR_CUDA, "cuTensorMapEncodeTiled(layer out-projection) failed");
    } else {
      memset(op.tmC, 0, sizeof op.tmC); memset(op.tmD, 0, sizeof op.tmD);
    }
    op.frag = fused ? frag_ok(C, L, cross) : true;
    y.cop_ln = (cop_ln && op.frag) ? 1 : 0; y.ln_eps = 1e-5f;
    if (cop_ln && !op.frag) raise(MDT_ERR_INVALID, "LayerNorm tail requested for an attention layer outside gemm_attn_frag.cu");
    // fp16 operands always take the f16 attention core (gemm_attn_frag.cu instantiates only that pairing)
    { const char* hf = getenv("MDT_ATTN_F16"); y.f16 = (op.frag && (pl.prec == MDT_PREC_F16 || !(hf && hf[0] == '0'))) ? 1 : 0; }
    if (cross && cross_layer >= 0) pl.cross[cross_layer].kperm = op.frag ? (y.f16 ? 2 : 1) : 0;
    emit(prog, op);
  }

  // ResnetBlock1d (modules.py:145-205).  Returns the output buffer (acquired from the pool).
  float* resnet(std::vector<Op>& prog, const std::string& pre, const Src& in, int L, int Cout, int groups,
                float* forced_out = nullptr, void** out_op = nullptr) {
    if (out_op) *out_op = nullptr;
    const int Cin = in.C();
    const int M = pl.cfg.mapping_features;
    // block1: GN(groups) -> SiLU -> conv3
    std::vector<float> aff1(2 * (size_t)Cin);
    memcpy(aff1.data(), T(pre + "block1.groupnorm.weight", Cin), Cin * sizeof(float));
    memcpy(aff1.data() + Cin, T(pre + "block1.groupnorm.bias", Cin), Cin * sizeof(float));
    const float* d_aff1 = upload(aff1);
    auto w1 = pack_conv(T(pre + "block1.project.weight", (int64_t)Cout * Cin * 3), Cout, Cin, 3);
    const float* d_w1 = upload(w1);
    const float* d_b1 = upload(T(pre + "block1.project.bias", Cout), Cout);
    const bool ok1 = tma_ok(Cin, L, Cout) && gn_apply_supported(L, Cin, groups);
    const bool ok2 = tma_ok(Cout, L, Cout) && gn_apply_supported(L, Cout, groups);
    const bool proj = has(pre + "to_out.weight");
    // Patcher / Unpatcher resnets with few channels (resnet_small.cu): the whole block in one kernel when the second conv is small
    // too (to_out), or everything up to the tensor-core conv2 (to_in) -- instead of GroupNorm passes and sliver-of-a-tile GEMMs
    const bool small_on = tma() && in.c1 == 0 && groups == 1 && !getenv("MDT_NO_RESNET_SMALL");
    const bool small_full = small_on && !out_op && Cout <= 32 && resnet_small_supported(L, Cin, Cout, groups, proj, 0);
    const bool small_head = small_on && !small_full && ok2 && Cin <= 32 && resnet_small_supported(L, Cin, Cout, groups, proj, 1);
    if (small_full || small_head) {
      Film f{};
      f.w_ss = upload(T(pre + "to_scale_shift.to_scale_shift.1.weight", (int64_t)2 * Cout * M), (size_t)2 * Cout * M);
      f.b_ss = upload(T(pre + "to_scale_shift.to_scale_shift.1.bias", 2 * Cout), 2 * Cout);
      f.gamma = upload(T(pre + "block2.groupnorm.weight", Cout), Cout);
      f.beta = upload(T(pre + "block2.groupnorm.bias", Cout), Cout);
      f.C = Cout;
      f.ss = dalloc((size_t)pl.max_calls * 2 * Cout);
      f.aff = dalloc((size_t)pl.max_calls * 2 * Cout);
      pl.films.push_back(f);
      auto w2 = pack_conv(T(pre + "block2.project.weight", (int64_t)Cout * Cout * 3), Cout, Cout, 3);
      const float* d_w2 = upload(w2);
      const float* d_b2 = upload(T(pre + "block2.project.bias", Cout), Cout);
      float* out = forced_out ? forced_out : acquire();
      Op op; op.type = OP_RESNET_SMALL;
      ResnetSmallParams& r = op.rs;
      r.x = in.p0; r.Cin = Cin; r.L = L; r.Cout = Cout; r.aff1 = d_aff1; r.w1 = d_w1; r.b1 = d_b1;
      r.ws = nullptr; r.bs = nullptr;
      if (proj) {
        r.ws = upload(T(pre + "to_out.weight", (int64_t)Cout * Cin), (size_t)Cout * Cin);
        r.bs = upload(T(pre + "to_out.bias", Cout), Cout);
      }
      r.aff2 = f.aff; r.aff2_stride = 2 * Cout; r.call_idx = pl.d_call; r.w2 = d_w2; r.b2 = d_b2; r.out = out;
      r.a2op = nullptr; r.kind = pl.prec; r.B = 0; r.mode = small_full ? 0 : 1; r.eps = 1e-5f;
      // 3xTF32 where the rounding would land straight on the sampler state: the first conv of the network input (to_in) and the last
      // conv of the network output (to_out); to_out's first conv / skip keep the single tf32 pass they had as TMA GEMMs
      r.split1 = small_full ? 0 : 1; r.split2 = 1;
      if (small_full) { emit(prog, op); return out; }
      float* a2f = acquire();
      op.rs.a2op = a2f;
      emit(prog, op);
      void* oc = nullptr;
      if (out_op) { oc = acquire(); *out_op = oc; }
      emit_gemm_tma(prog, a2f, Cout, L, 3, d_w2, d_b2, Cout, 0, out, out, oc);
      release(a2f);
      return out;
    }
    // block2's GroupNorm + FiLM + SiLU can run inside conv1's epilogue when a 32-row x 32-column epilogue block holds
    // whole (sample, group) sets: then h1 never goes to HBM and the separate normalisation pass disappears
    const int cpg2 = Cout / groups;
    const bool fuse_gn = ok1 && ok2 && !getenv("MDT_NO_FUSED_GN") && tma_pick_bn(Cout) >= 128 && L <= 32 && (32 % L) == 0 &&
                         (cpg2 == 16 || cpg2 == 32);
    float* h1 = fuse_gn ? nullptr : acquire();
    float* a1 = nullptr; float* raw = nullptr; float* a2f = nullptr;
    if (ok1) {
      a1 = acquire();
      if (proj) raw = acquire();
      emit_gn_apply(prog, in, L, groups, 1e-5f, d_aff1, 0, nullptr, 1, a1, raw);
      if (fuse_gn) {
        a2f = acquire();
        emit_gemm_tma(prog, a1, Cin, L, 3, d_w1, d_b1, Cout, 0, nullptr, nullptr, a2f);   // gn_* fields are patched below
      } else {
        emit_gemm_tma(prog, a1, Cin, L, 3, d_w1, d_b1, Cout, 0, nullptr, h1, nullptr);
      }
    } else {
      gn_stats(prog, in, L, groups, 1e-5f);
      ALoad a = make_aload(in, L, L, 3, 1, 1);
      a.stats = pl.gn_stats; a.stats_mode = 2; a.groups = groups; a.cpg = Cin / groups; a.aff = d_aff1; a.silu = 1;
      emit(prog, gemm_op(a, d_w1, tc_copy(d_w1, w1.size()), d_b1, Cout, 0, nullptr, h1, L));
    }
    const size_t conv1_index = prog.size() - 1;
    // FiLM table of this block
    Film f{};
    f.w_ss = upload(T(pre + "to_scale_shift.to_scale_shift.1.weight", (int64_t)2 * Cout * M), (size_t)2 * Cout * M);
    f.b_ss = upload(T(pre + "to_scale_shift.to_scale_shift.1.bias", 2 * Cout), 2 * Cout);
    f.gamma = upload(T(pre + "block2.groupnorm.weight", Cout), Cout);
    f.beta = upload(T(pre + "block2.groupnorm.bias", Cout), Cout);
    f.C = Cout;
    f.ss = dalloc((size_t)pl.max_calls * 2 * Cout);
    f.aff = dalloc((size_t)pl.max_calls * 2 * Cout);
    pl.films.push_back(f);
    if (fuse_gn) {
      TmaGemmParams& g1 = prog[conv1_index].tg;
      g1.gn_L = L; g1.gn_cpg = cpg2; g1.gn_eps = 1e-5f; g1.gn_aff = f.aff; g1.gn_aff_stride = 2 * Cout; g1.gn_call = pl.d_call;
    }
    // residual path
    float* out = forced_out ? forced_out : acquire();
    const float* res;
    if (proj) {
      const float* d_wo = upload(T(pre + "to_out.weight", (int64_t)Cout * Cin), (size_t)Cout * Cin);
      const float* d_bo = upload(T(pre + "to_out.bias", Cout), Cout);
      if (ok1) {
        emit_gemm_tma(prog, raw, Cin, L, 1, d_wo, d_bo, Cout, 0, nullptr, out, nullptr);
      } else {
        ALoad a = make_aload(in, L, L, 1, 1, 0);
        emit(prog, gemm_op(a, d_wo, tc_copy(d_wo, (size_t)Cout * Cin), d_bo, Cout, 0, nullptr, out, L));
      }
      res = out;
    } else {
      if (in.c1 != 0 || Cin != Cout) raise(MDT_ERR_INVALID, "resnet '%s': identity skip needs Cin == Cout", pre.c_str());
      res = in.p0;
    }
    // block2: GN -> FiLM -> SiLU -> conv3, + residual
    Src hs{h1, Cout, nullptr, 0, 1.f};
    auto w2 = pack_conv(T(pre + "block2.project.weight", (int64_t)Cout * Cout * 3), Cout, Cout, 3);
    const float* d_w2 = upload(w2);
    const float* d_b2 = upload(T(pre + "block2.project.bias", Cout), Cout);
    if (fuse_gn) {
      void* oc = nullptr;
      if (out_op) { oc = acquire(); *out_op = oc; }
      emit_gemm_tma(prog, a2f, Cout, L, 3, d_w2, d_b2, Cout, 0, res, out, oc);
      release(a2f);
    } else if (ok2) {
      float* a2 = a1 ? a1 : acquire();   // a1 is dead once conv1 has run
      emit_gn_apply(prog, hs, L, groups, 1e-5f, f.aff, 2 * Cout, pl.d_call, 1, a2, nullptr);
      void* oc = nullptr;
      if (out_op) { oc = acquire(); *out_op = oc; }
      emit_gemm_tma(prog, a2, Cout, L, 3, d_w2, d_b2, Cout, 0, res, out, oc);
      if (!a1) release(a2);
    } else {
      gn_stats(prog, hs, L, groups, 1e-5f);
      ALoad a = make_aload(hs, L, L, 3, 1, 1);
      a.stats = pl.gn_stats; a.stats_mode = 2; a.groups = groups; a.cpg = Cout / groups;
      a.aff = f.aff; a.aff_call_stride = 2 * Cout; a.call_idx = pl.d_call; a.silu = 1;
      emit(prog, gemm_op(a, d_w2, tc_copy(d_w2, w2.size()), d_b2, Cout, 0, res, out, L));
    }
    if (a1) release(a1);
    if (raw) release(raw);
    if (h1) release(h1);
    return out;
  }

  // Transformer1d (modules.py:469-524).  with_ctx == false -> self-attention only.
  // out_op (optional): receives an operand-dtype copy of the output when the tensor-core path is taken.
  float* transformer(std::vector<Op>& prog, const std::string& pre, const float* x, int L, int C, bool with_ctx,
                     void** out_op = nullptr) {
    const int Hd = pl.Hd, heads = pl.cfg.heads, d = pl.cfg.head_features, F = pl.F;
    const int mid = C * pl.cfg.ff_multiplier;
    if (out_op) *out_op = nullptr;
    const bool fast = tma_ok(C, L, C) && gn_apply_supported(L, C, 32) && tma_ok(C, L, 3 * Hd) && tma_ok(C, L, Hd) &&
                      tma_ok(Hd, L, C) && tma_ok(C, L, mid) && tma_ok(mid, L, C) && C <= 1024;
    const int akind = fast ? pl.prec : 0;
    const bool fuse_attn = Hd == heads * d && !getenv("MDT_NO_FUSED_ATTN");
    // to_in: GroupNorm(32, eps 1e-6) folded into the 1x1 conv
    std::vector<float> wi(T(pre + "to_in.1.weight", (int64_t)C * C), T(pre + "to_in.1.weight", (int64_t)C * C) + (size_t)C * C);
    std::vector<float> bi(T(pre + "to_in.1.bias", C), T(pre + "to_in.1.bias", C) + C);
    fold_affine(wi, bi, C, C, T(pre + "to_in.0.weight", C), T(pre + "to_in.0.bias", C));
    const float* d_wi = upload(wi); const float* d_bi = upload(bi);
    Src xs{x, C, nullptr, 0, 1.f};
    float* t = acquire();
    float* tn = fast ? acquire() : nullptr;   // normalised / raw operand copies of the token stream
    bool to_in_ln = false;
    if (fast) {
      emit_gn_apply(prog, xs, L, 32, 1e-6f, nullptr, 0, nullptr, 0, tn, nullptr);
      // block 0's self-attention reads LayerNorm(t): written by this GEMM's epilogue where a row block holds whole rows
      to_in_ln = ln_epilogue_ok(C) && has(pre + "blocks.0.attention.to_q.weight");
      emit_gemm_tma(prog, tn, C, L, 1, d_wi, d_bi, C, 0, nullptr, t, to_in_ln ? (void*)tn : nullptr, 0, to_in_ln);
    } else {
      gn_stats(prog, xs, L, 32, 1e-6f);
      ALoad a = make_aload(xs, L, L, 1, 1, 0);
      a.stats = pl.gn_stats; a.stats_mode = 2; a.groups = 32; a.cpg = C / 32;
      emit(prog, gemm_op(a, d_wi, tc_copy(d_wi, wi.size()), d_bi, C, 0, nullptr, t, L));
    }
    Src ts{t, C, nullptr, 0, 1.f};
    int nblocks = 0;
    while (has(pre + "blocks." + std::to_string(nblocks) + ".attention.to_q.weight")) ++nblocks;
    bool tn_is_ln = to_in_ln;   // tn already holds LayerNorm(t) (written by the epilogue that produced t)
    bool cross_ln_done = false; // same, for the cross-attention stage of the current block
    for (int i = 0; i < nblocks; ++i) {
      const std::string bp = pre + "blocks." + std::to_string(i) + ".";
      const bool has_cross = has(bp + "cross_attention.to_q.weight");
      // ---- self attention: fused QKV projection on LN statistics shared by norm / norm_context
      {
        const std::string ap = bp + "attention.";
        std::vector<float> w((size_t)3 * Hd * C), b((size_t)3 * Hd, 0.f);
        {
          std::vector<float> wq(T(ap + "to_q.weight", (int64_t)Hd * C), T(ap + "to_q.weight", (int64_t)Hd * C) + (size_t)Hd * C), bq(Hd, 0.f);
          fold_affine(wq, bq, Hd, C, T(ap + "norm.weight", C), T(ap + "norm.bias", C));
          std::vector<float> wkv(T(ap + "to_kv.weight", (int64_t)2 * Hd * C), T(ap + "to_kv.weight", (int64_t)2 * Hd * C) + (size_t)2 * Hd * C), bkv(2 * Hd, 0.f);
          fold_affine(wkv, bkv, 2 * Hd, C, T(ap + "norm_context.weight", C), T(ap + "norm_context.bias", C));
          std::copy(wq.begin(), wq.end(), w.begin()); std::copy(wkv.begin(), wkv.end(), w.begin() + (size_t)Hd * C);
          std::copy(bq.begin(), bq.end(), b.begin()); std::copy(bkv.begin(), bkv.end(), b.begin() + Hd);
        }
        const float* d_w = upload(w); const float* d_b = upload(b);
        const bool fuse_self = fast && fuse_attn && gemm_attn_supported(pl.prec, C, L, heads, d, 0, 0);
        cross_ln_done = false;
        const float* d_wr_self = nullptr; const float* d_bq_self = nullptr; bool layer_self = false;
        if (fuse_self) {
          // head-major repack: rows [h][q(64) | k(64) | v(64)]
          std::vector<float> wr((size_t)3 * Hd * C), br((size_t)3 * Hd);
          for (int h = 0; h < heads; ++h)
            for (int part = 0; part < 3; ++part)
              for (int r = 0; r < d; ++r) {
                const size_t src = (size_t)part * Hd + (size_t)h * d + r, dst = ((size_t)h * 3 + part) * d + r;
                memcpy(&wr[dst * C], &w[src * C], (size_t)C * sizeof(float));
                br[dst] = b[src];
              }
          d_wr_self = upload(wr);
          d_bq_self = upload(b.data(), Hd);      // q bias only (see gemm_attn.cu)
          if (!tn_is_ln) emit_ln_apply(prog, t, C, L, tn);
          layer_self = layer_ok(C, L, 0, false);
          if (!layer_self) {
            // wider levels: the per-(row block, head) variant of the TMEM-fragment kernel where it applies, else gemm_attn.cu
            if (frag_ok(C, L, 0, false)) emit_attn_layer(prog, tn, C, L, d_wr_self, d_bq_self, 0, -1, nullptr, nullptr, nullptr, nullptr, nullptr, t, nullptr, false);
            else emit_gemm_attn(prog, tn, C, L, d_wr_self, d_bq_self, 0, -1, nullptr, nullptr);
          }
        } else if (fast) {
          if (!tn_is_ln) emit_ln_apply(prog, t, C, L, tn);
          emit_gemm_tma(prog, tn, C, L, 1, d_w, d_b, 3 * Hd, 0, nullptr, nullptr, pl.qkv);
        } else {
          row_stats(prog, t, C, L, pl.row_stats);
          ALoad a = make_aload(ts, L, L, 1, 1, 0);
          a.stats = pl.row_stats; a.stats_mode = 1;
          emit(prog, gemm_op(a, d_w, tc_copy(d_w, w.size()), d_b, 3 * Hd, 0, nullptr, pl.qkv, L));
        }
        Op at; at.type = OP_ATTN; at.attn_kind = akind;
        at.at.q = pl.qkv; at.at.ldq = 3 * Hd;
        at.at.k = fast ? op_off(pl.qkv, Hd) : (void*)(pl.qkv + Hd);
        at.at.v = fast ? op_off(pl.qkv, 2 * Hd) : (void*)(pl.qkv + 2 * Hd);
        at.at.ldkv = 3 * Hd;
        at.at.kv_sample_stride = (long long)L * 3 * Hd; at.at.k_null = nullptr; at.at.v_null = nullptr; at.at.n_cond = 0;
        at.at.o = pl.att; at.at.ldo = Hd; at.at.nq = L; at.at.nk = L; at.at.heads = heads; at.at.d = d;
        at.at.scale = 1.0f / sqrtf((float)d);
        if (!fuse_self) emit(prog, at);
        const float* d_wo = upload(T(ap + "attention.to_out.weight", (int64_t)C * Hd), (size_t)C * Hd);
        std::vector<float> bo(T(ap + "attention.to_out.bias", C), T(ap + "attention.to_out.bias", C) + C);
        if (fuse_self) {
          // v bias passes through softmax-weighted averaging unchanged: out += Wo b_v
          const float* wo = T(ap + "attention.to_out.weight", (int64_t)C * Hd);
          for (int n = 0; n < C; ++n) {
            double acc = 0.0;
            for (int k = 0; k < Hd; ++k) acc += (double)wo[(size_t)n * Hd + k] * (double)b[(size_t)2 * Hd + k];
            bo[n] = (float)((double)bo[n] + acc);
          }
        }
        const float* d_bo = upload(bo);
        if (layer_self) {
          // with a cross-attention stage next, the fragment kernel's epilogue writes LayerNorm(t) for its q projection
          cross_ln_done = has_cross && !prefix && heads >= 4 && frag_ok(C, L, 0) && !getenv("MDT_NO_LN_EPILOGUE") && !getenv("MDT_NO_ATTN_LN");
          emit_attn_layer(prog, tn, C, L, d_wr_self, d_bq_self, 0, -1, nullptr, nullptr, nullptr, d_wo, d_bo, t,
                          (!has_cross || cross_ln_done) ? (void*)tn : nullptr, true, cross_ln_done);
        } else if (fast) {
          // without a cross-attention stage the raw operand copy of the new token stream feeds FF1 directly; with one, the epilogue
          // writes LayerNorm(t) for its q projection (not at the end of the CFG prefix: the null-branch rows are replicated there)
          cross_ln_done = has_cross && !prefix && ln_epilogue_ok(C);
          emit_gemm_tma(prog, pl.att, Hd, L, 1, d_wo, d_bo, C, 0, t, t, (!has_cross || cross_ln_done) ? (void*)tn : nullptr, 0, cross_ln_done);
        } else {
          Src as{pl.att, Hd, nullptr, 0, 1.f};
          emit(prog, gemm_op(make_aload(as, L, L, 1, 1, 0), d_wo, tc_copy(d_wo, (size_t)C * Hd), d_bo, C, 0, t, t, L));
        }
      }
      // ---- cross attention onto the conditioning embedding (K/V are loop invariant: precomputed per sample)
      if (has_cross) {
        if (!with_ctx) raise(MDT_ERR_INVALID, "unexpected cross_attention under '%s'", bp.c_str());
        end_prefix(prog, t, L, C);   // the branches diverge here (different K/V)
        const std::string ap = bp + "cross_attention.";
        std::vector<float> wq(T(ap + "to_q.weight", (int64_t)Hd * C), T(ap + "to_q.weight", (int64_t)Hd * C) + (size_t)Hd * C), bq(Hd, 0.f);
        fold_affine(wq, bq, Hd, C, T(ap + "norm.weight", C), T(ap + "norm.bias", C));
        const float* d_wq = upload(wq); const float* d_bq = upload(bq);
        std::vector<float> wkv(T(ap + "to_kv.weight", (int64_t)2 * Hd * F), T(ap + "to_kv.weight", (int64_t)2 * Hd * F) + (size_t)2 * Hd * F), bkv(2 * Hd, 0.f);
        fold_affine(wkv, bkv, 2 * Hd, F, T(ap + "norm_context.weight", F), T(ap + "norm_context.bias", F));
        CrossLayer cl{};
        cl.wkv = upload(wkv); cl.bkv = upload(bkv);
        cl.kv_cond = dalloc((size_t)pl.Bmax * pl.cfg.ctx_max_length * 2 * Hd);
        cl.kv_null = dalloc((size_t)pl.cfg.ctx_max_length * 2 * Hd);
        const bool fuse_cross = fast && fuse_attn && gemm_attn_supported(pl.prec, C, L, heads, d, 1, pl.cfg.ctx_max_length);
        // short contexts (<= 16 rows): K / V are read in mma.sync B-fragment order straight from a fragment-ordered tf32 cache and
        // 16 / L samples share one m16 tile (gemm_attn.cu mode 3); otherwise the kernels read row-major K / V in the operand dtype
        const char* np = getenv("MDT_NO_PACKED_CROSS");
        const char* ml = getenv("MDT_PACKED_CROSS_MAXL");
        const int maxl = ml ? atoi(ml) : 16;
        const bool packed = fuse_cross && (L == 4 || L == 8 || L == 16) && L <= maxl && pl.cfg.ctx_max_length <= 16 && d == 64 &&
                            !(np && np[0] != '0');
        if (packed) {
          cl.kvf_cond = dalloc((size_t)pl.Bmax * heads * 2048);
          cl.kvf_null = dalloc((size_t)heads * 2048);
        } else if (fast) {
          cl.kv_cond_op = dalloc((size_t)pl.Bmax * pl.cfg.ctx_max_length * 2 * Hd);   // sized in floats; bf16 uses half
          cl.kv_null_op = dalloc((size_t)pl.cfg.ctx_max_length * 2 * Hd);
        }
        const int layer = (int)pl.cross.size();
        pl.cross.push_back(cl);
        const bool layer_cross = fuse_cross && layer_ok(C, L, 1, packed);
        if (layer_cross) {
          if (!cross_ln_done) emit_ln_apply(prog, t, C, L, tn);
        } else if (fuse_cross && packed && frag_ok(C, L, 1, false)) {
          if (!cross_ln_done) emit_ln_apply(prog, t, C, L, tn);
          emit_attn_layer(prog, tn, C, L, d_wq, d_bq, 1, layer, cl.kv_null, cl.kvf_cond, cl.kvf_null, nullptr, nullptr, t, nullptr, false);
        } else if (fuse_cross) {
          if (!cross_ln_done) emit_ln_apply(prog, t, C, L, tn);
          // row-major variant: the fused kernel streams fp32 K/V with cp.async in both tensor-core modes (tf32: rounded copy, bf16: the
          // fp32 cache); the null-branch pointer doubles as the "has a null branch" flag, so it is passed in the packed variant too
          const bool op_copy = pl.prec == MDT_PREC_TF32 && !packed;
          emit_gemm_attn(prog, tn, C, L, d_wq, d_bq, 1, layer, op_copy ? cl.kv_cond_op : (void*)cl.kv_cond,
                         op_copy ? cl.kv_null_op : (void*)cl.kv_null, cl.kvf_cond, cl.kvf_null);
        } else if (fast) {
          if (!cross_ln_done) emit_ln_apply(prog, t, C, L, tn);
          emit_gemm_tma(prog, tn, C, L, 1, d_wq, d_bq, Hd, 0, nullptr, nullptr, pl.qc);
        } else {
          row_stats(prog, t, C, L, pl.row_stats);
          ALoad a = make_aload(ts, L, L, 1, 1, 0);
          a.stats = pl.row_stats; a.stats_mode = 1;
          emit(prog, gemm_op(a, d_wq, tc_copy(d_wq, wq.size()), d_bq, Hd, 0, nullptr, pl.qc, L));
        }
        Op at; at.type = OP_ATTN; at.cross = true; at.cross_layer = layer; at.attn_kind = akind;
        at.at.q = pl.qc; at.at.ldq = Hd; at.at.ldkv = 2 * Hd;
        if (fast) {
          at.at.k = cl.kv_cond_op; at.at.v = op_off(cl.kv_cond_op, Hd);
          at.at.k_null = cl.kv_null_op; at.at.v_null = op_off(cl.kv_null_op, Hd);
        } else {
          at.at.k = cl.kv_cond; at.at.v = cl.kv_cond + Hd; at.at.k_null = cl.kv_null; at.at.v_null = cl.kv_null + Hd;
        }
        at.at.o = pl.att; at.at.ldo = Hd; at.at.nq = L; at.at.heads = heads; at.at.d = d;
        at.at.scale = 1.0f / sqrtf((float)d);
        if (!fuse_cross) emit(prog, at);
        const float* d_wo = upload(T(ap + "attention.to_out.weight", (int64_t)C * Hd), (size_t)C * Hd);
        const float* d_bo = upload(T(ap + "attention.to_out.bias", C), C);
        if (layer_cross) {
          emit_attn_layer(prog, tn, C, L, d_wq, d_bq, 1, layer, cl.kv_null, cl.kvf_cond, cl.kvf_null, d_wo, d_bo, t, tn);
        } else if (fast) {
          emit_gemm_tma(prog, pl.att, Hd, L, 1, d_wo, d_bo, C, 0, t, t, tn);
        } else {
          Src as{pl.att, Hd, nullptr, 0, 1.f};
          emit(prog, gemm_op(make_aload(as, L, L, 1, 1, 0), d_wo, tc_copy(d_wo, (size_t)C * Hd), d_bo, C, 0, t, t, L));
        }
      }
      // ---- feed forward (no norm): Linear -> GELU -> Linear, + residual
      tn_is_ln = false;
      {
        const float* d_w0 = upload(T(bp + "feed_forward.0.weight", (int64_t)mid * C), (size_t)mid * C);
        const float* d_b0 = upload(T(bp + "feed_forward.0.bias", mid), mid);
        const float* d_w2 = upload(T(bp + "feed_forward.2.weight", (int64_t)C * mid), (size_t)C * mid);
        const float* d_b2 = upload(T(bp + "feed_forward.2.bias", C), C);
        // One chained kernel (gemm_chain.cu): the hidden activation goes through CTA-private L2 scratch instead of an HBM round
        // trip, and its epilogue hands the next block's self-attention its LayerNorm-ed operand (no separate ln_apply pass).
        // measured (profiles/README.md, B = 4096): level 1 (C = 128) 108 us vs 134 us for the two GEMMs; level 2 (C = 256: 256 row
        // blocks for 148 SMs, 1 MB of weights re-streamed per block) 91 vs 86 us, so wider layers keep the two-GEMM path by default
        const char* fmc = getenv("MDT_FF_CHAIN_MAXC");
        const bool chain = fast && pl.ff_scratch && (size_t)mid <= pl.ff_scratch_mid && !getenv("MDT_NO_FF_CHAIN") &&
                           C <= (fmc ? atoi(fmc) : 128) && ff_chain_supported(pl.prec, C, mid, L);
        if (chain) {
          const bool last = i == nblocks - 1;
          const bool ln_next = !last && !getenv("MDT_NO_CHAIN_LN");
          emit_ff_chain(prog, tn, C, mid, L, d_w0, d_b0, d_w2, d_b2, t, (last || ln_next) ? (void*)tn : nullptr, ln_next ? 1 : 0);
          tn_is_ln = ln_next;
        } else if (fast) {
          emit_gemm_tma(prog, tn, C, L, 1, d_w0, d_b0, mid, 1, nullptr, nullptr, pl.ff);
          // the last block's output is consumed by to_out as a raw operand: write the copy here
          // (other blocks: LayerNorm(t) for the next block's self-attention, from the same epilogue)
          const bool last = i == nblocks - 1;
          const bool ln_next = !last && ln_epilogue_ok(C);
          emit_gemm_tma(prog, pl.ff, mid, L, 1, d_w2, d_b2, C, 0, t, t, (last || ln_next) ? (void*)tn : nullptr, 0, ln_next);
          tn_is_ln = ln_next;
        } else {
          emit(prog, gemm_op(make_aload(ts, L, L, 1, 1, 0), d_w0, tc_copy(d_w0, (size_t)mid * C), d_b0, mid, 1, nullptr, pl.ff, L));
          Src fs{pl.ff, mid, nullptr, 0, 1.f};
          emit(prog, gemm_op(make_aload(fs, L, L, 1, 1, 0), d_w2, tc_copy(d_w2, (size_t)C * mid), d_b2, C, 0, t, t, L));
        }
      }
    }
    const float* d_wo = upload(T(pre + "to_out.1.weight", (int64_t)C * C), (size_t)C * C);
    const float* d_bo = upload(T(pre + "to_out.1.bias", C), C);
    float* out = acquire();
    if (fast && nblocks > 0) {
      void* oc = nullptr;
      if (out_op) { oc = acquire(); *out_op = oc; }
      emit_gemm_tma(prog, tn, C, L, 1, d_wo, d_bo, C, 0, nullptr, out, oc);
    } else {
      emit(prog, gemm_op(make_aload(ts, L, L, 1, 1, 0), d_wo, tc_copy(d_wo, (size_t)C * C), d_bo, C, 0, nullptr, out, L));
    }
    release(t);
    if (tn) release(tn);
    return out;
  }

  void build() {
    const mdt_config& c = pl.cfg;
    std::vector<Op>& prog = pl.unet;
    const int nlev = c.num_levels, p = c.patch_size;
    if (c.length % p) raise(MDT_ERR_INVALID, "length %d not divisible by patch_size %d", c.length, p);
    std::vector<int> Cl(nlev + 1), Ll(nlev + 1);
    Cl[0] = c.channels * c.multipliers[0]; Ll[0] = c.length / p;
    for (int i = 0; i < nlev; ++i) {
      if (Ll[i] % c.factors[i]) raise(MDT_ERR_INVALID, "level %d length %d not divisible by factor %d", i, Ll[i], c.factors[i]);
      if (c.factors[i] % 2) raise(MDT_ERR_INVALID, "odd resampling factor %d is not supported", c.factors[i]);
      Cl[i + 1] = c.channels * c.multipliers[i + 1]; Ll[i + 1] = Ll[i] / c.factors[i];
    }
    const int Cin0 = Cl[0] / p;  // channels of the to_in / to_out resnets (before patching)

    // ---- workspace sizing (floats per row-sample)
    size_t S = (size_t)c.length * std::max(Cin0, std::max(c.in_channels, c.out_channels));
    size_t Sqkv = 4, Satt = 4, Sff = 4, Sup = 4; int Lmax = c.length;
    for (int i = 0; i <= nlev; ++i) {
      S = std::max(S, (size_t)Ll[i] * Cl[i] * (i > 0 ? 2 : 1));   // up-block resnets see cat(x, skip): 2 * C channels
      if (i > 0) {
        Sqkv = std::max(Sqkv, (size_t)Ll[i] * 3 * pl.Hd);
        Satt = std::max(Satt, (size_t)Ll[i] * pl.Hd);
        Sff = std::max(Sff, (size_t)Ll[i] * Cl[i] * c.ff_multiplier);
        Sup = std::max(Sup, (size_t)Ll[i] * 2 * c.factors[i - 1] * Cl[i - 1]);
      }
    }
    pl.S_act = S;
    const size_t Be = pl.Beff_max;
    pl.qkv = dalloc(Sqkv * Be); pl.att = dalloc(Satt * Be); pl.qc = dalloc(Satt * Be); pl.ff = dalloc(Sff * Be);
    pl.upy = dalloc(Sup * Be);
    if (pl.prec != MDT_PREC_FP32) {
      size_t midmax = 0;
      for (int i = 1; i <= nlev; ++i) midmax = std::max(midmax, (size_t)Cl[i] * c.ff_multiplier);
      if (midmax >= 256 && midmax <= 2048) {
        const size_t bytes = ff_chain_scratch_bytes(pl.prec, (int)midmax);
        pl.ff_scratch = dalloc((bytes + 3) / 4);
        pl.ff_scratch_mid = midmax;
        CK(cudaMemset(pl.ff_scratch, 0, bytes));
      }
    }
    pl.gn_stats = dalloc((size_t)2 * 64 * Be);
    pl.row_stats = dalloc((size_t)2 * Lmax * Be);
    pl.xin = dalloc((size_t)c.length * c.in_channels * Be);
    pl.net_out = dalloc((size_t)c.length * c.out_channels * Be);
    pl.x = dalloc((size_t)c.length * c.in_channels * pl.Bmax);
    pl.xmid = dalloc((size_t)c.length * c.in_channels * pl.Bmax);
    pl.daux = dalloc((size_t)c.length * c.in_channels * pl.Bmax);
    const bool has_ctx = pl.F > 0;   // XUNet1d(type='base'): no conditioning embedding at all
    if (has_ctx) {
      pl.emb = dalloc((size_t)pl.Bmax * c.ctx_max_length * pl.F);
      pl.emb_stats = dalloc((size_t)pl.Bmax * c.ctx_max_length * 2);
      pl.emb_null_stats = dalloc((size_t)c.ctx_max_length * 2);
    }
    if (c.resnet_groups > 32) raise(MDT_ERR_INVALID, "resnet_groups > 32 unsupported");

    // ---- conditioning encoder + time path weights
    if (has_ctx) {
    pl.w_fc1 = upload(T("fc1.weight", c.text_embed_dim), c.text_embed_dim);
    pl.b_fc1 = upload(T("fc1.bias", c.text_embed_dim), c.text_embed_dim);
    if (c.pos_emb_fourier) {
      const int half = (c.embed_dim_position + 1) / 2;
      pl.inv_freq = upload(T("p_enc_1d.inv_freq", half), half);
      if (c.embed_dim_position % 2) raise(MDT_ERR_INVALID, "odd embed_dim_position unsupported");
      if (c.pos_emb_fourier_add && c.embed_dim_position != c.text_embed_dim)
        raise(MDT_ERR_INVALID, "pos_emb_fourier_add needs embed_dim_position == text_embed_dim");
      if (!c.pos_emb_fourier_add && c.embed_dim_position > c.text_embed_dim)
        raise(MDT_ERR_INVALID, "embed_dim_position > text_embed_dim truncates the encoding (transformer.py:3470)");
    }
    const int expectF = c.text_embed_dim + ((c.pos_emb_fourier && !c.pos_emb_fourier_add) ? c.embed_dim_position : 0);
    if (expectF != pl.F) raise(MDT_ERR_INVALID, "ctx_features %d != encoder width %d", pl.F, expectF);
    pl.w_null_emb = upload(T("unet.fixed_embedding.embedding.weight", (int64_t)c.ctx_max_length * pl.F),
                           (size_t)c.ctx_max_length * pl.F);
    }
    const int Mf = c.mapping_features, half = c.channels / 2;
    pl.w_time_freq = upload(T("unet.to_time.0.0.weights", half), half);
    pl.w_time = upload(T("unet.to_time.0.1.weight", (int64_t)Mf * (c.channels + 1)), (size_t)Mf * (c.channels + 1));
    pl.b_time = upload(T("unet.to_time.0.1.bias", Mf), Mf);
    pl.w_map0 = upload(T("unet.to_mapping.0.weight", (int64_t)Mf * Mf), (size_t)Mf * Mf);
    pl.b_map0 = upload(T("unet.to_mapping.0.bias", Mf), Mf);
    pl.w_map2 = upload(T("unet.to_mapping.2.weight", (int64_t)Mf * Mf), (size_t)Mf * Mf);
    pl.b_map2 = upload(T("unet.to_mapping.2.bias", Mf), Mf);
    pl.t_calls = dalloc(pl.max_calls);
    pl.t_feat = dalloc((size_t)pl.max_calls * (c.channels + 1));
    pl.t_a = dalloc((size_t)pl.max_calls * Mf);
    pl.t_b = dalloc((size_t)pl.max_calls * Mf);
    pl.t_map = dalloc((size_t)pl.max_calls * Mf);

    // ---- UNet program (UNet1d.forward, modules.py:1144-1180)
    const std::string U = "unet.";
    Src xin{pl.xin, c.in_channels, nullptr, 0, 1.f};
    void* cur_op = nullptr;   // operand-dtype copy of the current activation (feeds the TMA down-sampling conv)
    float* cur = resnet(prog, U + "to_in.block.", xin, c.length, Cin0, 1, nullptr, p == 1 ? &cur_op : nullptr);
    if (p > 1) {
      float* pt = acquire();
      Op op; op.type = OP_PERMUTE; op.in0 = cur; op.out = pt; op.i0 = Ll[0]; op.i1 = Cin0; op.i2 = p; op.i3 = 1;
      emit(prog, op);
      release(cur); cur = pt;
    }
    set_tap(prog, "to_in", cur, Ll[0], Cl[0]);
    const float* skip0 = cur;  // held until the end
    if (prefix) prefix_keep.push_back({cur, Ll[0], Cl[0]});
    std::vector<std::vector<const float*>> skips(nlev);
    const float* xcur = cur;
    bool xcur_owned = false;   // skip0 must not be released
    for (int i = 0; i < nlev; ++i) {
      const std::string dp = U + "downsamples." + std::to_string(i) + ".";
      const int f = c.factors[i], k = f * c.kernel_multiplier_downsample + 1, pad = f * (c.kernel_multiplier_downsample / 2);
      const int Ci = Cl[i], Co = Cl[i + 1], Lo = Ll[i + 1];
      const float* wraw = T(dp + "downsample.weight", (int64_t)Co * Ci * k);
      const float* d_bd = upload(T(dp + "downsample.bias", Co), Co);
      float* y = acquire();
      if (cur_op && k == 2 * f + 1 && pad == f && tma_ok(f * Ci, Lo, Co)) {
        // strided conv as a 3-tap stride-1 conv over f packed positions: [B][L][C] viewed as [B][L/f][f*C];
        // input f*o - f + kk = packed position o - 1 + kk / f, sub-position kk % f (the last tap uses sub-position 0 only)
        const int Cp = f * Ci;
        std::vector<float> wp((size_t)Co * 3 * Cp, 0.f);
        for (int co = 0; co < Co; ++co)
          for (int ci = 0; ci < Ci; ++ci)
            for (int kk = 0; kk < k; ++kk)
              wp[((size_t)co * 3 + kk / f) * Cp + (size_t)(kk % f) * Ci + ci] = wraw[((size_t)co * Ci + ci) * k + kk];
        const float* d_wp = upload(wp);
        emit_gemm_tma(prog, cur_op, Cp, Lo, 3, d_wp, d_bd, Co, 0, nullptr, y, nullptr);
      } else {
        auto wd = pack_conv(wraw, Co, Ci, k);
        const float* d_wd = upload(wd);
        Src xs{xcur, Ci, nullptr, 0, 1.f};
        emit(prog, gemm_op(make_aload(xs, Ll[i], Lo, k, f, pad), d_wd, tc_copy(d_wd, wd.size()), d_bd, Co, 0, nullptr, y, Lo));
      }
      if (cur_op) { release(reinterpret_cast<float*>(cur_op)); cur_op = nullptr; }
      set_tap(prog, "down" + std::to_string(i) + ".downsample", y, Lo, Co);
      if (xcur_owned) release(xcur);
      xcur = y; xcur_owned = true;
      if (has(dp + "pre_transformer_block.to_in.0.weight")) {
        float* t = transformer(prog, dp + "pre_transformer_block.", xcur, Lo, Co, false);
        set_tap(prog, "down" + std::to_string(i) + ".pre", t, Lo, Co);
        release(xcur); xcur = t;  // the pre-transformer skip is never consumed: not stored
      }
      for (int j = 0; j < c.num_blocks[i]; ++j) {
        Src s{xcur, Co, nullptr, 0, 1.f};
        float* r = resnet(prog, dp + "blocks." + std::to_string(j) + ".", s, Lo, Co, c.resnet_groups);
        set_tap(prog, "down" + std::to_string(i) + ".res" + std::to_string(j), r, Lo, Co);
        // previous xcur: release unless it is a stored skip
        if (j == 0) release(xcur);
        skips[i].push_back(r);
        if (prefix) prefix_keep.push_back({r, Lo, Co});
        xcur = r;
      }
      const bool more_levels = (i + 1 < nlev);
      if (c.attentions[i] > 0) {
        float* t = transformer(prog, dp + "transformer.", xcur, Lo, Co, true, more_levels ? &cur_op : nullptr);
        set_tap(prog, "down" + std::to_string(i) + ".tr", t, Lo, Co);
        skips[i].push_back(t);
        xcur = t;
      }
      xcur_owned = false;  // xcur is a stored skip now
      if (prefix) end_prefix(prog, const_cast<float*>(xcur), Lo, Co);
      if (c.num_blocks[i] == 0 && c.attentions[i] == 0) raise(MDT_ERR_INVALID, "level without blocks unsupported");
    }
    {
      const std::string bp = U + "bottleneck.";
      const int Cb = Cl[nlev], Lb = Ll[nlev];
      Src s{xcur, Cb, nullptr, 0, 1.f};
      float* r = resnet(prog, bp + "pre_block.", s, Lb, Cb, c.resnet_groups);
      set_tap(prog, "mid.pre", r, Lb, Cb);
      xcur = r;
      if (c.attentions[nlev] > 0 && has(bp + "transformer.to_in.0.weight")) {
        float* t = transformer(prog, bp + "transformer.", xcur, Lb, Cb, true);
        set_tap(prog, "mid.tr", t, Lb, Cb);
        release(xcur); xcur = t;
      }
      Src s2{xcur, Cb, nullptr, 0, 1.f};
      float* r2 = resnet(prog, bp + "post_block.", s2, Lb, Cb, c.resnet_groups);
      set_tap(prog, "mid.post", r2, Lb, Cb);
      release(xcur); xcur = r2;
    }
    const float skip_scale = c.use_skip_scale ? (float)pow(2.0, -0.5) : 1.0f;
    for (int u = 0; u < nlev; ++u) {
      const int i = nlev - 1 - u;
      const std::string up = U + "upsamples." + std::to_string(u) + ".";
      const int Ci = Cl[i + 1], Co = Cl[i], Li = Ll[i + 1], f = c.factors[i];
      const int nres = c.num_blocks[i] + (c.attentions[i] ? 1 : 0);
      if ((int)skips[i].size() != nres) raise(MDT_ERR_INVALID, "skip bookkeeping mismatch at level %d", i);
      for (int j = 0; j < nres; ++j) {
        const float* sk = skips[i].back(); skips[i].pop_back();
        Src s{xcur, Ci, sk, Ci, skip_scale};
        float* r = resnet(prog, up + "blocks." + std::to_string(j) + ".", s, Li, Ci, c.resnet_groups);
        set_tap(prog, "up" + std::to_string(u) + ".res" + std::to_string(j), r, Li, Ci);
        release(xcur); release(sk);
        xcur = r;
      }
      if (has(up + "pre_transformer_block.to_in.0.weight")) {
        float* t = transformer(prog, up + "pre_transformer_block.", xcur, Li, Ci, false);
        set_tap(prog, "up" + std::to_string(u) + ".pre", t, Li, Ci);
        release(xcur); xcur = t;
      }
      void* xop = nullptr;
      if (c.attentions[i] > 0) {
        float* t = transformer(prog, up + "transformer.", xcur, Li, Ci, true, &xop);
        set_tap(prog, "up" + std::to_string(u) + ".tr", t, Li, Ci);
        release(xcur); xcur = t;
      }
      // ConvTranspose1d(Ci -> Co, k = 2f, stride f, pad f/2): GEMM to [Li][2f*Co] then two-tap gather
      const int K2 = 2 * f;
      const float* wt = T(up + "upsample.weight", (int64_t)Ci * Co * K2);
      std::vector<float> wp((size_t)K2 * Co * Ci);
      for (int ci = 0; ci < Ci; ++ci)
        for (int co = 0; co < Co; ++co)
          for (int kk = 0; kk < K2; ++kk) wp[((size_t)kk * Co + co) * Ci + ci] = wt[((size_t)ci * Co + co) * K2 + kk];
      const float* d_wp = upload(wp);
      const float* d_bu = upload(T(up + "upsample.bias", Co), Co);
      if (xop && tma_ok(Ci, Li, K2 * Co)) {
        emit_gemm_tma(prog, xop, Ci, Li, 1, d_wp, nullptr, K2 * Co, 0, nullptr, pl.upy, nullptr);
      } else {
        Src s{xcur, Ci, nullptr, 0, 1.f};
        emit(prog, gemm_op(make_aload(s, Li, Li, 1, 1, 0), d_wp, tc_copy(d_wp, wp.size()), nullptr, K2 * Co, 0, nullptr, pl.upy, Li));
      }
      if (xop) release(reinterpret_cast<float*>(xop));
      float* y = acquire();
      Op op; op.type = OP_UPGATHER; op.in0 = pl.upy; op.in1 = d_bu; op.in2 = (i == 0) ? skip0 : nullptr; op.out = y;
      op.i0 = Li; op.i1 = Co; op.i2 = f;
      emit(prog, op);
      set_tap(prog, "up" + std::to_string(u) + ".upsample", y, Li * f, Co);
      release(xcur); xcur = y;
    }
    release(skip0);
    if (p > 1) {
      float* pt = acquire();
      Op op; op.type = OP_PERMUTE; op.in0 = xcur; op.out = pt; op.i0 = Ll[0]; op.i1 = Cin0; op.i2 = p; op.i3 = 0;
      emit(prog, op);
      release(xcur); xcur = pt;
    }
    {
      Src s{xcur, Cin0, nullptr, 0, 1.f};
      resnet(prog, U + "to_out.block.", s, c.length, c.out_channels, 1, pl.net_out);   // the program ends in pl.net_out
      set_tap(prog, "to_out", pl.net_out, c.length, c.out_channels);
      release(xcur);
    }
    CK(cudaDeviceSynchronize());
  }
};

// ------------------------------------------------------------------------------------------------
// execution
// ------------------------------------------------------------------------------------------------
static void launch_gemm(mdt_plan& pl, const GemmParams& g, cudaStream_t s) {
  if (g.M <= 0) return;
  if (pl.prec != MDT_PREC_FP32 && g.Wtc && gemm_tc_supported(g)) CK(launch_gemm_tc(g, pl.prec, s));
  else CK(launch_gemm_fp32(g, s));
  pl.launches++;
}

// MDT_OP_TIMES=1 (diagnostics; forces eager launches): per-op device time by shape, warm and back to back, printed when the
// plan is destroyed.  ncu's per-launch times are cold-cache and serialised; this is the in-situ complement.
struct OpTime { double ms = 0; long long n = 0; };
static std::map<std::string, OpTime> g_op_times;
static void dump_op_times();
static bool op_times_on() {
  static const bool on = [] { const bool v = getenv("MDT_OP_TIMES") != nullptr; if (v) atexit(dump_op_times); return v; }();
  return on;
}

static std::string describe(const Op& op, int Beff) {
  char b[160];
  const char* h = op.half ? " half" : "";
  switch (op.type) {
    case OP_GEMM_TMA: snprintf(b, sizeof b, "gemm_tma  M=%d N=%d K=%dx%d L=%d bn=%d%s%s%s%s", Beff * op.rps, op.tg.N, op.tg.taps, op.tg.C, op.tg.L, op.tg.BN,
                               op.tg.gn_L ? " +gn" : "", op.tg.res ? " +res" : "", op.tg.act ? " +act" : (op.tg.cop_ln ? " +ln" : ""), h); break;
    case OP_GEMM_ATTN: snprintf(b, sizeof b, "gemm_attn%s %s M=%d C=%d L=%d%s", op.umma_core ? "_umma" : "", op.cross ? "cross" : "self", Beff * op.rps, op.gat.C, op.gat.L, h); break;
    case OP_ATTN_LAYER: snprintf(b, sizeof b, "attn_%s%s %s M=%d C=%d L=%d%s%s", op.frag ? "frag " : "layer", op.al.fused ? "" : "(unfused)", op.cross ? "cross" : "self", Beff * op.rps, op.al.a.C, op.al.a.L, op.al.cop_ln ? " +ln" : (op.al.Cop ? " +cop" : ""), h); break;
    case OP_RESNET_SMALL: snprintf(b, sizeof b, "resnet_small %s B=%d L=%d %d->%d%s", op.rs.mode ? "head" : "full", Beff, op.rs.L, op.rs.Cin, op.rs.Cout, h); break;
    case OP_GEMM: snprintf(b, sizeof b, "gemm      M=%d N=%d K=%d taps=%d stride=%d%s", Beff * op.rps, op.g.N, op.g.K, op.g.a.taps, op.g.a.stride, h); break;
    case OP_GN_APPLY: snprintf(b, sizeof b, "gn_apply  B=%d L=%d C=%d%s%s", Beff, op.ga.L, op.ga.c0 + op.ga.c1, op.ga.raw ? " +raw" : "", h); break;
    case OP_LN_APPLY: snprintf(b, sizeof b, "ln_apply  rows=%d C=%d%s", Beff * op.rps, op.la.C, h); break;
    case OP_GEMM_FF: snprintf(b, sizeof b, "ff_chain  M=%d C=%d mid=%d%s%s", Beff * op.rps, op.ffc.C, op.ffc.mid, op.ffc.cop_ln ? " +ln" : (op.ffc.Cop ? " +cop" : ""), h); break;
    default: snprintf(b, sizeof b, "other(type %d)%s", (int)op.type, h); break;
  }
  return b;
}

static void dump_op_times() {
  if (g_op_times.empty()) return;
  double tot = 0;
  for (auto& kv : g_op_times) tot += kv.second.ms;
  std::vector<std::pair<std::string, OpTime>> v(g_op_times.begin(), g_op_times.end());
  std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.second.ms > b.second.ms; });
  fprintf(stderr, "[mdt] op times: %.3f ms total\n", tot);
  for (auto& kv : v)
    fprintf(stderr, "[mdt] %6.2f%% %9.3f ms n=%5lld avg=%8.1f us  %s\n", 100.0 * kv.second.ms / tot, kv.second.ms, kv.second.n,
            1e3 * kv.second.ms / (double)kv.second.n, kv.first.c_str());
  g_op_times.clear();
}

static void run_program(mdt_plan& pl, std::vector<Op>& prog, int Beff, int n_cond, int n_ctx, cudaStream_t s) {
  if (getenv("MDT_NO_SHARED_PREFIX")) for (Op& op : prog) op.half = false;
  const int Beff_full = Beff;
  // programmatic dependent launch only where every grid of the program leaves SMs free (launch.cuh): row blocks of the longest level
  {
    static int sms = 0;
    if (sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    pdl_auto() = ((long long)Beff_full * pl.L0 + 127) / 128 < sms ? 1 : 0;
  }
  const bool timed = op_times_on();
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (timed) { CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1)); }
  // serpentine traversal: each of the four streaming kernels walks the batch in the opposite direction to the previous one,
  // so a consumer starts on the rows its producer wrote last (still in L2) instead of the ones evicted first
  const bool serp = pl.serpentine;
  int rev = 0;
  for (Op& op : prog) {
    Beff = op.half ? n_cond : Beff_full;
    if (timed) CK(cudaEventRecord(ev0, s));
    const bool streams = op.type == OP_GEMM_TMA || op.type == OP_GEMM_ATTN || op.type == OP_GN_APPLY || op.type == OP_LN_APPLY ||
                         op.type == OP_ATTN_LAYER || op.type == OP_GEMM_FF;
    rev = (serp && streams) ? (rev ^ 1) : 0;
    switch (op.type) {
      case OP_DUP_ROWS: {
        if (Beff_full > n_cond) {
          const size_t n = (size_t)n_cond * op.i0 * op.i1;
          CK(cudaMemcpyAsync(op.out + n, op.out, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
          pl.launches++;
        }
        break;
      }
      case OP_GEMM: { GemmParams g = op.g; g.M = Beff * op.rps; launch_gemm(pl, g, s); break; }
      case OP_GN_STATS: { NormStatsParams n = op.ns; n.rows = Beff; CK(launch_groupnorm_stats(n, s)); pl.launches++; break; }
      case OP_ROW_STATS: { NormStatsParams n = op.ns; n.rows = Beff * op.rps; CK(launch_rownorm_stats(n, s)); pl.launches++; break; }
      case OP_ATTN: {
        AttnParams a = op.at; a.B = Beff;
        if (op.cross) { a.nk = n_ctx; a.kv_sample_stride = (long long)n_ctx * a.ldkv; a.n_cond = n_cond; }
        CK(launch_attention(a, op.attn_kind, s)); pl.launches++; break;
      }
      case OP_GN_APPLY: { GnApplyParams g = op.ga; g.B = Beff; g.rev = rev; CK(launch_gn_apply(g, pl.prec, s)); pl.launches++; break; }
      case OP_LN_APPLY: { LnApplyParams l = op.la; l.rows = (long long)Beff * op.rps; l.rev = rev; CK(launch_ln_apply(l, pl.prec, s)); pl.launches++; break; }
      case OP_GEMM_ATTN: {
        GemmAttnParams g = op.gat; g.M = Beff * op.rps; g.rev = rev;
        if (op.cross) { g.nk = n_ctx; g.kv_sample_stride = (long long)n_ctx * g.ldkv; g.n_cond = n_cond; }
        CK(op.umma_core ? launch_gemm_attn_umma(op.tmA, op.tmB, g, pl.prec, s) : launch_gemm_attn(op.tmA, op.tmB, g, pl.prec, s));
        pl.launches++; break;
      }
      case OP_ATTN_LAYER: {
        AttnLayerParams y = op.al; y.a.M = Beff * op.rps; y.a.rev = rev;
        if (op.cross) { y.a.nk = n_ctx; y.a.kv_sample_stride = (long long)n_ctx * y.a.ldkv; y.a.n_cond = n_cond; }
        CK(op.frag ? launch_attn_frag(op.tmA, op.tmB, op.tmC, op.tmD, y, pl.prec, s)
                   : launch_attn_layer(op.tmA, op.tmB, op.tmC, op.tmD, y, pl.prec, s));
        pl.launches++; break;
      }
      case OP_RESNET_SMALL: { ResnetSmallParams r = op.rs; r.B = Beff; CK(launch_resnet_small(r, s)); pl.launches++; break; }
      case OP_GEMM_FF: {
        FFChainParams g = op.ffc; g.M = Beff * op.rps; g.rev = rev;
        CK(launch_ff_chain(op.tmA, op.tmB, op.tmC, op.tmD, g, pl.prec, s)); pl.launches++; break;
      }
      case OP_GEMM_TMA: { TmaGemmParams g = op.tg; g.M = Beff * op.rps; g.rev = rev; CK(launch_gemm_tma(op.tmA, op.tmB, g, pl.prec, s)); pl.launches++; break; }
      case OP_UPGATHER: CK(launch_upsample_gather(op.in0, op.in1, op.in2, op.out, Beff, op.i0, op.i1, op.i2, s)); pl.launches++; break;
      case OP_PERMUTE: CK(launch_patch_permute(op.in0, op.out, Beff, op.i0, op.i1, op.i2, op.i3, s)); pl.launches++; break;
    }
    if (timed) {
      CK(cudaEventRecord(ev1, s));
      CK(cudaEventSynchronize(ev1));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, ev0, ev1));
      OpTime& t = g_op_times[describe(op, Beff)];
      t.ms += ms; t.n++;
    }
    if (pl.taps_on && !op.tap.empty()) {
      CK(cudaStreamSynchronize(s));
      std::vector<float>& dst = pl.tap_store[op.tap];
      dst.resize((size_t)Beff * op.tap_rps * op.tap_c);
      CK(cudaMemcpy(dst.data(), op.tap_ptr, dst.size() * sizeof(float), cudaMemcpyDeviceToHost));
    }
  }
  if (timed) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); }
}

static GemmParams dense(const float* A, int K, const float* W, const float* b, int N, int M, int act, float* C, bool silu_in) {
  GemmParams g{};
  g.a.src0 = A; g.a.src1 = nullptr; g.a.c0 = K; g.a.c1 = 0; g.a.C = K; g.a.scale1 = 1.f;
  g.a.L_in = 1; g.a.L_out = 1; g.a.taps = 1; g.a.stride = 1; g.a.pad = 0; g.a.stats = nullptr; g.a.stats_mode = 0;
  g.a.groups = 1; g.a.cpg = K; g.a.aff = nullptr; g.a.aff_call_stride = 0; g.a.call_idx = nullptr; g.a.silu = silu_in ? 1 : 0;
  g.W = W; g.Wtc = nullptr; g.bias = b; g.M = M; g.N = N; g.K = K; g.act = act; g.res = nullptr; g.ldres = N; g.C = C; g.ldc = N;
  return g;
}

// time -> mapping -> per-resnet FiLM tables for `rows` denoiser calls (UNet1d.get_mapping, modules.py:1123-1142;
// MappingToScaleShift, modules.py:138-142).  Always fp32: it is sigma-only, batch-invariant work.
static void run_time_tables(mdt_plan& pl, int rows, cudaStream_t s) {
  const mdt_config& c = pl.cfg;
  const int Mf = c.mapping_features;
  CK(launch_time_features(pl.t_calls, pl.w_time_freq, pl.t_feat, rows, c.channels / 2, s)); pl.launches++;
  CK(launch_gemm_fp32(dense(pl.t_feat, c.channels + 1, pl.w_time, pl.b_time, Mf, rows, 1, pl.t_a, false), s));
  CK(launch_gemm_fp32(dense(pl.t_a, Mf, pl.w_map0, pl.b_map0, Mf, rows, 1, pl.t_b, false), s));
  CK(launch_gemm_fp32(dense(pl.t_b, Mf, pl.w_map2, pl.b_map2, Mf, rows, 1, pl.t_map, false), s));
  pl.launches += 3;
  for (Film& f : pl.films) {
    CK(launch_gemm_fp32(dense(pl.t_map, Mf, f.w_ss, f.b_ss, 2 * f.C, rows, 0, f.ss, true), s));
    CK(launch_film_fold(f.ss, f.gamma, f.beta, f.aff, rows, f.C, s));
    pl.launches += 2;
  }
}

// conditioning embedding -> per-layer cross-attention K/V (loop invariant)
static void run_context(mdt_plan& pl, const float* cond_dev, int Bc, int n_ctx, bool cfg, cudaStream_t s) {
  const mdt_config& c = pl.cfg;
  if (pl.F == 0) return;   // unconditional UNet
  const int F = pl.F, Hd = pl.Hd;
  if (pl.ctx_pre_encoded)   // the caller hands over the embedding the wrapper would have computed (generative.py:838-850)
    CK(cudaMemcpyAsync(pl.emb, cond_dev, (size_t)Bc * n_ctx * F * sizeof(float), cudaMemcpyDeviceToDevice, s));
  else
    CK(launch_encode_cond(cond_dev, pl.w_fc1, pl.b_fc1, pl.inv_freq, pl.emb, Bc, n_ctx, c.text_embed_dim,
                          c.pos_emb_fourier ? c.embed_dim_position : 0, c.pos_emb_fourier_add, s));
  NormStatsParams n{}; n.src0 = pl.emb; n.c0 = F; n.scale1 = 1.f; n.L = 1; n.groups = 1; n.eps = 1e-5f; n.stats = pl.emb_stats; n.rows = Bc * n_ctx;
  CK(launch_rownorm_stats(n, s));
  pl.launches += 2;
  if (cfg) {
    n.src0 = pl.w_null_emb; n.stats = pl.emb_null_stats; n.rows = n_ctx;
    CK(launch_rownorm_stats(n, s)); pl.launches++;
  }
  for (CrossLayer& cl : pl.cross) {
    GemmParams g = dense(pl.emb, F, cl.wkv, cl.bkv, 2 * Hd, Bc * n_ctx, 0, cl.kv_cond, false);
    g.a.stats = pl.emb_stats; g.a.stats_mode = 1;
    CK(launch_gemm_fp32(g, s)); pl.launches++;
    if (cl.kv_cond_op) { CK(convert_weights_tc(cl.kv_cond, cl.kv_cond_op, (long long)Bc * n_ctx * 2 * Hd, pl.prec, s)); pl.launches++; }
    if (cl.kvf_cond) { CK(launch_kv_fragment_pack(cl.kv_cond, cl.kvf_cond, Bc, n_ctx, c.heads, c.head_features, cl.kperm, s)); pl.launches++; }
    if (cfg) {
      GemmParams gn = dense(pl.w_null_emb, F, cl.wkv, cl.bkv, 2 * Hd, n_ctx, 0, cl.kv_null, false);
      gn.a.stats = pl.emb_null_stats; gn.a.stats_mode = 1;
      CK(launch_gemm_fp32(gn, s)); pl.launches++;
      if (cl.kv_null_op) { CK(convert_weights_tc(cl.kv_null, cl.kv_null_op, (long long)n_ctx * 2 * Hd, pl.prec, s)); pl.launches++; }
      if (cl.kvf_null) { CK(launch_kv_fragment_pack(cl.kv_null, cl.kvf_null, 1, n_ctx, c.heads, c.head_features, cl.kperm, s)); pl.launches++; }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* mdt_last_error(void) { return g_err; }
int mdt_abi_version(void) { return MDT_ABI_VERSION; }

int mdt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++;
  }
  return ok;
}

int mdt_karras_sigmas(int num_steps, double sigma_min, double sigma_max, double rho, float* out) {
  if (num_steps < 2 || !out) return fail(MDT_ERR_INVALID, "num_steps must be >= 2");
  const double rho_inv = 1.0 / rho;
  const float a = (float)pow(sigma_max, rho_inv);
  const float d = (float)(pow(sigma_min, rho_inv) - pow(sigma_max, rho_inv));
  for (int i = 0; i < num_steps; ++i) {
    const float frac = (float)i / (float)(num_steps - 1);
    const float base = a + frac * d;
    out[i] = powf(base, (float)rho);
  }
  out[num_steps] = 0.f;
  return 0;
}

static void scale_weights(float s, double sigma_data, float* c_in, float* c_noise, float* c_skip, float* c_out) {
  const float sd2 = (float)(sigma_data * sigma_data);
  const float s2 = s * s;
  *c_noise = logf(s) * 0.25f;
  *c_skip = sd2 / (s2 + sd2);
  *c_out = (s * (float)sigma_data) * (1.0f / sqrtf(sd2 + s2));
  *c_in = 1.0f / sqrtf(s2 + sd2);
}

int mdt_adpm2_scalars(const float* sigmas, int n_iters, double rho, double sigma_data, mdt_iter_scalars* out) {
  if (!sigmas || !out || n_iters < 0) return fail(MDT_ERR_INVALID, "bad arguments");
  for (int i = 0; i < n_iters; ++i) {
    const float s = sigmas[i], sn = sigmas[i + 1];
    const float sn2 = sn * sn, s2 = s * s;
    const float q = sn2 * (s2 - sn2) / s2;
    const double up = sqrt((double)q);
    const float dn_in = sn2 - (float)(up * up);
    const double down = sqrt((double)dn_in);
    float mid;
    if (rho == 1.0) mid = (s + (float)down) / 2.0f;
    else mid = powf((powf(s, (float)(1.0 / rho)) + (float)pow(down, 1.0 / rho)) / 2.0f, (float)rho);
    mdt_iter_scalars& o = out[i];
    o.sigma = s; o.sigma_mid = mid;
    scale_weights(s, sigma_data, &o.c_in_a, &o.c_noise_a, &o.c_skip_a, &o.c_out_a);
    scale_weights(mid, sigma_data, &o.c_in_b, &o.c_noise_b, &o.c_skip_b, &o.c_out_b);
    o.dt_mid = mid - s;
    o.dt_down = (float)down - s;
    o.sigma_up = (float)up;
  }
  return 0;
}

int mdt_aeuler_scalars(const float* sigmas, int n_iters, double sigma_data, mdt_iter_scalars* out) {
  if (!sigmas || !out || n_iters < 0) return fail(MDT_ERR_INVALID, "bad arguments");
  for (int i = 0; i < n_iters; ++i) {
    const float s = sigmas[i], sn = sigmas[i + 1];
    const float sn2 = sn * sn, s2 = s * s;
    const double up = sqrt((double)(sn2 * (s2 - sn2) / s2));
    const double down = sqrt((double)(sn2 - (float)(up * up)));
    mdt_iter_scalars& o = out[i];
    o.sigma = s; o.sigma_mid = s;      // midpoint at the start: mdt_plan_sample runs one denoiser call per step for such rows
    scale_weights(s, sigma_data, &o.c_in_a, &o.c_noise_a, &o.c_skip_a, &o.c_out_a);
    o.c_in_b = o.c_in_a; o.c_noise_b = o.c_noise_a; o.c_skip_b = o.c_skip_a; o.c_out_b = o.c_out_a;
    o.dt_mid = 0.0f;
    o.dt_down = (float)down - s;
    o.sigma_up = (float)up;
  }
  return 0;
}

int mdt_plan_create(const mdt_config* cfg, const mdt_tensor* tensors, int64_t n_tensors, int device, mdt_plan** out) {
  if (!cfg || !tensors || !out) return fail(MDT_ERR_INVALID, "null argument");
  if (cfg->abi_version != MDT_ABI_VERSION) return fail(MDT_ERR_INVALID, "ABI version mismatch (%d != %d)", cfg->abi_version, MDT_ABI_VERSION);
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(MDT_ERR_NO_DEVICE, "no CUDA device visible: this library has no CPU fallback"); }
  if (device < 0 || device >= ndev) return fail(MDT_ERR_INVALID, "device %d out of range", device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(MDT_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(MDT_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  if (cfg->num_levels < 1 || cfg->num_levels > MDT_MAX_LEVELS) return fail(MDT_ERR_INVALID, "num_levels out of range");
  if (cfg->max_batch < 1) return fail(MDT_ERR_INVALID, "max_batch must be >= 1");
  if (cfg->precision < 0 || cfg->precision > 3) return fail(MDT_ERR_INVALID, "unknown precision %d", cfg->precision);
  mdt_plan* pl = new mdt_plan();
  int prev_device = -1;
  cudaGetDevice(&prev_device);   // creation must not change the caller's current device
  try {
    CK(cudaSetDevice(device));
    CK(init_kernels());
    CK(init_gemm_tc());
    CK(init_gemm_tma());
    CK(init_gemm_attn());
    CK(init_gemm_attn_umma());
    CK(init_attn_layer());
    CK(init_attn_frag());
    CK(init_resnet_small());
    CK(init_ff_chain());
    pl->cfg = *cfg; pl->device = device; pl->prec = cfg->precision;
    pl->P = cfg->in_channels; pl->L0 = cfg->length; pl->Hd = cfg->heads * cfg->head_features; pl->F = cfg->ctx_features;
    pl->Bmax = cfg->max_batch; pl->Beff_max = 2 * cfg->max_batch;
    const int max_steps = cfg->max_timesteps > 1 ? cfg->max_timesteps : 256;
    pl->max_calls = 2 * (max_steps - 1);
    const char* eg = getenv("MDT_GRAPH");
    pl->use_graph = !(eg && eg[0] == '0') && !op_times_on();
    const char* es = getenv("MDT_SERPENTINE");
    pl->serpentine = !(es && es[0] == '0');
    size_t total = 0;
    for (int64_t i = 0; i < n_tensors; ++i) {
      if (!tensors[i].name || !tensors[i].data) raise(MDT_ERR_INVALID, "tensor %lld has a null field", (long long)i);
      pl->tensors[tensors[i].name] = {tensors[i].data, tensors[i].numel};
      total += (size_t)tensors[i].numel;
    }
    pl->wcap = total * sizeof(float) * (pl->prec == MDT_PREC_FP32 ? 2 : 3) + (64u << 20);
    { void* p = nullptr; cudaError_t e = cudaMalloc(&p, pl->wcap); if (e != cudaSuccess) raise(MDT_ERR_OOM, "weight slab cudaMalloc(%zu) failed: %s", pl->wcap, cudaGetErrorString(e)); pl->wslab = (char*)p; }
    CK(cudaMalloc((void**)&pl->d_iters, sizeof(IterScalars) * (pl->max_calls / 2 + 1)));
    CK(cudaMallocHost((void**)&pl->h_iters, sizeof(IterScalars) * (pl->max_calls / 2 + 1)));
    CK(cudaMallocHost((void**)&pl->h_tcalls, sizeof(float) * (pl->max_calls + 2)));
    CK(cudaMalloc((void**)&pl->d_call, sizeof(int)));
    CK(cudaMemset(pl->d_call, 0, sizeof(int)));
    CK(cudaMalloc((void**)&pl->d_run, sizeof(RunParams)));
    CK(cudaEventCreateWithFlags(&pl->staged, cudaEventDisableTiming));
    Builder b(*pl);
    b.build();
    pl->tensors.clear();  // host pointers are not retained past creation
  } catch (const MdtError& e) {
    int code = e.code;
    snprintf(g_err, sizeof(g_err), "%s", e.msg.c_str());
    mdt_plan_destroy(pl);
    if (prev_device >= 0) cudaSetDevice(prev_device);
    return code;
  }
  if (prev_device >= 0) cudaSetDevice(prev_device);
  *out = pl;
  return 0;
}

void mdt_plan_destroy(mdt_plan* pl) {
  if (!pl) return;
  int prev_device = -1;
  cudaGetDevice(&prev_device);
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_device};
  cudaSetDevice(pl->device);
  cudaDeviceSynchronize();
  dump_op_times();
  for (auto& kv : pl->graphs) cudaGraphExecDestroy(kv.second.exec);
  for (void* p : pl->allocs) cudaFree(p);
  if (pl->wslab) cudaFree(pl->wslab);
  if (pl->d_iters) cudaFree(pl->d_iters);
  if (pl->h_iters) cudaFreeHost(pl->h_iters);
  if (pl->h_tcalls) cudaFreeHost(pl->h_tcalls);
  if (pl->d_call) cudaFree(pl->d_call);
  if (pl->d_run) cudaFree(pl->d_run);
  if (pl->staged) cudaEventDestroy(pl->staged);
  cudaGetLastError();
  delete pl;
}

int64_t mdt_plan_device_bytes(const mdt_plan* pl) { return pl ? (int64_t)(pl->wcap + pl->act_bytes) : 0; }
int64_t mdt_plan_launch_count(const mdt_plan* pl) { return pl ? pl->launches : 0; }

int mdt_plan_set_sampler_mode(mdt_plan* pl, int mode, float init_noise_scale, float init_sigma) {
  if (!pl) return fail(MDT_ERR_INVALID, "null plan");
  if (mode != 0 && mode != 1) return fail(MDT_ERR_INVALID, "unknown sampler mode %d", mode);
  pl->sampler_mode = mode;
  pl->init_noise_scale = init_noise_scale;
  pl->init_sigma = init_sigma;
  return 0;
}

int mdt_plan_set_context_mode(mdt_plan* pl, int pre_encoded) {
  if (!pl) return fail(MDT_ERR_INVALID, "null plan");
  pl->ctx_pre_encoded = pre_encoded != 0;
  return 0;
}

int mdt_plan_enable_taps(mdt_plan* pl, int enable) {
  if (!pl) return fail(MDT_ERR_INVALID, "null plan");
  pl->taps_on = enable != 0;
  pl->tap_store.clear();
  return 0;
}

int64_t mdt_plan_read_tap(mdt_plan* pl, const char* name, float* host_dst, int64_t capacity) {
  if (!pl || !name) return fail(MDT_ERR_INVALID, "null argument");
  auto it = pl->tap_store.find(name);
  if (it == pl->tap_store.end()) return fail(MDT_ERR_INVALID, "tap '%s' not recorded", name);
  const int64_t n = (int64_t)it->second.size();
  if (host_dst) {
    if (capacity < n) return fail(MDT_ERR_INVALID, "tap '%s' needs %lld floats", name, (long long)n);
    memcpy(host_dst, it->second.data(), n * sizeof(float));
  }
  return n;
}

static int check_ctx(mdt_plan* pl, int n_ctx) {
  if (pl->F == 0) return 0;   // XUNet1d(type='base'): the conditioning is ignored (generative.py:862-868)
  if (n_ctx < 1 || n_ctx > pl->cfg.ctx_max_length)
    return fail(MDT_ERR_INVALID, "Input sequence length must be <= max_length (%d > %d)", n_ctx, pl->cfg.ctx_max_length);
  return 0;
}

int mdt_plan_unet_forward(mdt_plan* pl, const float* x_dev, float time, const float* cond_dev, int32_t n_ctx,
                          int64_t B, float cond_scale, float* out_dev, void* stream) {
  if (!pl || !x_dev || !cond_dev || !out_dev) return fail(MDT_ERR_INVALID, "null argument");
  if (B < 1 || B > pl->Bmax) return fail(MDT_ERR_INVALID, "B=%lld exceeds plan max_batch=%d", (long long)B, pl->Bmax);
  if (check_ctx(pl, n_ctx)) return MDT_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  try {
    CK(cudaSetDevice(pl->device));
    const bool cfg = cond_scale != 1.0f && pl->F > 0;   // classifier-free guidance needs a conditioning embedding
    const int Bc = (int)B, Beff = cfg ? 2 * Bc : Bc;
    CK(cudaEventSynchronize(pl->staged));
    pl->h_tcalls[0] = time;
    CK(cudaMemcpyAsync(pl->t_calls, pl->h_tcalls, sizeof(float), cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(pl->staged, s));
    CK(launch_set_int(pl->d_call, 0, s));
    run_time_tables(*pl, 1, s);
    run_context(*pl, cond_dev, Bc, n_ctx, cfg, s);
    CK(launch_to_token_major(x_dev, pl->xin, Bc, pl->P, pl->L0, 1.0f, cfg ? 1 : 0, s));
    run_program(*pl, pl->unet, Beff, Bc, n_ctx, s);
    CK(launch_cfg_mix_to_bpl(pl->net_out, out_dev, Bc, pl->cfg.out_channels, pl->L0, cond_scale, cfg ? 1 : 0, s));
    pl->launches += 3;
  } catch (const MdtError& e) {
    snprintf(g_err, sizeof(g_err), "%s", e.msg.c_str());
    return e.code;
  }
  return 0;
}

// single_call: every row of the scalar table has its midpoint at the start (sigma_mid == sigma, dt_mid == 0), i.e. a first-order
// ancestral Euler step (AEulerSampler, diffusion.py:456-483).  Call A would evaluate the denoiser at (x, sigma) only to move by
// dt_mid = 0, so it is skipped: update B runs on (x, sigma) directly (x_mid aliases x) and the call counter still advances by two
// (the time tables keep one slot pair per iteration).
static void run_iteration(mdt_plan& pl, StepParams& sp, int Beff, int Bc, int n_ctx, bool single_call, cudaStream_t s) {
  if (single_call) {
    StepParams sb = sp;
    sb.xmid = sp.x;
    run_program(pl, pl.unet, Beff, Bc, n_ctx, s);
    CK(launch_step_update(1, sb, s));
    CK(launch_add_int(pl.d_call, 2, s));
    pl.launches += 2;
    return;
  }
  run_program(pl, pl.unet, Beff, Bc, n_ctx, s);
  CK(launch_step_update(0, sp, s));
  CK(launch_add_int(pl.d_call, 1, s));
  run_program(pl, pl.unet, Beff, Bc, n_ctx, s);
  CK(launch_step_update(1, sp, s));
  CK(launch_add_int(pl.d_call, 1, s));
  pl.launches += 4;
}

int mdt_plan_sample(mdt_plan* pl, const float* cond_dev, int32_t n_ctx, const float* noise0_dev,
                    const float* step_noise_dev, const mdt_iter_scalars* iters, int32_t n_iters, uint64_t seed,
                    uint64_t sample_offset, int64_t B, float cond_scale, int32_t clamp, float* out_dev,
                    uint8_t* tokens_dev, void* stream) {
  if (!pl || !cond_dev || !iters) return fail(MDT_ERR_INVALID, "null argument");
  if (!out_dev && !tokens_dev) return fail(MDT_ERR_INVALID, "one of out_dev / tokens_dev is required");
  if (B < 0) return fail(MDT_ERR_INVALID, "negative batch");
  if (n_iters < 1) return fail(MDT_ERR_INVALID, "timesteps must be >= 2");
  if (2 * n_iters > pl->max_calls) return fail(MDT_ERR_INVALID, "timesteps %d exceeds the plan's max_timesteps %d", n_iters + 1, pl->max_calls / 2 + 1);
  if (check_ctx(pl, n_ctx)) return MDT_ERR_INVALID;
  if (tokens_dev && pl->cfg.out_channels > 256) return fail(MDT_ERR_INVALID, "uint8 tokens need pred_dim <= 256");
  if (pl->cfg.in_channels != pl->cfg.out_channels) return fail(MDT_ERR_INVALID, "sampler needs in_channels == out_channels");
  if (B == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  static_assert(sizeof(IterScalars) == sizeof(mdt_iter_scalars), "scalar layout mismatch");
  try {
    CK(cudaSetDevice(pl->device));
    const bool cfg = cond_scale != 1.0f && pl->F > 0;   // classifier-free guidance needs a conditioning embedding
    const int P = pl->P, L = pl->L0;
    const size_t per = (size_t)P * L;
    // a previous call's async copies out of the pinned staging buffers must have drained (only those two copies: the
    // event sits right behind them, so a back-to-back caller does not wait for the previous sample() to finish)
    CK(cudaEventSynchronize(pl->staged));
    memcpy(pl->h_iters, iters, sizeof(IterScalars) * n_iters);
    for (int i = 0; i < n_iters; ++i) { pl->h_tcalls[2 * i] = iters[i].c_noise_a; pl->h_tcalls[2 * i + 1] = iters[i].c_noise_b; }
    bool single_call = true;   // first-order rows (see run_iteration)
    for (int i = 0; i < n_iters; ++i) single_call = single_call && iters[i].dt_mid == 0.0f && iters[i].sigma_mid == iters[i].sigma;
    const bool karras = pl->sampler_mode == 1;
    if (karras) single_call = false;
    CK(cudaMemcpyAsync(pl->d_iters, pl->h_iters, sizeof(IterScalars) * n_iters, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(pl->t_calls, pl->h_tcalls, sizeof(float) * 2 * n_iters, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(pl->staged, s));
    run_time_tables(*pl, 2 * n_iters, s);
    for (int64_t b0 = 0; b0 < B; b0 += pl->Bmax) {
      const int Bc = (int)std::min<int64_t>(pl->Bmax, B - b0);
      const int Beff = cfg ? 2 * Bc : Bc;
      run_context(*pl, cond_dev + (size_t)b0 * n_ctx * (pl->ctx_pre_encoded ? pl->F : 1), Bc, n_ctx, cfg, s);
      // KarrasSampler: x = sigmas[0] * noise while row 0 carries sigma_hat_0 (the x_hat / network input follow in karras_prenoise)
      CK(launch_step_init(noise0_dev ? noise0_dev + (size_t)b0 * per : nullptr, pl->x, pl->xin, pl->d_iters, seed,
                          sample_offset + (uint64_t)b0, Bc, P, L, cfg ? 1 : 0, karras ? pl->init_sigma : -1.0f, s));
      CK(launch_set_int(pl->d_call, 0, s));
      pl->launches += 2;
      if (karras) {   // x_hat of step 0: the first step noise goes in ahead of the first denoiser call (diffusion.py:425-426)
        CK(launch_karras_prenoise(pl->x, pl->xin, step_noise_dev ? step_noise_dev + (size_t)b0 * per : nullptr, pl->init_noise_scale,
                                  iters[0].c_in_a, seed, sample_offset + (uint64_t)b0, Bc, P, L, cfg ? 1 : 0, s));
        pl->launches++;
      }
      StepParams sp{};
      sp.iters = pl->d_iters; sp.call_idx = pl->d_call; sp.net = pl->net_out; sp.x = pl->x; sp.xmid = pl->xmid; sp.xin = pl->xin;
      sp.noise = step_noise_dev ? step_noise_dev + (size_t)b0 * per : nullptr;
      sp.noise_iter_stride = (long long)B * (long long)per;
      sp.seed = seed; sp.sample_offset = sample_offset + (uint64_t)b0; sp.cond_scale = cond_scale; sp.cfg = cfg ? 1 : 0;
      // seed / offset / injected-noise base / guidance scale live in device memory: one captured graph serves all of them
      sp.run = pl->d_run; sp.has_noise = sp.noise ? 1 : 0;
      RunParams rp{}; rp.seed = seed; rp.sample_offset = sample_offset + (uint64_t)b0; rp.noise = sp.noise;
      rp.noise_iter_stride = sp.noise_iter_stride; rp.cond_scale = cond_scale;
      CK(launch_set_run_params(pl->d_run, rp, s));
      sp.B = Bc; sp.P = P; sp.L = L; sp.n_iters = n_iters; sp.noise_stream = -1; sp.out = nullptr; sp.tokens = nullptr; sp.clamp = clamp;
      sp.karras = karras ? 1 : 0; sp.daux = pl->daux;
      if (pl->use_graph && !pl->taps_on) {
        // one captured iteration, replayed n_iters times; all per-iteration data is device resident
        std::vector<long long> key = {Bc, n_ctx, cfg ? 1 : 0, sp.noise ? 1 : 0, (long long)n_iters, single_call ? 1 : 0, karras ? 1 : 0};
        auto it = pl->graphs.find(key);
        if (it == pl->graphs.end()) {
          if (pl->graphs.size() > 16) { for (auto& kv : pl->graphs) cudaGraphExecDestroy(kv.second.exec); pl->graphs.clear(); }
          cudaStream_t cs;
          CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
          cudaGraph_t graph = nullptr;
          const long long before = pl->launches;
          CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
          try { run_iteration(*pl, sp, Beff, Bc, n_ctx, single_call, cs); }
          catch (...) { cudaStreamEndCapture(cs, &graph); if (graph) cudaGraphDestroy(graph); cudaStreamDestroy(cs); throw; }
          CK(cudaStreamEndCapture(cs, &graph));
          const long long captured = pl->launches - before;   // kernels (and copies) recorded into one iteration
          pl->launches = before;
          cudaGraphExec_t exec = nullptr;
          CK(cudaGraphInstantiate(&exec, graph, 0));
          CK(cudaGraphDestroy(graph));
          CK(cudaStreamDestroy(cs));
          it = pl->graphs.emplace(key, mdt_plan::GraphEntry{exec, captured}).first;
        }
        for (int i = 0; i < n_iters; ++i) CK(cudaGraphLaunch(it->second.exec, s));
        pl->launches += (long long)n_iters * it->second.launches;
      } else {
        for (int i = 0; i < n_iters; ++i) run_iteration(*pl, sp, Beff, Bc, n_ctx, single_call, s);
      }
      CK(launch_finalize(pl->x, out_dev ? out_dev + (size_t)b0 * per : nullptr,
                         tokens_dev ? tokens_dev + (size_t)b0 * L : nullptr, Bc, P, L, clamp, s));
      pl->launches++;
    }
  } catch (const MdtError& e) {
    snprintf(g_err, sizeof(g_err), "%s", e.msg.c_str());
    return e.code;
  }
  return 0;
}

int mdt_plan_inpaint(mdt_plan* pl, const float* cond_dev, int32_t n_ctx, const float* source_dev, const uint8_t* mask_dev,
                     const float* noise_dev, const mdt_iter_scalars* iters, const float* sigmas, int32_t n_iters,
                     int32_t num_resamples, uint64_t seed, uint64_t sample_offset, int64_t B, float cond_scale, float* out_dev,
                     void* stream) {
  if (!pl || !cond_dev || !source_dev || !mask_dev || !iters || !sigmas || !out_dev) return fail(MDT_ERR_INVALID, "null argument");
  if (B < 0 || n_iters < 1 || num_resamples < 1) return fail(MDT_ERR_INVALID, "bad batch / timesteps / num_resamples");
  if (2 * n_iters > pl->max_calls) return fail(MDT_ERR_INVALID, "timesteps %d exceeds the plan's max_timesteps %d", n_iters + 1, pl->max_calls / 2 + 1);
  if (check_ctx(pl, n_ctx)) return MDT_ERR_INVALID;
  if (pl->cfg.in_channels != pl->cfg.out_channels) return fail(MDT_ERR_INVALID, "sampler needs in_channels == out_channels");
  if (B == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  try {
    CK(cudaSetDevice(pl->device));
    const bool cfg = cond_scale != 1.0f && pl->F > 0;   // classifier-free guidance needs a conditioning embedding
    const int P = pl->P, L = pl->L0, R = num_resamples;
    const size_t per = (size_t)P * L;
    const long long draws_per_iter = 2LL * R;              // source noise + R step noises + (R - 1) re-noises
    CK(cudaEventSynchronize(pl->staged));
    memcpy(pl->h_iters, iters, sizeof(IterScalars) * n_iters);
    for (int i = 0; i < n_iters; ++i) { pl->h_tcalls[2 * i] = iters[i].c_noise_a; pl->h_tcalls[2 * i + 1] = iters[i].c_noise_b; }
    CK(cudaMemcpyAsync(pl->d_iters, pl->h_iters, sizeof(IterScalars) * n_iters, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(pl->t_calls, pl->h_tcalls, sizeof(float) * 2 * n_iters, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(pl->staged, s));
    run_time_tables(*pl, 2 * n_iters, s);
    for (int64_t b0 = 0; b0 < B; b0 += pl->Bmax) {
      const int Bc = (int)std::min<int64_t>(pl->Bmax, B - b0);
      const int Beff = cfg ? 2 * Bc : Bc;
      run_context(*pl, cond_dev + (size_t)b0 * n_ctx * (pl->ctx_pre_encoded ? pl->F : 1), Bc, n_ctx, cfg, s);
      const float* src = source_dev + (size_t)b0 * per;
      const uint8_t* msk = mask_dev + (size_t)b0 * per;
      // draw d of the reference's RNG sequence lives at noise_dev[d][B][P][L]
      auto nz = [&](long long d) -> const float* { return noise_dev ? noise_dev + ((size_t)d * B + (size_t)b0) * per : nullptr; };
      const uint64_t off = sample_offset + (uint64_t)b0;
      CK(launch_inpaint(0, pl->x, pl->xin, src, msk, nz(0), iters[0].sigma, 0.f, seed, off, 0, Bc, P, L, cfg, nullptr, s));
      StepParams sp{};
      sp.iters = pl->d_iters; sp.call_idx = pl->d_call; sp.net = pl->net_out; sp.x = pl->x; sp.xmid = pl->xmid; sp.xin = pl->xin;
      sp.noise_iter_stride = 0; sp.seed = seed; sp.sample_offset = off; sp.cond_scale = cond_scale; sp.cfg = cfg ? 1 : 0;
      sp.B = Bc; sp.P = P; sp.L = L; sp.n_iters = n_iters; sp.out = nullptr; sp.tokens = nullptr; sp.clamp = 0;
      for (int i = 0; i < n_iters; ++i) {
        long long d = 1 + (long long)i * draws_per_iter;
        const long long d_src = d++;
        for (int r = 0; r < R; ++r) {
          CK(launch_inpaint(1, pl->x, pl->xin, src, msk, nz(d_src), iters[i].sigma, iters[i].c_in_a, seed, off, (int)d_src, Bc, P, L,
                            cfg, nullptr, s));
          CK(launch_set_int(pl->d_call, 2 * i, s));
          const long long d_step = d++;
          sp.noise = nz(d_step); sp.noise_stream = (int)d_step;
          run_iteration(*pl, sp, Beff, Bc, n_ctx, false, s);
          pl->launches += 2;
          if (r < R - 1) {
            const long long d_re = d++;
            const float sg = (float)sqrt((double)(sigmas[i] * sigmas[i] - sigmas[i + 1] * sigmas[i + 1]));
            CK(launch_inpaint(2, pl->x, pl->xin, src, msk, nz(d_re), sg, 0.f, seed, off, (int)d_re, Bc, P, L, cfg, nullptr, s));
            pl->launches++;
          }
        }
      }
      CK(launch_inpaint(3, pl->x, pl->xin, src, msk, nullptr, 0.f, 0.f, seed, off, 0, Bc, P, L, cfg, out_dev + (size_t)b0 * per, s));
      pl->launches += 2;
    }
  } catch (const MdtError& e) {
    snprintf(g_err, sizeof(g_err), "%s", e.msg.c_str());
    return e.code;
  }
  return 0;
}

int mdt_op_linear(const float* a_dev, const float* w_dev, const float* bias_dev, const float* res_dev, float* c_dev,
                  int64_t M, int32_t N, int32_t K, int32_t act, int32_t precision, void* stream) {
  if (!a_dev || !w_dev || !c_dev) return fail(MDT_ERR_INVALID, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (init_kernels() != cudaSuccess || init_gemm_tc() != cudaSuccess) return fail(MDT_ERR_CUDA, "kernel attribute setup failed: %s", cudaGetErrorString(cudaGetLastError()));
  GemmParams g = dense(a_dev, K, w_dev, bias_dev, N, (int)M, act, c_dev, false);
  g.res = res_dev;
  cudaError_t e;
  if (precision == MDT_PREC_FP32) e = launch_gemm_fp32(g, s);
  else {
    if (!gemm_tc_supported(g)) return fail(MDT_ERR_INVALID, "shape M=%lld N=%d K=%d unsupported by the tcgen05 kernel", (long long)M, N, K);
    void* wtc = nullptr;
    const size_t n = (size_t)N * K;
    if (cudaMalloc(&wtc, n * 4) != cudaSuccess) return fail(MDT_ERR_OOM, "cudaMalloc failed");
    e = convert_weights_tc(w_dev, wtc, (long long)n, precision, s);
    g.Wtc = wtc;
    if (e == cudaSuccess) e = launch_gemm_tc(g, precision, s);
    cudaStreamSynchronize(s);
    cudaFree(wtc);
  }
  if (e != cudaSuccess) return fail(MDT_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return 0;
}

int mdt_op_step_update(int which, const float* net_dev, float* x_dev, float* xmid_dev, float* xin_dev,
                       const float* noise_dev, const mdt_iter_scalars* it, float cond_scale, int64_t B, int32_t P,
                       int32_t L, int cfg, void* stream) {
  if (!net_dev || !x_dev || !xmid_dev || !xin_dev || !it) return fail(MDT_ERR_INVALID, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  static IterScalars* d_it = nullptr;
  static int* d_call = nullptr;
  if (!d_it) {
    if (cudaMalloc((void**)&d_it, 2 * sizeof(IterScalars)) != cudaSuccess || cudaMalloc((void**)&d_call, sizeof(int)) != cudaSuccess)
      return fail(MDT_ERR_OOM, "cudaMalloc failed");
    cudaMemset(d_call, 0, sizeof(int));
  }
  IterScalars two[2];
  memcpy(&two[0], it, sizeof(IterScalars)); memcpy(&two[1], it, sizeof(IterScalars));
  cudaMemcpyAsync(d_it, two, sizeof(two), cudaMemcpyHostToDevice, s);
  cudaStreamSynchronize(s);
  StepParams sp{};
  sp.iters = d_it; sp.call_idx = d_call; sp.net = net_dev; sp.x = x_dev; sp.xmid = xmid_dev; sp.xin = xin_dev;
  sp.noise = noise_dev; sp.noise_iter_stride = 0; sp.seed = 0; sp.sample_offset = 0; sp.cond_scale = cond_scale; sp.cfg = cfg;
  sp.B = (int)B; sp.P = P; sp.L = L; sp.n_iters = 2; sp.noise_stream = -1; sp.clamp = 0;
  cudaError_t e = launch_step_update(which, sp, s);
  if (e != cudaSuccess) return fail(MDT_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return 0;
}

int mdt_op_decode_tokens(const uint8_t* tokens_dev, const uint8_t* lut_dev, uint8_t* out_dev, int32_t* lengths_dev, int64_t B,
                         int32_t L, void* stream) {
  if (B < 0 || L < 0) return fail(MDT_ERR_INVALID, "negative size");
  if (B == 0 || L == 0) return 0;
  if (!tokens_dev || !lut_dev || !out_dev) return fail(MDT_ERR_INVALID, "null argument");
  cudaError_t e = launch_decode_tokens(tokens_dev, lut_dev, out_dev, lengths_dev, (long long)B, L, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(MDT_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
