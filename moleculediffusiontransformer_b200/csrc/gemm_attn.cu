// gemm_attn.cu -- fused projection + attention for sm_100a: the per-head [q | k | v] projection runs on tcgen05
// with the accumulator in TMEM and the attention core runs in the epilogue, so the (rows x 3 * heads * d) q/k/v
// tensor never goes to HBM.
//
//   self  : tile = (128 rows, head h): D[128, 192] = LN(x)[128, C] * Wh[192, C]^T  (Wh = [Wq_h; Wk_h; Wv_h], LayerNorm
//           affine folded), epilogue adds the folded bias, stages q / k / v per sample in shared memory and runs
//           softmax(q k^T * scale) v per sample with mma.sync, writing only the (rows x d) head output.
//   cross : D[128, 64] = LN(x) * Wq_h^T; K / V of the conditioning come from the loop-invariant per-sample cache.
//
// Persistent CTA per SM, same three pipelines as gemm_tma.cu (smem ring fed by TMA, two TMEM accumulators, tile loop
// with the head index fastest so the eight heads of one row block share the activation tile in L2).
#include <cuda.h>
#include <cuda_bf16.h>
#include "attn_math.cuh"

namespace mdt {
namespace tc {

constexpr int A_TM = 128;
constexpr int A_STAGES = 2;           // operand ring depth of the modes that need the shared memory for q / k / v staging or K / V scratch
constexpr int A_STAGES_PACKED = 4;    // mode 3 stages only q: the ring can hold half a level-2 tile
__host__ __device__ constexpr int attn_stages(int mode) { return mode == 3 ? A_STAGES_PACKED : A_STAGES; }
constexpr int A_ABYTES = A_TM * 128;
constexpr int A_EPI_WARPS = 8;
constexpr int A_THREADS = 64 + 32 * A_EPI_WARPS;
#ifndef MDT_ATTN_SKIP_MATH
#define MDT_ATTN_SKIP_MATH 0
#endif
#ifndef MDT_ATTN_TRUNC
#define MDT_ATTN_TRUNC 0
#endif
constexpr int A_LD = 68;   // staged q/k/v row stride in floats (64 + 4: conflict-free fragment loads)

// MODE selects the attention core compiled into the epilogue (one path per instantiation keeps each within the register budget):
//   0 self, one sample per warp pass      1 self, 16 / L short samples packed into one block-diagonal m16 tile (L <= 8)
//   2 cross, K / V staged with cp.async   3 cross, K / V fragments straight from the fragment-ordered cache (n_ctx <= 16)
template <int KIND, int MODE>
__global__ void __launch_bounds__(A_THREADS, 1) gemm_attn_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const GemmAttnParams p, const uint32_t idesc) {
  constexpr int KCH = (KIND == 1) ? 32 : 64;
  constexpr bool CROSS = MODE >= 2;
  constexpr int NST = attn_stages(MODE);
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // SWIZZLE_128B operand tiles need 1024-byte alignment
  __shared__ __align__(8) uint64_t full_bar[NST];
  __shared__ __align__(8) uint64_t empty_bar[NST];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  // keep the pointer in the shared address space (an integer round trip would demote every access to generic LD/ST)
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = CROSS ? p.d : 3 * p.d;
  const int b_bytes = BN * 128;
  const int stage_bytes = A_ABYTES + ((b_bytes + 1023) & ~1023);
  const int m_tiles = (p.M + A_TM - 1) / A_TM;
  const int total_tiles = m_tiles * p.heads;
  const uint32_t tmem_cols = CROSS ? 128u : 512u;   // two accumulators of BN columns
  float* Qs = reinterpret_cast<float*>(smem + NST * stage_bytes);
  float* Ks = Qs + A_TM * A_LD;                       // self: staged k rows; cross: per-warp K scratch base
  float* Vs = Ks + A_TM * A_LD;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], A_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) tmem_alloc(&tmem_base_s, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_enter();                                  // the set-up above overlaps the previous grid's tail (launch.cuh)
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t tx = (uint32_t)(A_ABYTES + b_bytes);
      int c = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int te = p.rev ? total_tiles - 1 - t : t;
        const int mt = te / p.heads, h = te - mt * p.heads;
        for (int kc = 0; kc < p.kchunks; ++kc, ++c) {
          const int stage = c % NST;
          const uint32_t phase = (uint32_t)(c / NST) & 1u;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          tma_load_3d(sa, &tmA, &full_bar[stage], kc * KCH, 0, mt * p.Sb);
          tma_load_2d(sa + A_ABYTES, &tmB, &full_bar[stage], kc * KCH, h * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int c = 0, it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&acc_empty[buf], aphase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
      for (int k0 = 0; k0 < p.kchunks; ++k0, ++c) {
        const int stage = c % NST;
        const uint32_t phase = (uint32_t)(c / NST) & 1u;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + A_ABYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma<KIND>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((k0 | k) != 0));
          umma_commit(&empty_bar[stage]);
          if (k0 == p.kchunks - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue + attention (warps 2..9)
    const int ew = warp - 2;
    const int q = warp & 3;                 // TMEM lane quadrant
    const int half = ew >> 2;
    const int row = q * 32 + lane;          // tile row owned for the TMEM -> smem transfer
    const int cols_per_warp = BN / 2;       // 96 (self) or 32 (cross)
    const int L = p.L;
    // cross: warp-private K / V scratch [nk][A_LD] each, carved from the Ks region onwards
    float* Kw = Ks + (size_t)ew * 4 * p.nk * A_LD;   // two [K | V] buffers per warp (prefetch one sample ahead)
    float* Vw = Kw + (size_t)p.nk * A_LD;
    const int kvbuf = 2 * p.nk * A_LD;
    // sample ownership: with L <= 32 a quadrant (32 rows) holds 32 / L whole samples, shared by its two warps
    const bool quad_local = L <= 32;
    const int spq = quad_local ? 32 / L : 0;                       // samples per quadrant
    const int s_begin = quad_local ? q * spq : 0, s_end = quad_local ? (q + 1) * spq : p.Sb;
    const int s_lane = quad_local ? half : ew, s_step = quad_local ? 2 : A_EPI_WARPS;
    int pf_buf = 0;
    bool pf_primed = false;
    // cp.async copy of the K / V rows of (tile tt, sample ss) for this tile's head into scratch buffer `bufi`
    auto issue_kv = [&](int tt, int ss, int bufi) {
      if (p.rev) tt = total_tiles - 1 - tt;
      const int mtt = tt / p.heads, hh = tt - mtt * p.heads;
      const int b = (mtt * A_TM + ss * L) / L;
      const bool nul = p.kn && b >= p.n_cond;
      const float* kb = reinterpret_cast<const float*>(nul ? p.kn : p.kc) + (nul ? 0 : (size_t)b * p.kv_sample_stride) + (size_t)hh * p.d;
      float* kd = Kw + bufi * kvbuf;
      float* vd = Vw + bufi * kvbuf;
      for (int idx = lane; idx < p.nk * 16; idx += 32) {
        const int j = idx >> 4, c4 = (idx & 15) * 4;
        const float* src = kb + (size_t)j * p.ldkv + c4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(kd + j * A_LD + c4)), "l"(src) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(vd + j * A_LD + c4)), "l"(src + p.heads * p.d) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int te = p.rev ? total_tiles - 1 - t : t;
      const int mt = te / p.heads, h = te - mt * p.heads;
      const int m0 = mt * A_TM;
      const int buf = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&acc_full[buf], aphase);
      tc_fence_after();
      for (int cc = 0; cc < cols_per_warp; cc += 32) {
        uint32_t v[32];
        const int col0 = half * cols_per_warp + cc;
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + col0), v);
        if (m0 + row >= p.M) {
          // rows past the batch are in bounds for the TMA box and hold stale data; the packed tiles multiply them by P = 0
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        float* dst = (col0 < 64 ? Qs : (col0 < 128 ? Ks : Vs)) + (size_t)row * A_LD + (col0 & 63);
        if (col0 < 64) {
          // only q carries a bias here: the k bias shifts every score of a query equally (cancels in the softmax) and
          // the v bias passes through the row-stochastic attention matrix unchanged (folded into the out-projection bias)
          const float* bias = p.bias + h * p.d + col0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + j * 4));
            *reinterpret_cast<uint4*>(dst + j * 4) =
                make_uint4(to_tf32(__uint_as_float(v[4 * j]) + bv.x), to_tf32(__uint_as_float(v[4 * j + 1]) + bv.y),
                           to_tf32(__uint_as_float(v[4 * j + 2]) + bv.z), to_tf32(__uint_as_float(v[4 * j + 3]) + bv.w));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (MDT_ATTN_TRUNC)   // experiment: let mma.sync truncate k / v to tf32 instead of rounding to nearest here
              *reinterpret_cast<uint4*>(dst + j * 4) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            else
              *reinterpret_cast<uint4*>(dst + j * 4) =
                  make_uint4(to_tf32(__uint_as_float(v[4 * j])), to_tf32(__uint_as_float(v[4 * j + 1])),
                             to_tf32(__uint_as_float(v[4 * j + 2])), to_tf32(__uint_as_float(v[4 * j + 3])));
          }
        }
      }
      tc_fence_before();
      // rows 32q .. 32q + 31 (whole samples when L <= 32) are staged by the two warps of quadrant q only, so a
      // 64-thread named barrier per quadrant is enough; quadrants drift apart and overlap staging with attention
      if (quad_local) asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      else asm volatile("bar.sync 1, 256;" ::: "memory");
      if (lane == 0) mbar_arrive(&acc_empty[buf]);             // the accumulator may be overwritten now
      if constexpr (MODE == 3) {
        // packed path (L = 4, 8, 16): this warp's 16 rows = 16 / L whole samples in one m16 tile, K / V fragments straight from global
        const int r16 = q * 32 + half * 16;
        const int mrow = m0 + r16;
        if (mrow < p.M) {
          const int bs = mrow / L;
          const size_t ob = (size_t)mrow * p.ldo + (size_t)h * p.d;
          const int rows_valid = min(16, p.M - mrow);
          auto block_of = [&](int t) {
            const int b = (mrow + t * L < p.M) ? bs + t : bs;            // samples past the batch: any valid block, rows not stored
            const bool nul = p.kn && b >= p.n_cond;
            return reinterpret_cast<const uint2*>(nul ? p.kvf_n : p.kvf_c) + ((nul ? (size_t)0 : (size_t)b * p.heads) + h) * 1024;
          };
          if (L == 4) {
            const uint2* const kf[4] = {block_of(0), block_of(1), block_of(2), block_of(3)};
            attend_packed_cross<KIND, 4>(Qs + (size_t)r16 * A_LD, A_LD, kf, p.nk, p.scale, p.att, ob, p.ldo, rows_valid, lane);
          } else if (L == 8) {
            const uint2* const kf[2] = {block_of(0), block_of(1)};
            attend_packed_cross<KIND, 8>(Qs + (size_t)r16 * A_LD, A_LD, kf, p.nk, p.scale, p.att, ob, p.ldo, rows_valid, lane);
          } else {
            const uint2* const kf[1] = {block_of(0)};
            attend_packed_cross<KIND, 16>(Qs + (size_t)r16 * A_LD, A_LD, kf, p.nk, p.scale, p.att, ob, p.ldo, rows_valid, lane);
          }
        }
      } else if constexpr (MODE == 1) {
        // short samples (L <= 8): this warp's 16 staged rows = 16 / L samples as one block-diagonal 16 x 16 attention
        const int r16 = q * 32 + half * 16;
        const int mrow = m0 + r16;
        if (mrow < p.M)
          attend_head_mma_nt<1, KIND, 2>(Qs + (size_t)r16 * A_LD, A_LD, Ks + (size_t)r16 * A_LD, Vs + (size_t)r16 * A_LD, A_LD,
                                         min(16, p.M - mrow), 16, p.scale, p.att, (size_t)mrow * p.ldo + (size_t)h * p.d, p.ldo, lane, L);
      } else
      for (int s = s_begin + s_lane; s < s_end; s += s_step) {
        const int mrow = m0 + s * L;
        if (mrow >= p.M) break;
        const size_t ob = (size_t)mrow * p.ldo + (size_t)h * p.d;
        if (MDT_ATTN_SKIP_MATH) {
        } else if constexpr (!CROSS) {
          attend_head_mma<1, KIND>(Qs + (size_t)s * L * A_LD, A_LD, Ks + (size_t)s * L * A_LD, Vs + (size_t)s * L * A_LD, A_LD, L, L,
                                   p.scale, p.att, ob, p.ldo, lane);
        } else if (KIND == 1 || p.kv_fp32) {
          // fp32-storage K/V cache: the copy for this item was issued one item ago with cp.async; issue the next one now
          if (!pf_primed) { issue_kv(t, s, pf_buf); pf_primed = true; }
          int nt = t, ns = s + s_step;
          auto tile_m0 = [&](int tt) { return ((p.rev ? total_tiles - 1 - tt : tt) / p.heads) * A_TM; };
          if (ns >= s_end || tile_m0(nt) + ns * L >= p.M) { nt = t + gridDim.x; ns = s_begin + s_lane; }
          const bool has_next = nt < total_tiles && ns < s_end && tile_m0(nt) + ns * L < p.M;
          if (has_next) issue_kv(nt, ns, pf_buf ^ 1);
          if (has_next) asm volatile("cp.async.wait_group 1;" ::: "memory");
          else asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncwarp();
          attend_head_mma<1, KIND>(Qs + (size_t)s * L * A_LD, A_LD, Kw + pf_buf * kvbuf, Vw + pf_buf * kvbuf, A_LD, L, p.nk, p.scale,
                                   p.att, ob, p.ldo, lane);
          __syncwarp();
          pf_buf ^= 1;
        } else {
          const int b = mrow / L;
          const bool nul = p.kn && b >= p.n_cond;
          const size_t koff = nul ? 0 : (size_t)b * p.kv_sample_stride;
          const void* kb = nul ? p.kn : p.kc;
          __syncwarp();
          for (int idx = lane; idx < p.nk * 16; idx += 32) {
            const int j = idx >> 4, c4 = (idx & 15) * 4;
            const size_t g = koff + (size_t)j * p.ldkv + (size_t)h * p.d + c4;
            typedef SmemIO<(KIND == 1 ? 2 : KIND)> KVIO;     // 2-byte cache in the operand type of the mode
            const float4 kk = KVIO::ld4(reinterpret_cast<const typename KVIO::T*>(kb) + g);
            const float4 vv = KVIO::ld4(reinterpret_cast<const typename KVIO::T*>(kb) + g + p.heads * p.d);
            *reinterpret_cast<float4*>(Kw + j * A_LD + c4) = kk;
            *reinterpret_cast<float4*>(Vw + j * A_LD + c4) = vv;
          }
          __syncwarp();
          attend_head_mma<1, KIND>(Qs + (size_t)s * L * A_LD, A_LD, Kw, Vw, A_LD, L, p.nk, p.scale, p.att, ob, p.ldo, lane);
        }
      }
      if (quad_local) asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // staging is free for the next tile
      else asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace tc

static size_t gemm_attn_smem(const GemmAttnParams& p, int nk_max) {
  const int BN = p.cross ? p.d : 3 * p.d;
  const size_t stage = tc::A_ABYTES + (((size_t)BN * 128 + 1023) & ~(size_t)1023);
  size_t stg = (size_t)tc::A_TM * tc::A_LD * 4;   // Qs
  if (p.cross && p.kvf_c) {}                                                       // packed path: K / V fragments come from global
  else if (p.cross) stg += (size_t)tc::A_EPI_WARPS * 2 * 2 * nk_max * tc::A_LD * 4;   // per warp: two [K | V] scratch buffers
  else stg += 2 * (size_t)tc::A_TM * tc::A_LD * 4;
  return tc::attn_stages((p.cross && p.kvf_c) ? 3 : 0) * stage + stg + 1024;
}

bool gemm_attn_supported(int kind, int C, int L, int heads, int d, int cross, int nk_max) {
  const int kch = kind == 1 ? 32 : 64;
  if (d != 64 || C % kch || L < 1 || L > 64 || (128 % L) != 0) return false;
  if (!cross && L > 64) return false;
  GemmAttnParams p{}; p.cross = cross; p.d = d;
  return gemm_attn_smem(p, cross ? nk_max : 0) <= 220 * 1024 && (!cross || nk_max <= 64);
}

typedef void (*GemmAttnKernel)(const CUtensorMap, const CUtensorMap, const GemmAttnParams, const uint32_t);
static GemmAttnKernel gemm_attn_variant(int kind, int mode) {
  static const GemmAttnKernel tab[3][4] = {
      {tc::gemm_attn_kernel<1, 0>, tc::gemm_attn_kernel<1, 1>, tc::gemm_attn_kernel<1, 2>, tc::gemm_attn_kernel<1, 3>},
      {tc::gemm_attn_kernel<2, 0>, tc::gemm_attn_kernel<2, 1>, tc::gemm_attn_kernel<2, 2>, tc::gemm_attn_kernel<2, 3>},
      {tc::gemm_attn_kernel<3, 0>, tc::gemm_attn_kernel<3, 1>, tc::gemm_attn_kernel<3, 2>, tc::gemm_attn_kernel<3, 3>}};
  return tab[kind - 1][mode];
}

cudaError_t init_gemm_attn() {
  for (int kind = 1; kind <= 3; ++kind)
    for (int mode = 0; mode < 4; ++mode) {
      cudaError_t e = cudaFuncSetAttribute(gemm_attn_variant(kind, mode), cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      if (e != cudaSuccess) return e;
    }
  return cudaSuccess;
}

static int g_sms_ga = 0;

cudaError_t launch_gemm_attn(const void* tmA, const void* tmB, const GemmAttnParams& p, int kind, cudaStream_t s) {
  if (p.M <= 0) return cudaSuccess;
  if (g_sms_ga == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_ga, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms_ga <= 0) g_sms_ga = 148;
  }
  const int BN = p.cross ? p.d : 3 * p.d;
  const size_t smem = gemm_attn_smem(p, p.nk);
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  const uint32_t fmt = tc::umma_fmt(kind);
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(tc::A_TM >> 4) << 24);
  const long long tiles = (long long)((p.M + tc::A_TM - 1) / tc::A_TM) * p.heads;
  const unsigned grid = (unsigned)(tiles < g_sms_ga ? tiles : g_sms_ga);
  const CUtensorMap& a = *reinterpret_cast<const CUtensorMap*>(tmA);
  const CUtensorMap& b = *reinterpret_cast<const CUtensorMap*>(tmB);
  const int mode = p.cross ? (p.kvf_c ? 3 : 2) : ((p.pack_self && p.L <= 8) ? 1 : 0);
  return launch_k(gemm_attn_variant(kind, mode), grid, tc::A_THREADS, smem, s, a, b, p, idesc);
}

}  // namespace mdt
