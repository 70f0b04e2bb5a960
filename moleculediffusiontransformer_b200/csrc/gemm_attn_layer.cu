// gemm_attn_layer.cu -- one whole attention layer of TransformerBlock (modules.py:401-410, 350-364, 457-458) in one kernel
// for sm_100a: per-head projection (tcgen05) -> softmax attention (mma.sync in the epilogue warps) -> out-projection
// (tcgen05) + bias + residual, so neither q/k/v nor the (rows x heads*d) attention tensor ever reaches HBM.
//
// A CTA owns whole 128-row blocks (Sb = 128 / L samples) and walks the heads of a block one after the other:
//   J(k, h)  D_h[128, BN] = LN(x)[128, C] * Wh^T            (BN = 3d self / d cross; accumulator in TMEM, as gemm_attn.cu)
//   attn     epilogue warps stage q/k/v per sample and run softmax(q k^T * scale) v; the head output O_h[128, d] goes to a
//            CTA-private scratch slot in global memory (9 slots of 128 x d per CTA, 43 MB over all SMs: L2 resident, it is
//            rewritten every block and never needs to reach DRAM)
//   O(k)     OUT[128, Cout] = sum_h O_h * Wo[:, h*d:(h+1)*d]^T  on tcgen05 again: the A operand is TMA-loaded from the scratch slots,
//            the accumulator sits in its own TMEM columns across the eight heads
//   final    OUT + bias + residual -> fp32 token stream (and the operand-dtype copy the next GEMM reads)
// The out-projection of block k is issued behind the first head of block k+1 (ring order J(k+1,0), O(k), J(k+1,1), ...) and its
// epilogue runs after attention (k+1, 0), so its operand loads and MMAs hide under attention math instead of stalling the warps.
//
// Pipelines: smem ring (TMA producer <-> MMA issuer) shared by both GEMMs; TMEM projection accumulators (1 or 2) and the
// OUT accumulator (issuer <-> epilogue warps); att_ready (epilogue warps -> producer: the generic-proxy global writes of the
// scratch slots are fenced with fence.proxy.async before the async-proxy TMA reads).
#include <cuda.h>
#include <cuda_bf16.h>
#include "attn_math.cuh"

namespace mdt {
namespace tc {

constexpr int Y_TM = 128;
constexpr int Y_MAXST = 4;
constexpr int Y_ABYTES = Y_TM * 128;
constexpr int Y_EPI_WARPS = 8;
constexpr int Y_THREADS = 64 + 32 * Y_EPI_WARPS;
constexpr int Y_LD = 68;        // staged q/k/v row stride in floats (64 + 4: conflict-free fragment loads)
constexpr int Y_FLD = 36;       // final-epilogue transposition tile stride (32 + 4 floats: conflict-free float4 rows)
constexpr int Y_LA = 2;         // heads of block k + 1 that run ahead of block k's out-projection (scratch slots = heads + Y_LA)

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// MODE: 0 self, one sample per warp pass;  1 self, 16 / L short samples packed into one block-diagonal m16 tile (L <= 8);
//       3 cross, K / V fragments straight from the fragment-ordered cache (n_ctx <= 16).   (numbering of gemm_attn.cu)
template <int KIND, int MODE>
__global__ void __launch_bounds__(Y_THREADS, 1) attn_layer_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB,
                                                                 const __grid_constant__ CUtensorMap tmS,
                                                                 const __grid_constant__ CUtensorMap tmW,
                                                                 const AttnLayerParams p, const uint32_t idesc,
                                                                 const uint32_t idesc_o) {
  constexpr int KCH = (KIND == 1) ? 32 : 64;
  constexpr bool CROSS = MODE >= 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // SWIZZLE_128B operand tiles need 1024-byte alignment
  __shared__ __align__(8) uint64_t full_bar[Y_MAXST];
  __shared__ __align__(8) uint64_t empty_bar[Y_MAXST];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ __align__(8) uint64_t att_ready, out_full, out_empty;
  __shared__ uint32_t tmem_base_s;

  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const GemmAttnParams& a = p.a;
  const int heads = a.heads, d = a.d;
  const int BN = CROSS ? d : 3 * d;
  const int NST = p.nst, stage_bytes = p.stage_bytes, nacc = p.nacc;
  const int Cout = p.Cout;
  const int nblk = (a.M + Y_TM - 1) / Y_TM;
  const int nk_cta = (int)blockIdx.x < nblk ? (nblk - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int cph = d / KCH;                 // out-projection K chunks per head
  const int ochunks = heads * cph;
  float* Qs = reinterpret_cast<float*>(smem + NST * stage_bytes);
  float* Ks = Qs + Y_TM * Y_LD;
  float* Vs = Ks + Y_TM * Y_LD;
  auto block_of_k = [&](int k) { const int b = (int)blockIdx.x + k * (int)gridDim.x; return a.rev ? nblk - 1 - b : b; };

  if (tid == 0) {
    for (int s = 0; s < Y_MAXST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], Y_EPI_WARPS); }
    mbar_init(&att_ready, Y_EPI_WARPS * 32);
    mbar_init(&out_full, 1);
    mbar_init(&out_empty, Y_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmS); tma_prefetch_desc(&tmW); }
  if (warp == 1) tmem_alloc(&tmem_base_s, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_enter();                                  // the set-up above overlaps the previous grid's tail (launch.cuh)
  const uint32_t tmem_base = tmem_base_s;       // columns [0, Cout): OUT accumulator; [Cout + b * BN, ...): projection accumulators

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx_j = (uint32_t)(Y_ABYTES + BN * 128), tx_o = (uint32_t)(Y_ABYTES + Cout * 128);
      auto load_out = [&](int kk) {
        mbar_wait(&att_ready, (uint32_t)kk & 1u);        // every head output of block kk is in its scratch slot
        fence_proxy_async_all();
        const int j0 = kk * heads;
        for (int oc = 0; oc < ochunks; ++oc) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], tx_o);
          const int hh = oc / cph, sub = oc - hh * cph;
          const int slot = (j0 + hh) % p.nslot;
          tma_load_3d(sa, &tmS, &full_bar[stage], sub * KCH, slot * Y_TM, (int)blockIdx.x);
          tma_load_2d(sa + Y_ABYTES, &tmW, &full_bar[stage], oc * KCH, 0);
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      };
      for (int k = 0; k < nk_cta; ++k) {
        const int blk = block_of_k(k);
        for (int h = 0; h < heads; ++h) {
          if (h == Y_LA && k > 0) load_out(k - 1);
          for (int kc = 0; kc < a.kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * stage_bytes;
            mbar_arrive_expect_tx(&full_bar[stage], tx_j);
            tma_load_3d(sa, &tmA, &full_bar[stage], kc * KCH, 0, blk * a.Sb);
            tma_load_2d(sa + Y_ABYTES, &tmB, &full_bar[stage], kc * KCH, h * BN);
            if (++stage == NST) { stage = 0; phase ^= 1u; }
          }
        }
      }
      if (nk_cta > 0) load_out(nk_cta - 1);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int stage = 0; uint32_t phase = 0;
    int j = 0;
    auto mma_out = [&](int kk) {
      mbar_wait(&out_empty, ((uint32_t)kk & 1u) ^ 1u);   // the final epilogue of the previous block has drained OUT
      tc_fence_after();
      for (int oc = 0; oc < ochunks; ++oc) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + Y_ABYTES);
#pragma unroll
          for (int kq = 0; kq < 4; ++kq)
            umma<KIND>(tmem_base, adesc + (uint64_t)(2 * kq), bdesc + (uint64_t)(2 * kq), idesc_o, (uint32_t)((oc | kq) != 0));
          umma_commit(&empty_bar[stage]);
          if (oc == ochunks - 1) umma_commit(&out_full);
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
    };
    for (int k = 0; k < nk_cta; ++k) {
      for (int h = 0; h < heads; ++h, ++j) {
        if (h == Y_LA && k > 0) mma_out(k - 1);
        const int buf = nacc == 2 ? (j & 1) : 0;
        const uint32_t aphase = (uint32_t)(nacc == 2 ? (j >> 1) : j) & 1u;
        mbar_wait(&acc_empty[buf], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(Cout + buf * BN);
        for (int k0 = 0; k0 < a.kchunks; ++k0) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + stage * stage_bytes);
            const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + Y_ABYTES);
#pragma unroll
            for (int kq = 0; kq < 4; ++kq)
              umma<KIND>(tmem_d, adesc + (uint64_t)(2 * kq), bdesc + (uint64_t)(2 * kq), idesc, (uint32_t)((k0 | kq) != 0));
            umma_commit(&empty_bar[stage]);
            if (k0 == a.kchunks - 1) umma_commit(&acc_full[buf]);
          }
          __syncwarp();
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      }
    }
    if (nk_cta > 0) mma_out(nk_cta - 1);
  } else {
    // ------------------------------------------------------------------ epilogue + attention + final epilogue (warps 2..9)
    const int ew = warp - 2;
    const int q = warp & 3;                 // TMEM lane quadrant
    const int half = ew >> 2;
    const int row = q * 32 + lane;          // tile row owned for the TMEM -> smem transfers
    const int cols_per_warp = BN / 2;       // 96 (self) or 32 (cross)
    const int L = a.L;
    const int spq = 32 / L;                 // whole samples per quadrant (L <= 32), shared by its two warps
    const int s_begin = q * spq, s_end = (q + 1) * spq;
    // this warp's 32 x 32 transposition tile for the final epilogue: inside its quadrant's staged q (half 0) / k (half 1) rows, which are
    // dead between two heads; the cross modes stage q only and carry a dedicated region behind it
    float* fst = CROSS ? Ks + (size_t)ew * 32 * Y_FLD : (half ? Ks : Qs) + (size_t)q * 32 * Y_LD;
    const size_t cta_slot0 = (size_t)blockIdx.x * p.nslot;

    auto final_epilogue = [&](int kk) {
      const int m0 = block_of_k(kk) * Y_TM + q * 32;      // first token row of this warp's quadrant
      const int cols_w = Cout >> 1;
      const int cl = (lane & 7) * 4, r0 = lane >> 3;      // 8 lanes cover one 128-byte row segment, 4 rows per pass
      bool waited = false;
      for (int cc = 0; cc < cols_w; cc += 64) {
        // two 32-column chunks per pass; all residual loads are issued before the first use (HBM latency is paid once)
        const int nch = min(2, (cols_w - cc) >> 5);
        float4 r[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int no = half * cols_w + cc + u * 32 + cl;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int mo = m0 + r0 + i * 4;
            r[u][i] = (p.res && u < nch && mo < a.M) ? *reinterpret_cast<const float4*>(p.res + (size_t)mo * p.ldres + no)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (!waited) { mbar_wait(&out_full, (uint32_t)kk & 1u); tc_fence_after(); waited = true; }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u < nch) {
            const int col0 = half * cols_w + cc + u * 32;
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              *reinterpret_cast<uint4*>(fst + lane * Y_FLD + jj * 4) = make_uint4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
            __syncwarp();
            const int no = col0 + cl;
            const float4 bv = p.bias_o ? __ldg(reinterpret_cast<const float4*>(p.bias_o + no)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int mo = m0 + r0 + i * 4;
              if (mo < a.M) {
                float4 o = *reinterpret_cast<const float4*>(fst + (r0 + i * 4) * Y_FLD + cl);
                o.x += bv.x + r[u][i].x; o.y += bv.y + r[u][i].y; o.z += bv.z + r[u][i].z; o.w += bv.w + r[u][i].w;
                if (p.C32) *reinterpret_cast<float4*>(p.C32 + (size_t)mo * p.ldc + no) = o;
                if (p.Cop) {
                  if (KIND == 1)
                    *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.Cop) + (size_t)mo * p.ldcop + no) =
                        make_uint4(to_tf32(o.x), to_tf32(o.y), to_tf32(o.z), to_tf32(o.w));
                  else
                    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.Cop) + (size_t)mo * p.ldcop + no) =
                        make_uint2(pack_op2<KIND>(o.x, o.y), pack_op2<KIND>(o.z, o.w));
                }
              }
            }
            __syncwarp();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&out_empty);
    };

    int j = 0;
    for (int k = 0; k < nk_cta; ++k) {
      const int m0 = block_of_k(k) * Y_TM;
      for (int h = 0; h < heads; ++h, ++j) {
        const int buf = nacc == 2 ? (j & 1) : 0;
        const uint32_t aphase = (uint32_t)(nacc == 2 ? (j >> 1) : j) & 1u;
        mbar_wait(&acc_full[buf], aphase);
        tc_fence_after();
        for (int cc = 0; cc < cols_per_warp; cc += 32) {
          uint32_t v[32];
          const int col0 = half * cols_per_warp + cc;
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(Cout + buf * BN + col0), v);
          if (m0 + row >= a.M) {
            // rows past the batch are in bounds for the TMA box and hold stale data; the packed tiles multiply them by P = 0
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] = 0u;
          }
          float* dst = (col0 < 64 ? Qs : (col0 < 128 ? Ks : Vs)) + (size_t)row * Y_LD + (col0 & 63);
          if (col0 < 64) {
            // only q carries a bias: the k bias cancels in the softmax, the v bias is folded into the out-projection bias
            const float* bias = a.bias + h * d + col0;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + jj * 4));
              *reinterpret_cast<uint4*>(dst + jj * 4) =
                  make_uint4(to_tf32(__uint_as_float(v[4 * jj]) + bv.x), to_tf32(__uint_as_float(v[4 * jj + 1]) + bv.y),
                             to_tf32(__uint_as_float(v[4 * jj + 2]) + bv.z), to_tf32(__uint_as_float(v[4 * jj + 3]) + bv.w));
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              *reinterpret_cast<uint4*>(dst + jj * 4) =
                  make_uint4(to_tf32(__uint_as_float(v[4 * jj])), to_tf32(__uint_as_float(v[4 * jj + 1])),
                             to_tf32(__uint_as_float(v[4 * jj + 2])), to_tf32(__uint_as_float(v[4 * jj + 3])));
          }
        }
        tc_fence_before();
        // rows 32q .. 32q + 31 are whole samples staged by the two warps of quadrant q only: a 64-thread named barrier suffices
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        if (lane == 0) mbar_arrive(&acc_empty[buf]);             // the accumulator may be overwritten now
        const size_t slot_row0 = (cta_slot0 + (size_t)(j % p.nslot)) * Y_TM;   // first scratch row of this head's slot
        if constexpr (MODE == 3) {
          const int r16 = q * 32 + half * 16;
          const int mrow = m0 + r16;
          if (mrow < a.M) {
            const int bs = mrow / L;
            const size_t ob = (slot_row0 + (size_t)r16) * d;
            const int rows_valid = min(16, a.M - mrow);
            auto blockp = [&](int t) {
              const int b = (mrow + t * L < a.M) ? bs + t : bs;            // samples past the batch: any valid block, rows not stored
              const bool nul = a.kn && b >= a.n_cond;
              return reinterpret_cast<const uint2*>(nul ? a.kvf_n : a.kvf_c) + ((nul ? (size_t)0 : (size_t)b * heads) + h) * 1024;
            };
            if (L == 4) {
              const uint2* const kf[4] = {blockp(0), blockp(1), blockp(2), blockp(3)};
              attend_packed_cross<KIND, 4>(Qs + (size_t)r16 * Y_LD, Y_LD, kf, a.nk, a.scale, p.scratch, ob, d, rows_valid, lane);
            } else if (L == 8) {
              const uint2* const kf[2] = {blockp(0), blockp(1)};
              attend_packed_cross<KIND, 8>(Qs + (size_t)r16 * Y_LD, Y_LD, kf, a.nk, a.scale, p.scratch, ob, d, rows_valid, lane);
            } else {
              const uint2* const kf[1] = {blockp(0)};
              attend_packed_cross<KIND, 16>(Qs + (size_t)r16 * Y_LD, Y_LD, kf, a.nk, a.scale, p.scratch, ob, d, rows_valid, lane);
            }
          }
        } else if constexpr (MODE == 1) {
          const int r16 = q * 32 + half * 16;
          const int mrow = m0 + r16;
          if (mrow < a.M)
            attend_head_mma_nt<1, KIND, 2>(Qs + (size_t)r16 * Y_LD, Y_LD, Ks + (size_t)r16 * Y_LD, Vs + (size_t)r16 * Y_LD, Y_LD,
                                           min(16, a.M - mrow), 16, a.scale, p.scratch, (slot_row0 + (size_t)r16) * d, d, lane, L);
        } else {
          for (int s = s_begin + half; s < s_end; s += 2) {
            if (m0 + s * L >= a.M) break;
            attend_head_mma<1, KIND>(Qs + (size_t)s * L * Y_LD, Y_LD, Ks + (size_t)s * L * Y_LD, Vs + (size_t)s * L * Y_LD, Y_LD, L, L,
                                     a.scale, p.scratch, (slot_row0 + (size_t)s * L) * d, d, lane);
          }
        }
        if (h == heads - 1) {
          // publish this thread's head outputs of the whole block to the async proxy, then tell the producer
          fence_proxy_async_all();
          mbar_arrive(&att_ready);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // staging is free for the next head
        if (h == Y_LA - 1 && k > 0) {
          final_epilogue(k - 1);
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); // the transposition tiles live inside the q staging rows
        }
      }
    }
    if (nk_cta > 0) final_epilogue(nk_cta - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace tc

static int g_sms_layer = 0;
int attn_layer_sms() {
  if (g_sms_layer == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_layer, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms_layer <= 0) g_sms_layer = 148;
  }
  return g_sms_layer;
}

static const size_t Y_SMEM_LIMIT = 232448 - 1024;   // opt-in maximum minus room for the static barriers

static bool attn_layer_config(int d, int cross, int Cout, int* nst, int* stage_bytes, int* nacc, unsigned* tmem_cols, size_t* smem) {
  const int BN = cross ? d : 3 * d;
  const size_t sj = tc::Y_ABYTES + (((size_t)BN * 128 + 1023) & ~(size_t)1023), so = tc::Y_ABYTES + (size_t)Cout * 128;
  const size_t stage = sj > so ? sj : so;
  // staged q / k / v rows; the cross modes stage q only and add the eight final-epilogue transposition tiles behind it
  const size_t stg = cross ? (size_t)tc::Y_TM * tc::Y_LD * 4 + (size_t)tc::Y_EPI_WARPS * 32 * tc::Y_FLD * 4
                           : 3 * (size_t)tc::Y_TM * tc::Y_LD * 4;
  if (Y_SMEM_LIMIT < stg + 1024) return false;
  int n = (int)((Y_SMEM_LIMIT - stg - 1024) / stage);
  if (n > tc::Y_MAXST) n = tc::Y_MAXST;
  if (n < 2) return false;
  if (Cout + BN > 512) return false;
  const int na = (Cout + 2 * BN <= 512) ? 2 : 1;
  unsigned cols = 32;
  while ((int)cols < Cout + na * BN) cols <<= 1;
  *nst = n; *stage_bytes = (int)stage; *nacc = na; *tmem_cols = cols; *smem = (size_t)n * stage + stg + 1024;
  return true;
}

// cross != 0 needs the fragment-ordered K / V cache (packed path of gemm_attn.cu: L in {4, 8, 16}, n_ctx <= 16)
bool attn_layer_supported(int kind, int C, int L, int heads, int d, int cross, int Cout) {
  const int kch = kind == 1 ? 32 : 64;
  if (kind < 1 || kind > 3) return false;
  if (d != 64 || heads < 2 || C % kch || L < 1 || L > 32 || (128 % L) != 0) return false;
  if (cross && !(L == 4 || L == 8 || L == 16)) return false;
  if (Cout < 32 || Cout > 256 || Cout % 32) return false;
  int nst, sb, na; unsigned tc_; size_t sm;
  return attn_layer_config(d, cross, Cout, &nst, &sb, &na, &tc_, &sm);
}

int attn_layer_slots(int heads) { return heads + tc::Y_LA; }

size_t attn_layer_scratch_bytes(int kind, int heads, int d) {
  return (size_t)attn_layer_sms() * attn_layer_slots(heads) * tc::Y_TM * d * (kind == 1 ? 4 : 2);
}

typedef void (*AttnLayerKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnLayerParams,
                                const uint32_t, const uint32_t);
static AttnLayerKernel attn_layer_variant(int kind, int mode) {
  static const AttnLayerKernel tab[3][3] = {
      {tc::attn_layer_kernel<1, 0>, tc::attn_layer_kernel<1, 1>, tc::attn_layer_kernel<1, 3>},
      {tc::attn_layer_kernel<2, 0>, tc::attn_layer_kernel<2, 1>, tc::attn_layer_kernel<2, 3>},
      {tc::attn_layer_kernel<3, 0>, tc::attn_layer_kernel<3, 1>, tc::attn_layer_kernel<3, 3>}};
  return tab[kind - 1][mode == 3 ? 2 : mode];
}

cudaError_t init_attn_layer() {
  for (int kind = 1; kind <= 3; ++kind)
    for (int mode : {0, 1, 3}) {
      cudaError_t e = cudaFuncSetAttribute(attn_layer_variant(kind, mode), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Y_SMEM_LIMIT);
      if (e != cudaSuccess) return e;
    }
  return cudaSuccess;
}

cudaError_t launch_attn_layer(const void* tmA, const void* tmB, const void* tmS, const void* tmW, const AttnLayerParams& pin, int kind,
                              cudaStream_t s) {
  AttnLayerParams p = pin;
  const GemmAttnParams& a = p.a;
  if (a.M <= 0) return cudaSuccess;
  size_t smem = 0;
  if (!attn_layer_config(a.d, a.cross, p.Cout, &p.nst, &p.stage_bytes, &p.nacc, &p.tmem_cols, &smem)) return cudaErrorInvalidValue;
  if (a.cross && !a.kvf_c) return cudaErrorInvalidValue;
  p.nslot = a.heads + tc::Y_LA;
  const int BN = a.cross ? a.d : 3 * a.d;
  const uint32_t fmt = tc::umma_fmt(kind);
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(tc::Y_TM >> 4) << 24);
  const uint32_t idesc_o = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.Cout >> 3) << 17) | ((uint32_t)(tc::Y_TM >> 4) << 24);
  const int nblk = (a.M + tc::Y_TM - 1) / tc::Y_TM;
  const int sms = attn_layer_sms();
  const unsigned grid = (unsigned)(nblk < sms ? nblk : sms);
  const int mode = a.cross ? 3 : ((a.pack_self && a.L <= 8) ? 1 : 0);
  return launch_k(attn_layer_variant(kind, mode), grid, tc::Y_THREADS, smem, s, *reinterpret_cast<const CUtensorMap*>(tmA),
                  *reinterpret_cast<const CUtensorMap*>(tmB), *reinterpret_cast<const CUtensorMap*>(tmS),
                  *reinterpret_cast<const CUtensorMap*>(tmW), p, idesc, idesc_o);
}

}  // namespace mdt
