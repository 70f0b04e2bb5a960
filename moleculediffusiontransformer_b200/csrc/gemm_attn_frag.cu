// gemm_attn_frag.cu -- second generation of the whole-attention-layer kernel (gemm_attn_layer.cu): same pipelines (per-head
// projection on tcgen05 -> softmax attention -> out-projection on tcgen05 from CTA-private L2 scratch -> bias + residual), but the
// attention warps read their mma.sync operand fragments STRAIGHT FROM TMEM:
//
//   tcgen05.ld.16x256b delivers 16 accumulator rows x 8 columns per register quad in exactly the mma.m16n8 accumulator layout
//   (thread (g, q): row g and g + 8, columns 2q and 2q + 1).  With the k index relabelled consistently on both operands
//   (k = q <-> column 2q, k = q + 4 <-> column 2q + 1) that quad IS the A fragment of q and, read from the key rows, the two B
//   fragments of k.  So q and k never go through shared memory; only v (whose B fragment is transposed) passes through a
//   warp-private 16 x 64 tile.  No cross-warp barrier is left in the head loop, the staging shared memory shrinks from 104 KB to
//   68 KB, and SIXTEEN attention warps (two groups of eight, one per TMEM accumulator) keep two heads in flight.
//
// Supported: self-attention with L in {4, 8, 16} (short samples share one block-diagonal m16 tile) and cross-attention on the
// fragment-ordered K/V cache with L in {4, 8, 16}; everything else stays on gemm_attn_layer.cu.
#include <cuda.h>
#include <cuda_bf16.h>
#include "attn_math.cuh"

namespace mdt {
namespace tc {

constexpr int Z_TM = 128;
constexpr int Z_MAXST = 4;
constexpr int Z_ABYTES = Z_TM * 128;
constexpr int Z_EPI_WARPS = 16;
constexpr int Z_THREADS = 64 + 32 * Z_EPI_WARPS;
constexpr int Z_VLD = 68;       // warp-private v tile row stride in floats (64 + 4: conflict-free fragment loads)
constexpr int Z_LA = 2;         // heads of block k + 1 that run ahead of block k's out-projection (scratch slots = heads + Z_LA)

__device__ __forceinline__ void z_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// 16 TMEM lanes x 32 columns -> 16 registers: quad c = columns 8c .. 8c + 7 in the m16n8 accumulator layout (no wait)
__device__ __forceinline__ void tmem_ld16x256_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

__device__ __forceinline__ void mma_f16_16x8x16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two fp32 values -> one f16x2 register (low half = first): fp16 carries the same 11-bit significand as tf32, so the attention core
// keeps its precision while one m16n8k16 MMA replaces two m16n8k8 ones (operands here are O(1) after LayerNorm: no range issue)
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// softmax over <= 16 keys held as two m16n8 score tiles (thread: rows g / g + 8, keys 8t + 2q, 8t + 2q + 1); returns the
// probabilities as tf32 A fragments of P V (key permutation of attn_math.cuh).  bm: block-diagonal mask (~(blk - 1)) or 0.
template <bool F16>
__device__ __forceinline__ void softmax_2tiles(float (&sc)[2][4], int nk, float scale, int bm, int g, int q, uint32_t (&pa)[2][4]) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int j = t * 8 + 2 * q;
    const bool in0 = ((j ^ g) & bm) == 0, in1 = ((j ^ (g + 8)) & bm) == 0;
    sc[t][0] = (j < nk && in0) ? sc[t][0] * scale : -INFINITY;
    sc[t][1] = (j + 1 < nk && in0) ? sc[t][1] * scale : -INFINITY;
    sc[t][2] = (j < nk && in1) ? sc[t][2] * scale : -INFINITY;
    sc[t][3] = (j + 1 < nk && in1) ? sc[t][3] * scale : -INFINITY;
    m0 = fmaxf(m0, fmaxf(sc[t][0], sc[t][1]));
    m1 = fmaxf(m1, fmaxf(sc[t][2], sc[t][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    sc[t][0] = __expf(sc[t][0] - m0); sc[t][1] = __expf(sc[t][1] - m0);
    sc[t][2] = __expf(sc[t][2] - m1); sc[t][3] = __expf(sc[t][3] - m1);
    s0 += sc[t][0] + sc[t][1]; s1 += sc[t][2] + sc[t][3];
  }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float inv0 = 1.0f / s0, inv1 = 1.0f / s1;
  if (F16) {
    // one m16n8k16 A fragment over the 16 keys: a0 = (row g, keys 2q, 2q + 1), a1 = (row g + 8, same), a2 / a3 = keys + 8
    pa[0][0] = pack_f16(sc[0][0] * inv0, sc[0][1] * inv0);
    pa[0][1] = pack_f16(sc[0][2] * inv1, sc[0][3] * inv1);
    pa[0][2] = pack_f16(sc[1][0] * inv0, sc[1][1] * inv0);
    pa[0][3] = pack_f16(sc[1][2] * inv1, sc[1][3] * inv1);
    return;
  }
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    pa[t][0] = to_tf32(sc[t][0] * inv0);   // (row g,     key 8t + 2q)
    pa[t][1] = to_tf32(sc[t][2] * inv1);   // (row g + 8, key 8t + 2q)
    pa[t][2] = to_tf32(sc[t][1] * inv0);   // (row g,     key 8t + 2q + 1)
    pa[t][3] = to_tf32(sc[t][3] * inv1);   // (row g + 8, key 8t + 2q + 1)
  }
}

// q fragments (with the folded q bias) for the 16 rows at `tq`: two halves of 32 features each are loaded by the caller
template <int KIND>
__device__ __forceinline__ void add_q_bias(uint32_t (&qf)[16], const float* __restrict__ bias32, int q) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float2 b = __ldg(reinterpret_cast<const float2*>(bias32 + 8 * c + 2 * q));
    qf[4 * c + 0] = __float_as_uint(__uint_as_float(qf[4 * c + 0]) + b.x);
    qf[4 * c + 1] = __float_as_uint(__uint_as_float(qf[4 * c + 1]) + b.y);
    qf[4 * c + 2] = __float_as_uint(__uint_as_float(qf[4 * c + 2]) + b.x);
    qf[4 * c + 3] = __float_as_uint(__uint_as_float(qf[4 * c + 3]) + b.y);
  }
}

// hint != 0 (tf32 scratch slots): stores carry an L2 evict_last policy, the slot is read back by TMA within microseconds
template <int OK>
__device__ __forceinline__ void store_o_tiles(const float (&oc)[8][4], void* out, size_t out_base, int ldo, int rows_valid, int g, int q,
                                              uint64_t hint = 0) {
  typedef SmemIO<OK> OUT;
  const bool ok0 = g < rows_valid, ok1 = g + 8 < rows_valid;
  const size_t o0 = out_base + (size_t)g * ldo + 2 * q, o1 = o0 + (size_t)8 * ldo;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    if (OK == 1 && hint) {
      uint32_t* ob = reinterpret_cast<uint32_t*>(out);
      if (ok0) st_global_hint_b32x2(ob + o0 + n * 8, to_tf32(oc[n][0]), to_tf32(oc[n][1]), hint);
      if (ok1) st_global_hint_b32x2(ob + o1 + n * 8, to_tf32(oc[n][2]), to_tf32(oc[n][3]), hint);
    } else {
      if (ok0) OUT::st2(out, o0 + n * 8, oc[n][0], oc[n][1]);
      if (ok1) OUT::st2(out, o1 + n * 8, oc[n][2], oc[n][3]);
    }
  }
}

// MODE: 0 self (L in {4, 8, 16}: 16 rows = 16 / L samples, block-diagonal mask when L < 16);  4 / 8 / 16: cross-attention on the
// fragment-ordered cache with that many query rows per sample (compile time: 16 / MODE samples share the m16 tile)
// MATH: 0 = tf32 m16n8k8 attention core, 1 = f16 m16n8k16 (same 11-bit significand, half the MMAs and fragment loads)
template <int KIND, int MODE, int MATH>
__global__ void __launch_bounds__(Z_THREADS, 1) attn_frag_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ CUtensorMap tmS,
                                                                const __grid_constant__ CUtensorMap tmW,
                                                                const AttnLayerParams p, const uint32_t idesc,
                                                                const uint32_t idesc_o) {
  constexpr int KCH = (KIND == 1) ? 32 : 64;
  constexpr bool CROSS = MODE != 0;
  constexpr int LQ = CROSS ? MODE : 16;
  constexpr int G = 16 / LQ;                // samples sharing one m16 tile (cross)
  constexpr bool F16 = MATH == 1;
  constexpr int NKF = F16 ? 8 : 16;         // K (and V) fragments per (sample, head) block of the cross-attention cache
  constexpr int VOFF = F16 ? 256 : 512;     // uint2 offset of the V fragments inside a block
  constexpr int VKP = 24;                   // f16 v tile: halfs per feature row (16 keys + 8 padding: conflict-free 32-bit fragment loads)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[Z_MAXST];
  __shared__ __align__(8) uint64_t empty_bar[Z_MAXST];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ __align__(8) uint64_t att_ready, out_full, out_empty;
  __shared__ __align__(8) uint64_t a_full, a_empty;  // resident activation tile (p.ares_bytes != 0)
  __shared__ __align__(8) uint64_t ln_bar[4];       // cop_ln: per TMEM lane quadrant, "row statistics of this block are in shared memory"
  __shared__ uint32_t tmem_base_s;

  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const GemmAttnParams& a = p.a;
  const int heads = a.heads, d = a.d;
  const int BN = CROSS ? d : 3 * d;
  const int NST = p.nst, stage_bytes = p.stage_bytes;
  // Resident activation tile: the 128 x C operand tile of a row block is loaded ONCE into its own region and serves the projections of
  // all heads of the block (fused: the eight heads in sequence; per-head variant: a CTA owns a contiguous item range, so consecutive
  // items share the block) instead of being re-streamed through L2 with every head's weight slice.  The stages then carry only the
  // weight chunk (at offset 0) for projections; out-projection chunks keep [scratch A | W_o].  Chosen by the launcher where it fits.
  const bool ares = p.ares_bytes != 0;
  uint8_t* const stg0 = smem + p.ares_bytes;     // stage ring behind the resident tile
  const int boff = ares ? 0 : Z_ABYTES;          // weight chunk offset inside a projection stage
  const int Cout = p.Cout;
  const int nblk = (a.M + Z_TM - 1) / Z_TM;
  const int cph = d / KCH;
  const int ochunks = heads * cph;
  // fused: a work item is a whole row block (its heads in sequence, then the out-projection); otherwise one (row block, head) pair
  // whose head output goes to the global attention tensor (the out-projection is a separate GEMM) -- eight times more items, which is
  // what balances short levels (256 row blocks on 148 SMs) across the grid
  const bool fused = p.fused != 0;
  const int nitems = fused ? nblk : nblk * heads;
  // per-head variant with a resident tile: contiguous item ranges (items of one row block are consecutive); otherwise round robin
  const bool contig = ares && !fused;
  const int it_lo = contig ? (int)(((long long)blockIdx.x * nitems) / (int)gridDim.x) : 0;
  const int it_hi = contig ? (int)(((long long)(blockIdx.x + 1) * nitems) / (int)gridDim.x) : 0;
  const int nk_cta = contig ? it_hi - it_lo : ((int)blockIdx.x < nitems ? (nitems - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0);
  const int njobs = fused ? nk_cta * heads : nk_cta;
  auto item_of_k = [&](int k) { const int b = contig ? it_lo + k : (int)blockIdx.x + k * (int)gridDim.x; return a.rev ? nitems - 1 - b : b; };
  auto block_of_k = [&](int k) { return item_of_k(k); };      // fused mode: item == row block

  if (tid == 0) {
    for (int s = 0; s < Z_MAXST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], Z_EPI_WARPS / 2); }
    mbar_init(&att_ready, Z_EPI_WARPS * 32);
    mbar_init(&out_full, 1);
    mbar_init(&out_empty, Z_EPI_WARPS);
    mbar_init(&a_full, 1); mbar_init(&a_empty, 1);
    for (int s = 0; s < 4; ++s) mbar_init(&ln_bar[s], 4);      // the four warps that share a quadrant's rows (column quarters)
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    if (p.fused) { tma_prefetch_desc(&tmS); tma_prefetch_desc(&tmW); }     // the unfused variant carries no scratch / out-projection maps
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_enter();                                  // the set-up above overlaps the previous grid's tail (launch.cuh)
  const uint32_t tmem_base = tmem_base_s;       // columns [0, Cout): OUT accumulator; [Cout + b * BN, ...): two projection accumulators

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx_j = (uint32_t)(Z_ABYTES + BN * 128), tx_o = (uint32_t)(Z_ABYTES + Cout * 128);
      auto load_out = [&](int kk) {
        mbar_wait(&att_ready, (uint32_t)kk & 1u);
        z_fence_proxy_async();
        const int j0 = kk * heads;
        for (int oc = 0; oc < ochunks; ++oc) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = stg0 + stage * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], tx_o);
          const int hh = oc / cph, sub = oc - hh * cph;
          const int slot = (j0 + hh) % p.nslot;
          tma_load_3d(sa, &tmS, &full_bar[stage], sub * KCH, slot * Z_TM, (int)blockIdx.x);
          tma_load_2d(sa + Z_ABYTES, &tmW, &full_bar[stage], oc * KCH, 0);
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      };
      int k = 0, h = 0;
      int cur_blk = -1, n_a = 0;                    // resident tile: row block it holds, tiles loaded so far
      for (int j = 0; j < njobs; ++j) {
        int blk;
        if (fused) { blk = block_of_k(k); if (h == Z_LA && k > 0) load_out(k - 1); }
        else { const int it = item_of_k(j); blk = it / heads; h = it - blk * heads; }
        if (ares && blk != cur_blk) {
          if (n_a > 0) mbar_wait(&a_empty, (uint32_t)(n_a - 1) & 1u);        // every MMA that read the previous tile has completed
          mbar_arrive_expect_tx(&a_full, (uint32_t)(a.kchunks * Z_ABYTES));
          for (int kc = 0; kc < a.kchunks; ++kc) tma_load_3d(smem + kc * Z_ABYTES, &tmA, &a_full, kc * KCH, 0, blk * a.Sb);
          cur_blk = blk; ++n_a;
        }
        for (int kc = 0; kc < a.kchunks; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = stg0 + stage * stage_bytes;
          if (ares) {
            mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(BN * 128));
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], tx_j);
            tma_load_3d(sa, &tmA, &full_bar[stage], kc * KCH, 0, blk * a.Sb);
          }
          tma_load_2d(sa + boff, &tmB, &full_bar[stage], kc * KCH, h * BN);
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
        if (fused && ++h == heads) { h = 0; ++k; }
      }
      if (fused && nk_cta > 0) load_out(nk_cta - 1);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int stage = 0; uint32_t phase = 0;
    int j = 0;
    auto mma_out = [&](int kk) {
      mbar_wait(&out_empty, ((uint32_t)kk & 1u) ^ 1u);
      tc_fence_after();
      for (int oc = 0; oc < ochunks; ++oc) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(stg0 + stage * stage_bytes);
          const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + Z_ABYTES);
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
    int k = 0, h = 0;
    int cur_blk = -1, m_a = 0;
    const uint32_t ares_u32 = smem_u32(smem);
    for (; j < njobs; ++j) {
      {
        if (fused && h == Z_LA && k > 0) mma_out(k - 1);
        int blk = 0; bool last_use = false;
        if (ares) {
          if (fused) { blk = block_of_k(k); last_use = h == heads - 1; }
          else { blk = item_of_k(j) / heads; last_use = j == njobs - 1 || item_of_k(j + 1) / heads != blk; }
          if (blk != cur_blk) { mbar_wait(&a_full, (uint32_t)m_a & 1u); tc_fence_after(); cur_blk = blk; ++m_a; }
        }
        const int buf = j & 1;
        mbar_wait(&acc_empty[buf], ((uint32_t)(j >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(Cout + buf * BN);
        for (int k0 = 0; k0 < a.kchunks; ++k0) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(stg0 + stage * stage_bytes);
            const uint64_t adesc = make_desc(ares ? ares_u32 + (uint32_t)(k0 * Z_ABYTES) : sa), bdesc = make_desc(sa + (uint32_t)boff);
#pragma unroll
            for (int kq = 0; kq < 4; ++kq)
              umma<KIND>(tmem_d, adesc + (uint64_t)(2 * kq), bdesc + (uint64_t)(2 * kq), idesc, (uint32_t)((k0 | kq) != 0));
            umma_commit(&empty_bar[stage]);
            if (k0 == a.kchunks - 1) {
              umma_commit(&acc_full[buf]);
              if (last_use) umma_commit(&a_empty);      // the resident tile may be replaced once these MMAs have completed
            }
          }
          __syncwarp();
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
        if (fused && ++h == heads) { h = 0; ++k; }
      }
    }
    if (fused && nk_cta > 0) mma_out(nk_cta - 1);
  } else {
    // ------------------------------------------------------------------ attention warps (2..17): group gp owns accumulator gp
    const int ew = warp - 2;
    const int gp = ew >> 3;                 // head parity / projection accumulator of this warp
    const int qd = warp & 3;                // TMEM lane quadrant
    const int half = (ew >> 2) & 1;         // which 16 rows of the quadrant
    const int r16 = qd * 32 + half * 16;    // first tile row of this warp's m16 tile
    const int g = lane >> 2, q = lane & 3;
    const int L = a.L;
    const int bm = (!CROSS && L < 16) ? ~(L - 1) : 0;
    constexpr int VT_BYTES = F16 ? 64 * VKP * 2 : 16 * Z_VLD * 4;      // warp-private v tile (self only): f16 [64 features][24], tf32 [16 keys][68]
    float* vt = reinterpret_cast<float*>(stg0 + NST * stage_bytes + (size_t)ew * VT_BYTES);
    const size_t cta_slot0 = (size_t)blockIdx.x * p.nslot;
    const uint32_t lane_addr = (uint32_t)r16 << 16;
    const int sub = gp * 2 + half;          // this warp's column quarter of the OUT tile (four warps per quadrant)
    const uint64_t l2pol = p.l2_hint ? l2_policy_evict_last() : 0;
    // cop_ln: LayerNorm(row) of the layer output as the operand copy.  First pass (final_epilogue): every warp leaves (mean, M2) of its
    // column quarter per row in shared memory, double buffered by block parity; second pass (ln_pass2) one head later -- by then the
    // other warp group has long finished its first pass, so the mbarrier wait costs nothing -- merges the four quarters and normalises
    // the values it re-reads from its own C32 stores.
    float2* xch = reinterpret_cast<float2*>(stg0 + NST * stage_bytes + (CROSS ? 0 : (size_t)Z_EPI_WARPS * VT_BYTES));   // [2][128][4]
    const bool cop_ln = !CROSS && p.cop_ln != 0;     // self-attention layers only (the cross layer hands FeedForward a raw copy)

    auto final_epilogue = [&](int kk) {
      const int m0 = block_of_k(kk) * Z_TM + qd * 32;      // first token row of this warp's quadrant
      const int cols_w = Cout >> 2;
      bool waited = false;
      float mean_w[2][2], m2_w[2][2];                      // cop_ln: statistics of rows {g, g + 8} + 16 hh over this warp's columns
      for (int cc = 0; cc < cols_w; cc += 32) {
        const int col0 = sub * cols_w + cc;
        // residual first (its HBM latency hides behind the accumulator wait): rows {g, g + 8} + 16 * hh, column pairs 8c + 2q
        float2 r[2][2][4];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const int mo = m0 + hh * 16 + rr * 8 + g;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              r[hh][rr][c] = (p.res && mo < a.M) ? *reinterpret_cast<const float2*>(p.res + (size_t)mo * p.ldres + col0 + 8 * c + 2 * q)
                                                 : make_float2(0.f, 0.f);
          }
        if (!waited) { mbar_wait(&out_full, (uint32_t)kk & 1u); tc_fence_after(); waited = true; }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[16];
          tmem_ld16x256_x4(tmem_base + ((uint32_t)(qd * 32 + hh * 16) << 16) + (uint32_t)col0, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int no = col0 + 8 * c + 2 * q;
            const float2 bv = p.bias_o ? __ldg(reinterpret_cast<const float2*>(p.bias_o + no)) : make_float2(0.f, 0.f);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int mo = m0 + hh * 16 + rr * 8 + g;
              if (mo < a.M) {
                const float ox = __uint_as_float(v[4 * c + 2 * rr]) + bv.x + r[hh][rr][c].x;
                const float oy = __uint_as_float(v[4 * c + 2 * rr + 1]) + bv.y + r[hh][rr][c].y;
                if (p.C32) *reinterpret_cast<float2*>(p.C32 + (size_t)mo * p.ldc + no) = make_float2(ox, oy);
                if (p.Cop && !cop_ln) SmemIO<KIND>::st2(p.Cop, (size_t)mo * p.ldcop + no, ox, oy);
                r[hh][rr][c] = make_float2(ox, oy);
              } else {
                r[hh][rr][c] = make_float2(0.f, 0.f);
              }
            }
          }
        }
        if (cop_ln) {
          // exact mean / centred M2 of this 32-column chunk: eight values per row in each of the four lanes of a quad
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              float sm = 0.f;
#pragma unroll
              for (int c = 0; c < 4; ++c) sm += r[hh][rr][c].x + r[hh][rr][c].y;
              sm += __shfl_xor_sync(0xffffffffu, sm, 1); sm += __shfl_xor_sync(0xffffffffu, sm, 2);
              const float mean = sm * (1.0f / 32.0f);
              float m2 = 0.f;
#pragma unroll
              for (int c = 0; c < 4; ++c) { const float dx = r[hh][rr][c].x - mean, dy = r[hh][rr][c].y - mean; m2 = fmaf(dx, dx, fmaf(dy, dy, m2)); }
              m2 += __shfl_xor_sync(0xffffffffu, m2, 1); m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
              if (cc == 0) { mean_w[hh][rr] = mean; m2_w[hh][rr] = m2; }
              else {       // second chunk (Cout = 256): equal-count merge
                const float dm = mean - mean_w[hh][rr];
                mean_w[hh][rr] = 0.5f * (mean_w[hh][rr] + mean); m2_w[hh][rr] = m2_w[hh][rr] + m2 + 16.0f * dm * dm;
              }
            }
        }
      }
      if (cop_ln) {
        float2* xb = xch + (size_t)(kk & 1) * Z_TM * 4;
        if (q == 0) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) xb[(qd * 32 + hh * 16 + rr * 8 + g) * 4 + sub] = make_float2(mean_w[hh][rr], m2_w[hh][rr]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ln_bar[qd]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&out_empty);
    };

    auto ln_pass2 = [&](int kk) {
      const int m0 = block_of_k(kk) * Z_TM + qd * 32;
      const int cols_w = Cout >> 2;
      const float2* xb = xch + (size_t)(kk & 1) * Z_TM * 4;
      const float inv_n = 1.0f / (float)Cout, n_w = (float)cols_w;
      for (int cc = 0; cc < cols_w; cc += 32) {
        const int col0 = sub * cols_w + cc;
        float2 o[2][2][4];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const int mo = m0 + hh * 16 + rr * 8 + g;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              o[hh][rr][c] = mo < a.M ? *reinterpret_cast<const float2*>(p.C32 + (size_t)mo * p.ldc + col0 + 8 * c + 2 * q) : make_float2(0.f, 0.f);
          }
        if (cc == 0) mbar_wait(&ln_bar[qd], (uint32_t)kk & 1u);      // after the loads are in flight
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const int row = qd * 32 + hh * 16 + rr * 8 + g, mo = m0 + hh * 16 + rr * 8 + g;
            const float4 e01 = *reinterpret_cast<const float4*>(xb + row * 4), e23 = *reinterpret_cast<const float4*>(xb + row * 4 + 2);
            const float mean = 0.25f * ((e01.x + e01.z) + (e23.x + e23.z));
            const float d0 = e01.x - mean, d1 = e01.z - mean, d2 = e23.x - mean, d3 = e23.z - mean;
            const float dev = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
            const float rstd = rsqrtf(fmaf(n_w, dev, (e01.y + e01.w) + (e23.y + e23.w)) * inv_n + p.ln_eps);
            if (mo < a.M) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                SmemIO<KIND>::st2(p.Cop, (size_t)mo * p.ldcop + col0 + 8 * c + 2 * q, (o[hh][rr][c].x - mean) * rstd, (o[hh][rr][c].y - mean) * rstd);
            }
          }
      }
    };

    for (int j = gp; j < njobs; j += 2) {
      int k, h, m0;
      if (fused) { k = j / heads; h = j - k * heads; m0 = block_of_k(k) * Z_TM; }
      else { const int it = item_of_k(j); k = it / heads; h = it - k * heads; m0 = k * Z_TM; }
      const int mrow = m0 + r16;
      const int rows_valid = min(16, a.M - mrow);       // <= 0: this warp's rows are past the batch
      {
        const int buf = gp;
        // cross: the K / V fragment blocks of this tile's samples and head; with one sample per tile the K fragments are fetched
        // before the accumulator wait (they do not depend on the projection), so their L2 / HBM latency hides behind it
        const uint2* kvp[G];
        uint2 kb[NKF];
        if constexpr (CROSS) {
          const int bs = mrow / LQ;
#pragma unroll
          for (int t = 0; t < G; ++t) {
            const int b = (mrow + t * LQ < a.M) ? bs + t : bs;            // samples past the batch: any valid block, rows not stored
            const bool nul = a.kn && b >= a.n_cond;
            kvp[t] = reinterpret_cast<const uint2*>(nul ? a.kvf_n : a.kvf_c) + ((nul ? (size_t)0 : (size_t)b * heads) + h) * 1024;
          }
          if (G == 1 && rows_valid > 0) {
#pragma unroll
            for (int f = 0; f < NKF; ++f) kb[f] = __ldg(kvp[0] + f * 32 + lane);
          }
        }
        mbar_wait(&acc_full[buf], (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
        const uint32_t tq = tmem_base + lane_addr + (uint32_t)(Cout + buf * BN);
        // fused: this warp's rows of the head's scratch slot; otherwise its rows / head columns of the global attention tensor
        void* const obuf = fused ? p.scratch : a.att;
        const int ldo = fused ? d : a.ldo;
        const size_t ob = fused ? ((cta_slot0 + (size_t)(j % p.nslot)) * Z_TM + (size_t)r16) * d : (size_t)mrow * a.ldo + (size_t)h * d;
        float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        float oc[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) { oc[n][0] = 0.f; oc[n][1] = 0.f; oc[n][2] = 0.f; oc[n][3] = 0.f; }
        if constexpr (!CROSS) {
          // ---- v: 16 key rows x 64 features -> warp-private tile (rows past the batch zeroed: they meet P = 0 in the packed tiles)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t vf[16];
            tmem_ld16x256_x4(tq + (uint32_t)(2 * d + hf * 32), vf);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const bool z0 = g >= rows_valid, z1 = g + 8 >= rows_valid;
              if (F16) {
                // transposed f16 tile [feature][key]: a B fragment of P V (two consecutive keys of one feature) is one 32-bit load
                __half* vh = reinterpret_cast<__half*>(vt) + (hf * 32 + 8 * c + 2 * q) * VKP + g;
                vh[0] = __float2half_rn(z0 ? 0.f : __uint_as_float(vf[4 * c]));
                vh[VKP] = __float2half_rn(z0 ? 0.f : __uint_as_float(vf[4 * c + 1]));
                vh[8] = __float2half_rn(z1 ? 0.f : __uint_as_float(vf[4 * c + 2]));
                vh[VKP + 8] = __float2half_rn(z1 ? 0.f : __uint_as_float(vf[4 * c + 3]));
              } else {
                *reinterpret_cast<uint2*>(vt + g * Z_VLD + hf * 32 + 8 * c + 2 * q) =
                    make_uint2(z0 ? 0u : to_tf32(__uint_as_float(vf[4 * c])), z0 ? 0u : to_tf32(__uint_as_float(vf[4 * c + 1])));
                *reinterpret_cast<uint2*>(vt + (g + 8) * Z_VLD + hf * 32 + 8 * c + 2 * q) =
                    make_uint2(z1 ? 0u : to_tf32(__uint_as_float(vf[4 * c + 2])), z1 ? 0u : to_tf32(__uint_as_float(vf[4 * c + 3])));
              }
            }
          }
          // ---- scores: q and k fragments straight from TMEM, 32 features per pass
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t qf[16], kf[16];
            tmem_ld16x256_x4(tq + (uint32_t)(hf * 32), qf);
            tmem_ld16x256_x4(tq + (uint32_t)(d + hf * 32), kf);
            tmem_ld_wait();
            add_q_bias<KIND>(qf, a.bias + h * d + hf * 32, q);
            if (F16) {
#pragma unroll
              for (int c = 0; c < 4; c += 2) {     // one k16 step = two 8-column quads
                auto pk = [](const uint32_t (&r)[16], int i) { return pack_f16(__uint_as_float(r[i]), __uint_as_float(r[i + 1])); };
                const uint32_t af[4] = {pk(qf, 4 * c), pk(qf, 4 * c + 2), pk(qf, 4 * c + 4), pk(qf, 4 * c + 6)};
                mma_f16_16x8x16(sc[0], af, pk(kf, 4 * c), pk(kf, 4 * c + 4));          // keys g
                mma_f16_16x8x16(sc[1], af, pk(kf, 4 * c + 2), pk(kf, 4 * c + 6));      // keys g + 8
              }
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint32_t af[4] = {qf[4 * c], qf[4 * c + 2], qf[4 * c + 1], qf[4 * c + 3]};
                mma_tf32_16x8x8(sc[0], af, kf[4 * c], kf[4 * c + 1]);          // keys g
                mma_tf32_16x8x8(sc[1], af, kf[4 * c + 2], kf[4 * c + 3]);      // keys g + 8
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);       // every TMEM read of this warp is complete
          if (rows_valid > 0) {
            uint32_t pa[2][4];
            softmax_2tiles<F16>(sc, 16, a.scale, bm, g, q, pa);
            if (F16) {
              const uint32_t* v0 = reinterpret_cast<const uint32_t*>(vt) + g * (VKP / 2) + q;     // feature 8n + g, keys 2q, 2q + 1
#pragma unroll
              for (int n = 0; n < 8; ++n) mma_f16_16x8x16(oc[n], pa[0], v0[n * 8 * (VKP / 2)], v0[n * 8 * (VKP / 2) + 4]);
            } else {
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                const float* v0 = vt + (t * 8 + 2 * q) * Z_VLD + g;
#pragma unroll
                for (int n = 0; n < 8; ++n)
                  mma_tf32_16x8x8(oc[n], pa[t], __float_as_uint(v0[n * 8]), __float_as_uint(v0[Z_VLD + n * 8]));
              }
            }
            store_o_tiles<KIND>(oc, obuf, ob, ldo, rows_valid, g, q, fused ? l2pol : 0);
          }
          __syncwarp();                                     // the v tile is rewritten by the next head
        } else {
          // ---- cross: q fragments from TMEM, K / V fragments from the fragment-ordered cache (permuted-k packing)
          uint32_t qf[2][16];
          tmem_ld16x256_x4(tq, qf[0]);
          tmem_ld16x256_x4(tq + 32u, qf[1]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
          if (rows_valid > 0) {
            add_q_bias<KIND>(qf[0], a.bias + h * d, q);
            add_q_bias<KIND>(qf[1], a.bias + h * d + 32, q);
            float st[G][2][4];
#pragma unroll
            for (int t = 0; t < G; ++t)
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) { st[t][nt][0] = 0.f; st[t][nt][1] = 0.f; st[t][nt][2] = 0.f; st[t][nt][3] = 0.f; }
            if (F16) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {     // k16 steps
                const int c = 2 * (ks & 1);
                auto pk = [](const uint32_t (&r)[16], int i) { return pack_f16(__uint_as_float(r[i]), __uint_as_float(r[i + 1])); };
                const uint32_t af[4] = {pk(qf[ks >> 1], 4 * c), pk(qf[ks >> 1], 4 * c + 2), pk(qf[ks >> 1], 4 * c + 4), pk(qf[ks >> 1], 4 * c + 6)};
#pragma unroll
                for (int t = 0; t < G; ++t)
#pragma unroll
                  for (int nt = 0; nt < 2; ++nt) {
                    const uint2 bb = (G == 1) ? kb[ks * 2 + nt] : __ldg(kvp[t] + (ks * 2 + nt) * 32 + lane);
                    mma_f16_16x8x16(st[t][nt], af, bb.x, bb.y);
                  }
              }
            } else {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const int c = ks & 3;
                const uint32_t af[4] = {qf[ks >> 2][4 * c], qf[ks >> 2][4 * c + 2], qf[ks >> 2][4 * c + 1], qf[ks >> 2][4 * c + 3]};
#pragma unroll
                for (int t = 0; t < G; ++t)
#pragma unroll
                  for (int nt = 0; nt < 2; ++nt) {
                    const uint2 bb = (G == 1) ? kb[(ks * 2 + nt) % NKF] : __ldg(kvp[t] + (ks * 2 + nt) * 32 + lane);
                    mma_tf32_16x8x8(st[t][nt], af, bb.x, bb.y);
                  }
              }
            }
            // one sample per tile: fetch the V fragments now, their latency hides behind the softmax
            uint2 vb[NKF];
            if (G == 1) {
#pragma unroll
              for (int f = 0; f < NKF; ++f) vb[f] = __ldg(kvp[0] + VOFF + f * 32 + lane);
            }
            // row g belongs to sample t0 = g / LQ, row g + 8 to sample t1 = (g + 8) / LQ: pick their score blocks
            const int t0 = g / LQ, t1 = (g + 8) / LQ;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              float a0 = st[0][nt][0], a1 = st[0][nt][1], b0 = st[0][nt][2], b1 = st[0][nt][3];
#pragma unroll
              for (int t = 1; t < G; ++t) {
                if (t0 == t) { a0 = st[t][nt][0]; a1 = st[t][nt][1]; }
                if (t1 == t) { b0 = st[t][nt][2]; b1 = st[t][nt][3]; }
              }
              sc[nt][0] = a0; sc[nt][1] = a1; sc[nt][2] = b0; sc[nt][3] = b1;
            }
            uint32_t pa[2][4];
            softmax_2tiles<F16>(sc, a.nk, a.scale, 0, g, q, pa);
#pragma unroll
            for (int t = 0; t < G; ++t) {
              const bool own0 = t0 == t, own1 = t1 == t;
              if (F16) {
                const uint32_t af[4] = {own0 ? pa[0][0] : 0u, own1 ? pa[0][1] : 0u, own0 ? pa[0][2] : 0u, own1 ? pa[0][3] : 0u};
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                  const uint2 bb = (G == 1) ? vb[n] : __ldg(kvp[t] + VOFF + n * 32 + lane);
                  mma_f16_16x8x16(oc[n], af, bb.x, bb.y);
                }
              } else {
#pragma unroll
                for (int kt = 0; kt < 2; ++kt) {
                  const uint32_t af[4] = {own0 ? pa[kt][0] : 0u, own1 ? pa[kt][1] : 0u, own0 ? pa[kt][2] : 0u, own1 ? pa[kt][3] : 0u};
#pragma unroll
                  for (int n = 0; n < 8; ++n) {
                    const uint2 bb = (G == 1) ? vb[(kt * 8 + n) % NKF] : __ldg(kvp[t] + VOFF + (kt * 8 + n) * 32 + lane);
                    mma_tf32_16x8x8(oc[n], af, bb.x, bb.y);
                  }
                }
              }
            }
            store_o_tiles<KIND>(oc, obuf, ob, ldo, rows_valid, g, q, fused ? l2pol : 0);
          }
        }
        if (fused && h + 2 >= heads) {
          // this thread's last head of the block: publish its head outputs to the async proxy, then tell the producer
          z_fence_proxy_async();
          mbar_arrive(&att_ready);
        }
        if (fused && h < 2 && k > 0) final_epilogue(k - 1);   // after this warp's first head of the next block (Z_LA = 2)
        if (fused && cop_ln && (h == 2 || h == 3) && k > 0) ln_pass2(k - 1);   // one head later
      }
    }
    if (fused && nk_cta > 0) {
      final_epilogue(nk_cta - 1);
      if (cop_ln) ln_pass2(nk_cta - 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace tc

static const size_t Z_SMEM_LIMIT = 232448 - 1024;

// Cout = 0: unfused (no out-projection accumulator, no out-projection stages).  tile_bytes = 128 x C operand bytes of a row block's
// activation tile: kept resident (ares_bytes = tile_bytes) when at least three weight-only stages still fit beside it, else streamed
// with every head's weight slice as before (ares_bytes = 0).  f16core: the f16 attention core needs 3 KB of v tile per warp, not 4.25.
static bool attn_frag_config(int d, int cross, int Cout, size_t tile_bytes, int f16core, int* nst, int* stage_bytes, unsigned* tmem_cols,
                             size_t* smem, int* ares_bytes) {
  static const bool no_res = getenv("MDT_NO_A_RESIDENT") != nullptr;
  const int BN = cross ? d : 3 * d;
  const size_t bj = ((size_t)BN * 128 + 1023) & ~(size_t)1023, so = Cout ? tc::Z_ABYTES + (size_t)Cout * 128 : 0;
  // warp-private v tiles (self) + the LayerNorm exchange buffer [2][128][4] (mean, M2) of the fused variant
  const size_t vt = f16core ? 64 * 24 * 2 : 16 * tc::Z_VLD * 4;
  const size_t stg = (cross ? 0 : (size_t)tc::Z_EPI_WARPS * vt) + (Cout ? 2 * tc::Z_TM * 4 * 8 : 0);
  if (Cout + 2 * BN > 512) return false;                  // two projection accumulators (one per warp group) + OUT
  unsigned cols = 32;
  while ((int)cols < Cout + 2 * BN) cols <<= 1;
  *tmem_cols = cols;
  if (!no_res && tile_bytes > 0 && tile_bytes + stg + 1024 < Z_SMEM_LIMIT) {
    const size_t stage = bj > so ? bj : so;
    int n = (int)((Z_SMEM_LIMIT - tile_bytes - stg - 1024) / stage);
    if (n > tc::Z_MAXST) n = tc::Z_MAXST;
    if (n >= 3) {
      *nst = n; *stage_bytes = (int)stage; *ares_bytes = (int)tile_bytes; *smem = tile_bytes + (size_t)n * stage + stg + 1024;
      return true;
    }
  }
  const size_t sj = tc::Z_ABYTES + bj;
  const size_t stage = sj > so ? sj : so;
  int n = (int)((Z_SMEM_LIMIT - stg - 1024) / stage);
  if (n > tc::Z_MAXST) n = tc::Z_MAXST;
  if (n < 2) return false;
  *nst = n; *stage_bytes = (int)stage; *ares_bytes = 0; *smem = (size_t)n * stage + stg + 1024;
  return true;
}

// Cout = 0 asks for the unfused variant (head outputs to the global attention tensor, out-projection elsewhere)
bool attn_frag_supported(int kind, int C, int L, int heads, int d, int cross, int Cout) {
  const int kch = kind == 1 ? 32 : 64;
  if (kind < 1 || kind > 3) return false;
  if (d != 64 || heads < 2 || (heads & 1) || C % kch) return false;
  if (!(L == 4 || L == 8 || L == 16)) return false;
  if (Cout != 0 && (Cout < 128 || Cout > 256 || Cout % 128)) return false;   // four column quarters of >= 32 columns
  int nst, sb, ar; unsigned tc_; size_t sm;
  return attn_frag_config(d, cross, Cout, 0, 0, &nst, &sb, &tc_, &sm, &ar);     // the streaming layout with the larger v tiles is the tighter fit
}

typedef void (*AttnFragKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnLayerParams,
                               const uint32_t, const uint32_t);
// mode: 0 self; 4 / 8 / 16 cross with that many query rows per sample
static AttnFragKernel attn_frag_variant(int kind, int mode, int f16) {
  // fp16 operands (kind 3) always run the f16 attention core
  static const AttnFragKernel tab[3][2][4] = {
      {{tc::attn_frag_kernel<1, 0, 0>, tc::attn_frag_kernel<1, 4, 0>, tc::attn_frag_kernel<1, 8, 0>, tc::attn_frag_kernel<1, 16, 0>},
       {tc::attn_frag_kernel<1, 0, 1>, tc::attn_frag_kernel<1, 4, 1>, tc::attn_frag_kernel<1, 8, 1>, tc::attn_frag_kernel<1, 16, 1>}},
      {{tc::attn_frag_kernel<2, 0, 0>, tc::attn_frag_kernel<2, 4, 0>, tc::attn_frag_kernel<2, 8, 0>, tc::attn_frag_kernel<2, 16, 0>},
       {tc::attn_frag_kernel<2, 0, 1>, tc::attn_frag_kernel<2, 4, 1>, tc::attn_frag_kernel<2, 8, 1>, tc::attn_frag_kernel<2, 16, 1>}},
      {{tc::attn_frag_kernel<3, 0, 1>, tc::attn_frag_kernel<3, 4, 1>, tc::attn_frag_kernel<3, 8, 1>, tc::attn_frag_kernel<3, 16, 1>},
       {tc::attn_frag_kernel<3, 0, 1>, tc::attn_frag_kernel<3, 4, 1>, tc::attn_frag_kernel<3, 8, 1>, tc::attn_frag_kernel<3, 16, 1>}}};
  return tab[kind - 1][f16 ? 1 : 0][mode == 0 ? 0 : (mode == 4 ? 1 : (mode == 8 ? 2 : 3))];
}

cudaError_t init_attn_frag() {
  for (int kind = 1; kind <= 3; ++kind)
    for (int mode : {0, 4, 8, 16})
      for (int f16 = 0; f16 < 2; ++f16) {
        cudaError_t e = cudaFuncSetAttribute(attn_frag_variant(kind, mode, f16), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Z_SMEM_LIMIT);
        if (e != cudaSuccess) return e;
      }
  return cudaSuccess;
}

cudaError_t launch_attn_frag(const void* tmA, const void* tmB, const void* tmS, const void* tmW, const AttnLayerParams& pin, int kind,
                             cudaStream_t s) {
  AttnLayerParams p = pin;
  const GemmAttnParams& a = p.a;
  if (a.M <= 0) return cudaSuccess;
  size_t smem = 0;
  if (!p.fused) p.Cout = 0;
  const size_t tile_bytes = (size_t)tc::Z_TM * a.C * (kind == 1 ? 4 : 2);
  if (!attn_frag_config(a.d, a.cross, p.Cout, tile_bytes, p.f16, &p.nst, &p.stage_bytes, &p.tmem_cols, &smem, &p.ares_bytes)) return cudaErrorInvalidValue;
  if (a.cross && !a.kvf_c) return cudaErrorInvalidValue;
  p.nacc = 2;
  p.nslot = a.heads + tc::Z_LA;
  const int BN = a.cross ? a.d : 3 * a.d;
  const uint32_t fmt = tc::umma_fmt(kind);
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(tc::Z_TM >> 4) << 24);
  const uint32_t idesc_o = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.Cout >> 3) << 17) | ((uint32_t)(tc::Z_TM >> 4) << 24);
  if (p.cop_ln && (a.cross || !p.fused || a.heads < 4 || !p.Cop || !p.C32)) return cudaErrorInvalidValue;   // second pass runs one head after the first
  const int nitems = ((a.M + tc::Z_TM - 1) / tc::Z_TM) * (p.fused ? 1 : a.heads);
  const int sms = attn_layer_sms();
  const unsigned grid = (unsigned)(nitems < sms ? nitems : sms);
  return launch_k(attn_frag_variant(kind, a.cross ? a.L : 0, p.f16), grid, tc::Z_THREADS, smem, s, *reinterpret_cast<const CUtensorMap*>(tmA),
                  *reinterpret_cast<const CUtensorMap*>(tmB), *reinterpret_cast<const CUtensorMap*>(tmS),
                  *reinterpret_cast<const CUtensorMap*>(tmW), p, idesc, idesc_o);
}

}  // namespace mdt
