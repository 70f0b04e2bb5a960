// gemm_tma.cu -- TMA-fed tcgen05 / TMEM GEMM and implicit-GEMM Conv1d (stride 1) for sm_100a.
//
//   C[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ res)        A, W already in the MMA operand dtype
//
// A is a token-major activation tensor viewed as a 3-D TMA tensor [samples][L][C]; a conv tap is a
// shifted box in the L coordinate and TMA's out-of-bounds zero fill IS the conv's zero padding at
// the sample boundaries.  A CTA owns a 128-row x BN tile:
//   warp 0     one lane issues cp.async.bulk.tensor loads (A box 128 x 128 B, W box BN x 128 B) per stage
//   warp 1     one lane issues tcgen05.mma (accumulator in TMEM), tcgen05.commit frees stages; owns TMEM
//   warps 2-5  epilogue: tcgen05.ld -> smem staging -> coalesced bias / GELU / residual, fp32 and/or
//              operand-dtype stores
// Persistent, one CTA per SM: 4-stage smem ring, 16 epilogue warps, two TMEM accumulators so the epilogue of tile i overlaps tile i+1.
#include <cuda.h>
#include <stdlib.h>
#include <cuda_bf16.h>
#include "aload.cuh"
#include "tc_common.cuh"

namespace mdt {
namespace tc {

constexpr int T_TM = 128;
constexpr int T_STAGES = 4;
constexpr int T_A_BYTES = T_TM * 128;
constexpr int T_B_BYTES = 128 * 128;
constexpr int T_STAGE_BYTES = T_A_BYTES + T_B_BYTES;
constexpr int T_EPI_WARPS = 16;
constexpr int T_STG_LD = 36;                                  // 32 columns + 4 floats of padding per staged row
constexpr int T_STG_BYTES = T_EPI_WARPS * 32 * T_STG_LD * 4;  // per-warp private staging, 4.5 KB each
// LayerNorm epilogue (cop_ln): per-row (mean, M2) of each of the four column-group warps, double buffered by tile parity
constexpr int T_XCH_BYTES = 2 * T_TM * 4 * 8;
constexpr int T_SMEM_BYTES = T_STAGES * T_STAGE_BYTES + T_STG_BYTES + T_XCH_BYTES + 1024;
// BN = 256 (one n-tile for the 256-wide levels: the activation tile is streamed once per row block instead of twice, -25 % L2 -> SM
// operand bytes): 48 KB stages, three of them
constexpr int T_STAGES_WIDE = 3;
constexpr int T_STAGE_BYTES_WIDE = T_A_BYTES + 256 * 128;
constexpr int T_SMEM_BYTES_WIDE = T_STAGES_WIDE * T_STAGE_BYTES_WIDE + T_STG_BYTES + T_XCH_BYTES + 1024;
constexpr int T_THREADS = 64 + 32 * T_EPI_WARPS;

// Persistent kernel: every CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... with the N tile index
// fastest, so concurrently running CTAs share A tiles (and the whole weight matrix) in L2.  Three pipelines:
//   smem ring  (TMA producer  <-> MMA issuer)        full_bar / empty_bar       [T_STAGES]
//   TMEM ring  (MMA issuer    <-> epilogue warps)    acc_full / acc_empty       [2 accumulators of BN columns]
//   tile loop  (all roles derive the same tile sequence from blockIdx / gridDim)
// LN: the LayerNorm epilogue variant (its own instantiation: the register copy of the row block must not cost the plain GEMMs anything)
template <int KIND, bool LN>
__global__ void __launch_bounds__(T_THREADS, 1) gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const TmaGemmParams p, const uint32_t idesc) {
  constexpr int KCH = (KIND == 1) ? 32 : 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // SWIZZLE_128B operand tiles need 1024-byte alignment
  __shared__ __align__(8) uint64_t full_bar[T_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[T_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  // keep the pointer in the shared address space (an integer round trip would demote every access to generic LD/ST)
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.BN;
  const int n_tiles = p.N / BN;
  const int m_tiles = (p.M + T_TM - 1) / T_TM;
  const int total_tiles = m_tiles * n_tiles;
  const int num_chunks = p.taps * p.kchunks;
  const int NST = BN > 128 ? T_STAGES_WIDE : T_STAGES;
  const int stage_bytes = BN > 128 ? T_STAGE_BYTES_WIDE : T_STAGE_BYTES;
  // two accumulators; narrow tiles still read 32 columns per tcgen05.ld, so keep 32 columns of slack
  const uint32_t tmem_cols = BN <= 32 ? 64u : (BN <= 64 ? 128u : (BN <= 128 ? 256u : 512u));

  if (tid == 0) {
    for (int s = 0; s < T_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], T_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) tmem_alloc(&tmem_base_s, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_enter();                                  // set-up above overlaps the previous grid's tail (launch.cuh)
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t tx = (uint32_t)(T_A_BYTES + BN * 128);
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int te = p.rev ? total_tiles - 1 - t : t;
        const int mt = te / n_tiles, nt = te - mt * n_tiles;
        int b0, l0;
        if (p.L >= T_TM) { const int tps = p.L / T_TM; b0 = mt / tps; l0 = (mt - b0 * tps) * T_TM; }
        else { b0 = mt * p.Sb; l0 = 0; }
        for (int tap = 0; tap < p.taps; ++tap) {
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * stage_bytes;
            mbar_arrive_expect_tx(&full_bar[stage], tx);
            tma_load_3d(sa, &tmA, &full_bar[stage], kc * KCH, l0 + tap - p.pad, b0);
            tma_load_2d(sa + T_A_BYTES, &tmB, &full_bar[stage], tap * p.C + kc * KCH, nt * BN);
            if (++stage == NST) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int it = 0;
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&acc_empty[buf], aphase ^ 1u);      // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
      for (int k0 = 0; k0 < num_chunks; ++k0) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + T_A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma<KIND>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((k0 | k) != 0));
          umma_commit(&empty_bar[stage]);
          if (k0 == num_chunks - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int ew = warp - 2;             // 0..15
    const int q = warp & 3;              // TMEM lane quadrant this warp may access
    const int half = ew >> 2;            // column group (0..3) handled by this warp
    float* stg = reinterpret_cast<float*>(smem + NST * stage_bytes) + (size_t)ew * 32 * T_STG_LD;
    // BN = 128: four groups of 32 columns; BN = 64: two groups; narrower tiles: one group
    const int ngroups = BN >= 128 ? 4 : (BN >= 64 ? 2 : 1);
    const int cols_per_half = BN / ngroups;
    const bool active = half < ngroups;
    float2* xch = reinterpret_cast<float2*>(smem + NST * stage_bytes + T_STG_BYTES);       // [2][128 rows][4 column groups]
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int te = p.rev ? total_tiles - 1 - t : t;
      const int mt = te / n_tiles, nt = te - mt * n_tiles;
      const int buf = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      if (LN) {
        // ---- LayerNorm epilogue (launcher guarantees N == BN in {128, 256}: a row block holds whole rows, four warps per quadrant
        // own 32 or 64 columns each).  C32 = acc + bias (+ res); Cop = (C32 - mean) * rstd, no affine (folded into the consumer's
        // weights).  Exact mean / centred M2 per 32-column chunk inside the eight lanes that hold it, equal-count Chan merges across
        // chunks and across the four warps (one named barrier per tile).
        const int nch = cols_per_half >> 5;          // 1 or 2 chunks of 32 columns per warp
        const int cl = (lane & 7) * 4, r0 = lane >> 3;
        const int m0 = mt * T_TM + q * 32;
        float4 r[8];
        auto load_res = [&](int u) {
          const int no = half * cols_per_half + u * 32 + cl;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int mo = m0 + r0 + i * 4;
            r[i] = (p.res && mo < p.M) ? *reinterpret_cast<const float4*>(p.res + (size_t)mo * p.ldres + no) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        load_res(0);                                 // ahead of the accumulator wait: the residual's latency hides behind it
        mbar_wait(&acc_full[buf], aphase);
        tc_fence_after();
        float mean_w[8], m2_w[8];                    // per row of this lane: statistics over this warp's columns
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u < nch) {
            const int col0 = half * cols_per_half + u * 32;
            {
              uint32_t v[32];
              tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + col0), v);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(stg + lane * T_STG_LD + j * 4) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            __syncwarp();
            if (u > 0) load_res(u);      // after the staging registers are dead (register budget: 576 threads = 96 per thread)
            const int no = col0 + cl;
            const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + no)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int mo = m0 + r0 + i * 4;
              float4 o = *reinterpret_cast<const float4*>(stg + (r0 + i * 4) * T_STG_LD + cl);
              o.x += bv.x + r[i].x; o.y += bv.y + r[i].y; o.z += bv.z + r[i].z; o.w += bv.w + r[i].w;
              if (mo < p.M) *reinterpret_cast<float4*>(p.C32 + (size_t)mo * p.ldc + no) = o;
              // exact mean / centred M2 of this 32-column chunk inside the eight lanes that hold the row
              float sm = (o.x + o.y) + (o.z + o.w);
              sm += __shfl_xor_sync(0xffffffffu, sm, 1); sm += __shfl_xor_sync(0xffffffffu, sm, 2); sm += __shfl_xor_sync(0xffffffffu, sm, 4);
              const float mean = sm * (1.0f / 32.0f);
              const float dx = o.x - mean, dy = o.y - mean, dz = o.z - mean, dw = o.w - mean;
              float m2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
              m2 += __shfl_xor_sync(0xffffffffu, m2, 1); m2 += __shfl_xor_sync(0xffffffffu, m2, 2); m2 += __shfl_xor_sync(0xffffffffu, m2, 4);
              if (u == 0) { mean_w[i] = mean; m2_w[i] = m2; }
              else { const float dm = mean - mean_w[i]; mean_w[i] = 0.5f * (mean_w[i] + mean); m2_w[i] = m2_w[i] + m2 + 16.0f * dm * dm; }
            }
            __syncwarp();
          }
        }
        // the accumulator has been consumed: hand it back before the statistics exchange
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        float2* xb = xch + (size_t)(it & 1) * T_TM * 4;
        if ((lane & 7) == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) xb[(q * 32 + r0 + i * 4) * 4 + half] = make_float2(mean_w[i], m2_w[i]);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");      // the four column-group warps of this quadrant
        // second pass: the values come back from C32 (this thread's own stores, L2-resident) instead of living in 64 registers
        const float inv_n = 1.0f / (float)p.N, n_w = (float)cols_per_half;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u < nch) {
            const int no = half * cols_per_half + u * 32 + cl;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int mo = m0 + r0 + i * 4;
              r[i] = mo < p.M ? *reinterpret_cast<const float4*>(p.C32 + (size_t)mo * p.ldc + no) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 e01 = *reinterpret_cast<const float4*>(xb + (q * 32 + r0 + i * 4) * 4);
              const float4 e23 = *reinterpret_cast<const float4*>(xb + (q * 32 + r0 + i * 4) * 4 + 2);
              const float mean = 0.25f * ((e01.x + e01.z) + (e23.x + e23.z));
              const float d0 = e01.x - mean, d1 = e01.z - mean, d2 = e23.x - mean, d3 = e23.z - mean;
              const float dev = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
              const float rstd = rsqrtf(fmaf(n_w, dev, (e01.y + e01.w) + (e23.y + e23.w)) * inv_n + p.ln_eps);
              const int mo = m0 + r0 + i * 4;
              if (mo < p.M) {
                const float4 o = r[i];
                const float nx = (o.x - mean) * rstd, ny = (o.y - mean) * rstd, nz = (o.z - mean) * rstd, nw = (o.w - mean) * rstd;
                const size_t off = (size_t)mo * p.ldcop + no;
                if (KIND == 1) *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.Cop) + off) = make_uint4(to_tf32(nx), to_tf32(ny), to_tf32(nz), to_tf32(nw));
                else *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.Cop) + off) = make_uint2(pack_op2<KIND>(nx, ny), pack_op2<KIND>(nz, nw));
              }
            }
          }
        }
        // no second barrier: the next tile uses the other exchange buffer, and that tile's barrier orders the reuse of this one
        continue;
      }
      mbar_wait(&acc_full[buf], aphase);
      tc_fence_after();
      if (active) {
        const int col0 = half * cols_per_half;
        for (int cc = 0; cc < cols_per_half; cc += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + col0 + cc), v);
          const int ncol = min(32, cols_per_half - cc);
          if (p.gn_L > 0) {
            // ---- fused GroupNorm of (acc + bias): statistics over gn_L rows (lanes) x gn_cpg columns, all inside this block
            const int nb0 = nt * BN + col0 + cc;
            float x[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + nb0 + j * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
              x[4 * j] = __uint_as_float(v[4 * j]) + bv.x; x[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bv.y;
              x[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bv.z; x[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bv.w;
            }
            const float inv_n = 1.0f / (float)(p.gn_L * p.gn_cpg);
            if (p.gn_cpg == 32) {
              float s = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) s += x[j];
              for (int o = p.gn_L >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
              const float mean = s * inv_n;
              float sq = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) { const float d = x[j] - mean; sq = fmaf(d, d, sq); }
              for (int o = p.gn_L >> 1; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
              const float rstd = 1.0f / sqrtf(sq * inv_n + p.gn_eps);
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = (x[j] - mean) * rstd;
            } else {  // gn_cpg == 16: two groups per 32-column block
              float s0 = 0.f, s1 = 0.f;
#pragma unroll
              for (int j = 0; j < 16; ++j) { s0 += x[j]; s1 += x[16 + j]; }
              for (int o = p.gn_L >> 1; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
              const float m0 = s0 * inv_n, m1 = s1 * inv_n;
              float q0 = 0.f, q1 = 0.f;
#pragma unroll
              for (int j = 0; j < 16; ++j) { const float d0 = x[j] - m0, d1 = x[16 + j] - m1; q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); }
              for (int o = p.gn_L >> 1; o > 0; o >>= 1) { q0 += __shfl_xor_sync(0xffffffffu, q0, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o); }
              const float r0 = 1.0f / sqrtf(q0 * inv_n + p.gn_eps), r1 = 1.0f / sqrtf(q1 * inv_n + p.gn_eps);
#pragma unroll
              for (int j = 0; j < 16; ++j) { x[j] = (x[j] - m0) * r0; x[16 + j] = (x[16 + j] - m1) * r1; }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(stg + lane * T_STG_LD + j * 4) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            __syncwarp();
            const int cl = (lane & 7) * 4;
            const float* aff = p.gn_aff + (size_t)(p.gn_call ? *p.gn_call : 0) * p.gn_aff_stride;
            const float4 ga = __ldg(reinterpret_cast<const float4*>(aff + nb0 + cl));
            const float4 gb = __ldg(reinterpret_cast<const float4*>(aff + p.N + nb0 + cl));
#pragma unroll
            for (int rr = lane >> 3; rr < 32; rr += 4) {
              const int mo = mt * T_TM + q * 32 + rr, no = nb0 + cl;
              if (mo < p.M) {
                float4 o = *reinterpret_cast<const float4*>(stg + rr * T_STG_LD + cl);
                o.x = fmaf(o.x, ga.x, gb.x); o.y = fmaf(o.y, ga.y, gb.y); o.z = fmaf(o.z, ga.z, gb.z); o.w = fmaf(o.w, ga.w, gb.w);
                o.x = o.x * __fdividef(1.0f, 1.0f + __expf(-o.x)); o.y = o.y * __fdividef(1.0f, 1.0f + __expf(-o.y));
                o.z = o.z * __fdividef(1.0f, 1.0f + __expf(-o.z)); o.w = o.w * __fdividef(1.0f, 1.0f + __expf(-o.w));
                if (KIND == 1)
                  *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.Cop) + (size_t)mo * p.ldcop + no) =
                      make_uint4(to_tf32(o.x), to_tf32(o.y), to_tf32(o.z), to_tf32(o.w));
                else
                  *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.Cop) + (size_t)mo * p.ldcop + no) =
                      make_uint2(pack_op2<KIND>(o.x, o.y), pack_op2<KIND>(o.z, o.w));
              }
            }
            __syncwarp();
            continue;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * T_STG_LD + j * 4) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          // coalesced: 8 lanes cover one 128-byte row segment, 4 rows per pass
          const int cl = (lane & 7) * 4;
#pragma unroll
          for (int rr = lane >> 3; rr < 32; rr += 4) {
            const int mo = mt * T_TM + q * 32 + rr, no = nt * BN + col0 + cc + cl;
            if (mo < p.M && cl < ncol) {
              float4 o = *reinterpret_cast<const float4*>(stg + rr * T_STG_LD + cl);
              if (p.bias) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + no));
                o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
              }
              if (p.act == 1) { o.x = gelu_as(o.x); o.y = gelu_as(o.y); o.z = gelu_as(o.z); o.w = gelu_as(o.w); }
              if (p.res) {
                const float4 rv = *reinterpret_cast<const float4*>(p.res + (size_t)mo * p.ldres + no);
                o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
              }
              if (p.C32) *reinterpret_cast<float4*>(p.C32 + (size_t)mo * p.ldc + no) = o;
              if (p.Cop) {
                if (KIND == 1) {
                  *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.Cop) + (size_t)mo * p.ldcop + no) =
                      make_uint4(to_tf32(o.x), to_tf32(o.y), to_tf32(o.z), to_tf32(o.w));
                } else {
                  *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.Cop) + (size_t)mo * p.ldcop + no) =
                      make_uint2(pack_op2<KIND>(o.x, o.y), pack_op2<KIND>(o.z, o.w));
                }
              }
            }
          }
          __syncwarp();
        }
      }
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld inside tmem_ld32): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace tc

int tma_pick_bn(int N) {
  static const bool wide_off = getenv("MDT_NO_WIDE_BN") != nullptr;
  if (N % 256 == 0 && !wide_off) return 256;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  if (N % 32 == 0) return 32;
  if (N % 16 == 0) return 16;
  return 0;
}

bool gemm_tma_shape_ok(int kind, int C, int L, int N) {
  const int kch = kind == 1 ? 32 : 64;
  if (C % kch) return false;
  if (tma_pick_bn(N) == 0) return false;
  if (L <= 0 || (L & (L - 1))) return false;   // power of two: 128 % L == 0 or L % 128 == 0
  return true;
}

int make_tmap_act(void* map128, const void* base, int kind, int C, int L, long long samples) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) return -1;
  const int esz = kind == 1 ? 4 : 2, kch = kind == 1 ? 32 : 64;
  const CUtensorMapDataType dt = kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (kind == 3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  const int Lb = L >= 128 ? 128 : L, Sb = L >= 128 ? 1 : 128 / L;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)samples};
  cuuint64_t strides[2] = {(cuuint64_t)C * esz, (cuuint64_t)C * L * esz};
  cuuint32_t box[3] = {(cuuint32_t)kch, (cuuint32_t)Lb, (cuuint32_t)Sb};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(map128), dt, 3,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

int make_tmap_weight(void* map128, const void* base, int kind, long long K, int N, int BN) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) return -1;
  const int esz = kind == 1 ? 4 : 2, kch = kind == 1 ? 32 : 64;
  const CUtensorMapDataType dt = kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (kind == 3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)K * esz};
  cuuint32_t box[2] = {(cuuint32_t)kch, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(map128), dt, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

cudaError_t init_gemm_tma() {
  const int mx = tc::T_SMEM_BYTES_WIDE > tc::T_SMEM_BYTES ? tc::T_SMEM_BYTES_WIDE : tc::T_SMEM_BYTES;
  cudaError_t e = cudaFuncSetAttribute(tc::gemm_tma_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tma_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tma_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tma_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tma_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tma_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  return e;
}

static int g_num_sms = 0;

cudaError_t launch_gemm_tma(const void* tmA, const void* tmB, const TmaGemmParams& p, int kind, cudaStream_t s) {
  if (p.M <= 0 || p.N <= 0) return cudaSuccess;
  if (p.BN <= 0 || p.N % p.BN) return cudaErrorInvalidValue;
  if (p.cop_ln && (p.N != p.BN || p.BN < 128 || !p.Cop || !p.C32 || p.gn_L > 0 || p.act != 0)) return cudaErrorInvalidValue;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  const uint32_t fmt = tc::umma_fmt(kind);
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(tc::T_TM >> 4) << 24);
  const long long tiles = (long long)((p.M + tc::T_TM - 1) / tc::T_TM) * (p.N / p.BN);
  const unsigned grid = (unsigned)(tiles < g_num_sms ? tiles : g_num_sms);   // persistent: one CTA per SM
  const CUtensorMap& a = *reinterpret_cast<const CUtensorMap*>(tmA);
  const CUtensorMap& b = *reinterpret_cast<const CUtensorMap*>(tmB);
  const int smem = p.BN > 128 ? tc::T_SMEM_BYTES_WIDE : tc::T_SMEM_BYTES;
  if (p.cop_ln) {
    return launch_k(kind == 1 ? tc::gemm_tma_kernel<1, true> : (kind == 2 ? tc::gemm_tma_kernel<2, true> : tc::gemm_tma_kernel<3, true>), grid,
                    tc::T_THREADS, smem, s, a, b, p, idesc);
  }
  return launch_k(kind == 1 ? tc::gemm_tma_kernel<1, false> : (kind == 2 ? tc::gemm_tma_kernel<2, false> : tc::gemm_tma_kernel<3, false>), grid,
                  tc::T_THREADS, smem, s, a, b, p, idesc);
}

}  // namespace mdt
