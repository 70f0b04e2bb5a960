// gemm_attn_umma.cu -- fused per-head projection + attention for sm_100a with the attention core on tcgen05 as well.
//
// Same contract as gemm_attn.cu (tile = 128 activation rows x one head; the [q | k | v] projection accumulates in TMEM and
// only the (rows x d) head output goes to HBM), but softmax(q k^T) v no longer runs per sample on mma.sync.  A 128-row tile
// holds 128 / L whole samples, so one block-diagonal pair of UMMAs serves all of them at once:
//
//   S[128, 128] = Q[128, 64] K[128, 64]^T       A, B from shared memory (K-major SWIZZLE_128B tiles staged by the epilogue warps)
//   P           = softmax over the L columns of the row's own sample, everything else 0     (thread-local: TMEM lane = row)
//   O[128, 64]  = P[128, 128] V[128, 64]        A = P read from TMEM (written back over S with tcgen05.st), B = V^T from smem
//
// The off-diagonal blocks cost 8x the useful flops and are still ~1 k cycles per tile; what they buy is that the per-row softmax
// needs no shuffles and no fragment shuffling through shared memory, which is what bounded the mma.sync version (latency, not math).
// Cross-attention (q-only projection, K / V^T of the conditioning copied from the per-sample cache) uses the same core when the
// context fits the L key slots a sample owns in the tile (n_ctx <= L).
//
// Warp roles (320 threads, one persistent CTA per SM):
//   warp 0      TMA producer for the projection operands (2-stage ring)
//   warp 1      UMMA issuer: projection of tile i+1, then S and P V of tile i; owns TMEM (512 columns)
//   warps 2-9   TMEM quadrant q = warp % 4, column half = (warp - 2) / 4: stage q/k/v^T, softmax (half 0), read O, store
// TMEM columns: [0, 2 BN) two projection accumulators, [384, 512) S / P; O overwrites the first 64 columns of the projection
// accumulator it came from (self) or has its own 64 columns (cross).
// -DMDT_ATTN_DEBUG adds clock64 stamps around the epilogue phases of CTA 0 / warp 2 and prints cycles per tile after every launch.
#include <cuda.h>
#include <cuda_bf16.h>
#include "kernels.cuh"
#include "tc_common.cuh"
#include <stdio.h>
#include <string.h>

namespace mdt {
namespace tc {

#ifdef MDT_ATTN_DEBUG
__device__ unsigned long long g_dbg[16];
#define DBG_T(i) const long long dbg_t##i = clock64();
#else
#define DBG_T(i)
#endif

constexpr int U_TM = 128;
constexpr int U_STAGES = 3;
constexpr int U_ABYTES = U_TM * 128;
constexpr int U_EPI_WARPS = 8;
constexpr int U_THREADS = 64 + 32 * U_EPI_WARPS;
constexpr int U_QBYTES = 2 * U_TM * 128;    // q (or k): two K-chunk tiles of 128 rows x 32 tf32
constexpr int U_VBYTES = 4 * 64 * 128;      // v^T: four K-chunk tiles (32 keys each) of 64 feature rows
constexpr uint32_t U_S_COL = 384;           // S / P columns
constexpr uint32_t U_O_COL_CROSS = 128;     // cross: O columns (after the two 64-column q accumulators)

template <int KIND>
__global__ void __launch_bounds__(U_THREADS, 1) gemm_attn_umma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                     const __grid_constant__ CUtensorMap tmB,
                                                                     const GemmAttnParams p, const uint32_t idesc) {
  constexpr int KCH = (KIND == 1) ? 32 : 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[U_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[U_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ __align__(8) uint64_t staged_bar, s_full, p_ready, o_full;
  __shared__ uint32_t tmem_base_s;

  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.cross ? 64 : 192;
  const int b_bytes = BN * 128;
  const int stage_bytes = U_ABYTES + ((b_bytes + 1023) & ~1023);
  const int m_tiles = (p.M + U_TM - 1) / U_TM;
  const int total_tiles = m_tiles * p.heads;
  uint8_t* Qs = smem + U_STAGES * stage_bytes;
  uint8_t* Ks = Qs + U_QBYTES;
  uint8_t* Vt = Ks + U_QBYTES;

  if (tid == 0) {
    for (int s = 0; s < U_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], U_EPI_WARPS); }
    mbar_init(&staged_bar, U_EPI_WARPS); mbar_init(&s_full, 1); mbar_init(&p_ready, 4); mbar_init(&o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_s = tmem_base + U_S_COL;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (projection operands)
    if (lane == 0) {
      const uint32_t tx = (uint32_t)(U_ABYTES + b_bytes);
      int c = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int te = p.rev ? total_tiles - 1 - t : t;
        const int mt = te / p.heads, h = te - mt * p.heads;
        for (int kc = 0; kc < p.kchunks; ++kc, ++c) {
          const int stage = c % U_STAGES;
          const uint32_t phase = (uint32_t)(c / U_STAGES) & 1u;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          tma_load_3d(sa, &tmA, &full_bar[stage], kc * KCH, 0, mt * p.Sb);
          tma_load_2d(sa + U_ABYTES, &tmB, &full_bar[stage], kc * KCH, h * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer
    // attention-core instruction descriptors: tf32 x tf32 -> f32, K-major both, M = 128, N = 128 (S) / 64 (O)
    const uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // One thread multiplexes three in-order streams and issues whichever has its inputs ready, attention first (it is on the
    // epilogue warps' critical path; projection chunks of the next tile fill the gaps):
    //   S(i)   needs staged(i)        P V(i)  needs p_ready(i)        projection chunk needs its smem stage (+ a free accumulator)
    if (lane == 0) {
      const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      int pj = 0, pk = 0, c = 0;      // projection: tile, chunk within the tile, ring counter
      bool pacc = false;              // accumulator of tile pj acquired
      int at = 0, aph = 0;            // attention: tile, phase (0: S pending, 1: P V pending)
      const uint32_t qa = smem_u32(Qs), ka = smem_u32(Ks), va = smem_u32(Vt);
      while (at < my_tiles) {
        bool did = false;
        if (aph == 0) {
          if (mbar_test(&staged_bar, (uint32_t)at & 1u)) {
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma<1>(tmem_s, make_desc(qa + (kk >> 2) * (U_TM * 128)) + (uint64_t)(2 * (kk & 3)),
                      make_desc(ka + (kk >> 2) * (U_TM * 128)) + (uint64_t)(2 * (kk & 3)), idesc_s, (uint32_t)(kk != 0));
            umma_commit(&s_full);
            aph = 1; did = true;
          }
        } else if (mbar_test(&p_ready, (uint32_t)at & 1u)) {
          tc_fence_after();
          const uint32_t tmem_o = tmem_base + (p.cross ? U_O_COL_CROSS : (uint32_t)((at & 1) * BN));
#pragma unroll
          for (int kk = 0; kk < 16; ++kk)
            umma_ts_tf32(tmem_o, tmem_s + (uint32_t)(8 * kk), make_desc(va + (kk >> 2) * (64 * 128)) + (uint64_t)(2 * (kk & 3)), idesc_o,
                         (uint32_t)(kk != 0));
          umma_commit(&o_full);
          aph = 0; ++at; did = true;
        }
        if (pj < my_tiles) {
          const int buf = pj & 1;
          if (!pacc && mbar_test(&acc_empty[buf], ((uint32_t)(pj >> 1) & 1u) ^ 1u)) { tc_fence_after(); pacc = true; }
          if (pacc) {
            const int stage = c % U_STAGES;
            if (mbar_test(&full_bar[stage], (uint32_t)(c / U_STAGES) & 1u)) {
              tc_fence_after();
              const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
              const uint32_t sa = smem_u32(smem + stage * stage_bytes);
              const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + U_ABYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma<KIND>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((pk | k) != 0));
              umma_commit(&empty_bar[stage]);
              ++c;
              if (++pk == p.kchunks) { umma_commit(&acc_full[buf]); pk = 0; ++pj; pacc = false; }
              did = true;
            }
          }
        }
        (void)did;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ staging, softmax, output (warps 2..9)
    const int ew = warp - 2;
    const int q = warp & 3;                 // TMEM lane quadrant
    const int half = ew >> 2;               // which 32 of the 64 head features this warp moves
    const int row = q * 32 + lane;          // tile row = TMEM lane
    const int L = p.L;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t swz = (uint32_t)(lane & 7);
    uint8_t* qrow = Qs + half * (U_TM * 128) + row * 128;     // this lane's 128-byte row in the q K-chunk tile `half`
    uint8_t* krow = Ks + half * (U_TM * 128) + row * 128;
    // v^T: K-chunk tile q holds keys 32q .. 32q + 31; feature row f, this lane's key = column `lane`
    uint8_t* vcol = Vt + q * (64 * 128) + (lane & 3) * 4;
    float* ostg = reinterpret_cast<float*>(Qs + half * (U_TM * 128) + q * 32 * 128);   // private: the q rows this warp staged
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int te = p.rev ? total_tiles - 1 - t : t;
      const int mt = te / p.heads, h = te - mt * p.heads;
      const int m0 = mt * U_TM;
      const int buf = it & 1;
      const uint32_t par = (uint32_t)it & 1u;
      DBG_T(0)
      mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      DBG_T(1)
      const uint32_t acc = tmem_base + lane_addr + (uint32_t)(buf * BN + half * 32);
      uint32_t v[32];
      // Rows past M are in bounds for the TMA box (the activation map covers the plan's maximum batch) and hold stale data.
      // Per-sample kernels never mix them with valid rows; here P = 0 times a stale NaN in V would poison O, so they are zeroed.
      const bool row_ok = m0 + row < p.M;
      // ---- q (+ folded bias; the k bias cancels in the softmax, the v bias is folded into the out-projection bias)
      tmem_ld32(acc, v);
      {
        const float* bias = p.bias + h * p.d + half * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + j * 4));
          *reinterpret_cast<uint4*>(qrow + (((uint32_t)j ^ swz) << 4)) =
              row_ok ? make_uint4(to_tf32(__uint_as_float(v[4 * j]) + bv.x), to_tf32(__uint_as_float(v[4 * j + 1]) + bv.y),
                                  to_tf32(__uint_as_float(v[4 * j + 2]) + bv.z), to_tf32(__uint_as_float(v[4 * j + 3]) + bv.w))
                     : make_uint4(0u, 0u, 0u, 0u);
        }
      }
      if (!p.cross) {
        tmem_ld32(acc + 64u, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(krow + (((uint32_t)j ^ swz) << 4)) =
              row_ok ? make_uint4(to_tf32(__uint_as_float(v[4 * j])), to_tf32(__uint_as_float(v[4 * j + 1])),
                                  to_tf32(__uint_as_float(v[4 * j + 2])), to_tf32(__uint_as_float(v[4 * j + 3])))
                     : make_uint4(0u, 0u, 0u, 0u);
        tmem_ld32(acc + 128u, v);
#pragma unroll
        for (int f = 0; f < 32; ++f) {
          const uint32_t fg = (uint32_t)(half * 32 + f);
          *reinterpret_cast<uint32_t*>(vcol + fg * 128 + ((((uint32_t)lane >> 2) ^ (fg & 7u)) << 4)) = row_ok ? to_tf32(__uint_as_float(v[f])) : 0u;
        }
      }
      DBG_T(2)
      fence_proxy_async();     // generic-proxy smem writes -> visible to the UMMA (async proxy) reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&staged_bar);
      if (half == 0) {
        // ---- softmax of this quadrant's 32 rows: row r keeps the L columns of its own sample inside the quadrant's 32
        mbar_wait(&s_full, par);
        tc_fence_after();
        DBG_T(3)
        tmem_ld32(tmem_s + lane_addr + (uint32_t)(q * 32), v);
        const uint32_t blk = (uint32_t)lane & ~(uint32_t)(L - 1);   // first column of the lane's sample
        const int nk = p.cross ? p.nk : L;
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const bool ok = ((uint32_t)c & ~(uint32_t)(L - 1)) == blk && (c & (L - 1)) < nk;
          const float sv = ok ? __uint_as_float(v[c]) * p.scale : -INFINITY;
          v[c] = __float_as_uint(sv);
          mx = fmaxf(mx, sv);
        }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float e = __expf(__uint_as_float(v[c]) - mx);   // masked columns: exp(-inf) = 0
          v[c] = __float_as_uint(e);
          sum += e;
        }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = to_tf32(__uint_as_float(v[c]) * inv);
        uint32_t z[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) z[c] = 0u;
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) {
          if (jb == q) tmem_st32(tmem_s + lane_addr + (uint32_t)(jb * 32), v);
          else tmem_st32(tmem_s + lane_addr + (uint32_t)(jb * 32), z);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready);
        DBG_T(4)
#ifdef MDT_ATTN_DEBUG
        if (blockIdx.x == 0 && warp == 2 && lane == 0) { atomicAdd(&g_dbg[2], (unsigned long long)(dbg_t3 - dbg_t2)); atomicAdd(&g_dbg[3], (unsigned long long)(dbg_t4 - dbg_t3)); g_dbg[8] = (unsigned long long)dbg_t4; }
#endif
      }
      // ---- O: 32 of the 64 head features per warp
      mbar_wait(&o_full, par);
      tc_fence_after();
      DBG_T(5)
      tmem_ld32(tmem_base + lane_addr + (p.cross ? U_O_COL_CROSS : (uint32_t)(buf * BN)) + (uint32_t)(half * 32), v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);   // accumulator (and the O columns inside it) may be overwritten
      DBG_T(6)
      // coalesced store through the warp's private staging rows (the q rows it wrote; S = Q K^T has completed)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(ostg) + lane * 128 + (((uint32_t)j ^ swz) << 4)) =
            make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      __syncwarp();
      const int cl = lane & 7;
#pragma unroll
      for (int rr = lane >> 3; rr < 32; rr += 4) {
        const int mo = m0 + q * 32 + rr;
        if (mo < p.M) {
          const float4 o = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(ostg) + rr * 128 + ((cl ^ (rr & 7)) << 4));
          const size_t off = (size_t)mo * p.ldo + (size_t)h * p.d + half * 32 + cl * 4;
          if (KIND == 1)
            *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.att) + off) = make_uint4(to_tf32(o.x), to_tf32(o.y), to_tf32(o.z), to_tf32(o.w));
          else
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.att) + off) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
        }
      }
      __syncwarp();
#ifdef MDT_ATTN_DEBUG
      if (blockIdx.x == 0 && warp == 2 && lane == 0) {
        const long long dbg_t7 = clock64();
        atomicAdd(&g_dbg[0], (unsigned long long)(dbg_t1 - dbg_t0)); atomicAdd(&g_dbg[1], (unsigned long long)(dbg_t2 - dbg_t1));
        atomicAdd(&g_dbg[4], (unsigned long long)(dbg_t5 - (long long)g_dbg[8])); atomicAdd(&g_dbg[5], (unsigned long long)(dbg_t6 - dbg_t5));
        atomicAdd(&g_dbg[6], (unsigned long long)(dbg_t7 - dbg_t6)); atomicAdd(&g_dbg[7], 1ull);
      }
#endif
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

}  // namespace tc

static size_t gemm_attn_umma_smem(int cross) {
  const int BN = cross ? 64 : 192;
  const size_t stage = tc::U_ABYTES + (((size_t)BN * 128 + 1023) & ~(size_t)1023);
  return tc::U_STAGES * stage + 2 * tc::U_QBYTES + tc::U_VBYTES;
}

bool gemm_attn_umma_supported(int kind, int C, int L, int heads, int d, int cross, int nk_max) {
  const int kch = kind == 1 ? 32 : 64;
  if (d != 64 || C % kch || L < 1 || L > 32 || (L & (L - 1))) return false;
  if (kind == 3) return false;   // opt-in experiment: tf32 / bf16 operands only
  if (cross) return false;   // cross-attention stays on gemm_attn.cu (n_ctx key slots per sample would have to fit in L columns)
  (void)heads; (void)nk_max;
  return true;
}

cudaError_t init_gemm_attn_umma() {
  cudaError_t e = cudaFuncSetAttribute(tc::gemm_attn_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(tc::gemm_attn_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
}

static int g_sms_gu = 0;

#ifdef MDT_ATTN_DEBUG
void gemm_attn_umma_debug_dump() {
  unsigned long long h[16];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, tc::g_dbg, sizeof(h));
  const double n = h[7] ? (double)h[7] : 1.0;
  fprintf(stderr, "[attn dbg] tiles=%llu  wait_acc=%.0f stage=%.0f wait_S=%.0f softmax=%.0f wait_O=%.0f ld_O=%.0f store=%.0f (cycles / tile)\n", h[7],
          h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n);
  memset(h, 0, sizeof(h));
  cudaMemcpyToSymbol(tc::g_dbg, h, sizeof(h));
}
#endif

cudaError_t launch_gemm_attn_umma(const void* tmA, const void* tmB, const GemmAttnParams& p, int kind, cudaStream_t s) {
  if (p.M <= 0) return cudaSuccess;
  if (g_sms_gu == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_gu, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms_gu <= 0) g_sms_gu = 148;
  }
  const int BN = p.cross ? 64 : 192;
  const size_t smem = gemm_attn_umma_smem(p.cross);
  const uint32_t fmt = kind == 1 ? 2u : 1u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(tc::U_TM >> 4) << 24);
  const long long tiles = (long long)((p.M + tc::U_TM - 1) / tc::U_TM) * p.heads;
  const unsigned grid = (unsigned)(tiles < g_sms_gu ? tiles : g_sms_gu);
  const CUtensorMap& a = *reinterpret_cast<const CUtensorMap*>(tmA);
  const CUtensorMap& b = *reinterpret_cast<const CUtensorMap*>(tmB);
  if (kind == 1) tc::gemm_attn_umma_kernel<1><<<grid, tc::U_THREADS, smem, s>>>(a, b, p, idesc);
  else tc::gemm_attn_umma_kernel<2><<<grid, tc::U_THREADS, smem, s>>>(a, b, p, idesc);
#ifdef MDT_ATTN_DEBUG
  gemm_attn_umma_debug_dump();
#endif
  return cudaGetLastError();
}

}  // namespace mdt
