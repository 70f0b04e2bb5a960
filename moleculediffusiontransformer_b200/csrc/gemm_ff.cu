// gemm_ff.cu -- fused FeedForward (Linear -> GELU -> Linear, + residual; modules.py:314-320, 459) for sm_100a.
//
//   out[M, C] = x[M, C] + b2 + GELU(x_op[M, C] * W0[mid, C]^T + b0) * W2[C, mid]^T
//
// The (rows x mid) hidden activation never goes to HBM: a persistent CTA owns a 128-row block and walks the hidden
// dimension in slices of 64.  Per slice the first GEMM (N = 64) accumulates in TMEM, the epilogue warps apply bias +
// GELU and write the slice straight into shared memory in the K-major SWIZZLE_128B operand layout, and the second
// GEMM (K = 64, N = C) accumulates the output tile in a second TMEM region.  Chained pipelines:
//   smem ring     TMA producer <-> MMA issuer            (A + W0 chunks for GEMM 1, W2 chunks for GEMM 2)
//   acc1 ring     MMA issuer  <-> epilogue warps         two 64-column accumulators
//   operand ring  epilogue    <-> MMA issuer             two hidden-slice operand buffers
//   out           MMA issuer  <-> epilogue               one C-column accumulator per row block
#include <cuda.h>
#include <cuda_bf16.h>
#include "aload.cuh"
#include "tc_common.cuh"

namespace mdt {
namespace tc {

constexpr int F_TM = 128;
constexpr int F_ABYTES = F_TM * 128;
constexpr int F_EPI_WARPS = 16;    // two groups of 8: group 0 owns the even hidden slices, group 1 the odd ones
constexpr int F_THREADS = 64 + 32 * F_EPI_WARPS;
constexpr int F_SLICE = 64;       // hidden columns per slice

template <int KIND>
__global__ void __launch_bounds__(F_THREADS, 1) gemm_ff_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmW0,
                                                              const __grid_constant__ CUtensorMap tmW2, const GemmFFParams p,
                                                              const uint32_t idesc1, const uint32_t idesc2, const int stages,
                                                              const int stage_bytes) {
  constexpr int KCH = (KIND == 1) ? 32 : 64;
  constexpr int ESZ = (KIND == 1) ? 4 : 2;
  constexpr int OPCH = F_SLICE / KCH;                 // operand chunks per hidden slice (2 for tf32, 1 for bf16)
  constexpr int OP_BYTES = OPCH * F_ABYTES;           // one hidden-slice operand buffer
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[6];
  __shared__ __align__(8) uint64_t empty_bar[6];
  __shared__ __align__(8) uint64_t acc1_full[2], acc1_empty[2], op_full[2], op_empty[2], out_full, out_empty;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = p.C, mid = p.mid;
  const int S = mid / F_SLICE;                         // slices per row block
  const int k1c = C / KCH;                             // K chunks of GEMM 1
  const int m_tiles = (p.M + F_TM - 1) / F_TM;
  const uint32_t tmem_cols = C <= 128 ? 256u : 512u;   // [0, 128): two acc1 buffers, [128, 128 + C): output accumulator
  uint8_t* opbuf = smem + stages * stage_bytes;        // two operand buffers

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc1_full[s], 1); mbar_init(&acc1_empty[s], F_EPI_WARPS / 2);
      mbar_init(&op_full[s], F_EPI_WARPS / 2); mbar_init(&op_empty[s], 1);
    }
    mbar_init(&out_full, 1); mbar_init(&out_empty, F_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW0); tma_prefetch_desc(&tmW2); }
  if (warp == 1) tmem_alloc(&tmem_base_s, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_out = tmem_base + 128u;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int c = 0;
      auto g1 = [&](int mt, int s) {       // A chunk + W0 slice chunk
        for (int kc = 0; kc < k1c; ++kc, ++c) {
          const int stage = c % stages;
          mbar_wait(&empty_bar[stage], ((uint32_t)(c / stages) & 1u) ^ 1u);
          uint8_t* sa = smem + stage * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(F_ABYTES + F_SLICE * 128));
          tma_load_3d(sa, &tmA, &full_bar[stage], kc * KCH, 0, mt * p.Sb);
          tma_load_2d(sa + F_ABYTES, &tmW0, &full_bar[stage], kc * KCH, s * F_SLICE);
        }
      };
      auto g2 = [&](int s) {               // W2[:, slice] chunks (B operand only)
        for (int kc = 0; kc < OPCH; ++kc, ++c) {
          const int stage = c % stages;
          mbar_wait(&empty_bar[stage], ((uint32_t)(c / stages) & 1u) ^ 1u);
          uint8_t* sa = smem + stage * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(C * 128));
          tma_load_2d(sa + F_ABYTES, &tmW2, &full_bar[stage], s * F_SLICE + kc * KCH, 0);
        }
      };
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
        g1(mt, 0);
        for (int s = 1; s < S; ++s) { g1(mt, s); g2(s - 1); }
        g2(S - 1);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int c = 0, g = 0, ti = 0;              // ring chunk counter, global slice counter, row-block counter
    auto mma1 = [&](int gg) {
      const int buf = gg & 1;
      mbar_wait(&acc1_empty[buf], ((uint32_t)(gg >> 1) & 1u) ^ 1u);
      tc_fence_after();
      for (int kc = 0; kc < k1c; ++kc, ++c) {
        const int stage = c % stages;
        mbar_wait(&full_bar[stage], (uint32_t)(c / stages) & 1u);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + F_ABYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma<KIND>(tmem_base + (uint32_t)(buf * F_SLICE), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc1,
                       (uint32_t)((kc | k) != 0));
          umma_commit(&empty_bar[stage]);
          if (kc == k1c - 1) umma_commit(&acc1_full[buf]);
        }
        __syncwarp();
      }
    };
    auto mma2 = [&](int gg, bool first_of_tile, bool last_of_tile) {
      const int buf = gg & 1;
      if (first_of_tile) { mbar_wait(&out_empty, ((uint32_t)ti & 1u) ^ 1u); tc_fence_after(); }
      mbar_wait(&op_full[buf], (uint32_t)(gg >> 1) & 1u);
      tc_fence_after();
      for (int kc = 0; kc < OPCH; ++kc, ++c) {
        const int stage = c % stages;
        mbar_wait(&full_bar[stage], (uint32_t)(c / stages) & 1u);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = make_desc(smem_u32(opbuf + buf * OP_BYTES + kc * F_ABYTES));
          const uint64_t bdesc = make_desc(smem_u32(smem + stage * stage_bytes + F_ABYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma<KIND>(tmem_out, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc2,
                       (uint32_t)(!(first_of_tile && kc == 0 && k == 0)));
          umma_commit(&empty_bar[stage]);
          if (kc == OPCH - 1) { umma_commit(&op_empty[buf]); if (last_of_tile) umma_commit(&out_full); }
        }
        __syncwarp();
      }
    };
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++ti) {
      mma1(g);
      for (int s = 1; s < S; ++s) { mma1(g + s); mma2(g + s - 1, s == 1, false); }
      mma2(g + S - 1, S == 1, true);
      g += S;
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;                  // 0..15
    const int q = warp & 3;                   // TMEM lane quadrant
    const int grp = ew >> 3;                  // slice parity owned by this warp (S is even, so parity == buffer index)
    const int half = (ew & 7) >> 2;           // GEMM 1: which 32 of the 64 slice columns
    const int cg = ew >> 2;                   // GEMM 2: which quarter of the C output columns
    const int row = q * 32 + lane;
    int g0 = 0, ti = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++ti, g0 += S) {
      for (int s = grp; s < S; s += 2) {
        const int g = g0 + s;
        const int buf = grp;
        mbar_wait(&acc1_full[buf], (uint32_t)(g >> 1) & 1u);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * F_SLICE + half * 32), v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_empty[buf]);
        const float* b0 = p.b0 + s * F_SLICE + half * 32;
        float h[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(b0 + j * 4));
          h[4 * j] = gelu_as(__uint_as_float(v[4 * j]) + bv.x);
          h[4 * j + 1] = gelu_as(__uint_as_float(v[4 * j + 1]) + bv.y);
          h[4 * j + 2] = gelu_as(__uint_as_float(v[4 * j + 2]) + bv.z);
          h[4 * j + 3] = gelu_as(__uint_as_float(v[4 * j + 3]) + bv.w);
        }
        mbar_wait(&op_empty[buf], ((uint32_t)(g >> 1) & 1u) ^ 1u);   // GEMM 2 of the slice two steps back has drained it
        uint8_t* ob = opbuf + buf * OP_BYTES;
        const uint32_t roff = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        if (KIND == 1) {
          // 32 tf32 columns = this warp's whole 128-byte chunk row: chunk index = half
          uint8_t* dst = ob + half * F_ABYTES + roff;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(dst + ((j ^ (row & 7)) << 4)) =
                make_uint4(to_tf32(h[4 * j]), to_tf32(h[4 * j + 1]), to_tf32(h[4 * j + 2]), to_tf32(h[4 * j + 3]));
        } else {
          // 32 bf16 columns = 64 bytes = 16-byte units [4 * half, 4 * half + 4) of the single chunk
          uint8_t* dst = ob + roff;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(dst + (((half * 4 + j) ^ (row & 7)) << 4)) =
                make_uint4(pack_bf16(h[8 * j], h[8 * j + 1]), pack_bf16(h[8 * j + 2], h[8 * j + 3]),
                           pack_bf16(h[8 * j + 4], h[8 * j + 5]), pack_bf16(h[8 * j + 6], h[8 * j + 7]));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&op_full[buf]);
      }
      // ---- output tile (once per row block): + b2 + residual; each thread owns one row, 32 columns per pass
      mbar_wait(&out_full, (uint32_t)ti & 1u);
      tc_fence_after();
      const int cols_per_grp = C / 4;
      const int mo = mt * F_TM + row;
      for (int cc = 0; cc < cols_per_grp; cc += 32) {
        uint32_t v[32];
        const int n0 = cg * cols_per_grp + cc;
        tmem_ld32(tmem_out + ((uint32_t)(q * 32) << 16) + (uint32_t)n0, v);
        if (mo < p.M) {
          const float* rp = p.res + (size_t)mo * C + n0;
          float* op32 = p.out + (size_t)mo * C + n0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.b2 + n0 + j * 4));
            const float4 rv = *reinterpret_cast<const float4*>(rp + j * 4);
            float4 o;
            o.x = __uint_as_float(v[4 * j]) + bv.x + rv.x; o.y = __uint_as_float(v[4 * j + 1]) + bv.y + rv.y;
            o.z = __uint_as_float(v[4 * j + 2]) + bv.z + rv.z; o.w = __uint_as_float(v[4 * j + 3]) + bv.w + rv.w;
            *reinterpret_cast<float4*>(op32 + j * 4) = o;
            if (p.out_op) {
              if (KIND == 1)
                *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.out_op) + (size_t)mo * C + n0 + j * 4) =
                    make_uint4(to_tf32(o.x), to_tf32(o.y), to_tf32(o.z), to_tf32(o.w));
              else
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out_op) + (size_t)mo * C + n0 + j * 4) =
                    make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&out_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace tc

static void ff_smem(int kind, int C, int* stages, int* stage_bytes, size_t* total) {
  const int esz = kind == 1 ? 4 : 2;
  const int op_bytes = tc::F_TM * tc::F_SLICE * esz;                     // hidden-slice operand buffer
  const int b_bytes = (C * 128 > tc::F_SLICE * 128) ? C * 128 : tc::F_SLICE * 128;
  *stage_bytes = tc::F_ABYTES + b_bytes;
  const size_t fixed = 2 * (size_t)op_bytes;
  int st = (int)((220 * 1024 - fixed) / *stage_bytes);
  if (st > 6) st = 6;
  *stages = st;
  *total = fixed + (size_t)st * *stage_bytes;
}

bool gemm_ff_supported(int kind, int C, int mid, int L) {
  const int kch = kind == 1 ? 32 : 64;
  if (C % kch || mid % (2 * tc::F_SLICE) || (C != 128 && C != 256)) return false;   // even slice count; C = UMMA N of GEMM 2   // C = TMEM columns / UMMA N of GEMM 2
  if (L < 1 || L > 128 || (L & (L - 1))) return false;
  int st, sb; size_t tot;
  ff_smem(kind, C, &st, &sb, &tot);
  return st >= 2;
}

cudaError_t init_gemm_ff() {
  cudaError_t e = cudaFuncSetAttribute(tc::gemm_ff_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(tc::gemm_ff_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
}

static int g_sms_ff = 0;

cudaError_t launch_gemm_ff(const void* tmA, const void* tmW0, const void* tmW2, const GemmFFParams& p, int kind, cudaStream_t s) {
  if (p.M <= 0) return cudaSuccess;
  if (g_sms_ff == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_ff, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms_ff <= 0) g_sms_ff = 148;
  }
  int stages, stage_bytes; size_t smem;
  ff_smem(kind, p.C, &stages, &stage_bytes, &smem);
  if (stages < 2) return cudaErrorInvalidValue;
  const uint32_t fmt = kind == 1 ? 2u : 1u;
  const uint32_t base = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(tc::F_TM >> 4) << 24);
  const uint32_t idesc1 = base | ((uint32_t)(tc::F_SLICE >> 3) << 17);
  const uint32_t idesc2 = base | ((uint32_t)(p.C >> 3) << 17);
  const int m_tiles = (p.M + tc::F_TM - 1) / tc::F_TM;
  const unsigned grid = (unsigned)(m_tiles < g_sms_ff ? m_tiles : g_sms_ff);
  const CUtensorMap& a = *reinterpret_cast<const CUtensorMap*>(tmA);
  const CUtensorMap& w0 = *reinterpret_cast<const CUtensorMap*>(tmW0);
  const CUtensorMap& w2 = *reinterpret_cast<const CUtensorMap*>(tmW2);
  if (kind == 1) tc::gemm_ff_kernel<1><<<grid, tc::F_THREADS, smem, s>>>(a, w0, w2, p, idesc1, idesc2, stages, stage_bytes);
  else tc::gemm_ff_kernel<2><<<grid, tc::F_THREADS, smem, s>>>(a, w0, w2, p, idesc1, idesc2, stages, stage_bytes);
  return cudaGetLastError();
}

}  // namespace mdt
