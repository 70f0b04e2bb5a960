// gemm_chain.cu -- FeedForward of TransformerBlock (Linear -> GELU -> Linear, + residual; modules.py:314-320, 459) as ONE
// persistent tcgen05 kernel for sm_100a, optionally followed by the LayerNorm of the next attention layer (modules.py:384, 405):
//
//   t[M, C] = t + b2 + GELU(x_op[M, C] * W0[mid, C]^T + b0) * W2[C, mid]^T          cop = tf32/bf16(t)  or  LayerNorm(t)
//
// A CTA owns whole 128-row blocks.  Per block the first GEMM runs as mid / 128 tiles of 128 columns (accumulators in TMEM, two in
// flight); the epilogue warps apply bias + GELU and write the hidden tile to a CTA-private scratch buffer in global memory
// (2 x 128 x mid per CTA, L2 resident: it is rewritten every other block and never has to reach DRAM); the second GEMM then
// TMA-loads the hidden block from that scratch as its A operand and accumulates [128, C] in its own TMEM columns; the final
// epilogue adds bias and residual and writes the fp32 token stream plus the operand-dtype copy the next GEMM reads.  Neither the
// (rows x mid) hidden tensor nor a separate LayerNorm pass touches HBM.
// The second GEMM of block k is queued behind the first-GEMM tiles of block k + 1 (ring order J(k+1,0..T-1), O(k), J(k+2,0), ...)
// and its epilogue runs after theirs, so its operand loads stream while the epilogue warps are busy and nothing waits on L2 latency.
//
//   warp 0      TMA producer (one lane)            warp 1     tcgen05.mma issuer, owns TMEM
//   warps 2-17  epilogue: TMEM quadrant q = warp % 4, column group cg = (warp - 2) / 4 (32 columns of a 128-column tile)
#include <cuda.h>
#include <cuda_bf16.h>
#include "aload.cuh"
#include "tc_common.cuh"

namespace mdt {
namespace tc {

constexpr int C_TM = 128;
constexpr int C_MAXST = 4;
constexpr int C_ABYTES = C_TM * 128;
constexpr int C_EPI_WARPS = 16;
constexpr int C_THREADS = 64 + 32 * C_EPI_WARPS;
constexpr int C_LD = 36;                                    // transposition tile stride (32 + 4 floats)
constexpr int C_STG_BYTES = C_EPI_WARPS * 32 * C_LD * 4;    // one 32 x 32 tile per epilogue warp
constexpr int C_XCH_BYTES = 2 * C_TM * 8 * 2 * 4;           // LayerNorm exchange, double buffered by block parity: [2][128 rows][8 column chunks][mean, M2]

__device__ __forceinline__ void fence_proxy_async_glob() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int KIND>
__device__ __forceinline__ void store_op_f4(void* base, size_t idx, float4 v) {
  if (KIND == 1)
    *reinterpret_cast<uint4*>(reinterpret_cast<float*>(base) + idx) = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
  else
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + idx) = make_uint2(pack_op2<KIND>(v.x, v.y), pack_op2<KIND>(v.z, v.w));
}

// NCH = 32-column chunks of the output tile per epilogue warp: 1 for C = 128, 2 for C = 256
template <int KIND, int NCH>
__global__ void __launch_bounds__(C_THREADS, 1) ff_chain_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB0,
                                                               const __grid_constant__ CUtensorMap tmS,
                                                               const __grid_constant__ CUtensorMap tmW,
                                                               const FFChainParams p, const uint32_t idesc1,
                                                               const uint32_t idesc2) {
  constexpr int KCH = (KIND == 1) ? 32 : 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C_MAXST];
  __shared__ __align__(8) uint64_t empty_bar[C_MAXST];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ __align__(8) uint64_t h_ready, out_full, out_empty;
  __shared__ uint32_t tmem_base_s;

  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = p.C, mid = p.mid;
  const int NST = p.nst, stage_bytes = p.stage_bytes;
  const int T = mid / 128;                       // first-GEMM tiles per block
  const int k1c = C / KCH, k2c = mid / KCH;      // K chunks of the two GEMMs
  const int nblk = (p.M + C_TM - 1) / C_TM;
  const int nk_cta = (int)blockIdx.x < nblk ? (nblk - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto block_of_k = [&](int k) { const int b = (int)blockIdx.x + k * (int)gridDim.x; return p.rev ? nblk - 1 - b : b; };

  if (tid == 0) {
    for (int s = 0; s < C_MAXST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], C_EPI_WARPS); }
    mbar_init(&h_ready, C_EPI_WARPS * 32);
    mbar_init(&out_full, 1);
    mbar_init(&out_empty, C_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB0); tma_prefetch_desc(&tmS); tma_prefetch_desc(&tmW); }
  if (warp == 1) tmem_alloc(&tmem_base_s, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_enter();                                  // the set-up above overlaps the previous grid's tail (launch.cuh)
  const uint32_t tmem_base = tmem_base_s;        // columns [0, C): output accumulator; [C, C + 256): two first-GEMM accumulators

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx1 = (uint32_t)(C_ABYTES + 128 * 128), tx2 = (uint32_t)(C_ABYTES + C * 128);
      auto load_second = [&](int kk) {
        mbar_wait(&h_ready, (uint32_t)kk & 1u);      // the whole hidden block kk is in its scratch buffer
        fence_proxy_async_glob();
        for (int kc = 0; kc < k2c; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], tx2);
          tma_load_3d(sa, &tmS, &full_bar[stage], kc * KCH, (kk & 1) * C_TM, (int)blockIdx.x);
          tma_load_2d(sa + C_ABYTES, &tmW, &full_bar[stage], kc * KCH, 0);
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      };
      for (int k = 0; k < nk_cta; ++k) {
        const int blk = block_of_k(k);
        for (int nt = 0; nt < T; ++nt) {
          for (int kc = 0; kc < k1c; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * stage_bytes;
            mbar_arrive_expect_tx(&full_bar[stage], tx1);
            tma_load_3d(sa, &tmA, &full_bar[stage], kc * KCH, 0, blk * p.Sb);
            tma_load_2d(sa + C_ABYTES, &tmB0, &full_bar[stage], kc * KCH, nt * 128);
            if (++stage == NST) { stage = 0; phase ^= 1u; }
          }
        }
        if (k > 0) load_second(k - 1);     // one whole block behind: its operands stream while block k's tiles are in the epilogue
      }
      if (nk_cta > 0) load_second(nk_cta - 1);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int stage = 0; uint32_t phase = 0;
    int j = 0;
    auto mma_second = [&](int kk) {
      mbar_wait(&out_empty, ((uint32_t)kk & 1u) ^ 1u);
      tc_fence_after();
      for (int kc = 0; kc < k2c; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + C_ABYTES);
#pragma unroll
          for (int kq = 0; kq < 4; ++kq)
            umma<KIND>(tmem_base, adesc + (uint64_t)(2 * kq), bdesc + (uint64_t)(2 * kq), idesc2, (uint32_t)((kc | kq) != 0));
          umma_commit(&empty_bar[stage]);
          if (kc == k2c - 1) umma_commit(&out_full);
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
    };
    for (int k = 0; k < nk_cta; ++k) {
      for (int nt = 0; nt < T; ++nt, ++j) {
        const int buf = j & 1;
        mbar_wait(&acc_empty[buf], ((uint32_t)(j >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(C + buf * 128);
        for (int kc = 0; kc < k1c; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + stage * stage_bytes);
            const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + C_ABYTES);
#pragma unroll
            for (int kq = 0; kq < 4; ++kq)
              umma<KIND>(tmem_d, adesc + (uint64_t)(2 * kq), bdesc + (uint64_t)(2 * kq), idesc1, (uint32_t)((kc | kq) != 0));
            umma_commit(&empty_bar[stage]);
            if (kc == k1c - 1) umma_commit(&acc_full[buf]);
          }
          __syncwarp();
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      }
      if (k > 0) mma_second(k - 1);
    }
    if (nk_cta > 0) mma_second(nk_cta - 1);
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int q = warp & 3;                 // TMEM lane quadrant
    const int cg = ew >> 2;                 // column group 0..3
    float* stg = reinterpret_cast<float*>(smem + NST * stage_bytes) + (size_t)ew * 32 * C_LD;
    float2* xch = reinterpret_cast<float2*>(smem + NST * stage_bytes + C_STG_BYTES);    // [128][8]
    const int cl = (lane & 7) * 4, r0 = lane >> 3;      // coalesced layout: 8 lanes per 128-byte row segment, 4 rows per pass
    const size_t scr_cta = (size_t)blockIdx.x * 2 * C_TM;
    constexpr int nch = NCH;
    const int nchunks_row = C >> 5;

    auto final_epilogue = [&](int kk) {
      const int m0 = block_of_k(kk) * C_TM + q * 32;
      float4 r[NCH][8];
      auto load_res = [&](int u) {
        const int no = (cg * nch + u) * 32 + cl;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int mo = m0 + r0 + i * 4;
          r[u][i] = (p.res && mo < p.M) ? *reinterpret_cast<const float4*>(p.res + (size_t)mo * p.ldres + no) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      load_res(0);          // issued ahead of the accumulator wait: the HBM latency of the residual hides behind it
      mbar_wait(&out_full, (uint32_t)kk & 1u);
      tc_fence_after();
#pragma unroll
      for (int u = 0; u < NCH; ++u) {
        {
          if (u > 0) load_res(u);
          const int col0 = (cg * nch + u) * 32;
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            *reinterpret_cast<uint4*>(stg + lane * C_LD + jj * 4) = make_uint4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
          __syncwarp();
          const int no = col0 + cl;
          const float4 bv = p.b2 ? __ldg(reinterpret_cast<const float4*>(p.b2 + no)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int mo = m0 + r0 + i * 4;
            float4 o = *reinterpret_cast<const float4*>(stg + (r0 + i * 4) * C_LD + cl);
            o.x += bv.x + r[u][i].x; o.y += bv.y + r[u][i].y; o.z += bv.z + r[u][i].z; o.w += bv.w + r[u][i].w;
            r[u][i] = o;
            if (mo < p.M) {
              if (p.C32) *reinterpret_cast<float4*>(p.C32 + (size_t)mo * p.ldc + no) = o;
              if (p.Cop && !p.cop_ln) store_op_f4<KIND>(p.Cop, (size_t)mo * p.ldcop + no, o);
            }
          }
          __syncwarp();
        }
      }
      // the accumulator is in registers now: hand it back before the (optional) LayerNorm tail
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&out_empty);
      if (p.Cop && p.cop_ln) {
        // LayerNorm over the C columns of every row (no affine: folded into the consumer's weights).  Per 32-column chunk the
        // eight lanes of a row hold it entirely: exact local mean and centred sum of squares, then the equal-count form of
        // Chan's merge across chunks:  mean = avg(mean_c),  M2 = sum(M2_c) + 32 * sum((mean_c - mean)^2).
        float2* xb = xch + (size_t)(kk & 1) * C_TM * 8;
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
          {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 o = r[u][i];
              float s = (o.x + o.y) + (o.z + o.w);
              s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
              const float mean = s * (1.0f / 32.0f);
              const float dx = o.x - mean, dy = o.y - mean, dz = o.z - mean, dw = o.w - mean;
              float m2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
              m2 += __shfl_xor_sync(0xffffffffu, m2, 1); m2 += __shfl_xor_sync(0xffffffffu, m2, 2); m2 += __shfl_xor_sync(0xffffffffu, m2, 4);
              if ((lane & 7) == 0) xb[(q * 32 + r0 + i * 4) * 8 + cg * nch + u] = make_float2(mean, m2);
            }
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");      // the four column-group warps of this quadrant
        const float inv_chunks = 1.0f / (float)nchunks_row, inv_c = 1.0f / (float)C;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2* e = xb + (q * 32 + r0 + i * 4) * 8;
          float msum = 0.f, m2 = 0.f;
          for (int c = 0; c < nchunks_row; ++c) { msum += e[c].x; m2 += e[c].y; }
          const float mean = msum * inv_chunks;
          float dev = 0.f;
          for (int c = 0; c < nchunks_row; ++c) { const float dl = e[c].x - mean; dev = fmaf(dl, dl, dev); }
          const float rstd = rsqrtf(fmaf(32.0f, dev, m2) * inv_c + p.ln_eps);
          const int mo = m0 + r0 + i * 4;
#pragma unroll
          for (int u = 0; u < NCH; ++u) {
            if (mo < p.M) {
              const float4 o = r[u][i];
              store_op_f4<KIND>(p.Cop, (size_t)mo * p.ldcop + (cg * nch + u) * 32 + cl,
                                make_float4((o.x - mean) * rstd, (o.y - mean) * rstd, (o.z - mean) * rstd, (o.w - mean) * rstd));
            }
          }
        }
        // no second barrier: the exchange buffer of the other parity is used next, and the barrier of that block orders its reuse
      }
    };

    int j = 0;
    for (int k = 0; k < nk_cta; ++k) {
      const size_t scr_row0 = scr_cta + (size_t)(k & 1) * C_TM + (size_t)q * 32;
      for (int nt = 0; nt < T; ++nt, ++j) {
        const int buf = j & 1;
        mbar_wait(&acc_full[buf], (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(C + buf * 128 + cg * 32), v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);       // values are in registers: the accumulator may be overwritten
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          *reinterpret_cast<uint4*>(stg + lane * C_LD + jj * 4) = make_uint4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
        __syncwarp();
        const int no = nt * 128 + cg * 32 + cl;
        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.b0 + no));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 o = *reinterpret_cast<const float4*>(stg + (r0 + i * 4) * C_LD + cl);
          o.x = gelu_as(o.x + bv.x); o.y = gelu_as(o.y + bv.y); o.z = gelu_as(o.z + bv.z); o.w = gelu_as(o.w + bv.w);
          store_op_f4<KIND>(p.scratch, (scr_row0 + (size_t)(r0 + i * 4)) * mid + no, o);
        }
        __syncwarp();
        if (nt == T - 1) {
          fence_proxy_async_glob();      // publish this thread's hidden values of the block to the async proxy
          mbar_arrive(&h_ready);
        }
      }
      if (k > 0) final_epilogue(k - 1);
    }
    if (nk_cta > 0) final_epilogue(nk_cta - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace tc

static int g_sms_chain = 0;
int ff_chain_sms() {
  if (g_sms_chain == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_chain, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms_chain <= 0) g_sms_chain = 148;
  }
  return g_sms_chain;
}

static const size_t C_SMEM_LIMIT = 232448 - 1024;

static bool ff_chain_config(int C, int* nst, int* stage_bytes, unsigned* tmem_cols, size_t* smem) {
  const size_t s1 = tc::C_ABYTES + 128 * 128, s2 = tc::C_ABYTES + (size_t)C * 128;
  const size_t stage = s1 > s2 ? s1 : s2;
  const size_t fixed = tc::C_STG_BYTES + tc::C_XCH_BYTES + 1024;
  int n = (int)((C_SMEM_LIMIT - fixed) / stage);
  if (n > tc::C_MAXST) n = tc::C_MAXST;
  if (n < 2) return false;
  if (C + 256 > 512) return false;
  unsigned cols = 32;
  while ((int)cols < C + 256) cols <<= 1;
  *nst = n; *stage_bytes = (int)stage; *tmem_cols = cols; *smem = (size_t)n * stage + fixed;
  return true;
}

bool ff_chain_supported(int kind, int C, int mid, int L) {
  const int kch = kind == 1 ? 32 : 64;
  if (kind < 1 || kind > 3) return false;
  if (C != 128 && C != 256) return false;                  // output tile = one UMMA of N = C; LayerNorm chunks per warp = C / 128
  if (mid % 128 || mid < 256 || mid > 2048 || C % kch) return false;
  if (L < 1 || L > 128 || (128 % L) != 0) return false;
  int nst, sb; unsigned tc_; size_t sm;
  return ff_chain_config(C, &nst, &sb, &tc_, &sm);
}

size_t ff_chain_scratch_bytes(int kind, int mid) { return (size_t)ff_chain_sms() * 2 * tc::C_TM * mid * (kind == 1 ? 4 : 2); }

cudaError_t init_ff_chain() {
  cudaError_t e = cudaFuncSetAttribute(tc::ff_chain_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_SMEM_LIMIT);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::ff_chain_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_SMEM_LIMIT);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::ff_chain_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_SMEM_LIMIT);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::ff_chain_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_SMEM_LIMIT);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::ff_chain_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_SMEM_LIMIT);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::ff_chain_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_SMEM_LIMIT);
  return e;
}

cudaError_t launch_ff_chain(const void* tmA, const void* tmB0, const void* tmS, const void* tmW, const FFChainParams& pin, int kind,
                            cudaStream_t s) {
  FFChainParams p = pin;
  if (p.M <= 0) return cudaSuccess;
  size_t smem = 0;
  if (!ff_chain_config(p.C, &p.nst, &p.stage_bytes, &p.tmem_cols, &smem)) return cudaErrorInvalidValue;
  const uint32_t fmt = tc::umma_fmt(kind);
  const uint32_t idesc1 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(tc::C_TM >> 4) << 24);
  const uint32_t idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.C >> 3) << 17) | ((uint32_t)(tc::C_TM >> 4) << 24);
  const int nblk = (p.M + tc::C_TM - 1) / tc::C_TM;
  const int sms = ff_chain_sms();
  const unsigned grid = (unsigned)(nblk < sms ? nblk : sms);
  const CUtensorMap& a = *reinterpret_cast<const CUtensorMap*>(tmA);
  const CUtensorMap& b = *reinterpret_cast<const CUtensorMap*>(tmB0);
  const CUtensorMap& sc = *reinterpret_cast<const CUtensorMap*>(tmS);
  const CUtensorMap& w = *reinterpret_cast<const CUtensorMap*>(tmW);
  auto kern = kind == 1 ? (p.C == 128 ? tc::ff_chain_kernel<1, 1> : tc::ff_chain_kernel<1, 2>)
            : kind == 2 ? (p.C == 128 ? tc::ff_chain_kernel<2, 1> : tc::ff_chain_kernel<2, 2>)
                        : (p.C == 128 ? tc::ff_chain_kernel<3, 1> : tc::ff_chain_kernel<3, 2>);
  return launch_k(kern, grid, tc::C_THREADS, smem, s, a, b, sc, w, p, idesc1, idesc2);
}

}  // namespace mdt
