// gemm_tc.cu -- tcgen05 / TMEM GEMM and implicit-GEMM conv for sm_100a.
//
//   C[M,N] = act(A'[M,K] * W[N,K]^T + bias) (+ res)       A' produced by the fused loader (aload.cuh)
//
// One CTA computes a 128 x BN output tile (BN in {32, 64, 128}); the accumulator lives in TMEM.
// Warp roles:
//   warps 0-3  producers: thread r owns row r of the tile.  For every 128-byte K chunk it gathers
//              the row through the fused loader (conv taps, skip concat, GroupNorm/LayerNorm on load,
//              FiLM, SiLU), converts to the MMA operand type (tf32 round-to-nearest or bf16) and
//              stores it into the SWIZZLE_128B K-major canonical layout; threads r < BN also stage
//              the weight tile (pre-converted in HBM).  generic->async proxy fence, then mbarrier.
//              After the main loop the same warps run the epilogue: tcgen05.ld -> smem staging ->
//              coalesced bias / GELU / residual / store.
//   warp 4     lane 0 issues tcgen05.mma (4 per chunk), tcgen05.commit releases smem stages and
//              finally signals the epilogue; the warp owns TMEM alloc / dealloc.
// Three smem stages of (16 KB A + <=16 KB W) = 96 KB -> two CTAs per SM, so one CTA's epilogue
// overlaps the other's main loop.
#include <cuda_bf16.h>
#include "aload.cuh"
#include "tc_common.cuh"

namespace mdt {

namespace tc {

constexpr int TM = 128;            // tile rows (UMMA M)
constexpr int STAGES = 3;
constexpr int A_BYTES = TM * 128;  // 16 KB per stage
constexpr int B_BYTES = 128 * 128; // up to BN = 128 rows
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack

template <int KIND>  // 1 = tf32 (32 elements per 128-byte chunk), 2 = bf16, 3 = fp16 (64 elements)
__global__ void __launch_bounds__(160) gemm_tc_kernel(const GemmParams p, const int BN, const uint32_t idesc, const int num_chunks) {
  constexpr int KCH = (KIND == 1) ? 32 : 64;
  constexpr int ESZ = (KIND == 1) ? 4 : 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // SWIZZLE_128B operand tiles need 1024-byte alignment
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;

  // keep the pointer in the shared address space (an integer round trip would demote every access to generic LD/ST)
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * BN;
  const uint32_t tmem_cols = BN <= 32 ? 32u : (BN <= 64 ? 64u : 128u);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&tmem_base_s, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_enter();                                  // the set-up above overlaps the previous grid's tail (launch.cuh)
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 4) {
    // ------------------------------------------------------------------ producers
    const ALoad& a = p.a;
    const float* aff = aload_aff(a);
    const int r = tid;
    const int m = m0 + r;
    const bool row_ok = m < p.M;
    const int mm = row_ok ? m : 0;
    const int rb = mm / a.L_out, rlo = mm - rb * a.L_out;
    const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    const int sw = r & 7;
    const bool brow_ok = (r < BN) && (n0 + r < p.N);
    const uint8_t* wrow = reinterpret_cast<const uint8_t*>(p.Wtc) + (size_t)(n0 + r) * p.K * ESZ;

    for (int c = 0; c < num_chunks; ++c) {
      const int stage = c % STAGES;
      const uint32_t phase = (uint32_t)(c / STAGES) & 1u;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      uint8_t* sa = smem + stage * STAGE_BYTES;
      uint8_t* sb = sa + A_BYTES;
      const int k0 = c * KCH;
      // ---- A row chunk: KCH elements -> 8 x 16 bytes, swizzled
      if (KIND == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = k0 + j * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && k < p.K) v = aload4(a, aff, rb, rlo, k);
          uint4 o = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
          *reinterpret_cast<uint4*>(sa + row_off + ((j ^ sw) << 4)) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = k0 + j * 8;
          float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
          if (row_ok && k < p.K) v0 = aload4(a, aff, rb, rlo, k);
          if (row_ok && k + 4 < p.K) v1 = aload4(a, aff, rb, rlo, k + 4);
          uint4 o = make_uint4(pack_op2<KIND>(v0.x, v0.y), pack_op2<KIND>(v0.z, v0.w), pack_op2<KIND>(v1.x, v1.y), pack_op2<KIND>(v1.z, v1.w));
          *reinterpret_cast<uint4*>(sa + row_off + ((j ^ sw) << 4)) = o;
        }
      }
      // ---- W row chunk (pre-converted): rows 0..BN-1
      if (r < BN) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = k0 + j * (16 / ESZ);
          uint4 o = make_uint4(0u, 0u, 0u, 0u);
          if (brow_ok && k < p.K) o = __ldg(reinterpret_cast<const uint4*>(wrow + (size_t)k * ESZ));
          *reinterpret_cast<uint4*>(sb + row_off + ((j ^ sw) << 4)) = o;
        }
      }
      fence_proxy_async();
      mbar_arrive(&full_bar[stage]);
    }

    // ------------------------------------------------------------------ epilogue
    mbar_wait(&accum_bar, 0u);
    tc_fence_after();
    float* stg = reinterpret_cast<float*>(smem);  // [128][BN + 4] staging over the drained pipeline buffers
    const int ldst = BN + 4;
    for (int cc = 0; cc < BN / 32; ++cc) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cc * 32), v);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(stg + (size_t)r * ldst + cc * 32 + j * 4) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int n4 = BN >> 2;
    for (int idx = tid; idx < TM * n4; idx += 128) {
      const int rr = idx / n4, c4 = (idx - rr * n4) * 4;
      const int mo = m0 + rr, no = n0 + c4;
      if (mo >= p.M || no >= p.N) continue;
      float4 o = *reinterpret_cast<const float4*>(stg + (size_t)rr * ldst + c4);
      if (p.bias) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + no));
        o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
      }
      if (p.act == 1) { o.x = gelu_f(o.x); o.y = gelu_f(o.y); o.z = gelu_f(o.z); o.w = gelu_f(o.w); }
      if (p.res) {
        const float4 rv = *reinterpret_cast<const float4*>(p.res + (size_t)mo * p.ldres + no);
        o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
      }
      *reinterpret_cast<float4*>(p.C + (size_t)mo * p.ldc + no) = o;
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    for (int c = 0; c < num_chunks; ++c) {
      const int stage = c % STAGES;
      const uint32_t phase = (uint32_t)(c / STAGES) & 1u;
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 4 MMAs of 32 bytes of K each (K = 8 tf32 / 16 bf16); +32 B = +2 in desc units
          umma<KIND>(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((c | k) != 0));
        umma_commit(&empty_bar[stage]);
        if (c == num_chunks - 1) umma_commit(&accum_bar);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, tmem_cols);
}

__global__ void convert_tf32_kernel(const float* __restrict__ in, uint32_t* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = to_tf32(in[i]);
}
__global__ void convert_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}

__global__ void convert_f16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint16_t)(pack_f16s(in[i], 0.f) & 0xffffu);
}

static int pick_bn(int N) {
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  if (N % 32 == 0) return 32;
  return 0;
}

}  // namespace tc

bool gemm_tc_supported(const GemmParams& p) {
  if (tc::pick_bn(p.N) == 0) return false;
  if (!aload_vec4_ok(p.a) || p.K % 8 != 0) return false;
  if (p.ldc % 4 != 0 || (p.res && p.ldres % 4 != 0)) return false;
  return true;
}

cudaError_t convert_weights_tc(const float* W, void* Wtc, long long n, int kind, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (kind == 1) tc::convert_tf32_kernel<<<grid, 256, 0, s>>>(W, reinterpret_cast<uint32_t*>(Wtc), n);
  else if (kind == 2) tc::convert_bf16_kernel<<<grid, 256, 0, s>>>(W, reinterpret_cast<__nv_bfloat16*>(Wtc), n);
  else tc::convert_f16_kernel<<<grid, 256, 0, s>>>(W, reinterpret_cast<uint16_t*>(Wtc), n);
  return cudaGetLastError();
}

cudaError_t init_gemm_tc() {
  cudaError_t e = cudaFuncSetAttribute(tc::gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
  return e;
}

cudaError_t launch_gemm_tc(const GemmParams& p, int kind, cudaStream_t s) {
  if (p.M <= 0 || p.N <= 0) return cudaSuccess;
  const int BN = tc::pick_bn(p.N);
  if (BN == 0 || !p.Wtc) return cudaErrorInvalidValue;
  const int kch = kind == 1 ? 32 : 64;
  const int num_chunks = (p.K + kch - 1) / kch;
  // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6); a/b format [7,10)/[10,13): TF32 = 2, BF16 = 1;
  // K-major both; N >> 3 at [17,23); M >> 4 at [24,29)
  const uint32_t fmt = tc::umma_fmt(kind);
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(tc::TM >> 4) << 24);
  dim3 grid((p.M + tc::TM - 1) / tc::TM, p.N / BN);
  return launch_k(kind == 1 ? tc::gemm_tc_kernel<1> : (kind == 2 ? tc::gemm_tc_kernel<2> : tc::gemm_tc_kernel<3>), grid, 160, tc::SMEM_BYTES, s, p, BN, idesc, num_chunks);
}

}  // namespace mdt
