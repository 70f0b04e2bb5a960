// kernels.cu -- CUDA-core kernels of the sampling path (sm_100a).
//   * gemm_fp32_kernel      fp32 FFMA GEMM / implicit-GEMM conv with fused norm/FiLM/SiLU prologue and
//                           bias/GELU/residual epilogue (the "fp32" precision mode; also the fallback
//                           shape coverage for layers too small or odd for the tcgen05 kernel)
//   * attention_kernel      per-(sample, head) softmax attention, sequence lengths <= 64
//   * groupnorm/rownorm     two-pass statistics feeding the GEMM prologues
//   * sampler kernels       fused ADPM2 / EDM / classifier-free-guidance update (HBM-bound)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include "aload.cuh"

// launch through launch_k_light() (launch.cuh: programmatic dependent launch); a failed launch returns its error
#define MDT_CK_LAUNCH(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

namespace mdt {

// ================================================================================================
// fp32 GEMM:  C[M,N] = act(A'[M,K] * W[N,K]^T + bias) (+ res)
// Tile 128 x 64 x 16, 256 threads, 8 x 4 outputs per thread, register-prefetch double buffering.
// ================================================================================================
constexpr int BM = 128, BN = 64, BK = 16;
constexpr int APAD = 4, BPAD = 4;

template <int VEC>
__global__ void __launch_bounds__(256) gemm_fp32_kernel(const GemmParams p) {
  pdl_enter();
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN + BPAD];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const ALoad& a = p.a;
  const float* aff = aload_aff(a);

  // ---- per-thread load coordinates
  // VEC=4: A: 2 float4 per thread (row = idx/4, kq = idx%4), B: 1 float4 per thread
  // VEC=1: A: 8 scalars per thread (row = idx/16, kk = idx%16), B: 4 scalars per thread
  int arow_b[2], arow_lo[2];
  bool arow_ok[2];
  if (VEC == 4) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = (tid + i * 256) >> 2;
      const int m = m0 + r;
      arow_ok[i] = m < p.M;
      const int mm = arow_ok[i] ? m : 0;
      arow_b[i] = mm / a.L_out;
      arow_lo[i] = mm - arow_b[i] * a.L_out;
    }
  }

  float4 areg[2];
  float4 breg;
  float areg1[8], breg1[4];

  auto load_tile = [&](int k0) {
    if (VEC == 4) {
      const int kq = (tid & 3) * 4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        areg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (arow_ok[i] && k0 + kq < p.K) areg[i] = aload4(a, aff, arow_b[i], arow_lo[i], k0 + kq);
      }
      const int n = n0 + (tid >> 2);
      breg = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < p.N && k0 + kq < p.K) breg = __ldg(reinterpret_cast<const float4*>(p.W + (size_t)n * p.K + k0 + kq));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = tid + i * 256;
        const int r = idx >> 4, kk = idx & 15;
        const int m = m0 + r;
        float v = 0.f;
        if (m < p.M && k0 + kk < p.K) {
          const int b = m / a.L_out;
          v = aload1(a, aff, b, m - b * a.L_out, k0 + kk);
        }
        areg1[i] = v;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * 256;
        const int n = n0 + (idx >> 4), kk = idx & 15;
        breg1[i] = (n < p.N && k0 + kk < p.K) ? __ldg(p.W + (size_t)n * p.K + k0 + kk) : 0.f;
      }
    }
  };
  auto store_tile = [&](int buf) {
    if (VEC == 4) {
      const int kq = (tid & 3) * 4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = (tid + i * 256) >> 2;
        As[buf][kq + 0][r] = areg[i].x; As[buf][kq + 1][r] = areg[i].y;
        As[buf][kq + 2][r] = areg[i].z; As[buf][kq + 3][r] = areg[i].w;
      }
      const int n = tid >> 2;
      Bs[buf][kq + 0][n] = breg.x; Bs[buf][kq + 1][n] = breg.y;
      Bs[buf][kq + 2][n] = breg.z; Bs[buf][kq + 3][n] = breg.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const int idx = tid + i * 256; As[buf][idx & 15][idx >> 4] = areg1[i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) { const int idx = tid + i * 256; Bs[buf][idx & 15][idx >> 4] = breg1[i]; }
    }
  };

  const int tm = tid >> 4;  // 0..15 -> rows tm*8..+7
  const int tn = tid & 15;  // 0..15 -> cols tn*4..+3
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int T = (p.K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    const int buf = t & 1;
    if (t + 1 < T) load_tile((t + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < T) store_tile(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  const int nb = n0 + tn * 4;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (nb + j < p.N) bias[j] = __ldg(p.bias + nb + j);
  }
  const bool vec_ok = (nb + 3 < p.N) && (p.ldc % 4 == 0) && (!p.res || p.ldres % 4 == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + tm * 8 + i;
    if (m >= p.M) continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[i][j] + bias[j];
      if (p.act == 1) v = gelu_f(v);
      o[j] = v;
    }
    if (vec_ok) {
      if (p.res) {
        const float4 r = *reinterpret_cast<const float4*>(p.res + (size_t)m * p.ldres + nb);
        o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
      }
      *reinterpret_cast<float4*>(p.C + (size_t)m * p.ldc + nb) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (nb + j < p.N) {
          float v = o[j];
          if (p.res) v += p.res[(size_t)m * p.ldres + nb + j];
          p.C[(size_t)m * p.ldc + nb + j] = v;
        }
      }
    }
  }
}

cudaError_t launch_gemm_fp32(const GemmParams& p, cudaStream_t s) {
  if (p.M <= 0 || p.N <= 0) return cudaSuccess;
  dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN);
  const bool vec = aload_vec4_ok(p.a) && (p.K % 4 == 0);
  return launch_k_light(vec ? gemm_fp32_kernel<4> : gemm_fp32_kernel<1>, grid, 256, 0, s, p);
}

// ================================================================================================
// Attention core (AttentionBase.forward, modules.py:350-364): one warp per (sample, head).
// ================================================================================================
template <int KIND> struct AttnIO;
template <> struct AttnIO<0> {
  typedef float T;
  static __device__ __forceinline__ float4 ld4(const void* p, size_t i) { return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i); }
  static __device__ __forceinline__ void st(void* p, size_t i, float v) { reinterpret_cast<float*>(p)[i] = v; }
};
template <> struct AttnIO<1> {
  typedef float T;
  static __device__ __forceinline__ float4 ld4(const void* p, size_t i) { return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i); }
  static __device__ __forceinline__ void st(void* p, size_t i, float v) {
    uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    reinterpret_cast<uint32_t*>(p)[i] = r;
  }
};
template <> struct AttnIO<2> {
  typedef __nv_bfloat16 T;
  static __device__ __forceinline__ float4 ld4(const void* p, size_t i) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void st(void* p, size_t i, float v) { reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v); }
};

template <> struct AttnIO<3> {
  typedef __half T;
  static __device__ __forceinline__ float4 ld4(const void* p, size_t i) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p) + i);
    const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), fb = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void st(void* p, size_t i, float v) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(v));
    reinterpret_cast<uint16_t*>(p)[i] = (uint16_t)(r & 0xffffu);
  }
};

template <int KIND>
__global__ void attention_kernel(const AttnParams p, int warps_per_cta) {
  typedef AttnIO<KIND> IO;
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wg = (long long)blockIdx.x * warps_per_cta + warp;
  if (wg >= (long long)p.B * p.heads) return;
  const int b = (int)(wg / p.heads), h = (int)(wg % p.heads);
  const int d = p.d, dp = d + 4, nq = p.nq, nk = p.nk;
  const int per_warp = ((nq + nk) * dp + nk * d + nq * (nk + 1) + 3) & ~3;  // keep float4 alignment per warp
  float* qs = smem + (size_t)warp * per_warp;
  float* ks = qs + nq * dp;
  float* vs = ks + nk * dp;
  float* ss = vs + nk * d;

  const void* kbase = p.k;
  const void* vbase = p.v;
  size_t koff = (size_t)b * p.kv_sample_stride;
  if (p.k_null && b >= p.n_cond) { kbase = p.k_null; vbase = p.v_null; koff = 0; }
  const int d4 = d >> 2;
  for (int idx = lane; idx < nq * d4; idx += 32) {
    const int r = idx / d4, c = (idx - r * d4) * 4;
    *reinterpret_cast<float4*>(qs + r * dp + c) = IO::ld4(p.q, ((size_t)b * nq + r) * p.ldq + h * d + c);
  }
  for (int idx = lane; idx < nk * d4; idx += 32) {
    const int r = idx / d4, c = (idx - r * d4) * 4;
    *reinterpret_cast<float4*>(ks + r * dp + c) = IO::ld4(kbase, koff + (size_t)r * p.ldkv + h * d + c);
    *reinterpret_cast<float4*>(vs + r * d + c) = IO::ld4(vbase, koff + (size_t)r * p.ldkv + h * d + c);
  }
  __syncwarp();
  // S = (Q K^T) * scale : 2 x 4 register tiles (6 shared loads per 32 FMAs)
  {
    const int nbi = (nq + 1) >> 1, nbj = (nk + 3) >> 2;
    for (int blk = lane; blk < nbi * nbj; blk += 32) {
      const int i0 = (blk / nbj) * 2, j0 = (blk % nbj) * 4;
      const int i1 = min(i0 + 1, nq - 1);
      const int jj[4] = {j0, min(j0 + 1, nk - 1), min(j0 + 2, nk - 1), min(j0 + 3, nk - 1)};
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      for (int c = 0; c < d4; ++c) {
        const float4 qa = *reinterpret_cast<const float4*>(qs + i0 * dp + c * 4);
        const float4 qb = *reinterpret_cast<const float4*>(qs + i1 * dp + c * 4);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float4 kk = *reinterpret_cast<const float4*>(ks + jj[t] * dp + c * 4);
          acc[0][t] = fmaf(qa.x, kk.x, acc[0][t]); acc[0][t] = fmaf(qa.y, kk.y, acc[0][t]);
          acc[0][t] = fmaf(qa.z, kk.z, acc[0][t]); acc[0][t] = fmaf(qa.w, kk.w, acc[0][t]);
          acc[1][t] = fmaf(qb.x, kk.x, acc[1][t]); acc[1][t] = fmaf(qb.y, kk.y, acc[1][t]);
          acc[1][t] = fmaf(qb.z, kk.z, acc[1][t]); acc[1][t] = fmaf(qb.w, kk.w, acc[1][t]);
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (j0 + t < nk) {
          ss[i0 * (nk + 1) + j0 + t] = acc[0][t] * p.scale;
          if (i0 + 1 < nq) ss[(i0 + 1) * (nk + 1) + j0 + t] = acc[1][t] * p.scale;
        }
      }
    }
  }
  __syncwarp();
  // row softmax
  for (int i = lane; i < nq; i += 32) {
    float* row = ss + i * (nk + 1);
    float mx = row[0];
    for (int j = 1; j < nk; ++j) mx = fmaxf(mx, row[j]);
    float sum = 0.f;
    for (int j = 0; j < nk; ++j) { const float e = expf(row[j] - mx); row[j] = e; sum += e; }
    const float inv = 1.0f / sum;
    for (int j = 0; j < nk; ++j) row[j] *= inv;
  }
  __syncwarp();
  // O = P V : each lane owns DD = d / 32 output features; V chunk (16 keys) and 16 query rows live in registers
  if (d == 64) {
    for (int i0 = 0; i0 < nq; i0 += 16) {
      float acc[16][2];
#pragma unroll
      for (int ii = 0; ii < 16; ++ii) { acc[ii][0] = 0.f; acc[ii][1] = 0.f; }
      for (int j0 = 0; j0 < nk; j0 += 16) {
        float vr[16][2];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const bool ok = j0 + t < nk;
          vr[t][0] = ok ? vs[(j0 + t) * d + lane] : 0.f;
          vr[t][1] = ok ? vs[(j0 + t) * d + 32 + lane] : 0.f;
        }
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) {
          if (i0 + ii < nq) {
            const float* row = ss + (i0 + ii) * (nk + 1) + j0;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              const float pv = (j0 + t < nk) ? row[t] : 0.f;
              acc[ii][0] = fmaf(pv, vr[t][0], acc[ii][0]);
              acc[ii][1] = fmaf(pv, vr[t][1], acc[ii][1]);
            }
          }
        }
      }
#pragma unroll
      for (int ii = 0; ii < 16; ++ii) {
        if (i0 + ii < nq) {
          const size_t o = ((size_t)b * nq + i0 + ii) * p.ldo + h * d;
          IO::st(p.o, o + lane, acc[ii][0]);
          IO::st(p.o, o + 32 + lane, acc[ii][1]);
        }
      }
    }
  } else {
    for (int dd = lane; dd < d; dd += 32) {
      for (int i = 0; i < nq; ++i) {
        const float* row = ss + i * (nk + 1);
        float acc = 0.f;
        for (int j = 0; j < nk; ++j) acc = fmaf(row[j], vs[j * d + dd], acc);
        IO::st(p.o, ((size_t)b * nq + i) * p.ldo + h * d + dd, acc);
      }
    }
  }
}

cudaError_t init_prep();
static cudaError_t init_step_kernels();
// dynamic shared memory the sampler-state kernels may request for their (P, L) transposition tile (opt-in above 48 KB)
#define MDT_STEP_SMEM_MAX ((size_t)200 * 1024)
cudaError_t init_kernels() {
  cudaError_t e = cudaFuncSetAttribute(attention_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = init_prep();
  if (e == cudaSuccess) e = init_attention_bulk();
  if (e == cudaSuccess) e = init_step_kernels();
  return e;
}

cudaError_t launch_attention(const AttnParams& p, int kind, cudaStream_t s) {
  if (p.B <= 0) return cudaSuccess;
  if (attention_bulk_supported(p, kind)) return launch_attention_bulk(p, kind, s);
  if (p.d % 4 != 0 || p.nq > 128 || p.nk > 128) return cudaErrorInvalidValue;
  const size_t per_warp = ((((size_t)(p.nq + p.nk) * (p.d + 4) + (size_t)p.nk * p.d + (size_t)p.nq * (p.nk + 1)) + 3) & ~(size_t)3) * sizeof(float);
  int wpc = (int)((96 * 1024) / per_warp);
  if (wpc > 8) wpc = 8;
  if (wpc < 1) wpc = 1;
  const size_t smem = per_warp * wpc;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  const long long warps = (long long)p.B * p.heads;
  const unsigned grid = (unsigned)((warps + wpc - 1) / wpc);
  if (kind == 0) attention_kernel<0><<<grid, wpc * 32, smem, s>>>(p, wpc);
  else if (kind == 1) attention_kernel<1><<<grid, wpc * 32, smem, s>>>(p, wpc);
  else if (kind == 2) attention_kernel<2><<<grid, wpc * 32, smem, s>>>(p, wpc);
  else attention_kernel<3><<<grid, wpc * 32, smem, s>>>(p, wpc);
  return cudaGetLastError();
}

// ================================================================================================
// Normalisation statistics (two-pass, fp32): feed the GEMM prologues.
// ================================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per (sample, group).  nn.GroupNorm semantics: biased variance over (C/G) x L elements.
__global__ void groupnorm_stats_kernel(const NormStatsParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (wg >= (long long)p.rows * p.groups) return;
  const int b = (int)(wg / p.groups), g = (int)(wg % p.groups);
  const int C = p.c0 + p.c1, cpg = C / p.groups, n = cpg * p.L;
  auto fetch = [&](int idx) -> float {
    const int l = idx / cpg, c = g * cpg + (idx - l * cpg);
    const size_t row = (size_t)b * p.L + l;
    return (c < p.c0) ? __ldg(p.src0 + row * p.c0 + c) : __ldg(p.src1 + row * p.c1 + (c - p.c0)) * p.scale1;
  };
  float sum = 0.f;
  for (int idx = lane; idx < n; idx += 32) sum += fetch(idx);
  const float mean = warp_sum(sum) / (float)n;
  float sq = 0.f;
  for (int idx = lane; idx < n; idx += 32) { const float dlt = fetch(idx) - mean; sq = fmaf(dlt, dlt, sq); }
  const float var = warp_sum(sq) / (float)n;
  if (lane == 0) {
    p.stats[2 * wg] = mean;
    p.stats[2 * wg + 1] = 1.0f / sqrtf(var + p.eps);
  }
}

cudaError_t launch_groupnorm_stats(const NormStatsParams& p, cudaStream_t s) {
  const long long warps = (long long)p.rows * p.groups;
  if (warps <= 0) return cudaSuccess;
  groupnorm_stats_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, s>>>(p);
  return cudaGetLastError();
}

// One warp per row (nn.LayerNorm statistics over the feature axis).
__global__ void rownorm_stats_kernel(const NormStatsParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= p.rows) return;
  const int C = p.c0;
  const float* src = p.src0 + (size_t)row * C;
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += __ldg(src + c);
  const float mean = warp_sum(sum) / (float)C;
  float sq = 0.f;
  for (int c = lane; c < C; c += 32) { const float dlt = __ldg(src + c) - mean; sq = fmaf(dlt, dlt, sq); }
  const float var = warp_sum(sq) / (float)C;
  if (lane == 0) {
    p.stats[2 * row] = mean;
    p.stats[2 * row + 1] = 1.0f / sqrtf(var + p.eps);
  }
}

cudaError_t launch_rownorm_stats(const NormStatsParams& p, cudaStream_t s) {
  if (p.rows <= 0) return cudaSuccess;
  rownorm_stats_kernel<<<(unsigned)((p.rows + 7) / 8), 256, 0, s>>>(p);
  return cudaGetLastError();
}

// ================================================================================================
// ConvTranspose1d(k = 2f, stride f, pad f/2) gather stage (modules.py:74-81): Y = X @ Wp^T was
// computed by the GEMM with Wp[k * Cout + co][ci]; every output position has exactly two taps.
// ================================================================================================
__global__ void upsample_gather_kernel(const float* __restrict__ Y, const float* __restrict__ bias,
                                       const float* __restrict__ add, float* __restrict__ out, int B, int Lin,
                                       int Cout, int f) {
  pdl_enter();
  const int c4 = Cout >> 2;
  const long long total = (long long)B * Lin * f * c4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = (int)(idx % c4) * 4;
  const long long t = idx / c4;
  const int Lout = Lin * f;
  const int o = (int)(t % Lout), b = (int)(t / Lout);
  const int pad = f / 2;
  const int i0 = (o + pad) / f, k0 = (o + pad) - i0 * f;
  const int ldy = 2 * f * Cout;
  float4 acc = *reinterpret_cast<const float4*>(bias + co);
  if (i0 < Lin) {
    const float4 y = *reinterpret_cast<const float4*>(Y + ((size_t)b * Lin + i0) * ldy + (size_t)k0 * Cout + co);
    acc.x += y.x; acc.y += y.y; acc.z += y.z; acc.w += y.w;
  }
  if (i0 - 1 >= 0) {
    const float4 y = *reinterpret_cast<const float4*>(Y + ((size_t)b * Lin + i0 - 1) * ldy + (size_t)(k0 + f) * Cout + co);
    acc.x += y.x; acc.y += y.y; acc.z += y.z; acc.w += y.w;
  }
  const size_t oi = ((size_t)b * Lout + o) * Cout + co;
  if (add) {
    const float4 y = *reinterpret_cast<const float4*>(add + oi);
    acc.x += y.x; acc.y += y.y; acc.z += y.z; acc.w += y.w;
  }
  *reinterpret_cast<float4*>(out + oi) = acc;
}

cudaError_t launch_upsample_gather(const float* Y, const float* bias, const float* add, float* out, int B, int Lin,
                                   int Cout, int f, cudaStream_t s) {
  if (Cout % 4) return cudaErrorInvalidValue;
  const long long total = (long long)B * Lin * f * (Cout / 4);
  if (total <= 0) return cudaSuccess;
  MDT_CK_LAUNCH(launch_k_light(upsample_gather_kernel, (unsigned)((total + 255) / 256), 256, 0, s, Y, bias, add, out, B, Lin, Cout, f));
  return cudaGetLastError();
}

// Patcher: (b, c, l*p + q) -> (b, c*p + q, l); token-major: in[b][l*p + q][c] <-> out[b][l][c*p + q]
__global__ void patch_permute_kernel(const float* __restrict__ in, float* __restrict__ out, long long total, int L,
                                     int C, int p, int to_patched) {
  pdl_enter();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // idx enumerates the patched tensor [b][l][c*p + q]
  const int cp = C * p;
  const int j = (int)(idx % cp);
  const long long t = idx / cp;
  const int l = (int)(t % L);
  const long long b = t / L;
  const int c = j / p, q = j - c * p;
  const long long un = ((b * L + l) * p + q) * C + c;
  if (to_patched) out[idx] = in[un];
  else out[un] = in[idx];
}

cudaError_t launch_patch_permute(const float* in, float* out, int B, int L, int C, int p, int to_patched,
                                 cudaStream_t s) {
  const long long total = (long long)B * L * C * p;
  if (total <= 0) return cudaSuccess;
  MDT_CK_LAUNCH(launch_k_light(patch_permute_kernel, (unsigned)((total + 255) / 256), 256, 0, s, in, out, total, L, C, p, to_patched));
  return cudaGetLastError();
}

// ================================================================================================
// Conditioning encoder, time features, FiLM fold (all tiny; once per sample() call).
// ================================================================================================
__global__ void encode_cond_kernel(const float* __restrict__ seq, const float* __restrict__ w,
                                   const float* __restrict__ bias, const float* __restrict__ inv_freq,
                                   float* __restrict__ emb, long long total, int n, int text_dim, int pos_dim, int add) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int F = add ? text_dim : text_dim + pos_dim;
  const int f = (int)(idx % F);
  const long long t = idx / F;
  const int i = (int)(t % n);
  const int half = pos_dim / 2;
  auto pe = [&](int j) -> float {
    const float ang = (float)i * __ldg(inv_freq + (j < half ? j : j - half));
    return j < half ? sinf(ang) : cosf(ang);
  };
  float v;
  if (f < text_dim) {
    v = gelu_f(fmaf(__ldg(seq + t), __ldg(w + f), __ldg(bias + f)));
    if (add && f < pos_dim) v += pe(f);
  } else {
    v = pe(f - text_dim);
  }
  emb[idx] = v;
}

cudaError_t launch_encode_cond(const float* seq, const float* w, const float* bias, const float* inv_freq,
                               float* emb, int B, int n, int text_dim, int pos_dim, int add, cudaStream_t s) {
  const int F = add ? text_dim : text_dim + pos_dim;
  const long long total = (long long)B * n * F;
  if (total <= 0) return cudaSuccess;
  encode_cond_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(seq, w, bias, inv_freq, emb, total, n,
                                                                      text_dim, pos_dim, add);
  return cudaGetLastError();
}

__global__ void time_features_kernel(const float* __restrict__ t, const float* __restrict__ w,
                                     float* __restrict__ out, int rows, int half) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int dim = 2 * half + 1;
  if (idx >= rows * dim) return;
  const int r = idx / dim, j = idx - r * dim;
  const float tv = __ldg(t + r);
  float v;
  if (j == 0) v = tv;
  else {
    const int jj = (j - 1) < half ? (j - 1) : (j - 1 - half);
    float fr = tv * __ldg(w + jj);   // freqs = x * weights * 2 * pi, evaluated left to right in fp32
    fr = fr * 2.0f;
    fr = fr * 3.14159274101257324f;
    v = (j - 1) < half ? sinf(fr) : cosf(fr);
  }
  out[idx] = v;
}

cudaError_t launch_time_features(const float* t, const float* w, float* out, int rows, int half, cudaStream_t s) {
  const int total = rows * (2 * half + 1);
  if (total <= 0) return cudaSuccess;
  MDT_CK_LAUNCH(launch_k_light(time_features_kernel, (total + 127) / 128, 128, 0, s, t, w, out, rows, half));
  return cudaGetLastError();
}

__global__ void film_fold_kernel(const float* __restrict__ ss, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float* __restrict__ aff, int rows, int C) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * C) return;
  const int r = idx / C, c = idx - r * C;
  const float sc = ss[(size_t)r * 2 * C + c] + 1.0f;
  const float sh = ss[(size_t)r * 2 * C + C + c];
  aff[(size_t)r * 2 * C + c] = gamma[c] * sc;
  aff[(size_t)r * 2 * C + C + c] = fmaf(beta[c], sc, sh);
}

cudaError_t launch_film_fold(const float* ss, const float* gamma, const float* beta, float* aff, int rows, int C,
                             cudaStream_t s) {
  const int total = rows * C;
  if (total <= 0) return cudaSuccess;
  MDT_CK_LAUNCH(launch_k_light(film_fold_kernel, (total + 255) / 256, 256, 0, s, ss, gamma, beta, aff, rows, C));
  return cudaGetLastError();
}

// ================================================================================================
// Sampler: Philox4x32-10 + fused ADPM2 / EDM / CFG update.
// ================================================================================================
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0; k.y += W1;
  }
  return c;
}

// Four standard normals for (seed, global sample, stream id, group-of-4 index): sharding invariant.
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long sample,
                                                 unsigned stream, unsigned group) {
  const uint4 r = philox4x32_10(make_uint4(group, stream, (unsigned)sample, (unsigned)(sample >> 32)),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const float k = 5.9604644775390625e-8f;  // 2^-24
  const float u0 = ((r.x >> 8) + 0.5f) * k, u1 = ((r.y >> 8) + 0.5f) * k;
  const float u2 = ((r.z >> 8) + 0.5f) * k, u3 = ((r.w >> 8) + 0.5f) * k;
  const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincosf(6.28318530717958647692f * u1, &s0, &c0);
  sincosf(6.28318530717958647692f * u3, &s1, &c1);
  return make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
}

// x0 = sigma_0 * noise (diffusion.py:520) and the first network input c_in * x (diffusion.py:811).
// One CTA per sample; injected noise arrives in the reference's (B,P,L) layout and is transposed via smem.
__global__ void step_init_kernel(const float* __restrict__ noise0, float* __restrict__ x, float* __restrict__ xin,
                                 const IterScalars* __restrict__ iters, unsigned long long seed,
                                 unsigned long long sample_offset, int B, int P, int L, int cfg, float sigma0) {
  extern __shared__ float tile[];  // [P][L+1]
  const int b = blockIdx.x;
  const int n = P * L;
  IterScalars it = iters[0];
  if (sigma0 >= 0.f) it.sigma = sigma0;
  if (noise0) {
    for (int e = threadIdx.x; e < n; e += blockDim.x) { const int pp = e / L, l = e - pp * L; tile[pp * (L + 1) + l] = noise0[(size_t)b * n + e]; }
    __syncthreads();
  }
  for (int g = threadIdx.x; g < n / 4; g += blockDim.x) {
    float nz[4];
    if (noise0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int e = g * 4 + j; const int l = e / P, pp = e - l * P; nz[j] = tile[pp * (L + 1) + l]; }
    } else {
      const float4 z = philox_normal4(seed, sample_offset + b, 0u, (unsigned)g);
      nz[0] = z.x; nz[1] = z.y; nz[2] = z.z; nz[3] = z.w;
    }
    float4 xv, xi;
    xv.x = it.sigma * nz[0]; xv.y = it.sigma * nz[1]; xv.z = it.sigma * nz[2]; xv.w = it.sigma * nz[3];
    xi.x = it.c_in_a * xv.x; xi.y = it.c_in_a * xv.y; xi.z = it.c_in_a * xv.z; xi.w = it.c_in_a * xv.w;
    const size_t o = (size_t)b * n + (size_t)g * 4;
    *reinterpret_cast<float4*>(x + o) = xv;
    *reinterpret_cast<float4*>(xin + o) = xi;
    if (cfg) *reinterpret_cast<float4*>(xin + (size_t)B * n + o) = xi;
  }
}

cudaError_t launch_step_init(const float* noise0, float* x, float* xin, const IterScalars* iters,
                             unsigned long long seed, unsigned long long sample_offset, int B, int P, int L,
                             int cfg, float sigma0, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  if ((P * L) % 4) return cudaErrorInvalidValue;
  const size_t smem = noise0 ? (size_t)P * (L + 1) * sizeof(float) : 0;
  if (smem > MDT_STEP_SMEM_MAX) return cudaErrorInvalidValue;
  step_init_kernel<<<B, 256, smem, s>>>(noise0, x, xin, iters, seed, sample_offset, B, P, L, cfg, sigma0);
  return cudaGetLastError();
}

// which = 0: after denoiser call A (sigma)      -> x_mid and the input of call B
// which = 1: after denoiser call B (sigma_mid)  -> x_next (+ ancestral noise) and the input of the next call A;
//            on the last iteration also the final (B,P,L) result and the argmax tokens.
// KDiffusion_mod.denoise_fn (diffusion.py:809-814), UNetCFG1d mix (modules.py:1253), ADPM2Sampler.step (diffusion.py:506-514).
template <int WHICH>
__global__ void step_update_kernel(const StepParams p) {
  pdl_enter();
  extern __shared__ float tile[];  // injected noise [P][L+1] (WHICH == 1 only)
  const int b = blockIdx.x;
  const int P = p.P, L = p.L, n = P * L;
  const int iter = (*p.call_idx) >> 1;
  const IterScalars it = p.iters[iter];
  const bool last = (iter == p.n_iters - 1);
  const size_t half = (size_t)p.B * n;
  const float* noise = p.run ? p.run->noise : p.noise;
  const long long noise_stride = p.run ? p.run->noise_iter_stride : p.noise_iter_stride;
  const float cond_scale = p.run ? p.run->cond_scale : p.cond_scale;
  const bool add_noise = !(p.karras && last);            // karras: the noise added here belongs to the next step; none after the last
  if (WHICH == 1 && noise && add_noise) {
    const float* nz = noise + (size_t)(iter + (p.karras ? 1 : 0)) * noise_stride + (size_t)b * n;
    for (int e = threadIdx.x; e < n; e += blockDim.x) { const int pp = e / L, l = e - pp * L; tile[pp * (L + 1) + l] = nz[e]; }
    __syncthreads();
  }
  const float c_skip = WHICH == 0 ? it.c_skip_a : it.c_skip_b;
  const float c_out = WHICH == 0 ? it.c_out_a : it.c_out_b;
  const float sig = WHICH == 0 ? it.sigma : it.sigma_mid;
  float c_in_next = it.c_in_b;
  if (WHICH == 1) c_in_next = last ? 0.f : p.iters[iter + 1].c_in_a;
  for (int g = threadIdx.x; g < n / 4; g += blockDim.x) {
    const size_t o = (size_t)b * n + (size_t)g * 4;
    const float4 nc = *reinterpret_cast<const float4*>(p.net + o);
    float pred[4] = {nc.x, nc.y, nc.z, nc.w};
    if (p.cfg) {
      const float4 nn = *reinterpret_cast<const float4*>(p.net + half + o);
      const float un[4] = {nn.x, nn.y, nn.z, nn.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) pred[j] = un[j] + (pred[j] - un[j]) * cond_scale;
    }
    const float4 xv4 = *reinterpret_cast<const float4*>(p.x + o);
    const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
    float cur[4];
    if (WHICH == 0) { cur[0] = xv[0]; cur[1] = xv[1]; cur[2] = xv[2]; cur[3] = xv[3]; }
    else { const float4 m = *reinterpret_cast<const float4*>(p.xmid + o); cur[0] = m.x; cur[1] = m.y; cur[2] = m.z; cur[3] = m.w; }
    float res[4];
    float nz[4] = {0.f, 0.f, 0.f, 0.f};
    float dslope[4] = {0.f, 0.f, 0.f, 0.f};
    if (WHICH == 1 && p.karras) { const float4 d4 = *reinterpret_cast<const float4*>(p.daux + o); dslope[0] = d4.x; dslope[1] = d4.y; dslope[2] = d4.z; dslope[3] = d4.w; }
    if (WHICH == 1 && add_noise) {
      if (noise) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const int e = g * 4 + j; const int l = e / P, pp = e - l * P; nz[j] = tile[pp * (L + 1) + l]; }
      } else {
        const unsigned long long sd = p.run ? p.run->seed : p.seed, so = p.run ? p.run->sample_offset : p.sample_offset;
        const float4 z = philox_normal4(sd, so + b, p.noise_stream >= 0 ? (unsigned)p.noise_stream : (unsigned)(iter + 1 + (p.karras ? 1 : 0)), (unsigned)g);
        nz[0] = z.x; nz[1] = z.y; nz[2] = z.z; nz[3] = z.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x0 = c_skip * cur[j] + c_out * pred[j];
      x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
      const float dd = (cur[j] - x0) / sig;
      if (WHICH == 0) { res[j] = xv[j] + dd * it.dt_mid; dslope[j] = dd; }
      else if (p.karras) res[j] = (xv[j] + it.dt_down * (dslope[j] + dd)) + nz[j] * it.sigma_up;   // diffusion.py:433, then the next step's x_hat
      else res[j] = (xv[j] + dd * it.dt_down) + nz[j] * it.sigma_up;
    }
    if (p.karras && WHICH == 0) *reinterpret_cast<float4*>(p.daux + o) = make_float4(dslope[0], dslope[1], dslope[2], dslope[3]);
    const float4 r4 = make_float4(res[0], res[1], res[2], res[3]);
    if (WHICH == 0) *reinterpret_cast<float4*>(p.xmid + o) = r4;
    else *reinterpret_cast<float4*>(p.x + o) = r4;
    if (!(WHICH == 1 && last)) {
      const float4 xi = make_float4(c_in_next * res[0], c_in_next * res[1], c_in_next * res[2], c_in_next * res[3]);
      *reinterpret_cast<float4*>(p.xin + o) = xi;
      if (p.cfg) *reinterpret_cast<float4*>(p.xin + half + o) = xi;
    }
  }
}

cudaError_t launch_step_update(int which, const StepParams& p, cudaStream_t s) {
  if (p.B <= 0) return cudaSuccess;
  if ((p.P * p.L) % 4) return cudaErrorInvalidValue;
  if (which == 0) MDT_CK_LAUNCH(launch_k_light(step_update_kernel<0>, p.B, 256, 0, s, p));
  else {
    const bool tile = p.run ? p.has_noise != 0 : p.noise != nullptr;
    const size_t smem = tile ? (size_t)p.P * (p.L + 1) * sizeof(float) : 0;
    if (smem > MDT_STEP_SMEM_MAX) return cudaErrorInvalidValue;
    MDT_CK_LAUNCH(launch_k_light(step_update_kernel<1>, p.B, 256, smem, s, p));
  }
  return cudaGetLastError();
}

__global__ void set_run_params_kernel(RunParams* dst, const RunParams v) { *dst = v; }
cudaError_t launch_set_run_params(RunParams* dst, const RunParams& v, cudaStream_t s) {
  set_run_params_kernel<<<1, 1, 0, s>>>(dst, v);
  return cudaGetLastError();
}
__global__ void set_int_kernel(int* dst, int v) { *dst = v; }
__global__ void add_int_kernel(int* dst, int v) { *dst += v; }
cudaError_t launch_set_int(int* dst, int v, cudaStream_t s) { set_int_kernel<<<1, 1, 0, s>>>(dst, v); return cudaGetLastError(); }
cudaError_t launch_add_int(int* dst, int v, cudaStream_t s) { add_int_kernel<<<1, 1, 0, s>>>(dst, v); return cudaGetLastError(); }

// (B,P,L) -> token-major [B*L][P] scaled by mul; dup != 0 also writes the null-branch half.
__global__ void to_token_major_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int P, int L,
                                      float mul, int dup) {
  pdl_enter();
  const long long total = (long long)B * P * L;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int pp = (int)(idx % P);
  const long long t = idx / P;
  const int l = (int)(t % L);
  const long long b = t / L;
  const float v = in[(b * P + pp) * L + l] * mul;
  out[idx] = v;
  if (dup) out[total + idx] = v;
}

cudaError_t launch_to_token_major(const float* in, float* out, int B, int P, int L, float mul, int dup,
                                  cudaStream_t s) {
  const long long total = (long long)B * P * L;
  if (total <= 0) return cudaSuccess;
  MDT_CK_LAUNCH(launch_k_light(to_token_major_kernel, (unsigned)((total + 255) / 256), 256, 0, s, in, out, B, P, L, mul, dup));
  return cudaGetLastError();
}

// token-major network output -> (B,P,L) with the classifier-free mix (modules.py:1253); also used as the
// final transposition of the sampler state (cfg = 0, cond_scale ignored) with optional clamp via mul trick off.
__global__ void cfg_mix_to_bpl_kernel(const float* __restrict__ net, float* __restrict__ out, int B, int P, int L,
                                      float cond_scale, int cfg) {
  pdl_enter();
  const long long total = (long long)B * P * L;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int l = (int)(idx % L);
  const long long t = idx / L;
  const int pp = (int)(t % P);
  const long long b = t / P;
  const long long src = (b * L + l) * P + pp;
  float v = net[src];
  if (cfg) { const float u = net[total + src]; v = u + (v - u) * cond_scale; }
  out[idx] = v;
}

cudaError_t launch_cfg_mix_to_bpl(const float* net, float* out, int B, int P, int L, float cond_scale, int cfg,
                                  cudaStream_t s) {
  const long long total = (long long)B * P * L;
  if (total <= 0) return cudaSuccess;
  MDT_CK_LAUNCH(launch_k_light(cfg_mix_to_bpl_kernel, (unsigned)((total + 255) / 256), 256, 0, s, net, out, B, P, L, cond_scale, cfg));
  return cudaGetLastError();
}

// Final hand-back: sampler state (token-major) -> reference layout (B,P,L) with the optional final clamp
// (diffusion.py:590), plus argmax tokens over the class axis (generative.py:1212-1213; first max wins).
__global__ void finalize_kernel(const float* __restrict__ x, float* __restrict__ out, unsigned char* __restrict__ tokens,
                                int P, int L, int clamp) {
  extern __shared__ float tile[];  // [L][P+1]
  const int b = blockIdx.x, n = P * L;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int l = e / P, pp = e - l * P;
    float v = x[(size_t)b * n + e];
    if (clamp) v = fminf(fmaxf(v, -1.0f), 1.0f);
    tile[l * (P + 1) + pp] = v;
  }
  __syncthreads();
  if (out) {
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      const int pp = e / L, l = e - pp * L;
      out[(size_t)b * n + e] = tile[l * (P + 1) + pp];
    }
  }
  if (tokens) {
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      float best = tile[l * (P + 1)];
      int bi = 0;
      for (int pp = 1; pp < P; ++pp) { const float v = tile[l * (P + 1) + pp]; if (v > best) { best = v; bi = pp; } }
      tokens[(size_t)b * L + l] = (unsigned char)bi;
    }
  }
}

cudaError_t launch_finalize(const float* x, float* out, unsigned char* tokens, int B, int P, int L, int clamp,
                            cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  if ((size_t)L * (P + 1) * sizeof(float) > MDT_STEP_SMEM_MAX) return cudaErrorInvalidValue;
  finalize_kernel<<<B, 256, (size_t)L * (P + 1) * sizeof(float), s>>>(x, out, tokens, P, L, clamp);
  return cudaGetLastError();
}

// KarrasSampler step 0 (diffusion.py:425-426 on x = sigma_0 * noise): x_hat = x + scale * eps_0, next network input c_in * x_hat.
__global__ void karras_prenoise_kernel(float* __restrict__ x, float* __restrict__ xin, const float* __restrict__ noise, float scale,
                                       float c_in, unsigned long long seed, unsigned long long sample_offset, int B, int P, int L, int cfg) {
  const int b = blockIdx.x, n = P * L;
  for (int g = threadIdx.x; g < n / 4; g += blockDim.x) {
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!noise) z = philox_normal4(seed, sample_offset + b, 1u, (unsigned)g);
    const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = g * 4 + j;                 // token-major element: e = l * P + p
      const int l = e / P, pp = e - l * P;
      const size_t tm = (size_t)b * n + e;
      const float nzv = noise ? noise[(size_t)b * n + (size_t)pp * L + l] : zz[j];
      const float v = x[tm] + scale * nzv;
      x[tm] = v;
      xin[tm] = c_in * v;
      if (cfg) xin[(size_t)B * n + tm] = c_in * v;
    }
  }
}

cudaError_t launch_karras_prenoise(float* x, float* xin, const float* noise, float scale, float c_in, unsigned long long seed,
                                   unsigned long long sample_offset, int B, int P, int L, int cfg, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  if ((P * L) % 4) return cudaErrorInvalidValue;
  karras_prenoise_kernel<<<B, 256, 0, s>>>(x, xin, noise, scale, c_in, seed, sample_offset, B, P, L, cfg);
  return cudaGetLastError();
}

static cudaError_t init_step_kernels() {
  cudaError_t e = cudaFuncSetAttribute(step_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MDT_STEP_SMEM_MAX);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(step_update_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MDT_STEP_SMEM_MAX);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MDT_STEP_SMEM_MAX);
  return e;
}

// ---- inpainting (ADPM2Sampler.inpaint, diffusion.py:526-549) -----------------------------------------------------------
// mode 0: x = sigma * n                                   (initial state, diffusion.py:535)
// mode 1: x = mask ? source + sigma * n : x               (merge of the re-noised source, diffusion.py:539-542) + network input
// mode 2: x = x + sigma * n                               (re-noise between resamples, diffusion.py:545-547)
// mode 3: out(B,P,L) = mask ? source : x                  (diffusion.py:549)
// n is injected noise in the reference (B,P,L) layout or Philox(seed, sample, stream, element); x / xin are token-major.
__global__ void inpaint_kernel(int mode, float* __restrict__ x, float* __restrict__ xin, const float* __restrict__ source,
                               const unsigned char* __restrict__ mask, const float* __restrict__ noise, float sigma, float c_in,
                               unsigned long long seed, unsigned long long sample_offset, int stream, int B, int P, int L, int cfg,
                               float* __restrict__ out) {
  const int b = blockIdx.x, n = P * L;
  for (int g = threadIdx.x; g < n / 4; g += blockDim.x) {
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mode != 3 && !noise) z = philox_normal4(seed, sample_offset + b, (unsigned)stream, (unsigned)g);
    const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = g * 4 + j;                 // token-major element: e = l * P + p
      const int l = e / P, pp = e - l * P;
      const size_t bpl = (size_t)b * n + (size_t)pp * L + l;
      const size_t tm = (size_t)b * n + e;
      const float nz = (mode != 3 && noise) ? noise[bpl] : zz[j];
      float v;
      if (mode == 0) v = sigma * nz;
      else if (mode == 1) v = mask[bpl] ? source[bpl] + sigma * nz : x[tm];
      else if (mode == 2) v = x[tm] + sigma * nz;
      else { out[bpl] = mask[bpl] ? source[bpl] : x[tm]; continue; }
      x[tm] = v;
      if (mode == 1) {
        xin[tm] = c_in * v;
        if (cfg) xin[(size_t)B * n + tm] = c_in * v;
      }
    }
  }
}

cudaError_t launch_inpaint(int mode, float* x, float* xin, const float* source, const unsigned char* mask, const float* noise,
                           float sigma, float c_in, unsigned long long seed, unsigned long long sample_offset, int stream, int B,
                           int P, int L, int cfg, float* out, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  if ((P * L) % 4) return cudaErrorInvalidValue;
  inpaint_kernel<<<B, 256, 0, s>>>(mode, x, xin, source, mask, noise, sigma, c_in, seed, sample_offset, stream, B, P, L, cfg, out);
  return cudaGetLastError();
}

// ---- cross-attention K/V cache in mma.sync B-fragment order (gemm_attn.cu, packed path) ------------------------------------------
// out[(b * heads + h)][which = K|V][f = 0..15][lane] = uint2(b0, b1), the two B-fragment registers lane `lane` feeds to
// mma.m16n8k8 for fragment f:   K: f = kstep * 2 + ntile   b0 = K[8 ntile + g][8 kstep + q],  b1 = K[..][8 kstep + q + 4]
//                               V: f = ktile * 8 + ntile   b0 = V[8 ktile + 2q][8 ntile + g], b1 = V[8 ktile + 2q + 1][8 ntile + g]
// (g = lane / 4, q = lane % 4; the V key permutation matches the P -> A-fragment reuse in attn_math.cuh).  Keys >= nk read as 0.
// Values are rounded to tf32.  One 256-byte coalesced load per fragment replaces the cp.async staging + shared-memory fragment loads.
__global__ void kv_fragment_pack_kernel(const float* __restrict__ kv, uint2* __restrict__ out, long long B, int nk, int heads, int d, int kperm) {
  const long long bh = blockIdx.x;
  const long long b = bh / heads;
  const int h = (int)(bh - b * heads);
  if (b >= B) return;
  const int ldkv = 2 * heads * d;
  const float* base = kv + (size_t)b * nk * ldkv + (size_t)h * d;
  if (kperm == 2) {
    // f16 m16n8k16 B fragments: K: f = kstep16 * 2 + ntile, b0 = (K[8 ntile + g][16 ks + 2q], [.. + 1]), b1 = the pair 8 features on;
    // V: f = feature tile n, b0 = (V[2q][8n + g], V[2q + 1][8n + g]), b1 = the pair of keys 8 on.  512 uint2 per block.
    const float* v = base + heads * d;
    for (int idx = threadIdx.x; idx < 512; idx += blockDim.x) {
      const int which = idx >> 8, f = (idx >> 5) & 7, lane = idx & 31, g = lane >> 2, q = lane & 3;
      float e[4] = {0.f, 0.f, 0.f, 0.f};
      if (which == 0) {
        const int key = (f & 1) * 8 + g, c0 = (f >> 1) * 16 + 2 * q;
        if (key < nk) { e[0] = base[(size_t)key * ldkv + c0]; e[1] = base[(size_t)key * ldkv + c0 + 1]; e[2] = base[(size_t)key * ldkv + c0 + 8]; e[3] = base[(size_t)key * ldkv + c0 + 9]; }
      } else {
        const int col = f * 8 + g;
        const int ks[4] = {2 * q, 2 * q + 1, 2 * q + 8, 2 * q + 9};
        for (int i = 0; i < 4; ++i) if (ks[i] < nk) e[i] = v[(size_t)ks[i] * ldkv + col];
      }
      uint32_t r0, r1;
      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r0) : "f"(e[1]), "f"(e[0]));
      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r1) : "f"(e[3]), "f"(e[2]));
      out[(size_t)bh * 1024 + idx] = make_uint2(r0, r1);
    }
    return;
  }
  for (int idx = threadIdx.x; idx < 1024; idx += blockDim.x) {
    const int which = idx >> 9, f = (idx >> 5) & 15, lane = idx & 31, g = lane >> 2, q = lane & 3;
    float b0 = 0.f, b1 = 0.f;
    if (which == 0) {
      // kperm: the q fragments come from tcgen05.ld.16x256b (thread (g, q) holds columns 2q, 2q + 1 of every 8-column group), so the
      // k index is relabelled k = q <-> column 2q, k = q + 4 <-> column 2q + 1 on both operands (gemm_attn_frag.cu)
      const int key = (f & 1) * 8 + g, c0 = (f >> 1) * 8 + (kperm ? 2 * q : q), c1 = c0 + (kperm ? 1 : 4);
      if (key < nk) { b0 = base[(size_t)key * ldkv + c0]; b1 = base[(size_t)key * ldkv + c1]; }
    } else {
      const int j0 = (f >> 3) * 8 + 2 * q, col = (f & 7) * 8 + g;
      const float* v = base + heads * d;
      if (j0 < nk) b0 = v[(size_t)j0 * ldkv + col];
      if (j0 + 1 < nk) b1 = v[(size_t)(j0 + 1) * ldkv + col];
    }
    uint32_t r0, r1;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r0) : "f"(b0));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r1) : "f"(b1));
    out[(size_t)bh * 1024 + idx] = make_uint2(r0, r1);
  }
}

cudaError_t launch_kv_fragment_pack(const float* kv, void* out, long long B, int nk, int heads, int d, int kperm, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  if (d != 64 || nk > 16 || nk < 1) return cudaErrorInvalidValue;
  kv_fragment_pack_kernel<<<(unsigned)(B * heads), 256, 0, s>>>(kv, reinterpret_cast<uint2*>(out), B, nk, heads, d, kperm);
  return cudaGetLastError();
}

// ---- token ids -> text bytes (generative.py:1069-1078: Keras sequences_to_texts drops ids without a vocabulary entry -- padding id 0
// among them -- and the reference strips the separating spaces).  One warp per row: ballot-compaction, zero padded to L, length out.
__global__ void decode_tokens_kernel(const uint8_t* __restrict__ tokens, const uint8_t* __restrict__ lut, uint8_t* __restrict__ out,
                                     int* __restrict__ lengths, long long B, int L) {
  __shared__ uint8_t s_lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const uint8_t* src = tokens + row * L;
  uint8_t* dst = out + row * L;
  int base = 0;
  for (int i0 = 0; i0 < L; i0 += 32) {
    const int i = i0 + lane;
    const uint8_t ch = i < L ? s_lut[src[i]] : (uint8_t)0;
    const unsigned keep = __ballot_sync(0xffffffffu, ch != 0);
    if (ch != 0) dst[base + __popc(keep & ((1u << lane) - 1u))] = ch;
    base += __popc(keep);
  }
  for (int i = base + lane; i < L; i += 32) dst[i] = 0;
  if (lane == 0 && lengths) lengths[row] = base;
}

cudaError_t launch_decode_tokens(const uint8_t* tokens, const uint8_t* lut, uint8_t* out, int* lengths, long long B, int L, cudaStream_t s) {
  if (B <= 0 || L <= 0) return cudaSuccess;
  const int wpb = 8;
  decode_tokens_kernel<<<(unsigned)((B + wpb - 1) / wpb), wpb * 32, 0, s>>>(tokens, lut, out, lengths, B, L);
  return cudaGetLastError();
}

}  // namespace mdt
