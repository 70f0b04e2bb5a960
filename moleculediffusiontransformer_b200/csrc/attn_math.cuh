// attn_math.cuh -- per-(sample, head) softmax attention on operands staged in shared memory.
// Shared by the streaming attention kernel (attention_bulk.cu) and the fused projection+attention kernel (gemm_attn.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include "kernels.cuh"
#include "tc_common.cuh"

namespace mdt {

template <int KIND> struct SmemIO;
template <> struct SmemIO<0> {
  typedef float T;
  static __device__ __forceinline__ float4 ld4(const T* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ float ld1(const T* p) { return *p; }
  static __device__ __forceinline__ void st(void* o, size_t i, float v) { reinterpret_cast<float*>(o)[i] = v; }
  static __device__ __forceinline__ void st2(void* o, size_t i, float a, float b) { *reinterpret_cast<float2*>(reinterpret_cast<float*>(o) + i) = make_float2(a, b); }
};
template <> struct SmemIO<1> {
  typedef float T;
  static __device__ __forceinline__ float4 ld4(const T* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ float ld1(const T* p) { return *p; }
  static __device__ __forceinline__ void st(void* o, size_t i, float v) { reinterpret_cast<uint32_t*>(o)[i] = tc::to_tf32(v); }
  static __device__ __forceinline__ void st2(void* o, size_t i, float a, float b) { *reinterpret_cast<uint2*>(reinterpret_cast<uint32_t*>(o) + i) = make_uint2(tc::to_tf32(a), tc::to_tf32(b)); }
};
template <> struct SmemIO<2> {
  typedef __nv_bfloat16 T;
  static __device__ __forceinline__ float4 ld4(const T* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ float ld1(const T* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(void* o, size_t i, float v) { reinterpret_cast<__nv_bfloat16*>(o)[i] = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ void st2(void* o, size_t i, float a, float b) { *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(o) + i) = tc::pack_bf16(a, b); }
};

template <> struct SmemIO<3> {
  typedef __half T;
  static __device__ __forceinline__ float4 ld4(const T* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ float ld1(const T* p) { return __half2float(*p); }
  static __device__ __forceinline__ void st(void* o, size_t i, float v) { reinterpret_cast<uint16_t*>(o)[i] = (uint16_t)(tc::pack_f16s(v, 0.f) & 0xffffu); }
  static __device__ __forceinline__ void st2(void* o, size_t i, float a, float b) { *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(o) + i) = tc::pack_f16s(a, b); }
};

// One (sample, head): q rows at qs (stride sa), k rows at ks and v rows at vs (stride sb); ss = nq x (nk + 1) scratch.
template <int KIND>
__device__ __forceinline__ void attend_head(const typename SmemIO<KIND>::T* qs, int sa, const typename SmemIO<KIND>::T* ks,
                                            const typename SmemIO<KIND>::T* vs, int sb, float* ss, int nq, int nk, int d,
                                            float scale, void* out, size_t out_base, int ldo, int lane) {
  typedef SmemIO<KIND> IO;
  const int d4 = d >> 2;
  const int nbi = (nq + 1) >> 1, nbj = (nk + 3) >> 2;
  for (int blk = lane; blk < nbi * nbj; blk += 32) {
    const int i0 = (blk / nbj) * 2, j0 = (blk % nbj) * 4;
    const int i1 = min(i0 + 1, nq - 1);
    const int jj[4] = {j0, min(j0 + 1, nk - 1), min(j0 + 2, nk - 1), min(j0 + 3, nk - 1)};
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int c = 0; c < d4; ++c) {
      const float4 qa = IO::ld4(qs + i0 * sa + c * 4);
      const float4 qb = IO::ld4(qs + i1 * sa + c * 4);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4 kk = IO::ld4(ks + jj[t] * sb + c * 4);
        acc[0][t] = fmaf(qa.x, kk.x, acc[0][t]); acc[0][t] = fmaf(qa.y, kk.y, acc[0][t]);
        acc[0][t] = fmaf(qa.z, kk.z, acc[0][t]); acc[0][t] = fmaf(qa.w, kk.w, acc[0][t]);
        acc[1][t] = fmaf(qb.x, kk.x, acc[1][t]); acc[1][t] = fmaf(qb.y, kk.y, acc[1][t]);
        acc[1][t] = fmaf(qb.z, kk.z, acc[1][t]); acc[1][t] = fmaf(qb.w, kk.w, acc[1][t]);
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (j0 + t < nk) {
        ss[i0 * (nk + 1) + j0 + t] = acc[0][t] * scale;
        if (i0 + 1 < nq) ss[(i0 + 1) * (nk + 1) + j0 + t] = acc[1][t] * scale;
      }
    }
  }
  __syncwarp();
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
  // O = P V, d == 64: lane owns features lane and lane + 32; 8 query rows x 16 keys per register tile
  for (int i0 = 0; i0 < nq; i0 += 8) {
    float acc[8][2];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) { acc[ii][0] = 0.f; acc[ii][1] = 0.f; }
    for (int j0 = 0; j0 < nk; j0 += 16) {
      float vr[16][2];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const bool ok = j0 + t < nk;
        vr[t][0] = ok ? IO::ld1(vs + (j0 + t) * sb + lane) : 0.f;
        vr[t][1] = ok ? IO::ld1(vs + (j0 + t) * sb + 32 + lane) : 0.f;
      }
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
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
    for (int ii = 0; ii < 8; ++ii) {
      if (i0 + ii < nq) {
        const size_t o = out_base + (size_t)(i0 + ii) * ldo;
        IO::st(out, o + lane, acc[ii][0]);
        IO::st(out, o + 32 + lane, acc[ii][1]);
      }
    }
  }
  __syncwarp();
}

// ---- warp-level tensor-core variant for the tensor-core precisions (operands are already tf32 / bf16 values) ----
// S = Q K^T and O = P V with mma.sync.m16n8k8 (tf32 inputs, fp32 accumulate).  The row padding of 4 floats makes
// every fragment load hit 32 distinct banks.  P (C-fragment layout) feeds the second MMA as an A fragment by
// permuting the key index consistently on both operands (column q <-> key 8t + 2q, column q + 4 <-> key 8t + 2q + 1).
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NT = number of 8-key tiles held in registers (compile time, so no predicated-off MMA slots are issued).
template <int SK, int OK, int NT>
__device__ __forceinline__ void attend_head_mma_nt(const typename SmemIO<SK>::T* qs, int sa, const typename SmemIO<SK>::T* ks,
                                                   const typename SmemIO<SK>::T* vs, int sb, int nq, int nk, float scale, void* out,
                                                   size_t out_base, int ldo, int lane, int blk = 0) {
  // blk > 0 (power of two): the nq = nk <= 16 rows are 16 / blk short samples packed into one tile; a query sees only the keys of
  // its own block (block-diagonal mask), so one m16 tile serves all of them
  typedef SmemIO<SK> IO;
  typedef SmemIO<OK> OUT;
  const int g = lane >> 2, q = lane & 3;
  const int bm = blk > 0 ? ~(blk - 1) : 0;
  for (int i0 = 0; i0 < nq; i0 += 16) {
    const int r0 = min(i0 + g, nq - 1), r1 = min(i0 + g + 8, nq - 1);
    float sc[NT][4], sd[NT][4];   // two accumulator sets (even / odd k-steps): halves the dependent MMA chain
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      sc[t][0] = 0.f; sc[t][1] = 0.f; sc[t][2] = 0.f; sc[t][3] = 0.f;
      sd[t][0] = 0.f; sd[t][1] = 0.f; sd[t][2] = 0.f; sd[t][3] = 0.f;
    }
    int jr[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) jr[t] = min(t * 8 + g, nk - 1) * sb;
#pragma unroll
    for (int k0 = 0; k0 < 64; k0 += 8) {
      uint32_t a[4];
      a[0] = __float_as_uint(IO::ld1(qs + r0 * sa + k0 + q));
      a[1] = __float_as_uint(IO::ld1(qs + r1 * sa + k0 + q));
      a[2] = __float_as_uint(IO::ld1(qs + r0 * sa + k0 + q + 4));
      a[3] = __float_as_uint(IO::ld1(qs + r1 * sa + k0 + q + 4));
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const uint32_t b0 = __float_as_uint(IO::ld1(ks + jr[t] + k0 + q));
        const uint32_t b1 = __float_as_uint(IO::ld1(ks + jr[t] + k0 + q + 4));
        if ((k0 >> 3) & 1) mma_tf32_16x8x8(sd[t], a, b0, b1);
        else mma_tf32_16x8x8(sc[t], a, b0, b1);
      }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) { sc[t][0] += sd[t][0]; sc[t][1] += sd[t][1]; sc[t][2] += sd[t][2]; sc[t][3] += sd[t][3]; }
    // softmax over keys: thread holds columns 8t + 2q, 8t + 2q + 1 of rows g (c0, c1) and g + 8 (c2, c3)
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int j = t * 8 + 2 * q;
      const bool in0 = ((j ^ g) & bm) == 0, in1 = ((j ^ (g + 8)) & bm) == 0;   // j and j + 1 share a block (blk >= 2)
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
    for (int t = 0; t < NT; ++t) {
      // tensor-core modes only (operands are already tf32 / bf16): ex2.approx is ~2 ulp, far below operand rounding
      sc[t][0] = __expf(sc[t][0] - m0); sc[t][1] = __expf(sc[t][1] - m0);
      sc[t][2] = __expf(sc[t][2] - m1); sc[t][3] = __expf(sc[t][3] - m1);
      s0 += sc[t][0] + sc[t][1]; s1 += sc[t][2] + sc[t][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float inv0 = 1.0f / s0, inv1 = 1.0f / s1;
    // O = P V : 8 feature tiles of 8, one k-step per key tile
    float oc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { oc[n][0] = 0.f; oc[n][1] = 0.f; oc[n][2] = 0.f; oc[n][3] = 0.f; }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      uint32_t a[4];
      a[0] = tc::to_tf32(sc[t][0] * inv0);   // (row g,     key 8t + 2q)
      a[1] = tc::to_tf32(sc[t][2] * inv1);   // (row g + 8, key 8t + 2q)
      a[2] = tc::to_tf32(sc[t][1] * inv0);   // (row g,     key 8t + 2q + 1)
      a[3] = tc::to_tf32(sc[t][3] * inv1);   // (row g + 8, key 8t + 2q + 1)
      const int j0 = min(t * 8 + 2 * q, nk - 1) * sb, j1 = min(t * 8 + 2 * q + 1, nk - 1) * sb;   // masked keys carry p = 0
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const uint32_t b0 = __float_as_uint(IO::ld1(vs + j0 + n * 8 + g));
        const uint32_t b1 = __float_as_uint(IO::ld1(vs + j1 + n * 8 + g));
        mma_tf32_16x8x8(oc[n], a, b0, b1);
      }
    }
    const bool ok0 = i0 + g < nq, ok1 = i0 + g + 8 < nq;
    const size_t o0 = out_base + (size_t)(i0 + g) * ldo + 2 * q, o1 = o0 + (size_t)8 * ldo;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      if (ok0) OUT::st2(out, o0 + n * 8, oc[n][0], oc[n][1]);
      if (ok1) OUT::st2(out, o1 + n * 8, oc[n][2], oc[n][3]);
    }
  }
}

template <int SK, int OK>
__device__ __forceinline__ void attend_head_mma(const typename SmemIO<SK>::T* qs, int sa, const typename SmemIO<SK>::T* ks,
                                                const typename SmemIO<SK>::T* vs, int sb, int nq, int nk, float scale, void* out,
                                                size_t out_base, int ldo, int lane) {
  if (nk <= 8) attend_head_mma_nt<SK, OK, 1>(qs, sa, ks, vs, sb, nq, nk, scale, out, out_base, ldo, lane);
  else if (nk <= 16) attend_head_mma_nt<SK, OK, 2>(qs, sa, ks, vs, sb, nq, nk, scale, out, out_base, ldo, lane);
  else if (nk <= 32) attend_head_mma_nt<SK, OK, 4>(qs, sa, ks, vs, sb, nq, nk, scale, out, out_base, ldo, lane);
  else attend_head_mma_nt<SK, OK, 8>(qs, sa, ks, vs, sb, nq, nk, scale, out, out_base, ldo, lane);
}

// ---- packed cross-attention for short query blocks ----------------------------------------------------------------------------
// LQ = 4 or 8 queries per sample: G = 16 / LQ consecutive samples share one m16 tile (rows t * LQ .. of sample t), every sample
// brings its own context K / V.  K and V come straight from global memory in B-fragment order (kv_fragment_pack_kernel in
// kernels.cu: one coalesced 8-byte load per lane per fragment, no shared-memory staging); the G score blocks are independent MMA
// chains, the softmax runs once on the rows' own blocks, and P V accumulates the G samples into one output tile (rows of other
// samples enter as zeros).  nk <= 16.  kf[t] points at the 1024-uint2 [K | V] block of sample t and this tile's head.
template <int OK, int LQ>
__device__ __forceinline__ void attend_packed_cross(const float* qs, int sa, const uint2* const (&kf)[16 / LQ], int nk, float scale,
                                                    void* out, size_t out_base, int ldo, int rows_valid, int lane) {
  typedef SmemIO<OK> OUT;
  constexpr int G = 16 / LQ;
  const int g = lane >> 2, q = lane & 3;
  float sc[G][2][4];
#pragma unroll
  for (int t = 0; t < G; ++t)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) { sc[t][nt][0] = 0.f; sc[t][nt][1] = 0.f; sc[t][nt][2] = 0.f; sc[t][nt][3] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    uint32_t a[4];
    a[0] = __float_as_uint(qs[g * sa + ks * 8 + q]);
    a[1] = __float_as_uint(qs[(g + 8) * sa + ks * 8 + q]);
    a[2] = __float_as_uint(qs[g * sa + ks * 8 + q + 4]);
    a[3] = __float_as_uint(qs[(g + 8) * sa + ks * 8 + q + 4]);
#pragma unroll
    for (int t = 0; t < G; ++t)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const uint2 b = __ldg(kf[t] + (ks * 2 + nt) * 32 + lane);
        mma_tf32_16x8x8(sc[t][nt], a, b.x, b.y);
      }
  }
  // row g belongs to sample t0 = g / LQ, row g + 8 to sample t1 = (g + 8) / LQ: pick their score blocks
  const int t0 = g / LQ, t1 = (g + 8) / LQ;
  float s0[2][2], s1[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    if (LQ == 4) {
      s0[nt][0] = (g < 4) ? sc[0][nt][0] : sc[G > 1 ? 1 : 0][nt][0]; s0[nt][1] = (g < 4) ? sc[0][nt][1] : sc[G > 1 ? 1 : 0][nt][1];
      s1[nt][0] = (g < 4) ? sc[G > 2 ? 2 : 0][nt][2] : sc[G > 3 ? 3 : 0][nt][2];
      s1[nt][1] = (g < 4) ? sc[G > 2 ? 2 : 0][nt][3] : sc[G > 3 ? 3 : 0][nt][3];
    } else {
      s0[nt][0] = sc[0][nt][0]; s0[nt][1] = sc[0][nt][1];
      s1[nt][0] = sc[G > 1 ? 1 : 0][nt][2]; s1[nt][1] = sc[G > 1 ? 1 : 0][nt][3];
    }
  }
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int j = nt * 8 + 2 * q;
    s0[nt][0] = (j < nk) ? s0[nt][0] * scale : -INFINITY; s0[nt][1] = (j + 1 < nk) ? s0[nt][1] * scale : -INFINITY;
    s1[nt][0] = (j < nk) ? s1[nt][0] * scale : -INFINITY; s1[nt][1] = (j + 1 < nk) ? s1[nt][1] * scale : -INFINITY;
    m0 = fmaxf(m0, fmaxf(s0[nt][0], s0[nt][1]));
    m1 = fmaxf(m1, fmaxf(s1[nt][0], s1[nt][1]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float z0 = 0.f, z1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    s0[nt][0] = __expf(s0[nt][0] - m0); s0[nt][1] = __expf(s0[nt][1] - m0);
    s1[nt][0] = __expf(s1[nt][0] - m1); s1[nt][1] = __expf(s1[nt][1] - m1);
    z0 += s0[nt][0] + s0[nt][1]; z1 += s1[nt][0] + s1[nt][1];
  }
  z0 += __shfl_xor_sync(0xffffffffu, z0, 1); z0 += __shfl_xor_sync(0xffffffffu, z0, 2);
  z1 += __shfl_xor_sync(0xffffffffu, z1, 1); z1 += __shfl_xor_sync(0xffffffffu, z1, 2);
  const float inv0 = 1.0f / z0, inv1 = 1.0f / z1;
  uint32_t p0[2][2], p1[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    p0[nt][0] = tc::to_tf32(s0[nt][0] * inv0); p0[nt][1] = tc::to_tf32(s0[nt][1] * inv0);
    p1[nt][0] = tc::to_tf32(s1[nt][0] * inv1); p1[nt][1] = tc::to_tf32(s1[nt][1] * inv1);
  }
  float oc[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) { oc[n][0] = 0.f; oc[n][1] = 0.f; oc[n][2] = 0.f; oc[n][3] = 0.f; }
#pragma unroll
  for (int t = 0; t < G; ++t) {
    const bool own0 = t0 == t, own1 = t1 == t;
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      uint32_t a[4];
      a[0] = own0 ? p0[kt][0] : 0u;   // (row g,     key 8 kt + 2q)
      a[1] = own1 ? p1[kt][0] : 0u;   // (row g + 8, key 8 kt + 2q)
      a[2] = own0 ? p0[kt][1] : 0u;   // (row g,     key 8 kt + 2q + 1)
      a[3] = own1 ? p1[kt][1] : 0u;   // (row g + 8, key 8 kt + 2q + 1)
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const uint2 b = __ldg(kf[t] + 512 + (kt * 8 + n) * 32 + lane);
        mma_tf32_16x8x8(oc[n], a, b.x, b.y);
      }
    }
  }
  const bool ok0 = g < rows_valid, ok1 = g + 8 < rows_valid;
  const size_t o0 = out_base + (size_t)g * ldo + 2 * q, o1 = o0 + (size_t)8 * ldo;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    if (ok0) OUT::st2(out, o0 + n * 8, oc[n][0], oc[n][1]);
    if (ok1) OUT::st2(out, o1 + n * 8, oc[n][2], oc[n][3]);
  }
}

}  // namespace mdt
