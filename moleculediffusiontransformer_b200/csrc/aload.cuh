// aload.cuh -- device-side A-operand loader shared by the CUDA-core and the tcgen05 GEMM kernels.
#pragma once
#include "kernels.cuh"

namespace mdt {

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }                 // nn.SiLU
__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }  // nn.GELU() (erf form)

// exact-erf GELU evaluated with the Abramowitz-Stegun 7.1.26 rational form (|erf error| <= 1.5e-7, i.e. fp32 rounding level):
// one MUFU.EX2, one MUFU.RCP and a degree-5 Horner instead of the ~30-instruction erff.  Used only where the result is
// rounded to tf32 / bf16 right away (tensor-core modes); the fp32 mode keeps erff.
__device__ __forceinline__ float gelu_as(float v) {
  const float x = fabsf(v) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-x * x);          // erf(|v| / sqrt(2))
  return 0.5f * v * (1.0f + copysignf(e, v));
}

__device__ __forceinline__ const float* aload_aff(const ALoad& a) {
  if (!a.aff) return nullptr;
  int call = a.call_idx ? *a.call_idx : 0;
  return a.aff + (size_t)call * a.aff_call_stride;
}

// Four consecutive k (k % 4 == 0; requires C % 4 == 0, c0 % 4 == 0, cpg % 4 == 0 when grouped).
__device__ __forceinline__ float4 aload4(const ALoad& a, const float* __restrict__ aff, int b, int lo, int k) {
  int tap = 0, c = k;
  if (a.taps > 1) { tap = k / a.C; c = k - tap * a.C; }
  const int li = lo * a.stride + tap - a.pad;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (li < 0 || li >= a.L_in) return v;
  const size_t row = (size_t)b * a.L_in + li;
  if (c < a.c0) {
    v = __ldg(reinterpret_cast<const float4*>(a.src0 + row * a.c0 + c));
  } else {
    v = __ldg(reinterpret_cast<const float4*>(a.src1 + row * a.c1 + (c - a.c0)));
    v.x *= a.scale1; v.y *= a.scale1; v.z *= a.scale1; v.w *= a.scale1;
  }
  if (a.stats_mode) {
    const size_t si = (a.stats_mode == 1) ? row : ((size_t)b * a.groups + c / a.cpg);
    const float2 st = __ldg(reinterpret_cast<const float2*>(a.stats) + si);
    v.x = (v.x - st.x) * st.y; v.y = (v.y - st.x) * st.y; v.z = (v.z - st.x) * st.y; v.w = (v.w - st.x) * st.y;
  }
  if (aff) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(aff + c));
    const float4 h = __ldg(reinterpret_cast<const float4*>(aff + a.C + c));
    v.x = v.x * g.x + h.x; v.y = v.y * g.y + h.y; v.z = v.z * g.z + h.z; v.w = v.w * g.w + h.w;
  }
  if (a.silu) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
  return v;
}

// Scalar variant for layers whose channel count is not a multiple of 4 (pred_dim = 1, 22, ...).
__device__ __forceinline__ float aload1(const ALoad& a, const float* __restrict__ aff, int b, int lo, int k) {
  int tap = 0, c = k;
  if (a.taps > 1) { tap = k / a.C; c = k - tap * a.C; }
  const int li = lo * a.stride + tap - a.pad;
  if (li < 0 || li >= a.L_in) return 0.f;
  const size_t row = (size_t)b * a.L_in + li;
  float v = (c < a.c0) ? __ldg(a.src0 + row * a.c0 + c) : __ldg(a.src1 + row * a.c1 + (c - a.c0)) * a.scale1;
  if (a.stats_mode) {
    const size_t si = (a.stats_mode == 1) ? row : ((size_t)b * a.groups + c / a.cpg);
    const float2 st = __ldg(reinterpret_cast<const float2*>(a.stats) + si);
    v = (v - st.x) * st.y;
  }
  if (aff) v = v * __ldg(aff + c) + __ldg(aff + a.C + c);
  if (a.silu) v = silu_f(v);
  return v;
}

__host__ __device__ inline bool aload_vec4_ok(const ALoad& a) {
  return (a.C % 4 == 0) && (a.c0 % 4 == 0) && (a.stats_mode != 2 || a.cpg % 4 == 0);
}

}  // namespace mdt
