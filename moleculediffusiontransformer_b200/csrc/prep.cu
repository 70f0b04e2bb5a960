// prep.cu -- operand preparation kernels for the TMA-fed GEMMs (HBM-bound, one pass):
//   gn_apply : GroupNorm statistics + normalise + per-channel affine (gamma/beta with FiLM folded) + SiLU over a
//              (possibly concatenated) fp32 token-major tensor, written once in the MMA operand dtype.
//              One CTA owns whole samples, so statistics never leave shared memory.
//   ln_apply : LayerNorm (no affine: gamma/beta are folded into the following projection) per row.
#include <cuda_bf16.h>
#include <stdlib.h>
#include "aload.cuh"
#include "tc_common.cuh"

namespace mdt {

template <int KIND>
__device__ __forceinline__ void store_op4(void* base, size_t idx, float4 v) {
  if (KIND == 1) {
    *reinterpret_cast<uint4*>(reinterpret_cast<float*>(base) + idx) = make_uint4(tc::to_tf32(v.x), tc::to_tf32(v.y), tc::to_tf32(v.z), tc::to_tf32(v.w));
  } else {
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + idx) = make_uint2(tc::pack_op2<KIND>(v.x, v.y), tc::pack_op2<KIND>(v.z, v.w));
  }
}

// SiLU with ex2.approx / rcp.approx (~2 ulp each): these kernels only feed tf32 / bf16 MMA operands (rounded to 11 / 8 bits right
// after), and the precise expf + division form was a third of their issue slots (ncu: issue-active 36 %, DRAM 33 %)
__device__ __forceinline__ float silu_op(float v) { return v * __fdividef(1.0f, 1.0f + __expf(-v)); }

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int KIND>
__global__ void __launch_bounds__(256) gn_apply_kernel(const GnApplyParams p, const int spc) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];
  const int C = p.c0 + p.c1, L = p.L, LC = L * C, cpg = C / p.groups;
  float* data = sm;                                   // [spc][LC]
  float2* stats = reinterpret_cast<float2*>(sm + (size_t)spc * LC);  // [spc][groups]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b_first = (p.rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * spc;
  const int nb = min(spc, p.B - b_first);
  const int c4n = C >> 2;
  // ---- load (coalesced float4), concat + skip scale applied here
  for (int i = tid; i < nb * L * c4n; i += nthr) {
    const int c = (i % c4n) * 4;
    const int row = i / c4n;                 // local row: s * L + l
    const size_t grow = (size_t)b_first * L + row;
    float4 v;
    if (c < p.c0) v = __ldg(reinterpret_cast<const float4*>(p.src0 + grow * p.c0 + c));
    else {
      v = __ldg(reinterpret_cast<const float4*>(p.src1 + grow * p.c1 + (c - p.c0)));
      v.x *= p.scale1; v.y *= p.scale1; v.z *= p.scale1; v.w *= p.scale1;
    }
    *reinterpret_cast<float4*>(data + (size_t)row * C + c) = v;
  }
  __syncthreads();
  // ---- two-pass statistics per (sample, group)
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  const int n = cpg * L;
  const int ngroups = nb * p.groups;
  if (ngroups >= nwarps) {
    // one warp per (sample, group)
    for (int sg = warp; sg < ngroups; sg += nwarps) {
      const int s = sg / p.groups, g = sg - s * p.groups;
      const float* base = data + (size_t)s * LC + g * cpg;
      float sum = 0.f;
      for (int i = lane; i < n; i += 32) { const int l = i / cpg; sum += base[l * C + (i - l * cpg)]; }
      const float mean = warp_sum_f(sum) / (float)n;
      float sq = 0.f;
      for (int i = lane; i < n; i += 32) { const int l = i / cpg; const float d = base[l * C + (i - l * cpg)] - mean; sq = fmaf(d, d, sq); }
      const float var = warp_sum_f(sq) / (float)n;
      if (lane == 0) stats[sg] = make_float2(mean, 1.0f / sqrtf(var + p.eps));
    }
    __syncthreads();
  } else {
    // few large groups (Patcher / Unpatcher GroupNorm(1)): split every group over wpg warps
    __shared__ float part[8];
    const int wpg = nwarps / ngroups;
    const int sg = warp / wpg, sub = warp - sg * wpg;
    const bool on = sg < ngroups;
    const int s = on ? sg / p.groups : 0, g = on ? sg - s * p.groups : 0;
    const float* base = data + (size_t)s * LC + g * cpg;
    float sum = 0.f;
    if (on) for (int i = sub * 32 + lane; i < n; i += wpg * 32) { const int l = i / cpg; sum += base[l * C + (i - l * cpg)]; }
    sum = warp_sum_f(sum);
    if (lane == 0) part[warp] = sum;
    __syncthreads();
    float mean = 0.f;
    if (on) { for (int w = 0; w < wpg; ++w) mean += part[sg * wpg + w]; mean /= (float)n; }
    __syncthreads();
    float sq = 0.f;
    if (on) for (int i = sub * 32 + lane; i < n; i += wpg * 32) { const int l = i / cpg; const float d = base[l * C + (i - l * cpg)] - mean; sq = fmaf(d, d, sq); }
    sq = warp_sum_f(sq);
    if (lane == 0) part[warp] = sq;
    __syncthreads();
    if (on && sub == 0 && lane == 0) {
      float var = 0.f;
      for (int w = 0; w < wpg; ++w) var += part[sg * wpg + w];
      stats[sg] = make_float2(mean, 1.0f / sqrtf(var / (float)n + p.eps));
    }
    __syncthreads();
  }
  // ---- apply
  const float* aff = nullptr;
  if (p.aff) aff = p.aff + (size_t)(p.call_idx ? *p.call_idx : 0) * p.aff_call_stride;
  for (int i = tid; i < nb * L * c4n; i += nthr) {
    const int c = (i % c4n) * 4;
    const int row = i / c4n;
    const int s = row / L;
    const size_t gidx = ((size_t)b_first * L + row) * C + c;
    float4 v = *reinterpret_cast<const float4*>(data + (size_t)row * C + c);
    if (p.raw) store_op4<KIND>(p.raw, gidx, v);
    const float2 st = stats[s * p.groups + c / cpg];
    v.x = (v.x - st.x) * st.y; v.y = (v.y - st.x) * st.y; v.z = (v.z - st.x) * st.y; v.w = (v.w - st.x) * st.y;
    if (aff) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(aff + c));
      const float4 h = __ldg(reinterpret_cast<const float4*>(aff + C + c));
      v.x = v.x * g.x + h.x; v.y = v.y * g.y + h.y; v.z = v.z * g.z + h.z; v.w = v.w * g.w + h.w;
    }
    if (p.silu) { v.x = silu_op(v.x); v.y = silu_op(v.y); v.z = silu_op(v.z); v.w = silu_op(v.w); }
    store_op4<KIND>(p.out, gidx, v);
  }
}

// Register-resident variant for groups of <= 1024 elements (every ResnetBlock / Transformer1d GroupNorm below level 0):
// a segment of SEG lanes (power of two) owns one (sample, group); each lane keeps its float4 pieces in registers across the
// two statistics passes and the apply pass, so the tensor is read exactly once and no shared memory or block barrier is used.
template <int KIND, int MAXP>
__global__ void __launch_bounds__(256) gn_apply_reg_kernel(const GnApplyParams p, const int seg, const int pieces) {
  pdl_enter();
  const int C = p.c0 + p.c1, L = p.L, cpg = C / p.groups, q4 = cpg >> 2;   // q4 = float4 per row segment
  const int lane = threadIdx.x & 31;
  const int gpw = 32 / seg;                                               // (sample, group) items per warp
  const long long warp_global = (long long)(p.rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long item = warp_global * gpw + lane / seg;
  const int sl = lane % seg;
  const long long items = (long long)p.B * p.groups;
  const bool on = item < items;
  const int b = on ? (int)(item / p.groups) : 0, g = on ? (int)(item % p.groups) : 0;
  float4 v[MAXP];
  int rowc[MAXP];   // packed (l << 16) | channel
  float sum = 0.f;
#pragma unroll
  for (int u = 0; u < MAXP; ++u) {
    const int i = sl + u * seg;
    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    rowc[u] = -1;
    if (on && i < pieces) {
      const int l = i / q4, c = g * cpg + (i - l * q4) * 4;
      rowc[u] = (l << 16) | c;
      const size_t row = (size_t)b * L + l;
      if (c < p.c0) v[u] = __ldg(reinterpret_cast<const float4*>(p.src0 + row * p.c0 + c));
      else {
        v[u] = __ldg(reinterpret_cast<const float4*>(p.src1 + row * p.c1 + (c - p.c0)));
        v[u].x *= p.scale1; v[u].y *= p.scale1; v[u].z *= p.scale1; v[u].w *= p.scale1;
      }
      sum += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    }
  }
  for (int o = seg >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float n = (float)(cpg * L);
  const float mean = sum / n;
  float sq = 0.f;
#pragma unroll
  for (int u = 0; u < MAXP; ++u) {
    if (rowc[u] >= 0) {
      const float a = v[u].x - mean, bb = v[u].y - mean, cc = v[u].z - mean, d = v[u].w - mean;
      sq = fmaf(a, a, sq); sq = fmaf(bb, bb, sq); sq = fmaf(cc, cc, sq); sq = fmaf(d, d, sq);
    }
  }
  for (int o = seg >> 1; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = 1.0f / sqrtf(sq / n + p.eps);
  const float* aff = nullptr;
  if (p.aff) aff = p.aff + (size_t)(p.call_idx ? *p.call_idx : 0) * p.aff_call_stride;
#pragma unroll
  for (int u = 0; u < MAXP; ++u) {
    if (rowc[u] >= 0) {
      const int l = rowc[u] >> 16, c = rowc[u] & 0xffff;
      const size_t gidx = ((size_t)b * L + l) * C + c;
      float4 x = v[u];
      if (p.raw) store_op4<KIND>(p.raw, gidx, x);
      x.x = (x.x - mean) * rstd; x.y = (x.y - mean) * rstd; x.z = (x.z - mean) * rstd; x.w = (x.w - mean) * rstd;
      if (aff) {
        const float4 ga = __ldg(reinterpret_cast<const float4*>(aff + c));
        const float4 ha = __ldg(reinterpret_cast<const float4*>(aff + C + c));
        x.x = x.x * ga.x + ha.x; x.y = x.y * ga.y + ha.y; x.z = x.z * ga.z + ha.z; x.w = x.w * ga.w + ha.w;
      }
      if (p.silu) { x.x = silu_op(x.x); x.y = silu_op(x.y); x.z = silu_op(x.z); x.w = silu_op(x.w); }
      store_op4<KIND>(p.out, gidx, x);
    }
  }
}

// Slab variant for short samples (L <= 16): a warp owns (sample, 128-channel slab); lane = 4 channels, the L rows live in registers.
// Every load / store instruction moves 512 contiguous bytes per row, all L rows of the item are in flight at once, and a warp walks
// items grid-stride with the next item's loads issued before the current one is reduced.  Groups must not straddle a slab
// (cpg | 128) and a slab must not straddle the two concatenated sources (c0 % 128 == 0).
template <int KIND, int LMAX>
__global__ void __launch_bounds__(256) gn_apply_slab_kernel(const GnApplyParams p) {
  pdl_enter();
  const int C = p.c0 + p.c1, L = p.L, cpg = C / p.groups;
  const int lane = threadIdx.x & 31;
  const int slabs = C >> 7;
  const long long items = (long long)p.B * slabs;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int gl = cpg >> 2;                       // lanes per group (1 .. 32)
  const float inv_n = 1.0f / (float)(cpg * L);
  const float* aff = nullptr;
  if (p.aff) aff = p.aff + (size_t)(p.call_idx ? *p.call_idx : 0) * p.aff_call_stride;
  auto load_item = [&](long long item, float4 (&v)[LMAX]) {
    const long long it2 = p.rev ? items - 1 - item : item;
    const int b = (int)(it2 / slabs), c = (int)(it2 - (long long)b * slabs) * 128 + lane * 4;
    const bool second = c >= p.c0;
    const float* src = second ? p.src1 + (size_t)b * L * p.c1 + (c - p.c0) : p.src0 + (size_t)b * L * p.c0 + c;
    const int ld = second ? p.c1 : p.c0;
#pragma unroll
    for (int l = 0; l < LMAX; ++l)
      if (l < L) {
        v[l] = __ldg(reinterpret_cast<const float4*>(src + (size_t)l * ld));
        if (second) { v[l].x *= p.scale1; v[l].y *= p.scale1; v[l].z *= p.scale1; v[l].w *= p.scale1; }
      }
  };
  constexpr bool PREFETCH = LMAX <= 8;           // 16 rows in registers twice over would cost the occupancy that hides the latency
  float4 cur[LMAX], nxt[PREFETCH ? LMAX : 1];
  if (PREFETCH && w < items) load_item(w, cur);
  for (; w < items; w += nwarps) {
    const bool more = PREFETCH && w + nwarps < items;
    if constexpr (PREFETCH) { if (more) load_item(w + nwarps, nxt); }
    else load_item(w, cur);
    const long long it2 = p.rev ? items - 1 - w : w;
    const int b = (int)(it2 / slabs), c = (int)(it2 - (long long)b * slabs) * 128 + lane * 4;
    float sum = 0.f;
#pragma unroll
    for (int l = 0; l < LMAX; ++l) if (l < L) sum += (cur[l].x + cur[l].y) + (cur[l].z + cur[l].w);
    for (int o = gl >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_n;
    float sq = 0.f;
#pragma unroll
    for (int l = 0; l < LMAX; ++l)
      if (l < L) {
        const float a = cur[l].x - mean, bb = cur[l].y - mean, cc = cur[l].z - mean, d = cur[l].w - mean;
        sq = fmaf(a, a, sq); sq = fmaf(bb, bb, sq); sq = fmaf(cc, cc, sq); sq = fmaf(d, d, sq);
      }
    for (int o = gl >> 1; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = 1.0f / sqrtf(sq * inv_n + p.eps);
    float4 ga = make_float4(1.f, 1.f, 1.f, 1.f), ha = make_float4(0.f, 0.f, 0.f, 0.f);
    if (aff) { ga = __ldg(reinterpret_cast<const float4*>(aff + c)); ha = __ldg(reinterpret_cast<const float4*>(aff + C + c)); }
    const size_t g0 = (size_t)b * L * C + c;
#pragma unroll
    for (int l = 0; l < LMAX; ++l)
      if (l < L) {
        float4 x = cur[l];
        if (p.raw) store_op4<KIND>(p.raw, g0 + (size_t)l * C, x);
        x.x = (x.x - mean) * rstd * ga.x + ha.x; x.y = (x.y - mean) * rstd * ga.y + ha.y;
        x.z = (x.z - mean) * rstd * ga.z + ha.z; x.w = (x.w - mean) * rstd * ga.w + ha.w;
        if (p.silu) { x.x = silu_op(x.x); x.y = silu_op(x.y); x.z = silu_op(x.z); x.w = silu_op(x.w); }
        store_op4<KIND>(p.out, g0 + (size_t)l * C, x);
      }
    if constexpr (PREFETCH) {
      if (more) {
#pragma unroll
        for (int l = 0; l < LMAX; ++l) cur[l] = nxt[l];
      }
    }
  }
}

static bool gn_slab_ok(const GnApplyParams& p) {
  const int C = p.c0 + p.c1;
  if (p.L < 1 || p.L > 16 || (C & 127) || C % p.groups) return false;
  const int cpg = C / p.groups;
  if (cpg < 4 || cpg > 128 || (cpg & (cpg - 1)) || (128 % cpg)) return false;
  if (p.c1 && (p.c0 & 127)) return false;
  // measured (profiles/README.md, B = 8192): the slab mapping wins for short samples and single-slab rows (L = 4: 20 vs 34 us,
  // L = 16 / C = 128: 45 vs 59 us) and loses where a lane would hold 16 rows of a two-slab sample (L = 16 / C = 256: 111 vs 91 us)
  return p.L <= 8 || C <= 128;
}

static int gn_spc(int L, int C) {
  int spc = 4096 / (L * C);
  if (spc < 1) spc = 1;
  if (spc > 8) spc = 8;
  return spc;
}

bool gn_apply_supported(int L, int C, int groups) {
  if (C % 4 || C % groups || (C / groups) % 4) return false;
  const size_t bytes = ((size_t)gn_spc(L, C) * L * C + 2 * 8 * 32) * sizeof(float);
  return bytes <= 200 * 1024;
}

cudaError_t launch_gn_apply(const GnApplyParams& p, int kind, cudaStream_t s) {
  if (p.B <= 0) return cudaSuccess;
  const int C = p.c0 + p.c1;
  static const bool slab_off = getenv("MDT_NO_GN_SLAB") != nullptr;
  if (!slab_off && gn_slab_ok(p)) {
    static int sms = 0;
    if (sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const long long items = (long long)p.B * (C >> 7);
    const long long blocks_needed = (items + 7) / 8;
    // two items per warp on average keeps the prefetch useful; cap at 8 resident blocks per SM
    long long blocks = p.L <= 8 ? (blocks_needed + 1) / 2 : blocks_needed;
    const long long cap = (long long)sms * 8;
    if (p.L <= 8 && blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cudaError_t e = cudaSuccess;
#define MDT_SLAB_LAUNCH(K, LM) e = launch_k_light(gn_apply_slab_kernel<K, LM>, (unsigned)blocks, 256, 0, s, p)
    if (kind == 1) { if (p.L <= 4) MDT_SLAB_LAUNCH(1, 4); else if (p.L <= 8) MDT_SLAB_LAUNCH(1, 8); else MDT_SLAB_LAUNCH(1, 16); }
    else if (kind == 2) { if (p.L <= 4) MDT_SLAB_LAUNCH(2, 4); else if (p.L <= 8) MDT_SLAB_LAUNCH(2, 8); else MDT_SLAB_LAUNCH(2, 16); }
    else { if (p.L <= 4) MDT_SLAB_LAUNCH(3, 4); else if (p.L <= 8) MDT_SLAB_LAUNCH(3, 8); else MDT_SLAB_LAUNCH(3, 16); }
#undef MDT_SLAB_LAUNCH
    return e;
  }
  {
    const int cpg = C / p.groups;
    const int pieces = p.L * cpg / 4;                    // float4 per (sample, group)
    if (cpg % 4 == 0 && pieces <= 256 && C < 65536 && p.L < 32768) {
      int seg = 1;
      while (seg < pieces && seg < 32) seg <<= 1;
      const int gpw = 32 / seg;
      const long long items = (long long)p.B * p.groups;
      const long long warps = (items + gpw - 1) / gpw;
      const unsigned grid = (unsigned)((warps + 7) / 8);
      const int ppl = (pieces + seg - 1) / seg;          // float4 pieces per lane: sizes the register arrays
      cudaError_t e = cudaSuccess;
#define MDT_GN_LAUNCH(K, P) e = launch_k_light(gn_apply_reg_kernel<K, P>, grid, 256, 0, s, p, seg, pieces)
      if (kind == 1) { if (ppl <= 1) MDT_GN_LAUNCH(1, 1); else if (ppl <= 2) MDT_GN_LAUNCH(1, 2); else if (ppl <= 4) MDT_GN_LAUNCH(1, 4); else MDT_GN_LAUNCH(1, 8); }
      else if (kind == 2) { if (ppl <= 1) MDT_GN_LAUNCH(2, 1); else if (ppl <= 2) MDT_GN_LAUNCH(2, 2); else if (ppl <= 4) MDT_GN_LAUNCH(2, 4); else MDT_GN_LAUNCH(2, 8); }
      else { if (ppl <= 1) MDT_GN_LAUNCH(3, 1); else if (ppl <= 2) MDT_GN_LAUNCH(3, 2); else if (ppl <= 4) MDT_GN_LAUNCH(3, 4); else MDT_GN_LAUNCH(3, 8); }
#undef MDT_GN_LAUNCH
      return e;
    }
  }
  const int spc = gn_spc(p.L, C);
  const size_t smem = ((size_t)spc * p.L * C + 2 * (size_t)spc * p.groups) * sizeof(float);
  const unsigned grid = (unsigned)((p.B + spc - 1) / spc);
  return launch_k_light(kind == 1 ? gn_apply_kernel<1> : (kind == 2 ? gn_apply_kernel<2> : gn_apply_kernel<3>), grid, 256, smem, s, p, spc);
}

template <int KIND, int NV>   // NV = float4 per lane = C / 128 rounded up (register array size)
__global__ void __launch_bounds__(256) ln_apply_kernel(const LnApplyParams p) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)(p.rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * 8 + warp;
  if (row >= p.rows) return;
  const int C = p.C;
  const float* src = p.src + (size_t)row * C;
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    if (c < C) { v[i] = __ldg(reinterpret_cast<const float4*>(src + c)); sum += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
  }
  const float mean = warp_sum_f(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    if (c < C) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      sq = fmaf(a, a, sq); sq = fmaf(b, b, sq); sq = fmaf(cc, cc, sq); sq = fmaf(d, d, sq);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum_f(sq) / (float)C + p.eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    if (c < C) {
      float4 o = make_float4((v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd);
      store_op4<KIND>(p.out, (size_t)row * C + c, o);
    }
  }
}

cudaError_t launch_ln_apply(const LnApplyParams& p, int kind, cudaStream_t s) {
  if (p.rows <= 0) return cudaSuccess;
  if (p.C % 4 || p.C > 1024) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)((p.rows + 7) / 8);
  const int nv = (p.C + 127) / 128;
  cudaError_t e = cudaSuccess;
#define MDT_LN_LAUNCH(K, N) e = launch_k_light(ln_apply_kernel<K, N>, grid, 256, 0, s, p)
  if (kind == 1) { if (nv <= 1) MDT_LN_LAUNCH(1, 1); else if (nv <= 2) MDT_LN_LAUNCH(1, 2); else if (nv <= 4) MDT_LN_LAUNCH(1, 4); else MDT_LN_LAUNCH(1, 8); }
  else if (kind == 2) { if (nv <= 1) MDT_LN_LAUNCH(2, 1); else if (nv <= 2) MDT_LN_LAUNCH(2, 2); else if (nv <= 4) MDT_LN_LAUNCH(2, 4); else MDT_LN_LAUNCH(2, 8); }
  else { if (nv <= 1) MDT_LN_LAUNCH(3, 1); else if (nv <= 2) MDT_LN_LAUNCH(3, 2); else if (nv <= 4) MDT_LN_LAUNCH(3, 4); else MDT_LN_LAUNCH(3, 8); }
#undef MDT_LN_LAUNCH
  return e;
}

cudaError_t init_prep() {
  cudaError_t e = cudaFuncSetAttribute(gn_apply_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_apply_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_apply_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  return e;
}

}  // namespace mdt
