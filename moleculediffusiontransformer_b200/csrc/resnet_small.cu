// resnet_small.cu -- the Patcher / Unpatcher ResnetBlock1d at level 0 (modules.py:208-257, 145-205) for few channels, one kernel.
//
// At level 0 the README models have 16 <-> 64 channels over 64 positions: 2.8 MFLOP per sample, far too little to feed a tensor
// core tile, while the op-by-op path spends ~0.9 ms per denoiser call there (GroupNorm passes, N = 16 GEMMs at a fraction of a tile,
// a CUDA-core GEMM that re-evaluates normalise + FiLM + SiLU for each of its three taps).  Here a persistent CTA keeps the packed
// weights in shared memory and walks samples; everything of one sample stays on chip:
//
//   x[L][Cin] -> GroupNorm(1) + affine + SiLU -> conv3 -> h1 -> GroupNorm(1) + FiLM affine + SiLU -> a2     skip = 1x1 conv of x
//   FULL (to_out):  out = conv3(a2) + b2 + skip                                  written as fp32 [L][Cout]
//   HEAD (to_in):   a2 written in the MMA operand dtype (the 64 -> 64 conv3 that follows is tensor-core work), skip written as fp32
//
// A thread owns 4 consecutive positions x 8 output channels; activations sit in shared memory channel-major with a zero halo (the
// conv's padding), weights k-major so that a thread's output channels are two vector loads shared by the warp (broadcast).
// HBM traffic per sample is the input once and the outputs once.
#include <cuda_bf16.h>
#include "aload.cuh"
#include "tc_common.cuh"

namespace mdt {

constexpr int RS_THREADS = 128;
constexpr int RS_HALO = 4;           // floats of zero padding on either side of a channel row (keeps position 0 float4-aligned)

// Thread tile: 4 consecutive positions x 8 output channels (32 accumulators; per input channel 3 activation loads and 6 weight
// vector loads feed 96 FMAs).  TPS = threads per sample = (L / 4) * (Cout / 8): 32 (one warp per sample, four samples per CTA, warp
// shuffles and __syncwarp only) or 128 (one sample per CTA).
template <int TPS>
__global__ void __launch_bounds__(RS_THREADS) resnet_small_kernel(const ResnetSmallParams p) {
  extern __shared__ __align__(16) float rs_smem[];
  __shared__ float red[4];
  constexpr int SPC = RS_THREADS / TPS;          // samples per CTA pass
  const int L = p.L, Cin = p.Cin, Cout = p.Cout, LP = L + 2 * RS_HALO;
  const int tid = threadIdx.x;
  const int sl = tid / TPS, ts = tid % TPS;      // sample slot of this thread, thread index inside the sample
  const int l0 = 4 * (ts % (L >> 2)), co0 = 8 * (ts / (L >> 2));
  const size_t tile_f = (size_t)(2 * Cin + Cout) * LP;
  float* w1 = rs_smem;                           // [3 * Cin][Cout]
  float* ws = w1 + (size_t)3 * Cin * Cout;       // [Cin][Cout]        (only with a projection)
  float* w2 = ws + (p.ws ? (size_t)Cin * Cout : 0);   // [3 * Cout][Cout]   (mode 0 only)
  float* tiles = w2 + (p.mode == 0 ? (size_t)3 * Cout * Cout : 0);
  float* xs = tiles + (size_t)sl * tile_f;       // [Cin][LP]  raw input (skip path, statistics)
  float* as = xs + (size_t)Cin * LP;             // [Cin][LP]  activated input of conv1
  float* hs = as + (size_t)Cin * LP;             // [Cout][LP] a2 (input of conv2)

  // ---- once per CTA: weights (transposed to k-major) and zeroed tiles (the halos stay zero: they are the conv padding)
  for (int i = tid; i < 3 * Cin * Cout; i += RS_THREADS) { const int co = i / (3 * Cin), k = i - co * 3 * Cin; w1[k * Cout + co] = p.w1[i]; }
  if (p.ws) for (int i = tid; i < Cin * Cout; i += RS_THREADS) { const int co = i / Cin, k = i - co * Cin; ws[k * Cout + co] = p.ws[i]; }
  if (p.mode == 0) for (int i = tid; i < 3 * Cout * Cout; i += RS_THREADS) { const int co = i / (3 * Cout), k = i - co * 3 * Cout; w2[k * Cout + co] = p.w2[i]; }
  for (size_t i = tid; i < (size_t)SPC * tile_f; i += RS_THREADS) tiles[i] = 0.f;
  __syncthreads();

  auto group_sync = [&]() { if (TPS == 32) __syncwarp(); else __syncthreads(); };
  auto group_sum = [&](float v) -> float {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (TPS == 32) return v;
    __syncthreads();                             // red[] of the previous reduction has been consumed
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    return (red[0] + red[1]) + (red[2] + red[3]);
  };
  const float* aff2 = p.aff2 + (size_t)(p.call_idx ? *p.call_idx : 0) * p.aff2_stride;
  const int nin = L * Cin, nh = L * Cout;
  const int passes = (p.B + (int)gridDim.x * SPC - 1) / ((int)gridDim.x * SPC);

  for (int it = 0; it < passes; ++it) {
    const int b = (it * (int)gridDim.x + (int)blockIdx.x) * SPC + sl;
    const bool on = b < p.B;                     // (TPS == 128: uniform per CTA; TPS == 32: uniform per warp)
    if (TPS == 128 && !on) break;
    if (on) {
      // ---- 1. load x (coalesced, token-major) into the channel-major tile; statistics over the whole sample (GroupNorm(1))
      const float* xg = p.x + (size_t)b * nin;
      float s = 0.f;
      if ((Cin & 3) == 0) {
        // float4 loads, eight in flight per thread: the HBM latency of the whole sample is paid about once
        const float4* xg4 = reinterpret_cast<const float4*>(xg);
        const int n4 = nin >> 2, c4n = Cin >> 2;
        for (int e0 = ts; e0 < n4; e0 += 8 * TPS) {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { const int e = e0 + u * TPS; v[u] = e < n4 ? __ldg(xg4 + e) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * TPS;
            if (e < n4) {
              const int ll = e / c4n, c = (e - ll * c4n) * 4;
              float* d = xs + c * LP + RS_HALO + ll;
              d[0] = v[u].x; d[LP] = v[u].y; d[2 * LP] = v[u].z; d[3 * LP] = v[u].w;
              s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
            }
          }
        }
      } else {
        for (int e = ts; e < nin; e += TPS) { const float v = xg[e]; const int ll = e / Cin, c = e - ll * Cin; xs[c * LP + RS_HALO + ll] = v; s += v; }
      }
      const float mean1 = group_sum(s) / (float)nin;
      float sq = 0.f;
      for (int e = ts; e < nin; e += TPS) { const int c = e / L, ll = e - c * L; const float dv = xs[c * LP + RS_HALO + ll] - mean1; sq = fmaf(dv, dv, sq); }
      const float rstd1 = 1.0f / sqrtf(group_sum(sq) / (float)nin + p.eps);
      for (int e = ts; e < nin; e += TPS) {
        const int c = e / L, ll = e - c * L;
        const float v = (xs[c * LP + RS_HALO + ll] - mean1) * rstd1 * __ldg(p.aff1 + c) + __ldg(p.aff1 + Cin + c);
        as[c * LP + RS_HALO + ll] = silu_f(v);
      }
      group_sync();
      // ---- 2. conv1 (3 taps) and the 1x1 skip projection for positions l0 .. l0 + 3, channels co0 .. co0 + 7
      float h[4][8], sk[4][8];
      {
        const float4 ba = __ldg(reinterpret_cast<const float4*>(p.b1 + co0)), bb = __ldg(reinterpret_cast<const float4*>(p.b1 + co0 + 4));
#pragma unroll
        for (int i = 0; i < 4; ++i) { h[i][0] = ba.x; h[i][1] = ba.y; h[i][2] = ba.z; h[i][3] = ba.w; h[i][4] = bb.x; h[i][5] = bb.y; h[i][6] = bb.z; h[i][7] = bb.w; }
      }
      for (int c = 0; c < Cin; ++c) {
        const float* ar = as + c * LP + RS_HALO + l0;
        const float4 am = *reinterpret_cast<const float4*>(ar);
        const float av[6] = {ar[-1], am.x, am.y, am.z, am.w, ar[4]};
#pragma unroll
        for (int tap = 0; tap < 3; ++tap) {
          const float* wr = w1 + (size_t)(tap * Cin + c) * Cout + co0;
          const float4 wa = *reinterpret_cast<const float4*>(wr), wb = *reinterpret_cast<const float4*>(wr + 4);
          const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) h[i][j] = fmaf(av[i + tap], w[j], h[i][j]);
        }
      }
      if (p.ws) {
        const float4 ba = __ldg(reinterpret_cast<const float4*>(p.bs + co0)), bb = __ldg(reinterpret_cast<const float4*>(p.bs + co0 + 4));
#pragma unroll
        for (int i = 0; i < 4; ++i) { sk[i][0] = ba.x; sk[i][1] = ba.y; sk[i][2] = ba.z; sk[i][3] = ba.w; sk[i][4] = bb.x; sk[i][5] = bb.y; sk[i][6] = bb.z; sk[i][7] = bb.w; }
        for (int c = 0; c < Cin; ++c) {
          const float4 xm = *reinterpret_cast<const float4*>(xs + c * LP + RS_HALO + l0);
          const float xv[4] = {xm.x, xm.y, xm.z, xm.w};
          const float* wr = ws + (size_t)c * Cout + co0;
          const float4 wa = *reinterpret_cast<const float4*>(wr), wb = *reinterpret_cast<const float4*>(wr + 4);
          const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) sk[i][j] = fmaf(xv[i], w[j], sk[i][j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 xm = *reinterpret_cast<const float4*>(xs + (co0 + j) * LP + RS_HALO + l0);     // identity skip (Cin == Cout)
          sk[0][j] = xm.x; sk[1][j] = xm.y; sk[2][j] = xm.z; sk[3][j] = xm.w;
        }
      }
      // ---- 3. GroupNorm(1) of h1 (two passes over registers), FiLM affine, SiLU
      float s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s2 += h[i][j];
      const float mean2 = group_sum(s2) / (float)nh;
      float q2 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float dv = h[i][j] - mean2; q2 = fmaf(dv, dv, q2); }
      const float rstd2 = 1.0f / sqrtf(group_sum(q2) / (float)nh + p.eps);
      {
        float ga[8], gb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { ga[j] = __ldg(aff2 + co0 + j); gb[j] = __ldg(aff2 + Cout + co0 + j); }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) h[i][j] = silu_f(fmaf((h[i][j] - mean2) * rstd2, ga[j], gb[j]));
      }
      if (p.mode == 1) {
        // HEAD: a2 in the operand dtype for the tensor-core conv that follows, skip as the fp32 residual it adds
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const size_t orow = ((size_t)b * L + l0 + i) * Cout + co0;
          if (p.kind == 2) {
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.a2op) + orow) =
                make_uint4(tc::pack_bf16(h[i][0], h[i][1]), tc::pack_bf16(h[i][2], h[i][3]), tc::pack_bf16(h[i][4], h[i][5]), tc::pack_bf16(h[i][6], h[i][7]));
          } else {
            *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.a2op) + orow) = make_uint4(tc::to_tf32(h[i][0]), tc::to_tf32(h[i][1]), tc::to_tf32(h[i][2]), tc::to_tf32(h[i][3]));
            *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.a2op) + orow + 4) = make_uint4(tc::to_tf32(h[i][4]), tc::to_tf32(h[i][5]), tc::to_tf32(h[i][6]), tc::to_tf32(h[i][7]));
          }
          *reinterpret_cast<float4*>(p.out + orow) = make_float4(sk[i][0], sk[i][1], sk[i][2], sk[i][3]);
          *reinterpret_cast<float4*>(p.out + orow + 4) = make_float4(sk[i][4], sk[i][5], sk[i][6], sk[i][7]);
        }
      } else {
        // ---- 4. FULL: conv2 over a2 (through the channel-major tile for the neighbours), + bias + skip
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(hs + (co0 + j) * LP + RS_HALO + l0) = make_float4(h[0][j], h[1][j], h[2][j], h[3][j]);
        group_sync();
        {
          const float4 ba = __ldg(reinterpret_cast<const float4*>(p.b2 + co0)), bb = __ldg(reinterpret_cast<const float4*>(p.b2 + co0 + 4));
          const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) sk[i][j] += bv[j];
        }
        for (int c = 0; c < Cout; ++c) {
          const float* ar = hs + c * LP + RS_HALO + l0;
          const float4 am = *reinterpret_cast<const float4*>(ar);
          const float av[6] = {ar[-1], am.x, am.y, am.z, am.w, ar[4]};
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) {
            const float* wr = w2 + (size_t)(tap * Cout + c) * Cout + co0;
            const float4 wa = *reinterpret_cast<const float4*>(wr), wb = *reinterpret_cast<const float4*>(wr + 4);
            const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 8; ++j) sk[i][j] = fmaf(av[i + tap], w[j], sk[i][j]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const size_t orow = ((size_t)b * L + l0 + i) * Cout + co0;
          *reinterpret_cast<float4*>(p.out + orow) = make_float4(sk[i][0], sk[i][1], sk[i][2], sk[i][3]);
          *reinterpret_cast<float4*>(p.out + orow + 4) = make_float4(sk[i][4], sk[i][5], sk[i][6], sk[i][7]);
        }
      }
    }
    group_sync();          // the tiles are rewritten by the next sample
  }
}

static int resnet_small_tps(int L, int Cout) {
  if (L < 4 || (L & 3) || Cout < 8 || (Cout & 7)) return 0;
  const int tps = (L >> 2) * (Cout >> 3);
  return (tps == 32 || tps == 128) ? tps : 0;
}

static size_t resnet_small_smem(int L, int Cin, int Cout, bool proj, int mode) {
  const int tps = resnet_small_tps(L, Cout);
  if (!tps) return ~(size_t)0;
  const size_t LP = L + 2 * RS_HALO;
  size_t f = (size_t)(RS_THREADS / tps) * (2 * (size_t)Cin + Cout) * LP + (size_t)3 * Cin * Cout;
  if (proj) f += (size_t)Cin * Cout;
  if (mode == 0) f += (size_t)3 * Cout * Cout;
  return f * sizeof(float);
}

static const size_t RS_SMEM_MAX = 200 * 1024;

bool resnet_small_supported(int L, int Cin, int Cout, int groups, bool proj, int mode) {
  if (groups != 1 || Cin < 1 || Cin > 128 || (!proj && Cin != Cout)) return false;
  if (resnet_small_tps(L, Cout) == 0) return false;
  return resnet_small_smem(L, Cin, Cout, proj, mode) <= RS_SMEM_MAX;
}

cudaError_t init_resnet_small() {
  cudaError_t e = cudaFuncSetAttribute(resnet_small_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM_MAX);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(resnet_small_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM_MAX);
  return e;
}

static int g_sms_rs = 0;

cudaError_t launch_resnet_small(const ResnetSmallParams& p, cudaStream_t s) {
  if (p.B <= 0) return cudaSuccess;
  const int tps = resnet_small_tps(p.L, p.Cout);
  if (tps == 0) return cudaErrorInvalidValue;
  const size_t smem = resnet_small_smem(p.L, p.Cin, p.Cout, p.ws != nullptr, p.mode);
  if (smem > RS_SMEM_MAX) return cudaErrorInvalidValue;
  if (g_sms_rs == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_rs, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms_rs <= 0) g_sms_rs = 148;
  }
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const int spc = RS_THREADS / tps;
  const long long need = ((long long)p.B + spc - 1) / spc, want = (long long)g_sms_rs * per_sm;
  const unsigned grid = (unsigned)(need < want ? need : want);
  if (tps == 32) resnet_small_kernel<32><<<grid, RS_THREADS, smem, s>>>(p);
  else resnet_small_kernel<128><<<grid, RS_THREADS, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace mdt
