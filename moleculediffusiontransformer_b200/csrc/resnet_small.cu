// resnet_small.cu -- the Patcher / Unpatcher ResnetBlock1d at level 0 (modules.py:208-257, 145-205) for few channels, one kernel.
//
// At level 0 the README models have 16 <-> 64 channels over 64 positions: 2.8 MFLOP per sample, far too little to feed a tcgen05
// tile (128 x N), while the op-by-op path spends ~0.9 ms per denoiser call there (GroupNorm passes, N = 16 GEMMs at a sliver of a
// tile, a CUDA-core GEMM that re-evaluates normalise + FiLM + SiLU for each of its three taps).  Here a persistent CTA keeps the
// packed weights in shared memory and walks samples; everything of one sample stays on chip:
//
//   x[L][Cin] -> GroupNorm(1) + affine + SiLU -> conv3 -> h1 -> GroupNorm(1) + FiLM affine + SiLU -> a2     skip = 1x1 conv of x
//   FULL (to_out):  out = conv3(a2) + b2 + skip                                  written as fp32 [L][Cout]
//   HEAD (to_in):   a2 written in the MMA operand dtype (the 64 -> 64 conv3 that follows is tcgen05 work), skip written as fp32
//
// The convolutions are warp-level implicit GEMMs on mma.sync.m16n8k8 (tf32): M = positions (L / 16 tiles), N = 16 output channels
// per warp, K = tap * Cin + c.  Activations sit in shared memory channel-major with a zero halo (the conv's padding; row stride
// L + 8 makes the A-fragment loads conflict-free), weights k-major with a padded row (conflict-free B fragments).  Layers whose
// rounding would land directly on the network output or input (to_out's last conv, to_in's first) run as 3xTF32 (hi/lo split of
// both operands, three MMAs): fp32-grade results at a cost that is noise here.  HBM traffic per sample: input once, outputs once.
#include <cuda_bf16.h>
#include "aload.cuh"
#include "attn_math.cuh"
#include "tc_common.cuh"

namespace mdt {

constexpr int RS_THREADS = 128;
constexpr int RS_HALO = 4;           // floats of zero padding on either side of a channel row (row stride L + 8)
constexpr int RS_WPAD = 8;           // weight row padding in floats (row stride N + 8: conflict-free B fragments)

// SiLU with ex2.approx / rcp.approx (~2 ulp each): an order of magnitude below tf32 operand rounding, and still fp32-grade for the
// 3xTF32 layers; the precise expf + division version is most of this kernel's instruction count otherwise
__device__ __forceinline__ float silu_fast(float v) { return v * __fdividef(1.0f, 1.0f + __expf(-v)); }

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = tc::to_tf32(v);
  lo = tc::to_tf32(v - __uint_as_float(hi));
}

// acc[mt][nt][4] += A (positions x K) * W (K x 16 channels of this warp) for a `taps`-tap conv over a channel-major activation tile.
//   act: element (c, pos) at act[c * LP + RS_HALO + pos]; w: element (k, co) at w[k * WLD + co], k = tap * C + c
template <int MT, bool SPLIT>
__device__ __forceinline__ void conv_mma(float (&acc)[MT][2][4], const float* __restrict__ act, int LP, int C, const float* __restrict__ w,
                                         int WLD, int taps, int co_base, int lane) {
  const int g = lane >> 2, q = lane & 3;
  const int pad = taps >> 1;
  for (int tap = 0; tap < taps; ++tap) {
    for (int c0 = 0; c0 < C; c0 += 8) {
      const float* w0 = w + (size_t)(tap * C + c0 + q) * WLD + co_base + g;
      const float* a0p = act + (size_t)(c0 + q) * LP + RS_HALO + g + tap - pad;
      float bw[2][2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { bw[nt][0] = w0[8 * nt]; bw[nt][1] = w0[4 * WLD + 8 * nt]; }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const float av[4] = {a0p[16 * mt], a0p[16 * mt + 8], a0p[4 * LP + 16 * mt], a0p[4 * LP + 16 * mt + 8]};
        if (SPLIT) {
          uint32_t ah[4], al[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) split_tf32(av[i], ah[i], al[i]);
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(bw[nt][0], bh0, bl0); split_tf32(bw[nt][1], bh1, bl1);
            mma_tf32_16x8x8(acc[mt][nt], al, bh0, bh1);
            mma_tf32_16x8x8(acc[mt][nt], ah, bl0, bl1);
            mma_tf32_16x8x8(acc[mt][nt], ah, bh0, bh1);
          }
        } else {
          const uint32_t af[4] = {__float_as_uint(av[0]), __float_as_uint(av[1]), __float_as_uint(av[2]), __float_as_uint(av[3])};
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) mma_tf32_16x8x8(acc[mt][nt], af, __float_as_uint(bw[nt][0]), __float_as_uint(bw[nt][1]));
        }
      }
    }
  }
}

// WPS = warps per sample (each warp: all positions x 16 output channels): 1 for Cout = 16 (four samples per CTA), 4 for Cout = 64.
// MT = L / 16 position tiles.
template <int WPS, int MT>
__global__ void __launch_bounds__(RS_THREADS) resnet_small_kernel(const ResnetSmallParams p) {
  extern __shared__ __align__(16) float rs_smem[];
  __shared__ float red[4];
  pdl_enter();
  constexpr int TPS = 32 * WPS, SPC = 4 / WPS;
  constexpr int L = 16 * MT, LP = L + 2 * RS_HALO;
  const int Cin = p.Cin, Cout = p.Cout;
  const int WLD = Cout + RS_WPAD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sl = warp / WPS, ts = tid - sl * TPS;        // sample slot, thread index inside the sample's group
  const int co_base = 16 * (warp % WPS);
  const int g = lane >> 2, q = lane & 3;
  const int Ct = Cin > Cout ? Cin : Cout;                 // ONE tile per sample: raw input, then activated in place, then a2
  const size_t tile_f = (size_t)Ct * LP + 8;
  float* w1 = rs_smem;                                    // [3 * Cin][WLD]
  float* ws = w1 + (size_t)3 * Cin * WLD;                 // [Cin][WLD]        (only with a projection)
  float* w2 = ws + (p.ws ? (size_t)Cin * WLD : 0);        // [3 * Cout][WLD]   (mode 0 only)
  float* tiles = w2 + (p.mode == 0 ? (size_t)3 * Cout * WLD : 0);
  float* xs = tiles + (size_t)sl * tile_f;                // [Ct][LP] raw input (statistics, skip path) ...
  float* as = xs;                                         // ... activated in place once the skip projection has read it; later a2

  // ---- once per CTA: weights k-major (pre-rounded to tf32 where the layer runs single-pass) and zeroed tiles (halos = padding)
  for (int i = tid; i < 3 * Cin * Cout; i += RS_THREADS) {
    const int co = i / (3 * Cin), k = i - co * 3 * Cin; const float v = p.w1[i];
    w1[k * WLD + co] = p.split1 ? v : __uint_as_float(tc::to_tf32(v));
  }
  if (p.ws) for (int i = tid; i < Cin * Cout; i += RS_THREADS) {
    const int co = i / Cin, k = i - co * Cin; const float v = p.ws[i];
    ws[k * WLD + co] = p.split1 ? v : __uint_as_float(tc::to_tf32(v));
  }
  if (p.mode == 0) for (int i = tid; i < 3 * Cout * Cout; i += RS_THREADS) {
    const int co = i / (3 * Cout), k = i - co * 3 * Cout; const float v = p.w2[i];
    w2[k * WLD + co] = p.split2 ? v : __uint_as_float(tc::to_tf32(v));
  }
  for (size_t i = tid; i < (size_t)SPC * tile_f; i += RS_THREADS) tiles[i] = 0.f;
  __syncthreads();

  auto group_sync = [&]() { if (WPS == 1) __syncwarp(); else __syncthreads(); };
  auto group_sum = [&](float v) -> float {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (WPS == 1) return v;
    __syncthreads();                             // red[] of the previous reduction has been consumed
    if (lane == 0) red[warp] = v;
    __syncthreads();
    return (red[0] + red[1]) + (red[2] + red[3]);
  };
  const float* aff2 = p.aff2 + (size_t)(p.call_idx ? *p.call_idx : 0) * p.aff2_stride;
  const int nin = L * Cin, nh = L * Cout;
  const int passes = (p.B + (int)gridDim.x * SPC - 1) / ((int)gridDim.x * SPC);

  for (int it = 0; it < passes; ++it) {
    const int b = (it * (int)gridDim.x + (int)blockIdx.x) * SPC + sl;
    const bool on = b < p.B;                     // uniform per sample group
    if (WPS == 4 && !on) break;
    if (on) {
      // ---- 1. load x (coalesced float4, eight in flight) into the channel-major tile; GroupNorm(1) statistics; activation
      const float4* xg4 = reinterpret_cast<const float4*>(p.x + (size_t)b * nin);
      const int n4 = nin >> 2, c4n = Cin >> 2;
      float s = 0.f;
      for (int e0 = ts; e0 < n4; e0 += 8 * TPS) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int e = e0 + u * TPS; v[u] = e < n4 ? __ldg(xg4 + e) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = e0 + u * TPS;
          if (e < n4) {
            const int ll = e / c4n, c = (e - ll * c4n) * 4;
            float* dp = xs + c * LP + RS_HALO + ll;
            dp[0] = v[u].x; dp[LP] = v[u].y; dp[2 * LP] = v[u].z; dp[3 * LP] = v[u].w;
            s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
          }
        }
      }
      const float mean1 = group_sum(s) / (float)nin;
      float sq = 0.f;
#pragma unroll 8
      for (int e = ts; e < nin; e += TPS) { const int c = e / L, ll = e - c * L; const float dv = xs[c * LP + RS_HALO + ll] - mean1; sq = fmaf(dv, dv, sq); }
      const float rstd1 = 1.0f / sqrtf(group_sum(sq) / (float)nin + p.eps);
      // ---- 2a. the 1x1 skip projection reads the raw tile first: positions x 16 channels per warp, accumulators in mma C layout
      //          (thread (g, q): rows 16 mt + g and + 8, channels co_base + 8 nt + 2q and + 1)
      float h[MT][2][4], sk[MT][2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int co = co_base + 8 * nt + 2 * q;
        const float2 b1v = __ldg(reinterpret_cast<const float2*>(p.b1 + co));
        const float2 bsv = p.ws ? __ldg(reinterpret_cast<const float2*>(p.bs + co)) : make_float2(0.f, 0.f);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          h[mt][nt][0] = b1v.x; h[mt][nt][1] = b1v.y; h[mt][nt][2] = b1v.x; h[mt][nt][3] = b1v.y;
          sk[mt][nt][0] = bsv.x; sk[mt][nt][1] = bsv.y; sk[mt][nt][2] = bsv.x; sk[mt][nt][3] = bsv.y;
        }
      }
      if (p.ws) {
        if (p.split1) conv_mma<MT, true>(sk, xs, LP, Cin, ws, WLD, 1, co_base, lane);
        else conv_mma<MT, false>(sk, xs, LP, Cin, ws, WLD, 1, co_base, lane);
      } else {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const float* xr = xs + (size_t)(co_base + 8 * nt + 2 * q) * LP + RS_HALO + 16 * mt + g;      // identity skip (Cin == Cout)
            sk[mt][nt][0] = xr[0]; sk[mt][nt][1] = xr[LP]; sk[mt][nt][2] = xr[8]; sk[mt][nt][3] = xr[LP + 8];
          }
      }
      group_sync();                              // every warp of the sample has read the raw tile
      // ---- 2b. normalise + affine + SiLU in place, then conv1 (3 taps)
#pragma unroll 8
      for (int e = ts; e < nin; e += TPS) {
        const int c = e / L, ll = e - c * L;
        const float v = silu_fast((xs[c * LP + RS_HALO + ll] - mean1) * rstd1 * __ldg(p.aff1 + c) + __ldg(p.aff1 + Cin + c));
        as[c * LP + RS_HALO + ll] = p.split1 ? v : __uint_as_float(tc::to_tf32(v));
      }
      group_sync();
      if (p.split1) conv_mma<MT, true>(h, as, LP, Cin, w1, WLD, 3, co_base, lane);
      else conv_mma<MT, false>(h, as, LP, Cin, w1, WLD, 3, co_base, lane);
      // ---- 3. GroupNorm(1) of h1 (two passes over registers), FiLM affine, SiLU
      float s2 = 0.f;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) s2 += (h[mt][nt][0] + h[mt][nt][1]) + (h[mt][nt][2] + h[mt][nt][3]);
      const float mean2 = group_sum(s2) / (float)nh;
      float q2 = 0.f;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) { const float dv = h[mt][nt][i] - mean2; q2 = fmaf(dv, dv, q2); }
      const float rstd2 = 1.0f / sqrtf(group_sum(q2) / (float)nh + p.eps);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int co = co_base + 8 * nt + 2 * q;
        const float2 ga = __ldg(reinterpret_cast<const float2*>(aff2 + co)), gb = __ldg(reinterpret_cast<const float2*>(aff2 + Cout + co));
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          h[mt][nt][0] = silu_fast(fmaf((h[mt][nt][0] - mean2) * rstd2, ga.x, gb.x));
          h[mt][nt][1] = silu_fast(fmaf((h[mt][nt][1] - mean2) * rstd2, ga.y, gb.y));
          h[mt][nt][2] = silu_fast(fmaf((h[mt][nt][2] - mean2) * rstd2, ga.x, gb.x));
          h[mt][nt][3] = silu_fast(fmaf((h[mt][nt][3] - mean2) * rstd2, ga.y, gb.y));
        }
      }
      if (p.mode == 1) {
        // HEAD: a2 in the operand dtype for the tensor-core conv that follows, skip as the fp32 residual it adds
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const size_t o = ((size_t)b * L + 16 * mt + g + 8 * rr) * Cout + co_base + 8 * nt + 2 * q;
              if (p.kind >= 2) *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.a2op) + o) =
                  p.kind == 3 ? tc::pack_f16s(h[mt][nt][2 * rr], h[mt][nt][2 * rr + 1]) : tc::pack_bf16(h[mt][nt][2 * rr], h[mt][nt][2 * rr + 1]);
              else *reinterpret_cast<uint2*>(reinterpret_cast<float*>(p.a2op) + o) = make_uint2(tc::to_tf32(h[mt][nt][2 * rr]), tc::to_tf32(h[mt][nt][2 * rr + 1]));
              *reinterpret_cast<float2*>(p.out + o) = make_float2(sk[mt][nt][2 * rr], sk[mt][nt][2 * rr + 1]);
            }
      } else {
        // ---- 4. FULL: conv2 over a2 (through the channel-major tile: conv1's input is dead), + bias + skip
        group_sync();                            // every warp of the sample is done reading the tile as conv1's input
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            float* hr = as + (size_t)(co_base + 8 * nt + 2 * q) * LP + RS_HALO + 16 * mt + g;
            const float v0 = h[mt][nt][0], v1 = h[mt][nt][1], v2 = h[mt][nt][2], v3 = h[mt][nt][3];
            hr[0] = p.split2 ? v0 : __uint_as_float(tc::to_tf32(v0)); hr[LP] = p.split2 ? v1 : __uint_as_float(tc::to_tf32(v1));
            hr[8] = p.split2 ? v2 : __uint_as_float(tc::to_tf32(v2)); hr[LP + 8] = p.split2 ? v3 : __uint_as_float(tc::to_tf32(v3));
          }
        group_sync();
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const float2 b2v = __ldg(reinterpret_cast<const float2*>(p.b2 + co_base + 8 * nt + 2 * q));
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) { sk[mt][nt][0] += b2v.x; sk[mt][nt][1] += b2v.y; sk[mt][nt][2] += b2v.x; sk[mt][nt][3] += b2v.y; }
        }
        if (p.split2) conv_mma<MT, true>(sk, as, LP, Cout, w2, WLD, 3, co_base, lane);
        else conv_mma<MT, false>(sk, as, LP, Cout, w2, WLD, 3, co_base, lane);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const size_t o = ((size_t)b * L + 16 * mt + g + 8 * rr) * Cout + co_base + 8 * nt + 2 * q;
              *reinterpret_cast<float2*>(p.out + o) = make_float2(sk[mt][nt][2 * rr], sk[mt][nt][2 * rr + 1]);
            }
      }
    }
    group_sync();          // the tiles are rewritten by the next sample
  }
}

static int resnet_small_wps(int L, int Cout) {
  if (!(L == 16 || L == 32 || L == 64)) return 0;
  return Cout == 16 ? 1 : (Cout == 64 ? 4 : 0);
}

static size_t resnet_small_smem(int L, int Cin, int Cout, bool proj, int mode) {
  const int wps = resnet_small_wps(L, Cout);
  if (!wps) return ~(size_t)0;
  const size_t LP = L + 2 * RS_HALO, WLD = Cout + RS_WPAD;
  const size_t Ct = Cin > Cout ? Cin : Cout;
  size_t f = (size_t)(4 / wps) * (Ct * LP + 8) + (size_t)3 * Cin * WLD;
  if (proj) f += (size_t)Cin * WLD;
  if (mode == 0) f += (size_t)3 * Cout * WLD;
  return f * sizeof(float);
}

static const size_t RS_SMEM_MAX = 200 * 1024;

bool resnet_small_supported(int L, int Cin, int Cout, int groups, bool proj, int mode) {
  if (groups != 1 || Cin < 8 || (Cin & 7) || Cin > 128 || (!proj && Cin != Cout)) return false;
  if (resnet_small_wps(L, Cout) == 0) return false;
  return resnet_small_smem(L, Cin, Cout, proj, mode) <= RS_SMEM_MAX;
}

typedef void (*ResnetSmallKernel)(const ResnetSmallParams);
static ResnetSmallKernel resnet_small_variant(int wps, int L) {
  static const ResnetSmallKernel tab[2][3] = {{resnet_small_kernel<1, 1>, resnet_small_kernel<1, 2>, resnet_small_kernel<1, 4>},
                                              {resnet_small_kernel<4, 1>, resnet_small_kernel<4, 2>, resnet_small_kernel<4, 4>}};
  return tab[wps == 1 ? 0 : 1][L == 16 ? 0 : (L == 32 ? 1 : 2)];
}

cudaError_t init_resnet_small() {
  for (int wps : {1, 4})
    for (int L : {16, 32, 64}) {
      cudaError_t e = cudaFuncSetAttribute(resnet_small_variant(wps, L), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM_MAX);
      if (e != cudaSuccess) return e;
    }
  return cudaSuccess;
}

static int g_sms_rs = 0;

cudaError_t launch_resnet_small(const ResnetSmallParams& p, cudaStream_t s) {
  if (p.B <= 0) return cudaSuccess;
  const int wps = resnet_small_wps(p.L, p.Cout);
  if (wps == 0) return cudaErrorInvalidValue;
  const size_t smem = resnet_small_smem(p.L, p.Cin, p.Cout, p.ws != nullptr, p.mode);
  if (smem > RS_SMEM_MAX) return cudaErrorInvalidValue;
  if (g_sms_rs == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_rs, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms_rs <= 0) g_sms_rs = 148;
  }
  int per_sm = (int)((224 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const int spc = 4 / wps;
  const long long need = ((long long)p.B + spc - 1) / spc, want = (long long)g_sms_rs * per_sm;
  const unsigned grid = (unsigned)(need < want ? need : want);
  return launch_k(resnet_small_variant(wps, p.L), grid, RS_THREADS, smem, s, p);
}

}  // namespace mdt
