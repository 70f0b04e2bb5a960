// attention_bulk.cu -- softmax attention for short sequences, fed by TMA bulk copies (cp.async.bulk).
//
// AttentionBase.forward (modules.py:350-364) for nq, nk <= 64: the arithmetic is tiny (2.4 % of the UNet's FLOPs)
// so the kernel is built to stream q/k/v through shared memory at HBM speed.  A persistent CTA (8 warps) walks
// work items = (group of G samples) x (chunk of HC heads).  For every item one warp issues per-row bulk copies
// (global -> shared, mbarrier complete_tx) into one of two stage buffers; while the warps compute item i from
// buffer i & 1 the copies of item i + 1 are in flight.  Each warp owns one (sample, head) pair at a time:
// S = Q K^T in 2 x 4 register tiles, softmax, O = P V with V and the accumulators in registers.
#include <cuda_bf16.h>
#include "attn_math.cuh"

namespace mdt {

struct BulkAttnCfg {
  int G, HC;            // samples per stage, heads per stage
  int strideA, strideB; // padded row strides (elements) of the q rows and the [k | v] rows in shared memory
  int stage_elems;      // elements per stage buffer
  int groups, hchunks;  // work decomposition
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(256, 1) attention_bulk_kernel(const AttnParams p, const BulkAttnCfg c) {
  typedef typename SmemIO<KIND>::T T;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[2];
  T* stage0 = reinterpret_cast<T*>(smem_raw);
  float* scratch = reinterpret_cast<float*>(smem_raw + 2 * (size_t)c.stage_elems * sizeof(T));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nq = p.nq, nk = p.nk, d = p.d;
  const int items = c.groups * c.hchunks;
  float* ss = scratch + (size_t)warp * (((nq * (nk + 1)) + 3) & ~3);

  if (tid == 0) { tc::mbar_init(&full_bar[0], 1); tc::mbar_init(&full_bar[1], 1); tc::fence_barrier_init(); }
  __syncthreads();

  const uint32_t seg = (uint32_t)(c.HC * d * sizeof(T));   // bytes of one q / k / v row segment
  // warp 0 issues the copies of one item: (nq + 2 nk) row segments per sample
  auto issue = [&](int item, int buf) {
    const int grp = item / c.hchunks, hc = item - grp * c.hchunks;
    const int b0 = grp * c.G, gs = min(c.G, p.B - b0), h0 = hc * c.HC;
    T* A = stage0 + (size_t)buf * c.stage_elems;
    T* Bm = A + (size_t)c.G * nq * c.strideA;
    const int per_sample = nq + 2 * nk;
    if (lane == 0) tc::mbar_arrive_expect_tx(&full_bar[buf], (uint32_t)(gs * per_sample) * seg);
    __syncwarp();
    for (int e = lane; e < gs * per_sample; e += 32) {
      const int s = e / per_sample, r = e - s * per_sample;
      const int b = b0 + s;
      if (r < nq) {
        bulk_g2s(A + (size_t)(s * nq + r) * c.strideA, reinterpret_cast<const T*>(p.q) + ((size_t)b * nq + r) * p.ldq + h0 * d, seg,
                 &full_bar[buf]);
      } else {
        const int rr = r - nq, j = rr >> 1, isv = rr & 1;
        const T* base;
        if (p.k_null && b >= p.n_cond) base = reinterpret_cast<const T*>(isv ? p.v_null : p.k_null);
        else base = reinterpret_cast<const T*>(isv ? p.v : p.k) + (size_t)b * p.kv_sample_stride;
        bulk_g2s(Bm + (size_t)(s * nk + j) * c.strideB + isv * c.HC * d, base + (size_t)j * p.ldkv + h0 * d, seg, &full_bar[buf]);
      }
    }
  };

  int i = 0;
  if (warp == 0) {
    if ((int)blockIdx.x < items) issue(blockIdx.x, 0);
    if ((int)(blockIdx.x + gridDim.x) < items) issue(blockIdx.x + gridDim.x, 1);
  }
  for (int item = blockIdx.x; item < items; item += gridDim.x, ++i) {
    const int buf = i & 1;
    tc::mbar_wait(&full_bar[buf], (uint32_t)(i >> 1) & 1u);
    const int grp = item / c.hchunks, hc = item - grp * c.hchunks;
    const int b0 = grp * c.G, gs = min(c.G, p.B - b0), h0 = hc * c.HC;
    const T* A = stage0 + (size_t)buf * c.stage_elems;
    const T* Bm = A + (size_t)c.G * nq * c.strideA;
    for (int pr = warp; pr < gs * c.HC; pr += 8) {
      const int s = pr / c.HC, hl = pr - s * c.HC;
      const T* ks = Bm + (size_t)(s * nk) * c.strideB + hl * d;
      const size_t ob = ((size_t)(b0 + s) * nq) * p.ldo + (size_t)(h0 + hl) * d;
      if (KIND == 0)   // exact fp32 mode stays on the CUDA cores
        attend_head<KIND>(A + (size_t)(s * nq) * c.strideA + hl * d, c.strideA, ks, ks + c.HC * d, c.strideB, ss, nq, nk, d, p.scale,
                          p.o, ob, p.ldo, lane);
      else
        attend_head_mma<KIND, KIND>(A + (size_t)(s * nq) * c.strideA + hl * d, c.strideA, ks, ks + c.HC * d, c.strideB, nq, nk, p.scale,
                              p.o, ob, p.ldo, lane);
    }
    __syncthreads();   // every warp is done with this stage buffer
    const int nxt = item + 2 * gridDim.x;
    if (warp == 0 && nxt < items) issue(nxt, buf);
  }
}

static int g_sms_attn = 0;

// Returns false when the shape does not fit the two-stage shared-memory budget (caller falls back).
static bool bulk_cfg(const AttnParams& p, int kind, BulkAttnCfg* c, size_t* smem) {
  if (p.d != 64 || p.nq > 64 || p.nk > 64 || p.nq < 1 || p.nk < 1) return false;
  const int esz = kind >= 2 ? 2 : 4, pad = kind >= 2 ? 8 : 4;
  const size_t limit = 98 * 1024;
  for (int HC = p.heads; HC >= 1; HC >>= 1) {
    if (p.heads % HC) continue;
    const int sa = HC * p.d + pad, sb = 2 * HC * p.d + pad;
    const size_t per_sample = ((size_t)p.nq * sa + (size_t)p.nk * sb) * esz;
    if (per_sample > limit) continue;
    int G = (int)(limit / per_sample);
    if (G > 8) G = 8;
    c->G = G; c->HC = HC; c->strideA = sa; c->strideB = sb;
    c->stage_elems = (int)((((size_t)G * per_sample + 127) & ~(size_t)127) / esz);
    c->groups = (p.B + G - 1) / G; c->hchunks = p.heads / HC;
    *smem = 2 * (size_t)c->stage_elems * esz + 8 * (size_t)(((p.nq * (p.nk + 1)) + 3) & ~3) * sizeof(float);
    return *smem <= 220 * 1024;
  }
  return false;
}

bool attention_bulk_supported(const AttnParams& p, int kind) {
  BulkAttnCfg c; size_t smem;
  // every source segment must be 16-byte aligned for cp.async.bulk
  const int esz = kind >= 2 ? 2 : 4;
  if ((p.ldq * esz) % 16 || (p.ldkv * esz) % 16 || (p.d * esz) % 16) return false;
  return bulk_cfg(p, kind, &c, &smem);
}

cudaError_t init_attention_bulk() {
  cudaError_t e = cudaFuncSetAttribute(attention_bulk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_bulk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_bulk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_bulk_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  return e;
}

cudaError_t launch_attention_bulk(const AttnParams& p, int kind, cudaStream_t s) {
  if (p.B <= 0) return cudaSuccess;
  BulkAttnCfg c; size_t smem;
  if (!bulk_cfg(p, kind, &c, &smem)) return cudaErrorInvalidValue;
  if (g_sms_attn == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_attn, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms_attn <= 0) g_sms_attn = 148;
  }
  const int items = c.groups * c.hchunks;
  const unsigned grid = (unsigned)(items < g_sms_attn ? items : g_sms_attn);
  if (kind == 0) attention_bulk_kernel<0><<<grid, 256, smem, s>>>(p, c);
  else if (kind == 1) attention_bulk_kernel<1><<<grid, 256, smem, s>>>(p, c);
  else if (kind == 2) attention_bulk_kernel<2><<<grid, 256, smem, s>>>(p, c);
  else attention_bulk_kernel<3><<<grid, 256, smem, s>>>(p, c);
  return cudaGetLastError();
}

}  // namespace mdt
