// tmem_probe.cu -- empirical register <-> (lane, column) map of tcgen05.ld.16x256b on sm_100a (there is no GPU on the build box
// and the PTX figure is not at hand): every TMEM cell is filled with lane * 1000 + column through the well-understood 32x32b
// store, then read back with 16x256b.x4 at lane offsets 0 and 16 of each warp's quadrant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe/tmem_probe tools/probe/tmem_probe.cu && tools/probe/tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(float* out) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = base_s;
  uint32_t v[32];
  for (int j = 0; j < 32; ++j) v[j] = __float_as_uint((float)((warp * 32 + lane) * 1000 + j));
  const uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int half = 0; half < 2; ++half) {
    uint32_t r[16];
    const uint32_t a = base + ((uint32_t)(warp * 32 + half * 16) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(a));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[((warp * 2 + half) * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64u) : "memory");
}

int main() {
  float* d; const int n = 4 * 2 * 32 * 16;
  cudaMalloc(&d, n * sizeof(float));
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  static float h[4 * 2 * 32 * 16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int w = 0; w < 4; ++w) for (int half = 0; half < 2; ++half) for (int t = 0; t < 32; ++t) for (int i = 0; i < 16; ++i) {
    const int val = (int)h[((w * 2 + half) * 32 + t) * 16 + i];
    const int row = val / 1000, col = val % 1000;
    // expected (mma.m16n8 accumulator layout per 8-column group c = i / 4): regs 4c+0, 4c+1 -> row t/4, cols 8c + 2(t%4) + {0,1};
    // regs 4c+2, 4c+3 -> row t/4 + 8
    const int c = i / 4, k = i % 4;
    const int erow = w * 32 + half * 16 + t / 4 + (k >= 2 ? 8 : 0), ecol = 8 * c + 2 * (t % 4) + (k & 1);
    if (row != erow || col != ecol) { if (bad < 12) printf("w%d half%d t%2d r%2d: got (row %d, col %d) expected (%d, %d)\n", w, half, t, i, row, col, erow, ecol); ++bad; }
  }
  printf("tcgen05.ld.16x256b.x4 map: %s (%d mismatches)\n", bad ? "DIFFERS from the m16n8 accumulator layout" : "matches the m16n8 accumulator layout", bad);
  if (bad) for (int t = 0; t < 8; ++t) { printf("warp0 half0 t%d:", t); for (int i = 0; i < 16; ++i) printf(" %d", (int)h[t * 16 + i]); printf("\n"); }
  return 0;
}
