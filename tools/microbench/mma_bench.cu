// Microbenchmark: legacy mma.sync.m16n8k16 (fp16 in, fp32 accumulate) issue rate per SM on sm_100a,
// in the shape the merge screen uses it (4 row tiles x 2 k-steps per 8-representative group).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool RESET>
__global__ void __launch_bounds__(256, 2) k_mma(int iters, uint32_t seed, float* out, long long* cycles) {
  uint32_t af[4][2][4];
  for (int mt = 0; mt < 4; ++mt)
    for (int ks = 0; ks < 2; ++ks)
      for (int e = 0; e < 4; ++e) af[mt][ks][e] = 0x3c003c00u ^ (seed * (mt * 8 + ks * 4 + e + threadIdx.x));
  float c[4][4];
  for (int mt = 0; mt < 4; ++mt)
    for (int e = 0; e < 4; ++e) c[mt][e] = 0.f;
  uint32_t b0 = 0x38003800u ^ seed, b1 = 0x34003400u ^ seed;
  float keep = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      if (RESET)
        for (int e = 0; e < 4; ++e) c[mt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) mma_16816(c[mt], af[mt][ks], b0, b1);
      if (RESET)
        for (int e = 0; e < 4; ++e)
          if (!(c[mt][e] < 0.8f)) keep += 1.f;
    }
    b0 += 0x00010001u;
  }
  const long long t1 = clock64();
  float s = keep;
  for (int mt = 0; mt < 4; ++mt)
    for (int e = 0; e < 4; ++e) s += c[mt][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int grid = p.multiProcessorCount * 2, iters = 20000;
  float* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * grid * 256);
  cudaMalloc(&cyc, sizeof(long long) * grid);
  for (int variant = 0; variant < 2; ++variant) {
    for (int rep = 0; rep < 2; ++rep) {
      if (variant == 0) k_mma<false><<<grid, 256>>>(iters, 0u, out, cyc);
      else k_mma<true><<<grid, 256>>>(iters, 0u, out, cyc);
      cudaDeviceSynchronize();
    }
    long long h[4096];
    cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; ++i) avg += (double)h[i];
    avg /= grid;
    // per SM: 2 CTAs x 8 warps x iters x 8 mma, each 16*8*16 = 2048 FMA... (m16n8k16 = 2048 MACs)
    const double macs = 2.0 * 8 * iters * 8.0 * 2048.0;
    printf("%s: %.0f cycles/CTA -> %.1f MAC/clk/SM, %.2f cycles per mma per SM, %.1f (64x8 pair-groups, K=32)/kclk/SM\n",
           variant ? "screen-shaped (reset + threshold test)" : "accumulate only", avg, macs / avg, avg / (2.0 * 8 * iters * 8.0),
           2.0 * 8 * iters / avg * 1e3);
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
