// Microbenchmarks that size the merge kernels (run on the B200 box):
//   1. scalar FMUL+FADD (what __fmul_rn/__fadd_rn compile to) vs packed FFMA2 pairs issued as
//      fma(a,b,-0) ; fma(p,1,acc) with run-time constants (exactly mul.rn then add.rn, unfusable)
//   2. cost of cooperative grid.sync() and cluster.sync()
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

constexpr int ITERS = 4096;
__global__ void k_scalar(float* out, float w, float x0) {
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
  float x = x0 + threadIdx.x;
#pragma unroll 8
  for (int i = 0; i < ITERS; ++i) {
    a0 = __fadd_rn(a0, __fmul_rn(w, x)); a1 = __fadd_rn(a1, __fmul_rn(a0, x)); a2 = __fadd_rn(a2, __fmul_rn(w, a1)); a3 = __fadd_rn(a3, __fmul_rn(a2, x));
    a4 = __fadd_rn(a4, __fmul_rn(w, a3)); a5 = __fadd_rn(a5, __fmul_rn(a4, x)); a6 = __fadd_rn(a6, __fmul_rn(w, a5)); a7 = __fadd_rn(a7, __fmul_rn(a6, x));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// independent chains version (throughput, not latency): 8 accumulators, products of loop-invariant-ish inputs
__global__ void k_scalar_tp(float* out, const float* in) {
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float x[8];
  for (int k = 0; k < 8; ++k) x[k] = in[threadIdx.x + 32 * k];
  float w = in[0];
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __fadd_rn(a[k], __fmul_rn(x[k], w));
    w = __fadd_rn(w, 1e-9f);
  }
  float s = 0;
  for (int k = 0; k < 8; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed_tp(float* out, const float* in, float one, float nz) {
  uint64_t a[4];
  uint64_t x[4];
  for (int k = 0; k < 4; ++k) { a[k] = pack2(0.f, 0.f); x[k] = pack2(in[threadIdx.x + 64 * k], in[threadIdx.x + 64 * k + 32]); }
  const uint64_t one2 = pack2(one, one), nz2 = pack2(nz, nz);
  float w = in[0];
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
    const uint64_t w2 = pack2(w, w);
#pragma unroll
    for (int k = 0; k < 4; ++k) { uint64_t p = fma2(x[k], w2, nz2); a[k] = fma2(p, one2, a[k]); }
    w = __fadd_rn(w, 1e-9f);
  }
  float s = 0;
  for (int k = 0; k < 4; ++k) { float lo, hi; unpack2(a[k], lo, hi); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// bit-exactness check of the packed form against the scalar form
__global__ void k_check(const float* x, const float* y, int n, float one, float nz, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i * 2 + 1 >= n) return;
  float s0 = 0.f, s1 = 0.f;
  uint64_t acc = pack2(0.f, 0.f);
  const uint64_t one2 = pack2(one, one), nz2 = pack2(nz, nz);
  for (int k = 0; k < 64; ++k) {
    float a0 = x[(2 * i + k * 7) % n], a1 = x[(2 * i + 1 + k * 5) % n], b = y[(i + k * 3) % n];
    s0 = __fadd_rn(s0, __fmul_rn(a0, b));
    s1 = __fadd_rn(s1, __fmul_rn(a1, b));
    uint64_t p = fma2(pack2(a0, a1), pack2(b, b), nz2);
    acc = fma2(p, one2, acc);
  }
  float lo, hi; unpack2(acc, lo, hi);
  if (__float_as_uint(lo) != __float_as_uint(s0) || __float_as_uint(hi) != __float_as_uint(s1)) atomicAdd(bad, 1);
}
__global__ void k_gridsync(int n, long long* cyc) {
  cg::grid_group g = cg::this_grid();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) g.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = clock64() - t0;
}
__global__ void k_clustersync(int n, long long* cyc) {
  cg::cluster_group c = cg::this_cluster();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) c.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = clock64() - t0;
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  float *out, *in; cudaMalloc(&out, sizeof(float) * sms * 8 * 256); cudaMalloc(&in, sizeof(float) * 4096);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = (float)((i * 2654435761u) % 1000) / 997.f - 0.5f;
  cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  const int grid = sms * 8, block = 256;
  double pairs = (double)grid * block * ITERS * 8;  // (mul,add) pairs per launch
  float ms1 = timeit([&] { k_scalar_tp<<<grid, block>>>(out, in); });
  float ms2 = timeit([&] { k_packed_tp<<<grid, block>>>(out, in, 1.0f, -0.0f); });
  printf("SMs %d clock %.0f MHz\n", sms, p.clockRate / 1000.0);
  printf("scalar FMUL+FADD : %.3f ms -> %.2f T mul-add pairs/s\n", ms1, pairs / ms1 / 1e9);
  printf("packed FFMA2 x2  : %.3f ms -> %.2f T mul-add pairs/s\n", ms2, pairs / ms2 / 1e9);
  int* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
  k_check<<<64, 256>>>(in, in + 1000, 3000, 1.0f, -0.0f, bad);
  int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  printf("packed vs scalar mismatches: %d\n", hb);
  long long* cyc; cudaMalloc(&cyc, 8); long long hc;
  for (int per_sm = 1; per_sm <= 2; ++per_sm) {
    int n = 1000, g = sms * per_sm; void* args[] = {&n, &cyc};
    cudaLaunchCooperativeKernel((void*)k_gridsync, dim3(g), dim3(256), args); cudaDeviceSynchronize();
    cudaLaunchCooperativeKernel((void*)k_gridsync, dim3(g), dim3(256), args); cudaDeviceSynchronize();
    cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("grid.sync %d CTAs: %.0f cycles each (%s)\n", g, hc / 1000.0, cudaGetErrorString(cudaGetLastError()));
  }
  for (int cs : {2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(cs * 4); cfg.blockDim = dim3(256);
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cs > 8) cudaFuncSetAttribute(k_clustersync, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int n = 1000; void* args[] = {&n, &cyc};
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)k_clustersync, args); cudaDeviceSynchronize();
    cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    int maxc = 0; cudaOccupancyMaxActiveClusters(&maxc, (const void*)k_clustersync, &cfg);
    printf("cluster.sync size %d: %.0f cycles each (%s), max active clusters %d\n", cs, hc / 1000.0, cudaGetErrorString(e), maxc);
  }
  return 0;
}
