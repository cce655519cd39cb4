// Stand-alone check and timing of k_sign_umma (kmerlsh_b200/csrc/sign_umma.cuh): keys against the reference's
// mul-then-add chain computed on the host, then GB/s over repeated launches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -lineinfo -o sign_umma_test sign_umma_test.cu
//   [SIGN_PIPE=1] [SIGN_CTAS=n] ./sign_umma_test [rows] [D] [H] [indirect 0/1]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <random>
#include <vector>

#include "../../kmerlsh_b200/csrc/sign_umma.cuh"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);       \
      return 2;                                                                            \
    }                                                                                      \
  } while (0)

template <int KW, bool PIPE>
int run(uint64_t n, int D, int H, bool indirect) {
  const int ld = (D + 3) & ~3;
  std::mt19937 g(1234 + D * 7 + H);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<float> vals(n * ld, 0.f), planes((size_t)H * ld, 0.f);
  for (uint64_t r = 0; r < n; ++r)
    for (int j = 0; j < D; ++j) vals[r * ld + j] = nd(g) * ((r % 97 == 0) ? 1e-3f : 1.f) + ((r % 5 == 0) ? 2.f : 0.f);
  for (int h = 0; h < H; ++h)
    for (int j = 0; j < D; ++j) planes[(size_t)h * ld + j] = nd(g);
  // rows that sit on or next to a hyperplane: x orthogonal to plane 0 up to rounding
  for (uint64_t r = 3; r < n; r += 1000) {
    float dot = 0.f, ww = 0.f;
    for (int j = 0; j < D; ++j) { dot += planes[j] * vals[r * ld + j]; ww += planes[j] * planes[j]; }
    for (int j = 0; j < D; ++j) vals[r * ld + j] -= dot / ww * planes[j];
  }
  std::vector<uint32_t> rows(n);
  for (uint64_t r = 0; r < n; ++r) rows[r] = indirect ? (uint32_t)((r * 2654435761ull) % n) : (uint32_t)r;
  std::vector<uint32_t> want(n);
  for (uint64_t t = 0; t < n; ++t) {
    const float* x = &vals[(uint64_t)rows[t] * ld];
    uint32_t key = 0;
    for (int h = 0; h < H; ++h) {
      volatile float sum = 0.f;
      for (int j = 0; j < D; ++j) {
        volatile float p = planes[(size_t)h * ld + j] * x[j];
        sum = sum + p;
      }
      key = key * 2 + (sum >= 0.f ? 1u : 0u);
    }
    want[t] = key;
  }
  float *d_vals, *d_planes;
  uint32_t *d_rows, *d_keys, *d_rout;
  unsigned long long* d_eps;
  CK(cudaMalloc(&d_vals, vals.size() * 4));
  CK(cudaMalloc(&d_planes, planes.size() * 4));
  CK(cudaMalloc(&d_rows, n * 4));
  CK(cudaMalloc(&d_keys, n * 4));
  CK(cudaMalloc(&d_rout, n * 4));
  CK(cudaMalloc(&d_eps, 8));
  CK(cudaMemcpy(d_vals, vals.data(), vals.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_planes, planes.data(), planes.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_rows, rows.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_eps, 0, 8));
  CK(cudaMemset(d_keys, 0xFF, n * 4));
  auto fn = sign_umma::k_sign_umma<KW, PIPE>;
  const size_t smem = sign_umma::smem_bytes<KW>();
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0, sms = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, sign_umma::kThreads, smem));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const uint64_t ntiles = (n + 127) / 128;
  const int mult = getenv("SIGN_CTAS") ? atoi(getenv("SIGN_CTAS")) : sign_umma::ctas_per_sm<KW, PIPE>();
  const unsigned grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)sms * mult);
  printf("%s KW=%d D=%d H=%d n=%llu indirect=%d smem=%zu ctas/sm=%d grid=%u\n", PIPE ? "PIPE" : "PLAIN", KW, D, H, (unsigned long long)n, (int)indirect, smem, per_sm, grid);
  fn<<<grid, sign_umma::kThreads, smem>>>(d_vals, D, ld, indirect ? d_rows : nullptr, n, d_planes, H, d_keys, d_rout, d_eps, 0u, nullptr);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<uint32_t> got(n), rout(n);
  unsigned long long eps = 0;
  CK(cudaMemcpy(got.data(), d_keys, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(rout.data(), d_rout, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&eps, d_eps, 8, cudaMemcpyDeviceToHost));
  uint64_t bad = 0, badr = 0;
  for (uint64_t t = 0; t < n; ++t) {
    if (got[t] != want[t] && bad++ < 5) printf("  key mismatch at %llu: got %08x want %08x\n", (unsigned long long)t, got[t], want[t]);
    if (rout[t] != rows[t]) ++badr;
  }
  printf("  key mismatches %llu, row-index mismatches %llu, eps-margin rows %llu (%.3g of rows)\n", (unsigned long long)bad,
         (unsigned long long)badr, eps, (double)eps / (double)n);
  {
    unsigned long long* d_prof;
    CK(cudaMalloc(&d_prof, 64));
    CK(cudaMemset(d_prof, 0, 64));
    fn<<<grid, sign_umma::kThreads, smem>>>(d_vals, D, ld, indirect ? d_rows : nullptr, n, d_planes, H, d_keys, d_rout, d_eps, 0u, d_prof);
    CK(cudaDeviceSynchronize());
    unsigned long long hp[8];
    CK(cudaMemcpy(hp, d_prof, 64, cudaMemcpyDeviceToHost));
    if (hp[4])
      printf("  thread 0, cycles per tile: gather wait %.0f, split+barrier %.0f, products %.0f, epilogue %.0f (%llu tiles)\n",
             (double)hp[0] / hp[4], (double)hp[1] / hp[4], (double)hp[2] / hp[4], (double)hp[3] / hp[4], hp[4]);
    cudaFree(d_prof);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int reps = 20;
  cudaEventRecord(e0);
  for (int k = 0; k < reps; ++k) fn<<<grid, sign_umma::kThreads, smem>>>(d_vals, D, ld, indirect ? d_rows : nullptr, n, d_planes, H, d_keys, d_rout, d_eps, 0u, nullptr);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  printf("  %.3f ms per launch, %.1f GB/s algorithmic (4D+8 per row)\n", ms, (double)n * (4.0 * D + 8.0) / (ms * 1e-3) / 1e9);
  cudaFree(d_vals); cudaFree(d_planes); cudaFree(d_rows); cudaFree(d_keys); cudaFree(d_rout); cudaFree(d_eps);
  return (bad || badr) ? 1 : 0;
}

int main(int argc, char** argv) {
  const uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 4000000ull;
  const int D = argc > 2 ? atoi(argv[2]) : 32, H = argc > 3 ? atoi(argv[3]) : 25;
  const bool indirect = argc > 4 && atoi(argv[4]) != 0;
  const int ld = (D + 3) & ~3;
  const bool pipe = getenv("SIGN_PIPE") && atoi(getenv("SIGN_PIPE"));
  if (ld <= 32) return pipe ? run<32, true>(n, D, H, indirect) : run<32, false>(n, D, H, indirect);
  if (ld <= 64) return pipe ? run<64, true>(n, D, H, indirect) : run<64, false>(n, D, H, indirect);
  printf("D too wide for this kernel\n");
  return 3;
}
