// Mode E statistics on the clusters the hot path produced.
//
// Reference: app/kmerLSH.cc:541-585 reads the clustering result (IOMat::ReadClusterAll), calls
// AB::WRS (function/funcAB.cc:73-109) on every cluster — alglib::studentttest2 on the group-A and group-B
// halves of the centroid for clusters with more than size_thresh members, ids filed under group B when
// lefttail <= pvalue_thresh, else under group A when righttail <= pvalue_thresh — and then walks kmer_set.hex
// keeping the k-mers whose id is in either set.  Here:
//   klsh_ttest            k_ttest2: one thread per cluster of the context's working set
//   klsh_differential_ids chain ranking (kernels.cu) -> every member slot knows its cluster -> one label per k-mer id
//   klsh_select_kmers     order-preserving two-way compaction of the k-mer records by label
//
// Arithmetic.  The test statistic follows alglib-3.15.0 statistics.cpp:12502-12616 operation by operation in
// IEEE binary64 (__dadd_rn/__dmul_rn/__ddiv_rn/__dsqrt_rn, no contraction): it is bit-identical to the
// reference's.  The Student distribution follows specialfunctions.cpp:9559-9631 for t >= -2 (finite series,
// atan/sqrt); for t < -2 ALGLIB calls Cephes' incomplete beta function, which is evaluated here with the modified
// Lentz continued fraction instead.  The tail probabilities therefore agree with ALGLIB's to a relative 1e-9
// (measured: a few 1e-13), not bit for bit; a cluster whose tail lies within 1e-9 (relative) of the threshold
// is counted in klsh_ttest_stats.margin so that a caller can tell when a decision rested on that tolerance.
#include <algorithm>
#include <cstring>
#include <vector>

#include "klsh_internal.cuh"

namespace {

#define SLAUNCH(ctx)                                                                                     \
  do {                                                                                                   \
    (ctx)->launches++;                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                                \
    if (e__ != cudaSuccess)                                                                              \
      return klsh_fail((ctx), KLSH_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                       __FILE__, __LINE__);                                                              \
  } while (0)

__device__ double betacf(double a, double b, double x) {
  const double tiny = 1e-300;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 2000; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < 2e-16) break;
  }
  return h;
}

// regularised incomplete beta I_x(a, b)
__device__ double incbeta(double a, double b, double x) {
  if (x <= 0.0) return 0.0;
  if (x >= 1.0) return 1.0;
  const double lbt = lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log1p(-x);
  const double bt = exp(lbt);
  if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
  return 1.0 - bt * betacf(b, a, 1.0 - x) / b;
}

// alglib studenttdistribution(k, t), specialfunctions.cpp:9559-9631
__device__ double student_t_cdf(int k, double t) {
  if (t == 0.0) return 0.5;
  const double rk = (double)k;
  if (t < -2.0) {
    const double z = __ddiv_rn(rk, __dadd_rn(rk, __dmul_rn(t, t)));
    return 0.5 * incbeta(0.5 * rk, 0.5, z);
  }
  const double x = t < 0.0 ? -t : t;
  const double z = __dadd_rn(1.0, __ddiv_rn(__dmul_rn(x, x), rk));
  double p, f, tz;
  int j;
  if (k % 2 != 0) {
    const double xsqk = __ddiv_rn(x, __dsqrt_rn(rk));
    p = atan(xsqk);
    if (k > 1) {
      f = 1.0;
      tz = 1.0;
      j = 3;
      while (j <= k - 2 && __ddiv_rn(tz, f) > 5E-16) {
        tz = __dmul_rn(tz, __ddiv_rn((double)(j - 1), __dmul_rn(z, (double)j)));
        f = __dadd_rn(f, tz);
        j += 2;
      }
      p = __dadd_rn(p, __ddiv_rn(__dmul_rn(f, xsqk), z));
    }
    p = __ddiv_rn(__dmul_rn(p, 2.0), 3.14159265358979323846);
  } else {
    f = 1.0;
    tz = 1.0;
    j = 2;
    while (j <= k - 2 && __ddiv_rn(tz, f) > 5E-16) {
      tz = __dmul_rn(tz, __ddiv_rn((double)(j - 1), __dmul_rn(z, (double)j)));
      f = __dadd_rn(f, tz);
      j += 2;
    }
    p = __ddiv_rn(__dmul_rn(f, x), __dsqrt_rn(__dmul_rn(z, rk)));
  }
  if (t < 0.0) p = -p;
  return __dadd_rn(0.5, __dmul_rn(0.5, p));
}

// alglib studentttest2 on the two halves of one row, statistics.cpp:12502-12616 (x = row[0..n), y = row[n..n+m))
__device__ void ttest2_row(const float* __restrict__ row, int n, int m, double* left, double* right) {
  if (n <= 0 || m <= 0) {
    *left = 1.0;
    *right = 1.0;
    return;
  }
  double xmean = 0.0;
  const double x0 = (double)row[0];
  bool samex = true;
  for (int i = 0; i < n; ++i) {
    const double v = (double)row[i];
    xmean = __dadd_rn(xmean, v);
    samex = samex && (v == x0);
  }
  xmean = samex ? x0 : __ddiv_rn(xmean, (double)n);
  double ymean = 0.0;
  const double y0 = (double)row[n];
  bool samey = true;
  for (int i = 0; i < m; ++i) {
    const double v = (double)row[n + i];
    ymean = __dadd_rn(ymean, v);
    samey = samey && (v == y0);
  }
  ymean = samey ? y0 : __ddiv_rn(ymean, (double)m);
  double s = 0.0;
  if (n + m > 2) {
    for (int i = 0; i < n; ++i) {
      const double d = __dsub_rn((double)row[i], xmean);
      s = __dadd_rn(s, __dmul_rn(d, d));
    }
    for (int i = 0; i < m; ++i) {
      const double d = __dsub_rn((double)row[n + i], ymean);
      s = __dadd_rn(s, __dmul_rn(d, d));
    }
    const double w = __dadd_rn(__ddiv_rn(1.0, (double)n), __ddiv_rn(1.0, (double)m));
    s = __dsqrt_rn(__ddiv_rn(__dmul_rn(s, w), (double)(n + m - 2)));
  }
  if (s == 0.0) {
    *left = xmean >= ymean ? 1.0 : 0.0;
    *right = xmean <= ymean ? 1.0 : 0.0;
    return;
  }
  const double stat = __ddiv_rn(__dsub_rn(xmean, ymean), s);
  const double p = student_t_cdf(n + m - 2, stat);
  *left = p;
  *right = __dsub_rn(1.0, p);
}

// counts: {tested, rows in group A, rows in group B, ids in group A, ids in group B, margin}
__global__ void k_ttest2(const float* __restrict__ vals, int ld, const MetaCol cnt, const uint32_t* __restrict__ alive, uint64_t n,
                         int n1, int n2, double pthr, unsigned long long size_thr, uint8_t* group, double* left, double* right,
                         unsigned long long* counts) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t row = alive[r];
  const unsigned long long members = (unsigned long long)(uint32_t)cnt[row];
  uint8_t g = 0;
  double l = -1.0, rt = -1.0;
  if (members > size_thr) {  // `ids.size() > size_thresh`, the int converted to size_t (funcAB.cc:87)
    ttest2_row(vals + (uint64_t)row * ld, n1, n2, &l, &rt);
    if (l <= pthr) g = 2;        // funcAB.cc:100-101: the ids go to the second set
    else if (rt <= pthr) g = 1;  // :102-103
    atomicAdd(counts + 0, 1ull);
    if (g) {
      atomicAdd(counts + g, 1ull);
      atomicAdd(counts + 2 + g, members);
    }
    const double band = 1e-9 * fabs(pthr);
    if (fabs(l - pthr) <= band || fabs(rt - pthr) <= band) atomicAdd(counts + 5, 1ull);
  }
  group[r] = g;
  if (left) left[r] = l;
  if (right) right[r] = rt;
}

// every member slot of a labelled cluster marks its k-mer id: bit 0 = first set, bit 1 = second set
__global__ void k_label_ids(const int32_t* __restrict__ slot_row, const uint8_t* __restrict__ group, uint32_t m,
                            const uint64_t* __restrict__ ids, uint64_t id_base, uint64_t n_kmers, uint32_t* label_words) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= m) return;
  const int32_t r = slot_row[s];
  if (r < 0) return;
  const uint32_t g = group[r];
  if (!g) return;
  const uint64_t id = ids ? ids[s] : id_base + s;
  if (id >= n_kmers) return;  // the join only asks about ids below kmap_size (app/kmerLSH.cc:568)
  atomicOr(label_words + (id >> 2), g << (8u * (uint32_t)(id & 3u)));
}

// bits -> label with the precedence of the join (app/kmerLSH.cc:571-576): the first set wins
__global__ void k_label_finish(uint32_t* label_words, uint64_t n_words) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_words) return;
  const uint32_t w = label_words[i];
  if (!w) return;
  uint32_t out = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const uint32_t v = (w >> (8 * b)) & 0xFFu;
    out |= ((v & 1u) ? 1u : ((v & 2u) ? 2u : 0u)) << (8 * b);
  }
  label_words[i] = out;
}

constexpr int kSelTile = 1024;

__device__ __forceinline__ uint32_t block_scan2(uint32_t v, uint32_t* ws, uint32_t* total) {  // exclusive scan over kSelTile threads
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += t;
  }
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = ws[lane];
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (uint32_t)o) winc += t;
    }
    ws[lane] = winc - w;
    if (lane == 31) ws[32] = winc;
  }
  __syncthreads();
  const uint32_t res = ws[warp] + inc - v;
  *total = ws[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kSelTile) k_select_count(const uint8_t* __restrict__ label, uint64_t n, uint32_t* blk_a, uint32_t* blk_b) {
  __shared__ uint32_t ws[33];
  const uint64_t i = (uint64_t)blockIdx.x * kSelTile + threadIdx.x;
  const uint32_t l = i < n ? label[i] : 0u;
  uint32_t ta, tb;
  block_scan2(l == 1u ? 1u : 0u, ws, &ta);
  block_scan2(l == 2u ? 1u : 0u, ws, &tb);
  if (threadIdx.x == 0) {
    blk_a[blockIdx.x] = ta;
    blk_b[blockIdx.x] = tb;
  }
}

// exclusive scan of both block-count arrays by one block; totals to tot[0], tot[1]
__global__ void __launch_bounds__(kSelTile) k_select_scan(uint32_t* blk_a, uint32_t* blk_b, uint32_t nblk, unsigned long long* tot) {
  __shared__ uint32_t ws[33];
  uint32_t ca = 0, cb = 0;
  for (uint32_t base = 0; base < nblk; base += kSelTile) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t va = i < nblk ? blk_a[i] : 0u, vb = i < nblk ? blk_b[i] : 0u;
    uint32_t ta, tb;
    const uint32_t ea = block_scan2(va, ws, &ta);
    const uint32_t eb = block_scan2(vb, ws, &tb);
    if (i < nblk) {
      blk_a[i] = ca + ea;
      blk_b[i] = cb + eb;
    }
    ca += ta;
    cb += tb;
  }
  if (threadIdx.x == 0) {
    tot[0] = ca;
    tot[1] = cb;
  }
}

__global__ void __launch_bounds__(kSelTile) k_select_write(const uint8_t* __restrict__ rec, const uint8_t* __restrict__ label, uint64_t n,
                                                            int rb, const uint32_t* __restrict__ off_a, const uint32_t* __restrict__ off_b,
                                                            uint8_t* out_a, uint8_t* out_b) {
  __shared__ uint32_t ws[33];
  const uint64_t i = (uint64_t)blockIdx.x * kSelTile + threadIdx.x;
  const uint32_t l = i < n ? label[i] : 0u;
  uint32_t ta, tb;
  const uint32_t ea = block_scan2(l == 1u ? 1u : 0u, ws, &ta);
  const uint32_t eb = block_scan2(l == 2u ? 1u : 0u, ws, &tb);
  if (l != 1u && l != 2u) return;
  uint8_t* dst = l == 1u ? out_a + (uint64_t)(off_a[blockIdx.x] + ea) * rb : out_b + (uint64_t)(off_b[blockIdx.x] + eb) * rb;
  const uint8_t* src = rec + i * (uint64_t)rb;
  if (rb == 8) {  // Kmer::MAX_K = 32: MAX_K/4 bytes per record (kmer/Kmer.h:68-78)
    *reinterpret_cast<uint64_t*>(dst) = *reinterpret_cast<const uint64_t*>(src);
  } else {
    for (int b = 0; b < rb; ++b) dst[b] = src[b];
  }
}

int run_ttest(klsh_ctx* ctx, int n1, int n2, float pvalue_thresh, int size_thresh, bool want_tails) {
  const uint64_t n = ctx->cur.n_alive;
  if (n1 < 0 || n2 < 0 || (n && (int64_t)n1 + n2 > ctx->D))
    return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_ttest: %d + %d samples but rows have %d values", n1, n2, ctx->D);
  KTRY(dev_reserve(ctx, ctx->st_counts, sizeof(unsigned long long) * 8));
  KCUDA(ctx, cudaMemsetAsync(ctx->st_counts.p, 0, sizeof(unsigned long long) * 8, ctx->stream));
  if (!n) return KLSH_OK;
  KTRY(dev_reserve(ctx, ctx->st_group, n));
  if (want_tails) {
    KTRY(dev_reserve(ctx, ctx->st_left, sizeof(double) * n));
    KTRY(dev_reserve(ctx, ctx->st_right, sizeof(double) * n));
  }
  k_ttest2<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
      ctx->cur.vals.as<float>(), ctx->ld, ctx->cur.cnt(), ctx->cur.alive.as<uint32_t>(), n, n1, n2, (double)pvalue_thresh,
      (unsigned long long)(long long)size_thresh, ctx->st_group.as<uint8_t>(), want_tails ? ctx->st_left.as<double>() : nullptr,
      want_tails ? ctx->st_right.as<double>() : nullptr, ctx->st_counts.as<unsigned long long>());
  SLAUNCH(ctx);
  return KLSH_OK;
}

int fetch_stats(klsh_ctx* ctx, klsh_ttest_stats* stats) {
  if (!stats) return KLSH_OK;
  unsigned long long h[8];
  KCUDA(ctx, cudaMemcpyAsync(h, ctx->st_counts.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  stats->rows = ctx->cur.n_alive;
  stats->tested = h[0];
  stats->rows_a = h[1];
  stats->rows_b = h[2];
  stats->ids_a = h[3];
  stats->ids_b = h[4];
  stats->margin = h[5];
  return KLSH_OK;
}

}  // namespace

extern "C" int klsh_ttest(klsh_ctx* ctx, int num_sample1, int num_sample2, float pvalue_thresh, int size_thresh,
                          uint8_t* row_group, double* lefttail, double* righttail, klsh_ttest_stats* stats) {
  if (!ctx) return KLSH_ERR_ARG;
  KCUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t n = ctx->cur.n_alive;
  KTRY(run_ttest(ctx, num_sample1, num_sample2, pvalue_thresh, size_thresh, lefttail || righttail));
  if (n) {
    if (row_group) KCUDA(ctx, cudaMemcpyAsync(row_group, ctx->st_group.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (lefttail) KCUDA(ctx, cudaMemcpyAsync(lefttail, ctx->st_left.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (righttail) KCUDA(ctx, cudaMemcpyAsync(righttail, ctx->st_right.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return fetch_stats(ctx, stats);
}

extern "C" int klsh_differential_ids(klsh_ctx* ctx, int num_sample1, int num_sample2, float pvalue_thresh, int size_thresh,
                                     uint64_t n_kmers, uint8_t* id_label, klsh_ttest_stats* stats) {
  if (!ctx || (!id_label && n_kmers)) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_differential_ids: bad argument");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t n = ctx->cur.n_alive;
  KTRY(run_ttest(ctx, num_sample1, num_sample2, pvalue_thresh, size_thresh, false));
  if (n_kmers) {
    const uint64_t n_words = (n_kmers + 3) / 4;
    KTRY(dev_reserve(ctx, ctx->st_label, sizeof(uint32_t) * n_words));
    KCUDA(ctx, cudaMemsetAsync(ctx->st_label.p, 0, sizeof(uint32_t) * n_words, ctx->stream));
    const uint32_t m = (uint32_t)ctx->n_slots;
    if (n && m) {
      KTRY(dev_reserve(ctx, ctx->st_slot_row, sizeof(int32_t) * (size_t)m));
      KTRY(launch_rank_chains(ctx, n, nullptr, nullptr, ctx->st_slot_row.as<int32_t>()));
      const uint64_t* d_ids = nullptr;
      if (!ctx->ids_implicit) {
        if (ctx->ids.size() < m) return klsh_fail(ctx, KLSH_ERR_ARG, "internal: %zu member ids for %u slots", ctx->ids.size(), m);
        KTRY(dev_reserve(ctx, ctx->st_ids, sizeof(uint64_t) * (size_t)m));
        KCUDA(ctx, cudaMemcpyAsync(ctx->st_ids.p, ctx->ids.data(), sizeof(uint64_t) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        d_ids = ctx->st_ids.as<uint64_t>();
      }
      k_label_ids<<<(m + 255) / 256, 256, 0, ctx->stream>>>(ctx->st_slot_row.as<int32_t>(), ctx->st_group.as<uint8_t>(), m, d_ids,
                                                           ctx->id_base, n_kmers, ctx->st_label.as<uint32_t>());
      SLAUNCH(ctx);
      k_label_finish<<<(unsigned)((n_words + 255) / 256), 256, 0, ctx->stream>>>(ctx->st_label.as<uint32_t>(), n_words);
      SLAUNCH(ctx);
    }
    KCUDA(ctx, cudaMemcpyAsync(id_label, ctx->st_label.p, n_kmers, cudaMemcpyDeviceToHost, ctx->stream));
  }
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return fetch_stats(ctx, stats);
}

extern "C" int klsh_select_kmers(klsh_ctx* ctx, const uint8_t* records, uint64_t n_kmers, int record_bytes, const uint8_t* id_label,
                                 uint8_t* out_a, uint64_t* n_a, uint8_t* out_b, uint64_t* n_b) {
  if (!ctx || record_bytes <= 0 || !n_a || !n_b || (n_kmers && (!records || !id_label || !out_a || !out_b)))
    return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_select_kmers: bad argument");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  *n_a = 0;
  *n_b = 0;
  const uint64_t chunk = 8ull << 20;  // records per pass
  const uint64_t cap = std::min<uint64_t>(chunk, n_kmers);
  if (!cap) return KLSH_OK;
  const uint32_t nblk_max = (uint32_t)((cap + kSelTile - 1) / kSelTile);
  KTRY(dev_reserve(ctx, ctx->st_rec, cap * (uint64_t)record_bytes + 8));
  KTRY(dev_reserve(ctx, ctx->st_lab, cap));
  KTRY(dev_reserve(ctx, ctx->st_out_a, cap * (uint64_t)record_bytes + 8));
  KTRY(dev_reserve(ctx, ctx->st_out_b, cap * (uint64_t)record_bytes + 8));
  KTRY(dev_reserve(ctx, ctx->st_blk, sizeof(uint32_t) * 2 * (size_t)nblk_max + 16));
  KTRY(dev_reserve(ctx, ctx->st_counts, sizeof(unsigned long long) * 8));
  uint32_t* blk_a = ctx->st_blk.as<uint32_t>();
  uint32_t* blk_b = blk_a + nblk_max;
  unsigned long long* tot = ctx->st_counts.as<unsigned long long>() + 6;
  cudaStream_t st = ctx->stream;
  for (uint64_t base = 0; base < n_kmers; base += chunk) {
    const uint64_t c = std::min<uint64_t>(chunk, n_kmers - base);
    const uint32_t nblk = (uint32_t)((c + kSelTile - 1) / kSelTile);
    KCUDA(ctx, cudaMemcpyAsync(ctx->st_rec.p, records + base * (uint64_t)record_bytes, c * (uint64_t)record_bytes, cudaMemcpyHostToDevice, st));
    KCUDA(ctx, cudaMemcpyAsync(ctx->st_lab.p, id_label + base, c, cudaMemcpyHostToDevice, st));
    k_select_count<<<nblk, kSelTile, 0, st>>>(ctx->st_lab.as<uint8_t>(), c, blk_a, blk_b);
    SLAUNCH(ctx);
    k_select_scan<<<1, kSelTile, 0, st>>>(blk_a, blk_b, nblk, tot);
    SLAUNCH(ctx);
    k_select_write<<<nblk, kSelTile, 0, st>>>(ctx->st_rec.as<uint8_t>(), ctx->st_lab.as<uint8_t>(), c, record_bytes, blk_a, blk_b,
                                             ctx->st_out_a.as<uint8_t>(), ctx->st_out_b.as<uint8_t>());
    SLAUNCH(ctx);
    unsigned long long h[2];
    KCUDA(ctx, cudaMemcpyAsync(h, tot, sizeof h, cudaMemcpyDeviceToHost, st));
    KCUDA(ctx, cudaStreamSynchronize(st));
    if (h[0]) KCUDA(ctx, cudaMemcpyAsync(out_a + *n_a * (uint64_t)record_bytes, ctx->st_out_a.p, h[0] * (uint64_t)record_bytes, cudaMemcpyDeviceToHost, st));
    if (h[1]) KCUDA(ctx, cudaMemcpyAsync(out_b + *n_b * (uint64_t)record_bytes, ctx->st_out_b.p, h[1] * (uint64_t)record_bytes, cudaMemcpyDeviceToHost, st));
    KCUDA(ctx, cudaStreamSynchronize(st));
    *n_a += h[0];
    *n_b += h[1];
  }
  return KLSH_OK;
}
