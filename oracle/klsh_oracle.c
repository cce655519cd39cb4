/* TEST INFRASTRUCTURE — not product code.
 *
 * CPU restatement, in plain C, of the kmerLSH mode-C clustering hot path.  It exists
 * only so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can
 * check (never replace) the CUDA path in kmerlsh_b200/.  Nothing under kmerlsh_b200/
 * may include, link, import or execute this file.
 *
 * Every function cites the reference file:line (relative to the kmerLSH tree) whose
 * arithmetic it restates.  All floating point is IEEE binary32/binary64 with one
 * rounding per source-level operation and NO fused multiply-add — that is what the
 * reference's x86-64 -O3 build executes (SURVEY.md section 7 hard part 2) — so build with
 * -ffp-contract=off and without -ffast-math (oracle/Makefile does).
 *
 * Parity pinning: the reference ships no tests or golden vectors, so
 * this restatement is pinned against the reference ITSELF, compiled from its own
 * sources into oracle/_ref/ (oracle/Makefile, oracle/ref_harness.cc) and against the
 * fixtures minted from that build under tests/golden/ (tests/golden/make_golden.py).
 */
#define _POSIX_C_SOURCE 200809L
#include "klsh_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Random hyperplanes.
 * Reference: hash/lshash.cc:3-17 (generateNormalHashFunc): random_device rd; mt19937 gen(rd());
 * normal_distribution<> dis(0,1); function[i] = dis(gen) (double narrowed to float);
 * hash/lshash.cc:36-42 (generateHashTable): H such functions, in order.
 * Under oracle/seeded_rd.h, rd() yields (uint32)master() with master = mt19937_64(seed).
 * The three generators below restate the published algorithms the C++ standard fixes
 * (MT19937, MT19937-64) and libstdc++-13's generate_canonical / Marsaglia-polar
 * normal_distribution (bits/random.tcc), which is what the reference links.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t mt[624];
  int idx;
} mt32_t;

static void mt32_seed(mt32_t* g, uint32_t s) {
  g->mt[0] = s;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}

static uint32_t mt32_next(mt32_t* g) {
  if (g->idx >= 624) {
    for (int k = 0; k < 624; ++k) {
      uint32_t y = (g->mt[k] & 0x80000000u) | (g->mt[(k + 1) % 624] & 0x7fffffffu);
      g->mt[k] = g->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

typedef struct {
  uint64_t mt[312];
  int idx;
} mt64_t;

static void mt64_seed(mt64_t* g, uint64_t s) {
  g->mt[0] = s;
  for (int i = 1; i < 312; ++i)
    g->mt[i] = 6364136223846793005ULL * (g->mt[i - 1] ^ (g->mt[i - 1] >> 62)) + (uint64_t)i;
  g->idx = 312;
}

static uint64_t mt64_next(mt64_t* g) {
  if (g->idx >= 312) {
    for (int k = 0; k < 312; ++k) {
      uint64_t y = (g->mt[k] & 0xFFFFFFFF80000000ULL) | (g->mt[(k + 1) % 312] & 0x7FFFFFFFULL);
      g->mt[k] = g->mt[(k + 156) % 312] ^ (y >> 1) ^ ((y & 1ULL) ? 0xB5026F5AA96619E9ULL : 0ULL);
    }
    g->idx = 0;
  }
  uint64_t y = g->mt[g->idx++];
  y ^= (y >> 29) & 0x5555555555555555ULL;
  y ^= (y << 17) & 0x71D67FFFEDA60000ULL;
  y ^= (y << 37) & 0xFFF7EEE000000000ULL;
  y ^= (y >> 43);
  return y;
}

/* generate_canonical<double,53>(mt19937): two 32-bit draws, low word first. */
static double canonical53(mt32_t* g) {
  double sum = 0.0, tmp = 1.0;
  for (int k = 0; k < 2; ++k) {
    sum += (double)mt32_next(g) * tmp;
    tmp *= 4294967296.0;
  }
  double ret = sum / tmp;
  if (ret >= 1.0) ret = nextafter(1.0, 0.0);
  return ret;
}

struct klo_planes {
  mt64_t master;
  uint64_t draws;
};

klo_planes* klo_planes_new(uint64_t seed) {
  klo_planes* p = (klo_planes*)malloc(sizeof(klo_planes));
  klo_planes_reseed(p, seed);
  return p;
}
void klo_planes_free(klo_planes* p) { free(p); }
void klo_planes_reseed(klo_planes* p, uint64_t seed) {
  mt64_seed(&p->master, seed);
  p->draws = 0;
}
uint64_t klo_planes_draws(const klo_planes* p) { return p->draws; }

/* hash/lshash.cc:3-17: one hash function = fresh mt19937 + fresh normal_distribution. */
static void normal_hash_func(klo_planes* p, int D, float* out) {
  mt32_t gen;
  mt32_seed(&gen, (uint32_t)mt64_next(&p->master));
  p->draws++;
  int saved_available = 0;
  double saved = 0.0;
  for (int i = 0; i < D; ++i) {
    double ret;
    if (saved_available) {
      saved_available = 0;
      ret = saved;
    } else {
      double x, y, r2;
      do {
        x = 2.0 * canonical53(&gen) - 1.0;
        y = 2.0 * canonical53(&gen) - 1.0;
        r2 = x * x + y * y;
      } while (r2 > 1.0 || r2 == 0.0);
      double mult = sqrt(-2 * log(r2) / r2);
      saved = x * mult;
      saved_available = 1;
      ret = y * mult;
    }
    ret = ret * 1.0 + 0.0;
    out[i] = (float)ret;
  }
}

/* hash/lshash.cc:36-42 */
void klo_planes_table(klo_planes* p, int H, int D, float* out) {
  for (int h = 0; h < H; ++h) normal_hash_func(p, D, out + (size_t)h * D);
}

/* ------------------------------------------------------------------------------------------
 * Row transform.  io/ioMatrix.cc:372-392: value = float(log(cnt+1.0)) - v_kmers[j];
 * keep the row iff total_cnt > 0.1*tot_sample (uint64 vs double); ids = {batch_offset+i}.
 * app/kmerLSH.cc:477-481: v_kmers[j] = float(coverage_j) / kmap_size (float / size_t).
 * ---------------------------------------------------------------------------------------- */
void klo_log_lut(float* lut) {
  for (uint64_t c = 0; c < 65536; ++c) lut[c] = (float)log((double)c + 1.0);
}

void klo_vkmers(const float* coverage, uint64_t kmap_size, int D, float* out) {
  for (int j = 0; j < D; ++j) out[j] = coverage[j] / (float)kmap_size;
}

uint64_t klo_convert_counts(const uint16_t* counts, const float* v_kmers, int D, uint64_t batch_size,
                            uint64_t batch_offset, float* values_out, uint64_t* ids_out) {
  uint64_t kept = 0;
  for (uint64_t i = 0; i < batch_size; ++i) {
    uint64_t total = 0;
    float* dst = values_out + kept * (uint64_t)D;
    for (int j = 0; j < D; ++j) {
      uint64_t cnt = counts[(uint64_t)j * batch_size + i];
      total += cnt;
      dst[j] = (float)log((double)cnt + 1.0) - v_kmers[j];
    }
    if ((double)total > 0.1 * (double)D) {
      ids_out[kept] = batch_offset + i;
      ++kept;
    }
  }
  return kept;
}

/* ------------------------------------------------------------------------------------------
 * Scalar kernels.
 * ---------------------------------------------------------------------------------------- */
/* hash/lshash.cc:44-51 (one plane) and :53-59 (key = key*2 + bit; plane 0 = MSB). */
void klo_sign(const float* rows, uint64_t n, int D, const float* table, int H, uint32_t* keys) {
  for (uint64_t r = 0; r < n; ++r) {
    const float* v = rows + r * (uint64_t)D;
    uint32_t key = 0;
    for (int h = 0; h < H; ++h) {
      const float* f = table + (size_t)h * D;
      float sum = 0;
      for (int i = 0; i < D; ++i) sum += f[i] * v[i];
      key = key * 2u + (sum >= 0 ? 1u : 0u);
    }
    keys[r] = key;
  }
}

/* function/distance.cc:27-38: returns 1 - cos, all fp32, one pass, i ascending. */
float klo_cosine_distance(const float* lhs, const float* rhs, int D) {
  float similarity = 0;
  float magnitude_lhs = 0, magnitude_rhs = 0;
  for (int i = 0; i < D; ++i) {
    similarity += lhs[i] * rhs[i];
    magnitude_lhs += lhs[i] * lhs[i];
    magnitude_rhs += rhs[i] * rhs[i];
  }
  similarity /= sqrtf(magnitude_lhs) * sqrtf(magnitude_rhs);
  return 1 - similarity;
}

/* function/funcAB.cc:49-71: counts are int; v = cur*c1/all + cand*c2/all, left to right. */
void klo_consensus(const float* cur, int64_t cur_count, const float* cand, int64_t cand_count, int D,
                   float* out) {
  int c1 = (int)cur_count, c2 = (int)cand_count;
  int all = c1 + c2;
  for (int i = 0; i < D; ++i) out[i] = cur[i] * c1 / all + cand[i] * c2 / all;
}

/* function/cluster.cc:190-192, :330: fp32 recurrence threshold -= sim_step. */
float klo_threshold_after(float min_similarity, int iterations, int steps) {
  float max_similarity = 0.95;
  float sim_step = (max_similarity - min_similarity) / iterations;
  float threshold = max_similarity;
  for (int s = 0; s < steps; ++s) threshold -= sim_step;
  return threshold;
}

/* ------------------------------------------------------------------------------------------
 * Row sets: SoA restatement of vector<Abundance*> (common/abundance.h:18-54).
 * _values -> values[r*D..]; _ids -> a singly linked chain of member slots
 * head[r] -> next[] ... -> tail[r], payload member[slot]; count[r] = |_ids|.
 * Concatenation ids(cur) ++ ids(cand) (function/funcAB.cc:55) is next[tail(cur)] = head(cand).
 * ---------------------------------------------------------------------------------------- */
struct klo_rows {
  uint64_t n;
  int D;
  float* values;
  int64_t* count;
  int64_t* head;
  int64_t* tail;
  uint64_t m; /* member slots */
  int64_t* next;
  uint64_t* member;
};

klo_rows* klo_rows_new(const float* values, const uint64_t* id_offsets, const uint64_t* ids, uint64_t n,
                       int D) {
  klo_rows* r = (klo_rows*)calloc(1, sizeof(klo_rows));
  r->n = n;
  r->D = D;
  r->m = n ? id_offsets[n] : 0;
  r->values = (float*)malloc(sizeof(float) * (n * (uint64_t)D + 1));
  r->count = (int64_t*)malloc(sizeof(int64_t) * (n + 1));
  r->head = (int64_t*)malloc(sizeof(int64_t) * (n + 1));
  r->tail = (int64_t*)malloc(sizeof(int64_t) * (n + 1));
  r->next = (int64_t*)malloc(sizeof(int64_t) * (r->m + 1));
  r->member = (uint64_t*)malloc(sizeof(uint64_t) * (r->m + 1));
  if (n) memcpy(r->values, values, sizeof(float) * n * (uint64_t)D);
  if (r->m) memcpy(r->member, ids, sizeof(uint64_t) * r->m);
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t b = id_offsets[i], e = id_offsets[i + 1];
    r->count[i] = (int64_t)(e - b);
    r->head[i] = (e > b) ? (int64_t)b : -1;
    r->tail[i] = (e > b) ? (int64_t)(e - 1) : -1;
    for (uint64_t s = b; s < e; ++s) r->next[s] = (s + 1 < e) ? (int64_t)(s + 1) : -1;
  }
  return r;
}

void klo_rows_free(klo_rows* r) {
  if (!r) return;
  free(r->values);
  free(r->count);
  free(r->head);
  free(r->tail);
  free(r->next);
  free(r->member);
  free(r);
}

uint64_t klo_rows_count(const klo_rows* r) { return r->n; }
uint64_t klo_rows_members(const klo_rows* r) {
  uint64_t t = 0;
  for (uint64_t i = 0; i < r->n; ++i) t += (uint64_t)r->count[i];
  return t;
}
int klo_rows_dim(const klo_rows* r) { return r->D; }

void klo_rows_export(const klo_rows* r, float* values, uint64_t* id_offsets, uint64_t* ids) {
  uint64_t off = 0;
  if (r->n) memcpy(values, r->values, sizeof(float) * r->n * (uint64_t)r->D);
  for (uint64_t i = 0; i < r->n; ++i) {
    id_offsets[i] = off;
    for (int64_t s = r->head[i]; s >= 0; s = r->next[s]) ids[off++] = r->member[s];
  }
  id_offsets[r->n] = off;
}

void klo_rows_append(klo_rows* dst, klo_rows* src) {
  uint64_t n = dst->n + src->n, m = dst->m + src->m;
  int D = dst->D;
  dst->values = (float*)realloc(dst->values, sizeof(float) * (n * (uint64_t)D + 1));
  dst->count = (int64_t*)realloc(dst->count, sizeof(int64_t) * (n + 1));
  dst->head = (int64_t*)realloc(dst->head, sizeof(int64_t) * (n + 1));
  dst->tail = (int64_t*)realloc(dst->tail, sizeof(int64_t) * (n + 1));
  dst->next = (int64_t*)realloc(dst->next, sizeof(int64_t) * (m + 1));
  dst->member = (uint64_t*)realloc(dst->member, sizeof(uint64_t) * (m + 1));
  if (src->n) memcpy(dst->values + dst->n * (uint64_t)D, src->values, sizeof(float) * src->n * (uint64_t)D);
  for (uint64_t i = 0; i < src->n; ++i) {
    dst->count[dst->n + i] = src->count[i];
    dst->head[dst->n + i] = src->head[i] < 0 ? -1 : src->head[i] + (int64_t)dst->m;
    dst->tail[dst->n + i] = src->tail[i] < 0 ? -1 : src->tail[i] + (int64_t)dst->m;
  }
  for (uint64_t s = 0; s < src->m; ++s) {
    dst->next[dst->m + s] = src->next[s] < 0 ? -1 : src->next[s] + (int64_t)dst->m;
    dst->member[dst->m + s] = src->member[s];
  }
  dst->n = n;
  dst->m = m;
  src->n = 0;
  src->m = 0;
}

/* ------------------------------------------------------------------------------------------
 * Clustering.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint64_t compares, merges, buckets_nonempty, bucket_max, nested_calls;
} pass_stats;

/* function/cluster.cc:56-87 on the bucket c[0..size): greedy first-match merge with
 * swap-remove.  `c` holds row indices.  Merge of current=c[i] into candidate=c[j]:
 * consensus(current, candidate) replaces slot j (:70-74), slot i takes the tail (:75).
 * Returns the new size; survivors are c[0..size) in that order (:84). */
static uint64_t p_cluster_idx(klo_rows* r, uint64_t* c, uint64_t size, float threshold, pass_stats* st) {
  const int D = r->D;
  uint64_t i = 1, j = 0;
  while (i < size) {
    const uint64_t cur = c[i];
    const float* curv = r->values + cur * (uint64_t)D;
    for (j = 0; j < i; ++j) {
      const uint64_t cand = c[j];
      float* candv = r->values + cand * (uint64_t)D;
      float distance = klo_cosine_distance(curv, candv, D);
      if (st) st->compares++;
      if (1 - distance >= threshold) {
        /* the new Abundance lives in cand's storage; cur's storage dies */
        klo_consensus(curv, r->count[cur], candv, r->count[cand], D, candv);
        if (r->tail[cur] >= 0) { /* ids(cur) ++ ids(cand) */
          r->next[r->tail[cur]] = r->head[cand];
          r->head[cand] = r->head[cur];
          if (r->tail[cand] < 0) r->tail[cand] = r->tail[cur];
        }
        r->count[cand] += r->count[cur];
        c[i] = c[--size];
        if (st) st->merges++;
        break;
      }
    }
    if (j == i) ++i;
  }
  return size;
}

/* stable LSD radix sort of (key, idx) pairs by the low `bits` bits of key.
 * Equivalent to merge_hashtable's push_back into a dense 2^H table in row order
 * (function/cluster.cc:15-30) followed by visiting buckets in ascending key (:281-293). */
static void sort_pairs(uint32_t* keys, uint64_t* idx, uint64_t n, int bits) {
  if (n < 2 || bits <= 0) return;
  uint32_t* k2 = (uint32_t*)malloc(sizeof(uint32_t) * n);
  uint64_t* i2 = (uint64_t*)malloc(sizeof(uint64_t) * n);
  uint64_t* hist = (uint64_t*)malloc(sizeof(uint64_t) * 2048);
  for (int shift = 0; shift < bits; shift += 11) {
    memset(hist, 0, sizeof(uint64_t) * 2048);
    for (uint64_t t = 0; t < n; ++t) hist[(keys[t] >> shift) & 2047u]++;
    uint64_t acc = 0;
    for (int b = 0; b < 2048; ++b) {
      uint64_t h = hist[b];
      hist[b] = acc;
      acc += h;
    }
    for (uint64_t t = 0; t < n; ++t) {
      uint64_t p = hist[(keys[t] >> shift) & 2047u]++;
      k2[p] = keys[t];
      i2[p] = idx[t];
    }
    uint32_t* tk = keys; keys = k2; k2 = tk;
    uint64_t* ti = idx; idx = i2; i2 = ti;
  }
  /* odd number of passes: result sits in the scratch buffers */
  int passes = (bits + 10) / 11;
  if (passes & 1) {
    memcpy(k2, keys, sizeof(uint32_t) * n);
    memcpy(i2, idx, sizeof(uint64_t) * n);
    uint32_t* tk = keys; keys = k2; k2 = tk;
    uint64_t* ti = idx; idx = i2; i2 = ti;
  }
  free(k2);
  free(i2);
  free(hist);
}

static int floor_log2_ref(uint64_t n) { /* function/cluster.cc:194, :203: floor(log2(size)) */
  return (int)floor(log2((double)n));
}

/* One signing + grouping + per-bucket merge over the rows listed in idx[0..n).
 * nest_threshold < 0 disables nesting (that is nestedCluster's own inner pass,
 * function/cluster.cc:153-159, which always calls p_cluster).  Survivor indices are
 * written back to idx in canonical order; returns their number. */
static uint64_t cluster_pass(klo_rows* r, uint64_t* idx, uint64_t n, int H, const float* table,
                             float threshold, int64_t nest_threshold, klo_planes* planes, pass_stats* st) {
  const int D = r->D;
  uint32_t* keys = (uint32_t*)malloc(sizeof(uint32_t) * (n + 1));
  for (uint64_t t = 0; t < n; ++t) klo_sign(r->values + idx[t] * (uint64_t)D, 1, D, table, H, &keys[t]);
  sort_pairs(keys, idx, n, H);
  uint64_t out = 0, b = 0;
  while (b < n) {
    uint64_t e = b + 1;
    while (e < n && keys[e] == keys[b]) ++e;
    uint64_t size = e - b;
    if (st) {
      st->buckets_nonempty++;
      if (size > st->bucket_max) st->bucket_max = size;
    }
    uint64_t kept;
    if (nest_threshold >= 0 && size > (uint64_t)nest_threshold) {
      /* function/cluster.cc:286-288 -> nestedCluster :89-178 */
      int H2 = floor_log2_ref(size);
      float* t2 = (float*)malloc(sizeof(float) * ((size_t)H2 * D + 1));
      klo_planes_table(planes, H2, D, t2);
      if (st) st->nested_calls++;
      kept = cluster_pass(r, idx + b, size, H2, t2, threshold, -1, planes, st);
      free(t2);
    } else {
      kept = p_cluster_idx(r, idx + b, size, threshold, st);
    }
    memmove(idx + out, idx + b, sizeof(uint64_t) * kept);
    out += kept;
    b = e;
  }
  free(keys);
  return out;
}

/* gather survivors (row indices idx[0..k)) to the front, in order */
static void compact_rows(klo_rows* r, const uint64_t* idx, uint64_t k) {
  const int D = r->D;
  float* v = (float*)malloc(sizeof(float) * (k * (uint64_t)D + 1));
  int64_t* cnt = (int64_t*)malloc(sizeof(int64_t) * (k + 1));
  int64_t* hd = (int64_t*)malloc(sizeof(int64_t) * (k + 1));
  int64_t* tl = (int64_t*)malloc(sizeof(int64_t) * (k + 1));
  for (uint64_t t = 0; t < k; ++t) {
    memcpy(v + t * (uint64_t)D, r->values + idx[t] * (uint64_t)D, sizeof(float) * D);
    cnt[t] = r->count[idx[t]];
    hd[t] = r->head[idx[t]];
    tl[t] = r->tail[idx[t]];
  }
  free(r->values); free(r->count); free(r->head); free(r->tail);
  r->values = v; r->count = cnt; r->head = hd; r->tail = tl;
  r->n = k;
}

void klo_p_cluster(klo_rows* r, float threshold) {
  uint64_t n = r->n;
  uint64_t* idx = (uint64_t*)malloc(sizeof(uint64_t) * (n + 1));
  for (uint64_t t = 0; t < n; ++t) idx[t] = t;
  uint64_t k = p_cluster_idx(r, idx, n, threshold, NULL);
  compact_rows(r, idx, k);
  free(idx);
}

void klo_nested_cluster(klo_rows* r, float threshold, klo_planes* planes) {
  uint64_t n = r->n;
  if (n == 0) return;
  uint64_t* idx = (uint64_t*)malloc(sizeof(uint64_t) * (n + 1));
  for (uint64_t t = 0; t < n; ++t) idx[t] = t;
  int H = floor_log2_ref(n);
  float* table = (float*)malloc(sizeof(float) * ((size_t)H * r->D + 1));
  klo_planes_table(planes, H, r->D, table);
  uint64_t k = cluster_pass(r, idx, n, H, table, threshold, -1, planes, NULL);
  compact_rows(r, idx, k);
  free(table);
  free(idx);
}

/* function/cluster.cc:181-340 */
void klo_cluster(klo_rows* r, float min_similarity, int iterations, int64_t bucket_size_threshold,
                 klo_planes* planes, klo_iter_stats* stats) {
  float max_similarity = 0.95;
  float sim_step = (max_similarity - min_similarity) / iterations;
  float threshold = max_similarity;
  int iter = 0;
  while (iter++ < iterations) {
    uint64_t n = r->n;
    if (n == 0) break; /* reference: log2(0) -> undefined; we stop */
    int H = floor_log2_ref(n);
    float* table = (float*)malloc(sizeof(float) * ((size_t)H * r->D + 1));
    klo_planes_table(planes, H, r->D, table);
    uint64_t* idx = (uint64_t*)malloc(sizeof(uint64_t) * (n + 1));
    for (uint64_t t = 0; t < n; ++t) idx[t] = t;
    pass_stats st;
    memset(&st, 0, sizeof(st));
    uint64_t k = cluster_pass(r, idx, n, H, table, threshold, bucket_size_threshold, planes, &st);
    compact_rows(r, idx, k);
    if (stats) {
      klo_iter_stats* s = &stats[iter - 1];
      s->rows_in = n;
      s->rows_out = k;
      s->H = H;
      s->threshold = threshold;
      s->buckets_nonempty = st.buckets_nonempty;
      s->bucket_max = st.bucket_max;
      s->nested_calls = st.nested_calls;
      s->compares = st.compares;
      s->merges = st.merges;
    }
    free(idx);
    free(table);
    threshold -= sim_step;
  }
}

uint64_t klo_bucket_sizes(const klo_rows* r, const float* table, int H, uint64_t* sizes_out, uint64_t cap) {
  uint64_t n = r->n;
  uint32_t* keys = (uint32_t*)malloc(sizeof(uint32_t) * (n + 1));
  uint64_t* idx = (uint64_t*)malloc(sizeof(uint64_t) * (n + 1));
  klo_sign(r->values, n, r->D, table, H, keys);
  for (uint64_t t = 0; t < n; ++t) idx[t] = t;
  sort_pairs(keys, idx, n, H);
  uint64_t nb = 0, b = 0;
  while (b < n) {
    uint64_t e = b + 1;
    while (e < n && keys[e] == keys[b]) ++e;
    if (nb < cap) sizes_out[nb] = e - b;
    ++nb;
    b = e;
  }
  free(keys);
  free(idx);
  return nb;
}

/* ------------------------------------------------------------------------------------------
 * Files.
 * io/ioMatrix.cc:265-294 SaveResult: "<size>\t<id>\t<id>...\n" for rows with |ids| > ignore_small
 * io/ioMatrix.cc:322-351 SaveBinary: D raw float32 per such row, same order, no header
 * delfile -> remove() first; both open in append mode.
 * ---------------------------------------------------------------------------------------- */
int klo_save(const klo_rows* r, const char* bin_path, int delfile, int64_t ignore_small) {
  size_t L = strlen(bin_path);
  char* clust = (char*)malloc(L + 8);
  memcpy(clust, bin_path, L);
  memcpy(clust + L, ".clust", 7);
  if (delfile) {
    remove(clust);
    remove(bin_path);
  }
  FILE* ft = fopen(clust, "a");
  FILE* fb = fopen(bin_path, "ab");
  free(clust);
  if (!ft || !fb) {
    if (ft) fclose(ft);
    if (fb) fclose(fb);
    return -1;
  }
  for (uint64_t i = 0; i < r->n; ++i) {
    if (r->count[i] > ignore_small) {
      fprintf(ft, "%llu", (unsigned long long)r->count[i]);
      for (int64_t s = r->head[i]; s >= 0; s = r->next[s]) fprintf(ft, "\t%llu", (unsigned long long)r->member[s]);
      fputc('\n', ft);
      fwrite(r->values + i * (uint64_t)r->D, sizeof(float), (size_t)r->D, fb);
    }
  }
  fclose(ft);
  fclose(fb);
  return 0;
}

/* io/ioMatrix.cc:121-196 ReadCluster (start_line/num_lines) and :48-119 ReadClusterAll
 * (num_lines == 0): row k of the float matrix <-> line k of "<path>.clust"; the first token
 * of a line is the member count, the rest are ids (strtol). */
klo_rows* klo_read_cluster(const char* bin_path, int D, uint64_t start_line, uint64_t num_lines) {
  FILE* fb = fopen(bin_path, "rb");
  if (!fb) return NULL;
  fseek(fb, 0, SEEK_END);
  uint64_t total = (uint64_t)ftell(fb) / (sizeof(float) * (uint64_t)D);
  if (num_lines == 0) {
    start_line = 0;
    num_lines = total;
  }
  if (start_line + num_lines > total) num_lines = total > start_line ? total - start_line : 0;
  float* values = (float*)malloc(sizeof(float) * (num_lines * (uint64_t)D + 1));
  fseek(fb, (long)(start_line * sizeof(float) * (uint64_t)D), SEEK_SET);
  if (fread(values, sizeof(float) * (size_t)D, num_lines, fb) != num_lines) {
    fclose(fb);
    free(values);
    return NULL;
  }
  fclose(fb);

  size_t L = strlen(bin_path);
  char* clust = (char*)malloc(L + 8);
  memcpy(clust, bin_path, L);
  memcpy(clust + L, ".clust", 7);
  FILE* ft = fopen(clust, "r");
  free(clust);
  if (!ft) {
    free(values);
    return NULL;
  }
  uint64_t* offs = (uint64_t*)malloc(sizeof(uint64_t) * (num_lines + 1));
  uint64_t cap = num_lines * 2 + 16, m = 0;
  uint64_t* ids = (uint64_t*)malloc(sizeof(uint64_t) * cap);
  char* line = NULL;
  size_t linecap = 0;
  uint64_t lineno = 0, loc = 0;
  while (loc < num_lines && getline(&line, &linecap, ft) >= 0) {
    if (lineno++ < start_line) continue;
    char* end;
    const char* p = line;
    uint64_t cnt = (uint64_t)strtol(p, &end, 10);
    offs[loc] = m;
    if (m + cnt + 1 > cap) {
      cap = (m + cnt) * 2 + 16;
      ids = (uint64_t*)realloc(ids, sizeof(uint64_t) * cap);
    }
    for (uint64_t t = 0; t < cnt && p != end; ++t) {
      p = end;
      ids[m++] = (uint64_t)strtol(p, &end, 10);
    }
    ++loc;
  }
  free(line);
  fclose(ft);
  offs[loc] = m;
  klo_rows* r = klo_rows_new(values, offs, ids, loc, D);
  free(values);
  free(offs);
  free(ids);
  return r;
}

/* ------------------------------------------------------------------------------------------
 * Mode C (app/kmerLSH.cc:469-499 and init_clustering :278-430).
 * ---------------------------------------------------------------------------------------- */
static char* path_join(const char* dir, uint64_t k) {
  char* s = (char*)malloc(strlen(dir) + 40);
  sprintf(s, "%s%llu.bin", dir, (unsigned long long)k);
  return s;
}

int klo_mode_c(const char* count_bin, const char* count_log, int D, float min_similarity, int iterations,
               const char* tmp_dir, const char* out_path, uint64_t batch_thresh,
               int64_t phase2_bucket_threshold, uint64_t seed, klo_iter_stats* phase2_stats) {
  /* app/kmerLSH.cc:473-481 */
  FILE* fl = fopen(count_log, "r");
  if (!fl) return -1;
  unsigned long long kmap_size_ll = 0;
  if (fscanf(fl, "%llu", &kmap_size_ll) != 1) {
    fclose(fl);
    return -2;
  }
  uint64_t kmap_size = kmap_size_ll;
  float* cov = (float*)malloc(sizeof(float) * D);
  float* vk = (float*)malloc(sizeof(float) * D);
  for (int j = 0; j < D; ++j)
    if (fscanf(fl, "%f", &cov[j]) != 1) cov[j] = 0.f;
  fclose(fl);
  klo_vkmers(cov, kmap_size, D, vk);

  klo_planes* planes = klo_planes_new(seed);
  FILE* fb = fopen(count_bin, "rb");
  if (!fb) return -3;

  /* init_clustering first loop (:311-345): independent batches, Cluster(I=1, thr batch/1000) */
  uint64_t batch_offset = 0, total_size = 0, tmp = 0;
  uint64_t nbatch = kmap_size / batch_thresh;
  char* write_tmp = path_join(tmp_dir, tmp++);
  uint16_t* counts = (uint16_t*)malloc(sizeof(uint16_t) * ((size_t)D * batch_thresh + 1));
  float* vals = (float*)malloc(sizeof(float) * ((size_t)D * batch_thresh + 1));
  uint64_t* ids = (uint64_t*)malloc(sizeof(uint64_t) * (batch_thresh + 1));
  uint64_t* offs = (uint64_t*)malloc(sizeof(uint64_t) * (batch_thresh + 2));
  for (uint64_t i = 0; i <= nbatch; ++i) {
    uint64_t batch_size = (i == nbatch) ? kmap_size - batch_offset : batch_thresh;
    /* io/ioHT.cc:59-81 ReadHT: per sample seek (j*num_kmer + batch_offset)*2 */
    for (int j = 0; j < D; ++j) {
      fseek(fb, (long)(((uint64_t)j * kmap_size + batch_offset) * sizeof(uint16_t)), SEEK_SET);
      if (fread(counts + (size_t)j * batch_size, sizeof(uint16_t), batch_size, fb) != batch_size) return -4;
    }
    uint64_t kept = klo_convert_counts(counts, vk, D, batch_size, batch_offset, vals, ids);
    for (uint64_t t = 0; t <= kept; ++t) offs[t] = t;
    klo_rows* rows = klo_rows_new(vals, offs, ids, kept, D);
    klo_cluster(rows, min_similarity, 1, (int64_t)(batch_thresh / 1000), planes, NULL);
    total_size += rows->n;
    klo_save(rows, write_tmp, i == 0, 0);
    klo_rows_free(rows);
    batch_offset += batch_size;
  }
  free(counts); free(vals); free(ids); free(offs);
  fclose(fb);

  /* second loop (:354-411) */
  float similarity = min_similarity;
  while (total_size > batch_thresh) {
    similarity -= 0.001;
    batch_offset = 0;
    char* read_tmp = write_tmp;
    write_tmp = path_join(tmp_dir, tmp++);
    nbatch = total_size / batch_thresh;
    uint64_t rem = total_size;
    total_size = 0;
    for (uint64_t i = 0; i <= nbatch; ++i) {
      uint64_t batch_size = (i == nbatch) ? rem - batch_offset : batch_thresh;
      klo_rows* rows = klo_read_cluster(read_tmp, D, batch_offset, batch_size);
      if (!rows) return -5;
      if (batch_size == 0) { /* ReadCluster with 0 lines -> empty set; Cluster on it is UB upstream */
        klo_save(rows, write_tmp, i == 0, 0);
        klo_rows_free(rows);
        continue;
      }
      klo_cluster(rows, similarity, 1 + 4, (int64_t)(batch_thresh / 1000), planes, NULL); /* cluster_iteration+4, :377 */
      total_size += rows->n;
      klo_save(rows, write_tmp, i == 0, 0);
      klo_rows_free(rows);
      batch_offset += batch_size;
    }
    size_t L = strlen(read_tmp);
    char* rc = (char*)malloc(L + 8);
    memcpy(rc, read_tmp, L);
    memcpy(rc + L, ".clust", 7);
    remove(read_tmp);
    remove(rc);
    free(rc);
    free(read_tmp);
  }

  klo_rows* all = klo_read_cluster(write_tmp, D, 0, 0);
  free(write_tmp);
  if (!all) return -6;
  /* app/kmerLSH.cc:490: Cluster(all, N, I, T, dim, bucket_size_threshold = 1000000 (:440)) */
  klo_cluster(all, min_similarity, iterations, phase2_bucket_threshold, planes, phase2_stats);
  /* :498-499 */
  int rc = klo_save(all, out_path, 1, 5);
  klo_rows_free(all);
  klo_planes_free(planes);
  free(cov);
  free(vk);
  return rc;
}


/* ------------------------------------------------------------------------------------------
 * Mode E statistics.
 * Reference: AB::WRS (function/funcAB.cc:73-109) calls alglib::studentttest2 on the two halves of a
 * centroid (values widened to double) and files the cluster's ids under group B when
 * lefttail <= pvalue_thresh, else under group A when righttail <= pvalue_thresh; the caller
 * (app/kmerLSH.cc:541-585) then walks kmer_set.hex and keeps the k-mers whose id is in either set.
 *
 * ALGLIB 3.15.0 is vendored third-party code (utils/alglib-3.15.0, Cephes-derived).  The test
 * statistic below restates statistics.cpp:12502-12616 operation by operation and is bit-identical to
 * it; studenttdistribution's finite series for t >= -2 restates specialfunctions.cpp:9559-9631.  For
 * t < -2 ALGLIB evaluates 0.5*incompletebeta(k/2, 1/2, k/(k+t*t)) with Cephes' incbet (power series /
 * two continued fractions / gamma function, specialfunctions.cpp:6975-7088, :7584-7900); that code is
 * NOT restated: the same regularised incomplete beta function is evaluated here with the modified
 * Lentz continued fraction.  Parity of this part is therefore to a tolerance (relative 1e-9 on the tail
 * probabilities, pinned against ALGLIB itself through oracle/_ref/libklsh_ref.so and the fixtures in
 * tests/golden/ttest.npz), not bit for bit.
 * ---------------------------------------------------------------------------------------- */
static double klo_betacf(double a, double b, double x) {
  const double tiny = 1e-300;
  double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 2000; ++m) {
    double m2 = 2.0 * m;
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
    double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < 2e-16) break;
  }
  return h;
}

/* regularised incomplete beta I_x(a, b), 0 <= x <= 1 */
static double klo_incbeta(double a, double b, double x) {
  if (x <= 0.0) return 0.0;
  if (x >= 1.0) return 1.0;
  double lbt = lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log1p(-x);
  double bt = exp(lbt);
  if (x < (a + 1.0) / (a + b + 2.0)) return bt * klo_betacf(a, b, x) / a;
  return 1.0 - bt * klo_betacf(b, a, 1.0 - x) / b;
}

/* specialfunctions.cpp:9559-9631 */
static double klo_student_t_cdf(int k, double t) {
  if (t == 0.0) return 0.5;
  if (t < -2.0) {
    double rk = (double)k;
    double z = rk / (rk + t * t);
    return 0.5 * klo_incbeta(0.5 * rk, 0.5, z);
  }
  double x = t < 0.0 ? -t : t;
  double rk = (double)k;
  double z = 1.0 + x * x / rk;
  double p, f, tz;
  int j;
  if (k % 2 != 0) {
    double xsqk = x / sqrt(rk);
    p = atan(xsqk);
    if (k > 1) {
      f = 1.0;
      tz = 1.0;
      j = 3;
      while (j <= k - 2 && tz / f > 5E-16) { /* ae_machineepsilon */
        tz = tz * ((j - 1) / (z * j));
        f = f + tz;
        j = j + 2;
      }
      p = p + f * xsqk / z;
    }
    p = p * 2.0 / 3.14159265358979323846; /* ae_pi */
  } else {
    f = 1.0;
    tz = 1.0;
    j = 2;
    while (j <= k - 2 && tz / f > 5E-16) {
      tz = tz * ((j - 1) / (z * j));
      f = f + tz;
      j = j + 2;
    }
    p = f * x / sqrt(z * rk);
  }
  if (t < 0.0) p = -p;
  return 0.5 + 0.5 * p;
}

/* statistics.cpp:12502-12616 */
void klo_ttest2(const double* x, int n, const double* y, int m, double* bothtails, double* lefttail,
                double* righttail) {
  if (n <= 0 || m <= 0) {
    *bothtails = 1.0;
    *lefttail = 1.0;
    *righttail = 1.0;
    return;
  }
  double xmean = 0.0, x0 = x[0];
  int samex = 1;
  for (int i = 0; i < n; ++i) {
    xmean = xmean + x[i];
    samex = samex && (x[i] == x0);
  }
  xmean = samex ? x0 : xmean / n;
  double ymean = 0.0, y0 = y[0];
  int samey = 1;
  for (int i = 0; i < m; ++i) {
    ymean = ymean + y[i];
    samey = samey && (y[i] == y0);
  }
  ymean = samey ? y0 : ymean / m;
  double s = 0.0;
  if (n + m > 2) {
    for (int i = 0; i < n; ++i) s = s + (x[i] - xmean) * (x[i] - xmean);
    for (int i = 0; i < m; ++i) s = s + (y[i] - ymean) * (y[i] - ymean);
    s = sqrt(s * ((double)1 / (double)n + (double)1 / (double)m) / (n + m - 2));
  }
  if (s == 0.0) {
    *bothtails = xmean == ymean ? 1.0 : 0.0;
    *lefttail = xmean >= ymean ? 1.0 : 0.0;
    *righttail = xmean <= ymean ? 1.0 : 0.0;
    return;
  }
  double stat = (xmean - ymean) / s;
  double p = klo_student_t_cdf(n + m - 2, stat);
  *bothtails = 2 * (p < 1 - p ? p : 1 - p); /* ae_minreal */
  *lefttail = p;
  *righttail = 1 - p;
}

/* function/funcAB.cc:73-109, one call per row as in app/kmerLSH.cc:543-545 */
void klo_wrs_rows(const float* values, const uint64_t* id_offsets, uint64_t n, int D, int num_sample1,
                  int num_sample2, float pvalue_thresh, int size_thresh, uint8_t* group, double* lefttail,
                  double* righttail) {
  int n1 = num_sample1 > 0 ? num_sample1 : 0, n2 = num_sample2 > 0 ? num_sample2 : 0;
  double* a = (double*)malloc(sizeof(double) * (size_t)(n1 + 1));
  double* b = (double*)malloc(sizeof(double) * (size_t)(n2 + 1));
  for (uint64_t r = 0; r < n; ++r) {
    group[r] = 0;
    if (lefttail) lefttail[r] = -1.0;
    if (righttail) righttail[r] = -1.0;
    uint64_t members = id_offsets[r + 1] - id_offsets[r];
    /* `ids.size() > size_thresh`: size_t against int, the int is converted to size_t (:87) */
    if (!(members > (uint64_t)(int64_t)size_thresh)) continue;
    const float* v = values + r * (uint64_t)D;
    for (int i = 0; i < n1; ++i) a[i] = (double)v[i];
    for (int j = 0; j < n2; ++j) b[j] = (double)v[n1 + j];
    double both, left, right;
    klo_ttest2(a, num_sample1, b, num_sample2, &both, &left, &right);
    if (lefttail) lefttail[r] = left;
    if (righttail) righttail[r] = right;
    if (left <= pvalue_thresh)
      group[r] = 2;
    else if (right <= pvalue_thresh)
      group[r] = 1;
  }
  free(a);
  free(b);
}

/* app/kmerLSH.cc:543-545 (the id sets) with the precedence of :571-576 */
void klo_differential_ids(const uint8_t* group, const uint64_t* id_offsets, const uint64_t* ids, uint64_t n,
                          uint64_t n_kmers, uint8_t* id_label) {
  memset(id_label, 0, n_kmers);
  for (int pass = 2; pass >= 1; --pass) /* set 1 wins when an id sits in both */
    for (uint64_t r = 0; r < n; ++r)
      if (group[r] == pass)
        for (uint64_t k = id_offsets[r]; k < id_offsets[r + 1]; ++k)
          if (ids[k] < n_kmers) id_label[ids[k]] = (uint8_t)pass;
}

/* app/kmerLSH.cc:565-579 */
void klo_select_kmers(const uint8_t* records, uint64_t n_kmers, int record_bytes, const uint8_t* id_label,
                      uint8_t* out_a, uint64_t* n_a, uint8_t* out_b, uint64_t* n_b) {
  uint64_t ca = 0, cb = 0;
  for (uint64_t i = 0; i < n_kmers; ++i) {
    if (id_label[i] == 1) {
      memcpy(out_a + ca * (uint64_t)record_bytes, records + i * (uint64_t)record_bytes, (size_t)record_bytes);
      ++ca;
    } else if (id_label[i] == 2) {
      memcpy(out_b + cb * (uint64_t)record_bytes, records + i * (uint64_t)record_bytes, (size_t)record_bytes);
      ++cb;
    }
  }
  *n_a = ca;
  *n_b = cb;
}


/* ------------------------------------------------------------------------------------------
 * Read extraction votes.
 * Reference: IOFQ::CheckRead (io/ioFastQ.cc:5-75) slides a Kmer over every read of at least k+10 bases
 * (Kmer(const char*) kmer/Kmer.cc:131-150, forwardBase :213-230), takes the canonical form
 * `rep = (km < tw) ? km : tw` with tw = km.twin() (:160-185) and operator< = memcmp over the MAX_K/4 = 8
 * bytes (:98-100), counts the reps found in the unordered_set g_kmer, and marks the read when
 * kmer_count / (len - k + 1) > kmer_vote in float.  Byte layout as the reference: base i sits in byte i/4 at
 * bit 2*(i%4); characters other than A, C, G, T leave the bits at 0 (= A) in both set_kmer and forwardBase.
 * ---------------------------------------------------------------------------------------- */
enum { KLO_KB = 8 }; /* Kmer::MAX_K / 4 with MAX_KMER_SIZE 32 (kmer/Kmer.h:4-5, :68) */

static unsigned klo_base_code(char ch) {
  switch (ch) {
    case 'C': return 1u;
    case 'G': return 2u;
    case 'T': return 3u;
    default: return 0u; /* 'A' and everything else */
  }
}

void klo_kmer_from_string(const char* s, int k, uint8_t* out) { /* kmer/Kmer.cc:131-150 */
  memset(out, 0, KLO_KB);
  for (int i = 0; i < k; ++i) out[i / 4] |= (uint8_t)(klo_base_code(s[i]) << (2 * (i % 4)));
}

static void klo_kmer_forward(uint8_t* km, int k, char b) { /* kmer/Kmer.cc:213-230 */
  const int k_bytes = (k + 3) / 4;
  const unsigned k_modmask = (1u << (2 * ((k % 4) ? k % 4 : 4))) - 1u;
  for (int i = 0; i < k_bytes - 1; ++i) { /* shiftRight(2), :330-341 */
    km[i] >>= 2;
    km[i] |= (uint8_t)(km[i + 1] << 6);
  }
  km[k_bytes - 1] >>= 2;
  km[k_bytes - 1] &= (uint8_t)k_modmask;
  km[k_bytes - 1] |= (uint8_t)(klo_base_code(b) << (2 * ((k + 3) % 4)));
}

static void klo_kmer_twin(const uint8_t* km, int k, uint8_t* tw) {
  /* kmer/Kmer.cc:160-185 computes, by complementing, shifting and swapping bytes, the reverse complement in the
   * same layout: base i of the twin is 3 - base (k-1-i); bytes from k_bytes on are copied (zero). */
  const int k_bytes = (k + 3) / 4;
  memcpy(tw, km, KLO_KB);
  memset(tw, 0, (size_t)k_bytes);
  for (int i = 0; i < k; ++i) {
    const int j = k - 1 - i;
    const unsigned b = 3u - ((km[j / 4] >> (2 * (j % 4))) & 3u);
    tw[i / 4] |= (uint8_t)(b << (2 * (i % 4)));
  }
}

void klo_kmer_rep(const uint8_t* km, int k, uint8_t* out) {
  uint8_t tw[KLO_KB];
  klo_kmer_twin(km, k, tw);
  memcpy(out, memcmp(km, tw, KLO_KB) < 0 ? km : tw, KLO_KB);
}

static int klo_rec_cmp(const void* a, const void* b) { return memcmp(a, b, KLO_KB); }

void klo_check_reads(const uint8_t* kmers, uint64_t n_kmers, int k, const char* seq, const uint64_t* seq_offsets,
                     uint64_t n_reads, float kmer_vote, uint8_t* record, uint32_t* votes) {
  uint8_t* set = (uint8_t*)malloc((size_t)(n_kmers ? n_kmers : 1) * KLO_KB);
  memcpy(set, kmers, (size_t)n_kmers * KLO_KB);
  qsort(set, (size_t)n_kmers, KLO_KB, klo_rec_cmp); /* membership only: the reference's set is unordered */
  for (uint64_t r = 0; r < n_reads; ++r) {
    const char* s = seq + seq_offsets[r];
    const uint64_t len = seq_offsets[r + 1] - seq_offsets[r];
    record[r] = 0;
    if (votes) votes[r] = 0;
    if (len == 0 || *s == '\0') continue;      /* "abnormal read entry skipped", io/ioFastQ.cc:21-25 */
    if (len < (uint64_t)k + 10) continue;      /* :26 */
    uint8_t km[KLO_KB], rep[KLO_KB];
    klo_kmer_from_string(s, k, km);
    float kmer_count = 0;
    for (uint64_t j = 0; j <= len - (uint64_t)k; ++j) {
      if (j > 0) klo_kmer_forward(km, k, s[j + (uint64_t)k - 1]);
      klo_kmer_rep(km, k, rep);
      if (n_kmers && bsearch(rep, set, (size_t)n_kmers, KLO_KB, klo_rec_cmp)) kmer_count++;
    }
    const float tmp_ratio = kmer_count / (float)(len - (uint64_t)k + 1); /* float / size_t, :57 */
    if (tmp_ratio > kmer_vote) record[r] = 1;
    if (votes) votes[r] = (uint32_t)kmer_count;
  }
  free(set);
}
