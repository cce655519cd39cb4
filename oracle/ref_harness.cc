// TEST INFRASTRUCTURE — not product code.
//
// Function-level harness around the UNMODIFIED reference sources, compiled where
// they lie under /root/reference by oracle/Makefile into oracle/_ref/libklsh_ref.so.
// It exposes the reference's own hot-path functions through a flat C interface so
// that (a) the C restatement in oracle/klsh_oracle.c can be pinned against the real
// thing and (b) tests/golden/ fixtures can be minted (tests/golden/make_golden.py).
//
// Reference entry points wrapped here:
//   LSH::generateHashTable      hash/lshash.cc:36-42
//   LSH::random_projection      hash/lshash.cc:44-59
//   Distance::cosine            function/distance.cc:27-38
//   AB::SetConsensus            function/funcAB.cc:49-71
//   p_cluster                   function/cluster.cc:56-87
//   nestedCluster               function/cluster.cc:89-178
//   Cluster                     function/cluster.cc:181-340
//   IOMat::convertHTMat         io/ioMatrix.cc:353-408
//   IOMat::SaveResult/SaveBinary io/ioMatrix.cc:265-294, :322-351
//   IOMat::ReadClusterAll       io/ioMatrix.cc:48-119
//   AB::WRS                     function/funcAB.cc:73-109   (mode E statistics)
//   alglib::studentttest2       utils/alglib-3.15.0/src/statistics.cpp:2277
//   IOFQ::CheckRead             io/ioFastQ.cc:5-75          (read extraction votes)
//   Kmer(const char*), twin(), operator<, writeBytes   kmer/Kmer.cc
//
// Determinism: load with OMP_THREAD_LIMIT=1 in the environment and pass threads=1
// (SURVEY.md D7, D9).
#include "seeded_rd.h"
#undef random_device

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "function/cluster.h"
#include "function/funcAB.h"
#include "utils/alglib-3.15.0/src/statistics.h"
#include <unordered_set>
#include <cstdio>
#include "io/ioFastQ.h"

namespace {

std::vector<Abundance*> g_rows;  // result of the last call that yields rows

void clear_rows() {
  for (size_t i = 0; i < g_rows.size(); ++i) delete g_rows[i];
  g_rows.clear();
}

void build_rows(std::vector<Abundance*>* out, const float* values, const uint64_t* id_offsets,
                const uint64_t* ids, uint64_t n, int dim) {
  out->clear();
  out->reserve(n);
  for (uint64_t r = 0; r < n; ++r) {
    Abundance* ab = new Abundance();
    ab->_values.assign(values + r * dim, values + (r + 1) * dim);
    ab->_ids.assign(ids + id_offsets[r], ids + id_offsets[r + 1]);
    out->push_back(ab);
  }
}

}  // namespace

extern "C" {

void ref_reseed(unsigned long long seed) { klsh_oracle::reseed(seed); }
unsigned long long ref_master_draws() { return klsh_oracle::master().draws; }

void ref_generate_table(int H, int D, float* out) {
  hashTable t = LSH::generateHashTable(H, D);
  for (int h = 0; h < H; ++h) std::memcpy(out + (size_t)h * D, t[h].data(), sizeof(float) * D);
}

void ref_random_projection(const float* rows, uint64_t n, int D, const float* table, int H, int* keys) {
  hashTable t(H, hashFunction(D));
  for (int h = 0; h < H; ++h) t[h].assign(table + (size_t)h * D, table + (size_t)(h + 1) * D);
  for (uint64_t r = 0; r < n; ++r) {
    std::vector<float> v(rows + r * D, rows + (r + 1) * D);
    keys[r] = LSH::random_projection(v, t);
  }
}

float ref_cosine(const float* a, const float* b, int D) {
  return Distance::cosine(std::vector<float>(a, a + D), std::vector<float>(b, b + D));
}

void ref_set_consensus(const float* cur, uint64_t cur_count, const float* cand, uint64_t cand_count, int D,
                       float* out) {
  Abundance a, b, r;
  a._values.assign(cur, cur + D);
  a._ids.assign(cur_count, 0);
  b._values.assign(cand, cand + D);
  b._ids.assign(cand_count, 0);
  AB::SetConsensus(&r, a, b);
  std::memcpy(out, r._values.data(), sizeof(float) * D);
}

// --- calls that produce a row set; fetch it with ref_result_* -----------------

void ref_p_cluster(const float* values, const uint64_t* id_offsets, const uint64_t* ids, uint64_t n, int D,
                   float threshold) {
  clear_rows();
  std::vector<Abundance*> cand;
  build_rows(&cand, values, id_offsets, ids, n, D);
  p_cluster(&g_rows, &cand, threshold);
}

void ref_nested_cluster(const float* values, const uint64_t* id_offsets, const uint64_t* ids, uint64_t n,
                        int D, float threshold, int threads) {
  clear_rows();
  build_rows(&g_rows, values, id_offsets, ids, n, D);
  nestedCluster(&g_rows, threshold, D, threads, false);
}

void ref_cluster(const float* values, const uint64_t* id_offsets, const uint64_t* ids, uint64_t n, int D,
                 float min_similarity, int iterations, unsigned threads, int bucket_size_threshold,
                 int verbose) {
  clear_rows();
  build_rows(&g_rows, values, id_offsets, ids, n, D);
  Cluster(&g_rows, min_similarity, iterations, threads, D, bucket_size_threshold, verbose != 0);
}

// counts: sample-major uint16 [D][batch_size] (the layout ReadHT fills, io/ioHT.cc:59-81)
void ref_convert_ht_mat(const uint16_t* counts, const float* v_kmers, int D, uint64_t batch_size,
                        uint64_t batch_offset) {
  clear_rows();
  std::vector<uint16_t*> ptrs(D);
  for (int j = 0; j < D; ++j) ptrs[j] = const_cast<uint16_t*>(counts) + (size_t)j * batch_size;
  std::vector<float_t> vk(v_kmers, v_kmers + D);
  IOMat::convertHTMat(ptrs.data(), vk, D, false, batch_size, (streamoff)batch_offset, &g_rows);
}

void ref_read_cluster_all(const char* path, int D) {
  clear_rows();
  IOMat::ReadClusterAll(&g_rows, D, path, false);
}

uint64_t ref_result_rows() { return g_rows.size(); }
uint64_t ref_result_ids() {
  uint64_t t = 0;
  for (size_t i = 0; i < g_rows.size(); ++i) t += g_rows[i]->_ids.size();
  return t;
}
void ref_result_copy(float* values, uint64_t* id_offsets, uint64_t* ids, int D) {
  uint64_t off = 0;
  for (size_t r = 0; r < g_rows.size(); ++r) {
    std::memcpy(values + r * D, g_rows[r]->_values.data(), sizeof(float) * D);
    id_offsets[r] = off;
    std::memcpy(ids + off, g_rows[r]->_ids.data(), sizeof(uint64_t) * g_rows[r]->_ids.size());
    off += g_rows[r]->_ids.size();
  }
  id_offsets[g_rows.size()] = off;
}

// Write the current row set with the reference's own writers.
void ref_save(const char* bin_path, int delfile, int ignore_small) {
  IOMat::SaveResult(&g_rows, std::string(bin_path) + ".clust", delfile != 0, ignore_small, false);
  IOMat::SaveBinary(&g_rows, bin_path, delfile != 0, ignore_small, false);
}

// alglib::studentttest2 as AB::WRS calls it (function/funcAB.cc:95-97)
void ref_studentttest2(const double* x, int n, const double* y, int m, double* both, double* left, double* right) {
  alglib::real_1d_array ax, ay;
  ax.setcontent(n, x);
  ay.setcontent(m, y);
  alglib::studentttest2(ax, n, ay, m, *both, *left, *right);
}

// The loop of app/kmerLSH.cc:543-545 over a row set with the reference's own AB::WRS; the two id sets come
// back as one label per k-mer id (1: g_kmer_id1, 2: g_kmer_id2 only) and, per row, which set its ids went to.
void ref_wrs(const float* values, const uint64_t* id_offsets, const uint64_t* ids, uint64_t n, int D, int num_sample1,
             int num_sample2, float pvalue_thresh, int size_thresh, uint64_t n_kmers, uint8_t* row_group,
             uint8_t* id_label) {
  std::vector<Abundance*> rows;
  build_rows(&rows, values, id_offsets, ids, n, D);
  std::unordered_set<uint64_t> g1, g2;
  for (uint64_t r = 0; r < n; ++r) {
    const size_t b1 = g1.size(), b2 = g2.size();
    AB::WRS(&g1, &g2, rows[r], num_sample1, num_sample2, pvalue_thresh, size_thresh);
    row_group[r] = g1.size() != b1 ? 1 : (g2.size() != b2 ? 2 : 0);
  }
  std::memset(id_label, 0, n_kmers);
  for (uint64_t i = 0; i < n_kmers; ++i) {  // precedence of app/kmerLSH.cc:571-576
    if (g1.find(i) != g1.end()) id_label[i] = 1;
    else if (g2.find(i) != g2.end()) id_label[i] = 2;
  }
  for (size_t r = 0; r < rows.size(); ++r) delete rows[r];
}

// Kmer::set_k may be called once per process (kmer/Kmer.cc:343-351): the first k wins, another k is refused.
int ref_set_k(int k) {
  if (Kmer::k == 0) Kmer::set_k((unsigned)k);
  return (int)Kmer::k == k ? 0 : -1;
}

static void kmer_bytes(const Kmer& km, uint8_t* out) {  // the 8 bytes writeBytes emits (kmer/Kmer.cc:309-313)
  char* buf = nullptr;
  size_t len = 0;
  FILE* f = open_memstream(&buf, &len);
  km.writeBytes(f);
  fclose(f);
  std::memcpy(out, buf, Kmer::MAX_K / 4);
  free(buf);
}

// rep = (km < tw) ? km : tw of the first k characters of s, as the record kmer_set.hex holds
int ref_kmer_rep(const char* s, int k, uint8_t* km_out, uint8_t* rep_out) {
  if (ref_set_k(k) != 0) return -1;
  Kmer km(s);
  Kmer tw = km.twin();
  Kmer rep = (km < tw) ? km : tw;
  kmer_bytes(km, km_out);
  kmer_bytes(rep, rep_out);
  return 0;
}

// IOFQ::CheckRead with one thread (tid 0 of 1) on reads given as one character array + offsets
int ref_check_reads(const uint8_t* kmers, uint64_t n_kmers, int k, const char* seq, const uint64_t* offs, uint64_t n_reads,
                    float kmer_vote, uint8_t* record) {
  if (ref_set_k(k) != 0) return -1;
  uset_t set;
  for (uint64_t i = 0; i < n_kmers; ++i) set.insert(Kmer(const_cast<uint8_t*>(kmers + i * (Kmer::MAX_K / 4))));
  std::vector<ReadEntry> reads(n_reads);
  for (uint64_t r = 0; r < n_reads; ++r) {
    const uint64_t len = offs[r + 1] - offs[r];
    if (len >= sizeof(reads[r].s)) return -2;
    std::memcpy(reads[r].s, seq + offs[r], len);
    reads[r].s[len] = '\0';
    reads[r].len = len;
    std::snprintf(reads[r].name, sizeof(reads[r].name), "r%llu", (unsigned long long)r);
    reads[r].name_len = std::strlen(reads[r].name);
    reads[r].qual[0] = '\0';
  }
  std::vector<int> rec(n_reads, 0);
  Utility::IOFQ::CheckRead(&set, reads, rec, 1, 0, kmer_vote);
  for (uint64_t r = 0; r < n_reads; ++r) record[r] = (uint8_t)rec[r];
  return 0;
}

}  // extern "C"
