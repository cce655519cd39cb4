/* TEST INFRASTRUCTURE — not product code.  See klsh_oracle.c. */
#ifndef KLSH_ORACLE_H
#define KLSH_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- hyperplane source (hash/lshash.cc:3-17, :36-42 under oracle/seeded_rd.h) ---- */
typedef struct klo_planes klo_planes;
klo_planes* klo_planes_new(uint64_t seed);
void klo_planes_free(klo_planes* p);
void klo_planes_reseed(klo_planes* p, uint64_t seed);
uint64_t klo_planes_draws(const klo_planes* p); /* master draws consumed so far */
void klo_planes_table(klo_planes* p, int H, int D, float* out /* [H][D] */);

/* ---- row transform (io/ioMatrix.cc:353-408, app/kmerLSH.cc:473-481) ---- */
void klo_log_lut(float* lut /* [65536] */);
void klo_vkmers(const float* coverage, uint64_t kmap_size, int D, float* out);
uint64_t klo_convert_counts(const uint16_t* counts /* [D][batch] sample-major */, const float* v_kmers,
                            int D, uint64_t batch_size, uint64_t batch_offset, float* values_out,
                            uint64_t* ids_out);

/* ---- scalar kernels ---- */
void klo_sign(const float* rows, uint64_t n, int D, const float* table, int H, uint32_t* keys);
float klo_cosine_distance(const float* lhs, const float* rhs, int D);
void klo_consensus(const float* cur, int64_t cur_count, const float* cand, int64_t cand_count, int D,
                   float* out);
float klo_threshold_after(float min_similarity, int iterations, int steps);

/* ---- row sets ---- */
typedef struct klo_rows klo_rows;
klo_rows* klo_rows_new(const float* values, const uint64_t* id_offsets, const uint64_t* ids, uint64_t n,
                       int D);
void klo_rows_free(klo_rows* r);
uint64_t klo_rows_count(const klo_rows* r);
uint64_t klo_rows_members(const klo_rows* r);
int klo_rows_dim(const klo_rows* r);
void klo_rows_export(const klo_rows* r, float* values, uint64_t* id_offsets, uint64_t* ids);
/* append src's rows after dst's (batch concatenation in init_clustering); src is left empty */
void klo_rows_append(klo_rows* dst, klo_rows* src);

/* ---- clustering (function/cluster.cc) ---- */
typedef struct {
  uint64_t rows_in, rows_out;
  int H;
  float threshold;
  uint64_t buckets_nonempty, bucket_max, nested_calls;
  uint64_t compares, merges;
} klo_iter_stats;

void klo_p_cluster(klo_rows* r, float threshold); /* whole row set as ONE bucket */
void klo_nested_cluster(klo_rows* r, float threshold, klo_planes* planes);
/* stats may be NULL; otherwise room for `iterations` entries */
void klo_cluster(klo_rows* r, float min_similarity, int iterations, int64_t bucket_size_threshold,
                 klo_planes* planes, klo_iter_stats* stats);
/* keys of the LAST top-level signing pass done by klo_cluster (debug aid), length rows_in */
/* bucket size histogram probe: sizes of non-empty buckets for one signing of r (ascending key) */
uint64_t klo_bucket_sizes(const klo_rows* r, const float* table, int H, uint64_t* sizes_out, uint64_t cap);

/* ---- files (io/ioMatrix.cc:265-294, :322-351, :48-119) ---- */
int klo_save(const klo_rows* r, const char* bin_path, int delfile, int64_t ignore_small);
klo_rows* klo_read_cluster(const char* bin_path, int D, uint64_t start_line, uint64_t num_lines /*0=all*/);

/* ---- mode C end to end (app/kmerLSH.cc:278-430, :469-499) ---- */
/* cwd-independent: paths given explicitly.  Reference constants: batch_thresh = 100000000
 * (app/kmerLSH.cc:285), phase2_bucket_threshold = 1000000 (:440). */
int klo_mode_c(const char* count_bin, const char* count_log, int D, float min_similarity, int iterations,
               const char* tmp_dir, const char* out_path, uint64_t batch_thresh,
               int64_t phase2_bucket_threshold, uint64_t seed, klo_iter_stats* phase2_stats /* may be NULL */);

/* ---- mode E statistics (function/funcAB.cc:73-109, app/kmerLSH.cc:541-585) ---- */
/* alglib::studentttest2 (utils/alglib-3.15.0/src/statistics.cpp:12502-12616) on x[n], y[m]. */
void klo_ttest2(const double* x, int n, const double* y, int m, double* bothtails, double* lefttail,
                double* righttail);
/* AB::WRS for every row of a row set: group[r] = 2 if lefttail <= pvalue_thresh (the ids go to the group-B
 * set, funcAB.cc:100-101), 1 if righttail <= pvalue_thresh (group-A set, :102-103), else 0; rows with
 * |ids| <= size_thresh are not tested (0, tails reported as -1).  lefttail/righttail may be NULL. */
void klo_wrs_rows(const float* values, const uint64_t* id_offsets, uint64_t n, int D, int num_sample1,
                  int num_sample2, float pvalue_thresh, int size_thresh, uint8_t* group, double* lefttail,
                  double* righttail);
/* The two id sets of app/kmerLSH.cc:543-545 as one label per k-mer id < n_kmers (1: in g_kmer_id1,
 * 2: only in g_kmer_id2, 0: neither — the precedence of the join at :571-576). */
void klo_differential_ids(const uint8_t* group, const uint64_t* id_offsets, const uint64_t* ids, uint64_t n,
                          uint64_t n_kmers, uint8_t* id_label);
/* The join over kmer_set.hex (app/kmerLSH.cc:565-579): records of `record_bytes` bytes, record i belongs
 * to k-mer id i; out_a / out_b receive the records labelled 1 / 2 in id order; returns the two counts. */
void klo_select_kmers(const uint8_t* records, uint64_t n_kmers, int record_bytes, const uint8_t* id_label,
                      uint8_t* out_a, uint64_t* n_a, uint8_t* out_b, uint64_t* n_b);

/* ---- read extraction votes (io/ioFastQ.cc:5-75, kmer/Kmer.cc) ---- */
/* Kmer(const char*) (kmer/Kmer.cc:131-150): the first k characters packed 2 bits per base, base i in byte i/4 at
 * bit 2*(i%4), A=0 C=1 G=2 T=3, any other character 0; out is one 8-byte record (Kmer::MAX_K/4, zero padded). */
void klo_kmer_from_string(const char* s, int k, uint8_t* out /* [8] */);
/* Kmer::twin() (kmer/Kmer.cc:160-185) followed by the `rep = (km < tw) ? km : tw` of the callers (memcmp order). */
void klo_kmer_rep(const uint8_t* km /* [8] */, int k, uint8_t* out /* [8] */);
/* IOFQ::CheckRead (io/ioFastQ.cc:5-75) over reads given as one character array and n_reads+1 offsets:
 * record[r] = 1 iff the read has at least k+10 characters, does not start with a NUL, and
 * float(k-mers whose canonical form is in the set) / float(len-k+1) > kmer_vote.  kmers: n_kmers 8-byte records
 * (any order, the set g_kmer).  votes (optional): the count per read (0 for skipped reads). */
void klo_check_reads(const uint8_t* kmers, uint64_t n_kmers, int k, const char* seq, const uint64_t* seq_offsets,
                     uint64_t n_reads, float kmer_vote, uint8_t* record, uint32_t* votes);

#ifdef __cplusplus
}
#endif
#endif
