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

#ifdef __cplusplus
}
#endif
#endif
