/* klsh.h — C ABI of the B200-native kmerLSH mode-C clustering hot path.
 *
 * The reference (wthanone/kmerLSH) has no plugin/FFI interface: the seam of its hot path is the
 * C++ free function
 *     void Cluster(vector<Abundance*>*, float min_similarity, int cluster_iteration,
 *                  unsigned threads_to_use, int dim, int bucket_size_threshold, bool verbose)
 * (reference function/cluster.h:42; call sites app/kmerLSH.cc:323, :377, :490) plus the files
 * either side of it.  This header is what a binding for that seam binds: plain C types, caller-
 * owned host buffers, library-owned device memory, `int` status (0 = ok) instead of exit().
 * INTEGRATION.md shows the reference-side shim that routes Cluster() through it.
 *
 * One context = one GPU = one driving host thread (not re-entrant).  The library never prints
 * (exception: with KLSH_DEBUG=1 in the environment it writes per-kernel diagnostics to stderr).
 * There is no CPU fallback: every entry point fails with KLSH_ERR_CUDA when no sm_100 device is
 * usable.
 * Row width: any D up to 768 samples.  Up to D = 280 the in-bucket merge is the windowed tensor-core
 * kernel; wider rows take a block-per-bucket kernel automatically (exact, much slower).  Beyond 768 the
 * signing kernel's plane tables no longer fit in shared memory and klsh_cluster returns KLSH_ERR_ARG.
 */
#ifndef KLSH_H
#define KLSH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KLSH_OK 0
#define KLSH_ERR_ARG 1   /* bad argument / wrong state */
#define KLSH_ERR_CUDA 2  /* CUDA runtime failure (see klsh_last_error) */
#define KLSH_ERR_IO 3    /* file could not be opened / short read */
#define KLSH_ERR_NOMEM 4 /* device or host allocation failed */

typedef struct klsh_ctx klsh_ctx;

/* Fills out[H][D] with the hyperplanes of one hash table.  Replaces
 * LSH::generateHashTable(H, D) (reference hash/lshash.cc:36-42); a reference-side binding passes a
 * trampoline to that very function so both sides draw from the reference's own generator. */
typedef void (*klsh_plane_fn)(void* user, int H, int D, float* out);

/* Per-iteration record; mirrors the reference's --verbose lines (function/cluster.cc:209-211,
 * :263, :307, :325-326) plus device timings. */
typedef struct {
  uint64_t rows_in, rows_out;
  int32_t H;
  float threshold;
  uint64_t buckets, bucket_max, nested_calls;
  uint64_t eps_margin_rows; /* rows whose key needed the exact re-evaluation path (0 = exact everywhere) */
  float ms_sign, ms_group, ms_merge, ms_compact, ms_total; /* CUDA-event times on the context's stream */
  /* compare work of the in-bucket merge: (window candidate, representative) pairs screened on the tensor
   * cores and the pairs that went on to the reference's exact fp32 test */
  uint64_t screen_pairs, exact_pairs;
} klsh_iter_stats;

/* ---- lifetime -------------------------------------------------------------------------------- */
int klsh_create(int device, klsh_ctx** out);
void klsh_destroy(klsh_ctx* ctx);
/* Message of the last failure on ctx (ctx == NULL: last klsh_create failure).  Never NULL. */
const char* klsh_last_error(const klsh_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t klsh_launch_count(const klsh_ctx* ctx);

/* ---- hyperplane source (reference hash/lshash.cc:3-17, :36-42) -------------------------------- */
/* Built-in source: one process-wide-style master std::mt19937_64(seed); every hash function is
 * std::mt19937(uint32(master())) + std::normal_distribution<double>(0,1) narrowed to float —
 * the reference's generator with its std::random_device replaced by a seeded one. */
int klsh_set_seed(klsh_ctx* ctx, uint64_t seed);
int klsh_set_plane_source(klsh_ctx* ctx, klsh_plane_fn fn, void* user);
/* Position of the built-in source: its seed and the number of hash functions drawn since (every table
 * of H functions advances it by H).  klsh_plane_seek(seed, drawn) puts any context at that position,
 * e.g. to give several contexts identical copies of one stream. */
int klsh_plane_tell(const klsh_ctx* ctx, uint64_t* seed, uint64_t* drawn);
int klsh_plane_seek(klsh_ctx* ctx, uint64_t seed, uint64_t drawn);
/* fn(user) is called once per klsh_cluster call, from inside it, as soon as the call has drawn its LAST
 * hyperplane table (the table of its last iteration and those of that iteration's oversized buckets) —
 * before that iteration's merge runs.  A driver that serialises several contexts on one generator stream
 * (the reference draws its tables from one stream, batch after batch: app/kmerLSH.cc:311-345) uses it to
 * hand the stream to the next context while this one is still merging. */
typedef void (*klsh_done_fn)(void* user);
int klsh_set_draws_done_callback(klsh_ctx* ctx, klsh_done_fn fn, void* user);
/* Draw one table from the current source (advances it), e.g. to replay what a run used. */
int klsh_draw_table(klsh_ctx* ctx, int H, int D, float* out);

/* ---- rows in --------------------------------------------------------------------------------- */
/* = ReadHT (reference io/ioHT.cc:59-81) output handed to IOMat::convertHTMat
 * (io/ioMatrix.cc:353-408): counts is the sample-major uint16 block [D][batch_size] of
 * kmer_count.bin, v_kmers[j] = float(coverage_j)/kmap_size (app/kmerLSH.cc:480).  Row i becomes
 * value_j = float(log(cnt+1.0)) - v_kmers[j], kept iff sum_j cnt > 0.1*D, with id batch_offset+i.
 * Replaces the context's row set. */
int klsh_load_counts(klsh_ctx* ctx, const uint16_t* counts, const float* v_kmers, int D, uint64_t batch_size,
                     uint64_t batch_offset);
/* = a vector<Abundance*> handed to Cluster: values[n][D], ids of row r = ids[id_offsets[r] ..
 * id_offsets[r+1]).  Replaces the context's row set. */
int klsh_set_rows(klsh_ctx* ctx, const float* values, const uint64_t* id_offsets, const uint64_t* ids,
                  uint64_t n, int D);
/* = IOMat::ReadCluster / ReadClusterAll (io/ioMatrix.cc:121-196, :48-119): rows
 * [start_line, start_line+num_lines) of <bin_path> and <bin_path>.clust; num_lines == 0 reads all. */
int klsh_load_cluster_file(klsh_ctx* ctx, const char* bin_path, int D, uint64_t start_line, uint64_t num_lines);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* = Cluster(rows, min_similarity, cluster_iteration, T, dim, bucket_size_threshold, verbose)
 * (function/cluster.cc:181-340) on the context's row set, in place.  threads_to_use has no
 * meaning here; results equal the reference's T=1 run (its only deterministic schedule).
 * stats may be NULL, else has room for `iterations` records. */
int klsh_cluster(klsh_ctx* ctx, float min_similarity, int iterations, int64_t bucket_size_threshold,
                 klsh_iter_stats* stats);

/* Function-level entry points (parity tests call these; they are the pieces of Cluster): */
/* LSH::random_projection(row, table) for n host rows (hash/lshash.cc:44-59); key bit order as the
 * reference: plane 0 is the most significant of the H bits. */
int klsh_sign(klsh_ctx* ctx, const float* rows, uint64_t n, int D, const float* table, int H, uint64_t* keys_out);
/* p_cluster (function/cluster.cc:56-87) on the context's whole row set as ONE bucket. */
int klsh_p_cluster(klsh_ctx* ctx, float threshold);
/* nestedCluster (function/cluster.cc:89-178) on the context's whole row set. */
int klsh_nested_cluster(klsh_ctx* ctx, float threshold);

/* Distance::cosine(left, right) (function/distance.cc:27-38) for n pairs of host rows [n][D]:
 * out[k] = 1 - dot / (sqrtf(|left|^2) * sqrtf(|right|^2)), every sum accumulated in index order. */
int klsh_cosine_distance(klsh_ctx* ctx, const float* left, const float* right, uint64_t n, int D, float* out);
/* The values AB::SetConsensus(current, candidate) produces (function/funcAB.cc:49-71) for rows
 * with n_current and n_candidate member ids: out[i] = current[i]*n_current/all + candidate[i]*n_candidate/all,
 * counts converted int -> float as the reference's cvtsi2ss does. */
int klsh_set_consensus(klsh_ctx* ctx, const float* current, int64_t n_current, const float* candidate, int64_t n_candidate,
                       int D, float* out);

/* ---- rows out -------------------------------------------------------------------------------- */
int klsh_row_count(klsh_ctx* ctx, uint64_t* n_rows, uint64_t* n_ids);
int klsh_get_rows(klsh_ctx* ctx, float* values, uint64_t* id_offsets, uint64_t* ids);
/* = IOMat::SaveResult(rows, path+".clust", delfile, ignore_small) + IOMat::SaveBinary(rows, path,
 * delfile, ignore_small) (io/ioMatrix.cc:265-294, :322-351). */
int klsh_save(klsh_ctx* ctx, const char* bin_path, int delfile, int64_t ignore_small);
/* Member-list file format used by klsh_save and klsh_load_cluster_file from now on: 0 = the reference's text
 * <F>.clust (default), 1 = binary <F>.clust.bin — per cluster a uint64 count followed by that many uint64 ids, host
 * endianness, same cluster order as <F>.  No reference counterpart (opt-in): it spares the
 * decimal formatting and parsing of the spill files' 10^8 ids per batch.  The centroid file <F> is the same either way. */
int klsh_set_id_format(klsh_ctx* ctx, int format);

/* ---- mode E statistics on the clusters -------------------------------------
 * = the loop of app/kmerLSH.cc:541-545 over the clustering result: AB::WRS (function/funcAB.cc:73-109) per
 * cluster of the context's current row set (after klsh_cluster, or klsh_load_cluster_file for a result on disk).
 * A cluster with more than size_thresh member ids is tested with alglib::studentttest2 on values[0 .. num_sample1)
 * against values[num_sample1 .. num_sample1+num_sample2) (widened to double); row_group[r] = 2 when
 * lefttail <= pvalue_thresh (its ids join the group-B set, :100-101), else 1 when righttail <= pvalue_thresh (group-A
 * set, :102-103), else 0.  lefttail/righttail (optional, [rows]) receive the two probabilities, -1 for untested rows.
 * The test statistic is bit-identical to ALGLIB's; the tail probabilities agree to a relative 1e-9 (ALGLIB's Cephes
 * incomplete beta is replaced by a continued fraction, see csrc/stats.cu) and `margin` counts the clusters whose
 * decision lies inside that tolerance of the threshold. */
typedef struct {
  uint64_t rows, tested;     /* clusters in the row set / clusters above size_thresh */
  uint64_t rows_a, rows_b;   /* clusters filed under group A (row_group 1) / group B (row_group 2) */
  uint64_t ids_a, ids_b;     /* their member ids */
  uint64_t margin;           /* tested clusters with a tail probability within 1e-9 (relative) of pvalue_thresh */
} klsh_ttest_stats;
int klsh_ttest(klsh_ctx* ctx, int num_sample1, int num_sample2, float pvalue_thresh, int size_thresh,
               uint8_t* row_group /* [rows] or NULL */, double* lefttail /* [rows] or NULL */,
               double* righttail /* [rows] or NULL */, klsh_ttest_stats* stats /* or NULL */);
/* The two id sets g_kmer_id1 / g_kmer_id2 the same loop fills (app/kmerLSH.cc:543-545), as one label per k-mer id
 * below n_kmers (= kmap_size): 1 = in the group-A set, 2 = in the group-B set only, 0 = in neither — the
 * precedence of the join at :571-576.  Runs the test itself (same arguments as klsh_ttest). */
int klsh_differential_ids(klsh_ctx* ctx, int num_sample1, int num_sample2, float pvalue_thresh, int size_thresh,
                          uint64_t n_kmers, uint8_t* id_label /* [n_kmers] */, klsh_ttest_stats* stats /* or NULL */);
/* = the join over kmer_set.hex (app/kmerLSH.cc:565-579): records[n_kmers][record_bytes] are the k-mers in id order
 * (record_bytes = Kmer::MAX_K/4 = 8 in the reference build, kmer/Kmer.h:68-78); out_a / out_b (room for n_kmers
 * records each) receive the records labelled 1 / 2 in id order — the contents of g_kmer1 / g_kmer2. */
int klsh_select_kmers(klsh_ctx* ctx, const uint8_t* records, uint64_t n_kmers, int record_bytes, const uint8_t* id_label,
                      uint8_t* out_a, uint64_t* n_a, uint8_t* out_b, uint64_t* n_b);

/* ---- read extraction votes ------------------------------------------------
 * = IOFQ::CheckRead (io/ioFastQ.cc:5-75), the test IOFQ::ReadExtract (:77-158) applies to every read of a FASTQ part.
 * klsh_kmer_set_load makes records[n_kmers][8] (Kmer::MAX_K/4 bytes each, the layout of kmer_set.hex and of
 * klsh_select_kmers' output) the context's set of differential k-mers (g_kmer1 or g_kmer2, app/kmerLSH.cc:583-584).
 * klsh_check_reads tests n_reads reads given as one character array and n_reads+1 offsets into it:
 * record[r] = 1 iff the read has at least k+10 characters, does not start with a NUL character, and
 * float(#k-mers of the read whose canonical form is in the set) / float(len-k+1) > kmer_vote (float arithmetic as the
 * reference's; characters other than A, C, G, T count as A, kmer/Kmer.cc:139-144).  votes (optional) = the counts. */
int klsh_kmer_set_load(klsh_ctx* ctx, const uint8_t* records, uint64_t n_kmers, int record_bytes /* 8 */);
int klsh_check_reads(klsh_ctx* ctx, int k, const char* seq, const uint64_t* seq_offsets /* [n_reads+1] */, uint64_t n_reads,
                     float kmer_vote, uint8_t* record /* [n_reads] */, uint32_t* votes /* [n_reads] or NULL */);

/* ---- device-resident state control (benchmarks; multi-batch phase 1) --------------------------- */
/* Remember / restore the current row set on the device (no host traffic). */
int klsh_snapshot(klsh_ctx* ctx);
int klsh_restore(klsh_ctx* ctx);
/* Block until all work queued on the context's stream is done. */
int klsh_sync(klsh_ctx* ctx);

/* ---- survivors of several batches, resident on the device ----------------------------------------
 * The reference appends every phase-1 batch's survivors to tmp/0.bin(.clust) and reads the files back
 * (app/kmerLSH.cc:326-335, :415).  klsh_stash_rows appends the context's current working set to a
 * device-resident stash instead (call it after each batch's klsh_cluster); klsh_unstash_rows makes the stash —
 * all batches in append order — the current row set.  Same rows, same order, same member lists as the
 * file round trip.  Member ids stay implicit when the batches came from klsh_load_counts with contiguous
 * offsets. */
int klsh_stash_rows(klsh_ctx* ctx);
int klsh_stash_count(const klsh_ctx* ctx, uint64_t* n_rows);
int klsh_unstash_rows(klsh_ctx* ctx);

/* ---- multi-GPU building blocks ------------------------------------------------------------------
 * No reference counterpart (the reference is one process).  One context per rank; every rank holds
 * the same row set (load it identically on all ranks) and the same hyperplane source.  Per LSH
 * iteration of Cluster (function/cluster.cc:199-331) the caller runs, on every rank,
 *   klsh_mg_pass_begin  sign + group ALL rows (replicated, no communication)
 *   klsh_mg_plan        contiguous bucket ranges, one per rank, balanced by row count
 *   klsh_mg_merge       p_cluster / nestedCluster on the rank's own bucket range; logs what changed
 *   klsh_mg_export      survivors of the range + modified rows + member-chain writes -> device buffers
 *   (all-gather of those buffers over NCCL — kmerlsh_b200/distributed.py)
 *   klsh_mg_apply       replay the other ranks' logs on this replica
 *   klsh_mg_set_alive   new working set = survivor lists concatenated in rank order
 * Results are identical to klsh_cluster on one GPU.  d_* arguments are DEVICE pointers. */
int klsh_row_stride(const klsh_ctx* ctx); /* floats per row in exported/applied row blocks (D rounded up to 4) */
int klsh_mg_pass_begin(klsh_ctx* ctx, uint64_t* n_rows, int32_t* H, uint64_t* n_buckets);
int klsh_mg_plan(klsh_ctx* ctx, int world, uint32_t* splits_out /* host, world+1 bucket indices */);
int klsh_mg_merge(klsh_ctx* ctx, uint32_t bucket_lo, uint32_t bucket_hi, float threshold, int64_t bucket_size_threshold,
                  uint64_t* n_survivors, uint64_t* n_modified_rows, uint64_t* n_chain_writes);
int klsh_mg_export(klsh_ctx* ctx, uint32_t* d_survivors, uint32_t* d_mod_rows, float* d_mod_vals /* [n][stride] */,
                   int32_t* d_mod_meta /* [n][3] count, head, tail */, uint32_t* d_chain_slots, int32_t* d_chain_vals);
int klsh_mg_apply(klsh_ctx* ctx, const uint32_t* d_mod_rows, const float* d_mod_vals, const int32_t* d_mod_meta,
                  uint64_t n_modified_rows, const uint32_t* d_chain_slots, const int32_t* d_chain_vals,
                  uint64_t n_chain_writes);
int klsh_mg_set_alive(klsh_ctx* ctx, const uint32_t* d_alive, uint64_t n);


/* ---- multi-GPU, NCCL inside the library (kmerlsh_b200/csrc/multi.cu) ---------------------------
 * One context per GPU/rank.  klsh_nccl_unique_id (rank 0) makes the 128-byte NCCL id the caller hands
 * to every rank by its own means (threads of one process: a shared variable; processes: any
 * broadcast).  klsh_mg_init joins the communicator (collective: every rank calls it concurrently). */
int klsh_nccl_unique_id(void* out, uint64_t bytes /* >= 128 */);
int klsh_mg_init(klsh_ctx* ctx, int rank, int world, const void* unique_id, uint64_t bytes);
int klsh_mg_finalize(klsh_ctx* ctx);
int klsh_mg_rank(const klsh_ctx* ctx, int* rank, int* world);
/* = Cluster(...) (function/cluster.cc:181-340) over ALL ranks' GPUs: every rank holds the same row set
 * and the same hyperplane source and calls this collectively; the loop over the klsh_mg_* building
 * blocks above and the NCCL exchange (sizes: one ncclAllGather; survivors + modified rows + chain
 * writes: one grouped ncclBroadcast per rank) run inside the library.  Every rank ends with the
 * clusters klsh_cluster computes on one GPU, bit for bit. */
int klsh_mg_cluster(klsh_ctx* ctx, float min_similarity, int iterations, int64_t bucket_size_threshold,
                    klsh_iter_stats* stats);
/* All-gather of row sets over NVLink: every rank contributes its working set (the survivors of its
 * phase-1 batch, app/kmerLSH.cc:311-345) and ends with the concatenation in rank order — the vector
 * the reference gets by appending batch after batch to tmp/0.bin and reading it back (:326-335, :415).
 * Rows must come from klsh_load_counts with batch offsets contiguous in rank order. */
int klsh_mg_gather_rows(klsh_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
