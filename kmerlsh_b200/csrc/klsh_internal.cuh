// Internal declarations shared by the CUDA translation units of libklsh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/klsh.h"

#define KLSH_SENTINEL 0xFFFFFFFFu
#define KLSH_SMALL_MAX 32  // buckets of 2..32 rows: one warp each
#define KLSH_BIG 4096      // buckets at least this large are scheduled first

struct PlaneSource;  // planes.cc
struct MgComm;       // multi.cu

struct HostBuf {  // grow-only pinned host buffer
  void* p = nullptr;
  size_t bytes = 0;
};

// Grow-only device buffer.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// Scratch for one signing+grouping+merge pass over a list of rows.  The top-level pass and the
// nested pass (reference nestedCluster) each own one.
struct PassScratch {
  DevBuf keys_a, keys_b, rows_a, rows_b;  // sort ping-pong (uint32 each)
  DevBuf hist;                            // radix histograms
  DevBuf blkcnt;                          // per-block counts for compaction-style kernels
  DevBuf bstart;                          // bucket start offsets (uint32, nb+1)
  DevBuf list_small, list_nested;   // bucket ids
  DevBuf list_big, list_large;      // windowed-merge work items {bucket, i, size}: >= KLSH_BIG rows / the rest
  DevBuf list_direct;               // ... buckets of >= direct_min rows: straight to cluster teams on the second stream
  DevBuf esc1, esc2, esc3;          // buckets handed on: CTA -> cluster -> large cluster -> grid
  DevBuf escb2, escb3;              // the same hand-over lists of the direct pipeline
  DevBuf pos_nrm;   // norm of the representative at each sorted position (merge scratch)
  DevBuf pos_h;     // unit-norm fp16 copy of the representative at each sorted position (merge scratch, D <= 64)
  DevBuf planes;    // H*ld floats
  DevBuf counters;  // device counters (see PassCounters)
};

struct PassCounters {  // lives in device memory, mirrored to pinned host memory
  uint32_t n_buckets;
  uint32_t n_small;    // 2..KLSH_SMALL_MAX rows: one warp each
  uint32_t n_large;    // > KLSH_SMALL_MAX rows: windowed merge, starts on one CTA
  uint32_t n_big;      // ... the ones with >= KLSH_BIG rows (scheduled first)
  uint32_t n_esc1;     // handed on to a thread-block cluster
  uint32_t n_esc2;     // handed on to a large (non-portable size) cluster
  uint32_t n_esc3;     // handed on to the whole cooperative grid
  uint32_t n_nested;   // above bucket_size_threshold: reference nestedCluster
  uint32_t n_out;
  uint32_t bucket_max;
  uint32_t large_cursor;
  uint32_t cluster_cursor;
  uint32_t cluster2_cursor;
  // direct pipeline (buckets of >= direct_min rows, second stream): its own lists and cursors
  uint32_t n_direct;
  uint32_t nb_esc2, nb_esc3;
  uint32_t b_cursor1, b_cursor2;
  uint32_t direct_thr;     // the threshold k_classify chose (rows)
  uint32_t pad[1];
  uint32_t size_hist[32];  // buckets to merge by floor(log2(rows)), classes above KLSH_SMALL_MAX
};

// Multi-GPU update logs (all null in single-GPU runs): rows whose values/metadata changed and the
// member-chain pointer writes, appended by the merge kernels.  counts[0] = rows, counts[1] = next writes.
struct MgLog {
  uint32_t* counts = nullptr;
  uint32_t* mod_rows = nullptr;
  uint32_t* next_slot = nullptr;
  int32_t* next_val = nullptr;
};

// One column of the per-row member metadata.  Count, head and tail of a row share ONE 16-byte record
// {cnt, head, tail, 0}, so a gathered row touches one 32-byte sector of metadata instead of three.
struct MetaCol {
  int32_t* p;
  __host__ __device__ __forceinline__ int32_t& operator[](size_t i) const { return p[4 * i]; }
  __host__ __device__ __forceinline__ MetaCol operator+(size_t rows) const { return MetaCol{p + 4 * rows}; }
};

struct RowState {  // everything klsh_snapshot copies
  DevBuf vals, meta, next, alive;  // meta: [n_born] records {cnt, head, tail, 0}
  uint64_t n_alive = 0;
  MetaCol cnt() const { return MetaCol{meta.as<int32_t>()}; }       // member count of the row (cluster size)
  MetaCol head() const { return MetaCol{meta.as<int32_t>() + 1}; }  // first member slot of its id chain
  MetaCol tail() const { return MetaCol{meta.as<int32_t>() + 2}; }  // last member slot
};

struct klsh_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // the direct pipeline of the merge runs here, concurrently with the rest
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  uint64_t launches = 0;
  int sm_count = 148;
  int max_smem_optin = 0;

  // row arena: rows are born once and never move; a merge overwrites the surviving row in place.
  int D = 0, ld = 0;       // ld = D rounded up to a multiple of 4 floats
  uint64_t n_born = 0;     // rows in the arena
  uint64_t n_slots = 0;    // member slots (id list nodes)
  RowState cur, snap, stash;  // stash: survivors of earlier batches, appended on the device (klsh_stash_rows)
  uint64_t stash_rows = 0, stash_slots = 0, stash_id_base = 0;
  bool stash_implicit = true;
  int stash_D = 0;
  std::vector<uint64_t> stash_ids;  // explicit member ids of the stash (when not implicit)
  int id_format = 0;  // klsh_set_id_format: 0 text <F>.clust (the reference's), 1 binary <F>.clust.bin
  bool has_snap = false;
  uint64_t snap_born = 0, snap_slots = 0;
  // id payload of member slots: explicit (ids.size()==n_slots) or implicit id = id_base + slot
  std::vector<uint64_t> ids;
  uint64_t id_base = 0;
  bool ids_implicit = true;
  std::vector<uint64_t> snap_ids;
  uint64_t snap_id_base = 0;
  bool snap_ids_implicit = true;

  PassScratch top, nested;
  DevBuf lut;       // 65536 floats: float(log(c+1.0)) computed by the host libm
  bool lut_ready = false;
  DevBuf io_a, io_b;  // staging for loads/exports
  DevBuf alive_alt;   // the other half of the alive-list ping-pong
  DevBuf nested_out;  // survivors of one nested pass
  DevBuf team_ctl, team_ctl_b;  // per-team control blocks of the windowed merge (one set per pipeline)
  DevBuf pool_ctl, pool_pub;  // screen pool of the single-CTA teams: control word + board, per-leader blocks
  DevBuf pool_ctl_b, pool_pub_b;  // ... of the direct pipeline's cluster teams
  DevBuf exp_vals, exp_cnt, exp_head;  // export staging (device)
  DevBuf rank_buf, exp_offs, exp_slots; // chain ranking scratch, offsets and flat slot order (device)
  HostBuf h_slots;                     // flat slot order (pinned host)
  HostBuf h_cnt, h_head;               // export staging (pinned host)
  DevBuf st_group, st_left, st_right, st_counts, st_label, st_ids, st_slot_row;  // mode E statistics (stats.cu)
  DevBuf st_rec, st_lab, st_out_a, st_out_b, st_blk;
  DevBuf rd_table, rd_seq, rd_offs, rd_rec, rd_votes;  // read extraction votes (reads.cu): k-mer hash set, staged reads
  uint64_t rd_cap = 0, rd_n = 0;
  DevBuf eps_counter; // rows whose key needed the exact re-evaluation of at least one plane (cumulative)
  MgComm* comm = nullptr;  // NCCL communicator + exchange buffers (klsh_mg_init)
  MgLog mg;           // multi-GPU update logs (null pointers unless a sharded pass is running)
  DevBuf mg_counts, mg_mod_rows, mg_next_slot, mg_next_val, mg_splits, mg_surv;
  uint32_t* mg_keys_sorted = nullptr;  // state of the sharded pass between klsh_mg_* calls
  uint32_t* mg_rows_sorted = nullptr;
  uint64_t mg_n = 0;
  uint32_t mg_nb = 0;
  int mg_H = 0;
  int mg_stage = 0;  // 0: no sharded pass running, 1: after klsh_mg_pass_begin, 2: after klsh_mg_merge
  // escalation of the windowed merge: a bucket leaves its CTA for a cluster once it has more than
  // cta_max representatives, and the cluster for the whole grid above cluster_max
  uint32_t cta_max = 4096, cluster_max = 65536, cluster2_max = 1000000;  // measured on C2 (profiles/README.md)
  // buckets with at least this many rows skip the single-CTA stage: they start on cluster teams on a
  // second stream, so the long window chains of the few largest buckets run beside everything else
  uint32_t direct_min = 8192;
  uint32_t max_direct = 32;  // at most this many buckets take the direct pipeline per pass (about one per cluster team)
  int cluster_size = 8, cluster2_size = 16;
  int cluster_ctas_per_sm = 2;
  bool debug = false;     // KLSH_DEBUG=1
  bool cpool = false;     // KLSH_CPOOL=1: the direct pipeline's cluster teams split the screen of a window into chunks claimed by their CTAs (and by idle teams, KLSH_CPOOL_HELPERS); measured slower than the static split on C2
  uint32_t cpool_min = 65536;       // KLSH_CPOOL_MIN: representatives from which a cluster team opens its screen to helpers
  uint32_t cpool_helper_ctas = 0;   // KLSH_CPOOL_HELPERS: CTAs (in whole teams) of the direct pipeline that stay resident as helpers once they run out of buckets
  bool pool = false;      // KLSH_POOL=1: single-CTA teams with a screen pool for every bucket instead of the escalation stages and the direct pipeline (measured slower)
  uint32_t pool_min = 4096;  // KLSH_POOL_MIN: representatives from which a window's screen goes to the pool
  bool timeline = false;  // KLSH_TIMELINE=1: per pass, when each of the two merge pipelines ended (stderr)
  DevBuf dbg;
  // The speculative resolver's scan as a warp-parallel prefix computation (spec_scan_par): exact and tested, but
  // measured no faster than the scalar scan on C2 (10-20 k cycles per window against 12-37 k, and slower on the
  // single-CTA stage), so it is opt-in: KLSH_PAR_SCAN=1
  int scan_mode = 3;  // KLSH_SCAN: the speculative resolver's scan: 0 scalar replay, 1 closed-form order (KLSH_PAR_SCAN=1), 2 lean replay + parallel records, 3 = 0 on single-CTA teams, 2 on cluster teams (measured best on C2)
  bool no_spec = false;   // KLSH_NO_SPEC=1: windows are resolved by the sequential loop only
  bool merge_v1 = false;  // KLSH_MERGE_V1=1: first-generation block-per-bucket kernel (A/B checks)
  PassCounters* h_counters = nullptr;  // pinned

  PlaneSource* planes = nullptr;
  klsh_done_fn draws_done_fn = nullptr;  // see klsh_set_draws_done_callback
  void* draws_done_user = nullptr;
  bool last_iteration = false, draws_done_fired = false;  // state of the running klsh_cluster call
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// ---- error handling -------------------------------------------------------------------------
int klsh_fail(klsh_ctx* ctx, int code, const char* fmt, ...);
#define KCUDA(ctx, call)                                                                         \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return klsh_fail((ctx), KLSH_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                       __FILE__, __LINE__);                                                      \
  } while (0)
#define KTRY(expr)            \
  do {                        \
    int rc__ = (expr);        \
    if (rc__ != KLSH_OK) return rc__; \
  } while (0)

int dev_reserve(klsh_ctx* ctx, DevBuf& b, size_t bytes);

// ---- planes.cc ---------------------------------------------------------------------------------
PlaneSource* planes_new();
void planes_free(PlaneSource* p);
void planes_seed(PlaneSource* p, uint64_t seed);
void planes_callback(PlaneSource* p, klsh_plane_fn fn, void* user);
void planes_draw(PlaneSource* p, int H, int D, float* out);
void planes_tell(const PlaneSource* p, uint64_t* seed, uint64_t* drawn);
void planes_seek(PlaneSource* p, uint64_t seed, uint64_t drawn);

// ---- io.cc -------------------------------------------------------------------------------------
int io_save(const char* bin_path, int delfile, int64_t ignore_small, const float* values, int D,
            const uint64_t* id_offsets, const uint64_t* ids, uint64_t n, int id_format = 0);
int io_read_cluster(const char* bin_path, int D, uint64_t start_line, uint64_t num_lines,
                    std::vector<float>& values, std::vector<uint64_t>& id_offsets, std::vector<uint64_t>& ids, int id_format = 0);

// ---- kernels.cu (launch wrappers; all enqueue on ctx->stream) ---------------------------------
int launch_transform(klsh_ctx* ctx, const uint16_t* d_counts, const float* d_vk, uint64_t batch,
                     uint64_t* kept_out);
int launch_sign(klsh_ctx* ctx, const float* vals, int D, int ld, const uint32_t* rows, uint64_t n,
                const float* d_planes, int H, uint32_t* keys_out, uint32_t* rows_out, uint32_t key_or = 0);
int launch_sort_pairs(klsh_ctx* ctx, PassScratch& s, uint64_t n, int bits, uint32_t** keys_sorted,
                      uint32_t** rows_sorted);
int launch_bounds(klsh_ctx* ctx, PassScratch& s, const uint32_t* keys_sorted, uint64_t n);
int launch_classify(klsh_ctx* ctx, PassScratch& s, uint64_t n, int64_t nest_threshold, uint32_t b_lo, uint32_t b_hi);
int launch_rank_chains(klsh_ctx* ctx, uint64_t n, const uint32_t* d_offs, uint32_t* slot_out, int32_t* slot_row_out = nullptr);
int launch_find_splits(klsh_ctx* ctx, PassScratch& s, uint32_t nb, uint64_t n, int world, uint32_t* d_splits);
int launch_gather_mod(klsh_ctx* ctx, const uint32_t* rows, uint32_t n, float* out_vals, int32_t* out_meta);
int launch_apply_mod(klsh_ctx* ctx, const uint32_t* rows, uint32_t n, const float* in_vals, const int32_t* in_meta,
                     const uint32_t* slots, const int32_t* nvals, uint32_t n_next);
int launch_merge(klsh_ctx* ctx, PassScratch& s, uint32_t* rows_sorted, float threshold, const PassCounters& c);
int launch_merge_window(klsh_ctx* ctx, PassScratch& s, uint32_t* rows_sorted, float threshold, uint32_t n_items_host,
                        uint32_t bucket_max_host);
int launch_merge_direct(klsh_ctx* ctx, PassScratch& s, uint32_t* rows_sorted, float threshold, uint32_t n_direct_host,
                        uint32_t bucket_max_host);
bool launch_merge_uses_fallback(const klsh_ctx* ctx);
bool launch_merge_uses_pool(const klsh_ctx* ctx);
int launch_merge_pool(klsh_ctx* ctx, PassScratch& s, uint32_t* rows_sorted, float threshold);
int launch_pool_reset(klsh_ctx* ctx);
size_t merge_window_smem_bytes(int ld);
void merge_window_preload();
int launch_merge_one(klsh_ctx* ctx, PassScratch& s, uint32_t* rows_sorted, uint64_t n, float threshold);
int launch_compact(klsh_ctx* ctx, PassScratch& s, const uint32_t* rows_sorted, uint64_t n, uint32_t* out);
int launch_iota(klsh_ctx* ctx, uint32_t* out, uint64_t n, uint32_t base);
int launch_fill_tail(klsh_ctx* ctx, uint32_t* seg, uint64_t from, uint64_t to);
int launch_init_meta(klsh_ctx* ctx, uint64_t n);
int launch_cosine_pairs(klsh_ctx* ctx, const float* left, const float* right, uint64_t n, int ld, float* out);
int launch_consensus(klsh_ctx* ctx, const float* cur, int c1, const float* cand, int c2, int D, float* out);
int launch_sum_counts(klsh_ctx* ctx, const uint32_t* rows, uint64_t n, unsigned long long* total_dev);
int launch_gather_rows(klsh_ctx* ctx, const uint32_t* rows, uint64_t n, float* out_vals, int32_t* out_cnt,
                       int32_t* out_head);
