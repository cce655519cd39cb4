// C ABI (include/klsh.h) and the host-side iteration driver of the clustering loop.
//
// Device layout (DESIGN.md "Layout"): rows live in an arena vals[n_born][ld] and never move.
// The working set of an iteration is the list `alive` of row indices in the reference's canonical
// order (ascending bucket key, in-bucket order as left by p_cluster).  A merge overwrites the
// surviving representative's row in place and splices the member chains (head/tail per row, next
// per member slot).  Per iteration:
//   sign (gather rows by alive[]) -> stable radix sort of (key,row) -> bucket bounds + size classes
//   -> greedy merge per bucket (in place in the sorted row list) -> nested passes for oversized
//   buckets -> order-preserving compaction of the sorted row list = next alive[].
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <thread>

#include "klsh_internal.cuh"

static std::string g_create_error;

int klsh_fail(klsh_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  else g_create_error = buf;
  return code;
}

int dev_reserve(klsh_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.bytes) return KLSH_OK;
  size_t want = std::max(bytes, b.bytes + b.bytes / 2);
  want = (want + 255) & ~(size_t)255;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    want = (bytes + 255) & ~(size_t)255;
    e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return klsh_fail(ctx, KLSH_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    }
  }
  if (b.p) {
    // contents are preserved (grow)
    cudaError_t ec = cudaMemcpyAsync(p, b.p, b.bytes, cudaMemcpyDeviceToDevice, ctx->stream);
    if (ec == cudaSuccess) ec = cudaStreamSynchronize(ctx->stream);
    if (ec != cudaSuccess) {
      cudaFree(p);
      return klsh_fail(ctx, KLSH_ERR_CUDA, "growing a device buffer to %zu bytes failed: %s", want, cudaGetErrorString(ec));
    }
    cudaFree(b.p);
  }
  b.p = p;
  b.bytes = want;
  return KLSH_OK;
}

static void dev_free(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
}

static void free_scratch(PassScratch& s) {
  dev_free(s.keys_a); dev_free(s.keys_b); dev_free(s.rows_a); dev_free(s.rows_b);
  dev_free(s.hist); dev_free(s.blkcnt); dev_free(s.bstart);
  dev_free(s.list_small); dev_free(s.list_large); dev_free(s.list_big); dev_free(s.list_direct); dev_free(s.escb2); dev_free(s.escb3); dev_free(s.esc1); dev_free(s.esc2); dev_free(s.esc3); dev_free(s.list_nested);
  dev_free(s.pos_nrm); dev_free(s.pos_h);
  dev_free(s.planes); dev_free(s.counters);
}

static void free_state(RowState& r) {
  dev_free(r.vals); dev_free(r.meta); dev_free(r.next); dev_free(r.alive);
  r.n_alive = 0;
}

static int floor_log2_u64(uint64_t n) {  // floor(log2(n)), reference function/cluster.cc:194, :203
  int h = 0;
  while ((n >> (h + 1)) != 0) ++h;
  return h;
}

extern "C" {

int klsh_create(int device, klsh_ctx** out) {
  if (!out) return klsh_fail(nullptr, KLSH_ERR_ARG, "klsh_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    (void)cudaGetLastError();
    return klsh_fail(nullptr, KLSH_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  if (device < 0 || device >= count) return klsh_fail(nullptr, KLSH_ERR_ARG, "device %d out of range [0,%d)", device, count);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return klsh_fail(nullptr, KLSH_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return klsh_fail(nullptr, KLSH_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                     prop.major, prop.minor);
  if ((e = cudaSetDevice(device)) != cudaSuccess)
    return klsh_fail(nullptr, KLSH_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  klsh_ctx* ctx = new klsh_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    delete ctx;
    return klsh_fail(nullptr, KLSH_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  if ((e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking)) != cudaSuccess) {
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return klsh_fail(nullptr, KLSH_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  for (auto& ev : ctx->ev) cudaEventCreate(&ev);
  cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  if ((e = cudaMallocHost(&ctx->h_counters, sizeof(PassCounters) * 2)) != cudaSuccess) {
    delete ctx;
    return klsh_fail(nullptr, KLSH_ERR_NOMEM, "cudaMallocHost: %s", cudaGetErrorString(e));
  }
  ctx->planes = planes_new();
  merge_window_preload();
  // development knobs (bucket size classes of the windowed merge)
  if (const char* e = std::getenv("KLSH_DEBUG")) ctx->debug = std::atoi(e) != 0;
  if (const char* e = std::getenv("KLSH_TIMELINE")) ctx->timeline = std::atoi(e) != 0;
  if (const char* e = std::getenv("KLSH_POOL")) ctx->pool = std::atoi(e) != 0;
  if (const char* e = std::getenv("KLSH_CPOOL")) ctx->cpool = std::atoi(e) != 0;
  if (const char* e = std::getenv("KLSH_CPOOL_MIN")) ctx->cpool_min = (uint32_t)std::max(64, std::atoi(e));
  if (const char* e = std::getenv("KLSH_CPOOL_HELPERS")) ctx->cpool_helper_ctas = (uint32_t)std::max(0, std::atoi(e));
  if (const char* e = std::getenv("KLSH_POOL_MIN")) ctx->pool_min = (uint32_t)std::max(64, std::atoi(e));
  if (const char* e = std::getenv("KLSH_MERGE_V1")) ctx->merge_v1 = std::atoi(e) != 0;
  if (const char* e = std::getenv("KLSH_NO_SPEC")) ctx->no_spec = std::atoi(e) != 0;
  if (const char* e = std::getenv("KLSH_PAR_SCAN")) ctx->scan_mode = std::atoi(e) != 0 ? 1 : 0;
  if (const char* e = std::getenv("KLSH_SCAN")) ctx->scan_mode = std::max(0, std::min(3, std::atoi(e)));
  if (const char* e = std::getenv("KLSH_CTA_MAX")) ctx->cta_max = (uint32_t)std::strtoul(e, nullptr, 10);
  if (const char* e = std::getenv("KLSH_CLUSTER_MAX")) ctx->cluster_max = (uint32_t)std::strtoul(e, nullptr, 10);
  if (const char* e = std::getenv("KLSH_CLUSTER_SIZE")) ctx->cluster_size = std::atoi(e);
  if (const char* e = std::getenv("KLSH_CLUSTER2_MAX")) ctx->cluster2_max = (uint32_t)std::strtoul(e, nullptr, 10);
  if (const char* e = std::getenv("KLSH_CLUSTER2_SIZE")) ctx->cluster2_size = std::atoi(e);
  if (const char* e = std::getenv("KLSH_CLUSTER_CTAS_PER_SM")) ctx->cluster_ctas_per_sm = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("KLSH_DIRECT_MIN")) ctx->direct_min = (uint32_t)std::strtoul(e, nullptr, 10);
  if (ctx->direct_min < 2) ctx->direct_min = 2;
  if (const char* e = std::getenv("KLSH_MAX_DIRECT")) ctx->max_direct = (uint32_t)std::strtoul(e, nullptr, 10);
  if (ctx->cta_max < 1) ctx->cta_max = 1;
  if (ctx->cluster_max < ctx->cta_max) ctx->cluster_max = ctx->cta_max;
  if (ctx->cluster2_max < ctx->cluster_max) ctx->cluster2_max = ctx->cluster_max;
  *out = ctx;
  return KLSH_OK;
}

void klsh_destroy(klsh_ctx* ctx) {
  if (!ctx) return;
  klsh_mg_finalize(ctx);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->stream2);
  free_state(ctx->cur);
  free_state(ctx->snap);
  free_state(ctx->stash);
  free_scratch(ctx->top);
  free_scratch(ctx->nested);
  dev_free(ctx->lut);
  dev_free(ctx->io_a);
  dev_free(ctx->io_b);
  dev_free(ctx->alive_alt);
  dev_free(ctx->nested_out);
  dev_free(ctx->team_ctl);
  dev_free(ctx->team_ctl_b);
  dev_free(ctx->pool_ctl); dev_free(ctx->pool_pub);
  dev_free(ctx->pool_ctl_b); dev_free(ctx->pool_pub_b);
  dev_free(ctx->exp_vals); dev_free(ctx->exp_cnt); dev_free(ctx->exp_head);
  dev_free(ctx->rank_buf); dev_free(ctx->exp_offs); dev_free(ctx->exp_slots);
  if (ctx->h_slots.p) cudaFreeHost(ctx->h_slots.p);
  if (ctx->h_cnt.p) cudaFreeHost(ctx->h_cnt.p);
  if (ctx->h_head.p) cudaFreeHost(ctx->h_head.p);
  dev_free(ctx->eps_counter);
  dev_free(ctx->st_group); dev_free(ctx->st_left); dev_free(ctx->st_right); dev_free(ctx->st_counts); dev_free(ctx->st_label);
  dev_free(ctx->st_ids); dev_free(ctx->st_slot_row); dev_free(ctx->st_rec); dev_free(ctx->st_lab); dev_free(ctx->st_out_a);
  dev_free(ctx->st_out_b); dev_free(ctx->st_blk);
  dev_free(ctx->rd_table); dev_free(ctx->rd_seq); dev_free(ctx->rd_offs); dev_free(ctx->rd_rec); dev_free(ctx->rd_votes);
  dev_free(ctx->dbg);
  dev_free(ctx->mg_counts); dev_free(ctx->mg_mod_rows); dev_free(ctx->mg_next_slot); dev_free(ctx->mg_next_val);
  dev_free(ctx->mg_splits); dev_free(ctx->mg_surv);
  if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
  for (auto& ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  cudaStreamDestroy(ctx->stream2);
  cudaStreamDestroy(ctx->stream);
  planes_free(ctx->planes);
  delete ctx;
}

const char* klsh_last_error(const klsh_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
uint64_t klsh_launch_count(const klsh_ctx* ctx) { return ctx ? ctx->launches : 0; }

int klsh_set_seed(klsh_ctx* ctx, uint64_t seed) {
  if (!ctx) return KLSH_ERR_ARG;
  planes_seed(ctx->planes, seed);
  return KLSH_OK;
}

int klsh_set_plane_source(klsh_ctx* ctx, klsh_plane_fn fn, void* user) {
  if (!ctx || !fn) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_set_plane_source: NULL argument");
  planes_callback(ctx->planes, fn, user);
  return KLSH_OK;
}

int klsh_plane_tell(const klsh_ctx* ctx, uint64_t* seed, uint64_t* drawn) {
  if (!ctx || !seed || !drawn) return KLSH_ERR_ARG;
  planes_tell(ctx->planes, seed, drawn);
  return KLSH_OK;
}

int klsh_plane_seek(klsh_ctx* ctx, uint64_t seed, uint64_t drawn) {
  if (!ctx) return KLSH_ERR_ARG;
  planes_seek(ctx->planes, seed, drawn);
  return KLSH_OK;
}

int klsh_set_draws_done_callback(klsh_ctx* ctx, klsh_done_fn fn, void* user) {
  if (!ctx) return KLSH_ERR_ARG;
  ctx->draws_done_fn = fn;
  ctx->draws_done_user = user;
  return KLSH_OK;
}

int klsh_draw_table(klsh_ctx* ctx, int H, int D, float* out) {
  if (!ctx || H < 0 || D <= 0 || (!out && H > 0)) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_draw_table: bad argument");
  planes_draw(ctx->planes, H, D, out);
  return KLSH_OK;
}

int klsh_sync(klsh_ctx* ctx) {
  if (!ctx) return KLSH_ERR_ARG;
  KCUDA(ctx, cudaSetDevice(ctx->device));
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KLSH_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Row-set setup
// ------------------------------------------------------------------------------------------------
static int reserve_rows(klsh_ctx* ctx, uint64_t n_rows, uint64_t n_slots, int D) {
  if (n_rows >= 0xFFFFFFF0ull || n_slots >= 0x7FFFFFF0ull)
    return klsh_fail(ctx, KLSH_ERR_ARG, "row set too large for one GPU context (%llu rows, %llu ids)",
                     (unsigned long long)n_rows, (unsigned long long)n_slots);
  ctx->D = D;
  ctx->ld = (D + 3) & ~3;
  KTRY(dev_reserve(ctx, ctx->cur.vals, sizeof(float) * (n_rows * (uint64_t)ctx->ld + 4)));
  KTRY(dev_reserve(ctx, ctx->cur.meta, sizeof(int32_t) * 4 * (n_rows + 1)));
  KTRY(dev_reserve(ctx, ctx->cur.next, sizeof(int32_t) * (n_slots + 1)));
  KTRY(dev_reserve(ctx, ctx->cur.alive, sizeof(uint32_t) * (n_rows + 1)));
  return KLSH_OK;
}

static int ensure_lut(klsh_ctx* ctx) {
  if (ctx->lut_ready) return KLSH_OK;
  // float(log(cnt+1.0)) for every uint16 count, by the host libm the reference itself would call
  // (io/ioMatrix.cc:378)
  std::vector<float> lut(65536);
  for (uint32_t c = 0; c < 65536; ++c) lut[c] = (float)std::log((double)c + 1.0);
  KTRY(dev_reserve(ctx, ctx->lut, sizeof(float) * 65536));
  KCUDA(ctx, cudaMemcpyAsync(ctx->lut.p, lut.data(), sizeof(float) * 65536, cudaMemcpyHostToDevice, ctx->stream));
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->lut_ready = true;
  return KLSH_OK;
}

extern "C" int klsh_load_counts(klsh_ctx* ctx, const uint16_t* counts, const float* v_kmers, int D, uint64_t batch_size,
                                uint64_t batch_offset) {
  if (!ctx || !counts || !v_kmers || D <= 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_load_counts: bad argument");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  ctx->has_snap = false;
  KTRY(ensure_lut(ctx));
  KTRY(reserve_rows(ctx, batch_size, batch_size, D));
  KTRY(dev_reserve(ctx, ctx->io_a, sizeof(uint16_t) * ((uint64_t)D * batch_size + 8)));
  KTRY(dev_reserve(ctx, ctx->io_b, sizeof(float) * (D + 1)));
  KCUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, counts, sizeof(uint16_t) * (uint64_t)D * batch_size, cudaMemcpyHostToDevice,
                             ctx->stream));
  KCUDA(ctx, cudaMemcpyAsync(ctx->io_b.p, v_kmers, sizeof(float) * D, cudaMemcpyHostToDevice, ctx->stream));
  uint64_t kept = 0;
  if (batch_size) KTRY(launch_transform(ctx, ctx->io_a.as<uint16_t>(), ctx->io_b.as<float>(), batch_size, &kept));
  ctx->n_born = kept;
  ctx->n_slots = batch_size;
  ctx->ids.clear();
  ctx->ids_implicit = true;
  ctx->id_base = batch_offset;
  KTRY(launch_init_meta(ctx, ctx->n_slots));
  KTRY(launch_iota(ctx, ctx->cur.alive.as<uint32_t>(), kept, 0));
  ctx->cur.n_alive = kept;
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KLSH_OK;
}

extern "C" int klsh_set_rows(klsh_ctx* ctx, const float* values, const uint64_t* id_offsets, const uint64_t* ids,
                             uint64_t n, int D) {
  if (!ctx || D <= 0 || (n && (!values || !id_offsets || !ids)))
    return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_set_rows: bad argument");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  ctx->has_snap = false;
  const uint64_t m = n ? id_offsets[n] : 0;
  for (uint64_t r = 0; r < n; ++r)
    if (id_offsets[r + 1] < id_offsets[r] || id_offsets[r + 1] - id_offsets[r] > 0x7FFFFFFFull)
      return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_set_rows: id_offsets must be non-decreasing with at most 2^31-1 ids per row (row %llu)",
                       (unsigned long long)r);
  if (n && id_offsets[0] != 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_set_rows: id_offsets[0] must be 0");
  KTRY(reserve_rows(ctx, n, m, D));
  const int ld = ctx->ld;
  // values: pad rows to ld
  if (n) {
    if (ld == D) {
      KCUDA(ctx, cudaMemcpyAsync(ctx->cur.vals.p, values, sizeof(float) * n * (uint64_t)D, cudaMemcpyHostToDevice, ctx->stream));
    } else {
      KCUDA(ctx, cudaMemsetAsync(ctx->cur.vals.p, 0, sizeof(float) * n * (uint64_t)ld, ctx->stream));
      KCUDA(ctx, cudaMemcpy2DAsync(ctx->cur.vals.p, sizeof(float) * ld, values, sizeof(float) * D, sizeof(float) * D, n,
                                   cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  std::vector<int32_t> meta(4 * n), next(m);  // records {cnt, head, tail, 0}
  for (uint64_t r = 0; r < n; ++r) {
    uint64_t b = id_offsets[r], e = id_offsets[r + 1];
    meta[4 * r] = (int32_t)(e - b);
    meta[4 * r + 1] = e > b ? (int32_t)b : -1;
    meta[4 * r + 2] = e > b ? (int32_t)(e - 1) : -1;
    meta[4 * r + 3] = 0;
    for (uint64_t s = b; s < e; ++s) next[s] = (s + 1 < e) ? (int32_t)(s + 1) : -1;
  }
  if (n) KCUDA(ctx, cudaMemcpyAsync(ctx->cur.meta.p, meta.data(), sizeof(int32_t) * 4 * n, cudaMemcpyHostToDevice, ctx->stream));
  if (m) KCUDA(ctx, cudaMemcpyAsync(ctx->cur.next.p, next.data(), sizeof(int32_t) * m, cudaMemcpyHostToDevice, ctx->stream));
  ctx->ids.assign(ids, ids + m);
  ctx->ids_implicit = false;
  ctx->id_base = 0;
  ctx->n_born = n;
  ctx->n_slots = m;
  KTRY(launch_iota(ctx, ctx->cur.alive.as<uint32_t>(), n, 0));
  ctx->cur.n_alive = n;
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KLSH_OK;
}

extern "C" int klsh_load_cluster_file(klsh_ctx* ctx, const char* bin_path, int D, uint64_t start_line, uint64_t num_lines) {
  if (!ctx || !bin_path || D <= 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_load_cluster_file: bad argument");
  std::vector<float> values;
  std::vector<uint64_t> offs, ids;
  int rc = io_read_cluster(bin_path, D, start_line, num_lines, values, offs, ids, ctx->id_format);
  if (rc != KLSH_OK) return klsh_fail(ctx, rc, "cannot read %s(.clust)", bin_path);
  return klsh_set_rows(ctx, values.data(), offs.data(), ids.data(), offs.size() - 1, D);
}

// ------------------------------------------------------------------------------------------------
// One signing + grouping + merge pass over rows_in[0..n).  Survivors are written, in canonical
// order, to out[0..*n_out).  nest_threshold < 0 disables nesting (the pass nestedCluster itself
// runs, reference function/cluster.cc:153-159).
// ------------------------------------------------------------------------------------------------
struct PassInfo {
  uint64_t buckets = 0, bucket_max = 0, nested_calls = 0;
  float ms_sign = 0, ms_group = 0, ms_merge = 0, ms_compact = 0;
};

static int reserve_scratch(klsh_ctx* ctx, PassScratch& s, uint64_t n, int H) {
  KTRY(dev_reserve(ctx, s.keys_a, sizeof(uint32_t) * (n + 1)));
  KTRY(dev_reserve(ctx, s.keys_b, sizeof(uint32_t) * (n + 1)));
  KTRY(dev_reserve(ctx, s.rows_a, sizeof(uint32_t) * (n + 1)));
  KTRY(dev_reserve(ctx, s.rows_b, sizeof(uint32_t) * (n + 1)));
  KTRY(dev_reserve(ctx, s.planes, sizeof(float) * ((size_t)H * ctx->ld + 4)));
  KTRY(dev_reserve(ctx, s.counters, sizeof(PassCounters)));
  return KLSH_OK;
}

static int upload_planes(klsh_ctx* ctx, PassScratch& s, int H) {
  const int D = ctx->D, ld = ctx->ld;
  std::vector<float> t((size_t)H * D + 1), padded((size_t)H * ld + 1, 0.f);
  planes_draw(ctx->planes, H, D, t.data());
  for (int h = 0; h < H; ++h) std::memcpy(&padded[(size_t)h * ld], &t[(size_t)h * D], sizeof(float) * D);
  if (H) {
    KCUDA(ctx, cudaMemcpyAsync(s.planes.p, padded.data(), sizeof(float) * (size_t)H * ld, cudaMemcpyHostToDevice, ctx->stream));
    KCUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `padded` is pageable and goes out of scope
  }
  return KLSH_OK;
}

static int run_pass(klsh_ctx* ctx, PassScratch& s, const uint32_t* rows_in, uint64_t n, int H, float threshold,
                    int64_t nest_threshold, uint32_t* out, uint64_t* n_out, PassInfo* info, bool timed) {
  cudaStream_t st = ctx->stream;
  KTRY(reserve_scratch(ctx, s, n, H));
  KTRY(upload_planes(ctx, s, H));
  if (timed) cudaEventRecord(ctx->ev[0], st);
  KTRY(launch_sign(ctx, ctx->cur.vals.as<float>(), ctx->D, ctx->ld, rows_in, n, s.planes.as<float>(), H,
                   s.keys_a.as<uint32_t>(), s.rows_a.as<uint32_t>()));
  if (timed) cudaEventRecord(ctx->ev[1], st);
  uint32_t *keys_sorted, *rows_sorted;
  KTRY(launch_sort_pairs(ctx, s, n, H, &keys_sorted, &rows_sorted));
  KTRY(launch_bounds(ctx, s, keys_sorted, n));
  KTRY(launch_classify(ctx, s, n, nest_threshold, 0u, 0xFFFFFFFFu));
  PassCounters* hc = ctx->h_counters + (&s == &ctx->nested ? 1 : 0);
  KCUDA(ctx, cudaMemcpyAsync(hc, s.counters.p, sizeof(PassCounters), cudaMemcpyDeviceToHost, st));
  if (timed) cudaEventRecord(ctx->ev[2], st);
  KCUDA(ctx, cudaStreamSynchronize(st));
  const PassCounters c = *hc;
  ctx->h_counters->bucket_max = c.bucket_max;  // launch_merge sizes its spill slab from slot 0
  if (ctx->debug && c.n_buckets) {  // KLSH_DEBUG=1: bucket-size histogram (rows per power-of-two size class)
    std::vector<uint32_t> hb(c.n_buckets + 1);
    cudaMemcpy(hb.data(), s.bstart.p, sizeof(uint32_t) * (c.n_buckets + 1), cudaMemcpyDeviceToHost);
    uint64_t nb_h[33] = {0}, rows_h[33] = {0};
    for (uint32_t b = 0; b < c.n_buckets; ++b) {
      const uint32_t sz = hb[b + 1] - hb[b];
      int k = 0;
      while ((1u << k) < sz) ++k;  // size in (2^(k-1), 2^k]
      nb_h[k]++;
      rows_h[k] += sz;
    }
    fprintf(stderr, "[klsh] pass n=%llu H=%d buckets=%u small=%u large=%u big=%u direct=%u nested=%u | size<=2^k: buckets/rows",
            (unsigned long long)n, H, c.n_buckets, c.n_small, c.n_large, c.n_big, c.n_direct, c.n_nested);
    for (int k = 0; k < 33; ++k)
      if (nb_h[k]) fprintf(stderr, " %d:%llu/%llu", k, (unsigned long long)nb_h[k], (unsigned long long)rows_h[k]);
    fprintf(stderr, "\n");
  }
  // Oversized buckets (reference nestedCluster): their sizes are known now, so their hash tables are
  // drawn BEFORE the top-level merge is launched — the host never waits for the merge to learn what to
  // draw, and a multi-context driver can pass the hyperplane stream on (draws-done callback).
  uint64_t nested_calls = 0;
  std::vector<uint32_t> nb, bs;
  std::vector<int> H2;
  std::vector<float> padded;
  int Hmax = 0, pb = 0;
  uint64_t total = 0;
  bool combined = false;
  if (c.n_nested) {
    if (&s == &ctx->nested) return klsh_fail(ctx, KLSH_ERR_ARG, "internal: nested pass may not nest");
    // ascending key order (= ascending bucket index), like the reference's serial bucket loop; each
    // bucket consumes one fresh table from the plane source
    nb.resize(c.n_nested);
    KCUDA(ctx, cudaMemcpyAsync(nb.data(), s.list_nested.p, sizeof(uint32_t) * c.n_nested, cudaMemcpyDeviceToHost, st));
    KCUDA(ctx, cudaStreamSynchronize(st));
    std::sort(nb.begin(), nb.end());
    bs.resize(2 * nb.size());
    for (size_t k = 0; k < nb.size(); ++k)
      KCUDA(ctx, cudaMemcpyAsync(&bs[2 * k], s.bstart.as<uint32_t>() + nb[k], sizeof(uint32_t) * 2, cudaMemcpyDeviceToHost, st));
    KCUDA(ctx, cudaStreamSynchronize(st));
    H2.resize(nb.size());
    for (size_t k = 0; k < nb.size(); ++k) {
      const uint64_t seg_n = bs[2 * k + 1] - bs[2 * k];
      H2[k] = floor_log2_u64(seg_n);
      Hmax = std::max(Hmax, H2[k]);
      total += seg_n;
    }
    while ((1ull << pb) < nb.size()) ++pb;
    combined = Hmax + pb <= 32 && total < 0xFFFFFFF0ull;
    if (combined) {
      const int D = ctx->D, ld = ctx->ld;
      size_t plane_floats = 0;
      for (int h : H2) plane_floats += (size_t)h * ld;
      padded.assign(plane_floats + 1, 0.f);
      std::vector<float> t;
      size_t po = 0;
      for (size_t k = 0; k < nb.size(); ++k) {  // hyperplane stream order = bucket order
        t.resize((size_t)H2[k] * D + 1);
        planes_draw(ctx->planes, H2[k], D, t.data());
        for (int h = 0; h < H2[k]; ++h) std::memcpy(&padded[po + (size_t)h * ld], &t[(size_t)h * D], sizeof(float) * D);
        po += (size_t)H2[k] * ld;
      }
    }
  }
  if (&s == &ctx->top && ctx->last_iteration && (!c.n_nested || combined) && !ctx->draws_done_fired) {
    ctx->draws_done_fired = true;  // nothing after this point draws from the plane source
    if (ctx->draws_done_fn) ctx->draws_done_fn(ctx->draws_done_user);
  }
  KTRY(launch_merge(ctx, s, rows_sorted, threshold, c));
  if (c.n_nested) {
    // All oversized buckets go through ONE combined nested pass: bucket k's rows are signed with
    // bucket k's own fresh table and get the key prefix k, so one sort / bounds / merge handles every
    // sub-bucket of every nested bucket and their greedy chains run side by side.  The merged, still
    // sentinel-holed, row lists are copied back over the parent's segments; the parent's compaction
    // drops the sentinels.
    if (combined) {
      PassScratch& ns = ctx->nested;
      const int D = ctx->D, ld = ctx->ld;
      KTRY(reserve_scratch(ctx, ns, total, Hmax));
      const size_t plane_floats = padded.size() - 1;
      KTRY(dev_reserve(ctx, ns.planes, sizeof(float) * (plane_floats + 4)));
      if (plane_floats) {
        KCUDA(ctx, cudaMemcpyAsync(ns.planes.p, padded.data(), sizeof(float) * plane_floats, cudaMemcpyHostToDevice, st));
        KCUDA(ctx, cudaStreamSynchronize(st));  // `padded` is pageable
      }
      uint64_t off = 0;
      size_t po = 0;
      for (size_t k = 0; k < nb.size(); ++k) {
        const uint64_t seg_n = bs[2 * k + 1] - bs[2 * k];
        KTRY(launch_sign(ctx, ctx->cur.vals.as<float>(), D, ld, rows_sorted + bs[2 * k], seg_n, ns.planes.as<float>() + po, H2[k],
                         ns.keys_a.as<uint32_t>() + off, ns.rows_a.as<uint32_t>() + off,
                         Hmax >= 32 ? 0u : (uint32_t)(k << Hmax)));
        off += seg_n;
        po += (size_t)H2[k] * ld;
      }
      uint32_t *nkeys, *nrows;
      KTRY(launch_sort_pairs(ctx, ns, total, Hmax + pb, &nkeys, &nrows));
      KTRY(launch_bounds(ctx, ns, nkeys, total));
      KTRY(launch_classify(ctx, ns, total, -1, 0u, 0xFFFFFFFFu));
      PassCounters* hc2 = ctx->h_counters + 1;
      KCUDA(ctx, cudaMemcpyAsync(hc2, ns.counters.p, sizeof(PassCounters), cudaMemcpyDeviceToHost, st));
      KCUDA(ctx, cudaStreamSynchronize(st));
      const PassCounters c2 = *hc2;
      KTRY(launch_merge(ctx, ns, nrows, threshold, c2));
      off = 0;
      for (size_t k = 0; k < nb.size(); ++k) {
        const uint64_t seg_n = bs[2 * k + 1] - bs[2 * k];
        KCUDA(ctx, cudaMemcpyAsync(rows_sorted + bs[2 * k], nrows + off, sizeof(uint32_t) * seg_n, cudaMemcpyDeviceToDevice, st));
        off += seg_n;
      }
      nested_calls = nb.size();
    } else {
      for (size_t k = 0; k < nb.size(); ++k) {
        const uint64_t seg_n = bs[2 * k + 1] - bs[2 * k];
        uint32_t* seg = rows_sorted + bs[2 * k];
        uint64_t kept = 0;
        KTRY(reserve_scratch(ctx, ctx->nested, seg_n, H2[k]));
        KTRY(dev_reserve(ctx, ctx->nested_out, sizeof(uint32_t) * (seg_n + 2)));
        uint32_t* tmp_out = ctx->nested_out.as<uint32_t>();
        KTRY(run_pass(ctx, ctx->nested, seg, seg_n, H2[k], threshold, -1, tmp_out, &kept, nullptr, false));
        KCUDA(ctx, cudaMemcpyAsync(seg, tmp_out, sizeof(uint32_t) * kept, cudaMemcpyDeviceToDevice, st));
        KTRY(launch_fill_tail(ctx, seg, kept, seg_n));
        ++nested_calls;
      }
    }
  }
  if (timed) cudaEventRecord(ctx->ev[3], st);
  KTRY(launch_compact(ctx, s, rows_sorted, n, out));
  KCUDA(ctx, cudaMemcpyAsync(&hc->n_out, &s.counters.as<PassCounters>()->n_out, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  if (timed) cudaEventRecord(ctx->ev[4], st);
  KCUDA(ctx, cudaStreamSynchronize(st));
  *n_out = hc->n_out;
  if (info) {
    info->buckets = c.n_buckets;
    info->bucket_max = c.bucket_max;
    info->nested_calls = nested_calls;
    if (timed) {
      cudaEventElapsedTime(&info->ms_sign, ctx->ev[0], ctx->ev[1]);
      cudaEventElapsedTime(&info->ms_group, ctx->ev[1], ctx->ev[2]);
      cudaEventElapsedTime(&info->ms_merge, ctx->ev[2], ctx->ev[3]);
      cudaEventElapsedTime(&info->ms_compact, ctx->ev[3], ctx->ev[4]);
    }
  }
  return KLSH_OK;
}

extern "C" int klsh_cluster(klsh_ctx* ctx, float min_similarity, int iterations, int64_t bucket_size_threshold,
                            klsh_iter_stats* stats) {
  if (!ctx || iterations <= 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_cluster: bad argument");
  if (ctx->D <= 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_cluster: no rows loaded");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  if (stats) std::memset(stats, 0, sizeof(klsh_iter_stats) * (size_t)iterations);
  // reference function/cluster.cc:190-192
  float max_similarity = 0.95f;
  float sim_step = (max_similarity - min_similarity) / iterations;
  float threshold = max_similarity;
  KTRY(dev_reserve(ctx, ctx->eps_counter, sizeof(unsigned long long) * 4));
  KCUDA(ctx, cudaMemsetAsync(ctx->eps_counter.p, 0, sizeof(unsigned long long) * 4, ctx->stream));
  unsigned long long eps_prev = 0, pairs_prev = 0, exact_prev = 0;
  ctx->draws_done_fired = false;
  struct DoneGuard {  // the callback fires exactly once per call, whatever path the call takes
    klsh_ctx* c;
    ~DoneGuard() {
      if (!c->draws_done_fired && c->draws_done_fn) c->draws_done_fn(c->draws_done_user);
      c->draws_done_fired = true;
      c->last_iteration = false;
    }
  } done_guard{ctx};
  for (int iter = 1; iter <= iterations; ++iter) {
    const uint64_t n = ctx->cur.n_alive;
    if (n == 0) break;  // the reference takes log2(0) here (undefined); an empty set stays empty
    ctx->last_iteration = iter == iterations;
    const int H = floor_log2_u64(n);
    PassInfo info;
    uint64_t kept = 0;
    // survivors are compacted into a second list, then the lists swap roles
    KTRY(dev_reserve(ctx, ctx->alive_alt, sizeof(uint32_t) * (n + 1)));
    KTRY(run_pass(ctx, ctx->top, ctx->cur.alive.as<uint32_t>(), n, H, threshold,
                  bucket_size_threshold < 0 ? -1 : bucket_size_threshold, ctx->alive_alt.as<uint32_t>(), &kept, &info, true));
    std::swap(ctx->cur.alive, ctx->alive_alt);
    ctx->cur.n_alive = kept;
    if (stats) {
      klsh_iter_stats& s = stats[iter - 1];
      s.rows_in = n;
      s.rows_out = kept;
      s.H = H;
      s.threshold = threshold;
      s.buckets = info.buckets;
      s.bucket_max = info.bucket_max;
      s.nested_calls = info.nested_calls;
      unsigned long long cnt_now[4] = {0, 0, 0, 0};
      KCUDA(ctx, cudaMemcpyAsync(cnt_now, ctx->eps_counter.p, sizeof cnt_now, cudaMemcpyDeviceToHost, ctx->stream));
      KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
      s.eps_margin_rows = cnt_now[0] - eps_prev;  // includes the signing passes of nested buckets
      eps_prev = cnt_now[0];
      s.screen_pairs = cnt_now[2] - pairs_prev;
      pairs_prev = cnt_now[2];
      s.exact_pairs = cnt_now[3] - exact_prev;
      exact_prev = cnt_now[3];
      s.ms_sign = info.ms_sign;
      s.ms_group = info.ms_group;
      s.ms_merge = info.ms_merge;
      s.ms_compact = info.ms_compact;
      s.ms_total = info.ms_sign + info.ms_group + info.ms_merge + info.ms_compact;
    }
    threshold -= sim_step;  // fp32 recurrence, reference function/cluster.cc:330
  }
  return KLSH_OK;
}

extern "C" int klsh_sign(klsh_ctx* ctx, const float* rows, uint64_t n, int D, const float* table, int H, uint64_t* keys_out) {
  if (!ctx || D <= 0 || H < 0 || H > 32 || (n && (!rows || !keys_out)) || (H && !table))
    return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_sign: bad argument (H must be in [0,32])");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  if (!n) return KLSH_OK;
  const int ld = (D + 3) & ~3;
  DevBuf dv, dp, dk, dr;
  int rc = KLSH_OK;
  std::vector<float> padded((size_t)H * ld + 1, 0.f);
  for (int h = 0; h < H; ++h) std::memcpy(&padded[(size_t)h * ld], table + (size_t)h * D, sizeof(float) * D);
  std::vector<uint32_t> k32(n);
  do {
    if ((rc = dev_reserve(ctx, dv, sizeof(float) * (n * (uint64_t)ld + 4)))) break;
    if ((rc = dev_reserve(ctx, dp, sizeof(float) * ((size_t)H * ld + 4)))) break;
    if ((rc = dev_reserve(ctx, dk, sizeof(uint32_t) * n))) break;
    if ((rc = dev_reserve(ctx, dr, sizeof(uint32_t) * n))) break;
    cudaError_t e = cudaMemsetAsync(dv.p, 0, sizeof(float) * n * (uint64_t)ld, ctx->stream);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(dv.p, sizeof(float) * ld, rows, sizeof(float) * D, sizeof(float) * D, n, cudaMemcpyHostToDevice,
                            ctx->stream);
    if (e == cudaSuccess && H)
      e = cudaMemcpyAsync(dp.p, padded.data(), sizeof(float) * (size_t)H * ld, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
      rc = klsh_fail(ctx, KLSH_ERR_CUDA, "klsh_sign: upload failed: %s", cudaGetErrorString(e));
      break;
    }
    if ((rc = launch_sign(ctx, dv.as<float>(), D, ld, nullptr, n, dp.as<float>(), H, dk.as<uint32_t>(), dr.as<uint32_t>())))
      break;
    e = cudaMemcpyAsync(k32.data(), dk.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = klsh_fail(ctx, KLSH_ERR_CUDA, "klsh_sign: %s", cudaGetErrorString(e));
  } while (0);
  dev_free(dv); dev_free(dp); dev_free(dk); dev_free(dr);
  if (rc == KLSH_OK)
    for (uint64_t i = 0; i < n; ++i) keys_out[i] = k32[i];
  return rc;
}

extern "C" int klsh_p_cluster(klsh_ctx* ctx, float threshold) {
  if (!ctx || ctx->D <= 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_p_cluster: no rows loaded");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t n = ctx->cur.n_alive;
  if (n < 2) return KLSH_OK;
  KTRY(dev_reserve(ctx, ctx->top.counters, sizeof(PassCounters)));
  KTRY(dev_reserve(ctx, ctx->top.rows_a, sizeof(uint32_t) * (n + 1)));
  uint32_t* seg = ctx->top.rows_a.as<uint32_t>();
  KCUDA(ctx, cudaMemcpyAsync(seg, ctx->cur.alive.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  KTRY(launch_merge_one(ctx, ctx->top, seg, n, threshold));
  KTRY(launch_compact(ctx, ctx->top, seg, n, ctx->cur.alive.as<uint32_t>()));
  KCUDA(ctx, cudaMemcpyAsync(&ctx->h_counters->n_out, &ctx->top.counters.as<PassCounters>()->n_out, sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, ctx->stream));
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->cur.n_alive = ctx->h_counters->n_out;
  return KLSH_OK;
}

extern "C" int klsh_nested_cluster(klsh_ctx* ctx, float threshold) {
  if (!ctx || ctx->D <= 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_nested_cluster: no rows loaded");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t n = ctx->cur.n_alive;
  if (n == 0) return KLSH_OK;
  const int H = floor_log2_u64(n);
  uint64_t kept = 0;
  KTRY(dev_reserve(ctx, ctx->alive_alt, sizeof(uint32_t) * (n + 1)));
  KTRY(run_pass(ctx, ctx->top, ctx->cur.alive.as<uint32_t>(), n, H, threshold, -1, ctx->alive_alt.as<uint32_t>(), &kept,
                nullptr, false));
  std::swap(ctx->cur.alive, ctx->alive_alt);
  ctx->cur.n_alive = kept;
  return KLSH_OK;
}

extern "C" int klsh_cosine_distance(klsh_ctx* ctx, const float* left, const float* right, uint64_t n, int D, float* out) {
  if (!ctx || D <= 0 || (n && (!left || !right || !out))) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_cosine_distance: bad argument");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  if (!n) return KLSH_OK;
  const int ld = (D + 3) & ~3;
  const size_t rb = sizeof(float) * n * (uint64_t)ld;
  DevBuf dl, dr, dout;
  int rc = KLSH_OK;
  do {
    if ((rc = dev_reserve(ctx, dl, rb))) break;
    if ((rc = dev_reserve(ctx, dr, rb))) break;
    if ((rc = dev_reserve(ctx, dout, sizeof(float) * n))) break;
    cudaError_t e = cudaMemsetAsync(dl.p, 0, rb, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(dr.p, 0, rb, ctx->stream);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(dl.p, sizeof(float) * ld, left, sizeof(float) * D, sizeof(float) * D, n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(dr.p, sizeof(float) * ld, right, sizeof(float) * D, sizeof(float) * D, n, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
      rc = klsh_fail(ctx, KLSH_ERR_CUDA, "klsh_cosine_distance: upload failed: %s", cudaGetErrorString(e));
      break;
    }
    if ((rc = launch_cosine_pairs(ctx, dl.as<float>(), dr.as<float>(), n, ld, dout.as<float>()))) break;
    e = cudaMemcpyAsync(out, dout.p, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = klsh_fail(ctx, KLSH_ERR_CUDA, "klsh_cosine_distance: %s", cudaGetErrorString(e));
  } while (0);
  dev_free(dl); dev_free(dr); dev_free(dout);
  return rc;
}

extern "C" int klsh_set_consensus(klsh_ctx* ctx, const float* current, int64_t n_current, const float* candidate,
                                  int64_t n_candidate, int D, float* out) {
  if (!ctx || D <= 0 || !current || !candidate || !out || n_current < 0 || n_candidate < 0 || n_current + n_candidate > 0x7FFFFFFFll)
    return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_set_consensus: bad argument (member counts are the reference's `int`)");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf buf;
  int rc = KLSH_OK;
  do {
    if ((rc = dev_reserve(ctx, buf, sizeof(float) * 3 * (size_t)D))) break;
    float* d = buf.as<float>();
    cudaError_t e = cudaMemcpyAsync(d, current, sizeof(float) * D, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + D, candidate, sizeof(float) * D, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
      rc = klsh_fail(ctx, KLSH_ERR_CUDA, "klsh_set_consensus: upload failed: %s", cudaGetErrorString(e));
      break;
    }
    if ((rc = launch_consensus(ctx, d, (int)n_current, d + D, (int)n_candidate, D, d + 2 * D))) break;
    e = cudaMemcpyAsync(out, d + 2 * D, sizeof(float) * D, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = klsh_fail(ctx, KLSH_ERR_CUDA, "klsh_set_consensus: %s", cudaGetErrorString(e));
  } while (0);
  dev_free(buf);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// Rows out
// ------------------------------------------------------------------------------------------------
// Grow-only pinned host buffer (export staging: D2H at full PCIe rate, no page faults per call).
static int host_reserve(klsh_ctx* ctx, HostBuf& b, size_t bytes) {
  if (bytes <= b.bytes) return KLSH_OK;
  if (b.p) cudaFreeHost(b.p);
  b.p = nullptr;
  b.bytes = 0;
  size_t want = (bytes + (bytes >> 3) + 4095) & ~(size_t)4095;
  if (cudaMallocHost(&b.p, want) != cudaSuccess) {
    (void)cudaGetLastError();
    return klsh_fail(ctx, KLSH_ERR_NOMEM, "cudaMallocHost(%zu bytes) failed", want);
  }
  b.bytes = want;
  return KLSH_OK;
}

// Survivors in order.  values_out / ids_out may be NULL; offs_out has n_alive+1 entries.
static int export_rows(klsh_ctx* ctx, float* values_out, uint64_t* offs_out, uint64_t* ids_out) {
  const uint64_t n = ctx->cur.n_alive;
  const int D = ctx->D;
  offs_out[0] = 0;
  if (!n) return KLSH_OK;
  cudaStream_t st = ctx->stream;
  const auto t_begin = std::chrono::high_resolution_clock::now();
  KTRY(dev_reserve(ctx, ctx->exp_vals, sizeof(float) * n * (uint64_t)D));
  KTRY(dev_reserve(ctx, ctx->exp_cnt, sizeof(int32_t) * n));
  KTRY(dev_reserve(ctx, ctx->exp_head, sizeof(int32_t) * n));
  KTRY(host_reserve(ctx, ctx->h_cnt, sizeof(int32_t) * n));
  KTRY(host_reserve(ctx, ctx->h_head, sizeof(int32_t) * n));
  KTRY(launch_gather_rows(ctx, ctx->cur.alive.as<uint32_t>(), n, ctx->exp_vals.as<float>(), ctx->exp_cnt.as<int32_t>(),
                          ctx->exp_head.as<int32_t>()));
  KCUDA(ctx, cudaMemcpyAsync(ctx->h_cnt.p, ctx->exp_cnt.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  if (values_out)
    KCUDA(ctx, cudaMemcpyAsync(values_out, ctx->exp_vals.p, sizeof(float) * n * (uint64_t)D, cudaMemcpyDeviceToHost, st));
  KCUDA(ctx, cudaStreamSynchronize(st));
  const auto t_copied = std::chrono::high_resolution_clock::now();
  const int32_t* cnt = static_cast<const int32_t*>(ctx->h_cnt.p);
  for (uint64_t r = 0; r < n; ++r) offs_out[r + 1] = offs_out[r] + (uint64_t)cnt[r];
  if (ids_out) {
    const uint64_t total = offs_out[n];
    if (total >= 0xFFFFFFF0ull) return klsh_fail(ctx, KLSH_ERR_ARG, "too many member ids for one export (%llu)", (unsigned long long)total);
    // flat member order on the device (pointer jumping over the chains), then one streaming map slot -> id
    std::vector<uint32_t> offs32(n + 1);
    for (uint64_t r = 0; r <= n; ++r) offs32[r] = (uint32_t)offs_out[r];
    KTRY(dev_reserve(ctx, ctx->exp_offs, sizeof(uint32_t) * (n + 1)));
    KTRY(dev_reserve(ctx, ctx->exp_slots, sizeof(uint32_t) * (total + 1)));
    KTRY(host_reserve(ctx, ctx->h_slots, sizeof(uint32_t) * (total + 1)));
    KCUDA(ctx, cudaMemcpyAsync(ctx->exp_offs.p, offs32.data(), sizeof(uint32_t) * (n + 1), cudaMemcpyHostToDevice, st));
    KCUDA(ctx, cudaMemsetAsync(ctx->exp_slots.p, 0xFF, sizeof(uint32_t) * total, st));
    KTRY(launch_rank_chains(ctx, n, ctx->exp_offs.as<uint32_t>(), ctx->exp_slots.as<uint32_t>()));
    KCUDA(ctx, cudaMemcpyAsync(ctx->h_slots.p, ctx->exp_slots.p, sizeof(uint32_t) * total, cudaMemcpyDeviceToHost, st));
    KCUDA(ctx, cudaStreamSynchronize(st));
    const uint32_t* slots = static_cast<const uint32_t*>(ctx->h_slots.p);
    const unsigned hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    const unsigned nt = (unsigned)std::min<uint64_t>(hw, std::max<uint64_t>(1, total / 1048576));
    std::vector<uint64_t> bad(nt, UINT64_MAX);
    auto work = [&](unsigned w) {
      const uint64_t lo = total / nt * w, hi = (w + 1 == nt) ? total : total / nt * (w + 1);
      for (uint64_t i = lo; i < hi; ++i) {
        const uint32_t s = slots[i];
        if (s == 0xFFFFFFFFu) {
          if (bad[w] == UINT64_MAX) bad[w] = i;
          continue;
        }
        ids_out[i] = ctx->ids_implicit ? ctx->id_base + (uint64_t)s : ctx->ids[s];
      }
    };
    std::vector<std::thread> pool;
    for (unsigned w = 1; w < nt; ++w) pool.emplace_back(work, w);
    work(0);
    for (auto& t : pool) t.join();
    for (unsigned w = 0; w < nt; ++w)
      if (bad[w] != UINT64_MAX)
        return klsh_fail(ctx, KLSH_ERR_ARG, "internal: member chains are inconsistent (output position %llu unfilled)",
                         (unsigned long long)bad[w]);
    if (ctx->debug) {
      const auto t_end = std::chrono::high_resolution_clock::now();
      fprintf(stderr, "[klsh] export: gather+D2H %.1f ms, offsets+chain ranking+id map %.1f ms (%u threads, %llu ids)\n",
              std::chrono::duration<double, std::milli>(t_copied - t_begin).count(),
              std::chrono::duration<double, std::milli>(t_end - t_copied).count(), nt, (unsigned long long)total);
    }
  }
  return KLSH_OK;
}

// sum of the survivors' member counts, reduced on the device (no row gather, 8 bytes back)
static int count_ids(klsh_ctx* ctx, uint64_t* n_ids) {
  *n_ids = 0;
  const uint64_t n = ctx->cur.n_alive;
  if (!n) return KLSH_OK;
  KTRY(dev_reserve(ctx, ctx->eps_counter, sizeof(unsigned long long) * 4));
  unsigned long long* d_total = ctx->eps_counter.as<unsigned long long>() + 1;
  KTRY(launch_sum_counts(ctx, ctx->cur.alive.as<uint32_t>(), n, d_total));
  unsigned long long h = 0;
  KCUDA(ctx, cudaMemcpyAsync(&h, d_total, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *n_ids = h;
  return KLSH_OK;
}

extern "C" int klsh_row_count(klsh_ctx* ctx, uint64_t* n_rows, uint64_t* n_ids) {
  if (!ctx) return KLSH_ERR_ARG;
  KCUDA(ctx, cudaSetDevice(ctx->device));
  if (n_rows) *n_rows = ctx->cur.n_alive;
  if (n_ids) KTRY(count_ids(ctx, n_ids));
  return KLSH_OK;
}

extern "C" int klsh_get_rows(klsh_ctx* ctx, float* values, uint64_t* id_offsets, uint64_t* ids) {
  if (!ctx || !id_offsets) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_get_rows: bad argument");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  return export_rows(ctx, values, id_offsets, ids);
}

extern "C" int klsh_set_id_format(klsh_ctx* ctx, int format) {
  if (!ctx || (format != 0 && format != 1)) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_set_id_format: format must be 0 (text) or 1 (binary)");
  ctx->id_format = format;
  return KLSH_OK;
}

extern "C" int klsh_save(klsh_ctx* ctx, const char* bin_path, int delfile, int64_t ignore_small) {
  if (!ctx || !bin_path) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_save: bad argument");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t n = ctx->cur.n_alive;
  uint64_t n_ids = 0;
  KTRY(count_ids(ctx, &n_ids));
  std::vector<uint64_t> offs(n + 1, 0);
  std::vector<float> v(n * (uint64_t)ctx->D + 1);
  std::vector<uint64_t> idv(n_ids + 1);
  KTRY(export_rows(ctx, v.data(), offs.data(), idv.data()));
  int rc = io_save(bin_path, delfile, ignore_small, v.data(), ctx->D, offs.data(), idv.data(), n, ctx->id_format);
  if (rc != KLSH_OK) return klsh_fail(ctx, rc, "cannot write %s(.clust)", bin_path);
  return KLSH_OK;
}

// ------------------------------------------------------------------------------------------------
// Snapshot / restore (device to device)
// ------------------------------------------------------------------------------------------------
static int copy_state(klsh_ctx* ctx, RowState& dst, const RowState& src, uint64_t n_born, uint64_t n_slots) {
  const size_t vb = sizeof(float) * (n_born * (uint64_t)ctx->ld + 4);
  KTRY(dev_reserve(ctx, dst.vals, vb));
  KTRY(dev_reserve(ctx, dst.meta, sizeof(int32_t) * 4 * (n_born + 1)));
  KTRY(dev_reserve(ctx, dst.next, sizeof(int32_t) * (n_slots + 1)));
  KTRY(dev_reserve(ctx, dst.alive, sizeof(uint32_t) * (n_born + 1)));
  cudaStream_t st = ctx->stream;
  KCUDA(ctx, cudaMemcpyAsync(dst.vals.p, src.vals.p, sizeof(float) * n_born * (uint64_t)ctx->ld, cudaMemcpyDeviceToDevice, st));
  KCUDA(ctx, cudaMemcpyAsync(dst.meta.p, src.meta.p, sizeof(int32_t) * 4 * n_born, cudaMemcpyDeviceToDevice, st));
  KCUDA(ctx, cudaMemcpyAsync(dst.next.p, src.next.p, sizeof(int32_t) * n_slots, cudaMemcpyDeviceToDevice, st));
  KCUDA(ctx, cudaMemcpyAsync(dst.alive.p, src.alive.p, sizeof(uint32_t) * src.n_alive, cudaMemcpyDeviceToDevice, st));
  dst.n_alive = src.n_alive;
  KCUDA(ctx, cudaStreamSynchronize(st));
  return KLSH_OK;
}

extern "C" int klsh_snapshot(klsh_ctx* ctx) {
  if (!ctx || ctx->D <= 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_snapshot: no rows loaded");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  KTRY(copy_state(ctx, ctx->snap, ctx->cur, ctx->n_born, ctx->n_slots));
  ctx->snap_born = ctx->n_born;
  ctx->snap_slots = ctx->n_slots;
  ctx->snap_ids = ctx->ids;
  ctx->snap_id_base = ctx->id_base;
  ctx->snap_ids_implicit = ctx->ids_implicit;
  ctx->has_snap = true;
  return KLSH_OK;
}

extern "C" int klsh_restore(klsh_ctx* ctx) {
  if (!ctx || !ctx->has_snap) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_restore: no snapshot");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  KTRY(copy_state(ctx, ctx->cur, ctx->snap, ctx->snap_born, ctx->snap_slots));
  ctx->n_born = ctx->snap_born;
  ctx->n_slots = ctx->snap_slots;
  ctx->ids = ctx->snap_ids;
  ctx->id_base = ctx->snap_id_base;
  ctx->ids_implicit = ctx->snap_ids_implicit;
  return KLSH_OK;
}


// ------------------------------------------------------------------------------------------------
// Multi-GPU building blocks: one context per rank, REPLICATED row state,
// PARTITIONED merge work.  Every rank signs, sorts and bounds all rows (cheap, and it keeps every
// replica's view identical without communication); the buckets are split into `world` contiguous
// ranges balanced by row count; each rank merges its range and logs what it changed; the logs and
// the per-range survivor lists are exchanged by the caller (NCCL all-gather) and applied on every
// replica.  Contiguous ranges in bucket order keep the reference's canonical row order: the new
// working set is the concatenation of the ranks' survivor lists in rank order.
// Pointers named d_* are DEVICE pointers owned by the caller.
// ------------------------------------------------------------------------------------------------
extern "C" int klsh_row_stride(const klsh_ctx* ctx) { return ctx ? ctx->ld : 0; }

extern "C" int klsh_mg_pass_begin(klsh_ctx* ctx, uint64_t* n_rows, int32_t* H_out, uint64_t* n_buckets) {
  if (!ctx || ctx->D <= 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_pass_begin: no rows loaded");
  if (launch_merge_uses_fallback(ctx))
    return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_*: not available with the block-per-bucket fallback merge (KLSH_MERGE_V1 or D > ~280)");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t n = ctx->cur.n_alive;
  ctx->mg_n = n;
  ctx->mg_nb = 0;
  ctx->mg_stage = 1;
  if (n_rows) *n_rows = n;
  if (n == 0) {
    if (H_out) *H_out = 0;
    if (n_buckets) *n_buckets = 0;
    return KLSH_OK;
  }
  const int H = floor_log2_u64(n);
  ctx->mg_H = H;
  PassScratch& s = ctx->top;
  KTRY(reserve_scratch(ctx, s, n, H));
  KTRY(upload_planes(ctx, s, H));
  KTRY(launch_sign(ctx, ctx->cur.vals.as<float>(), ctx->D, ctx->ld, ctx->cur.alive.as<uint32_t>(), n, s.planes.as<float>(), H,
                   s.keys_a.as<uint32_t>(), s.rows_a.as<uint32_t>()));
  KTRY(launch_sort_pairs(ctx, s, n, H, &ctx->mg_keys_sorted, &ctx->mg_rows_sorted));
  KTRY(launch_bounds(ctx, s, ctx->mg_keys_sorted, n));
  KCUDA(ctx, cudaMemcpyAsync(ctx->h_counters, s.counters.p, sizeof(PassCounters), cudaMemcpyDeviceToHost, ctx->stream));
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->mg_nb = ctx->h_counters->n_buckets;
  if (H_out) *H_out = H;
  if (n_buckets) *n_buckets = ctx->mg_nb;
  return KLSH_OK;
}

extern "C" int klsh_mg_plan(klsh_ctx* ctx, int world, uint32_t* splits_out) {
  if (!ctx || world < 1 || world > 63 || !splits_out) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_plan: bad argument");
  if (ctx->mg_stage < 1) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_plan: call klsh_mg_pass_begin first");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->mg_n == 0) {
    for (int r = 0; r <= world; ++r) splits_out[r] = 0;
    return KLSH_OK;
  }
  KTRY(dev_reserve(ctx, ctx->mg_splits, sizeof(uint32_t) * 64));
  KTRY(launch_find_splits(ctx, ctx->top, ctx->mg_nb, ctx->mg_n, world, ctx->mg_splits.as<uint32_t>()));
  KCUDA(ctx, cudaMemcpyAsync(splits_out, ctx->mg_splits.p, sizeof(uint32_t) * (world + 1), cudaMemcpyDeviceToHost, ctx->stream));
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  splits_out[0] = 0;
  return KLSH_OK;
}

extern "C" int klsh_mg_merge(klsh_ctx* ctx, uint32_t b_lo, uint32_t b_hi, float threshold, int64_t bucket_size_threshold,
                             uint64_t* n_surv, uint64_t* n_mod, uint64_t* n_next) {
  if (!ctx || b_lo > b_hi) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_merge: bad argument");
  if (ctx->mg_stage != 1) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_merge: call klsh_mg_pass_begin first (once per pass)");
  if (ctx->mg_n && b_hi > ctx->mg_nb)
    return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_merge: bucket range [%u, %u) exceeds the pass's %u buckets", b_lo, b_hi, ctx->mg_nb);
  ctx->mg_stage = 2;
  KCUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t n = ctx->mg_n;
  if (n_surv) *n_surv = 0;
  if (n_mod) *n_mod = 0;
  if (n_next) *n_next = 0;
  if (n == 0) return KLSH_OK;
  cudaStream_t st = ctx->stream;
  PassScratch& s = ctx->top;
  uint32_t* rows_sorted = ctx->mg_rows_sorted;
  // logs: a row is logged once per window flush, a chain pointer once per merge -> both bounded by n
  KTRY(dev_reserve(ctx, ctx->mg_counts, sizeof(uint32_t) * 4));
  KTRY(dev_reserve(ctx, ctx->mg_mod_rows, sizeof(uint32_t) * (n + 1)));
  KTRY(dev_reserve(ctx, ctx->mg_next_slot, sizeof(uint32_t) * (n + 1)));
  KTRY(dev_reserve(ctx, ctx->mg_next_val, sizeof(int32_t) * (n + 1)));
  KTRY(dev_reserve(ctx, ctx->mg_surv, sizeof(uint32_t) * (n + 1)));
  KCUDA(ctx, cudaMemsetAsync(ctx->mg_counts.p, 0, sizeof(uint32_t) * 4, st));
  ctx->mg.counts = ctx->mg_counts.as<uint32_t>();
  ctx->mg.mod_rows = ctx->mg_mod_rows.as<uint32_t>();
  ctx->mg.next_slot = ctx->mg_next_slot.as<uint32_t>();
  ctx->mg.next_val = ctx->mg_next_val.as<int32_t>();
  struct Reset {
    klsh_ctx* c;
    ~Reset() { c->mg = MgLog(); }
  } reset{ctx};

  const int64_t nest = bucket_size_threshold < 0 ? -1 : bucket_size_threshold;
  KTRY(launch_classify(ctx, s, n, nest, b_lo, b_hi));
  KCUDA(ctx, cudaMemcpyAsync(ctx->h_counters, s.counters.p, sizeof(PassCounters), cudaMemcpyDeviceToHost, st));
  KCUDA(ctx, cudaStreamSynchronize(st));
  const PassCounters c = *ctx->h_counters;
  KTRY(launch_merge(ctx, s, rows_sorted, threshold, c));
  if (c.n_nested) {
    // every rank draws the table of every oversized bucket, in bucket order, to stay in step with
    // the single-process hyperplane stream; only the owner runs the nested pass
    std::vector<uint32_t> nb(c.n_nested);
    KCUDA(ctx, cudaMemcpyAsync(nb.data(), s.list_nested.p, sizeof(uint32_t) * c.n_nested, cudaMemcpyDeviceToHost, st));
    KCUDA(ctx, cudaStreamSynchronize(st));
    std::sort(nb.begin(), nb.end());
    uint32_t bs[2];
    std::vector<float> discard;
    for (uint32_t b : nb) {
      KCUDA(ctx, cudaMemcpyAsync(bs, s.bstart.as<uint32_t>() + b, sizeof(uint32_t) * 2, cudaMemcpyDeviceToHost, st));
      KCUDA(ctx, cudaStreamSynchronize(st));
      const uint64_t seg_n = bs[1] - bs[0];
      const int H2 = floor_log2_u64(seg_n);
      if (b < b_lo || b >= b_hi) {
        discard.resize((size_t)H2 * ctx->D + 1);
        planes_draw(ctx->planes, H2, ctx->D, discard.data());
        continue;
      }
      uint32_t* seg = rows_sorted + bs[0];
      uint64_t kept = 0;
      KTRY(reserve_scratch(ctx, ctx->nested, seg_n, H2));
      KTRY(dev_reserve(ctx, ctx->nested_out, sizeof(uint32_t) * (seg_n + 2)));
      uint32_t* tmp_out = ctx->nested_out.as<uint32_t>();
      KTRY(run_pass(ctx, ctx->nested, seg, seg_n, H2, threshold, -1, tmp_out, &kept, nullptr, false));
      KCUDA(ctx, cudaMemcpyAsync(seg, tmp_out, sizeof(uint32_t) * kept, cudaMemcpyDeviceToDevice, st));
      KTRY(launch_fill_tail(ctx, seg, kept, seg_n));
    }
  }
  // survivors of my range, in order
  uint32_t pr[2] = {0, 0};
  KCUDA(ctx, cudaMemcpyAsync(&pr[0], s.bstart.as<uint32_t>() + b_lo, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  KCUDA(ctx, cudaMemcpyAsync(&pr[1], s.bstart.as<uint32_t>() + b_hi, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  KCUDA(ctx, cudaStreamSynchronize(st));
  if (b_lo >= ctx->mg_nb) pr[0] = (uint32_t)n;
  if (b_hi >= ctx->mg_nb) pr[1] = (uint32_t)n;
  uint64_t surv = 0;
  if (pr[1] > pr[0]) {
    KTRY(launch_compact(ctx, s, rows_sorted + pr[0], pr[1] - pr[0], ctx->mg_surv.as<uint32_t>()));
    KCUDA(ctx, cudaMemcpyAsync(&ctx->h_counters->n_out, &s.counters.as<PassCounters>()->n_out, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    KCUDA(ctx, cudaStreamSynchronize(st));
    surv = ctx->h_counters->n_out;
  }
  uint32_t counts[2];
  KCUDA(ctx, cudaMemcpyAsync(counts, ctx->mg_counts.p, sizeof counts, cudaMemcpyDeviceToHost, st));
  KCUDA(ctx, cudaStreamSynchronize(st));
  if (n_surv) *n_surv = surv;
  if (n_mod) *n_mod = counts[0];
  if (n_next) *n_next = counts[1];
  ctx->h_counters[1].n_out = (uint32_t)surv;  // remembered for klsh_mg_export
  ctx->h_counters[1].n_small = counts[0];
  ctx->h_counters[1].n_large = counts[1];
  return KLSH_OK;
}

extern "C" int klsh_mg_export(klsh_ctx* ctx, uint32_t* d_surv, uint32_t* d_mod_rows, float* d_mod_vals, int32_t* d_mod_meta,
                              uint32_t* d_next_slot, int32_t* d_next_val) {
  if (!ctx) return KLSH_ERR_ARG;
  if (ctx->mg_stage != 2) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_export: call klsh_mg_merge first");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint32_t surv = ctx->h_counters[1].n_out, nmod = ctx->h_counters[1].n_small, nnext = ctx->h_counters[1].n_large;
  if ((surv && !d_surv) || (nmod && (!d_mod_rows || !d_mod_vals || !d_mod_meta)) || (nnext && (!d_next_slot || !d_next_val)))
    return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_export: NULL output buffer for a non-empty log");
  if (surv) KCUDA(ctx, cudaMemcpyAsync(d_surv, ctx->mg_surv.p, sizeof(uint32_t) * surv, cudaMemcpyDeviceToDevice, st));
  if (nmod) {
    KCUDA(ctx, cudaMemcpyAsync(d_mod_rows, ctx->mg_mod_rows.p, sizeof(uint32_t) * nmod, cudaMemcpyDeviceToDevice, st));
    KTRY(launch_gather_mod(ctx, ctx->mg_mod_rows.as<uint32_t>(), nmod, d_mod_vals, d_mod_meta));
  }
  if (nnext) {
    KCUDA(ctx, cudaMemcpyAsync(d_next_slot, ctx->mg_next_slot.p, sizeof(uint32_t) * nnext, cudaMemcpyDeviceToDevice, st));
    KCUDA(ctx, cudaMemcpyAsync(d_next_val, ctx->mg_next_val.p, sizeof(int32_t) * nnext, cudaMemcpyDeviceToDevice, st));
  }
  KCUDA(ctx, cudaStreamSynchronize(st));
  return KLSH_OK;
}

extern "C" int klsh_mg_apply(klsh_ctx* ctx, const uint32_t* d_mod_rows, const float* d_mod_vals, const int32_t* d_mod_meta,
                             uint64_t n_mod, const uint32_t* d_next_slot, const int32_t* d_next_val, uint64_t n_next) {
  if (!ctx) return KLSH_ERR_ARG;
  if (ctx->mg_stage < 1) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_apply: no sharded pass is running");
  if (n_mod > ctx->n_born || n_next > ctx->n_slots || (n_mod && (!d_mod_rows || !d_mod_vals || !d_mod_meta)) ||
      (n_next && (!d_next_slot || !d_next_val)))
    return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_apply: log sizes (%llu rows, %llu chain writes) do not fit the row set, or NULL buffers",
                     (unsigned long long)n_mod, (unsigned long long)n_next);
  KCUDA(ctx, cudaSetDevice(ctx->device));
  KTRY(launch_apply_mod(ctx, d_mod_rows, (uint32_t)n_mod, d_mod_vals, d_mod_meta, d_next_slot, d_next_val, (uint32_t)n_next));
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KLSH_OK;
}

extern "C" int klsh_mg_set_alive(klsh_ctx* ctx, const uint32_t* d_alive, uint64_t n) {
  if (!ctx || n > ctx->n_born || (n && !d_alive)) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_set_alive: bad argument");
  ctx->mg_stage = 0;
  KCUDA(ctx, cudaSetDevice(ctx->device));
  KTRY(dev_reserve(ctx, ctx->cur.alive, sizeof(uint32_t) * (n + 1)));
  if (n) KCUDA(ctx, cudaMemcpyAsync(ctx->cur.alive.p, d_alive, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->cur.n_alive = n;
  return KLSH_OK;
}
