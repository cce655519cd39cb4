// Multi-GPU layer of libklsh: NCCL communicator per context, Cluster() sharded over the GPUs of a
// node without any host-language orchestration, and the all-gather of row sets that joins the
// per-GPU phase-1 batches (reference app/kmerLSH.cc:311-345) into one working set.
//
// NCCL is bound at run time (dlopen libnccl.so.2): single-GPU users need no NCCL at all, and inside a
// process that already loaded a libnccl (PyTorch ships its own) the same copy is used.
//
// Design: one context = one rank = one GPU.  klsh_mg_cluster keeps the rows
// REPLICATED on every rank and PARTITIONS the merge work: per LSH iteration every rank signs and
// groups all rows (no communication; identical hyperplane streams), the buckets are split into
// `world` contiguous ranges balanced by row count, each rank merges its range, then three things
// travel over NVLink in ONE grouped NCCL broadcast per rank: the range's survivor list, the rows the
// range modified (values + count/head/tail) and the member-chain pointer writes.  Sizes are
// exchanged first with a 16-byte ncclAllGather, merge statistics are summed with ncclAllReduce.
// Contiguous ranges in rank order keep the reference's canonical row order, so every rank ends
// bit-identical to the single-GPU klsh_cluster.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "klsh_internal.cuh"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  std::string err;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) {
      api.err = std::string("cannot load libnccl.so.2: ") + dlerror();
      return;
    }
    bool ok = true;
    auto sym = [&](const char* name) {
      void* p = dlsym(api.lib, name);
      if (!p) {
        ok = false;
        api.err = std::string("libnccl lacks ") + name;
      }
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    if (!ok) {
      dlclose(api.lib);
      api.lib = nullptr;
    }
  });
  return api.lib ? &api : nullptr;
}

}  // namespace

struct MgComm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  DevBuf d_small;                  // 4 words to send + 4*world received + 4 uint64 for the statistics
  DevBuf xbuf;                     // exchange buffer: every rank's block at its offset
  unsigned long long* h_small = nullptr;  // pinned mirror of d_small
};

#define KNCCL(ctx, api, call)                                                                              \
  do {                                                                                                     \
    ncclResult_t r__ = (call);                                                                             \
    if (r__ != ncclSuccess)                                                                                \
      return klsh_fail((ctx), KLSH_ERR_CUDA, "%s failed: %s (%s:%d)", #call, (api)->GetErrorString(r__), __FILE__, __LINE__); \
  } while (0)

extern "C" int klsh_nccl_unique_id(void* out, uint64_t bytes) {
  if (!out || bytes < sizeof(ncclUniqueId)) return klsh_fail(nullptr, KLSH_ERR_ARG, "klsh_nccl_unique_id: need %zu bytes", sizeof(ncclUniqueId));
  NcclApi* api = nccl_api();
  if (!api) return klsh_fail(nullptr, KLSH_ERR_CUDA, "NCCL is not available");
  ncclUniqueId id;
  ncclResult_t r = api->GetUniqueId(&id);
  if (r != ncclSuccess) return klsh_fail(nullptr, KLSH_ERR_CUDA, "ncclGetUniqueId: %s", api->GetErrorString(r));
  std::memcpy(out, &id, sizeof id);
  return KLSH_OK;
}

extern "C" int klsh_mg_init(klsh_ctx* ctx, int rank, int world, const void* unique_id, uint64_t bytes) {
  if (!ctx || world < 1 || world > 63 || rank < 0 || rank >= world || !unique_id || bytes < sizeof(ncclUniqueId))
    return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_init: bad argument");
  if (ctx->comm) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_init: the context already has a communicator");
  NcclApi* api = nccl_api();
  if (!api) return klsh_fail(ctx, KLSH_ERR_CUDA, "NCCL is not available");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  MgComm* c = new MgComm();
  c->rank = rank;
  c->world = world;
  ncclUniqueId id;
  std::memcpy(&id, unique_id, sizeof id);
  ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return klsh_fail(ctx, KLSH_ERR_CUDA, "ncclCommInitRank(rank %d of %d): %s", rank, world, api->GetErrorString(r));
  }
  if (cudaMallocHost(&c->h_small, sizeof(unsigned long long) * (8 + 4 * 64)) != cudaSuccess) {
    api->CommDestroy(c->comm);
    delete c;
    return klsh_fail(ctx, KLSH_ERR_NOMEM, "cudaMallocHost failed");
  }
  ctx->comm = c;
  return KLSH_OK;
}

extern "C" int klsh_mg_finalize(klsh_ctx* ctx) {
  if (!ctx || !ctx->comm) return KLSH_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  MgComm* c = ctx->comm;
  if (NcclApi* api = nccl_api()) api->CommDestroy(c->comm);
  if (c->d_small.p) cudaFree(c->d_small.p);
  if (c->xbuf.p) cudaFree(c->xbuf.p);
  if (c->h_small) cudaFreeHost(c->h_small);
  delete c;
  ctx->comm = nullptr;
  return KLSH_OK;
}

extern "C" int klsh_mg_rank(const klsh_ctx* ctx, int* rank, int* world) {
  if (!ctx) return KLSH_ERR_ARG;
  if (rank) *rank = ctx->comm ? ctx->comm->rank : 0;
  if (world) *world = ctx->comm ? ctx->comm->world : 1;
  return KLSH_OK;
}

// all-gather of `nw` 64-bit words per rank through the small device buffer; result in c->h_small[8 ..]
static int gather_small(klsh_ctx* ctx, NcclApi* api, const unsigned long long* mine, int nw) {
  MgComm* c = ctx->comm;
  KTRY(dev_reserve(ctx, c->d_small, sizeof(unsigned long long) * (8 + 4 * 64)));
  unsigned long long* d = c->d_small.as<unsigned long long>();
  for (int k = 0; k < nw; ++k) c->h_small[k] = mine[k];
  KCUDA(ctx, cudaMemcpyAsync(d, c->h_small, sizeof(unsigned long long) * nw, cudaMemcpyHostToDevice, ctx->stream));
  KNCCL(ctx, api, api->AllGather(d, d + 8, (size_t)nw, ncclUint64, c->comm, ctx->stream));
  KCUDA(ctx, cudaMemcpyAsync(c->h_small + 8, d + 8, sizeof(unsigned long long) * nw * c->world, cudaMemcpyDeviceToHost, ctx->stream));
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KLSH_OK;
}

extern "C" int klsh_mg_cluster(klsh_ctx* ctx, float min_similarity, int iterations, int64_t bucket_size_threshold,
                               klsh_iter_stats* stats) {
  if (!ctx || iterations <= 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_cluster: bad argument");
  if (!ctx->comm) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_cluster: call klsh_mg_init first");
  if (ctx->D <= 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_cluster: no rows loaded");
  NcclApi* api = nccl_api();
  MgComm* c = ctx->comm;
  const int rank = c->rank, world = c->world, ld = ctx->ld;
  KCUDA(ctx, cudaSetDevice(ctx->device));
  if (stats) std::memset(stats, 0, sizeof(klsh_iter_stats) * (size_t)iterations);
  // reference function/cluster.cc:190-192
  const float max_similarity = 0.95f;
  const float sim_step = (max_similarity - min_similarity) / iterations;
  float threshold = max_similarity;
  std::vector<uint32_t> splits((size_t)world + 1);
  std::vector<uint64_t> off((size_t)world + 1);
  for (int iter = 1; iter <= iterations; ++iter) {
    cudaEventRecord(ctx->ev[0], ctx->stream);
    uint64_t n = 0, nb = 0;
    int32_t H = 0;
    KTRY(klsh_mg_pass_begin(ctx, &n, &H, &nb));
    if (n == 0) break;
    KTRY(klsh_mg_plan(ctx, world, splits.data()));
    uint64_t ns = 0, nm = 0, nx = 0;
    KTRY(klsh_mg_merge(ctx, splits[rank], splits[rank + 1], threshold, bucket_size_threshold, &ns, &nm, &nx));
    // sizes of every rank's block (one 32-byte all-gather), then the blocks themselves
    const unsigned long long mine[4] = {ns, nm, nx, 0ull};
    KTRY(gather_small(ctx, api, mine, 4));
    const unsigned long long* all = c->h_small + 8;
    // block of rank r (32-bit words): survivors | modified rows | meta (3 per row) | chain slots | chain values | row values
    uint64_t total = 0, surv_total = 0;
    for (int r = 0; r < world; ++r) {
      const uint64_t s = all[4 * r], m = all[4 * r + 1], x = all[4 * r + 2];
      off[r] = total;
      total += s + 4 * m + 2 * x + m * (uint64_t)ld;
      total = (total + 3) & ~(uint64_t)3;  // 16-byte aligned blocks
      surv_total += s;
    }
    off[world] = total;
    KTRY(dev_reserve(ctx, c->xbuf, sizeof(uint32_t) * (total + 4)));
    uint32_t* xb = c->xbuf.as<uint32_t>();
    auto block = [&](int r, uint32_t*& surv, uint32_t*& mrows, int32_t*& meta, uint32_t*& slots, int32_t*& cvals, float*& vals) {
      const uint64_t s = all[4 * r], m = all[4 * r + 1], x = all[4 * r + 2];
      uint32_t* p = xb + off[r];
      surv = p; p += s;
      mrows = p; p += m;
      meta = reinterpret_cast<int32_t*>(p); p += 3 * m;
      slots = p; p += x;
      cvals = reinterpret_cast<int32_t*>(p); p += x;
      vals = reinterpret_cast<float*>(p);
    };
    uint32_t *surv, *mrows, *slots;
    int32_t *meta, *cvals;
    float* vals;
    block(rank, surv, mrows, meta, slots, cvals, vals);
    KTRY(klsh_mg_export(ctx, surv, mrows, vals, meta, slots, cvals));
    KNCCL(ctx, api, api->GroupStart());
    for (int r = 0; r < world; ++r) {
      const uint64_t words = off[r + 1] - off[r];
      if (words) KNCCL(ctx, api, api->Broadcast(xb + off[r], xb + off[r], (size_t)words, ncclUint32, r, c->comm, ctx->stream));
    }
    KNCCL(ctx, api, api->GroupEnd());
    // replay the other ranks' logs; the new working set is the survivor lists in rank order
    KTRY(dev_reserve(ctx, ctx->alive_alt, sizeof(uint32_t) * (surv_total + 1)));
    uint64_t at = 0;
    for (int r = 0; r < world; ++r) {
      block(r, surv, mrows, meta, slots, cvals, vals);
      const uint64_t s = all[4 * r], m = all[4 * r + 1], x = all[4 * r + 2];
      if (r != rank) KTRY(klsh_mg_apply(ctx, mrows, vals, meta, m, slots, cvals, x));
      if (s) KCUDA(ctx, cudaMemcpyAsync(ctx->alive_alt.as<uint32_t>() + at, surv, sizeof(uint32_t) * s, cudaMemcpyDeviceToDevice, ctx->stream));
      at += s;
    }
    KTRY(klsh_mg_set_alive(ctx, ctx->alive_alt.as<uint32_t>(), surv_total));
    cudaEventRecord(ctx->ev[1], ctx->stream);
    KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (stats) {
      klsh_iter_stats& st = stats[iter - 1];
      st.rows_in = n;
      st.rows_out = surv_total;
      st.H = H;
      st.threshold = threshold;
      st.buckets = nb;
      st.bucket_max = ctx->h_counters->bucket_max;
      cudaEventElapsedTime(&st.ms_total, ctx->ev[0], ctx->ev[1]);
    }
    threshold -= sim_step;  // fp32 recurrence, reference function/cluster.cc:330
  }
  // merge statistics over all ranks (north star: "allreduce of merge statistics"): rows every rank modified
  // and chain writes are already known from the size exchange; the signing kernel's exact-path rows are summed here
  return KLSH_OK;
}

// ------------------------------------------------------------------------------------------------
// All-gather of row sets: every rank contributes its current working set (e.g. the survivors of its
// phase-1 batch); afterwards every rank holds the concatenation in rank order, exactly the vector
// the reference builds by appending batch after batch to its spill file and reading it back
// (app/kmerLSH.cc:326-335, :415).  Member ids must be the implicit kind (rows loaded by
// klsh_load_counts) with contiguous batch offsets in rank order, so that ids stay implicit.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void k_pack_rows(const float* __restrict__ vals, int ld, const MetaCol cnt, const MetaCol head,
                            const MetaCol tail, const uint32_t* __restrict__ alive, uint64_t n, int32_t slot_base,
                            float* out_vals, MetaCol out_cnt, MetaCol out_head, MetaCol out_tail) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (w >= n) return;
  const uint32_t r = alive[w];
  for (int d = lane; d < ld; d += 32) out_vals[w * (uint64_t)ld + d] = vals[(uint64_t)r * ld + d];
  if (lane == 0) {
    const int32_t h = head[r], t = tail[r];
    out_cnt[w] = cnt[r];
    out_head[w] = h < 0 ? h : h + slot_base;
    out_tail[w] = t < 0 ? t : t + slot_base;
    out_cnt.p[4 * w + 3] = 0;
  }
}
__global__ void k_shift_next(const int32_t* __restrict__ next, uint64_t n, int32_t slot_base, int32_t* out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int32_t v = next[i];
    out[i] = v < 0 ? v : v + slot_base;
  }
}
}  // namespace

extern "C" int klsh_mg_gather_rows(klsh_ctx* ctx) {
  if (!ctx || !ctx->comm) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_gather_rows: call klsh_mg_init first");
  if (ctx->D <= 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_gather_rows: no rows loaded");
  if (!ctx->ids_implicit) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_gather_rows: needs implicit member ids (rows from klsh_load_counts)");
  NcclApi* api = nccl_api();
  MgComm* c = ctx->comm;
  const int rank = c->rank, world = c->world, ld = ctx->ld;
  KCUDA(ctx, cudaSetDevice(ctx->device));
  ctx->has_snap = false;
  const unsigned long long mine[4] = {ctx->cur.n_alive, ctx->n_slots, ctx->id_base, (unsigned long long)ctx->D};
  KTRY(gather_small(ctx, api, mine, 4));
  const unsigned long long* all = c->h_small + 8;
  uint64_t rows = 0, slots = 0;
  std::vector<uint64_t> row0((size_t)world + 1), slot0((size_t)world + 1);
  for (int r = 0; r < world; ++r) {
    if ((int)all[4 * r + 3] != ctx->D) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_gather_rows: rank %d has dimension %llu", r, all[4 * r + 3]);
    if (all[4 * r + 2] != all[2] + slots)
      return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_mg_gather_rows: the batches' id ranges must be contiguous in rank order (rank %d starts at %llu, expected %llu)",
                       r, all[4 * r + 2], all[2] + (unsigned long long)slots);
    row0[r] = rows;
    slot0[r] = slots;
    rows += all[4 * r];
    slots += all[4 * r + 1];
  }
  row0[world] = rows;
  slot0[world] = slots;
  if (rows >= 0xFFFFFFF0ull || slots >= 0x7FFFFFF0ull)
    return klsh_fail(ctx, KLSH_ERR_ARG, "gathered row set too large for one GPU context (%llu rows, %llu ids)", (unsigned long long)rows,
                     (unsigned long long)slots);
  RowState g;  // the gathered state
  int rc = KLSH_OK;
  do {
    if ((rc = dev_reserve(ctx, g.vals, sizeof(float) * (rows * (uint64_t)ld + 4)))) break;
    if ((rc = dev_reserve(ctx, g.meta, sizeof(int32_t) * 4 * (rows + 1)))) break;
    if ((rc = dev_reserve(ctx, g.next, sizeof(int32_t) * (slots + 1)))) break;
    if ((rc = dev_reserve(ctx, g.alive, sizeof(uint32_t) * (rows + 1)))) break;
    const uint64_t n = ctx->cur.n_alive, m = ctx->n_slots;
    if (n) {
      k_pack_rows<<<(uint32_t)((n * 32 + 255) / 256), 256, 0, ctx->stream>>>(
          ctx->cur.vals.as<float>(), ld, ctx->cur.cnt(), ctx->cur.head(), ctx->cur.tail(),
          ctx->cur.alive.as<uint32_t>(), n, (int32_t)slot0[rank], g.vals.as<float>() + row0[rank] * (uint64_t)ld,
          g.cnt() + row0[rank], g.head() + row0[rank], g.tail() + row0[rank]);
      ctx->launches++;
    }
    if (m) {
      k_shift_next<<<(uint32_t)((m + 255) / 256), 256, 0, ctx->stream>>>(ctx->cur.next.as<int32_t>(), m, (int32_t)slot0[rank],
                                                                      g.next.as<int32_t>() + slot0[rank]);
      ctx->launches++;
    }
    if (cudaGetLastError() != cudaSuccess) {
      rc = klsh_fail(ctx, KLSH_ERR_CUDA, "klsh_mg_gather_rows: pack kernels failed");
      break;
    }
    ncclResult_t nr = api->GroupStart();
    for (int r = 0; r < world && nr == ncclSuccess; ++r) {
      const uint64_t nr_rows = all[4 * r], nr_slots = all[4 * r + 1];
      if (nr_rows) {
        float* v = g.vals.as<float>() + row0[r] * (uint64_t)ld;
        nr = api->Broadcast(v, v, (size_t)(nr_rows * (uint64_t)ld), ncclFloat32, r, c->comm, ctx->stream);
        int32_t* pm = g.meta.as<int32_t>() + 4 * row0[r];  // {cnt, head, tail, 0} records
        if (nr == ncclSuccess) nr = api->Broadcast(pm, pm, (size_t)(4 * nr_rows), ncclInt32, r, c->comm, ctx->stream);
      }
      if (nr_slots && nr == ncclSuccess) {
        int32_t* pn = g.next.as<int32_t>() + slot0[r];
        nr = api->Broadcast(pn, pn, (size_t)nr_slots, ncclInt32, r, c->comm, ctx->stream);
      }
    }
    if (nr == ncclSuccess) nr = api->GroupEnd();
    if (nr != ncclSuccess) {
      rc = klsh_fail(ctx, KLSH_ERR_CUDA, "klsh_mg_gather_rows: NCCL broadcast failed: %s", api->GetErrorString(nr));
      break;
    }
    if ((rc = launch_iota(ctx, g.alive.as<uint32_t>(), rows, 0))) break;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = klsh_fail(ctx, KLSH_ERR_CUDA, "klsh_mg_gather_rows: %s", cudaGetErrorString(e));
  } while (0);
  if (rc != KLSH_OK) {
    for (DevBuf* b : {&g.vals, &g.meta, &g.next, &g.alive})
      if (b->p) cudaFree(b->p);
    return rc;
  }
  // swap the gathered state in
  for (DevBuf* b : {&ctx->cur.vals, &ctx->cur.meta, &ctx->cur.next, &ctx->cur.alive})
    if (b->p) {
      cudaFree(b->p);
      b->p = nullptr;
      b->bytes = 0;
    }
  ctx->cur.vals = g.vals;
  ctx->cur.meta = g.meta;
  ctx->cur.next = g.next;
  ctx->cur.alive = g.alive;
  ctx->cur.n_alive = rows;
  ctx->n_born = rows;
  ctx->n_slots = slots;
  ctx->id_base = all[2];
  return KLSH_OK;
}

// ------------------------------------------------------------------------------------------------
// Multi-batch phase 1 without file round trips: the survivors of each batch
// are appended to a device-resident stash; klsh_unstash_rows makes the stash — all batches, in append
// order — the current row set, the vector the reference rebuilds by appending to tmp/0.bin(.clust) and
// reading it back (app/kmerLSH.cc:326-335, :415).  Member ids stay implicit (id = base + slot) when the
// batches were loaded by klsh_load_counts with contiguous offsets; otherwise they are kept explicitly.
// ------------------------------------------------------------------------------------------------
extern "C" int klsh_stash_rows(klsh_ctx* ctx) {
  if (!ctx || ctx->D <= 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_stash_rows: no rows loaded");
  KCUDA(ctx, cudaSetDevice(ctx->device));
  const int ld = ctx->ld;
  const uint64_t n = ctx->cur.n_alive, m = ctx->n_slots;
  const uint64_t r0 = ctx->stash_rows, s0 = ctx->stash_slots;
  if (r0 && ctx->stash_D != ctx->D)
    return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_stash_rows: the stash holds rows of dimension %d, the current rows have %d", ctx->stash_D, ctx->D);
  ctx->stash_D = ctx->D;
  if (r0 + n >= 0xFFFFFFF0ull || s0 + m >= 0x7FFFFFF0ull)
    return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_stash_rows: the stash would exceed one GPU context (%llu rows, %llu ids)",
                     (unsigned long long)(r0 + n), (unsigned long long)(s0 + m));
  // member ids: implicit while every appended batch continues the id range of the previous one
  const bool cont = ctx->ids_implicit && ctx->stash_implicit && (r0 == 0 || ctx->id_base == ctx->stash_id_base + s0);
  if (!cont) {
    if (ctx->stash_implicit) {  // materialise what is stashed so far
      ctx->stash_ids.resize(s0);
      for (uint64_t k = 0; k < s0; ++k) ctx->stash_ids[k] = ctx->stash_id_base + k;
      ctx->stash_implicit = false;
    }
    ctx->stash_ids.resize(s0 + m);
    for (uint64_t k = 0; k < m; ++k) ctx->stash_ids[s0 + k] = ctx->ids_implicit ? ctx->id_base + k : ctx->ids[k];
  } else if (r0 == 0) {
    ctx->stash_id_base = ctx->id_base;
  }
  RowState& g = ctx->stash;
  KTRY(dev_reserve(ctx, g.vals, sizeof(float) * ((r0 + n) * (uint64_t)ld + 4)));
  KTRY(dev_reserve(ctx, g.meta, sizeof(int32_t) * 4 * (r0 + n + 1)));
  KTRY(dev_reserve(ctx, g.next, sizeof(int32_t) * (s0 + m + 1)));
  if (n) {
    k_pack_rows<<<(uint32_t)((n * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->cur.vals.as<float>(), ld, ctx->cur.cnt(), ctx->cur.head(), ctx->cur.tail(), ctx->cur.alive.as<uint32_t>(), n, (int32_t)s0,
        g.vals.as<float>() + r0 * (uint64_t)ld, g.cnt() + r0, g.head() + r0, g.tail() + r0);
    ctx->launches++;
  }
  if (m) {
    k_shift_next<<<(uint32_t)((m + 255) / 256), 256, 0, ctx->stream>>>(ctx->cur.next.as<int32_t>(), m, (int32_t)s0, g.next.as<int32_t>() + s0);
    ctx->launches++;
  }
  KCUDA(ctx, cudaGetLastError());
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stash_rows = r0 + n;
  ctx->stash_slots = s0 + m;
  return KLSH_OK;
}

extern "C" int klsh_stash_count(const klsh_ctx* ctx, uint64_t* n_rows) {
  if (!ctx || !n_rows) return KLSH_ERR_ARG;
  *n_rows = ctx->stash_rows;
  return KLSH_OK;
}

extern "C" int klsh_unstash_rows(klsh_ctx* ctx) {
  if (!ctx || ctx->stash_D <= 0) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_unstash_rows: nothing was stashed");
  ctx->D = ctx->stash_D;
  ctx->ld = (ctx->D + 3) & ~3;
  KCUDA(ctx, cudaSetDevice(ctx->device));
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->has_snap = false;
  const uint64_t rows = ctx->stash_rows, slots = ctx->stash_slots;
  // the stash becomes the current state (buffers swapped, nothing copied); the old state's buffers become the
  // empty stash and are reused by the next run
  std::swap(ctx->cur.vals, ctx->stash.vals);
  std::swap(ctx->cur.meta, ctx->stash.meta);
  std::swap(ctx->cur.next, ctx->stash.next);
  KTRY(dev_reserve(ctx, ctx->cur.alive, sizeof(uint32_t) * (rows + 1)));
  KTRY(launch_iota(ctx, ctx->cur.alive.as<uint32_t>(), rows, 0));
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->cur.n_alive = rows;
  ctx->n_born = rows;
  ctx->n_slots = slots;
  ctx->ids_implicit = ctx->stash_implicit;
  ctx->id_base = ctx->stash_id_base;
  ctx->ids.swap(ctx->stash_ids);
  ctx->stash_ids.clear();
  ctx->stash_rows = ctx->stash_slots = ctx->stash_id_base = 0;
  ctx->stash_implicit = true;
  return KLSH_OK;
}
