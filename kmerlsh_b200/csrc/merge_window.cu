// Windowed greedy merge for buckets of more than KLSH_SMALL_MAX rows.
//
// Reference semantics (p_cluster, function/cluster.cc:56-87): candidates are examined one at a
// time; the candidate at position i is compared with the representatives at positions 0..i-1 in
// order and merged into the FIRST one whose cosine similarity reaches the threshold (the merged
// representative becomes the count-weighted consensus, AB::SetConsensus funcAB.cc:49-71; the tail
// position is swapped into i and examined next); otherwise it becomes representative i.
//
// The work is candidates x representatives x D, strictly sequential in the reference.  Here a
// TEAM processes a WINDOW of up to 64 upcoming candidates at once:
//   parallel phase : the window (fp16, unit norm, A fragments in registers) is screened against all
//                    representatives on the tensor cores (mma.sync; the representatives' fp16 copies are
//                    streamed in fragment order through a cp.async ring); pairs that pass are re-tested
//                    with the reference's exact fp32 arithmetic, recording per candidate the first
//                    matching representative as of the window start (atomicMin), plus the candidate x
//                    candidate match bits;
//   resolver       : one warp replays the reference's sequential order over the window using
//                    those results; representatives modified inside the window live in a small
//                    "dirty" cache and are re-compared exactly with their current values; when a
//                    decision cannot be proven from what was precomputed the window is truncated
//                    and the candidate is re-examined in the next window.
// Teams escalate with the amount of compare work: every bucket starts on ONE CTA; when its
// representative count passes max_reps the bucket's state (it lives entirely in global memory) is
// handed to a thread-block CLUSTER, and from there to the whole cooperative GRID.  Merge-heavy
// buckets (few representatives, long dependent chains) therefore stay on one SM each and run side
// by side, while compare-heavy buckets get more SMs as their representative set grows.
// Every decision and every centroid is bit-identical to the sequential algorithm (DESIGN.md
// "Windowed merge: why it is exact").
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include <chrono>
#include <cstdio>

#include "exact_math.cuh"
#include "klsh_internal.cuh"

#ifndef KLSH_CTA_THREADS
#define KLSH_CTA_THREADS 128
#endif

namespace cg = cooperative_groups;

namespace {

constexpr int kW = 64;    // window capacity
constexpr int kKD = 32;   // dirty-cache entries
// Threads per CTA and CTAs per SM by team kind.  Single-CTA teams are paced by per-window latencies
// (staging, the resolver warp), so more, smaller CTAs per SM raise throughput; cluster and grid teams
// carry the compare-heavy buckets and keep the wider CTAs.
template <int TEAM>
struct Shape {
  static constexpr int kMT = TEAM == 0 ? KLSH_CTA_THREADS : 256;
  static constexpr int kCtasPerSm = TEAM == 0 ? (512 / KLSH_CTA_THREADS) : 2;
};
constexpr int kWbMax = 62;
constexpr int kSurvCap = 1023;  // deferred screen survivors per CTA and window
constexpr int kRing = 8;  // representative groups in flight per warp in the screen (cp.async ring)
constexpr uint32_t kInf = 0x7fffffffu;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// ---- tensor-core prefilter ----------------------------------------------------------------------------
// The parallel phase is a contraction (window candidates x representatives x D).  Its exact form is
// bound by FP32 issue, so it is screened on the tensor cores first: rows are scaled to unit norm,
// rounded to fp16 and multiplied with mma.sync (fp32 accumulate).  With unit-norm inputs the result
// approximates the cosine with absolute error < 1.1e-3 (fp16 input rounding 2*2^-11 per product,
// Cauchy-Schwarz; accumulation and the reference's own fp32 rounding are orders of magnitude below).
// A match needs cos >= thr - 2^-21, so only pairs with approx >= thr - 2e-3 can match: those few are
// re-evaluated with the reference's exact arithmetic and nothing else decides anything.  Zero, NaN or
// infinite norms give NaN/inf fragments, which never compare "below" and therefore reach the exact test.
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Fragment-order fp16 copy of a unit-norm row.  Chunk c (16 bytes) holds what lane tg = c & 3 of an
// mma B fragment needs for the k-steps 2*(c>>2) and 2*(c>>2)+1: halves {k0+2tg, k0+2tg+1, k0+8+2tg, ...}.
__device__ __forceinline__ float unit_scale(float x, float nrm) { return __fdividef(x, nrm); }
__device__ __forceinline__ uint4 h16_chunk_from_row(const float* row, int ld, float nrm, int c) {
  const int tg = c & 3, k0 = (c >> 2) * 32;
  uint32_t w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int off = k0 + (q >> 1) * 16 + (q & 1) * 8 + tg * 2;
    const float x0 = off < ld ? unit_scale(row[off], nrm) : 0.f;
    const float x1 = off + 1 < ld ? unit_scale(row[off + 1], nrm) : 0.f;
    w[q] = pack_half2(x0, x1);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 h16_chunk_from_htile(const __half* hrow, int c) {
  const int tg = c & 3, k0 = (c >> 2) * 32;
  uint32_t w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) w[q] = *reinterpret_cast<const uint32_t*>(hrow + k0 + (q >> 1) * 16 + (q & 1) * 8 + tg * 2);
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- teams -----------------------------------------------------------------------------------------
template <int TEAM>
struct Team;
template <>
struct Team<0> {  // one CTA
  static __device__ __forceinline__ void sync() { __syncthreads(); }
  static __device__ __forceinline__ uint32_t rank() { return 0; }
  static __device__ __forceinline__ uint32_t ncta() { return 1; }
  static __device__ __forceinline__ uint32_t id() { return blockIdx.x; }
};
template <>
struct Team<1> {  // one thread-block cluster
  static __device__ __forceinline__ void sync() { cg::this_cluster().sync(); }
  static __device__ __forceinline__ uint32_t rank() { return cg::this_cluster().block_rank(); }
  static __device__ __forceinline__ uint32_t ncta() { return cg::this_cluster().num_blocks(); }
  static __device__ __forceinline__ uint32_t id() { return blockIdx.x / cg::this_cluster().num_blocks(); }
};
template <>
struct Team<2> {  // the whole cooperative grid
  static __device__ __forceinline__ void sync() { cg::this_grid().sync(); }
  static __device__ __forceinline__ uint32_t rank() { return blockIdx.x; }
  static __device__ __forceinline__ uint32_t ncta() { return gridDim.x; }
  static __device__ __forceinline__ uint32_t id() { return 0; }
};

struct TeamCtl {  // global memory, one per team
  uint32_t f[kW];  // per window candidate: first matching old representative (position), kInf if none
  uint32_t i, size, wb, work;
  uint32_t mode;  // 1: the next window is resolved by the sequential loop (after a mispredicted speculative window)
  uint32_t big;   // cluster pool: this team's bucket is counted in PoolCtl::big_active
  uint32_t pad[2];
};

// ---- screen pool (single-CTA teams) -------------------------------------------------------------------
// A bucket's windows are a sequential chain led by ONE CTA, but the screen of a window against the bucket's
// representatives is independent work.  With the pool a leader whose bucket has passed pool_min representatives
// publishes its window (fp16 A fragments, row indices, norms) and opens its screen as CHUNKS of kPoolChunk
// representatives.  A chunk is claimed with one atomicAdd on the leader's own counter — by the leader itself
// (which screens it from its shared-memory copy) or by any CTA of the launch with nothing better to do: leaders
// waiting for their last chunks and CTAs that ran out of buckets.  Helpers find open windows on a board of one
// word per leader.  A helper's result is the per-candidate first exact match in its chunk, folded into the
// leader's block with atomicMin; the leader continues when its done counter reaches the number of chunks.
// Nothing ever waits for a helper to show up: the leader claims chunks until none is left, so the board is a hint.
constexpr int kPoolChunk = 2048;
constexpr uint32_t kPoolCountBits = 20;  // claim word = window epoch << 20 | chunks claimed so far
struct PoolCtl {
  uint32_t open;       // leaders with unclaimed chunks (hint: helpers scan the board only when it is non-zero)
  uint32_t finished;   // teams of this launch that ran out of buckets
  uint32_t helpers;    // cluster teams that stayed on as helpers after running out of buckets
  uint32_t big_active; // cluster teams whose current bucket is on its way to pooled windows (pool_min / 4 representatives)
  uint32_t pad[4];
};
struct PoolPub {  // one per leader CTA
  uint4 af[kW * 8];      // the window's fp16 copy in A-fragment order: [(row tile * KS16 + k step) * 32 + lane]
  uint32_t ridx[kW];
  float cnorm[kW];
  uint32_t f[kW];        // first exact match per candidate over the chunks helpers have served
  uint32_t W, st, i0;
  uint32_t meta;         // window epoch << 20 | number of chunks: ONE word, so a helper validates its claim against a snapshot
  uint32_t next;         // claim word: window epoch << 20 | chunks claimed so far
  uint32_t done;         // chunks helpers have finished
  uint32_t pad[2];
};

struct MergeArgs {
  float* vals;
  int D, ld;
  MetaCol cnt, head, tail;
  int32_t* next;
  uint32_t* rows_sorted;
  const uint32_t* bstart;
  // work items are triples {bucket, i, size}; i == 0 means "not started".  Two lists, walked in order.
  const uint32_t* list_a;
  const uint32_t* n_a;
  const uint32_t* list_b;
  const uint32_t* n_b;
  uint32_t* cursor;
  // buckets whose representative count passes max_reps are appended here for the next, larger team
  uint32_t* esc_list;
  uint32_t* esc_count;
  uint32_t max_reps;
  float* pos_nrm;  // norm of the representative at each sorted position (scratch, N floats)
  // unit-norm fp16 copy of the representative at each sorted position, in tensor-core fragment order:
  // the screen streams it with one coalesced 16-byte load per lane instead of chasing
  // row index -> row and converting on the fly
  uint4* pos_h;
  TeamCtl* ctl;
  // screen pool (nullptr: off)
  PoolCtl* pool;
  uint32_t* pool_board;  // [grid] window epoch + 1 while the leader has unclaimed chunks, else 0 (hint)
  PoolPub* pool_pub;
  uint32_t pool_min;
  uint32_t pool_n;        // board entries = teams of the launch
  uint32_t pool_helpers;  // cluster teams: how many idle teams stay resident as helpers (the others exit and free their SMs)
  // a third work list, walked before list_a (the pass's largest buckets when the pool is on)
  const uint32_t* list_0;
  const uint32_t* n_0;
  MgLog mg;
  unsigned long long* work;  // [2] device counters: pairs screened on the tensor cores, pairs re-tested exactly
  int scan_mode;  // speculative scan, KLSH_SCAN: 0 scalar replay with records, 1 closed-form order, 2 lean replay + parallel records, 3 (default) 0 on single-CTA teams and 2 on cluster teams
  int no_spec;   // KLSH_NO_SPEC=1: sequential resolution only (A/B checks)
  unsigned long long* dbg;  // [32] 26: speculative windows, 27: ... cut short; 0..7: windows, candidates, merges, undecidable, cache_full, back_exhausted, accepted, escalated; 8..13: leader cycles in stage/parallel/sync1/prefetch/decide/flush+sync2; 18..21: staging detail
  float threshold;
};

struct Smem {
  float* tile;   // [kW][ts]
  float* dvals;  // [kKD][ts]
  float* pre;    // [kW][ts]  row of each candidate's first-match representative, prefetched for the resolver
  float* cnorm;  // [kW]
  float* dnorm;  // [kKD]
  __half* htile; // [kW][hs] unit-norm fp16 copy of the window (tensor-core prefilter)
  uint32_t* surv;  // [kSurvCap + 1] deferred (representative << 6 | candidate) pairs that passed the screen; [kSurvCap] = count
  uint4* ring;   // per warp: kRing slots x 32 lanes x (width/32) 16-byte chunks of representative fp16 copies in flight
  int hs;
  uint32_t* ridx;   // [kW]
  int32_t* ccnt;    // [kW]
  int32_t* chead;   // [kW]
  int32_t* ctail;   // [kW]
  uint32_t* s_f;    // [kW]
  uint32_t* pair;   // [kW][2]
  uint32_t* acc;    // [kW]
  uint32_t* dpos;   // [kKD]
  int32_t* dcnt;    // [kKD]
  int32_t* dhead;   // [kKD]
  int32_t* dtail;   // [kKD]
  uint32_t* dridx;  // [kKD]
  uint32_t* pridx;  // [kW]
  int32_t* pcnt;    // [kW]
  int32_t* phead;   // [kW]
  int32_t* ptail;   // [kW]
  int32_t* mprev;   // [kW] merge log: previous candidate merged into the same entry (-1: none)
  int32_t* ment;    // [kW] merge log: dirty entry the candidate was merged into (-1: not merged)
  // decision warp <-> mask helper warps (seqlock per dirty entry, see resolve_decide / mask_helpers)
  uint32_t* dver;   // [kKD] version of the entry's value: odd while it is being written
  uint32_t* mver;   // [kKD] version the published match mask belongs to
  uint32_t* dmlo;   // [kKD] published match mask, window candidates 0..31
  uint32_t* dmhi;   // [kKD] ... candidates 32..63
  int32_t* dlast;   // [kKD] last candidate merged into the entry that carried ids (-1: none)
  uint32_t* ro;     // [16] what the decision loop hands to the flush (RO_*)
  uint32_t* w_f;    // [kW] pool: first exact match per candidate of the task being served
  uint32_t* ptask;  // [8] pool: broadcast slots {claimed?, leader, chunk, board value / scan result, own claim, done, epoch}
  int ts;
};

enum { RO_A = 0, RO_ND, RO_I, RO_SIZE, RO_FROM_BACK, RO_BI, RO_MERGES, RO_BACK_EXH, RO_UNDEC, RO_FULL, RO_DONE, RO_EXAMINED, RO_WORDS = 16 };

__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void st_vol(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }

// width (halfs) of the fp16 window copy: the k extent the kernel variant for this ld multiplies over
__host__ __device__ inline int tc_width(int ld) { return ld <= 32 ? 32 : (ld <= 64 ? 64 : ((ld + 31) & ~31)); }

// floats per staged row in shared memory: rows are zero-padded to the kernel variant's compile-time width
// (32 or 64) so that the decision warp's chains have compile-time trip counts; adding the padded +0
// products leaves every reference sum bit-identical
__host__ __device__ inline int row_width(int ld) { return ld <= 32 ? 32 : (ld <= 64 ? 64 : ld); }

__host__ __device__ inline size_t spec_bytes_for(int ld);
__host__ __device__ inline size_t ring_bytes_for(int ld, int threads) {
  if (ld > 64) return 0;
  const size_t ring = (size_t)(threads / 32) * kRing * 32 * (tc_width(ld) / 32) * 16;
  const size_t spec = spec_bytes_for(ld);  // the speculative resolver's scratch lives in the idle ring
  return ring > spec ? ring : ((spec + 15) & ~(size_t)15);
}
__host__ __device__ inline size_t smem_bytes_for(int ld, int threads) {
  const int hs = tc_width(ld) + 8;
  return sizeof(float) * ((size_t)(2 * kW + kKD) * (row_width(ld) + 4) + kW + kKD) + sizeof(uint32_t) * (kW * 15 + kKD * 10 + RO_WORDS + 8) + 64 +
         sizeof(__half) * (size_t)kW * hs + 16 + sizeof(uint32_t) * (kSurvCap + 1) + ring_bytes_for(ld, threads);
}

__device__ __forceinline__ void carve(Smem& s, float* base, int ld) {
  s.ts = row_width(ld) + 4;
  s.tile = base;
  s.dvals = s.tile + (size_t)kW * s.ts;
  s.pre = s.dvals + (size_t)kKD * s.ts;
  s.cnorm = s.pre + (size_t)kW * s.ts;
  s.dnorm = s.cnorm + kW;
  uint32_t* u = reinterpret_cast<uint32_t*>(s.dnorm + kKD);
  s.ridx = u; u += kW;
  s.ccnt = reinterpret_cast<int32_t*>(u); u += kW;
  s.chead = reinterpret_cast<int32_t*>(u); u += kW;
  s.ctail = reinterpret_cast<int32_t*>(u); u += kW;
  s.s_f = u; u += kW;
  s.pair = u; u += 2 * kW;
  s.acc = u; u += kW;
  s.dpos = u; u += kKD;
  s.dcnt = reinterpret_cast<int32_t*>(u); u += kKD;
  s.dhead = reinterpret_cast<int32_t*>(u); u += kKD;
  s.dtail = reinterpret_cast<int32_t*>(u); u += kKD;
  s.dridx = u; u += kKD;
  s.pridx = u; u += kW;
  s.pcnt = reinterpret_cast<int32_t*>(u); u += kW;
  s.phead = reinterpret_cast<int32_t*>(u); u += kW;
  s.ptail = reinterpret_cast<int32_t*>(u); u += kW;
  s.mprev = reinterpret_cast<int32_t*>(u); u += kW;
  s.ment = reinterpret_cast<int32_t*>(u); u += kW;
  s.dver = u; u += kKD;
  s.mver = u; u += kKD;
  s.dmlo = u; u += kKD;
  s.dmhi = u; u += kKD;
  s.dlast = reinterpret_cast<int32_t*>(u); u += kKD;
  s.ro = u; u += RO_WORDS;
  s.w_f = u; u += kW;
  s.ptask = u; u += 8;
  s.hs = tc_width(ld) + 8;  // +8 halfs: rows 16 bytes apart modulo 128 -> conflict-free fragment loads
  s.htile = reinterpret_cast<__half*>(u + 4);
  s.surv = reinterpret_cast<uint32_t*>(s.htile + (size_t)kW * s.hs);  // kW*hs halves: a multiple of 16 bytes
  s.ring = reinterpret_cast<uint4*>(s.surv + kSurvCap + 1);
}

// Exact evaluation of one (candidate t, representative) pair that survived a prefilter.
template <bool SM>
__device__ __forceinline__ bool exact_pair(const Smem& s, int t, const float* rowp, float rn, int nq, float threshold) {
  const float4* c4 = reinterpret_cast<const float4*>(s.tile + (size_t)t * s.ts);
  float dx = 0.f;
  for (int q = 0; q < nq; ++q) {
    const float4 x = c4[q];
    const float4 y = SM ? reinterpret_cast<const float4*>(rowp)[q] : __ldcg(reinterpret_cast<const float4*>(rowp) + q);
    dx = __fadd_rn(dx, __fmul_rn(x.x, y.x));
    dx = __fadd_rn(dx, __fmul_rn(x.y, y.y));
    dx = __fadd_rn(dx, __fmul_rn(x.z, y.z));
    dx = __fadd_rn(dx, __fmul_rn(x.w, y.w));
  }
  return cos_match(dx, s.cnorm[t], rn, threshold);
}

// The same test served for another CTA's window (pool): the candidate's row and norm come from global memory
// (the leader staged its tile from exactly these rows; nothing writes them until its flush, which waits for us).
__device__ __forceinline__ bool exact_pair_remote(const MergeArgs& A, const PoolPub* pub, int t, const float* rowp, float rn, int nq) {
  const float4* c4 = reinterpret_cast<const float4*>(A.vals + (uint64_t)__ldcg(&pub->ridx[t]) * A.ld);
  float dx = 0.f;
  for (int q = 0; q < nq; ++q) {
    const float4 x = __ldcg(c4 + q);
    const float4 y = __ldcg(reinterpret_cast<const float4*>(rowp) + q);
    dx = __fadd_rn(dx, __fmul_rn(x.x, y.x));
    dx = __fadd_rn(dx, __fmul_rn(x.y, y.y));
    dx = __fadd_rn(dx, __fmul_rn(x.z, y.z));
    dx = __fadd_rn(dx, __fmul_rn(x.w, y.w));
  }
  return cos_match(dx, __ldcg(&pub->cnorm[t]), rn, A.threshold);
}

// Tensor-core screened comparison of the window with representatives [j_begin, j_end) (SELF == false)
// or with the window's own rows (SELF == true: candidate x candidate bits).  One warp handles 8
// representatives per step; KS16 = number of 16-wide k steps (rows are zero-padded to 16*KS16).
// Representatives come from the fragment-order fp16 copy seg_h (KS16/2 16-byte chunks per lane and
// step, coalesced), fetched several steps ahead so that the tensor pipe, not L2 latency, paces the loop.
// MODE 0: this CTA's window against representatives, 1: against the window's own rows (SELF), 2: ANOTHER CTA's
// window (pub) against representatives — the pool's worker side: fragments come from the leader's published copy,
// verified matches go to s.w_f, exact tests read the candidate from global memory.
template <int KS16, int MODE>
__device__ __forceinline__ void tc_compare(const MergeArgs& A, const uint32_t* seg, const float* pos_nrm, const uint4* seg_h,
                                           Smem& s, int W, uint32_t j_begin, uint32_t j_end, uint32_t warp_rank,
                                           uint32_t n_warps, int nq, const PoolPub* pub = nullptr) {
  constexpr bool SELF = MODE == 1;
  uint32_t* const filt = MODE == 2 ? s.w_f : s.s_f;
  const uint32_t lane = lane_id(), g = lane >> 2, tg = lane & 3;
  const float thr_tc = A.threshold - 2e-3f;
  const int ld = A.ld;
  constexpr int NV = KS16 / 2;      // 16-byte chunks per lane per step
  constexpr int QH = KS16 * 2;      // 16-byte chunks per representative
  // A fragments of the whole window stay in registers: 4 row tiles x KS16 k-steps
  uint32_t af[4][KS16][4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int ks = 0; ks < KS16; ++ks) {
      if (MODE == 2) {
        const uint4 x = __ldcg(pub->af + (mt * KS16 + ks) * 32 + lane);
        af[mt][ks][0] = x.x;
        af[mt][ks][1] = x.y;
        af[mt][ks][2] = x.z;
        af[mt][ks][3] = x.w;
      } else {
        const __half* r0 = s.htile + (size_t)(mt * 16 + g) * s.hs + ks * 16 + tg * 2;
        const __half* r1 = r0 + 8 * s.hs;
        af[mt][ks][0] = *reinterpret_cast<const uint32_t*>(r0);
        af[mt][ks][1] = *reinterpret_cast<const uint32_t*>(r1);
        af[mt][ks][2] = *reinterpret_cast<const uint32_t*>(r0 + 8);
        af[mt][ks][3] = *reinterpret_cast<const uint32_t*>(r1 + 8);
      }
    }
  // screen all 64 x 8 pairs of one step, then every lane works through its own survivors: the exact
  // evaluations of different lanes run side by side instead of one (row tile, element) slot after the other
  auto screen = [&](const uint32_t (&bf)[KS16][2], uint32_t jb) {
    uint32_t pend = 0;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS16; ++ks) mma_16816(c, af[mt][ks], bf[ks][0], bf[ks][1]);
      // c[0],c[1]: candidate mt*16+g vs representatives jb+2*tg, +1 ; c[2],c[3]: candidate +8
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (!(c[e] < thr_tc)) pend |= 1u << (mt * 4 + e);  // NaN/inf do not compare below: they go to the exact test
    }
    while (__any_sync(0xffffffffu, pend != 0u)) {
      if (pend != 0u) {
        const int slot = __ffs(pend) - 1;
        pend &= pend - 1;
        const int mt = slot >> 2, e = slot & 3;
        const int t = mt * 16 + (int)g + ((e & 2) ? 8 : 0);
        const uint32_t jj = jb + tg * 2 + (e & 1);
        if (t < W && jj < j_end) {
          if (SELF) {
            if ((int)jj != t && exact_pair<true>(s, t, s.tile + (size_t)jj * s.ts, s.cnorm[jj], nq, A.threshold))
              atomicOr(&s.pair[2 * t + (jj >> 5)], 1u << (jj & 31));
          } else if (!(filt[t] < jj)) {  // skip when an earlier match is already recorded
            // The exact test needs the representative's fp32 row (two dependent global round trips):
            // park the pair and test all parked pairs of the CTA together after the streaming loop.
            uint32_t k = kSurvCap;
            if (jj < (1u << 26)) k = atomicAdd(&s.surv[kSurvCap], 1u);
            if (k < (uint32_t)kSurvCap) {
              s.surv[k] = (jj << 6) | (uint32_t)t;
            } else {
              const uint32_t rr2 = __ldcg(seg + jj);
              const float* rowp = A.vals + (uint64_t)rr2 * ld;
              const float rn = __ldcg(pos_nrm + jj);
              if (MODE == 2 ? exact_pair_remote(A, pub, t, rowp, rn, nq) : exact_pair<false>(s, t, rowp, rn, nq, A.threshold))
                atomicMin(&filt[t], jj);
            }
          }
        }
      }
    }
  };
  if (SELF) {
    for (uint32_t jb = j_begin + warp_rank * 8; jb < j_end; jb += n_warps * 8) {
      const uint32_t j = jb + g;
      const __half* hr = s.htile + (size_t)(j < j_end ? j : 0) * s.hs + tg * 2;
      uint32_t bf[KS16][2];
#pragma unroll
      for (int ks = 0; ks < KS16; ++ks) {
        bf[ks][0] = *reinterpret_cast<const uint32_t*>(hr + ks * 16);
        bf[ks][1] = *reinterpret_cast<const uint32_t*>(hr + ks * 16 + 8);
      }
      screen(bf, jb);
    }
    return;
  }
  // The fp16 copies travel global -> shared with cp.async through a ring private to the warp (every
  // lane copies and later reads its own 16-byte chunks, so no barrier is involved): kRing steps are in
  // flight per warp without holding registers.
  const uint32_t stride = n_warps * 8;
  const uint32_t jb0 = j_begin + warp_rank * 8;
  uint4* ring = s.ring + (size_t)(threadIdx.x >> 5) * (kRing * 32 * NV) + lane;
  auto fetch = [&](int slot, uint32_t jb) {
    const uint32_t j = jb + g;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      uint4* dst = ring + (slot * NV + v) * 32;
      if (j < j_end) cp_async16(dst, seg_h + (size_t)j * QH + v * 4 + tg);
      else *dst = make_uint4(0u, 0u, 0u, 0u);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int u = 0; u < kRing; ++u) fetch(u, jb0 + (uint32_t)u * stride);
  int slot = 0;
  for (uint32_t jb = jb0; jb < j_end; jb += stride) {
    cp_async_wait_group<kRing - 1>();
    uint32_t bf[KS16][2];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const uint4 x = ring[(slot * NV + v) * 32];
      bf[2 * v][0] = x.x;
      bf[2 * v][1] = x.y;
      bf[2 * v + 1][0] = x.z;
      bf[2 * v + 1][1] = x.w;
    }
    screen(bf, jb);
    fetch(slot, jb + stride * kRing);  // after the screen: the slot's values have been consumed
    slot = (slot + 1 == kRing) ? 0 : slot + 1;
  }
  cp_async_wait_all();
}

// Same screen for rows wider than 64 floats: the window's fragments are re-read from shared memory
// per k-step instead of living in registers (ks16 = number of 16-wide k steps, run time).
template <bool SELF>
__device__ __forceinline__ void tc_compare_wide(const MergeArgs& A, const uint32_t* seg, const float* pos_nrm, const uint4* seg_h,
                                                Smem& s, int W, uint32_t j_begin, uint32_t j_end, uint32_t warp_rank,
                                                uint32_t n_warps, int nq, int ks16) {
  const uint32_t lane = lane_id(), g = lane >> 2, tg = lane & 3;
  const float thr_tc = A.threshold - 2e-3f;
  const int ld = A.ld;
  const int qh = ks16 * 2;  // 16-byte chunks of the fp16 copy per representative (ks16 is even)
  for (uint32_t jb = j_begin + warp_rank * 8; jb < j_end; jb += n_warps * 8) {
    const uint32_t j = jb + g;
    const bool valid = j < j_end;
    const __half* hrow = SELF ? s.htile + (size_t)(valid ? j : 0) * s.hs + tg * 2 : nullptr;
    const uint4* hp = SELF ? nullptr : seg_h + (size_t)(valid ? j : 0) * qh + tg;
    float c[4][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int e = 0; e < 4; ++e) c[mt][e] = 0.f;
#pragma unroll 2
    for (int ks2 = 0; ks2 < (ks16 >> 1); ++ks2) {
      uint32_t b[2][2];
      if (SELF) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          b[h][0] = *reinterpret_cast<const uint32_t*>(hrow + (2 * ks2 + h) * 16);
          b[h][1] = *reinterpret_cast<const uint32_t*>(hrow + (2 * ks2 + h) * 16 + 8);
        }
      } else {
        const uint4 x = valid ? __ldcg(hp + ks2 * 4) : make_uint4(0u, 0u, 0u, 0u);
        b[0][0] = x.x; b[0][1] = x.y; b[1][0] = x.z; b[1][1] = x.w;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ks = 2 * ks2 + h;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          const __half* r0 = s.htile + (size_t)(mt * 16 + g) * s.hs + ks * 16 + tg * 2;
          const __half* r1 = r0 + 8 * s.hs;
          uint32_t af[4];
          af[0] = *reinterpret_cast<const uint32_t*>(r0);
          af[1] = *reinterpret_cast<const uint32_t*>(r1);
          af[2] = *reinterpret_cast<const uint32_t*>(r0 + 8);
          af[3] = *reinterpret_cast<const uint32_t*>(r1 + 8);
          mma_16816(c[mt], af, b[h][0], b[h][1]);
        }
      }
    }
    uint32_t pend = 0;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (!(c[mt][e] < thr_tc)) pend |= 1u << (mt * 4 + e);
    while (__any_sync(0xffffffffu, pend != 0u)) {
      if (pend != 0u) {
        const int slot = __ffs(pend) - 1;
        pend &= pend - 1;
        const int mt = slot >> 2, e = slot & 3;
        const int t = mt * 16 + (int)g + ((e & 2) ? 8 : 0);
        const uint32_t jj = jb + tg * 2 + (e & 1);
        if (t < W && jj < j_end) {
          if (SELF) {
            if ((int)jj != t && exact_pair<true>(s, t, s.tile + (size_t)jj * s.ts, s.cnorm[jj], nq, A.threshold))
              atomicOr(&s.pair[2 * t + (jj >> 5)], 1u << (jj & 31));
          } else if (!(s.s_f[t] < jj)) {
            uint32_t k = kSurvCap;
            if (jj < (1u << 26)) k = atomicAdd(&s.surv[kSurvCap], 1u);
            if (k < (uint32_t)kSurvCap) {
              s.surv[k] = (jj << 6) | (uint32_t)t;
            } else {
              const uint32_t rr2 = __ldcg(seg + jj);
              if (exact_pair<false>(s, t, A.vals + (uint64_t)rr2 * ld, __ldcg(pos_nrm + jj), nq, A.threshold))
                atomicMin(&s.s_f[t], jj);
            }
          }
        }
      }
    }
  }
}

// ---- window resolution in the team's leader CTA ---------------------------------------------------------
// Warp 0 replays the reference's sequential order over the window (resolve_decide).  Representatives
// modified inside the window live in a dirty cache of kKD entries; lane e of the decision warp owns
// entry e: its position, member count and a 64-bit mask of the window candidates that match the
// entry's CURRENT value.  A mask goes stale whenever the entry is modified.  Stale masks are NOT
// rebuilt by the decision warp: the CTA's other warps (mask_helpers) rebuild them concurrently under a
// per-entry seqlock (dver odd = being written; a mask is valid iff its version equals the entry's),
// while the decision warp compares each candidate with the stale entries' current values exactly,
// lane per entry, which costs one mul+add chain per candidate however many entries are stale.  Runs of
// candidates that cannot merge (no old match, no pair bit, no dirty match) are accepted in one step as
// soon as every mask is valid again.  Member-chain splices, swap-remove writes and the modified
// rows are independent of the decisions: they are logged and applied by the whole CTA afterwards
// (flush_window).
// Fast pre-test of candidate x entry: dot and |entry|^2 with fused multiply-adds in four independent
// accumulators.  Against the reference's mul-then-add chains the cosine it yields is off by less than
// (4D+40)*2^-24 (products and D-1 additions rounded once each on both sides: 2*gamma_D on the dot,
// gamma_D/2 + u per norm, rsqrt and the final divisions a few u, the reference's 1-(1-sim) at most
// 2^-23), provided the norms are far from the denormal and overflow ranges.  Outside that band the
// decision is therefore the reference's; inside it (or with unsafe norms, NaN, inf) the exact chain decides.
template <int NQ>
__device__ __forceinline__ void fast_dot_nn(const float4* c4, const float4* r4, int nq, float& dot, float& nn) {
  float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
  if (NQ > 0) {
#pragma unroll 4
    for (int q = 0; q < NQ; ++q) {
      const float4 x = c4[q], y = r4[q];
      d0 = __fmaf_rn(x.x, y.x, d0); n0 = __fmaf_rn(y.x, y.x, n0);
      d1 = __fmaf_rn(x.y, y.y, d1); n1 = __fmaf_rn(y.y, y.y, n1);
      d2 = __fmaf_rn(x.z, y.z, d2); n2 = __fmaf_rn(y.z, y.z, n2);
      d3 = __fmaf_rn(x.w, y.w, d3); n3 = __fmaf_rn(y.w, y.w, n3);
    }
  } else {
#pragma unroll 4
    for (int q = 0; q < nq; ++q) {
      const float4 x = c4[q], y = r4[q];
      d0 = __fmaf_rn(x.x, y.x, d0); n0 = __fmaf_rn(y.x, y.x, n0);
      d1 = __fmaf_rn(x.y, y.y, d1); n1 = __fmaf_rn(y.y, y.y, n1);
      d2 = __fmaf_rn(x.z, y.z, d2); n2 = __fmaf_rn(y.z, y.z, n2);
      d3 = __fmaf_rn(x.w, y.w, d3); n3 = __fmaf_rn(y.w, y.w, n3);
    }
  }
  dot = (d0 + d1) + (d2 + d3);
  nn = (n0 + n1) + (n2 + n3);
}

template <int TEAM, int DR>
__device__ void resolve_decide(const MergeArgs& A, Smem& s, int W, int wf, int wb, bool tail_mode, uint32_t i0, uint32_t size0) {
  constexpr int NQ = DR / 4;  // 0: run-time width
  const int D = A.D, ld = A.ld, nq = NQ > 0 ? NQ : (ld >> 2);
  const int ts = NQ > 0 ? DR + 4 : s.ts;
  const uint32_t lane = lane_id();
  const float thr = A.threshold;
  const float band = (4.f * (float)D + 40.f) * 5.9604645e-8f;
  const float thr_hi = thr + band, thr_lo = thr - band;
  int nd = 0, a = 0, fi = 0, bi = 0, merges = 0;
  uint32_t i = i0, size = size0;
  bool from_back = false, back_exhausted = false;
  int dbg_undec = 0, dbg_full = 0;
  uint32_t accd_lo = 0, accd_hi = 0;  // accepted-in-window candidates that have since been modified
  uint32_t accm_lo = 0, accm_hi = 0;  // accepted-in-window candidates (tile indices)
  uint32_t my_dpos = kInf, my_dm_lo = 0, my_dm_hi = 0;  // this lane's dirty entry
  uint32_t my_ver = 0;                                  // version of its value (even)
  bool my_valid = true;                                 // its mask belongs to that version (lanes without an entry: true)
  int my_dcnt = 0, my_last = -1;                        // its member count; last candidate merged into it
  const bool rprof = A.dbg != nullptr;
  long long p_top = 0, p_cmp = 0, p_bc = 0, p_mrg = 0, tq0 = 0, tq1 = 0;
  int n_run = 0, n_cmp = 0, n_it = 0, n_exact = 0;
  // Candidates with no old match and no match bit against any other window row ("easy") can only merge
  // into a representative modified in this window; while every dirty mask is valid that is one bit
  // test, so whole runs of them are accepted at once instead of one loop trip each.
  uint32_t easy_lo, easy_hi;
  {
    const int t0 = (int)lane, t1 = (int)lane + 32;
    const bool e0 = t0 < wf && s.s_f[t0] == kInf && (s.pair[2 * t0] | s.pair[2 * t0 + 1]) == 0u;
    const bool e1 = t1 < wf && s.s_f[t1] == kInf && (s.pair[2 * t1] | s.pair[2 * t1 + 1]) == 0u;
    easy_lo = __ballot_sync(0xffffffffu, e0);
    easy_hi = __ballot_sync(0xffffffffu, e1);
  }
  while (i < size) {
    if (rprof) tq0 = clock64();
    ++n_it;
    // masks the helper warps have finished since the last look
    if (!my_valid && ld_vol(s.mver + lane) == my_ver) {
      fence_cta();
      my_dm_lo = ld_vol(s.dmlo + lane);
      my_dm_hi = ld_vol(s.dmhi + lane);
      my_valid = true;
    }
    const uint32_t inval = __ballot_sync(0xffffffffu, !my_valid);  // entries whose match mask is stale
    if (!from_back && inval == 0u && fi < wf) {
      uint32_t any_lo = 0u, any_hi = 0u;
      if (nd > 0) {
        any_lo = __reduce_or_sync(0xffffffffu, my_dm_lo);
        any_hi = __reduce_or_sync(0xffffffffu, my_dm_hi);
      }
      const unsigned long long ok = ((unsigned long long)(easy_hi & ~any_hi) << 32) | (unsigned long long)(easy_lo & ~any_lo);
      const unsigned long long stop = ~(ok >> fi);
      int k = stop ? (__ffsll((long long)stop) - 1) : 64;
      k = min(k, min(wf - fi, (int)(size - i)));
      if (k > 0) {
        for (int j = (int)lane; j < k; j += 32) s.acc[a + j] = (uint32_t)(fi + j);
        const unsigned long long bits = ((k >= 64) ? ~0ull : ((1ull << k) - 1ull)) << fi;
        accm_lo |= (uint32_t)bits;
        accm_hi |= (uint32_t)(bits >> 32);
        a += k;
        i += (uint32_t)k;
        fi += k;
        ++n_run;
        if (rprof) p_top += clock64() - tq0;
        continue;
      }
    }
    int t;
    if (from_back) {
      if (!tail_mode && bi >= wb) { back_exhausted = true; break; }
      t = tail_mode ? (wf - 1 - bi) : (wf + bi);
    } else {
      if (!tail_mode && fi >= wf) break;
      t = fi;
    }
    // everything the trip needs about the candidate, requested together
    const uint32_t fpos = s.s_f[t];
    const uint32_t plo = s.pair[2 * t], phi = s.pair[2 * t + 1];
    const float cn = s.cnorm[t];
    const int c1 = s.ccnt[t];
    const int ct_tail = s.ctail[t];
    const float4* c4 = reinterpret_cast<const float4*>(s.tile + (size_t)t * ts);
    const uint32_t tbit_lo = (t < 32) ? (1u << t) : 0u, tbit_hi = (t < 32) ? 0u : (1u << (t - 32));
    if (from_back) ++bi; else ++fi;
    uint32_t best = kInf;
    if (rprof) {
      tq1 = clock64();
      p_top += tq1 - tq0;
    }
    if (nd > 0) {
      // (a) representatives modified in this window: match bits against their current values where the
      // mask is valid; where it is stale, lane e tests the candidate against entry e's current value
      bool hit = my_valid && ((my_dm_lo & tbit_lo) | (my_dm_hi & tbit_hi)) != 0u;
      if (inval != 0u) {
        ++n_cmp;
        const bool mine = !my_valid;
        const float4* r4 = reinterpret_cast<const float4*>(s.dvals + (size_t)(mine ? lane : 0u) * ts);
        float dot, nn;
        fast_dot_nn<NQ>(c4, r4, nq, dot, nn);
        const float rn = sqrtf(nn);
        const float prod = cn * rn;
        const float sim = __fdividef(dot, prod);
        const bool safe = prod > 1e-18f && prod < 1e18f;
        bool amb = false;
        if (mine) {
          hit = safe && sim >= thr_hi;
          amb = !hit && !(safe && sim < thr_lo);
        }
        if (__any_sync(0xffffffffu, amb)) {  // inside the error band: the reference's own arithmetic decides
          ++n_exact;
          float de = 0.f, ne = 0.f;
#pragma unroll 4
          for (int q = 0; q < nq; ++q) {
            const float4 x = c4[q], y = r4[q];
            de = __fadd_rn(de, __fmul_rn(x.x, y.x)); ne = __fadd_rn(ne, __fmul_rn(y.x, y.x));
            de = __fadd_rn(de, __fmul_rn(x.y, y.y)); ne = __fadd_rn(ne, __fmul_rn(y.y, y.y));
            de = __fadd_rn(de, __fmul_rn(x.z, y.z)); ne = __fadd_rn(ne, __fmul_rn(y.z, y.z));
            de = __fadd_rn(de, __fmul_rn(x.w, y.w)); ne = __fadd_rn(ne, __fmul_rn(y.w, y.w));
          }
          if (amb) hit = cos_match(de, cn, __fsqrt_rn(ne), thr);
        }
      }
      best = __reduce_min_sync(0xffffffffu, hit ? my_dpos : kInf);
    }
    if (rprof) {
      tq0 = clock64();
      p_cmp += tq0 - tq1;
    }
    // (b) first matching old representative as of the window start
    if (fpos != kInf) {
      const bool is_dirty = nd > 0 && __any_sync(0xffffffffu, my_dpos == fpos);
      if (!is_dirty) {
        best = min(best, fpos);
      } else if (best > fpos) {
        // fpos changed since the window start and no modified representative at or before it
        // matches: a clean match between fpos and `best` cannot be ruled out -> next window
        if (from_back) --bi; else --fi;
        dbg_undec = 1;
        break;
      }
    }
    // (c) representatives accepted in this window and not modified since: precomputed pair bits
    if (best >= i0 && (((plo & accm_lo & ~accd_lo) | (phi & accm_hi & ~accd_hi)) != 0u)) {
      __syncwarp();  // s.acc[] is written by lane 0
      for (int k0 = 0; k0 < a; k0 += 32) {
        const int k = k0 + (int)lane;
        bool ok = false;
        if (k < a) {
          const uint32_t u = s.acc[k];
          const bool bit = (u < 32) ? ((plo >> u) & 1u) : ((phi >> (u - 32)) & 1u);
          const bool dirty = (u < 32) ? ((accd_lo >> u) & 1u) : ((accd_hi >> (u - 32)) & 1u);
          ok = bit && !dirty;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (m) {
          best = min(best, i0 + (uint32_t)k0 + (uint32_t)(__ffs(m) - 1));
          break;
        }
      }
    }
    if (best == kInf) {
      // no merge: the candidate becomes representative i (its norm is published with the flush)
      if (lane == 0) s.acc[a] = (uint32_t)t;
      accm_lo |= tbit_lo;
      accm_hi |= tbit_hi;
      ++a;
      ++i;
      from_back = false;
      if (rprof) p_bc += clock64() - tq0;
      continue;
    }
    // merge the candidate into the representative at position `best`
    if (rprof) {
      tq1 = clock64();
      p_bc += tq1 - tq0;
    }
    const uint32_t p = best;
    const uint32_t em = __ballot_sync(0xffffffffu, my_dpos == p);
    const bool fresh = em == 0u;
    const int e = fresh ? nd : (__ffs(em) - 1);
    // seqlock: the entry's version is odd while its value is being written
    const uint32_t ver_e = __shfl_sync(0xffffffffu, my_ver, e);
    if (lane == 0) st_vol(s.dver + e, ver_e + 1u);
    fence_cta();
    // the representative's value before this merge: its cache entry, or (first modification in this
    // window) the prefetched row of an old representative / the window row of an accepted candidate
    const float* src;
    int c2;
    if (!fresh) {
      src = s.dvals + (size_t)e * ts;
      c2 = __shfl_sync(0xffffffffu, my_dcnt, e);
    } else if (p < i0) {
      // an old representative enters the cache only as this candidate's precomputed first match
      // (any other old position in `best` is already dirty), so its row was prefetched
      src = s.pre + (size_t)t * ts;
      c2 = s.pcnt[t];
      if (lane == 0) {
        s.dridx[e] = s.pridx[t];
        s.dhead[e] = s.phead[t];
        s.dtail[e] = s.ptail[t];
        s.dpos[e] = p;
      }
    } else {
      __syncwarp();  // s.acc[] is written by lane 0
      const uint32_t u = s.acc[p - i0];
      src = s.tile + (size_t)u * ts;
      c2 = s.ccnt[u];
      if (lane == 0) {
        s.dridx[e] = s.ridx[u];
        s.dhead[e] = s.chead[u];
        s.dtail[e] = s.ctail[u];
        s.dpos[e] = p;
      }
      if (u < 32) accd_lo |= 1u << u; else accd_hi |= 1u << (u - 32);
    }
    {
      const float* c = s.tile + (size_t)t * ts;
      float* r = s.dvals + (size_t)e * ts;
      for (int d = lane; d < ld; d += 32) {
        const float v = src[d];
        r[d] = d < D ? consensus1(c[d], c1, v, c2) : v;
      }
    }
    // merge log: which candidate was merged into this entry just before t (applied by the flush)
    {
      const int prev = fresh ? -1 : __shfl_sync(0xffffffffu, my_last, e);
      if (lane == 0) {
        s.mprev[t] = prev;
        s.ment[t] = e;
      }
      if ((int)lane == e) {
        if (fresh) {
          my_dpos = p;
          my_last = -1;
        }
        my_dcnt = c1 + c2;
        if (ct_tail >= 0) my_last = t;  // only candidates that carry ids take part in the chain
        my_ver = ver_e + 2u;
        my_valid = false;
      }
    }
    fence_cta();
    __syncwarp();
    if (lane == 0) {
      st_vol(s.dver + e, ver_e + 2u);
      if (fresh) {
        fence_cta();
        st_vol(s.ro + RO_ND, (uint32_t)(nd + 1));  // the helpers may look at the entry from now on
      }
    }
    if (fresh) ++nd;
    --size;
    ++merges;
    from_back = true;
    if (rprof) p_mrg += clock64() - tq1;
    if (nd == kKD) { dbg_full = 1; break; }  // dirty cache full: flush and start a new window
  }
  __syncwarp();
  if (rprof && lane == 0) {
    atomicAdd(A.dbg + 14, (unsigned long long)p_top);
    atomicAdd(A.dbg + 15, (unsigned long long)p_cmp);
    atomicAdd(A.dbg + 16, (unsigned long long)p_bc);
    atomicAdd(A.dbg + 17, (unsigned long long)p_mrg);
    atomicAdd(A.dbg + 22, (unsigned long long)n_it);
    atomicAdd(A.dbg + 23, (unsigned long long)n_run);
    atomicAdd(A.dbg + 24, (unsigned long long)n_cmp);
    atomicAdd(A.dbg + 25, (unsigned long long)n_exact);
  }
  if ((int)lane < nd) {
    s.dcnt[lane] = my_dcnt;
    s.dlast[lane] = my_last;
  }
  if (lane == 0) {
    s.ro[RO_A] = (uint32_t)a;
    s.ro[RO_I] = i;
    s.ro[RO_SIZE] = size;
    s.ro[RO_FROM_BACK] = from_back ? 1u : 0u;
    s.ro[RO_BI] = (uint32_t)bi;
    s.ro[RO_MERGES] = (uint32_t)merges;
    s.ro[RO_BACK_EXH] = back_exhausted ? 1u : 0u;
    s.ro[RO_UNDEC] = (uint32_t)dbg_undec;
    s.ro[RO_FULL] = (uint32_t)dbg_full;
    s.ro[RO_EXAMINED] = (uint32_t)(fi + bi);
  }
  fence_cta();
  __syncwarp();
  if (lane == 0) st_vol(s.ro + RO_DONE, 1u);
}

// The CTA's other warps while warp 0 decides: rebuild the match masks of modified entries.  Helper
// hw looks after entries hw, hw + nh, ...  A mask is computed for ALL window candidates against one
// version of the entry's value and published only if the version is unchanged afterwards.
__device__ void mask_helpers(const MergeArgs& A, Smem& s, int W, int hw, int nh) {
  const int nq = A.ld >> 2;
  const uint32_t lane = lane_id();
  for (;;) {
    bool did = false;
    const int nd = (int)__shfl_sync(0xffffffffu, ld_vol(s.ro + RO_ND), 0);
    for (int e = hw; e < nd; e += nh) {
      const uint32_t v1 = __shfl_sync(0xffffffffu, ld_vol(s.dver + e), 0);
      const uint32_t mv = __shfl_sync(0xffffffffu, ld_vol(s.mver + e), 0);
      if ((v1 & 1u) || v1 == mv) continue;
      fence_cta();
      const float4* r4 = reinterpret_cast<const float4*>(s.dvals + (size_t)e * s.ts);
      uint32_t w[2] = {0u, 0u};
      float rn = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (half * 32 >= W) break;
        const int tc = half * 32 + (int)lane;
        const float4* c4 = reinterpret_cast<const float4*>(s.tile + (size_t)min(tc, W - 1) * s.ts);
        float dot = 0.f, nn = 0.f;
        if (half == 0) {  // the entry's norm comes out of the first pass (every lane walks the same chain)
#pragma unroll 4
          for (int q = 0; q < nq; ++q) {
            const float4 x = c4[q], y = r4[q];
            dot = __fadd_rn(dot, __fmul_rn(x.x, y.x)); nn = __fadd_rn(nn, __fmul_rn(y.x, y.x));
            dot = __fadd_rn(dot, __fmul_rn(x.y, y.y)); nn = __fadd_rn(nn, __fmul_rn(y.y, y.y));
            dot = __fadd_rn(dot, __fmul_rn(x.z, y.z)); nn = __fadd_rn(nn, __fmul_rn(y.z, y.z));
            dot = __fadd_rn(dot, __fmul_rn(x.w, y.w)); nn = __fadd_rn(nn, __fmul_rn(y.w, y.w));
          }
          rn = __fsqrt_rn(nn);
        } else {
          dot = dot_seq(c4, r4, nq);
        }
        const bool mt = tc < W && cos_match(dot, s.cnorm[min(tc, W - 1)], rn, A.threshold);
        w[half] = __ballot_sync(0xffffffffu, mt);
      }
      fence_cta();
      const uint32_t v2 = __shfl_sync(0xffffffffu, ld_vol(s.dver + e), 0);
      if (v2 == v1) {
        if (lane == 0) {
          st_vol(s.dmlo + e, w[0]);
          st_vol(s.dmhi + e, w[1]);
          fence_cta();
          st_vol(s.mver + e, v1);
        }
        did = true;
      }
    }
    if (__shfl_sync(0xffffffffu, ld_vol(s.ro + RO_DONE), 0)) break;
    if (!did) __nanosleep(64);
  }
}

// ---- speculative window resolution -----------------------------------------------------------------------
// The sequential loop above is bound by the latency of one warp's dependent instructions (about 1.5 k
// cycles per candidate).  Almost all of that is floating point that rarely changes a decision: a
// candidate merges into the first representative that matched it at the window start, or into an earlier
// window candidate it has a pair bit with, or becomes a representative.  So the window is resolved in
// four parallel-friendly passes and the sequential loop is only the fallback:
//   scan    (warp 0, integers only) replays the reference's order with PREDICTED decisions: first old
//           match, else earliest accepted candidate with a pair bit, else accept.  It records, per
//           examined candidate x, everything the prediction depended on, and per merge a version link;
//   versions (all warps, one entry's chain per warp) computes the consensus values every predicted merge
//           produces, with the reference's exact arithmetic;
//   match   (all warps) tests every examined candidate against the CURRENT version, at its turn, of every
//           representative modified before it (fast test + exact chain inside its error band);
//   verify  (warp 0, lane per candidate) derives the TRUE decision each candidate would get if everything
//           before it went as predicted.  The first candidate whose true decision differs from the
//           prediction cuts the window: everything before it is exactly what the sequential algorithm
//           does (by induction over the examine order) and is committed; the rest is re-examined by the
//           next window, which then runs the sequential loop once (so a hard spot cannot stall progress).
struct Spec {
  float* vers;        // [kW][ts] value after predicted merge k
  unsigned long long* emask;  // [kKD] merges (bit k) that went into the entry, in merge order
  unsigned long long* cmask;  // [2 kW] per target id: examine indices of its id-carrying merged candidates — parallel scan
  int32_t* vcnt;      // [kW] member count after merge k
  uint32_t* xtarget;  // [kW] predicted position (kInf: accept) per examined candidate
  uint32_t* acc_lo;   // [kW] accepted candidates before x (tile index bits 0..31)
  uint32_t* acc_hi;   // [kW]
  uint32_t* accd_lo;  // [kW] ... of those, the ones modified before x
  uint32_t* accd_hi;  // [kW]
  uint32_t* dmatch;   // [kW] entries whose current version matches the candidate examined at x
  uint32_t* dbase;    // [kKD] where the entry's value before its first merge lives: t (prefetched row of t's first match) or 0x100|u (window row u)
  int32_t* dbcnt;     // [kKD] its member count before the first merge
  uint8_t* xcand;     // [kW] candidate examined at x
  uint8_t* xep;       // [kW+1] merges before x (the state before x follows from it: a = x - merges, i = i0 + a, size = size0 - merges)
  uint8_t* xfi;       // [kW+1] front candidates examined before x (back candidates: x - xfi)
  uint8_t* xfe;       // [kW] entry that holds the candidate's first old match if that representative was modified before x, else 0xFF
  uint8_t* vcand;     // [kW] candidate merged by k
  int8_t* vlast;      // [kW] last id-carrying candidate merged into the entry up to and including k (-1: none)
  uint8_t* arank;     // [kW] acceptance rank of a candidate
  uint8_t* cls;       // [kW] class of a candidate's first old match: the smallest candidate index with the same one
  uint8_t* dense;     // [2 kW] target id (class, or kW + accepted candidate) -> cache entry, 0xFF: none yet
  int8_t* elast;      // [kKD] last id-carrying candidate merged into the entry so far
  uint8_t* sx;        // [kW] examine index of a candidate (0xFF: not examined) — parallel scan
  uint8_t* xcls;      // [kW] target id of the candidate examined at x (0xFF: accept) — parallel scan
  uint32_t* cfirst;   // [2 kW] examine index of the first merge into a target id (0xFF: none) — parallel scan
};
__host__ __device__ inline size_t spec_bytes_for(int ld) {
  return sizeof(float) * (size_t)kW * (row_width(ld) + 4) + 8 * (size_t)(kKD + 2 * kW) + 4 * (size_t)(kW * 9 + kKD * 2) + (size_t)kW * 12 + 2 + kKD + 16;
}
__device__ __forceinline__ void carve_spec(Spec& sp, void* base, int ts) {
  sp.vers = reinterpret_cast<float*>(base);
  sp.emask = reinterpret_cast<unsigned long long*>(sp.vers + (size_t)kW * ts);  // kW*ts floats: a multiple of 16 bytes
  sp.cmask = sp.emask + kKD;
  uint32_t* u = reinterpret_cast<uint32_t*>(sp.cmask + 2 * kW);
  sp.vcnt = reinterpret_cast<int32_t*>(u); u += kW;
  sp.xtarget = u; u += kW;
  sp.acc_lo = u; u += kW;
  sp.acc_hi = u; u += kW;
  sp.accd_lo = u; u += kW;
  sp.accd_hi = u; u += kW;
  sp.dmatch = u; u += kW;
  sp.dbase = u; u += kKD;
  sp.dbcnt = reinterpret_cast<int32_t*>(u); u += kKD;
  sp.cfirst = u; u += 2 * kW;
  uint8_t* b = reinterpret_cast<uint8_t*>(u);
  sp.xcand = b; b += kW;
  sp.xep = b; b += kW + 1;
  sp.xfi = b; b += kW + 1;
  sp.xfe = b; b += kW;
  sp.vcand = b; b += kW;
  sp.vlast = reinterpret_cast<int8_t*>(b); b += kW;
  sp.arank = b; b += kW;
  sp.cls = b; b += kW;
  sp.dense = b; b += 2 * kW;
  sp.elast = reinterpret_cast<int8_t*>(b); b += kKD;
  sp.sx = b; b += kW;
  sp.xcls = b;
}

// scan: returns the number of examined candidates.  Warp 0 only.  The replay is scalar integer code on lane 0
// (measured: the same loop run warp-synchronously with shuffles and votes, or uniformly on all lanes, costs
// up to three times as much per candidate); only runs of candidates that are accepted for sure (no old
// match, no pair bit) are spread over the lanes, one candidate each.
__device__ int spec_scan(Smem& s, Spec& sp, int W, int wf, int wb, bool tail_mode, uint32_t i0, uint32_t size0) {
  const uint32_t lane = lane_id();
  // candidates whose first old match is the same representative share a cache entry: class = the
  // smallest candidate index with that first match
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int t = half * 32 + (int)lane;
    if (t < W) {
      const uint32_t f = s.s_f[t];
      int c = t;
      if (f != kInf)
        for (int u = 0; u < t; ++u)
          if (s.s_f[u] == f) {
            c = u;
            break;
          }
      sp.cls[t] = (uint8_t)c;
    }
    sp.dense[half * 32 + lane] = 0xFF;
    sp.dense[kW + half * 32 + lane] = 0xFF;
  }
  sp.emask[lane] = 0ull;
  unsigned long long easy;
  {
    const int t0 = (int)lane, t1 = (int)lane + 32;
    const bool e0 = t0 < wf && s.s_f[t0] == kInf && (s.pair[2 * t0] | s.pair[2 * t0 + 1]) == 0u;
    const bool e1 = t1 < wf && s.s_f[t1] == kInf && (s.pair[2 * t1] | s.pair[2 * t1 + 1]) == 0u;
    easy = ((unsigned long long)__ballot_sync(0xffffffffu, e1) << 32) | (unsigned long long)__ballot_sync(0xffffffffu, e0);
  }
  __syncwarp();
  // Lane 0 replays the order candidate by candidate (scalar integer code; its registers hold the state).  When
  // it reaches a run of front candidates that are accepted for sure, the state is broadcast and the run is
  // recorded by all lanes, one candidate each.
  int nd = 0, a = 0, fi = 0, merges = 0, x = 0;
  uint32_t i = i0, size = size0;
  bool from_back = false, back_exhausted = false, full = false;
  unsigned long long accm = 0ull, accd = 0ull;
  for (;;) {
    int mode = 2;  // 1: an easy run starts at fi, 2: the window is over
    if (lane == 0) {
      while (i < size) {
        if (!from_back && fi < wf && ((easy >> fi) & 1ull)) {
          mode = 1;
          break;
        }
        const int bi = x - fi;
        int t;
        if (from_back) {
          if (!tail_mode && bi >= wb) { back_exhausted = true; break; }
          t = tail_mode ? (wf - 1 - bi) : (wf + bi);
        } else {
          if (!tail_mode && fi >= wf) break;
          t = fi;
        }
        // what this candidate's decision will be checked against
        sp.xcand[x] = (uint8_t)t;
        sp.xep[x] = (uint8_t)merges;
        sp.xfi[x] = (uint8_t)fi;
        sp.acc_lo[x] = (uint32_t)accm;
        sp.acc_hi[x] = (uint32_t)(accm >> 32);
        sp.accd_lo[x] = (uint32_t)accd;
        sp.accd_hi[x] = (uint32_t)(accd >> 32);
        // predicted decision: first old match, else the earliest accepted candidate with a pair bit, else accept
        uint32_t P = kInf;
        int id = -1;
        const uint32_t fpos = s.s_f[t];
        if (fpos != kInf) {
          P = fpos;
          id = sp.cls[t];
        } else {
          unsigned long long pm = (((unsigned long long)s.pair[2 * t + 1] << 32) | (unsigned long long)s.pair[2 * t]) & accm;
          uint32_t rank = 0xFFu;
          while (pm) {
            const int u = __ffsll((long long)pm) - 1;
            pm &= pm - 1ull;
            rank = min(rank, (uint32_t)sp.arank[u]);
          }
          if (rank != 0xFFu) {
            P = i0 + rank;
            id = kW + (int)s.acc[rank];
          }
        }
        if (!from_back) ++fi;
        sp.xtarget[x] = P;
        if (P == kInf) {
          s.acc[a] = (uint32_t)t;
          sp.arank[t] = (uint8_t)a;
          sp.xfe[x] = 0xFF;
          accm |= 1ull << t;
          ++a;
          ++i;
          ++x;
          from_back = false;
          continue;
        }
        int e = sp.dense[id];
        const bool fresh = e == 0xFF;
        int prev_last = -1;
        unsigned long long em = 1ull << merges;
        if (fresh) {
          e = nd++;
          sp.dense[id] = (uint8_t)e;
          s.dpos[e] = P;
          if (P < i0) {
            sp.dbase[e] = (uint32_t)t;  // the prefetched row of t's first match
            sp.dbcnt[e] = s.pcnt[t];
            s.dridx[e] = s.pridx[t];
            s.dhead[e] = s.phead[t];
            s.dtail[e] = s.ptail[t];
          } else {
            const uint32_t u = (uint32_t)(id - kW);
            sp.dbase[e] = 0x100u | u;
            sp.dbcnt[e] = s.ccnt[u];
            s.dridx[e] = s.ridx[u];
            s.dhead[e] = s.chead[u];
            s.dtail[e] = s.ctail[u];
            accd |= 1ull << u;
          }
        } else {
          prev_last = (int)sp.elast[e];
          em |= sp.emask[e];
        }
        const int new_last = s.ctail[t] >= 0 ? t : prev_last;  // only candidates that carry ids take part in the chain
        sp.xfe[x] = (fpos != kInf && !fresh) ? (uint8_t)e : (uint8_t)0xFF;
        sp.vcand[merges] = (uint8_t)t;
        sp.vlast[merges] = (int8_t)new_last;
        sp.elast[e] = (int8_t)new_last;
        sp.emask[e] = em;
        s.mprev[t] = prev_last;
        s.ment[t] = e;
        ++merges;
        --size;
        ++x;
        from_back = true;
        if (nd == kKD) { full = true; break; }
      }
    }
    __syncwarp();
    mode = __shfl_sync(0xffffffffu, mode, 0);
    x = __shfl_sync(0xffffffffu, x, 0);
    if (mode == 2) break;
    fi = __shfl_sync(0xffffffffu, fi, 0);
    a = __shfl_sync(0xffffffffu, a, 0);
    merges = __shfl_sync(0xffffffffu, merges, 0);
    i = __shfl_sync(0xffffffffu, i, 0);
    size = __shfl_sync(0xffffffffu, size, 0);
    // the run: front candidates fi .. fi + k - 1, one lane per candidate
    const unsigned long long stop = ~(easy >> fi);
    int k = stop ? (__ffsll((long long)stop) - 1) : 64;
    k = min(k, min(wf - fi, (int)(size - i)));
    for (int j = (int)lane; j < k; j += 32) {
      const int t = fi + j, xx = x + j;
      s.acc[a + j] = (uint32_t)t;
      sp.arank[t] = (uint8_t)(a + j);
      sp.xcand[xx] = (uint8_t)t;
      sp.xep[xx] = (uint8_t)merges;
      sp.xfi[xx] = (uint8_t)t;
      sp.xfe[xx] = 0xFF;
      sp.xtarget[xx] = kInf;
    }
    accm |= ((k >= 64) ? ~0ull : ((1ull << k) - 1ull)) << fi;  // lane 0's copy is the one that is used
    a += k;
    i += (uint32_t)k;
    fi += k;
    x += k;
    __syncwarp();
  }
  if (lane == 0) {
    sp.xep[x] = (uint8_t)merges;
    sp.xfi[x] = (uint8_t)fi;
    s.ro[RO_BACK_EXH] = back_exhausted ? 1u : 0u;
    s.ro[RO_FULL] = full ? 1u : 0u;
    s.ro[RO_ND] = (uint32_t)nd;       // entries the scan allocated (match pass)
    s.ro[RO_MERGES] = (uint32_t)merges;
  }
  return x;
}

// ---- the scan as a prefix computation ---------------------------------------------------------------------
// The predicted outcome of almost every candidate is known before the replay: a candidate with an old match
// merges, a candidate with no old match and no pair bit is accepted; only a candidate whose possible matches
// are window candidates ("unsure") depends on what was accepted before it.  Given the outcomes, the examine
// order is closed form: front candidates are examined in order, and every merging front candidate pulls one
// BURST of back candidates — consecutive back candidates up to and including the first accepted one.  So
//   front j  is examined at  j + (back candidates in the bursts of the merging fronts before j),
//   back  k  is examined at  (index of the merging front that owns k's burst) + 1 + k,
// both popcounts and n-th-set-bit selections on two 64-bit masks, one lane per candidate.  Unsure candidates
// start as accepts and are corrected by iterating (each round fixes at least the first wrong one); windows
// that do not settle in a few rounds, and tail windows (one stream eaten from both ends), use the scalar scan.
__device__ __forceinline__ int sel32(uint32_t m, int n) {  // index of the n-th (0-based) set bit; n < popc(m)
  int pos = 0;
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const int c = __popc(m & ((1u << w) - 1u));
    if (n >= c) {
      n -= c;
      m >>= w;
      pos += w;
    }
  }
  return pos;
}
__device__ __forceinline__ int sel64(unsigned long long m, int n) {  // index of the n-th (0-based) set bit; 64: none
  const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
  const int pl = __popc(lo);
  if (n < pl) return sel32(lo, n);
  n -= pl;
  if (n < __popc(hi)) return 32 + sel32(hi, n);
  return 64;
}
__device__ __forceinline__ unsigned long long below64(int x) { return x >= 64 ? ~0ull : ((1ull << x) - 1ull); }
__device__ __forceinline__ unsigned long long ballot64(bool p0, bool p1) {
  return ((unsigned long long)__ballot_sync(0xffffffffu, p1) << 32) | (unsigned long long)__ballot_sync(0xffffffffu, p0);
}

// Returns the number of examined candidates, or -1 if the window has to take the scalar scan.  Tail windows (one
// stream of W candidates eaten from both ends: back index k is candidate W-1-k) follow the same formulas with both
// masks over the same candidates; a candidate is examined by whichever end reaches it first, all W are examined.
// LEAN: the examine order and the outcomes come from a replay on lane 0 that carries nothing but the order
// (two counters, the accepted mask in a register, one shared-memory load per candidate); every record is then
// built by the lanes in parallel exactly as for the closed form.
template <bool LEAN>
__device__ int spec_scan_par(const MergeArgs& A, Smem& s, Spec& sp, int W, int wf, int wb, bool tail_mode, uint32_t i0) {
  const uint32_t lane = lane_id();
  const int INF = 255;
  const int wb_in = wb;
  long long tq0 = 0, tq1 = 0;
  if (A.dbg && lane == 0) tq0 = clock64();
  if (tail_mode) wb = wf;
  // per-candidate facts (two candidates per lane: t = lane and lane + 32)
  bool valid[2], hasf[2], unsure[2];
  uint32_t fpos[2], plo[2], phi[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int t = h * 32 + (int)lane;
    valid[h] = t < W;
    fpos[h] = valid[h] ? s.s_f[t] : kInf;
    plo[h] = valid[h] ? s.pair[2 * t] : 0u;
    phi[h] = valid[h] ? s.pair[2 * t + 1] : 0u;
    hasf[h] = fpos[h] != kInf;
    unsure[h] = valid[h] && !hasf[h] && (plo[h] | phi[h]) != 0u;
    sp.cfirst[h * 32 + lane] = 0xFFu;
    sp.cfirst[kW + h * 32 + lane] = 0xFFu;
    sp.cmask[h * 32 + lane] = 0ull;
    sp.cmask[kW + h * 32 + lane] = 0ull;
  }
  sp.emask[lane] = 0ull;
  // target ids: candidates whose first old match is the same representative share one — the smallest candidate
  // index with that first match (match within each half, then the second half looks its value up in the first)
  {
    const uint32_t m0 = __match_any_sync(0xffffffffu, fpos[0]);
    const uint32_t m1 = __match_any_sync(0xffffffffu, fpos[1]);
    const int c0 = __ffs(m0) - 1;
    int c1 = 32 + __ffs(m1) - 1;
    if (hasf[1] && (int)lane == __ffs(m1) - 1) {  // one lane per distinct value of the second half
      for (int u = 0; u < 32 && u < W; ++u)
        if (s.s_f[u] == fpos[1]) {
          c1 = u;
          break;
        }
    }
    c1 = __shfl_sync(0xffffffffu, c1, __ffs(m1) - 1);
    if (valid[0]) sp.cls[lane] = (uint8_t)(hasf[0] ? c0 : (int)lane);
    if (valid[1]) sp.cls[32 + lane] = (uint8_t)(hasf[1] ? c1 : 32 + (int)lane);
  }
  const unsigned long long fm = below64(wf), bm = below64(wb);
  unsigned long long OM = ballot64(hasf[0], hasf[1]);  // predicted merges, by candidate index
  int xs[2] = {INF, INF}, ustar[2] = {-1, -1}, n_ex = 0;
  bool asfront[2] = {true, true};
  bool back_exhausted = false;
  if (LEAN) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (valid[h]) sp.sx[h * 32 + lane] = 0xFF;
    __syncwarp();
    const unsigned long long HF = OM;
    unsigned long long om = 0ull, fr = 0ull;
    int x = 0;
    if (lane == 0) {
      unsigned long long accm = 0ull;
      int fi = 0, bi = 0;
      bool from_back = false;
      const uint2* pairs = reinterpret_cast<const uint2*>(s.pair);
      while (x < W) {
        int t;
        if (from_back) {
          if (!tail_mode && bi >= wb_in) {
            back_exhausted = true;
            break;
          }
          t = tail_mode ? (wf - 1 - bi) : (wf + bi);
          ++bi;
        } else {
          if (!tail_mode && fi >= wf) break;
          t = fi;
          fr |= 1ull << t;
          ++fi;
        }
        const uint2 pr = pairs[t];
        const unsigned long long pm = ((unsigned long long)pr.y << 32) | (unsigned long long)pr.x;
        const bool merge = ((HF >> t) & 1ull) != 0ull || (pm & accm) != 0ull;
        sp.sx[t] = (uint8_t)x;
        if (merge) om |= 1ull << t;
        else accm |= 1ull << t;
        from_back = merge;
        ++x;
      }
    }
    OM = __shfl_sync(0xffffffffu, om, 0);
    fr = __shfl_sync(0xffffffffu, fr, 0);
    n_ex = __shfl_sync(0xffffffffu, x, 0);
    back_exhausted = __shfl_sync(0xffffffffu, back_exhausted ? 1 : 0, 0) != 0;
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = h * 32 + (int)lane;
      xs[h] = valid[h] ? (int)sp.sx[t] : INF;
      if (xs[h] == 0xFF) xs[h] = INF;
      asfront[h] = ((fr >> t) & 1ull) != 0ull;
      // the earliest accepted mate of an unsure candidate that merges
      if (unsure[h] && xs[h] < n_ex && ((OM >> t) & 1ull)) {
        unsigned long long pm = (((unsigned long long)phi[h] << 32) | (unsigned long long)plo[h]) & ~OM;
        int best = INF, bu = -1;
        while (pm) {
          const int u = __ffsll((long long)pm) - 1;
          pm &= pm - 1ull;
          const int xu = sp.sx[u];
          if (xu < xs[h] && xu < best) {
            best = xu;
            bu = u;
          }
        }
        ustar[h] = bu;
      }
    }
  }
  int round = 0;
  for (; !LEAN; ++round) {
    if (round == 6) return -1;
    const unsigned long long MF = OM & fm;  // merging fronts, by front index
    // accepted backs, by back index
    const unsigned long long AB = tail_mode ? ((__brevll(~OM) >> (64 - W)) & bm) : ((~OM >> wf) & bm);
    const int popMF = __popcll(MF), popAB = __popcll(AB);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = h * 32 + (int)lane;
      int xf = INF, xb = INF;
      if (valid[h]) {
        if (t < wf) {
          const int m = __popcll(MF & below64(t));  // bursts before this front
          if (m == 0) xf = t;
          else if (m - 1 < popAB) xf = t + sel64(AB, m - 1) + 1;
        }
        if (tail_mode || t >= wf) {
          const int k = tail_mode ? (W - 1 - t) : (t - wf);
          const int g = __popcll(AB & below64(k));   // the burst this back belongs to
          if (g < popMF) xb = sel64(MF, g) + 1 + k;
        }
      }
      xs[h] = min(xf, xb);
      asfront[h] = xf <= xb;
    }
    // where the window ends: the next front or back the replay would need does not exist
    if (tail_mode) {
      n_ex = W;
      back_exhausted = false;
    } else {
      int t1 = INF, t2 = INF;
      if (popMF == 0) t1 = wf;
      else if (popMF - 1 < popAB) t1 = wf + sel64(AB, popMF - 1) + 1;
      if (popAB < popMF) t2 = sel64(MF, popAB) + 1 + wb;
      n_ex = min(t1, t2);
      back_exhausted = t2 < t1;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (valid[h]) sp.sx[h * 32 + lane] = (uint8_t)(xs[h] < n_ex ? xs[h] : 0xFF);
    __syncwarp();
    // unsure candidates: merge iff a pair-bit mate was accepted before their turn
    bool changed = false, newm[2] = {false, false};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (unsure[h] && xs[h] < n_ex) {
        unsigned long long pm = ((unsigned long long)phi[h] << 32) | (unsigned long long)plo[h];
        pm &= ~OM;  // accepted mates only
        int best = INF, bu = -1;
        while (pm) {
          const int u = __ffsll((long long)pm) - 1;
          pm &= pm - 1ull;
          const int xu = sp.sx[u];
          if (xu < xs[h] && xu < best) {
            best = xu;
            bu = u;
          }
        }
        newm[h] = bu >= 0;
        ustar[h] = bu;
      }
      const int t = h * 32 + (int)lane;
      if (unsure[h] && newm[h] != (((OM >> t) & 1ull) != 0ull)) changed = true;
    }
    __syncwarp();
    if (!__any_sync(0xffffffffu, changed)) break;
    const unsigned long long U = ballot64(unsure[0], unsure[1]);
    OM = (OM & ~U) | ballot64(unsure[0] && newm[0], unsure[1] && newm[1]);
  }
  if (A.dbg && lane == 0) tq1 = clock64();
  // ---- records in examine order ----
#pragma unroll
  for (int h = 0; h < 2; ++h)
    if (valid[h]) {
      const int t = h * 32 + (int)lane;
      if (xs[h] < n_ex) sp.xcand[xs[h]] = (uint8_t)t;
      sp.dense[t] = asfront[h] ? 1 : 0;  // (the scalar scan's table, free here) examined as a front candidate
      // the earliest accepted mate of an unsure merge, for the lanes that stand for examine indices below
      sp.xcls[t] = (uint8_t)(ustar[h] >= 0 && xs[h] < n_ex && ((OM >> t) & 1ull) ? ustar[h] : 0xFF);
    }
  __syncwarp();
  // lanes now stand for examine indices x = lane and lane + 32
  int tx[2], cid[2];
  bool ism[2], on[2], fr[2], idc[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int x = h * 32 + (int)lane;
    on[h] = x < n_ex;
    tx[h] = on[h] ? sp.xcand[x] : 0;
    ism[h] = on[h] && ((OM >> tx[h]) & 1ull) != 0ull;
    fr[h] = on[h] && sp.dense[tx[h]] != 0;
    idc[h] = on[h] && s.ctail[tx[h]] >= 0;
    cid[h] = 0xFF;
    if (ism[h]) {
      cid[h] = (s.s_f[tx[h]] != kInf) ? (int)sp.cls[tx[h]] : (kW + (int)sp.xcls[tx[h]]);
      atomicMin(&sp.cfirst[cid[h]], (uint32_t)x);
    }
  }
  __syncwarp();
  bool first[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) first[h] = ism[h] && sp.cfirst[cid[h]] == (uint32_t)(h * 32 + lane);
  unsigned long long FX = ballot64(first[0], first[1]);
  bool full = false;
  if (__popcll(FX) >= kKD) {  // the dirty cache fills up: the window ends right after the merge that takes its last entry
    n_ex = sel64(FX, kKD - 1) + 1;
    full = true;
    back_exhausted = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int x = h * 32 + (int)lane;
      on[h] = x < n_ex;
      ism[h] = ism[h] && on[h];
      first[h] = first[h] && on[h];
      fr[h] = fr[h] && on[h];
    }
    FX &= below64(n_ex);
  }
  const unsigned long long MX = ballot64(ism[0], ism[1]);                      // merges, by examine index
  const unsigned long long AX = ballot64(on[0] && !ism[0], on[1] && !ism[1]);  // accepts
  const unsigned long long FRX = ballot64(fr[0], fr[1]);
  const int nd = __popcll(FX), merges = __popcll(MX);
  // id-carrying merged candidates per target, by examine index (for the member-chain log)
#pragma unroll
  for (int h = 0; h < 2; ++h)
    if (ism[h] && idc[h]) atomicOr(&sp.cmask[cid[h]], 1ull << (h * 32 + lane));
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int x = h * 32 + (int)lane;
    if (!on[h]) continue;
    const int t = tx[h];
    const unsigned long long bx = below64(x);
    const int k = __popcll(MX & bx);
    sp.xep[x] = (uint8_t)k;
    sp.xfi[x] = (uint8_t)__popcll(FRX & bx);
    // accepted mates before x (and which of them were modified before x): only pair bits matter to the verify pass
    {
      unsigned long long pm = (((unsigned long long)s.pair[2 * t + 1] << 32) | (unsigned long long)s.pair[2 * t]) & ~OM;
      unsigned long long am = 0ull, dm = 0ull;
      while (pm) {
        const int u = __ffsll((long long)pm) - 1;
        pm &= pm - 1ull;
        if ((int)sp.sx[u] < x) {
          am |= 1ull << u;
          if (sp.cfirst[kW + u] < (uint32_t)x) dm |= 1ull << u;
        }
      }
      sp.acc_lo[x] = (uint32_t)am;
      sp.acc_hi[x] = (uint32_t)(am >> 32);
      sp.accd_lo[x] = (uint32_t)dm;
      sp.accd_hi[x] = (uint32_t)(dm >> 32);
    }
    if (!ism[h]) {
      const int r = __popcll(AX & bx);
      s.acc[r] = (uint32_t)t;
      sp.arank[t] = (uint8_t)r;
      sp.xtarget[x] = kInf;
      sp.xfe[x] = 0xFF;
      continue;
    }
    const int xf = (int)sp.cfirst[cid[h]];
    const int e = __popcll(FX & below64(xf));
    const bool hf = s.s_f[t] != kInf;
    sp.xfe[x] = (hf && xf != x) ? (uint8_t)e : (uint8_t)0xFF;
    sp.vcand[k] = (uint8_t)t;
    s.ment[t] = e;
    atomicOr(&sp.emask[e], 1ull << k);
    // the previous id-carrying member of the same target in examine order
    const unsigned long long pmask = sp.cmask[cid[h]] & bx;
    const int prev_last = pmask ? (int)sp.xcand[63 - __clzll((long long)pmask)] : -1;
    s.mprev[t] = prev_last;
    sp.vlast[k] = (int8_t)(idc[h] ? t : prev_last);
    if (first[h]) {  // the entry's value before its first merge
      if (hf) {
        s.dpos[e] = s.s_f[t];
        sp.dbase[e] = (uint32_t)t;
        sp.dbcnt[e] = s.pcnt[t];
        s.dridx[e] = s.pridx[t];
        s.dhead[e] = s.phead[t];
        s.dtail[e] = s.ptail[t];
      } else {
        const uint32_t u = (uint32_t)(cid[h] - kW);
        sp.dbase[e] = 0x100u | u;
        sp.dbcnt[e] = s.ccnt[u];
        s.dridx[e] = s.ridx[u];
        s.dhead[e] = s.chead[u];
        s.dtail[e] = s.ctail[u];
      }
    }
  }
  __syncwarp();
  // targets that are window candidates: position = i0 + acceptance rank of the mate (known now)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int x = h * 32 + (int)lane;
    if (!on[h] || !ism[h]) continue;
    const int t = tx[h];
    if (s.s_f[t] != kInf) {
      sp.xtarget[x] = s.s_f[t];
    } else {
      const uint32_t P = i0 + (uint32_t)sp.arank[cid[h] - kW];
      sp.xtarget[x] = P;
      if (first[h]) s.dpos[__popcll(FX & below64(x))] = P;
    }
  }
  if (lane == 0) {
    sp.xep[n_ex] = (uint8_t)merges;
    sp.xfi[n_ex] = (uint8_t)__popcll(FRX & below64(n_ex));
    s.ro[RO_BACK_EXH] = back_exhausted ? 1u : 0u;
    s.ro[RO_FULL] = full ? 1u : 0u;
    s.ro[RO_ND] = (uint32_t)nd;
    s.ro[RO_MERGES] = (uint32_t)merges;
    if (A.dbg) {
      atomicAdd(A.dbg + 36, (unsigned long long)(tq1 - tq0));
      atomicAdd(A.dbg + 37, (unsigned long long)(clock64() - tq1));
    }
  }
  __syncwarp();
  return n_ex;
}

// versions: the consensus value every predicted merge produces.  One entry's chain per warp, the entry's
// value in registers (lane d holds dimensions d and d + 32), the next candidate's operands requested
// before the current consensus is computed: the chain costs one multiply-divide-add per merge.
__device__ void spec_versions(const MergeArgs& A, Smem& s, Spec& sp, int warp, int nwarps) {
  const int D = A.D, ts = s.ts, rw = s.ts - 4;  // rw: 32 or 64 (speculation is off for wider rows)
  const uint32_t lane = lane_id();
  const int nd = (int)s.ro[RO_ND];
  const bool two = rw > 32;
  const int d0 = (int)lane, d1 = (int)lane + 32;
  for (int e = warp; e < nd; e += nwarps) {
    const uint32_t base = sp.dbase[e];
    const float* src = (base & 0x100u) ? s.tile + (size_t)(base & 0xFFu) * ts : s.pre + (size_t)base * ts;
    float r0 = src[d0], r1 = two ? src[d1] : 0.f;
    int c2 = sp.dbcnt[e];
    unsigned long long m = sp.emask[e];
    int k = __ffsll((long long)m) - 1;
    int t = sp.vcand[k];
    int c1 = s.ccnt[t];
    float x0 = s.tile[(size_t)t * ts + d0], x1 = two ? s.tile[(size_t)t * ts + d1] : 0.f;
    while (m) {
      m &= m - 1ull;
      // operands of the next merge of this entry
      const int kn = m ? (__ffsll((long long)m) - 1) : k;
      const int tn = sp.vcand[kn];
      const int c1n = s.ccnt[tn];
      const float x0n = s.tile[(size_t)tn * ts + d0], x1n = two ? s.tile[(size_t)tn * ts + d1] : 0.f;
      if (d0 < D) r0 = consensus1(x0, c1, r0, c2);
      if (two && d1 < D) r1 = consensus1(x1, c1, r1, c2);
      c2 += c1;
      float* dst = sp.vers + (size_t)k * ts;
      dst[d0] = r0;
      if (two) dst[d1] = r1;
      if (lane == 0) sp.vcnt[k] = c2;
      k = kn; t = tn; c1 = c1n; x0 = x0n; x1 = x1n;
    }
  }
}

// match: candidate examined at x against the current version of every entry modified before x.
template <int DR>
__device__ void spec_match(const MergeArgs& A, Smem& s, Spec& sp, int n_ex, int warp, int nwarps) {
  constexpr int NQ = DR / 4;
  const int D = A.D, ts = s.ts, nq = NQ > 0 ? NQ : (A.ld >> 2);
  const uint32_t lane = lane_id();
  const float thr = A.threshold;
  const float band = (4.f * (float)D + 40.f) * 5.9604645e-8f;
  const float thr_hi = thr + band, thr_lo = thr - band;
  const unsigned long long my_mask = sp.emask[lane];  // merges into this lane's entry
  for (int x = warp; x < n_ex; x += nwarps) {
    const int ep = sp.xep[x];
    const unsigned long long before = my_mask & ((1ull << ep) - 1ull);  // ep <= 63: a window has at most kW - 1 merges before a candidate
    const bool mine = before != 0ull;
    if (!__any_sync(0xffffffffu, mine)) {  // nothing modified yet
      if (lane == 0) sp.dmatch[x] = 0u;
      continue;
    }
    const int t = sp.xcand[x];
    const int k = mine ? (63 - __clzll((long long)before)) : 0;  // the entry's current version at x's turn (row 0 is valid filler)
    const float4* c4 = reinterpret_cast<const float4*>(s.tile + (size_t)t * ts);
    const float4* r4 = reinterpret_cast<const float4*>(sp.vers + (size_t)k * ts);
    const float cn = s.cnorm[t];
    float dot, nn;
    fast_dot_nn<NQ>(c4, r4, nq, dot, nn);
    const float rn = sqrtf(nn);
    const float prod = cn * rn;
    const float sim = __fdividef(dot, prod);
    const bool safe = prod > 1e-18f && prod < 1e18f;
    bool hit = false, amb = false;
    if (mine) {
      hit = safe && sim >= thr_hi;
      amb = !hit && !(safe && sim < thr_lo);
    }
    if (__any_sync(0xffffffffu, amb)) {  // inside the error band: the reference's own arithmetic decides
      float de = 0.f, ne = 0.f;
#pragma unroll 4
      for (int q = 0; q < nq; ++q) {
        const float4 xx = c4[q], y = r4[q];
        de = __fadd_rn(de, __fmul_rn(xx.x, y.x)); ne = __fadd_rn(ne, __fmul_rn(y.x, y.x));
        de = __fadd_rn(de, __fmul_rn(xx.y, y.y)); ne = __fadd_rn(ne, __fmul_rn(y.y, y.y));
        de = __fadd_rn(de, __fmul_rn(xx.z, y.z)); ne = __fadd_rn(ne, __fmul_rn(y.z, y.z));
        de = __fadd_rn(de, __fmul_rn(xx.w, y.w)); ne = __fadd_rn(ne, __fmul_rn(y.w, y.w));
      }
      if (amb) hit = cos_match(de, cn, __fsqrt_rn(ne), thr);
    }
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) sp.dmatch[x] = m;
  }
}

// verify + cut: returns the number of leading examined candidates whose prediction is the true decision.
__device__ int spec_verify(Smem& s, Spec& sp, int n_ex, uint32_t i0) {
  const uint32_t lane = lane_id();
  int cut = n_ex;
  for (int half = 0; half < 2; ++half) {
    const int x = half * 32 + (int)lane;
    bool ok = true;
    if (x < n_ex) {
      const int t = sp.xcand[x];
      const uint32_t P = sp.xtarget[x];
      uint32_t best = kInf;
      uint32_t dm = sp.dmatch[x];
      const uint32_t dm_all = dm;
      while (dm) {  // modified representatives whose current value matches
        const int e = __ffs(dm) - 1;
        dm &= dm - 1;
        best = min(best, s.dpos[e]);
      }
      const uint32_t fpos = s.s_f[t];
      bool undecidable = false;
      if (fpos != kInf) {
        const int fe = sp.xfe[x];
        if (fe == 0xFF) best = min(best, fpos);  // the first old match is unmodified: it still matches
        else if (!((dm_all >> fe) & 1u) && best > fpos) undecidable = true;  // a clean match behind it cannot be ruled out
      }
      // accepted in this window before x and not modified since: the precomputed pair bits hold
      uint32_t plo = s.pair[2 * t] & sp.acc_lo[x] & ~sp.accd_lo[x];
      uint32_t phi = s.pair[2 * t + 1] & sp.acc_hi[x] & ~sp.accd_hi[x];
      uint32_t rank = 0xFFu;
      while (plo) {
        const int u = __ffs(plo) - 1;
        plo &= plo - 1;
        rank = min(rank, (uint32_t)sp.arank[u]);
      }
      while (phi) {
        const int u = 32 + __ffs(phi) - 1;
        phi &= phi - 1;
        rank = min(rank, (uint32_t)sp.arank[u]);
      }
      if (rank != 0xFFu) best = min(best, i0 + rank);
      ok = !undecidable && best == P;
    }
    const uint32_t bad = __ballot_sync(0xffffffffu, !ok);
    if (bad && cut == n_ex) cut = half * 32 + (__ffs(bad) - 1);
  }
  return cut;
}

// commit the verified prefix [0, cut): the state the sequential loop would have reached there, in the form
// flush_window expects.  All threads of the CTA.
template <int TEAM>
__device__ void spec_commit(const MergeArgs& A, Smem& s, Spec& sp, int n_ex, int cut, uint32_t i0, uint32_t size0) {
  constexpr int kMT = Shape<TEAM>::kMT;
  const int tid = threadIdx.x, ts = s.ts, ld = A.ld;
  // the state before examine index `cut`: merges and front candidates so far say everything
  const int merges = (int)sp.xep[cut], fi = (int)sp.xfi[cut], all_merges = (int)s.ro[RO_MERGES], nd_all = (int)s.ro[RO_ND];
  const int a = cut - merges, bi = cut - fi;
  const bool from_back = cut > 0 && sp.xtarget[cut - 1] != kInf;
  const unsigned long long below = merges >= 64 ? ~0ull : ((1ull << merges) - 1ull);
  // entries are allocated in the order of their first merge: those with a merge before the cut exist
  int nd = 0;
  for (int e = 0; e < nd_all; ++e) nd += (sp.emask[e] & below) != 0ull ? 1 : 0;
  // merges beyond the cut did not happen
  for (int k = merges + tid; k < all_merges; k += kMT) s.ment[sp.vcand[k]] = -1;
  if (tid < nd) {
    const int k = 63 - __clzll((long long)(sp.emask[tid] & below));  // the entry's last version before the cut
    s.dcnt[tid] = sp.vcnt[k];
    s.dlast[tid] = (int32_t)sp.vlast[k];
  }
  for (int idx = tid; idx < nd * ld; idx += kMT) {
    const int e = idx / ld, d = idx - e * ld;
    const int k = 63 - __clzll((long long)(sp.emask[e] & below));
    s.dvals[(size_t)e * ts + d] = sp.vers[(size_t)k * ts + d];
  }
  __syncthreads();  // everybody has read RO_ND / RO_MERGES before they are rewritten
  if (tid == 0) {
    const bool whole = cut == n_ex;
    s.ro[RO_A] = (uint32_t)a;
    s.ro[RO_ND] = (uint32_t)nd;
    s.ro[RO_I] = i0 + (uint32_t)a;
    s.ro[RO_SIZE] = size0 - (uint32_t)merges;
    s.ro[RO_FROM_BACK] = from_back ? 1u : 0u;
    s.ro[RO_BI] = (uint32_t)bi;
    s.ro[RO_MERGES] = (uint32_t)merges;
    if (!whole) {
      s.ro[RO_BACK_EXH] = 0u;
      s.ro[RO_FULL] = 0u;
    }
    s.ro[RO_UNDEC] = whole ? 0u : 1u;
    s.ro[RO_EXAMINED] = (uint32_t)cut;
  }
}

// Apply the window's effects to global memory with the whole CTA.
template <int TEAM>
__device__ void flush_window(const MergeArgs& A, uint32_t* seg, float* pos_nrm, uint4* seg_h, int qh, TeamCtl* ctl, Smem& s,
                             int W, int wf, int wb, bool tail_mode, uint32_t i0, uint32_t size0) {
  constexpr int kMT = Shape<TEAM>::kMT;
  const int D = A.D, ld = A.ld, nq = ld >> 2;
  const int tid = threadIdx.x;
  const int a = (int)s.ro[RO_A], nd = (int)s.ro[RO_ND], bi = (int)s.ro[RO_BI], merges = (int)s.ro[RO_MERGES];
  const uint32_t i = s.ro[RO_I], size = s.ro[RO_SIZE];
  const bool from_back = s.ro[RO_FROM_BACK] != 0u, back_exhausted = s.ro[RO_BACK_EXH] != 0u;
  // norms of the modified representatives' final values
  if (tid < nd) s.dnorm[tid] = norm_seq(reinterpret_cast<const float4*>(s.dvals + (size_t)tid * s.ts), nq);
  // a moved tail element waiting at position i (the window ended right after a merge); its source may
  // lie in the range the sentinels overwrite below, so it is read before the barrier
  if (tid == kMT - 1 && from_back && i < size) {
    uint32_t moved;
    if (tail_mode) moved = s.ridx[wf - 1 - bi];
    else if (bi < wb) moved = s.ridx[wf + bi];
    else moved = __ldcg(seg + size);  // the element that was at the old tail position
    seg[i] = moved;
  }
  // accepted candidates: positions i0.. in acceptance order (row index + norm + fp16 copy)
  for (int k = tid; k < a; k += kMT) {
    const uint32_t u = s.acc[k];
    seg[i0 + k] = s.ridx[u];
    pos_nrm[i0 + k] = s.cnorm[u];
  }
  if (qh) {
    for (int idx = tid; idx < a * qh; idx += kMT) {
      const int k = idx / qh, c = idx - k * qh;
      seg_h[(size_t)(i0 + k) * qh + c] = h16_chunk_from_htile(s.htile + (size_t)s.acc[k] * s.hs, c);
    }
  }
  __syncthreads();  // an accepted candidate modified later in the window is rewritten below
  for (uint32_t k = size + (uint32_t)tid; k < size0; k += kMT) seg[k] = KLSH_SENTINEL;
  if (qh) {
    // modified representatives are re-scaled from their current value and norm
    for (int idx = tid; idx < nd * qh; idx += kMT) {
      const int e = idx / qh, c = idx - e * qh;
      seg_h[(size_t)s.dpos[e] * qh + c] = h16_chunk_from_row(s.dvals + (size_t)e * s.ts, ld, s.dnorm[e], c);
    }
  }
  // member chains: ids(current) ++ ids(candidate) for every logged merge
  if (tid < W && s.ment[tid] >= 0) {
    const int e = s.ment[tid], prev = s.mprev[tid];
    const int t1 = s.ctail[tid];
    if (t1 >= 0) {
      const int nv = (prev < 0) ? s.dhead[e] : s.chead[prev];
      A.next[t1] = nv;
      if (A.mg.counts) {
        const uint32_t k = atomicAdd(A.mg.counts + 1, 1u);
        A.mg.next_slot[k] = (uint32_t)t1;
        A.mg.next_val[k] = nv;
      }
    }
  }
  // modified representatives: count, head, tail, norm ...
  if (tid < nd) {
    const int e = tid;
    const uint32_t rr = s.dridx[e];
    const int last = s.dlast[e];
    if (A.mg.counts) A.mg.mod_rows[atomicAdd(A.mg.counts, 1u)] = rr;
    A.cnt[rr] = s.dcnt[e];
    // each merge prepends the candidate's members: the chain now starts with the LAST merged
    // candidate that carried ids
    if (last >= 0) A.head[rr] = s.chead[last];
    if (s.dtail[e] < 0 && last >= 0) {  // the representative had no ids: its tail is the EARLIEST such candidate's
      int first_t = last;
      for (int tt = s.mprev[last]; tt >= 0; tt = s.mprev[tt]) first_t = tt;
      A.tail[rr] = s.ctail[first_t];
    }
    pos_nrm[s.dpos[e]] = s.dnorm[e];
  }
  // ... and values
  for (int idx = tid; idx < nd * D; idx += kMT) {
    const int e = idx / D, d = idx - e * D;
    A.vals[(uint64_t)s.dridx[e] * ld + d] = s.dvals[(size_t)e * s.ts + d];
  }
  if (tid == 0) {
    ctl->i = i;
    ctl->size = size;
    // back candidates for the next window: grow fast when they ran out, shrink slowly otherwise (a
    // window that runs out of them throws away the screen of its remaining front candidates)
    ctl->wb = back_exhausted ? (uint32_t)min(kWbMax, max(wb * 2, 4))
                             : (uint32_t)min(kWbMax, max(max(2, merges + merges / 2 + 2), wb - (wb + 3) / 4));
    if (A.dbg) {
      atomicAdd(A.dbg + 0, 1ull);
      atomicAdd(A.dbg + 1, (unsigned long long)s.ro[RO_EXAMINED]);
      atomicAdd(A.dbg + 2, (unsigned long long)merges);
      atomicAdd(A.dbg + 3, (unsigned long long)s.ro[RO_UNDEC]);
      atomicAdd(A.dbg + 4, (unsigned long long)s.ro[RO_FULL]);
      atomicAdd(A.dbg + 5, (unsigned long long)(back_exhausted ? 1 : 0));
      atomicAdd(A.dbg + 6, (unsigned long long)a);
    }
  }
}

// Window staging.  Rows go global -> shared with cp.async (no registers, no wait until the whole
// batch is in flight), so a stage costs two dependent round trips (row index -> row / metadata)
// however many rows a thread moves.

// ---- screen pool: worker side ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_acq_rel(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
// the board entry is taken down by whoever sees the last chunk go (or a stale window): only the value that was seen
__device__ __forceinline__ void pool_close(const MergeArgs& A, uint32_t leader, uint32_t board_val) {
  if (atomicCAS(A.pool_board + leader, board_val, 0u) == board_val) atomicSub(&A.pool->open, 1u);
}

// Exact tests of the pairs a screen parked in s.surv: all of the CTA's threads at once, so the dependent row
// fetches of different pairs overlap.  REMOTE: the window belongs to another CTA (pub), results go to s.w_f.
template <int KMT, bool REMOTE>
__device__ __forceinline__ void exact_parked(const MergeArgs& A, Smem& s, const uint32_t* seg, const float* pos_nrm, int nq,
                                             const PoolPub* pub) {
  const int tid = threadIdx.x, ld = A.ld;
  uint32_t* const filt = REMOTE ? s.w_f : s.s_f;
  const uint32_t ns = min(s.surv[kSurvCap], (uint32_t)kSurvCap);
  if (tid == 0 && ns && A.work) atomicAdd(A.work + 1, (unsigned long long)ns);
  for (uint32_t k0 = 0; k0 < ns; k0 += 2 * KMT) {
    uint32_t ent[2], rr2[2];
    bool on[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t k = k0 + u * KMT + tid;
      on[u] = k < ns;
      ent[u] = on[u] ? s.surv[k] : 0u;
      on[u] = on[u] && !(filt[ent[u] & 63u] < (ent[u] >> 6));
      rr2[u] = on[u] ? __ldcg(seg + (ent[u] >> 6)) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (on[u]) {
        const uint32_t jj = ent[u] >> 6;
        const int t = (int)(ent[u] & 63u);
        const float* rowp = A.vals + (uint64_t)rr2[u] * ld;
        const float rn = __ldcg(pos_nrm + jj);
        if (REMOTE ? exact_pair_remote(A, pub, t, rowp, rn, nq) : exact_pair<false>(s, t, rowp, rn, nq, A.threshold))
          atomicMin(&filt[t], jj);
      }
  }
  if (ns) __syncthreads();
}

// Serve one chunk of another leader's window.  All threads.
template <int KMT, int DR>
__device__ void pool_serve(const MergeArgs& A, Smem& s, uint32_t leader, uint32_t chunk) {
  constexpr int KS16 = DR / 16, QH = DR / 8;
  const int tid = threadIdx.x, nq = A.ld >> 2;
  PoolPub* pub = A.pool_pub + leader;
  const int W = (int)__ldcg(&pub->W);
  const uint32_t st = __ldcg(&pub->st), i0 = __ldcg(&pub->i0);
  const uint32_t j_begin = chunk * kPoolChunk, j_end = min(i0, (chunk + 1u) * kPoolChunk);
  const uint32_t* seg = A.rows_sorted + st;
  const float* pos_nrm = A.pos_nrm + st;
  const uint4* seg_h = A.pos_h + (size_t)st * QH;
  if (tid < kW) s.w_f[tid] = kInf;
  if (tid == 0) s.surv[kSurvCap] = 0u;
  __syncthreads();
  tc_compare<KS16, 2>(A, seg, pos_nrm, seg_h, s, W, j_begin, j_end, (uint32_t)(tid >> 5), KMT / 32, nq, pub);
  __syncthreads();
  exact_parked<KMT, true>(A, s, seg, pos_nrm, nq, pub);
  if (tid < W && s.w_f[tid] != kInf) atomicMin(&pub->f[tid], s.w_f[tid]);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    atom_add_acq_rel(&pub->done, 1u);
    if (A.dbg) atomicAdd(A.dbg + 34, 1ull);
  }
}

// One step of a CTA that has nothing of its own to do: find an open window on the board, claim a chunk, serve it.
// Returns false when there was nothing to serve.  All threads; s.ptask holds the broadcast slots.
template <int KMT, int DR>
__device__ __forceinline__ bool pool_help(const MergeArgs& A, Smem& s, uint32_t self) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    s.ptask[0] = 0u;
    s.ptask[3] = ld_acquire(&A.pool->open) ? 0xFFFFFFFFu : 0xFFFFFFFEu;  // FFFFFFFE: nothing open, do not scan
  }
  __syncthreads();
  if (s.ptask[3] == 0xFFFFFFFEu) {
    __syncthreads();
    return false;
  }
  // scan the board from a CTA-specific offset so that helpers spread over the open windows
  const uint32_t grid = A.pool_n, off = (blockIdx.x * 37u) % grid;
  for (uint32_t k = (uint32_t)tid; k < grid; k += KMT) {
    uint32_t l = k + off;
    if (l >= grid) l -= grid;
    if (l != self && __ldcg(A.pool_board + l) != 0u) atomicMin(&s.ptask[3], k);
  }
  __syncthreads();
  if (tid == 0 && s.ptask[3] < grid) {
    uint32_t l = s.ptask[3] + off;
    if (l >= grid) l -= grid;
    PoolPub* pub = A.pool_pub + l;
    const uint32_t bv = ld_acquire(A.pool_board + l);
    if (bv != 0u) {
      const uint32_t v = atom_add_acq_rel(&pub->next, 1u);
      const uint32_t e = v >> kPoolCountBits, k = v & ((1u << kPoolCountBits) - 1u);
      const uint32_t meta = ld_acquire(&pub->meta);
      const uint32_t nt = meta & ((1u << kPoolCountBits) - 1u);
      // a claim counts only in the window it was made in; then the leader is waiting for it and the block is stable
      const bool mine = (meta >> kPoolCountBits) == e && k < nt;
      if (!mine || k + 1u >= nt) pool_close(A, l, bv);
      if (mine) {
        s.ptask[0] = 1u;
        s.ptask[1] = l;
        s.ptask[2] = k;
      }
    }
  }
  __syncthreads();
  const uint32_t got = s.ptask[0], leader = s.ptask[1], chunk = s.ptask[2];
  __syncthreads();
  if (!got) return false;
  pool_serve<KMT, DR>(A, s, leader, chunk);
  return true;
}

// ---- one bucket, one team.  Returns true if the bucket was handed on to the next team. -----------------
template <int TEAM, int DR>
__device__ bool merge_team(const MergeArgs& A, uint32_t bucket, uint32_t start_i, uint32_t start_size, TeamCtl* ctl, Smem& s) {
  constexpr int kMT = Shape<TEAM>::kMT;
  const int ld = A.ld, nq = ld >> 2;
  const int tid = threadIdx.x;
  const uint32_t lane = lane_id(), warp = tid >> 5;
  const bool leader = Team<TEAM>::rank() == 0;
  const uint32_t st = A.bstart[bucket];
  uint32_t* seg = A.rows_sorted + st;
  float* pos_nrm = A.pos_nrm + st;
  const int QH = DR > 0 ? DR / 8 : (s.hs - 8) / 8;  // 16-byte chunks of the fp16 copy per representative
  uint4* seg_h = A.pos_h + (size_t)st * QH;
  // every CTA must have left the previous bucket's loop before the control block is reused
  if (TEAM == 2) Team<TEAM>::sync();
  if (leader && warp == 0) {
    if (start_i == 0) {  // fresh bucket: representative 0 is the first row
      const uint32_t r0 = seg[0];
      const float* src = A.vals + (uint64_t)r0 * ld;
      for (int d = lane; d < ld; d += 32) s.dvals[d] = src[d];
      __syncwarp();
      const float n0 = norm_seq(reinterpret_cast<const float4*>(s.dvals), nq);
      for (int c = (int)lane; c < QH; c += 32) seg_h[c] = h16_chunk_from_row(s.dvals, ld, n0, c);
      if (lane == 0) {
        pos_nrm[0] = n0;
        ctl->i = 1;
        ctl->size = A.bstart[bucket + 1] - st;
        ctl->wb = 4;
        ctl->mode = 0u;
        ctl->big = 0u;
      }
    } else if (lane == 0) {
      ctl->i = start_i;
      ctl->size = start_size;
      ctl->wb = 4;
      ctl->mode = 0u;
      ctl->big = 0u;
    }
    if (TEAM != 0)
      for (int t = lane; t < kW; t += 32) ctl->f[t] = kInf;
  }
  __threadfence();
  Team<TEAM>::sync();
  for (;;) {
    const uint32_t i0 = __ldcg(&ctl->i), size0 = __ldcg(&ctl->size);
    const uint32_t mode0 = __ldcg(&ctl->mode);  // read here by everybody: the leader rewrites it while resolving
    if constexpr (TEAM == 1 && DR > 0) {
      // helpers are worth keeping resident only for buckets that collect many representatives: announce early
      if (A.pool && leader && tid == 0) {
        const bool over = i0 >= size0 || (i0 > A.max_reps && A.esc_list);
        if (over && ctl->big) {
          atomicSub(&A.pool->big_active, 1u);
          ctl->big = 0u;
        } else if (!over && !ctl->big && i0 >= A.pool_min / 4u) {
          atomicAdd(&A.pool->big_active, 1u);
          ctl->big = 1u;
        }
      }
    }
    if (i0 >= size0) return false;
    const uint32_t remaining = size0 - i0;
    if (i0 > A.max_reps && A.esc_list) {  // more compare work than this team should carry: hand on
      if (leader && tid == 0) {
        const uint32_t k = atomicAdd(A.esc_count, 1u);
        A.esc_list[3 * k] = bucket;
        A.esc_list[3 * k + 1] = i0;
        A.esc_list[3 * k + 2] = size0;
        if (A.dbg) atomicAdd(A.dbg + 7, 1ull);
      }
      return true;
    }
    bool tail_mode;
    int wf, wb;
    if (remaining <= (uint32_t)kW) {
      tail_mode = true;
      wf = (int)remaining;
      wb = 0;
    } else {
      tail_mode = false;
      wb = (int)min(__ldcg(&ctl->wb), (uint32_t)kWbMax);
      wf = kW - wb;
    }
    const int W = wf + wb;
    if (leader && tid == 0 && A.work) atomicAdd(A.work, (unsigned long long)W * ((unsigned long long)i0 + (unsigned long long)W));
    long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0, tk4 = 0, tk5 = 0;
    const bool prof = A.dbg && leader && tid == 0;
    if (prof) tk0 = clock64();
    // ---- stage the window: rows, metadata, norms ----
    // Every load is issued before the first result is used, so the staging costs two dependent
    // round trips (row index -> row / metadata) however many rows a thread moves.
    long long ts1 = 0, ts2 = 0, ts3 = 0;
    {
      auto pos_of = [&](int t) { return (t < wf) ? (i0 + (uint32_t)t) : (size0 - 1 - (uint32_t)(t - wf)); };
      uint32_t mr = 0u;
      if (tid < W) mr = __ldcg(seg + pos_of(tid));
      const int items = W * nq;
      for (int base = 0; base < items; base += 4 * kMT) {
        uint32_t rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int v = base + u * kMT + tid;
          rr[u] = (v < items) ? __ldcg(seg + pos_of(v / nq)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int v = base + u * kMT + tid;
          if (v < items) {
            const int t = v / nq, q = v - t * nq;
            cp_async16(reinterpret_cast<float4*>(s.tile + (size_t)t * s.ts) + q,
                       reinterpret_cast<const float4*>(A.vals + (uint64_t)rr[u] * ld) + q);
          }
        }
      }
      if (tid < W) {
        s.ridx[tid] = mr;
        if (leader) {
          const int4 m = __ldcg(reinterpret_cast<const int4*>(A.cnt.p) + mr);  // one 16-byte record {cnt, head, tail, 0}
          s.ccnt[tid] = m.x;
          s.chead[tid] = m.y;
          s.ctail[tid] = m.z;
        }
      }
      cp_async_wait_all();
    }
    if (tid == 0) s.surv[kSurvCap] = 0u;
    if (tid < kW) {
      s.s_f[tid] = kInf;
      s.pair[2 * tid] = 0u;
      s.pair[2 * tid + 1] = 0u;
      s.ment[tid] = -1;
      s.mprev[tid] = -1;
    }
    if (tid < kKD) {
      s.dver[tid] = 0u;
      s.mver[tid] = 0u;
    }
    if (tid == 0) {
      s.ro[RO_ND] = 0u;
      s.ro[RO_DONE] = 0u;
    }
    __syncthreads();
    if (prof) ts1 = clock64();
    if (prof) ts2 = ts1;
    if (tid < W) s.cnorm[tid] = norm_seq(reinterpret_cast<const float4*>(s.tile + (size_t)tid * s.ts), nq);
    __syncthreads();
    if (prof) ts3 = clock64();
    {
      // unit-norm fp16 copy of the window, zero-padded to the k-step width (tensor-core screen);
      // one thread converts 8 consecutive elements and stores them with one 16-byte write
      const int kw = DR > 0 ? DR : s.hs - 8;
      const int cpr = kw >> 3;
      for (int v = tid; v < kW * cpr; v += kMT) {
        const int t = v / cpr, d0 = (v - t * cpr) * 8;
        const float nrm = t < W ? s.cnorm[t] : 1.f;
        const float* src = s.tile + (size_t)t * s.ts + d0;
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int d = d0 + 2 * q;
          const float x0 = (t < W && d < ld) ? unit_scale(src[2 * q], nrm) : 0.f;
          const float x1 = (t < W && d + 1 < ld) ? unit_scale(src[2 * q + 1], nrm) : 0.f;
          w[q] = pack_half2(x0, x1);
        }
        *reinterpret_cast<uint4*>(s.htile + (size_t)t * s.hs + d0) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      __syncthreads();
    }
    if (prof) tk1 = clock64();
    // ---- parallel phase: old representatives [0, i0) across the team, then candidate x candidate bits ----
    bool pooled = false;
    if constexpr (DR > 0 && TEAM == 0) pooled = A.pool != nullptr && i0 >= A.pool_min;
    if constexpr (DR > 0 && TEAM == 0) {
      if (pooled) {
        // publish the window and open its screen; then the candidate x candidate bits, this window's chunks as
        // long as there are unclaimed ones, and other windows' chunks until all of this window's are done
        constexpr int KS16 = DR / 16;
        PoolPub* pub = A.pool_pub + blockIdx.x;
        const uint32_t ntask = (i0 + kPoolChunk - 1) / kPoolChunk;
        for (int idx = tid; idx < 4 * KS16 * 32; idx += kMT) {
          const int l = idx & 31, ks = (idx >> 5) % KS16, mt = (idx >> 5) / KS16;
          const __half* r0 = s.htile + (size_t)(mt * 16 + (l >> 2)) * s.hs + ks * 16 + (l & 3) * 2;
          const __half* r1 = r0 + 8 * s.hs;
          pub->af[idx] = make_uint4(*reinterpret_cast<const uint32_t*>(r0), *reinterpret_cast<const uint32_t*>(r1),
                                    *reinterpret_cast<const uint32_t*>(r0 + 8), *reinterpret_cast<const uint32_t*>(r1 + 8));
        }
        if (tid < kW) {
          pub->ridx[tid] = tid < W ? s.ridx[tid] : 0u;
          pub->cnorm[tid] = tid < W ? s.cnorm[tid] : 1.f;
          pub->f[tid] = kInf;
        }
        if (tid == 0) {
          pub->W = (uint32_t)W;
          pub->st = st;
          pub->i0 = i0;
          pub->done = 0u;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
          const uint32_t epoch = (s.ptask[6] + 1u) & 0xFFFu;
          s.ptask[6] = epoch;
          st_release(&pub->meta, (epoch << kPoolCountBits) | ntask);
          st_release(&pub->next, epoch << kPoolCountBits);
          if (ntask > 1u) {
            st_release(A.pool_board + blockIdx.x, epoch + 1u);
            atomicAdd(&A.pool->open, 1u);
          }
          if (A.dbg) atomicAdd(A.dbg + 35, (unsigned long long)ntask);
        }
        tc_compare<KS16, 1>(A, seg, pos_nrm, seg_h, s, W, 0u, (uint32_t)W, warp, kMT / 32, nq);
        __syncthreads();
        uint32_t own = 0;
        for (;;) {  // own chunks, screened from shared memory
          if (tid == 0) s.ptask[4] = atom_add_acq_rel(&pub->next, 1u) & ((1u << kPoolCountBits) - 1u);
          __syncthreads();
          const uint32_t k = s.ptask[4];
          __syncthreads();
          if (k + 1u >= ntask && ntask > 1u && tid == 0) pool_close(A, blockIdx.x, s.ptask[6] + 1u);
          if (k >= ntask) break;
          tc_compare<KS16, 0>(A, seg, pos_nrm, seg_h, s, W, k * kPoolChunk, min(i0, (k + 1u) * kPoolChunk), warp, kMT / 32, nq);
          ++own;
        }
        __syncthreads();
        exact_parked<kMT, false>(A, s, seg, pos_nrm, nq, nullptr);
        __syncthreads();
        for (;;) {
          if (tid == 0) s.ptask[5] = ld_acquire(&pub->done);
          __syncthreads();
          const bool all_done = s.ptask[5] + own == ntask;
          __syncthreads();
          if (all_done) break;
          if (!pool_help<kMT, DR>(A, s, blockIdx.x)) __nanosleep(60);
        }
        if (tid < W) s.s_f[tid] = min(s.s_f[tid], __ldcg(&pub->f[tid]));
        if (tid == 0) s.surv[kSurvCap] = 0u;
      }
    }
    uint32_t pool_ntask = 0u;
    if constexpr (DR > 0 && TEAM == 1) {
      pooled = A.pool != nullptr && i0 >= A.pool_min;
      if (pooled) {
        // The cluster's CTAs claim chunks of the screen like everybody else; idle teams of the launch that stayed
        // on as helpers take the rest.  Every CTA of the cluster counts the pooled windows itself (s.ptask[6]).
        constexpr int KS16 = DR / 16;
        PoolPub* pub = A.pool_pub + Team<TEAM>::id();
        const uint32_t ntask = (i0 + kPoolChunk - 1) / kPoolChunk;
        pool_ntask = ntask;
        const uint32_t epoch = (s.ptask[6] + 1u) & 0xFFFu;
        __syncthreads();
        if (tid == 0) s.ptask[6] = epoch;
        if (leader) {
          for (int idx = tid; idx < 4 * KS16 * 32; idx += kMT) {
            const int l = idx & 31, ks = (idx >> 5) % KS16, mt = (idx >> 5) / KS16;
            const __half* r0 = s.htile + (size_t)(mt * 16 + (l >> 2)) * s.hs + ks * 16 + (l & 3) * 2;
            const __half* r1 = r0 + 8 * s.hs;
            pub->af[idx] = make_uint4(*reinterpret_cast<const uint32_t*>(r0), *reinterpret_cast<const uint32_t*>(r1),
                                      *reinterpret_cast<const uint32_t*>(r0 + 8), *reinterpret_cast<const uint32_t*>(r1 + 8));
          }
          if (tid < kW) {
            pub->ridx[tid] = tid < W ? s.ridx[tid] : 0u;
            pub->cnorm[tid] = tid < W ? s.cnorm[tid] : 1.f;
            pub->f[tid] = kInf;
          }
          if (tid == 0) {
            pub->W = (uint32_t)W;
            pub->st = st;
            pub->i0 = i0;
            pub->done = 0u;
          }
          __threadfence();
          __syncthreads();
          if (tid == 0) {
            st_release(&pub->meta, (epoch << kPoolCountBits) | ntask);
            st_release(&pub->next, epoch << kPoolCountBits);
            st_release(A.pool_board + Team<TEAM>::id(), epoch + 1u);
            atomicAdd(&A.pool->open, 1u);
            if (A.dbg) atomicAdd(A.dbg + 35, (unsigned long long)ntask);
          }
          tc_compare<KS16, 1>(A, seg, pos_nrm, seg_h, s, W, 0u, (uint32_t)W, warp, kMT / 32, nq);
          __syncthreads();
        } else {
          if (tid == 0)
            while ((ld_acquire(&pub->meta) >> kPoolCountBits) != epoch) __nanosleep(20);
          __syncthreads();
        }
        uint32_t own = 0;
        for (;;) {
          if (tid == 0) {
            uint32_t v = atom_add_acq_rel(&pub->next, 1u);
            while ((v >> kPoolCountBits) != epoch) {  // the leader has published the window but not yet reset the claim word
              __nanosleep(20);
              v = atom_add_acq_rel(&pub->next, 1u);
            }
            s.ptask[4] = v & ((1u << kPoolCountBits) - 1u);
          }
          __syncthreads();
          const uint32_t k = s.ptask[4];
          __syncthreads();
          if (k + 1u >= ntask && tid == 0) pool_close(A, Team<TEAM>::id(), epoch + 1u);
          if (k >= ntask) break;
          tc_compare<KS16, 0>(A, seg, pos_nrm, seg_h, s, W, k * kPoolChunk, min(i0, (k + 1u) * kPoolChunk), warp, kMT / 32, nq);
          ++own;
        }
        if (own) {
          __threadfence();
          if (tid == 0) s.ptask[5] = own;
        } else if (tid == 0) {
          s.ptask[5] = 0u;
        }
      }
    }
    if (pooled) {
    } else if constexpr (DR > 0) {
      constexpr int KS16 = DR / 16;
      tc_compare<KS16, 0>(A, seg, pos_nrm, seg_h, s, W, 0u, i0, Team<TEAM>::rank() * (kMT / 32) + warp,
                              Team<TEAM>::ncta() * (kMT / 32), nq);
      if (leader) tc_compare<KS16, 1>(A, seg, pos_nrm, seg_h, s, W, 0u, (uint32_t)W, warp, kMT / 32, nq);
    } else {
      const int ks16 = (s.hs - 8) >> 4;
      tc_compare_wide<false>(A, seg, pos_nrm, seg_h, s, W, 0u, i0, Team<TEAM>::rank() * (kMT / 32) + warp,
                             Team<TEAM>::ncta() * (kMT / 32), nq, ks16);
      if (leader) tc_compare_wide<true>(A, seg, pos_nrm, seg_h, s, W, 0u, (uint32_t)W, warp, kMT / 32, nq, ks16);
    }
    __syncthreads();
    exact_parked<kMT, false>(A, s, seg, pos_nrm, nq, nullptr);
    if (prof) tk2 = clock64();
    if (TEAM != 0) {
      if (tid < W && s.s_f[tid] != kInf) atomicMin(&ctl->f[tid], s.s_f[tid]);
      __threadfence();
      if constexpr (DR > 0 && TEAM == 1) {
        if (pooled) {  // this CTA's chunks are finished (their matches are in ctl->f above)
          __syncthreads();
          if (tid == 0 && s.ptask[5]) atom_add_acq_rel(&(A.pool_pub + Team<TEAM>::id())->done, s.ptask[5]);
        }
      }
      Team<TEAM>::sync();
    }
    if (prof) tk3 = clock64();
    if (leader) {
      if (TEAM != 0) {
        if constexpr (DR > 0 && TEAM == 1) {
          if (pooled) {  // chunks taken by helpers outside the cluster
            const PoolPub* pub = A.pool_pub + Team<TEAM>::id();
            if (tid == 0)
              while (ld_acquire(&pub->done) != pool_ntask) __nanosleep(40);
            __syncthreads();
            if (tid < W) atomicMin(&ctl->f[tid], __ldcg(&pub->f[tid]));
            __syncthreads();
          }
        }
        if (tid < W) {
          s.s_f[tid] = __ldcg(&ctl->f[tid]);
          ctl->f[tid] = kInf;
        }
        __syncthreads();
      }
      // prefetch every candidate's first-match representative (row + member metadata) in parallel
      {
        const bool has = tid < W && s.s_f[tid] != kInf;
        uint32_t mr = 0u;
        if (has) mr = __ldcg(seg + s.s_f[tid]);
        const int items = W * nq;
        for (int base = 0; base < items; base += 4 * kMT) {
          uint32_t rr[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int v = base + u * kMT + tid;
            const uint32_t p = (v < items) ? s.s_f[v / nq] : kInf;
            rr[u] = (p != kInf) ? __ldcg(seg + p) : 0xFFFFFFFFu;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int v = base + u * kMT + tid;
            if (rr[u] != 0xFFFFFFFFu) {
              const int t = v / nq, q = v - t * nq;
              cp_async16(reinterpret_cast<float4*>(s.pre + (size_t)t * s.ts) + q,
                         reinterpret_cast<const float4*>(A.vals + (uint64_t)rr[u] * ld) + q);
            }
          }
        }
        if (has) {
          s.pridx[tid] = mr;
          const int4 m = __ldcg(reinterpret_cast<const int4*>(A.cnt.p) + mr);
          s.pcnt[tid] = m.x;
          s.phead[tid] = m.y;
          s.ptail[tid] = m.z;
        }
        cp_async_wait_all();
      }
      __syncthreads();
      if (prof) tk4 = clock64();
      const bool sequential = DR == 0 || mode0 != 0u || A.no_spec;
      if (!sequential) {
        // speculative resolution (see above); the cp.async ring of the screen is idle now and holds its scratch
        Spec sp;
        carve_spec(sp, s.ring, s.ts);
        long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0, tp4 = 0;
        if (prof) tp0 = clock64();
        if (warp == 0) {
          int n = -1;
          const int sm = A.scan_mode == 3 ? (TEAM == 0 ? 0 : 2) : A.scan_mode;
          if (sm == 2) n = spec_scan_par<true>(A, s, sp, W, wf, wb, tail_mode, i0);
          else if (sm == 1) n = spec_scan_par<false>(A, s, sp, W, wf, wb, tail_mode, i0);
          if (n < 0) n = spec_scan(s, sp, W, wf, wb, tail_mode, i0, size0);
          else if (A.dbg && lane == 0) atomicAdd(A.dbg + 32, 1ull);
          if (lane == 0) s.ro[RO_EXAMINED] = (uint32_t)n;
        }
        __syncthreads();
        if (prof) tp1 = clock64();
        const int n_ex = (int)s.ro[RO_EXAMINED];
        spec_versions(A, s, sp, (int)warp, kMT / 32);
        __syncthreads();
        if (prof) tp2 = clock64();
        spec_match<DR>(A, s, sp, n_ex, (int)warp, kMT / 32);
        __syncthreads();
        if (prof) tp3 = clock64();
        if (warp == 0) {
          const int cut = spec_verify(s, sp, n_ex, i0);
          if (lane == 0) {
            s.ro[RO_DONE] = (uint32_t)cut;
            if (cut < n_ex) ctl->mode = 1u;  // a misprediction: the next window takes the sequential loop
            if (A.dbg) {
              atomicAdd(A.dbg + 26, 1ull);
              atomicAdd(A.dbg + 27, (unsigned long long)(cut < n_ex ? 1 : 0));
            }
          }
        }
        __syncthreads();
        if (prof) tp4 = clock64();
        spec_commit<TEAM>(A, s, sp, n_ex, (int)s.ro[RO_DONE], i0, size0);
        if (prof) {
          atomicAdd(A.dbg + 28, (unsigned long long)(tp1 - tp0));
          atomicAdd(A.dbg + 29, (unsigned long long)(tp2 - tp1));
          atomicAdd(A.dbg + 30, (unsigned long long)(tp3 - tp2));
          atomicAdd(A.dbg + 31, (unsigned long long)(tp4 - tp3));
        }
      } else {
        if (tid == 0) ctl->mode = 0u;
        if (warp == 0) resolve_decide<TEAM, DR>(A, s, W, wf, wb, tail_mode, i0, size0);
        else mask_helpers(A, s, W, (int)warp - 1, kMT / 32 - 1);
      }
      __syncthreads();
      if (prof) tk5 = clock64();
      flush_window<TEAM>(A, seg, pos_nrm, seg_h, QH, ctl, s, W, wf, wb, tail_mode, i0, size0);
    }
    __threadfence();
    Team<TEAM>::sync();
    if (prof) {
      const long long tk6 = clock64();
      atomicAdd(A.dbg + 8, (unsigned long long)(tk1 - tk0));
      atomicAdd(A.dbg + 9, (unsigned long long)(tk2 - tk1));
      atomicAdd(A.dbg + 10, (unsigned long long)(tk3 - tk2));
      atomicAdd(A.dbg + 11, (unsigned long long)(tk4 - tk3));
      atomicAdd(A.dbg + 12, (unsigned long long)(tk5 - tk4));
      atomicAdd(A.dbg + 13, (unsigned long long)(tk6 - tk5));
      atomicAdd(A.dbg + 18, (unsigned long long)(ts1 - tk0));
      atomicAdd(A.dbg + 19, (unsigned long long)(ts2 - ts1));
      atomicAdd(A.dbg + 20, (unsigned long long)(ts3 - ts2));
      atomicAdd(A.dbg + 21, (unsigned long long)(tk1 - ts3));
    }
  }
}

template <int TEAM, int DR>
__global__ void __launch_bounds__(Shape<TEAM>::kMT, Shape<TEAM>::kCtasPerSm) k_merge_window(MergeArgs A) {
  extern __shared__ __align__(16) float smem_raw[];
  __shared__ uint32_t s_work;
  Smem s;
  carve(s, smem_raw, A.ld);
  // pool: window epochs of this team continue where the team slot's previous user (an earlier launch) stopped —
  // the per-team blocks persist, and a CTA of the team must never take a leftover block for the window it waits for
  if (threadIdx.x == 0) {
    uint32_t e0 = 0u;
    if constexpr (DR > 0 && TEAM != 2)
      if (A.pool) e0 = ld_acquire(&(A.pool_pub + Team<TEAM>::id())->meta) >> kPoolCountBits;
    s.ptask[6] = e0;
  }
  // staged rows are zero beyond the row's own width (row_width): nothing ever writes there
  for (int v = threadIdx.x; v < (2 * kW + kKD) * s.ts; v += blockDim.x) s.tile[v] = 0.f;
  __syncthreads();
  TeamCtl* ctl = A.ctl + Team<TEAM>::id();
  const uint32_t n0 = A.n_0 ? *A.n_0 : 0u, na = A.n_a ? *A.n_a : 0u, nb = A.n_b ? *A.n_b : 0u;
  auto item = [&](uint32_t w) {
    return (w < n0) ? (A.list_0 + 3 * (size_t)w) : (w - n0 < na) ? (A.list_a + 3 * (size_t)(w - n0)) : (A.list_b + 3 * (size_t)(w - n0 - na));
  };
  if (TEAM == 2) {  // the grid walks the lists together
    for (uint32_t w = 0; w < n0 + na + nb; ++w) {
      const uint32_t* it = item(w);
      merge_team<TEAM, DR>(A, it[0], it[1], it[2], ctl, s);
    }
    return;
  }
  for (;;) {
    uint32_t w;
    if (TEAM == 0) {
      if (threadIdx.x == 0) s_work = atomicAdd(A.cursor, 1u);
      __syncthreads();
      w = s_work;
      __syncthreads();
    } else {
      if (Team<TEAM>::rank() == 0 && threadIdx.x == 0) ctl->work = atomicAdd(A.cursor, 1u);
      __threadfence();
      Team<TEAM>::sync();
      w = __ldcg(&ctl->work);
      Team<TEAM>::sync();
    }
    if (w >= n0 + na + nb) break;
    const uint32_t* it = item(w);
    merge_team<TEAM, DR>(A, it[0], it[1], it[2], ctl, s);
  }
  if constexpr (TEAM == 1 && DR > 0) {
    if (A.pool) {
      // out of buckets: the first pool_helpers teams to get here stay on and serve chunks (each CTA on its own)
      // until every team of the launch is out of buckets; the others leave and free their SMs
      if (Team<TEAM>::rank() == 0 && threadIdx.x == 0) {
        atomicAdd(&A.pool->finished, 1u);
        ctl->work = atomicAdd(&A.pool->helpers, 1u);
      }
      __threadfence();
      Team<TEAM>::sync();
      const bool stay = __ldcg(&ctl->work) < A.pool_helpers;
      while (stay) {
        if (pool_help<Shape<TEAM>::kMT, DR>(A, s, 0xFFFFFFFFu)) continue;
        if (threadIdx.x == 0) s.ptask[5] = (ld_acquire(&A.pool->finished) >= A.pool_n || ld_acquire(&A.pool->big_active) == 0u) ? A.pool_n : 0u;
        __syncthreads();
        const bool fin = s.ptask[5] >= A.pool_n;
        __syncthreads();
        if (fin) break;
        __nanosleep(200);
      }
    }
  }
  if constexpr (TEAM == 0 && DR > 0) {
    if (A.pool) {  // out of buckets: serve screen tasks until every CTA of the launch is out of buckets
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd(&A.pool->finished, 1u);
      for (;;) {
        if (pool_help<Shape<TEAM>::kMT, DR>(A, s, blockIdx.x)) continue;
        if (threadIdx.x == 0) s.ptask[5] = ld_acquire(&A.pool->finished);
        __syncthreads();
        const bool fin = s.ptask[5] >= A.pool_n;
        __syncthreads();
        if (fin) break;
        __nanosleep(400);
      }
    }
  }
}

template <int TEAM>
const void* kernel_for(int ld) {
  if (ld <= 32) return (const void*)k_merge_window<TEAM, 32>;
  if (ld <= 64) return (const void*)k_merge_window<TEAM, 64>;
  return (const void*)k_merge_window<TEAM, 0>;
}

}  // namespace

// ================================================================================================
// Launch: stage 0 (CTA teams) over the classified lists, then the escalation stages.
// ================================================================================================
size_t merge_window_smem_bytes(int ld) { return smem_bytes_for(ld, Shape<1>::kMT); }

// The merge kernels are large; with CUDA's lazy module loading the first launch of each pays its load while the
// GPU waits between the launches of the first pass.  Touch them once when the context is created instead.
void merge_window_preload() {
  cudaFuncAttributes fa;
  const void* fns[] = {kernel_for<0>(32), kernel_for<0>(64), kernel_for<0>(128), kernel_for<1>(32), kernel_for<1>(64), kernel_for<1>(128),
                       kernel_for<2>(32), kernel_for<2>(64), kernel_for<2>(128)};
  for (const void* f : fns)
    if (cudaFuncGetAttributes(&fa, f) != cudaSuccess) (void)cudaGetLastError();
}

static int launch_stage(klsh_ctx* ctx, cudaStream_t stream, DevBuf& ctl_buf, int team, int csize, MergeArgs& A,
                        uint32_t host_items /* upper bound, 0 = unknown */) {
  const int ld = ctx->ld;
  const size_t smem = smem_bytes_for(ld, team == 0 ? Shape<0>::kMT : Shape<1>::kMT);
  if (smem > (size_t)ctx->max_smem_optin)
    return klsh_fail(ctx, KLSH_ERR_ARG, "dimension %d needs %zu bytes of shared memory per CTA (limit %d)", ctx->D, smem,
                     ctx->max_smem_optin);
  const void* fn = team == 0 ? kernel_for<0>(ld) : team == 1 ? kernel_for<1>(ld) : kernel_for<2>(ld);
  KCUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (team == 1 && csize > 8) {
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      (void)cudaGetLastError();
      csize = 8;
    }
  }
  const int kMT = team == 0 ? Shape<0>::kMT : Shape<1>::kMT;
  const int kCtasPerSm = Shape<0>::kCtasPerSm;
  int per_sm = 1;
  KCUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kMT, smem));
  if (per_sm < 1) per_sm = 1;
  uint32_t grid, nteams;
  if (team == 0) {
    nteams = (uint32_t)ctx->sm_count * std::min(per_sm, kCtasPerSm);
    if (host_items) nteams = std::min(nteams, host_items);
    grid = nteams;
  } else if (team == 1) {
    nteams = (uint32_t)std::max(1, ctx->sm_count * std::min(per_sm, ctx->cluster_ctas_per_sm) / csize);
    if (host_items) nteams = std::min(nteams, host_items);
    grid = nteams * csize;
  } else {
    nteams = 1;
    grid = (uint32_t)ctx->sm_count * std::min(per_sm, ctx->cluster_ctas_per_sm);
  }
  KTRY(dev_reserve(ctx, ctl_buf, sizeof(TeamCtl) * (size_t)std::max<uint32_t>(nteams, 1)));
  A.ctl = ctl_buf.as<TeamCtl>();

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kMT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  if (team == 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.numAttrs = 1;
  } else if (team == 2) {
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.numAttrs = 1;
  }
  std::chrono::high_resolution_clock::time_point t0;
  if (ctx->debug) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->stream2);
    cudaMemset(ctx->dbg.p, 0, sizeof(unsigned long long) * 40);
    t0 = std::chrono::high_resolution_clock::now();
  }
  if (A.pool) {
    A.pool_n = nteams;
    KCUDA(ctx, cudaMemsetAsync(&A.pool->finished, 0, 3 * sizeof(uint32_t), stream));  // finished, helpers, big_active
  }
  void* args[] = {&A};
  cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
  ctx->launches++;
  if (e != cudaSuccess)
    return klsh_fail(ctx, KLSH_ERR_CUDA, "merge kernel launch (team %d, grid %u, smem %zu) failed: %s", team, grid, smem,
                     cudaGetErrorString(e));
  if (ctx->debug) {  // KLSH_DEBUG=1: per-stage timing and window statistics on stderr
    unsigned long long h[40];
    cudaStreamSynchronize(stream);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
    cudaMemcpy(h, ctx->dbg.p, sizeof h, cudaMemcpyDeviceToHost);
    if (h[0])
      fprintf(stderr,
              "[klsh] merge team=%d csize=%d grid=%u: %.3f ms; windows %llu cands %llu merges %llu accepted %llu escalated %llu | trunc: "
              "undecidable %llu cache_full %llu back_exhausted %llu\n",
              team, csize, grid, ms, h[0], h[1], h[2], h[6], h[7], h[3], h[4], h[5]);
    if (h[0])
      fprintf(stderr, "[klsh]   leader kcycles/window: stage %.1f parallel %.1f sync1 %.1f prefetch %.1f decide %.1f flush+sync2 %.1f\n",
              h[8] / 1e3 / h[0], h[9] / 1e3 / h[0], h[10] / 1e3 / h[0], h[11] / 1e3 / h[0], h[12] / 1e3 / h[0], h[13] / 1e3 / h[0]);
    if (h[0])
      fprintf(stderr, "[klsh]   stage kcycles/window: index+meta %.1f rows %.1f norms %.1f fp16 %.1f\n", h[18] / 1e3 / h[0],
              h[19] / 1e3 / h[0], h[20] / 1e3 / h[0], h[21] / 1e3 / h[0]);
    if (h[0])
      fprintf(stderr, "[klsh]   decide kcycles/window: select %.1f dirty-compare %.1f old+accepted %.1f merge %.1f | per window: loop trips %.1f runs %.1f dirty tests %.1f\n",
              h[14] / 1e3 / h[0], h[15] / 1e3 / h[0], h[16] / 1e3 / h[0], h[17] / 1e3 / h[0], (double)h[22] / h[0], (double)h[23] / h[0], (double)h[24] / h[0]);
    if (h[26])
      fprintf(stderr, "[klsh]   speculative windows %llu (parallel scan %llu), cut short by a misprediction %llu; kcycles per speculative window: scan %.1f (order %.1f records %.1f) versions %.1f match %.1f verify %.1f\n",
              h[26], h[32], h[27], h[28] / 1e3 / h[26], h[36] / 1e3 / h[26], h[37] / 1e3 / h[26], h[29] / 1e3 / h[26], h[30] / 1e3 / h[26], h[31] / 1e3 / h[26]);
    if (h[35]) fprintf(stderr, "[klsh]   screen pool: %llu tasks posted, %llu served\n", h[35], h[34]);
    if (h[0]) fprintf(stderr, "[klsh]   dirty tests decided by the exact chain (inside the fast test's error band): %.3f per window\n", (double)h[25] / h[0]);
  }
  return KLSH_OK;
}

static MergeArgs base_args(klsh_ctx* ctx, PassScratch& s, uint32_t* rows_sorted, float threshold) {
  MergeArgs A;
  A.vals = ctx->cur.vals.as<float>();
  A.D = ctx->D;
  A.ld = ctx->ld;
  A.cnt = ctx->cur.cnt();
  A.head = ctx->cur.head();
  A.tail = ctx->cur.tail();
  A.next = ctx->cur.next.as<int32_t>();
  A.rows_sorted = rows_sorted;
  A.bstart = s.bstart.as<uint32_t>();
  A.pos_nrm = s.pos_nrm.as<float>();
  A.pos_h = s.pos_h.as<uint4>();
  A.dbg = ctx->debug ? ctx->dbg.as<unsigned long long>() : nullptr;
  A.mg = ctx->mg;
  A.threshold = threshold;
  A.no_spec = ctx->no_spec ? 1 : 0;
  A.scan_mode = ctx->scan_mode;
  A.work = ctx->eps_counter.p ? ctx->eps_counter.as<unsigned long long>() + 2 : nullptr;
  A.list_a = A.list_b = A.list_0 = nullptr;
  A.n_a = A.n_b = A.n_0 = nullptr;
  A.pool = nullptr;
  A.pool_board = nullptr;
  A.pool_pub = nullptr;
  A.pool_min = 0xFFFFFFFFu;
  A.pool_n = 0u;
  A.pool_helpers = 0u;
  A.cursor = nullptr;
  A.esc_list = nullptr;
  A.esc_count = nullptr;
  A.max_reps = 0xFFFFFFFFu;
  A.ctl = nullptr;
  return A;
}

// Work items {bucket, 0, 0} come from k_classify (list_big first, then list_large); their counts
// live in the pass counters on the device.
int launch_merge_window(klsh_ctx* ctx, PassScratch& s, uint32_t* rows_sorted, float threshold, uint32_t n_items_host,
                        uint32_t bucket_max_host) {
  if (n_items_host == 0) return KLSH_OK;
  PassCounters* dc = s.counters.as<PassCounters>();
  // escalation lists: every large bucket can escalate at most once per stage
  KTRY(dev_reserve(ctx, s.esc1, sizeof(uint32_t) * 3 * ((size_t)n_items_host + 1)));
  KTRY(dev_reserve(ctx, s.esc2, sizeof(uint32_t) * 3 * ((size_t)n_items_host + 1)));
  KTRY(dev_reserve(ctx, s.esc3, sizeof(uint32_t) * 3 * ((size_t)n_items_host + 1)));
  if (ctx->debug) KTRY(dev_reserve(ctx, ctx->dbg, sizeof(unsigned long long) * 40));

  MergeArgs A = base_args(ctx, s, rows_sorted, threshold);

  // stage 0: one CTA per bucket, biggest buckets first
  A.list_a = s.list_big.as<uint32_t>();
  A.n_a = &dc->n_big;
  A.list_b = s.list_large.as<uint32_t>();
  A.n_b = &dc->n_large;
  A.cursor = &dc->large_cursor;
  A.esc_list = s.esc1.as<uint32_t>();
  A.esc_count = &dc->n_esc1;
  A.max_reps = ctx->cta_max;
  KTRY(launch_stage(ctx, ctx->stream, ctx->team_ctl, 0, 1, A, n_items_host));
  if (bucket_max_host <= ctx->cta_max) return KLSH_OK;  // nothing can have escalated

  // stage 1: one cluster per escalated bucket
  A.list_a = s.esc1.as<uint32_t>();
  A.n_a = &dc->n_esc1;
  A.list_b = nullptr;
  A.n_b = nullptr;
  A.cursor = &dc->cluster_cursor;
  A.esc_list = s.esc2.as<uint32_t>();
  A.esc_count = &dc->n_esc2;
  A.max_reps = ctx->cluster_max;
  KTRY(launch_stage(ctx, ctx->stream, ctx->team_ctl, 1, ctx->cluster_size, A, 0));
  if (bucket_max_host <= ctx->cluster_max) return KLSH_OK;

  // stage 2: one large cluster per bucket
  A.list_a = s.esc2.as<uint32_t>();
  A.n_a = &dc->n_esc2;
  A.cursor = &dc->cluster2_cursor;
  A.esc_list = s.esc3.as<uint32_t>();
  A.esc_count = &dc->n_esc3;
  A.max_reps = ctx->cluster2_max;
  KTRY(launch_stage(ctx, ctx->stream, ctx->team_ctl, 1, ctx->cluster2_size, A, 0));
  if (bucket_max_host <= ctx->cluster2_max) return KLSH_OK;

  // stage 3: the whole grid per bucket
  A.list_a = s.esc3.as<uint32_t>();
  A.n_a = &dc->n_esc3;
  A.cursor = nullptr;
  A.esc_list = nullptr;
  A.esc_count = nullptr;
  A.max_reps = 0xFFFFFFFFu;
  KTRY(launch_stage(ctx, ctx->stream, ctx->team_ctl, 2, 1, A, 0));
  return KLSH_OK;
}

// ---- screen pool launch ---------------------------------------------------------------------------------------
bool launch_merge_uses_pool(const klsh_ctx* ctx) { return ctx->pool && ctx->ld <= 64 && !launch_merge_uses_fallback(ctx); }

// One launch of single-CTA teams over every bucket of more than KLSH_SMALL_MAX rows, the pass's largest first
// (list_direct, then list_big, then list_large); screens of buckets past pool_min representatives go through the pool.
int launch_merge_pool(klsh_ctx* ctx, PassScratch& s, uint32_t* rows_sorted, float threshold) {
  PassCounters* dc = s.counters.as<PassCounters>();
  if (ctx->debug) KTRY(dev_reserve(ctx, ctx->dbg, sizeof(unsigned long long) * 40));
  const uint32_t grid_max = (uint32_t)ctx->sm_count * (uint32_t)Shape<0>::kCtasPerSm;
  if (!ctx->pool_ctl.p) {  // control word and board start as zeros and every launch leaves them that way
    KTRY(dev_reserve(ctx, ctx->pool_ctl, sizeof(PoolCtl) + sizeof(uint32_t) * (size_t)grid_max));
    KCUDA(ctx, cudaMemsetAsync(ctx->pool_ctl.p, 0, ctx->pool_ctl.bytes, ctx->stream));
    KTRY(dev_reserve(ctx, ctx->pool_pub, sizeof(PoolPub) * (size_t)grid_max));
    KCUDA(ctx, cudaMemsetAsync(ctx->pool_pub.p, 0, ctx->pool_pub.bytes, ctx->stream));
  }
  MergeArgs A = base_args(ctx, s, rows_sorted, threshold);
  A.list_0 = s.list_direct.as<uint32_t>();
  A.n_0 = &dc->n_direct;
  A.list_a = s.list_big.as<uint32_t>();
  A.n_a = &dc->n_big;
  A.list_b = s.list_large.as<uint32_t>();
  A.n_b = &dc->n_large;
  A.cursor = &dc->large_cursor;
  A.pool = ctx->pool_ctl.as<PoolCtl>();
  A.pool_board = reinterpret_cast<uint32_t*>(ctx->pool_ctl.as<PoolCtl>() + 1);
  A.pool_pub = ctx->pool_pub.as<PoolPub>();
  A.pool_min = ctx->pool_min;
  return launch_stage(ctx, ctx->stream, ctx->team_ctl, 0, 1, A, 0);
}

static bool cluster_pool_on(const klsh_ctx* ctx) { return ctx->cpool && ctx->ld <= 64; }

// Called once per pass BEFORE the fork: allocates the pool's control block, board and per-team blocks.
int launch_pool_reset(klsh_ctx* ctx) {
  if (!cluster_pool_on(ctx)) return KLSH_OK;
  const uint32_t teams_max = (uint32_t)ctx->sm_count * 2u;
  if (!ctx->pool_ctl_b.p) {  // control block + board and the per-team blocks start as zeros and every pass leaves them that way
    KTRY(dev_reserve(ctx, ctx->pool_ctl_b, sizeof(PoolCtl) + sizeof(uint32_t) * (size_t)teams_max));
    KCUDA(ctx, cudaMemsetAsync(ctx->pool_ctl_b.p, 0, ctx->pool_ctl_b.bytes, ctx->stream));
    KTRY(dev_reserve(ctx, ctx->pool_pub_b, sizeof(PoolPub) * (size_t)teams_max));
    KCUDA(ctx, cudaMemsetAsync(ctx->pool_pub_b.p, 0, ctx->pool_pub_b.bytes, ctx->stream));
  }
  return KLSH_OK;
}

// The direct pipeline: buckets of at least direct_min rows start on cluster teams (no single-CTA stage)
// and escalate to large clusters and the grid like the others.  Everything is enqueued on the
// context's second stream with its own work lists, cursors and team control blocks, so it runs beside
// the small buckets and the single-CTA stage: the step ends when the longer of the two ends, not their sum.
int launch_merge_direct(klsh_ctx* ctx, PassScratch& s, uint32_t* rows_sorted, float threshold, uint32_t n_direct_host,
                        uint32_t bucket_max_host) {
  if (n_direct_host == 0) return KLSH_OK;
  PassCounters* dc = s.counters.as<PassCounters>();
  KTRY(dev_reserve(ctx, s.escb2, sizeof(uint32_t) * 3 * ((size_t)n_direct_host + 1)));
  KTRY(dev_reserve(ctx, s.escb3, sizeof(uint32_t) * 3 * ((size_t)n_direct_host + 1)));
  if (ctx->debug) KTRY(dev_reserve(ctx, ctx->dbg, sizeof(unsigned long long) * 40));
  MergeArgs A = base_args(ctx, s, rows_sorted, threshold);
  cudaStream_t st = ctx->stream2;

  A.list_a = s.list_direct.as<uint32_t>();
  A.n_a = &dc->n_direct;
  A.cursor = &dc->b_cursor1;
  A.esc_list = s.escb2.as<uint32_t>();
  A.esc_count = &dc->nb_esc2;
  A.max_reps = ctx->cluster_max;
  // few direct buckets: give each a large (16-CTA) cluster from the start — their screen is spread over twice
  // the SMs and nobody queues; many: the portable 8-CTA clusters, so that more of them run side by side
  const int per_sm = std::max(1, ctx->cluster_ctas_per_sm);
  const int csize1 = ((uint64_t)n_direct_host * (uint64_t)ctx->cluster2_size <= (uint64_t)ctx->sm_count * per_sm) ? ctx->cluster2_size : ctx->cluster_size;
  const bool cpool = cluster_pool_on(ctx);
  if (cpool) {
    // screen pool of the cluster teams (opt-in): their windows' screens are split into chunks that the cluster's own
    // CTAs claim, and that teams of this launch told to stay on after running out of buckets may claim too
    A.pool = ctx->pool_ctl_b.as<PoolCtl>();
    A.pool_board = reinterpret_cast<uint32_t*>(ctx->pool_ctl_b.as<PoolCtl>() + 1);
    A.pool_pub = ctx->pool_pub_b.as<PoolPub>();
    A.pool_min = ctx->cpool_min;
    A.pool_helpers = ctx->cpool_helper_ctas / (uint32_t)csize1;
  }
  KTRY(launch_stage(ctx, st, ctx->team_ctl_b, 1, csize1, A, n_direct_host));
  if (bucket_max_host <= ctx->cluster_max) return KLSH_OK;

  A.list_a = s.escb2.as<uint32_t>();
  A.n_a = &dc->nb_esc2;
  A.cursor = &dc->b_cursor2;
  A.esc_list = s.escb3.as<uint32_t>();
  A.esc_count = &dc->nb_esc3;
  A.max_reps = ctx->cluster2_max;
  if (cpool) {
    A.pool_helpers = ctx->cpool_helper_ctas / (uint32_t)ctx->cluster2_size;
    // with the pool the large clusters carry any bucket: no grid stage
    A.esc_list = nullptr;
    A.esc_count = nullptr;
    A.max_reps = 0xFFFFFFFFu;
  }
  KTRY(launch_stage(ctx, st, ctx->team_ctl_b, 1, ctx->cluster2_size, A, 0));
  if (cpool || bucket_max_host <= ctx->cluster2_max) return KLSH_OK;

  A.pool = nullptr;
  A.list_a = s.escb3.as<uint32_t>();
  A.n_a = &dc->nb_esc3;
  A.cursor = nullptr;
  A.esc_list = nullptr;
  A.esc_count = nullptr;
  A.max_reps = 0xFFFFFFFFu;
  KTRY(launch_stage(ctx, st, ctx->team_ctl_b, 2, 1, A, 0));
  return KLSH_OK;
}
