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
  uint32_t pad[4];
};

struct MergeArgs {
  float* vals;
  int D, ld;
  int32_t *cnt, *head, *tail, *next;
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
  MgLog mg;
  unsigned long long* dbg;  // [16] (8..13: leader cycles in stage/parallel/sync1/prefetch/resolve/sync2) windows, candidates, merges, undecidable, cache_full, back_exhausted, accepted, escalated
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
  int ts;
};

// width (halfs) of the fp16 window copy: the k extent the kernel variant for this ld multiplies over
__host__ __device__ inline int tc_width(int ld) { return ld <= 32 ? 32 : (ld <= 64 ? 64 : ((ld + 31) & ~31)); }

__host__ __device__ inline size_t ring_bytes_for(int ld, int threads) {
  return ld <= 64 ? (size_t)(threads / 32) * kRing * 32 * (tc_width(ld) / 32) * 16 : 0;
}
__host__ __device__ inline size_t smem_bytes_for(int ld, int threads) {
  const int hs = tc_width(ld) + 8;
  return sizeof(float) * ((size_t)(2 * kW + kKD) * (ld + 4) + kW + kKD) + sizeof(uint32_t) * (kW * 14 + kKD * 5) + 64 +
         sizeof(__half) * (size_t)kW * hs + 16 + sizeof(uint32_t) * (kSurvCap + 1) + ring_bytes_for(ld, threads);
}

__device__ __forceinline__ void carve(Smem& s, float* base, int ld) {
  s.ts = ld + 4;
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

// Tensor-core screened comparison of the window with representatives [j_begin, j_end) (SELF == false)
// or with the window's own rows (SELF == true: candidate x candidate bits).  One warp handles 8
// representatives per step; KS16 = number of 16-wide k steps (rows are zero-padded to 16*KS16).
// Representatives come from the fragment-order fp16 copy seg_h (KS16/2 16-byte chunks per lane and
// step, coalesced), fetched several steps ahead so that the tensor pipe, not L2 latency, paces the loop.
template <int KS16, bool SELF>
__device__ __forceinline__ void tc_compare(const MergeArgs& A, const uint32_t* seg, const float* pos_nrm, const uint4* seg_h,
                                           Smem& s, int W, uint32_t j_begin, uint32_t j_end, uint32_t warp_rank,
                                           uint32_t n_warps, int nq) {
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
      const __half* r0 = s.htile + (size_t)(mt * 16 + g) * s.hs + ks * 16 + tg * 2;
      const __half* r1 = r0 + 8 * s.hs;
      af[mt][ks][0] = *reinterpret_cast<const uint32_t*>(r0);
      af[mt][ks][1] = *reinterpret_cast<const uint32_t*>(r1);
      af[mt][ks][2] = *reinterpret_cast<const uint32_t*>(r0 + 8);
      af[mt][ks][3] = *reinterpret_cast<const uint32_t*>(r1 + 8);
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
          } else if (!(s.s_f[t] < jj)) {  // skip when an earlier match is already recorded
            // The exact test needs the representative's fp32 row (two dependent global round trips):
            // park the pair and test all parked pairs of the CTA together after the streaming loop.
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

// ---- the resolver: warp 0 of the team's leader CTA ---------------------------------------------------
// Lane e owns dirty-cache entry e: its position, member count and a 64-bit mask of the window
// candidates that match the entry's CURRENT value (recomputed, for all unexamined candidates at once,
// whenever the entry changes; the new norm comes out of the same pass).  Examining a candidate is
// then a handful of bit tests; the floating-point work happens once per merge.  Member-chain
// splices and the swap-remove writes are independent of the decisions, so they are logged and
// applied in parallel when the window ends.
template <int TEAM>
__device__ void resolve_window(const MergeArgs& A, uint32_t* seg, float* pos_nrm, uint4* seg_h, int qh, TeamCtl* ctl, Smem& s,
                               int W, int wf, int wb, bool tail_mode, uint32_t i0, uint32_t size0) {
  const int D = A.D, ld = A.ld, nq = ld >> 2;
  const uint32_t lane = lane_id();
  int nd = 0, a = 0, fi = 0, bi = 0, merges = 0;
  uint32_t i = i0, size = size0;
  bool from_back = false, back_exhausted = false;
  int dbg_undec = 0, dbg_full = 0;
  uint32_t accd_lo = 0, accd_hi = 0;  // accepted-in-window candidates that have since been modified
  uint32_t accm_lo = 0, accm_hi = 0;  // accepted-in-window candidates (tile indices)
  uint32_t my_dpos = kInf, my_dm_lo = 0, my_dm_hi = 0;  // this lane's dirty entry
  int my_dcnt = 0, my_last = -1;                        // its member count; last candidate merged into it
  // The entry modified by the latest merge keeps an INVALID mask: runs of candidates merging into the
  // same representative (the common shape of a merge-heavy bucket) then cost one exact comparison
  // each instead of a whole-window mask rebuild.  The mask is rebuilt when another entry is modified
  // or a candidate is accepted.
  uint32_t inval = 0;  // entries whose match mask is stale (modified since it was built)
  int pend = -1;       // the entry whose norm is stale too (modified by the latest merge)
  // which unexamined candidates match entry e's current value; its norm falls out of the same pass
  long long pv = 0, pa = 0, pm = 0;
  int nval = 0;
  const bool rprof = A.dbg != nullptr;
  auto validate = [&](int e) {
    const long long tv0 = rprof ? clock64() : 0;
    ++nval;
    const float4* r4 = reinterpret_cast<const float4*>(s.dvals + (size_t)e * s.ts);
    const int nfront = tail_mode ? max(0, wf - fi - bi) : (wf - fi);
    const int nback = tail_mode ? 0 : (wb - bi);
    const int nun = nfront + nback;
    uint32_t w0 = 0, w1 = 0;
    float rn = 0.f;
    for (int base = 0; base < max(nun, 1); base += 32) {
      const int k = base + (int)lane;
      const bool act = k < nun;
      const int tc = act ? (k < nfront ? fi + k : wf + bi + (k - nfront)) : 0;
      const float4* c4 = reinterpret_cast<const float4*>(s.tile + (size_t)tc * s.ts);
      float dot = 0.f, nn = 0.f;
      if (base == 0) {  // the entry's norm comes out of the first pass
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
#pragma unroll 4
        for (int q = 0; q < nq; ++q) {
          const float4 x = c4[q], y = r4[q];
          dot = __fadd_rn(dot, __fmul_rn(x.x, y.x));
          dot = __fadd_rn(dot, __fmul_rn(x.y, y.y));
          dot = __fadd_rn(dot, __fmul_rn(x.z, y.z));
          dot = __fadd_rn(dot, __fmul_rn(x.w, y.w));
        }
      }
      const bool mt = act && cos_match(dot, s.cnorm[tc], rn, A.threshold);
      const uint32_t blo = (mt && tc < 32) ? (1u << tc) : 0u, bhi = (mt && tc >= 32) ? (1u << (tc - 32)) : 0u;
      w0 |= __reduce_or_sync(0xffffffffu, blo);
      w1 |= __reduce_or_sync(0xffffffffu, bhi);
    }
    if ((int)lane == e) {
      my_dm_lo = w0;
      my_dm_hi = w1;
    }
    if (lane == 0) s.dnorm[e] = rn;
    if (rprof) pv += clock64() - tv0;
  };
  // s.pair is dead once a candidate has been examined; reuse the per-candidate slots s.acc/... no:
  // the merge log lives in s.pridx (prefetch row index, dead after the entry was allocated) as
  // "previous candidate merged into the same entry" (kInf: none) and s.pcnt as the entry index.
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
    const uint32_t fpos = s.s_f[t];
    const uint32_t plo = s.pair[2 * t], phi = s.pair[2 * t + 1];
    const uint32_t tbit_lo = (t < 32) ? (1u << t) : 0u, tbit_hi = (t < 32) ? 0u : (1u << (t - 32));
    if (inval != 0u) {
      // Stale masks still predict well.  If nothing suggests that t merges, it will be accepted and the
      // masks would be rebuilt right after its exact comparisons: rebuild them first (t is still in
      // the unexamined set, so its bits come out of the same pass) and skip those comparisons.
      const bool stale_hit = ((my_dm_lo & tbit_lo) | (my_dm_hi & tbit_hi)) != 0u;
      const bool likely = fpos != kInf || ((plo & accm_lo) | (phi & accm_hi)) != 0u || __any_sync(0xffffffffu, stale_hit);
      if (!likely) {
        while (inval != 0u) {
          const int e2 = __ffs(inval) - 1;
          inval &= inval - 1;
          validate(e2);
        }
        pend = -1;
      }
    }
    if (from_back) ++bi; else ++fi;
    uint32_t best = kInf;
    if (nd > 0) {
      // (a) representatives modified in this window: match bits against their current values; the
      // entry with an invalid mask is compared exactly (every lane computes the same comparison)
      bool hit = ((my_dm_lo & tbit_lo) | (my_dm_hi & tbit_hi)) != 0u;
      if (inval != 0u) {
        // entries with a stale mask: lane e compares the candidate with entry e's current value.  Only
        // the entry modified last (pend) needs a new norm as well: a lane that owns no stale entry
        // accumulates entry x entry in the same instruction stream, so the warp walks ONE chain of
        // mul+add per lane instead of two.
        const bool mine = (inval >> lane) & 1u;
        const uint32_t spare = ~inval;
        const int nl = (pend >= 0 && spare != 0u) ? (__ffs(spare) - 1) : -1;
        const bool normer = (int)lane == nl;
        const bool inline_nn = pend >= 0 && nl < 0;  // all 32 entries stale: the owner walks both chains
        const int er = mine ? (int)lane : (normer ? pend : 0);
        const float4* r4 = reinterpret_cast<const float4*>(s.dvals + (size_t)er * s.ts);
        const float4* c4 = normer ? r4 : reinterpret_cast<const float4*>(s.tile + (size_t)t * s.ts);
        float dot = 0.f, nn = 0.f;
#pragma unroll 4
        for (int q = 0; q < nq; ++q) {
          const float4 x = c4[q], y = r4[q];
          dot = __fadd_rn(dot, __fmul_rn(x.x, y.x));
          dot = __fadd_rn(dot, __fmul_rn(x.y, y.y));
          dot = __fadd_rn(dot, __fmul_rn(x.z, y.z));
          dot = __fadd_rn(dot, __fmul_rn(x.w, y.w));
          if (inline_nn) {
            nn = __fadd_rn(nn, __fmul_rn(y.x, y.x));
            nn = __fadd_rn(nn, __fmul_rn(y.y, y.y));
            nn = __fadd_rn(nn, __fmul_rn(y.z, y.z));
            nn = __fadd_rn(nn, __fmul_rn(y.w, y.w));
          }
        }
        float rn_p = 0.f;
        if (pend >= 0) rn_p = __shfl_sync(0xffffffffu, __fsqrt_rn(inline_nn ? nn : dot), inline_nn ? pend : nl);
        if (mine) {
          const float rn = ((int)lane == pend) ? rn_p : s.dnorm[lane];
          hit = cos_match(dot, s.cnorm[t], rn, A.threshold);
          if ((int)lane == pend) s.dnorm[lane] = rn_p;
        }
        pend = -1;  // the latest entry's norm is now published
      }
      best = __reduce_min_sync(0xffffffffu, hit ? my_dpos : kInf);
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
      // an accept usually means more accepts follow: rebuild the stale masks once so that the
      // following candidates are decided by bit tests alone
      while (inval != 0u) {
        const int e2 = __ffs(inval) - 1;
        inval &= inval - 1;
        validate(e2);
      }
      pend = -1;
      if (lane == 0) s.acc[a] = (uint32_t)t;
      accm_lo |= tbit_lo;
      accm_hi |= tbit_hi;
      ++a;
      ++i;
      from_back = false;
      continue;
    }
    // merge the candidate into the representative at position `best`
    const long long tm0 = rprof ? clock64() : 0;
    const uint32_t p = best;
    int e;
    {
      const uint32_t m = __ballot_sync(0xffffffffu, my_dpos == p);
      if (m) {
        e = __ffs(m) - 1;
      } else {
        e = nd++;
        float* dst = s.dvals + (size_t)e * s.ts;
        int cnt_e;
        if (p < i0) {
          // an old representative enters the cache only as this candidate's precomputed first match
          // (any other old position in `best` is already dirty), so its row was prefetched
          const float* src = s.pre + (size_t)t * s.ts;
          for (int d = lane; d < ld; d += 32) dst[d] = src[d];
          cnt_e = s.pcnt[t];
          if (lane == 0) {
            s.dridx[e] = s.pridx[t];
            s.dhead[e] = s.phead[t];
            s.dtail[e] = s.ptail[t];
          }
        } else {
          __syncwarp();
          const uint32_t u = s.acc[p - i0];
          const float* src = s.tile + (size_t)u * s.ts;
          for (int d = lane; d < ld; d += 32) dst[d] = src[d];
          cnt_e = s.ccnt[u];
          if (lane == 0) {
            s.dridx[e] = s.ridx[u];
            s.dhead[e] = s.chead[u];
            s.dtail[e] = s.ctail[u];
          }
          if (u < 32) accd_lo |= 1u << u; else accd_hi |= 1u << (u - 32);
        }
        if (lane == 0) s.dpos[e] = p;
        if ((int)lane == e) {
          my_dpos = p;
          my_dcnt = cnt_e;
          my_last = -1;
        }
        __syncwarp();
      }
    }
    const int c1 = s.ccnt[t];
    const int c2 = __shfl_sync(0xffffffffu, my_dcnt, e);
    {
      const float* c = s.tile + (size_t)t * s.ts;
      float* r = s.dvals + (size_t)e * s.ts;
      for (int d = lane; d < D; d += 32) r[d] = consensus1(c[d], c1, r[d], c2);
    }
    // merge log: which candidate was merged into this entry just before t (applied at window end)
    {
      const int prev = __shfl_sync(0xffffffffu, my_last, e);
      if (lane == 0) {
        s.mprev[t] = prev;
        s.ment[t] = e;
      }
      if ((int)lane == e) {
        my_dcnt = c1 + c2;
        if (s.ctail[t] >= 0) my_last = t;  // only candidates that carry ids take part in the chain
      }
    }
    --size;
    ++merges;
    from_back = true;
    __syncwarp();
    inval |= 1u << e;
    pend = e;
    if (rprof) pm += clock64() - tm0;
    if (nd == kKD) { dbg_full = 1; break; }  // dirty cache full: flush and start a new window
  }
  __syncwarp();
  if (pend >= 0) validate(pend);  // publishes the latest entry's norm
  __syncwarp();
  const long long ta0 = rprof ? clock64() : 0;
  // ---- apply the window's effects to global memory, in parallel ----
  // accepted candidates: positions i0.. in acceptance order (row index + norm)
  for (int k = lane; k < a; k += 32) {
    const uint32_t u = s.acc[k];
    seg[i0 + k] = s.ridx[u];
    pos_nrm[i0 + k] = s.cnorm[u];
  }
  if (qh) {
    // fp16 copies: accepted candidates take theirs from the window's fp16 tile, modified
    // representatives are re-scaled from their current value and norm
    for (int idx = lane; idx < a * qh; idx += 32) {
      const int k = idx / qh, c = idx - k * qh;
      seg_h[(size_t)(i0 + k) * qh + c] = h16_chunk_from_htile(s.htile + (size_t)s.acc[k] * s.hs, c);
    }
    __syncwarp();  // an accepted candidate modified later in the window is rewritten below
    for (int idx = lane; idx < nd * qh; idx += 32) {
      const int e = idx / qh, c = idx - e * qh;
      seg_h[(size_t)s.dpos[e] * qh + c] = h16_chunk_from_row(s.dvals + (size_t)e * s.ts, ld, s.dnorm[e], c);
    }
  }
  // a moved tail element waiting at position i (the window ended right after a merge)
  if (from_back && i < size && lane == 0) {
    uint32_t moved;
    if (tail_mode) moved = s.ridx[wf - 1 - bi];
    else if (bi < wb) moved = s.ridx[wf + bi];
    else moved = __ldcg(seg + size);  // the element that was at the old tail position
    seg[i] = moved;
  }
  __syncwarp();
  for (uint32_t k = size + lane; k < size0; k += 32) seg[k] = KLSH_SENTINEL;
  // member chains: ids(current) ++ ids(candidate) for every logged merge
  for (int round = 0; round < 2; ++round) {
    const int t = round * 32 + (int)lane;
    if (t < W) {
      const bool examined_merge = s.ment[t] >= 0;
      if (examined_merge) {
        const int e = s.ment[t], prev = s.mprev[t];
        const int t1 = s.ctail[t];
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
    }
  }
  __syncwarp();
  // modified representatives: values, count, head, tail, norm
  {
    const int e = (int)lane;
    if (e < nd) {
      const uint32_t rr = s.dridx[e];
      if (A.mg.counts) A.mg.mod_rows[atomicAdd(A.mg.counts, 1u)] = rr;
      A.cnt[rr] = my_dcnt;
      // each merge prepends the candidate's members: the chain now starts with the LAST merged
      // candidate that carried ids
      if (my_last >= 0) A.head[rr] = s.chead[my_last];
      if (s.dtail[e] < 0 && my_last >= 0) {  // the representative had no ids: its tail is the EARLIEST such candidate's
        int first_t = my_last;
        for (int tt = s.mprev[my_last]; tt >= 0; tt = s.mprev[tt]) first_t = tt;
        A.tail[rr] = s.ctail[first_t];
      }
      pos_nrm[my_dpos] = s.dnorm[e];
    }
  }
  for (int e = 0; e < nd; ++e) {
    float* dst = A.vals + (uint64_t)s.dridx[e] * ld;
    const float* src = s.dvals + (size_t)e * s.ts;
    for (int d = lane; d < D; d += 32) dst[d] = src[d];
  }
  if (rprof) pa = clock64() - ta0;
  if (lane == 0) {
    ctl->i = i;
    ctl->size = size;
    // back candidates for the next window: grow fast when they ran out, shrink slowly otherwise (a
    // window that runs out of them throws away the screen of its remaining front candidates)
    ctl->wb = back_exhausted ? (uint32_t)min(kWbMax, max(wb * 2, 4))
                             : (uint32_t)min(kWbMax, max(max(2, merges + merges / 2 + 2), wb - (wb + 3) / 4));
    if (A.dbg) {
      atomicAdd(A.dbg + 0, 1ull);
      atomicAdd(A.dbg + 1, (unsigned long long)(fi + bi));
      atomicAdd(A.dbg + 2, (unsigned long long)merges);
      atomicAdd(A.dbg + 3, (unsigned long long)dbg_undec);
      atomicAdd(A.dbg + 4, (unsigned long long)dbg_full);
      atomicAdd(A.dbg + 5, (unsigned long long)(back_exhausted ? 1 : 0));
      atomicAdd(A.dbg + 6, (unsigned long long)a);
      atomicAdd(A.dbg + 14, (unsigned long long)pv);
      atomicAdd(A.dbg + 15, (unsigned long long)pm);
      atomicAdd(A.dbg + 16, (unsigned long long)pa);
      atomicAdd(A.dbg + 17, (unsigned long long)nval);
    }
  }
  __syncwarp();
}

// Window staging.  Rows go global -> shared with cp.async (no registers, no wait until the whole
// batch is in flight), so a stage costs two dependent round trips (row index -> row / metadata)
// however many rows a thread moves.

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
      }
    } else if (lane == 0) {
      ctl->i = start_i;
      ctl->size = start_size;
      ctl->wb = 4;
    }
    if (TEAM != 0)
      for (int t = lane; t < kW; t += 32) ctl->f[t] = kInf;
  }
  __threadfence();
  Team<TEAM>::sync();
  for (;;) {
    const uint32_t i0 = __ldcg(&ctl->i), size0 = __ldcg(&ctl->size);
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
          s.ccnt[tid] = A.cnt[mr];
          s.chead[tid] = A.head[mr];
          s.ctail[tid] = A.tail[mr];
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
    if constexpr (DR > 0) {
      constexpr int KS16 = DR / 16;
      tc_compare<KS16, false>(A, seg, pos_nrm, seg_h, s, W, 0u, i0, Team<TEAM>::rank() * (kMT / 32) + warp,
                              Team<TEAM>::ncta() * (kMT / 32), nq);
      if (leader) tc_compare<KS16, true>(A, seg, pos_nrm, seg_h, s, W, 0u, (uint32_t)W, warp, kMT / 32, nq);
    } else {
      const int ks16 = (s.hs - 8) >> 4;
      tc_compare_wide<false>(A, seg, pos_nrm, seg_h, s, W, 0u, i0, Team<TEAM>::rank() * (kMT / 32) + warp,
                             Team<TEAM>::ncta() * (kMT / 32), nq, ks16);
      if (leader) tc_compare_wide<true>(A, seg, pos_nrm, seg_h, s, W, 0u, (uint32_t)W, warp, kMT / 32, nq, ks16);
    }
    __syncthreads();
    {
      // exact tests of the pairs the screen parked: all of the CTA's threads at once, so the
      // dependent row fetches of different pairs overlap
      const uint32_t ns = min(s.surv[kSurvCap], (uint32_t)kSurvCap);
      for (uint32_t k0 = 0; k0 < ns; k0 += 2 * kMT) {
        uint32_t ent[2], rr2[2];
        bool on[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t k = k0 + u * kMT + tid;
          on[u] = k < ns;
          ent[u] = on[u] ? s.surv[k] : 0u;
          on[u] = on[u] && !(s.s_f[ent[u] & 63u] < (ent[u] >> 6));
          rr2[u] = on[u] ? __ldcg(seg + (ent[u] >> 6)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (on[u]) {
            const uint32_t jj = ent[u] >> 6;
            const int t = (int)(ent[u] & 63u);
            if (exact_pair<false>(s, t, A.vals + (uint64_t)rr2[u] * ld, __ldcg(pos_nrm + jj), nq, A.threshold))
              atomicMin(&s.s_f[t], jj);
          }
      }
      if (ns) __syncthreads();
    }
    if (prof) tk2 = clock64();
    if (TEAM != 0) {
      if (tid < W && s.s_f[tid] != kInf) atomicMin(&ctl->f[tid], s.s_f[tid]);
      __threadfence();
      Team<TEAM>::sync();
    }
    if (prof) tk3 = clock64();
    if (leader) {
      if (TEAM != 0) {
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
          s.pcnt[tid] = A.cnt[mr];
          s.phead[tid] = A.head[mr];
          s.ptail[tid] = A.tail[mr];
        }
        cp_async_wait_all();
      }
      __syncthreads();
      if (prof) tk4 = clock64();
      if (warp == 0) resolve_window<TEAM>(A, seg, pos_nrm, seg_h, QH, ctl, s, W, wf, wb, tail_mode, i0, size0);
      if (prof) tk5 = clock64();
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
  TeamCtl* ctl = A.ctl + Team<TEAM>::id();
  const uint32_t na = A.n_a ? *A.n_a : 0u, nb = A.n_b ? *A.n_b : 0u;
  if (TEAM == 2) {  // the grid walks the lists together
    for (uint32_t w = 0; w < na + nb; ++w) {
      const uint32_t* it = (w < na) ? (A.list_a + 3 * (size_t)w) : (A.list_b + 3 * (size_t)(w - na));
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
    if (w >= na + nb) break;
    const uint32_t* it = (w < na) ? (A.list_a + 3 * (size_t)w) : (A.list_b + 3 * (size_t)(w - na));
    merge_team<TEAM, DR>(A, it[0], it[1], it[2], ctl, s);
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
static int launch_stage(klsh_ctx* ctx, int team, int csize, MergeArgs& A, uint32_t host_items /* upper bound, 0 = unknown */) {
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
  KTRY(dev_reserve(ctx, ctx->team_ctl, sizeof(TeamCtl) * (size_t)std::max<uint32_t>(nteams, 1)));
  A.ctl = ctx->team_ctl.as<TeamCtl>();

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kMT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
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
    cudaMemset(ctx->dbg.p, 0, sizeof(unsigned long long) * 24);
    t0 = std::chrono::high_resolution_clock::now();
  }
  void* args[] = {&A};
  cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
  ctx->launches++;
  if (e != cudaSuccess)
    return klsh_fail(ctx, KLSH_ERR_CUDA, "merge kernel launch (team %d, grid %u, smem %zu) failed: %s", team, grid, smem,
                     cudaGetErrorString(e));
  if (ctx->debug) {  // KLSH_DEBUG=1: per-stage timing and window statistics on stderr
    unsigned long long h[24];
    cudaStreamSynchronize(ctx->stream);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
    cudaMemcpy(h, ctx->dbg.p, sizeof h, cudaMemcpyDeviceToHost);
    if (h[0])
      fprintf(stderr,
              "[klsh] merge team=%d csize=%d grid=%u: %.3f ms; windows %llu cands %llu merges %llu accepted %llu escalated %llu | trunc: "
              "undecidable %llu cache_full %llu back_exhausted %llu\n",
              team, csize, grid, ms, h[0], h[1], h[2], h[6], h[7], h[3], h[4], h[5]);
    if (h[0])
      fprintf(stderr, "[klsh]   leader kcycles/window: stage %.1f parallel %.1f sync1 %.1f prefetch %.1f resolve %.1f sync2 %.1f\n",
              h[8] / 1e3 / h[0], h[9] / 1e3 / h[0], h[10] / 1e3 / h[0], h[11] / 1e3 / h[0], h[12] / 1e3 / h[0], h[13] / 1e3 / h[0]);
    if (h[0])
      fprintf(stderr, "[klsh]   stage kcycles/window: index+meta %.1f rows %.1f norms %.1f fp16 %.1f\n", h[18] / 1e3 / h[0],
              h[19] / 1e3 / h[0], h[20] / 1e3 / h[0], h[21] / 1e3 / h[0]);
    if (h[0])
      fprintf(stderr, "[klsh]   resolver kcycles/window: validate %.1f (%.2f calls) merge-step %.1f apply %.1f\n", h[14] / 1e3 / h[0],
              (double)h[17] / h[0], h[15] / 1e3 / h[0], h[16] / 1e3 / h[0]);

  }
  return KLSH_OK;
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
  if (ctx->debug) KTRY(dev_reserve(ctx, ctx->dbg, sizeof(unsigned long long) * 24));

  MergeArgs A;
  A.vals = ctx->cur.vals.as<float>();
  A.D = ctx->D;
  A.ld = ctx->ld;
  A.cnt = ctx->cur.cnt.as<int32_t>();
  A.head = ctx->cur.head.as<int32_t>();
  A.tail = ctx->cur.tail.as<int32_t>();
  A.next = ctx->cur.next.as<int32_t>();
  A.rows_sorted = rows_sorted;
  A.bstart = s.bstart.as<uint32_t>();
  A.pos_nrm = s.pos_nrm.as<float>();
  A.pos_h = s.pos_h.as<uint4>();
  A.dbg = ctx->debug ? ctx->dbg.as<unsigned long long>() : nullptr;
  A.mg = ctx->mg;
  A.threshold = threshold;

  // stage 0: one CTA per bucket, biggest buckets first
  A.list_a = s.list_big.as<uint32_t>();
  A.n_a = &dc->n_big;
  A.list_b = s.list_large.as<uint32_t>();
  A.n_b = &dc->n_large;
  A.cursor = &dc->large_cursor;
  A.esc_list = s.esc1.as<uint32_t>();
  A.esc_count = &dc->n_esc1;
  A.max_reps = ctx->cta_max;
  KTRY(launch_stage(ctx, 0, 1, A, n_items_host));
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
  KTRY(launch_stage(ctx, 1, ctx->cluster_size, A, 0));
  if (bucket_max_host <= ctx->cluster_max) return KLSH_OK;

  // stage 2: one large cluster per bucket
  A.list_a = s.esc2.as<uint32_t>();
  A.n_a = &dc->n_esc2;
  A.cursor = &dc->cluster2_cursor;
  A.esc_list = s.esc3.as<uint32_t>();
  A.esc_count = &dc->n_esc3;
  A.max_reps = ctx->cluster2_max;
  KTRY(launch_stage(ctx, 1, ctx->cluster2_size, A, 0));
  if (bucket_max_host <= ctx->cluster2_max) return KLSH_OK;

  // stage 3: the whole grid per bucket
  A.list_a = s.esc3.as<uint32_t>();
  A.n_a = &dc->n_esc3;
  A.cursor = nullptr;
  A.esc_list = nullptr;
  A.esc_count = nullptr;
  A.max_reps = 0xFFFFFFFFu;
  KTRY(launch_stage(ctx, 2, 1, A, 0));
  return KLSH_OK;
}
