// Hand-written sm_100a kernels of the kmerLSH mode-C clustering hot path.
//
// Arithmetic contract (DESIGN.md "Exactness"): every floating-point operation that decides a
// key bit, a merge, or a centroid value is an IEEE binary32 operation with one rounding, in the
// reference's source order, and NEVER a fused multiply-add — the reference's x86-64 build has no
// FMA (SURVEY.md section 7 hard part 2).  This file is compiled with -fmad=false and uses the
// explicit-rounding intrinsics (__fmul_rn, __fadd_rn, __fdiv_rn, __fsqrt_rn), which ptxas never
// contracts.
#include <cstddef>
#include <cstring>

#include "exact_math.cuh"
#include "klsh_internal.cuh"
#include "sign_umma.cuh"

namespace {

constexpr int kScanTile = 4096;  // elements per block in the count/scatter style kernels

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// ================================================================================================
// Generic helpers: block-level exclusive scan of one value per thread (blockDim.x <= 1024).
// ================================================================================================
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums /*[33]*/, uint32_t* total) {
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = (lane < nwarp) ? warp_sums[lane] : 0u;
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (uint32_t)o) winc += t;
    }
    warp_sums[lane] = winc - w;  // exclusive
    if (lane == 31) warp_sums[32] = winc;
  }
  __syncthreads();
  uint32_t res = warp_sums[warp] + inc - v;
  if (total) *total = warp_sums[32];
  __syncthreads();
  return res;
}

// Exclusive scan of a small array (block counts) by ONE block; writes the grand total to *total.
__global__ void k_scan_single(uint32_t* data, uint32_t n, uint32_t* total) {
  __shared__ uint32_t ws[33];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = (i < n) ? data[i] : 0u;
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(v, ws, &tot);
    if (i < n) data[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void k_iota(uint32_t* out, uint64_t n, uint32_t base) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = base + (uint32_t)i;
}

__global__ void k_fill(uint32_t* out, uint64_t from, uint64_t to, uint32_t v) {
  uint64_t i = from + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < to) out[i] = v;
}

// ================================================================================================
// Row transform (reference IOMat::convertHTMat, io/ioMatrix.cc:353-408).
//   counts: sample-major uint16 [D][batch]; lut[c] = float(log(c+1.0)) from the host libm.
//   pass 1: keep flag per row (sum_j cnt > 0.1*D, uint64 vs double) -> per-block counts
//   pass 2: kept rows written densely, value_j = lut[cnt] - v_kmers[j]; slot = original index i
// ================================================================================================
__device__ __forceinline__ bool row_kept(const uint16_t* counts, uint64_t batch, int D, uint64_t i) {
  unsigned long long total = 0;
  for (int j = 0; j < D; ++j) total += counts[(uint64_t)j * batch + i];
  return (double)total > 0.1 * (double)D;
}

__global__ void k_transform_count(const uint16_t* __restrict__ counts, uint64_t batch, int D, uint32_t* blkcnt) {
  __shared__ uint32_t ws[33];
  uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t c = 0;
  for (int k = 0; k < kScanTile / 256; ++k) {
    uint64_t i = base + (uint64_t)k * 256 + threadIdx.x;
    if (i < batch && row_kept(counts, batch, D, i)) ++c;
  }
  uint32_t tot;
  block_exclusive_scan(c, ws, &tot);
  if (threadIdx.x == 0) blkcnt[blockIdx.x] = tot;
}

__global__ void k_transform_write(const uint16_t* __restrict__ counts, const float* __restrict__ lut,
                                  const float* __restrict__ vk, uint64_t batch, int D, int ld,
                                  const uint32_t* __restrict__ blkoff, float* vals, MetaCol cnt, MetaCol head,
                                  MetaCol tail, uint64_t row_base) {
  __shared__ uint32_t ws[33];
  uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t run = blkoff[blockIdx.x];
  for (int k = 0; k < kScanTile / 256; ++k) {
    uint64_t i = base + (uint64_t)k * 256 + threadIdx.x;
    bool keep = (i < batch) && row_kept(counts, batch, D, i);
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(keep ? 1u : 0u, ws, &tot);
    if (keep) {
      uint64_t r = row_base + run + ex;
      float* dst = vals + r * (uint64_t)ld;
      for (int j = 0; j < D; ++j) {
        uint16_t cc = counts[(uint64_t)j * batch + i];
        dst[j] = __fsub_rn(lut[cc], vk[j]);
      }
      for (int j = D; j < ld; ++j) dst[j] = 0.f;
      cnt[r] = 1;
      head[r] = (int32_t)i;  // member slot = original index within the batch
      tail[r] = (int32_t)i;
    }
    run += tot;
  }
}

__global__ void k_fill_i32(int32_t* p, uint64_t n, int32_t v) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ================================================================================================
// Signing (reference LSH::random_projection, hash/lshash.cc:44-59):
//   bit_h = (sum_h >= 0), sum_h = fl(fl(... fl(0 + fl(w_h0*x_0)) ...) + fl(w_h,D-1 * x_D-1))
//   key = ((bit_0*2 + bit_1)*2 + ...) — plane 0 is the most significant bit.
// D <= 64: sign_umma::k_sign_umma (tcgen05.mma with the row operand and the sums in tensor memory, sign_umma.cuh).
// D > 64: k_sign_tc_wide below (mma.sync, the same product walked in 64-column chunks).  Both decide every bit with
// the reference's exact mul-then-add chain whenever the fast sum is inside its error margin, and count those rows.
// ================================================================================================

// Tensor-core signing.  ncu on the FMA kernel this replaced, at C2 scale: issue
// slots 59 % busy at 44 % occupancy, 2.7 k warp instructions per 32 rows, 1.08 TB/s — contraction
// bound, H*D fused multiply-adds per row against 4*D bytes (profiles/).  The projection is a
// [rows x D] x [D x H] product, so it goes to the tensor cores as 3xTF32: every operand is split
// into hi = tf32(v) and lo = tf32(v - hi) and the sum hi*hi + hi*lo + lo*hi is accumulated in fp32
// (tcgen05.mma kind::tf32 / mma.sync.m16n8k8, K = 8 per instruction either way).  Error budget in units of
// 2^-24 * sum|w_i x_i| (<= 2^-24 * |w| * |x|):
//   28   the dropped lo*lo products and the split residues: rows are split by truncation (residue < 2^-20 |x|),
//        planes by rounding (2^-22 |w|): (2^-21 + 2^-20 + 2^-22) per term;
//   10   per mma for the tensor core's truncating fp32 accumulation (alignment of the 8 products to the
//        largest exponent with a few guard bits + the final truncation; a worst-case figure, the
//        measured behaviour is far better), 3*ceil(D/8) mma per sum;
//   D    the reference chain's own rounding (its sign is what has to be reproduced).
//    2   the planes are multiplied as unit vectors w/|w| (a positive factor keeps the sign; the scaling rounds each
//        element once, and |w/|w|| <= 1 + 2^-22), which makes the margin the same for every plane of a row.
// So   eps = (D + 34 + 30*ceil(D/8)) * 2^-24 * |x|   (D = 32: 186 * 2^-24 = 1.1e-5) on the sum with the unit plane
// guarantees that a sum outside it has the sign of the reference's mul-then-add chain; a sum inside it
// is re-evaluated with the reference's exact arithmetic (and the row counted), so every key bit is
// the reference's.
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// Row operands are split by masking: hi = the top 19 bits (what a tf32 operand keeps; mma ignores the low 13
// bits of a .tf32 register), lo = x - hi, exact in fp32 and itself truncated by the mma.  One LOP and one
// FADD per element instead of two cvt.rna and an FADD; the residue is < 2^-20 |x| instead of 2^-22 |x|,
// which the error budget below accounts for.
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void sign_cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// Rows wider than 64 floats: the same 3xTF32 product, walked in chunks of 64 columns.  The accumulators
// stay in registers across the chunks, the 32 x 64 chunk tiles are double-buffered through cp.async
// (chunk after chunk, tile after tile, one flat sequence), and the split planes for the whole width
// sit in shared memory in B-fragment order.  The exact re-evaluation reads row and plane from global
// memory (rare).
__global__ void __launch_bounds__(256)
k_sign_tc_wide(const float* __restrict__ vals, int D, int ld, const uint32_t* __restrict__ rows, uint64_t n,
               const float* __restrict__ planes, int H, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ rows_out,
               unsigned long long* eps_rows, uint32_t key_or) {
  constexpr int CW = 64, TS = CW + 4, KS8 = CW / 8;
  const int nch = (ld + CW - 1) / CW;
  const int nwarp = blockDim.x >> 5;
  extern __shared__ __align__(16) float smem[];
  float* pn = smem;                                            // eps factor per plane [32]
  uint4* bfrag = reinterpret_cast<uint4*>(pn + 32);            // [nch][KS8][4][32] {hi b0, hi b1, lo b0, lo b1}
  float* tiles = reinterpret_cast<float*>(bfrag + (size_t)nch * KS8 * 4 * 32);  // [nwarp][2][32][TS]
  for (int i = threadIdx.x; i < nwarp * 2 * 32 * TS; i += blockDim.x) tiles[i] = 0.f;
  for (int h = threadIdx.x; h < 32; h += blockDim.x) {
    float m = 0.f;
    if (h < H)
      for (int i = 0; i < D; ++i) m = __fmaf_rn(planes[h * ld + i], planes[h * ld + i], m);
    pn[h] = sqrtf(m) * (((float)D + 32.f + 240.f + (float)nch) * 5.9604645e-8f);  // 24 mma per 64-column chunk
  }
  for (int i = threadIdx.x; i < nch * KS8 * 4 * 32; i += blockDim.x) {
    const int l = i & 31, nt = (i >> 5) & 3, ks = i >> 7;  // ks runs over the whole width
    const int h = nt * 8 + (l >> 2), k0 = ks * 8 + (l & 3);
    const float w0 = (h < H && k0 < ld) ? planes[h * ld + k0] : 0.f;
    const float w1 = (h < H && k0 + 4 < ld) ? planes[h * ld + k0 + 4] : 0.f;
    const uint32_t h0 = tf32_rna(w0), h1 = tf32_rna(w1);
    bfrag[i] = make_uint4(h0, h1, tf32_rna(w0 - __uint_as_float(h0)), tf32_rna(w1 - __uint_as_float(h1)));
  }
  __syncthreads();
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3;
  float* wt = tiles + (size_t)warp * 2 * 32 * TS;
  const uint64_t nwarps_total = (uint64_t)gridDim.x * nwarp;
  uint32_t my_eps = 0;
  float pnr[8];
  uint32_t vmask = 0u;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int h = (q >> 1) * 8 + 2 * (int)tg + (q & 1);
    pnr[q] = pn[h];
    if (h < H) vmask |= 1u << q;
  }
  // gather chunk ch of the tile starting at t0 into buffer b (r = this lane's row index)
  auto issue = [&](uint64_t t0, uint32_t r, int ch, int b) {
    if (t0 < n) {
      const int nrow = (int)min((uint64_t)32, n - t0);
      const int c0 = ch * CW;
      const int vpr = min(CW, ld - c0) >> 2;  // 16-byte pieces per row in this chunk
      const int total = nrow * vpr;
      float* tile = wt + (size_t)b * 32 * TS;
      for (int v0 = 0; v0 < total; v0 += 32) {
        const int v = v0 + (int)lane;
        const int rr = min(v, total - 1) / vpr, cc = v - rr * vpr;
        const uint32_t ri = __shfl_sync(0xffffffffu, r, rr);
        if (v < total) sign_cp_async16(tile + rr * TS + cc * 4, vals + (uint64_t)ri * ld + c0 + cc * 4);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto row_of = [&](uint64_t t0) -> uint32_t {
    const uint64_t t = t0 + lane;
    return (t < n) ? (rows ? rows[t] : (uint32_t)t) : 0u;
  };
  uint64_t t0 = ((uint64_t)blockIdx.x * nwarp + warp) * 32;
  uint32_t r_cur = row_of(t0);
  int buf = 0;
  issue(t0, r_cur, 0, 0);
  for (; t0 < n; t0 += nwarps_total * 32) {
    const uint64_t t0n = t0 + nwarps_total * 32;
    const uint32_t r_next = row_of(t0n);
    // every chunk is accumulated from zero and added to the running sum in fp32: the tensor core's
    // accumulation error is then relative to the chunk's own magnitude (24 mma), not to the whole row's
    float c[2][4][4], ct[2][4][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) ct[m][nt][e] = 0.f;
    float xx = 0.f;
    for (int ch = 0; ch < nch; ++ch, buf ^= 1) {
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) c[m][nt][e] = 0.f;
      if (ch + 1 < nch) issue(t0, r_cur, ch + 1, buf ^ 1);
      else issue(t0n, r_next, 0, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncwarp();
      const float* tile = wt + (size_t)buf * 32 * TS;
      {
        const float4* x4 = reinterpret_cast<const float4*>(tile + lane * TS);
        const int nq = min(CW, ld - ch * CW) >> 2;
        for (int q = 0; q < nq; ++q) {
          const float4 x = x4[q];
          xx = __fmaf_rn(x.x, x.x, xx);
          xx = __fmaf_rn(x.y, x.y, xx);
          xx = __fmaf_rn(x.z, x.z, xx);
          xx = __fmaf_rn(x.w, x.w, xx);
        }
      }
#pragma unroll
      for (int ks = 0; ks < KS8; ++ks) {
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const float* p0 = tile + (m * 16 + g) * TS + ks * 8 + tg;
          const float a0 = p0[0], a1 = p0[8 * TS], a2 = p0[4], a3 = p0[8 * TS + 4];
          tf32_split(a0, ahi[m][0], alo[m][0]);
          tf32_split(a1, ahi[m][1], alo[m][1]);
          tf32_split(a2, ahi[m][2], alo[m][2]);
          tf32_split(a3, ahi[m][3], alo[m][3]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          if (nt * 8 < H) {  // warp-uniform
            const uint4 b = bfrag[((size_t)(ch * KS8 + ks) * 4 + nt) * 32 + lane];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              mma_tf32(c[m][nt], alo[m], b.x, b.y);
              mma_tf32(c[m][nt], ahi[m], b.z, b.w);
              mma_tf32(c[m][nt], ahi[m], b.x, b.y);
            }
          }
      }
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) ct[m][nt][e] = __fadd_rn(ct[m][nt][e], c[m][nt][e]);
      __syncwarp();
    }
    const float xn = sqrtf(xx);
    uint32_t part[4];
    uint32_t slow = 0u;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int row = m * 16 + half * 8 + (int)g;
        const float xnr = __shfl_sync(0xffffffffu, xn, row);
        const uint32_t rrow = __shfl_sync(0xffffffffu, r_cur, row);
        uint32_t bits = 0u, flag = 0u;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float sum = ct[m][q >> 1][2 * half + (q & 1)];
          const float a = fabsf(sum);
          bits |= (sum >= 0.f ? 1u : 0u) << q;
          flag |= ((a > pnr[q] * xnr && a <= 3.0e38f) ? 0u : 1u) << q;
        }
        flag &= vmask;
        if (t0 + row >= n) flag = 0u;  // rows past the end of the last tile hold stale data
        while (flag) {  // reference arithmetic: sum = fl(sum + fl(w_i * x_i)), i ascending (hash/lshash.cc:44-51)
          const int q = __ffs(flag) - 1;
          flag &= flag - 1;
          const float* w = planes + ((q >> 1) * 8 + 2 * (int)tg + (q & 1)) * ld;
          const float* x = vals + (uint64_t)rrow * ld;
          float sum = 0.f;
          for (int j = 0; j < D; ++j) sum = __fadd_rn(sum, __fmul_rn(__ldg(w + j), __ldg(x + j)));
          bits = (bits & ~(1u << q)) | ((sum >= 0.f ? 1u : 0u) << q);
          slow |= 1u << (m * 2 + half);
        }
        bits &= vmask;
        uint32_t byplane = 0u;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) byplane |= ((bits >> (2 * nt)) & 3u) << (nt * 8);
        byplane <<= 2 * tg;
        part[m * 2 + half] = H ? (__brev(byplane) >> (32 - H)) : 0u;
      }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      part[q] |= __shfl_xor_sync(0xffffffffu, part[q], 1);
      part[q] |= __shfl_xor_sync(0xffffffffu, part[q], 2);
    }
    slow |= __shfl_xor_sync(0xffffffffu, slow, 1);
    slow |= __shfl_xor_sync(0xffffffffu, slow, 2);
    {
      const uint32_t key = tg == 0 ? part[0] : (tg == 1 ? part[1] : (tg == 2 ? part[2] : part[3]));
      const uint64_t t = t0 + tg * 8 + g;
      if (t < n) {
        keys_out[t] = key | key_or;
        my_eps += (slow >> tg) & 1u;
      }
      const uint64_t tl = t0 + lane;
      if (tl < n) rows_out[tl] = r_cur;
    }
    r_cur = r_next;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (eps_rows) {
    const uint32_t tot = __reduce_add_sync(0xffffffffu, my_eps);
    if (lane == 0 && tot) atomicAdd(eps_rows, (unsigned long long)tot);
  }
}

// ================================================================================================
// Stable LSD radix sort of (key, row) pairs, 8 bits per pass.
// Equivalent to the reference's merge_hashtable (function/cluster.cc:15-30): push rows in input
// order into a dense table indexed by key, then visit keys ascending.
//   k_radix_hist    : per-block digit histogram          -> hist[digit][block]
//   k_radix_scan    : per digit, exclusive scan over blocks; digit totals -> dtot[digit]
//   k_radix_scatter : block re-ranks its tile stably and scatters
// A block's tile is 8 warps x 16 rounds x 32 lanes, warp-major, so that (warp, round, lane) order
// is input order.
// ================================================================================================
constexpr int kRadixWarps = 8;
constexpr int kRadixRounds = 16;
constexpr int kRadixTile = kRadixWarps * kRadixRounds * 32;  // 4096

__global__ void __launch_bounds__(kRadixWarps * 32)
k_radix_hist(const uint32_t* __restrict__ keys, uint64_t n, int shift, uint32_t* __restrict__ hist, uint32_t nblk) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  uint64_t base = (uint64_t)blockIdx.x * kRadixTile;
  for (int k = 0; k < kRadixTile / 256; ++k) {
    uint64_t i = base + (uint64_t)k * 256 + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// one block per digit
__global__ void k_radix_scan(uint32_t* hist, uint32_t nblk, uint32_t* dtot) {
  __shared__ uint32_t ws[33];
  uint32_t* row = hist + (uint64_t)blockIdx.x * nblk;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nblk; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = (i < nblk) ? row[i] : 0u;
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(v, ws, &tot);
    if (i < nblk) row[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) dtot[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(kRadixWarps * 32)
k_radix_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rows, uint64_t n, int shift,
                const uint32_t* __restrict__ hist, const uint32_t* __restrict__ dtot, uint32_t nblk,
                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ rows_out) {
  __shared__ uint32_t wc[kRadixWarps][256];  // per warp and digit: count, then exclusive prefix over warps
  __shared__ uint32_t dbase[256];            // global position of this block's first element of each digit
  __shared__ uint32_t tstart[256];           // position of each digit's run inside the block's sorted tile
  __shared__ uint32_t skey[kRadixTile], srow[kRadixTile];
  __shared__ uint32_t ws[33];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  for (int w = 0; w < kRadixWarps; ++w) wc[w][threadIdx.x] = 0;
  {
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(dtot[threadIdx.x], ws, &tot);
    dbase[threadIdx.x] = ex + hist[(uint64_t)threadIdx.x * nblk + blockIdx.x];
  }
  __syncthreads();
  uint32_t k[kRadixRounds], r[kRadixRounds];
  uint16_t rank[kRadixRounds];
  const uint64_t tile_base = (uint64_t)blockIdx.x * kRadixTile;
  const uint64_t base = tile_base + (uint64_t)warp * (kRadixRounds * 32);
#pragma unroll
  for (int rd = 0; rd < kRadixRounds; ++rd) {
    uint64_t i = base + (uint64_t)rd * 32 + lane;
    bool valid = i < n;
    k[rd] = valid ? keys[i] : 0u;
    r[rd] = valid ? rows[i] : 0u;
    uint32_t d = valid ? ((k[rd] >> shift) & 255u) : 0xFFFFFFFFu;
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t before = __popc(peers & lanemask_lt());
    uint32_t old = 0;
    if (valid) {
      int leader = __ffs(peers) - 1;
      if ((int)lane == leader) {
        old = wc[warp][d];
        wc[warp][d] = old + __popc(peers);
      }
      old = __shfl_sync(peers, old, leader);
    }
    rank[rd] = (uint16_t)(old + before);
    __syncwarp();
  }
  __syncthreads();
  // per digit: exclusive prefix over warps (input order inside the tile is warp-major) and the tile total
  uint32_t tile_cnt;
  {
    uint32_t run = 0;
    for (int w = 0; w < kRadixWarps; ++w) {
      uint32_t c = wc[w][threadIdx.x];
      wc[w][threadIdx.x] = run;
      run += c;
    }
    tile_cnt = run;
  }
  {
    uint32_t tot;
    tstart[threadIdx.x] = block_exclusive_scan(tile_cnt, ws, &tot);
  }
  __syncthreads();
  // stage the tile in digit order in shared memory ...
#pragma unroll
  for (int rd = 0; rd < kRadixRounds; ++rd) {
    uint64_t i = base + (uint64_t)rd * 32 + lane;
    if (i < n) {
      uint32_t d = (k[rd] >> shift) & 255u;
      uint32_t lp = tstart[d] + wc[warp][d] + rank[rd];
      skey[lp] = k[rd];
      srow[lp] = r[rd];
    }
  }
  __syncthreads();
  // ... and write every digit's run to its global position with consecutive threads
  const uint32_t tile_n = (uint32_t)min((uint64_t)kRadixTile, n - tile_base);
  for (uint32_t i = threadIdx.x; i < tile_n; i += kRadixWarps * 32) {
    const uint32_t kk = skey[i];
    const uint32_t d = (kk >> shift) & 255u;
    const uint32_t dst = dbase[d] + (i - tstart[d]);
    keys_out[dst] = kk;
    rows_out[dst] = srow[i];
  }
}

// ================================================================================================
// Bucket boundaries (run-length encode of sorted keys) and size classes.
// ================================================================================================
__device__ __forceinline__ bool is_head(const uint32_t* keys, uint64_t i) { return i == 0 || keys[i] != keys[i - 1]; }

__global__ void k_heads_count(const uint32_t* __restrict__ keys, uint64_t n, uint32_t* blkcnt) {
  __shared__ uint32_t ws[33];
  uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t c = 0;
  for (int k = 0; k < kScanTile / 256; ++k) {
    uint64_t i = base + (uint64_t)k * 256 + threadIdx.x;
    if (i < n && is_head(keys, i)) ++c;
  }
  uint32_t tot;
  block_exclusive_scan(c, ws, &tot);
  if (threadIdx.x == 0) blkcnt[blockIdx.x] = tot;
}

__global__ void k_heads_write(const uint32_t* __restrict__ keys, uint64_t n, const uint32_t* __restrict__ blkoff,
                              uint32_t* bstart) {
  __shared__ uint32_t ws[33];
  uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t run = blkoff[blockIdx.x];
  for (int k = 0; k < kScanTile / 256; ++k) {
    uint64_t i = base + (uint64_t)k * 256 + threadIdx.x;
    bool hd = (i < n) && is_head(keys, i);
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(hd ? 1u : 0u, ws, &tot);
    if (hd) bstart[run + ex] = (uint32_t)i;
    run += tot;
  }
}

// nest_threshold < 0: nesting disabled.  Reads the bucket count from counters->n_buckets.
// Appends are warp-aggregated: one atomic per warp and class.
__device__ __forceinline__ void warp_append(bool pred, uint32_t* counter, uint32_t* list, uint32_t value) {
  const uint32_t m = __ballot_sync(0xffffffffu, pred);
  if (m == 0u) return;
  uint32_t base = 0;
  const int leader = __ffs(m) - 1;
  if ((int)lane_id() == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) list[base + __popc(m & lanemask_lt())] = value;
}

__device__ __forceinline__ void warp_append_item(bool pred, uint32_t* counter, uint32_t* list, uint32_t bucket) {
  const uint32_t m = __ballot_sync(0xffffffffu, pred);
  if (m == 0u) return;
  uint32_t base = 0;
  const int leader = __ffs(m) - 1;
  if ((int)lane_id() == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) {
    uint32_t* it = list + 3 * (size_t)(base + __popc(m & lanemask_lt()));
    it[0] = bucket;
    it[1] = 0u;  // not started
    it[2] = 0u;
  }
}

// Bucket-size histogram by power-of-two class (only classes that can be "direct" matter): the input of
// the direct threshold below.
__global__ void k_size_hist(const uint32_t* __restrict__ bstart, uint64_t n, uint32_t b_lo, uint32_t b_hi, long long nest_threshold,
                            PassCounters* counters) {
  const uint32_t nb = counters->n_buckets;
  for (uint32_t b0 = blockIdx.x * blockDim.x; b0 < nb; b0 += gridDim.x * blockDim.x) {
    const uint32_t b = b0 + threadIdx.x;
    uint32_t size = 0;
    if (b < nb && b >= b_lo && b < b_hi) {
      const uint32_t e = (b + 1 < nb) ? bstart[b + 1] : (uint32_t)n;
      size = e - bstart[b];
      if (nest_threshold >= 0 && (long long)size > nest_threshold) size = 0;  // oversized buckets are not merged here
    }
    if (size > KLSH_SMALL_MAX) atomicAdd(&counters->size_hist[31 - __clz(size)], 1u);  // few buckets are this large
  }
}

// Buckets with at least direct_threshold() rows start on cluster teams on the second stream.  The
// threshold is the smallest power of two >= direct_min that leaves at most max_direct such buckets
// (there are only a few dozen cluster teams: more items than that would queue behind each other
// while the single-CTA stage runs hundreds of buckets side by side).
__device__ __forceinline__ uint32_t direct_threshold(const PassCounters* c, uint32_t direct_min, uint32_t max_direct) {
  uint32_t k = 31u - (uint32_t)__clz(max(direct_min, 2u));
  if ((1u << k) < direct_min) ++k;  // ceil(log2(direct_min))
  for (; k < 32u; ++k) {
    uint32_t above = 0;
    for (uint32_t j = k; j < 32u; ++j) above += c->size_hist[j];
    if (above <= max_direct) break;
  }
  return k >= 32u ? 0xFFFFFFFFu : max(1u << k, direct_min);
}

// Only buckets in [b_lo, b_hi) are queued for merging (multi-GPU: each rank owns a contiguous
// range); oversized buckets are listed regardless of the range because every rank has to draw their
// hash tables to keep the hyperplane stream in step.
__global__ void k_classify(uint32_t* bstart, uint64_t n, long long nest_threshold, uint32_t b_lo, uint32_t b_hi, uint32_t direct_min,
                           uint32_t max_direct, PassCounters* counters, uint32_t* list_small, uint32_t* list_large, uint32_t* list_big,
                           uint32_t* list_direct, uint32_t* list_nested) {
  const uint32_t nb = counters->n_buckets;
  const uint32_t dthr = direct_threshold(counters, direct_min, max_direct);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    bstart[nb] = (uint32_t)n;
    counters->direct_thr = dthr;
  }
  uint32_t tmax = 0;
  // grid-stride over the buckets (their number is only known on the device); whole warps iterate together
  for (uint32_t b0 = blockIdx.x * blockDim.x; b0 < nb; b0 += gridDim.x * blockDim.x) {
    const uint32_t b = b0 + threadIdx.x;
    uint32_t size = 0;
    const bool mine = b >= b_lo && b < b_hi;
    if (b < nb) {
      const uint32_t s = bstart[b];
      const uint32_t e = (b + 1 < nb) ? bstart[b + 1] : (uint32_t)n;
      size = e - s;
    }
    tmax = max(tmax, size);
    const bool nested = nest_threshold >= 0 && (long long)size > nest_threshold && size >= 2;
    const bool merged = mine && !nested;
    const bool small = merged && size >= 2 && size <= KLSH_SMALL_MAX;
    const bool direct = merged && size > KLSH_SMALL_MAX && size >= dthr;
    const bool large = merged && size > KLSH_SMALL_MAX && size < KLSH_BIG && !direct;
    const bool big = merged && size >= KLSH_BIG && !direct;
    warp_append(nested, &counters->n_nested, list_nested, b);
    warp_append(small, &counters->n_small, list_small, b);
    warp_append_item(large, &counters->n_large, list_large, b);
    warp_append_item(big, &counters->n_big, list_big, b);
    warp_append_item(direct, &counters->n_direct, list_direct, b);
  }
  const uint32_t wmax = __reduce_max_sync(0xffffffffu, tmax);
  if (lane_id() == 0 && wmax > 0) atomicMax(&counters->bucket_max, wmax);
}

// The direct list in order of decreasing bucket size: its buckets are the pass's longest window chains, and
// a long chain picked up late by a team that first worked through a shorter one ends the pass late.
__global__ void k_order_direct(const uint32_t* __restrict__ bstart, const PassCounters* counters, uint32_t* list_direct) {
  __shared__ uint32_t sb[1024], ss[1024];
  const uint32_t n = min(counters->n_direct, 1024u);  // a longer list keeps its tail as it is
  const uint32_t i = threadIdx.x;
  if (i < n) {
    sb[i] = list_direct[3 * i];
    ss[i] = bstart[sb[i] + 1] - bstart[sb[i]];
  }
  __syncthreads();
  if (i < n) {
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; ++j) rank += (ss[j] > ss[i] || (ss[j] == ss[i] && sb[j] < sb[i])) ? 1u : 0u;
    list_direct[3 * rank] = sb[i];
  }
}

// ================================================================================================
// Greedy in-bucket merge (reference p_cluster, function/cluster.cc:56-87) with
//   Distance::cosine   (function/distance.cc:27-38)
//   AB::SetConsensus   (function/funcAB.cc:49-71)
// Bucket = positions [s, e) of rows_sorted (row indices in bucket order).  On return
// rows_sorted[s .. s+size) are the survivors in the reference's order and rows_sorted[s+size .. e)
// hold KLSH_SENTINEL.  A merge of current (position i) into candidate (position j<i) overwrites
// candidate's row in the arena with the consensus, prepends current's member chain to
// candidate's, and moves the tail position into i (swap-remove).
// ================================================================================================

__device__ __forceinline__ float row_norm(const float* v, int D) {  // unpadded row (fallback kernel)
  float m = 0.f;
  for (int i = 0; i < D; ++i) m = __fadd_rn(m, __fmul_rn(v[i], v[i]));
  return __fsqrt_rn(m);
}

// ---- small buckets: one warp per bucket, rows resident in shared memory -------------------------
__global__ void __launch_bounds__(128)
k_merge_small(float* __restrict__ vals, int D, int ld, MetaCol cnt, MetaCol head,
              MetaCol tail, int32_t* __restrict__ next, uint32_t* __restrict__ rows_sorted,
              const uint32_t* __restrict__ bstart, const uint32_t* __restrict__ list, const PassCounters* counters,
              float threshold, MgLog mg) {
  extern __shared__ __align__(16) float smem[];
  const int stride = ld + 4;  // rows 16 bytes apart modulo 128: per-lane float4 walks are conflict-free
  const int nq = ld >> 2;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* tile = smem + (size_t)warp * KLSH_SMALL_MAX * stride;
  const uint32_t nlist = counters->n_small;
  for (uint32_t w = blockIdx.x * wpb + warp; w < nlist; w += gridDim.x * wpb) {
    const uint32_t b = list[w];
    const uint32_t s = bstart[b];
    const int n = (int)(bstart[b + 1] - s);
    // slot = lane: row index and metadata of the row originally at position `lane`
    uint32_t ridx = (lane < (uint32_t)n) ? rows_sorted[s + lane] : 0u;
    // all rows of the bucket are requested before any is waited for (cp.async, 16 bytes per lane)
    {
      const int total = n * nq;
      for (int v0 = 0; v0 < total; v0 += 32) {
        const int v = v0 + (int)lane;
        const int rr = min(v, total - 1) / nq, cc = v - rr * nq;
        const uint32_t rk = __shfl_sync(0xffffffffu, ridx, rr);
        if (v < total) {
          const uint32_t dsts = (uint32_t)__cvta_generic_to_shared(tile + rr * stride + cc * 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dsts), "l"(vals + (uint64_t)rk * ld + cc * 4) : "memory");
        }
      }
    }
    int my_cnt = 0, my_head = -1, my_tail = -1;
    float my_nrm = 0.f;
    bool my_dirty = false;
    if (lane < (uint32_t)n) {
      const int4 m = __ldcg(reinterpret_cast<const int4*>(cnt.p) + ridx);  // one 16-byte record {cnt, head, tail, 0}
      my_cnt = m.x;
      my_head = m.y;
      my_tail = m.z;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    if (lane < (uint32_t)n) my_nrm = norm_seq(reinterpret_cast<const float4*>(tile + lane * stride), nq);
    int pos_slot = (int)lane;  // slot held by position `lane`
    int size = n, i = 1;
    while (i < size) {
      const int cs = __shfl_sync(0xffffffffu, pos_slot, i);
      const float cn = __shfl_sync(0xffffffffu, my_nrm, cs);
      const float rn = __shfl_sync(0xffffffffu, my_nrm, pos_slot & 31);
      bool match = false;
      if ((int)lane < i) {
        const float4* c = reinterpret_cast<const float4*>(tile + cs * stride);
        const float4* r = reinterpret_cast<const float4*>(tile + pos_slot * stride);
        float dot = 0.f;
        for (int q = 0; q < nq; ++q) {  // padded products are +0: the D-term sum of the reference
          const float4 x = c[q], y = r[q];
          dot = __fadd_rn(dot, __fmul_rn(x.x, y.x));
          dot = __fadd_rn(dot, __fmul_rn(x.y, y.y));
          dot = __fadd_rn(dot, __fmul_rn(x.z, y.z));
          dot = __fadd_rn(dot, __fmul_rn(x.w, y.w));
        }
        match = cos_match(dot, cn, rn, threshold);
      }
      const uint32_t m = __ballot_sync(0xffffffffu, match);
      if (m == 0u) {
        ++i;
        continue;
      }
      const int j = __ffs(m) - 1;
      const int rs = __shfl_sync(0xffffffffu, pos_slot, j);
      const int c1 = __shfl_sync(0xffffffffu, my_cnt, cs);
      const int c2 = __shfl_sync(0xffffffffu, my_cnt, rs);
      const int h1 = __shfl_sync(0xffffffffu, my_head, cs);
      const int t1 = __shfl_sync(0xffffffffu, my_tail, cs);
      const int h2 = __shfl_sync(0xffffffffu, my_head, rs);
      {
        float* c = tile + cs * stride;
        float* r = tile + rs * stride;
        for (int d = lane; d < D; d += 32) r[d] = consensus1(c[d], c1, r[d], c2);
      }
      __syncwarp();
      if ((int)lane == rs) {
        // ids(current) ++ ids(candidate): current's chain goes first
        if (t1 >= 0) {
          next[t1] = h2;
          if (mg.counts) {
            const uint32_t k = atomicAdd(mg.counts + 1, 1u);
            mg.next_slot[k] = (uint32_t)t1;
            mg.next_val[k] = h2;
          }
          my_head = h1;
          if (my_tail < 0) my_tail = t1;
        }
        my_cnt = c1 + c2;
        my_nrm = norm_seq(reinterpret_cast<const float4*>(tile + rs * stride), nq);
        my_dirty = true;
      }
      // position i takes the tail position's slot
      const int last = __shfl_sync(0xffffffffu, pos_slot, size - 1);
      if ((int)lane == i) pos_slot = last;
      --size;
      __syncwarp();
    }
    // write back: survivors' order, dirty rows and metadata
    {
      const uint32_t my_ridx_for_pos = __shfl_sync(0xffffffffu, ridx, pos_slot & 31);
      if (lane < (uint32_t)n) rows_sorted[s + lane] = ((int)lane < size) ? my_ridx_for_pos : KLSH_SENTINEL;
    }
    const uint32_t dirty = __ballot_sync(0xffffffffu, my_dirty);
    if (my_dirty) {
      *(reinterpret_cast<int4*>(cnt.p) + ridx) = make_int4(my_cnt, my_head, my_tail, 0);
      if (mg.counts) mg.mod_rows[atomicAdd(mg.counts, 1u)] = ridx;
    }
    uint32_t dm = dirty;
    while (dm) {
      const int sl = __ffs(dm) - 1;
      dm &= dm - 1;
      const uint32_t rk = __shfl_sync(0xffffffffu, ridx, sl);
      float* dst = vals + (uint64_t)rk * ld;
      for (int d = lane; d < D; d += 32) dst[d] = tile[sl * stride + d];
    }
    __syncwarp();
  }
}

// ---- large buckets: one block per bucket, representatives cached in shared memory ---------------
constexpr int kLargeThreads = 256;

struct LargeShared {
  int match_j;
  int c1, c2, h1, t1, h2;
  float cand_nrm;
  uint32_t cand_ridx;
  uint32_t work;
};

// merges ONE bucket [s, e) with the whole block; rep_cap = representatives that fit in smem
__device__ void merge_large_bucket(float* vals, int D, int ld, MetaCol cnt, MetaCol head, MetaCol tail,
                                   int32_t* next, uint32_t* seg, uint32_t n, float threshold, float* cand,
                                   float* rep_nrm, float* rep_rows, int rep_cap, float* nrm_spill, LargeShared* sh) {
  const int stride = ld + 1;
  const int tid = threadIdx.x;
  // representative 0
  {
    const uint32_t r0 = seg[0];
    const float* src = vals + (uint64_t)r0 * ld;
    for (int d = tid; d < D; d += blockDim.x) rep_rows[d] = src[d];
    __syncthreads();
    if (tid == 0) rep_nrm[0] = row_norm(rep_rows, D);
  }
  uint32_t size = n, i = 1;
  __syncthreads();
  while (i < size) {
    // stage the candidate
    const uint32_t cr = seg[i];
    {
      const float* src = vals + (uint64_t)cr * ld;
      for (int d = tid; d < D; d += blockDim.x) cand[d] = src[d];
      if (tid == 0) sh->match_j = 0x7fffffff;
    }
    __syncthreads();
    if (tid == 0) sh->cand_nrm = row_norm(cand, D);
    __syncthreads();
    const float cn = sh->cand_nrm;
    int best = 0x7fffffff;
    for (uint32_t j = tid; j < i; j += blockDim.x) {
      float dot = 0.f, rn;
      if ((int)j < rep_cap) {
        const float* r = rep_rows + (size_t)j * stride;
        for (int d = 0; d < D; ++d) dot = __fadd_rn(dot, __fmul_rn(cand[d], r[d]));
        rn = rep_nrm[j];
      } else {
        const float* r = vals + (uint64_t)seg[j] * ld;
        for (int d = 0; d < D; ++d) dot = __fadd_rn(dot, __fmul_rn(cand[d], r[d]));
        rn = nrm_spill[j - rep_cap];
      }
      if (cos_match(dot, cn, rn, threshold)) {
        best = (int)j;
        break;  // this thread's later j are larger
      }
    }
    // block-wide minimum
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane_id() == 0 && best != 0x7fffffff) atomicMin(&sh->match_j, best);
    __syncthreads();
    const int j = sh->match_j;
    if (j == 0x7fffffff) {
      // no merge: the candidate becomes representative i
      if ((int)i < rep_cap) {
        float* dst = rep_rows + (size_t)i * stride;
        for (int d = tid; d < D; d += blockDim.x) dst[d] = cand[d];
        if (tid == 0) rep_nrm[i] = cn;
      } else if (tid == 0) {
        nrm_spill[i - rep_cap] = cn;
      }
      ++i;
      __syncthreads();
      continue;
    }
    const uint32_t rr = seg[j];
    if (tid == 0) {
      sh->c1 = cnt[cr];
      sh->c2 = cnt[rr];
      sh->h1 = head[cr];
      sh->t1 = tail[cr];
      sh->h2 = head[rr];
    }
    __syncthreads();
    {
      const int c1 = sh->c1, c2 = sh->c2;
      float* g = vals + (uint64_t)rr * ld;
      if (j < rep_cap) {
        float* r = rep_rows + (size_t)j * stride;
        for (int d = tid; d < D; d += blockDim.x) {
          float v = consensus1(cand[d], c1, r[d], c2);
          r[d] = v;
          g[d] = v;
        }
      } else {
        for (int d = tid; d < D; d += blockDim.x) g[d] = consensus1(cand[d], c1, g[d], c2);
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (sh->t1 >= 0) {
        next[sh->t1] = sh->h2;
        head[rr] = sh->h1;
        if (tail[rr] < 0) tail[rr] = sh->t1;
      }
      cnt[rr] = sh->c1 + sh->c2;
      if (j < rep_cap) rep_nrm[j] = row_norm(rep_rows + (size_t)j * stride, D);
      else nrm_spill[j - rep_cap] = row_norm(vals + (uint64_t)rr * ld, D);
      seg[i] = seg[size - 1];
      seg[size - 1] = KLSH_SENTINEL;
    }
    --size;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kLargeThreads)
k_merge_large(float* vals, int D, int ld, MetaCol cnt, MetaCol head, MetaCol tail, int32_t* next,
              uint32_t* rows_sorted, const uint32_t* __restrict__ bstart, const uint32_t* __restrict__ list,
              const uint32_t* __restrict__ list_big, PassCounters* counters, float threshold, int rep_cap,
              float* nrm_spill_all, uint64_t spill_stride, int single_bucket_n) {
  extern __shared__ float smem[];
  __shared__ LargeShared sh;
  float* cand = smem;                 // [ld]
  float* rep_nrm = cand + ld;         // [rep_cap]
  float* rep_rows = rep_nrm + rep_cap;  // [rep_cap][ld+1]
  float* nrm_spill = nrm_spill_all + (uint64_t)blockIdx.x * spill_stride;
  if (single_bucket_n >= 0) {  // klsh_p_cluster: the whole list is one bucket
    if (blockIdx.x == 0 && single_bucket_n >= 2)
      merge_large_bucket(vals, D, ld, cnt, head, tail, next, rows_sorted, (uint32_t)single_bucket_n, threshold, cand,
                         rep_nrm, rep_rows, rep_cap, nrm_spill, &sh);
    return;
  }
  const uint32_t nbig = counters->n_big, nlist = nbig + counters->n_large;
  for (;;) {
    if (threadIdx.x == 0) sh.work = atomicAdd(&counters->large_cursor, 1u);
    __syncthreads();
    const uint32_t w = sh.work;
    __syncthreads();
    if (w >= nlist) break;
    const uint32_t b = (w < nbig) ? list_big[3 * (size_t)w] : list[3 * (size_t)(w - nbig)];
    const uint32_t s = bstart[b];
    merge_large_bucket(vals, D, ld, cnt, head, tail, next, rows_sorted + s, bstart[b + 1] - s, threshold, cand, rep_nrm,
                       rep_rows, rep_cap, nrm_spill, &sh);
    __syncthreads();
  }
}

// ================================================================================================
// Survivor compaction: keep entries != sentinel, order preserved.
// ================================================================================================
__global__ void k_alive_count(const uint32_t* __restrict__ rows, uint64_t n, uint32_t* blkcnt) {
  __shared__ uint32_t ws[33];
  uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t c = 0;
  for (int k = 0; k < kScanTile / 256; ++k) {
    uint64_t i = base + (uint64_t)k * 256 + threadIdx.x;
    if (i < n && rows[i] != KLSH_SENTINEL) ++c;
  }
  uint32_t tot;
  block_exclusive_scan(c, ws, &tot);
  if (threadIdx.x == 0) blkcnt[blockIdx.x] = tot;
}

__global__ void k_alive_write(const uint32_t* __restrict__ rows, uint64_t n, const uint32_t* __restrict__ blkoff,
                              uint32_t* out) {
  __shared__ uint32_t ws[33];
  uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t run = blkoff[blockIdx.x];
  for (int k = 0; k < kScanTile / 256; ++k) {
    uint64_t i = base + (uint64_t)k * 256 + threadIdx.x;
    uint32_t r = (i < n) ? rows[i] : KLSH_SENTINEL;
    bool keep = r != KLSH_SENTINEL;
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(keep ? 1u : 0u, ws, &tot);
    if (keep) out[run + ex] = r;
    run += tot;
  }
}

// ================================================================================================
// Export: gather surviving rows densely.
// ================================================================================================
__global__ void k_gather_rows(const float* __restrict__ vals, int D, int ld, const MetaCol cnt,
                              const MetaCol head, const uint32_t* __restrict__ rows, uint64_t n,
                              float* out_vals, int32_t* out_cnt, int32_t* out_head) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = lane_id();
  if (w >= n) return;
  const uint32_t r = rows[w];
  const float* src = vals + (uint64_t)r * ld;
  float* dst = out_vals + w * (uint64_t)D;
  for (int d = lane; d < D; d += 32) dst[d] = src[d];
  if (lane == 0) {
    out_cnt[w] = cnt[r];
    out_head[w] = head[r];
  }
}


// ================================================================================================
// Function-level entry points (parity tests): Distance::cosine and AB::SetConsensus on their own.
// ================================================================================================
// out[k] = Distance::cosine(left[k], right[k]) (function/distance.cc:27-38); rows padded to ld.
__global__ void k_cosine_pairs(const float* __restrict__ left, const float* __restrict__ right, uint64_t n, int ld, float* out) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float4* l4 = reinterpret_cast<const float4*>(left + k * (uint64_t)ld);
  const float4* r4 = reinterpret_cast<const float4*>(right + k * (uint64_t)ld);
  const int nq = ld >> 2;
  out[k] = cosine_distance(dot_seq(l4, r4, nq), norm_seq(l4, nq), norm_seq(r4, nq));
}
// out[d] = the value AB::SetConsensus gives dimension d (function/funcAB.cc:58-63)
__global__ void k_consensus(const float* __restrict__ cur, int c1, const float* __restrict__ cand, int c2, int D, float* out) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < D) out[d] = consensus1(cur[d], c1, cand[d], c2);
}
// sum of the member counts of the rows in `rows` (export sizing)
__global__ void k_sum_counts(const MetaCol cnt, const uint32_t* __restrict__ rows, uint64_t n, unsigned long long* total) {
  unsigned long long acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) acc += (unsigned long long)cnt[rows[i]];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane_id() == 0 && acc) atomicAdd(total, acc);
}

}  // namespace

// ================================================================================================
// Launch wrappers
// ================================================================================================
#define KLAUNCH(ctx)                                                                                   \
  do {                                                                                                 \
    (ctx)->launches++;                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                                              \
    if (e__ != cudaSuccess)                                                                            \
      return klsh_fail((ctx), KLSH_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                       __FILE__, __LINE__);                                                            \
  } while (0)

static inline uint32_t cdiv64(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

static int scan_blkcnt(klsh_ctx* ctx, uint32_t* blkcnt, uint32_t nblk, uint32_t* total_dev) {
  k_scan_single<<<1, 1024, 0, ctx->stream>>>(blkcnt, nblk, total_dev);
  KLAUNCH(ctx);
  return KLSH_OK;
}

int launch_iota(klsh_ctx* ctx, uint32_t* out, uint64_t n, uint32_t base) {
  if (!n) return KLSH_OK;
  k_iota<<<cdiv64(n, 256), 256, 0, ctx->stream>>>(out, n, base);
  KLAUNCH(ctx);
  return KLSH_OK;
}

int launch_fill_tail(klsh_ctx* ctx, uint32_t* seg, uint64_t from, uint64_t to) {
  if (to <= from) return KLSH_OK;
  k_fill<<<cdiv64(to - from, 256), 256, 0, ctx->stream>>>(seg, from, to, KLSH_SENTINEL);
  KLAUNCH(ctx);
  return KLSH_OK;
}

int launch_init_meta(klsh_ctx* ctx, uint64_t n) {
  // next[] = -1 for n member slots
  if (!n) return KLSH_OK;
  k_fill_i32<<<cdiv64(n, 256), 256, 0, ctx->stream>>>(ctx->cur.next.as<int32_t>(), n, -1);
  KLAUNCH(ctx);
  return KLSH_OK;
}

int launch_transform(klsh_ctx* ctx, const uint16_t* d_counts, const float* d_vk, uint64_t batch, uint64_t* kept_out) {
  PassScratch& s = ctx->top;
  uint32_t nblk = cdiv64(batch, kScanTile);
  KTRY(dev_reserve(ctx, s.blkcnt, sizeof(uint32_t) * (nblk + 1)));
  KTRY(dev_reserve(ctx, s.counters, sizeof(PassCounters)));
  uint32_t* blk = s.blkcnt.as<uint32_t>();
  PassCounters* dc = s.counters.as<PassCounters>();
  k_transform_count<<<nblk, 256, 0, ctx->stream>>>(d_counts, batch, ctx->D, blk);
  KLAUNCH(ctx);
  KTRY(scan_blkcnt(ctx, blk, nblk, &dc->n_out));
  k_transform_write<<<nblk, 256, 0, ctx->stream>>>(d_counts, ctx->lut.as<float>(), d_vk, batch, ctx->D, ctx->ld, blk,
                                                   ctx->cur.vals.as<float>(), ctx->cur.cnt(),
                                                   ctx->cur.head(), ctx->cur.tail(), 0);
  KLAUNCH(ctx);
  KCUDA(ctx, cudaMemcpyAsync(&ctx->h_counters->n_out, &dc->n_out, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *kept_out = ctx->h_counters->n_out;
  return KLSH_OK;
}

int launch_sign(klsh_ctx* ctx, const float* vals, int D, int ld, const uint32_t* rows, uint64_t n,
                const float* d_planes, int H, uint32_t* keys_out, uint32_t* rows_out, uint32_t key_or) {
  if (!n) return KLSH_OK;
  KTRY(dev_reserve(ctx, ctx->eps_counter, sizeof(unsigned long long) * 4));
  if (ld <= 64) {  // tcgen05 + tensor memory (sign_umma.cuh); H <= 32 by the ABI
    if (H > 32) return klsh_fail(ctx, KLSH_ERR_ARG, "more than 32 hyperplanes (%d)", H);
    auto fn = ld <= 32 ? sign_umma::k_sign_umma<32, sign_umma::kPipe> : sign_umma::k_sign_umma<64, sign_umma::kPipe>;
    const size_t smem = ld <= 32 ? sign_umma::smem_bytes<32>() : sign_umma::smem_bytes<64>();
    KCUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // CTAs per SM by shared memory, registers and tensor-memory columns (the occupancy API reports 1 for this kernel)
    const int per_sm = ld <= 32 ? sign_umma::ctas_per_sm<32, sign_umma::kPipe>() : sign_umma::ctas_per_sm<64, sign_umma::kPipe>();
    const uint64_t ntiles = (n + 127) / 128;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(ntiles, (uint64_t)ctx->sm_count * per_sm);
    fn<<<grid, sign_umma::kThreads, smem, ctx->stream>>>(vals, D, ld, rows, n, d_planes, H, keys_out, rows_out,
                                                         ctx->eps_counter.as<unsigned long long>(), key_or, nullptr);
    KLAUNCH(ctx);
    return KLSH_OK;
  }
  const int nch = (ld + 63) / 64;
  const size_t fixed = sizeof(float) * 32 + 16 * (size_t)nch * 8 * 4 * 32;
  const size_t per_warp = sizeof(float) * 2 * 32 * 68;
  int warps = 8;
  while (warps > 1 && fixed + per_warp * warps > (size_t)ctx->max_smem_optin) warps >>= 1;
  const size_t smem = fixed + per_warp * warps;
  if (smem > (size_t)ctx->max_smem_optin) return klsh_fail(ctx, KLSH_ERR_ARG, "dimension %d too large for the signing kernel", D);
  KCUDA(ctx, cudaFuncSetAttribute(k_sign_tc_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const uint64_t want = (n + (uint64_t)warps * 32 - 1) / ((uint64_t)warps * 32);
  const uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)ctx->sm_count);
  k_sign_tc_wide<<<grid, warps * 32, smem, ctx->stream>>>(vals, D, ld, rows, n, d_planes, H, keys_out, rows_out,
                                                          ctx->eps_counter.as<unsigned long long>(), key_or);
  KLAUNCH(ctx);
  return KLSH_OK;
}

int launch_sort_pairs(klsh_ctx* ctx, PassScratch& s, uint64_t n, int bits, uint32_t** keys_sorted,
                      uint32_t** rows_sorted) {
  uint32_t* ka = s.keys_a.as<uint32_t>();
  uint32_t* kb = s.keys_b.as<uint32_t>();
  uint32_t* ra = s.rows_a.as<uint32_t>();
  uint32_t* rb = s.rows_b.as<uint32_t>();
  if (n >= 2 && bits > 0) {
    uint32_t nblk = cdiv64(n, kRadixTile);
    KTRY(dev_reserve(ctx, s.hist, sizeof(uint32_t) * ((size_t)256 * nblk + 256)));
    uint32_t* hist = s.hist.as<uint32_t>();
    uint32_t* dtot = hist + (size_t)256 * nblk;
    for (int shift = 0; shift < bits; shift += 8) {
      k_radix_hist<<<nblk, kRadixWarps * 32, 0, ctx->stream>>>(ka, n, shift, hist, nblk);
      KLAUNCH(ctx);
      k_radix_scan<<<256, 256, 0, ctx->stream>>>(hist, nblk, dtot);
      KLAUNCH(ctx);
      k_radix_scatter<<<nblk, kRadixWarps * 32, 0, ctx->stream>>>(ka, ra, n, shift, hist, dtot, nblk, kb, rb);
      KLAUNCH(ctx);
      std::swap(ka, kb);
      std::swap(ra, rb);
    }
  }
  *keys_sorted = ka;
  *rows_sorted = ra;
  return KLSH_OK;
}

int launch_bounds(klsh_ctx* ctx, PassScratch& s, const uint32_t* keys_sorted, uint64_t n) {
  uint32_t nblk = cdiv64(n, kScanTile);
  KTRY(dev_reserve(ctx, s.blkcnt, sizeof(uint32_t) * (nblk + 1)));
  KTRY(dev_reserve(ctx, s.bstart, sizeof(uint32_t) * (n + 2)));
  KTRY(dev_reserve(ctx, s.list_small, sizeof(uint32_t) * (n / 2 + 2)));
  KTRY(dev_reserve(ctx, s.list_large, sizeof(uint32_t) * 3 * (n / (KLSH_SMALL_MAX + 1) + 2)));
  KTRY(dev_reserve(ctx, s.list_big, sizeof(uint32_t) * 3 * (n / KLSH_BIG + 2)));
  KTRY(dev_reserve(ctx, s.list_direct, sizeof(uint32_t) * 3 * ((size_t)ctx->max_direct + 2)));
  KTRY(dev_reserve(ctx, s.list_nested, sizeof(uint32_t) * (n / 2 + 2)));
  KTRY(dev_reserve(ctx, s.pos_nrm, sizeof(float) * (n + 2)));
  KTRY(dev_reserve(ctx, s.pos_h, (size_t)(ctx->ld <= 32 ? 64 : (ctx->ld <= 64 ? 128 : 2 * ((ctx->ld + 31) & ~31))) * (n + 2)));
  KTRY(dev_reserve(ctx, s.counters, sizeof(PassCounters)));
  PassCounters* dc = s.counters.as<PassCounters>();
  KCUDA(ctx, cudaMemsetAsync(dc, 0, sizeof(PassCounters), ctx->stream));
  uint32_t* blk = s.blkcnt.as<uint32_t>();
  k_heads_count<<<nblk, 256, 0, ctx->stream>>>(keys_sorted, n, blk);
  KLAUNCH(ctx);
  KTRY(scan_blkcnt(ctx, blk, nblk, &dc->n_buckets));
  k_heads_write<<<nblk, 256, 0, ctx->stream>>>(keys_sorted, n, blk, s.bstart.as<uint32_t>());
  KLAUNCH(ctx);
  return KLSH_OK;
}

// Size classes of the buckets in [b_lo, b_hi) (everything after n_buckets in the counters is reset).
int launch_classify(klsh_ctx* ctx, PassScratch& s, uint64_t n, int64_t nest_threshold, uint32_t b_lo, uint32_t b_hi) {
  PassCounters* dc = s.counters.as<PassCounters>();
  KCUDA(ctx, cudaMemsetAsync(&dc->n_small, 0, sizeof(PassCounters) - offsetof(PassCounters, n_small), ctx->stream));
  // bucket count is on the device; launch enough threads for the worst case (n buckets)
  const uint32_t grid = std::min<uint32_t>(cdiv64(n, 256), (uint32_t)ctx->sm_count * 8);
  const bool fallback = launch_merge_uses_fallback(ctx);
  if (!fallback) {
    k_size_hist<<<grid, 256, 0, ctx->stream>>>(s.bstart.as<uint32_t>(), n, b_lo, b_hi, (long long)nest_threshold, dc);
    KLAUNCH(ctx);
  }
  k_classify<<<grid, 256, 0, ctx->stream>>>(s.bstart.as<uint32_t>(), n, (long long)nest_threshold, b_lo, b_hi,
                                            fallback ? 0xFFFFFFFFu : ctx->direct_min, ctx->max_direct, dc,
                                            s.list_small.as<uint32_t>(), s.list_large.as<uint32_t>(), s.list_big.as<uint32_t>(),
                                            s.list_direct.as<uint32_t>(), s.list_nested.as<uint32_t>());
  KLAUNCH(ctx);
  if (!fallback) {
    k_order_direct<<<1, 1024, 0, ctx->stream>>>(s.bstart.as<uint32_t>(), dc, s.list_direct.as<uint32_t>());
    KLAUNCH(ctx);
  }
  return KLSH_OK;
}

// ================================================================================================
// Multi-GPU helpers: bucket range splits, update export / apply
// ================================================================================================
namespace {
// splits[r] = first bucket whose start offset is >= r*n/world (buckets are never split)
__global__ void k_find_splits(const uint32_t* __restrict__ bstart, uint32_t nb, uint64_t n, int world, uint32_t* splits) {
  const int r = threadIdx.x;
  if (r > world) return;
  if (r == world) { splits[r] = nb; return; }
  const uint64_t target = (uint64_t)r * n / (uint64_t)world;
  uint32_t lo = 0, hi = nb;
  while (lo < hi) {
    const uint32_t mid = lo + (hi - lo) / 2;
    if ((uint64_t)bstart[mid] < target) lo = mid + 1; else hi = mid;
  }
  splits[r] = lo;
}
__global__ void k_gather_mod(const float* __restrict__ vals, int ld, const MetaCol cnt,
                             const MetaCol head, const MetaCol tail,
                             const uint32_t* __restrict__ rows, uint32_t n, float* out_vals, int32_t* out_meta) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
  if (w >= n) return;
  const uint32_t r = rows[w];
  for (int d = lane; d < ld; d += 32) out_vals[(uint64_t)w * ld + d] = vals[(uint64_t)r * ld + d];
  if (lane == 0) {
    out_meta[3 * (uint64_t)w] = cnt[r];
    out_meta[3 * (uint64_t)w + 1] = head[r];
    out_meta[3 * (uint64_t)w + 2] = tail[r];
  }
}
__global__ void k_apply_mod(float* vals, int ld, MetaCol cnt, MetaCol head, MetaCol tail, const uint32_t* __restrict__ rows,
                            uint32_t n, const float* __restrict__ in_vals, const int32_t* __restrict__ in_meta) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
  if (w >= n) return;
  const uint32_t r = rows[w];
  for (int d = lane; d < ld; d += 32) vals[(uint64_t)r * ld + d] = in_vals[(uint64_t)w * ld + d];
  if (lane == 0) {
    cnt[r] = in_meta[3 * (uint64_t)w];
    head[r] = in_meta[3 * (uint64_t)w + 1];
    tail[r] = in_meta[3 * (uint64_t)w + 2];
  }
}
__global__ void k_apply_next(int32_t* next, const uint32_t* __restrict__ slots, const int32_t* __restrict__ v, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) next[slots[i]] = v[i];
}
}  // namespace

// ================================================================================================
// Member chains -> flat id order on the device (export).  A host walk of next[] is one dependent
// cache miss per id and cannot be split inside one chain (clusters grow to millions of members), so
// the chains are ranked by pointer jumping: after ceil(log2(longest chain)) rounds every member slot
// knows its distance to the chain's tail and the tail itself; the tail identifies the cluster, and
// the distance gives the member's position inside the cluster's id list.
// ================================================================================================
namespace {
__global__ void k_rank_init(const int32_t* __restrict__ next, uint32_t m, int32_t* succ, uint32_t* dist, uint32_t* tailof) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= m) return;
  succ[s] = next[s];
  dist[s] = 1u;
  tailof[s] = s;
}
__global__ void k_rank_step(const int32_t* __restrict__ succ_in, const uint32_t* __restrict__ dist_in,
                            const uint32_t* __restrict__ tail_in, uint32_t m, int32_t* succ_out, uint32_t* dist_out,
                            uint32_t* tail_out, uint32_t* changed) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= m) return;
  const int32_t q = succ_in[s];
  if (q >= 0) {
    dist_out[s] = dist_in[s] + dist_in[q];
    tail_out[s] = tail_in[q];
    succ_out[s] = succ_in[q];
    *changed = 1u;  // benign race: everybody writes the same value
  } else {
    dist_out[s] = dist_in[s];
    tail_out[s] = tail_in[s];
    succ_out[s] = -1;
  }
}
__global__ void k_rank_owner(const MetaCol tail, const uint32_t* __restrict__ alive, uint32_t n, int32_t* owner) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int32_t t = tail[alive[r]];
  if (t >= 0) owner[t] = (int32_t)r;
}
__global__ void k_rank_emit(const uint32_t* __restrict__ dist, const uint32_t* __restrict__ tailof,
                            const int32_t* __restrict__ owner, const uint32_t* __restrict__ offs /* n+1 */, uint32_t m,
                            uint32_t* slot_out) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= m) return;
  const int32_t r = owner[tailof[s]];
  if (r < 0) return;  // slot of a row that never entered the set (dropped by the keep filter)
  const uint32_t end = offs[r + 1];
  const uint32_t d = dist[s];
  if (d <= end - offs[r]) slot_out[end - d] = s;
}
__global__ void k_rank_slot_row(const uint32_t* __restrict__ tailof, const int32_t* __restrict__ owner, uint32_t m, int32_t* slot_row) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < m) slot_row[s] = owner[tailof[s]];
}
}  // namespace

// slot_out[0 .. total_ids): member slots in output order (cluster after cluster, chain order inside).
// d_offs: exclusive prefix of the survivors' member counts (n+1 entries, uint32) on the device.
// With slot_row_out instead (d_offs, slot_out NULL): slot_row_out[s] = position in the working set of the row
// whose member chain holds slot s, -1 for slots of rows that are not in the set.
int launch_rank_chains(klsh_ctx* ctx, uint64_t n, const uint32_t* d_offs, uint32_t* slot_out, int32_t* slot_row_out) {
  const uint32_t m = (uint32_t)ctx->n_slots;
  if (!m || !n) return KLSH_OK;
  KTRY(dev_reserve(ctx, ctx->rank_buf, sizeof(uint32_t) * ((size_t)m * 7 + 64)));
  uint32_t* base = ctx->rank_buf.as<uint32_t>();
  int32_t* succ[2] = {reinterpret_cast<int32_t*>(base), reinterpret_cast<int32_t*>(base + (size_t)m)};
  uint32_t* dist[2] = {base + (size_t)2 * m, base + (size_t)3 * m};
  uint32_t* tl[2] = {base + (size_t)4 * m, base + (size_t)5 * m};
  int32_t* owner = reinterpret_cast<int32_t*>(base + (size_t)6 * m);
  uint32_t* changed = base + (size_t)7 * m;
  const uint32_t grid = cdiv64(m, 256);
  k_rank_init<<<grid, 256, 0, ctx->stream>>>(ctx->cur.next.as<int32_t>(), m, succ[0], dist[0], tl[0]);
  KLAUNCH(ctx);
  KCUDA(ctx, cudaMemsetAsync(owner, 0xFF, sizeof(int32_t) * m, ctx->stream));
  k_rank_owner<<<cdiv64(n, 256), 256, 0, ctx->stream>>>(ctx->cur.tail(), ctx->cur.alive.as<uint32_t>(), (uint32_t)n, owner);
  KLAUNCH(ctx);
  int cur = 0;
  for (int round = 0; round < 34; ++round) {
    KCUDA(ctx, cudaMemsetAsync(changed, 0, sizeof(uint32_t), ctx->stream));
    k_rank_step<<<grid, 256, 0, ctx->stream>>>(succ[cur], dist[cur], tl[cur], m, succ[cur ^ 1], dist[cur ^ 1], tl[cur ^ 1], changed);
    KLAUNCH(ctx);
    cur ^= 1;
    if ((round & 3) == 3) {  // look at the flag every fourth round only (each look is a host sync)
      uint32_t h = 0;
      KCUDA(ctx, cudaMemcpyAsync(&h, changed, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
      KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (!h) break;
    }
  }
  if (slot_row_out) {
    k_rank_slot_row<<<grid, 256, 0, ctx->stream>>>(tl[cur], owner, m, slot_row_out);
    KLAUNCH(ctx);
    return KLSH_OK;
  }
  k_rank_emit<<<grid, 256, 0, ctx->stream>>>(dist[cur], tl[cur], owner, d_offs, m, slot_out);
  KLAUNCH(ctx);
  return KLSH_OK;
}

int launch_find_splits(klsh_ctx* ctx, PassScratch& s, uint32_t nb, uint64_t n, int world, uint32_t* d_splits) {
  k_find_splits<<<1, 64, 0, ctx->stream>>>(s.bstart.as<uint32_t>(), nb, n, world, d_splits);
  KLAUNCH(ctx);
  return KLSH_OK;
}
int launch_gather_mod(klsh_ctx* ctx, const uint32_t* rows, uint32_t n, float* out_vals, int32_t* out_meta) {
  if (!n) return KLSH_OK;
  k_gather_mod<<<cdiv64((uint64_t)n * 32, 256), 256, 0, ctx->stream>>>(ctx->cur.vals.as<float>(), ctx->ld, ctx->cur.cnt(),
                                                                     ctx->cur.head(), ctx->cur.tail(), rows, n,
                                                                     out_vals, out_meta);
  KLAUNCH(ctx);
  return KLSH_OK;
}
int launch_apply_mod(klsh_ctx* ctx, const uint32_t* rows, uint32_t n, const float* in_vals, const int32_t* in_meta,
                     const uint32_t* slots, const int32_t* nvals, uint32_t n_next) {
  if (n) {
    k_apply_mod<<<cdiv64((uint64_t)n * 32, 256), 256, 0, ctx->stream>>>(ctx->cur.vals.as<float>(), ctx->ld, ctx->cur.cnt(),
                                                                      ctx->cur.head(), ctx->cur.tail(), rows, n,
                                                                      in_vals, in_meta);
    KLAUNCH(ctx);
  }
  if (n_next) {
    k_apply_next<<<cdiv64(n_next, 256), 256, 0, ctx->stream>>>(ctx->cur.next.as<int32_t>(), slots, nvals, n_next);
    KLAUNCH(ctx);
  }
  return KLSH_OK;
}

static int large_smem_config(klsh_ctx* ctx, int ld, int* rep_cap, size_t* smem) {
  size_t budget = (size_t)ctx->max_smem_optin - 1024;
  size_t per_rep = sizeof(float) * (size_t)(ld + 1 + 1);
  size_t fixed = sizeof(float) * (size_t)ld;
  int cap = (int)((budget - fixed) / per_rep);
  if (cap < 1) return klsh_fail(ctx, KLSH_ERR_ARG, "dimension %d too large for the merge kernel", ld);
  *rep_cap = cap;
  *smem = fixed + per_rep * (size_t)cap;
  // function attributes are per device: set it on every call (cheap), never cache it process-wide
  KCUDA(ctx, cudaFuncSetAttribute(k_merge_large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
  return KLSH_OK;
}

// The windowed merge stages its window in shared memory (about 768 bytes per float of row width): rows
// wider than that allows (D > ~280) and KLSH_MERGE_V1=1 go to the block-per-bucket kernel, which sizes
// its representative cache from whatever shared memory there is and is exact for any D (and slow).
bool launch_merge_uses_fallback(const klsh_ctx* ctx) {
  return ctx->merge_v1 || merge_window_smem_bytes(ctx->ld) > (size_t)ctx->max_smem_optin;
}

static int launch_merge_fallback(klsh_ctx* ctx, PassScratch& s, uint32_t* rows_sorted, float threshold, const PassCounters& c) {
  int rep_cap;
  size_t smem;
  KTRY(large_smem_config(ctx, ctx->ld, &rep_cap, &smem));
  const uint32_t n_large = c.n_large + c.n_big;
  uint32_t grid = std::min<uint32_t>(n_large, (uint32_t)ctx->sm_count);
  uint64_t spill_stride = 0;
  if ((int64_t)c.bucket_max > rep_cap) spill_stride = c.bucket_max - rep_cap;
  KTRY(dev_reserve(ctx, ctx->io_b, sizeof(float) * (spill_stride * grid + 1)));
  k_merge_large<<<grid, kLargeThreads, smem, ctx->stream>>>(ctx->cur.vals.as<float>(), ctx->D, ctx->ld, ctx->cur.cnt(),
                                                            ctx->cur.head(), ctx->cur.tail(),
                                                            ctx->cur.next.as<int32_t>(), rows_sorted, s.bstart.as<uint32_t>(),
                                                            s.list_large.as<uint32_t>(), s.list_big.as<uint32_t>(),
                                                            s.counters.as<PassCounters>(), threshold, rep_cap, ctx->io_b.as<float>(),
                                                            spill_stride, -1);
  KLAUNCH(ctx);
  return KLSH_OK;
}

int launch_merge(klsh_ctx* ctx, PassScratch& s, uint32_t* rows_sorted, float threshold, const PassCounters& c) {
  const int D = ctx->D, ld = ctx->ld;
  float* vals = ctx->cur.vals.as<float>();
  const MetaCol cnt = ctx->cur.cnt(), head = ctx->cur.head(), tail = ctx->cur.tail();
  int32_t* next = ctx->cur.next.as<int32_t>();
  PassCounters* dc = s.counters.as<PassCounters>();
  const uint32_t n_small = c.n_small, n_large = c.n_large + c.n_big, n_direct = c.n_direct;
  const bool v1 = launch_merge_uses_fallback(ctx);
  const bool pool = launch_merge_uses_pool(ctx);
  // The few largest buckets are long sequential window chains: they start right away on cluster
  // teams on the second stream and run beside the small buckets and the single-CTA stage.
  bool forked = false;
  static thread_local cudaEvent_t tl[3] = {nullptr, nullptr, nullptr};  // KLSH_TIMELINE: start, end of the main pipeline, end of the direct one
  if (ctx->timeline) {
    for (auto& e : tl)
      if (!e) cudaEventCreate(&e);
    cudaEventRecord(tl[0], ctx->stream);
  }
  if (n_direct && !v1 && !pool) {
    KTRY(launch_pool_reset(ctx));
    KCUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    KCUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    forked = true;
    int rc = launch_merge_direct(ctx, s, rows_sorted, threshold, n_direct, c.bucket_max);
    if (rc != KLSH_OK) {
      cudaStreamSynchronize(ctx->stream2);
      return rc;
    }
    KCUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->stream2));
    if (ctx->timeline) cudaEventRecord(tl[2], ctx->stream2);
  }
  int rc = KLSH_OK;
  do {
    if (n_small) {
      size_t per_warp = sizeof(float) * (size_t)KLSH_SMALL_MAX * (ld + 4);
      int wpb = 4;
      while (wpb > 1 && per_warp * wpb > (size_t)ctx->max_smem_optin - 1024) wpb >>= 1;
      size_t smem = per_warp * wpb;
      if (smem > (size_t)ctx->max_smem_optin) {
        rc = klsh_fail(ctx, KLSH_ERR_ARG, "dimension %d too large", D);
        break;
      }
      if (smem > 48 * 1024 &&  // per device, so not cached process-wide
          cudaFuncSetAttribute(k_merge_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        rc = klsh_fail(ctx, KLSH_ERR_CUDA, "cudaFuncSetAttribute(k_merge_small) failed");
        break;
      }
      uint32_t grid = std::min<uint32_t>((n_small + wpb - 1) / wpb, (uint32_t)ctx->sm_count * 32);
      k_merge_small<<<grid, wpb * 32, smem, ctx->stream>>>(vals, D, ld, cnt, head, tail, next, rows_sorted,
                                                           s.bstart.as<uint32_t>(), s.list_small.as<uint32_t>(), dc,
                                                           threshold, ctx->mg);
      ctx->launches++;
      if (cudaGetLastError() != cudaSuccess) {
        rc = klsh_fail(ctx, KLSH_ERR_CUDA, "k_merge_small launch failed");
        break;
      }
    }
    if (!(n_large + ((v1 || pool) ? n_direct : 0u))) break;
    if (pool) {
      rc = launch_merge_pool(ctx, s, rows_sorted, threshold);
      break;
    }
    if (!v1) {
      rc = launch_merge_window(ctx, s, rows_sorted, threshold, n_large, c.n_direct ? std::min(c.bucket_max, c.direct_thr - 1u) : c.bucket_max);
      break;
    }
    rc = launch_merge_fallback(ctx, s, rows_sorted, threshold, c);
  } while (0);
  if (ctx->timeline) cudaEventRecord(tl[1], ctx->stream);
  if (forked) {
    cudaError_t e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
    if (e != cudaSuccess && rc == KLSH_OK) rc = klsh_fail(ctx, KLSH_ERR_CUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(e));
    if (rc != KLSH_OK) cudaStreamSynchronize(ctx->stream2);
  }
  if (ctx->timeline && rc == KLSH_OK) {
    cudaStreamSynchronize(ctx->stream);
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, tl[0], tl[1]);
    if (forked) cudaEventElapsedTime(&b, tl[0], tl[2]);
    fprintf(stderr, "[klsh] timeline: main pipeline (small + single-CTA + escalations) ended at %.3f ms, direct pipeline (%u buckets) at %.3f ms; largest bucket %u rows\n",
            a, n_direct, b, c.bucket_max);
  }
  return rc;
}

// klsh_p_cluster: rows_sorted[0..n) is ONE bucket
int launch_merge_one(klsh_ctx* ctx, PassScratch& s, uint32_t* rows_sorted, uint64_t n, float threshold) {
  if (n < 2) return KLSH_OK;
  KTRY(dev_reserve(ctx, s.counters, sizeof(PassCounters)));
  if (!launch_merge_uses_fallback(ctx)) {
    KTRY(dev_reserve(ctx, s.pos_nrm, sizeof(float) * (n + 2)));
    KTRY(dev_reserve(ctx, s.pos_h, (size_t)(ctx->ld <= 32 ? 64 : (ctx->ld <= 64 ? 128 : 2 * ((ctx->ld + 31) & ~31))) * (n + 2)));
    KTRY(dev_reserve(ctx, s.bstart, sizeof(uint32_t) * 4));
    KTRY(dev_reserve(ctx, s.list_big, sizeof(uint32_t) * 4));
    KTRY(dev_reserve(ctx, s.list_large, sizeof(uint32_t) * 4));
    PassCounters hc;
    std::memset(&hc, 0, sizeof hc);
    hc.n_buckets = 1;
    hc.n_big = 1;
    hc.bucket_max = (uint32_t)n;
    const uint32_t bs[2] = {0u, (uint32_t)n};
    const uint32_t item[3] = {0u, 0u, 0u};
    KCUDA(ctx, cudaMemcpyAsync(s.counters.p, &hc, sizeof hc, cudaMemcpyHostToDevice, ctx->stream));
    KCUDA(ctx, cudaMemcpyAsync(s.bstart.p, bs, sizeof bs, cudaMemcpyHostToDevice, ctx->stream));
    KCUDA(ctx, cudaMemcpyAsync(s.list_big.p, item, sizeof item, cudaMemcpyHostToDevice, ctx->stream));
    KCUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the sources above are on the stack
    return launch_merge_window(ctx, s, rows_sorted, threshold, 1, (uint32_t)n);
  }
  int rep_cap;
  size_t smem;
  KTRY(large_smem_config(ctx, ctx->ld, &rep_cap, &smem));
  uint64_t spill = (n > (uint64_t)rep_cap) ? n - rep_cap : 0;
  KTRY(dev_reserve(ctx, ctx->io_b, sizeof(float) * (spill + 1)));
  k_merge_large<<<1, kLargeThreads, smem, ctx->stream>>>(
      ctx->cur.vals.as<float>(), ctx->D, ctx->ld, ctx->cur.cnt(), ctx->cur.head(),
      ctx->cur.tail(), ctx->cur.next.as<int32_t>(), rows_sorted, nullptr, nullptr, nullptr,
      s.counters.as<PassCounters>(), threshold, rep_cap, ctx->io_b.as<float>(), spill, (int)n);
  KLAUNCH(ctx);
  return KLSH_OK;
}

int launch_compact(klsh_ctx* ctx, PassScratch& s, const uint32_t* rows_sorted, uint64_t n, uint32_t* out) {
  if (!n) return KLSH_OK;
  uint32_t nblk = cdiv64(n, kScanTile);
  KTRY(dev_reserve(ctx, s.blkcnt, sizeof(uint32_t) * (nblk + 1)));
  PassCounters* dc = s.counters.as<PassCounters>();
  uint32_t* blk = s.blkcnt.as<uint32_t>();
  k_alive_count<<<nblk, 256, 0, ctx->stream>>>(rows_sorted, n, blk);
  KLAUNCH(ctx);
  KTRY(scan_blkcnt(ctx, blk, nblk, &dc->n_out));
  k_alive_write<<<nblk, 256, 0, ctx->stream>>>(rows_sorted, n, blk, out);
  KLAUNCH(ctx);
  return KLSH_OK;
}

int launch_gather_rows(klsh_ctx* ctx, const uint32_t* rows, uint64_t n, float* out_vals, int32_t* out_cnt,
                       int32_t* out_head) {
  if (!n) return KLSH_OK;
  k_gather_rows<<<cdiv64(n * 32, 256), 256, 0, ctx->stream>>>(ctx->cur.vals.as<float>(), ctx->D, ctx->ld,
                                                             ctx->cur.cnt(), ctx->cur.head(),
                                                             rows, n, out_vals, out_cnt, out_head);
  KLAUNCH(ctx);
  return KLSH_OK;
}

int launch_cosine_pairs(klsh_ctx* ctx, const float* left, const float* right, uint64_t n, int ld, float* out) {
  if (!n) return KLSH_OK;
  k_cosine_pairs<<<cdiv64(n, 128), 128, 0, ctx->stream>>>(left, right, n, ld, out);
  KLAUNCH(ctx);
  return KLSH_OK;
}

int launch_consensus(klsh_ctx* ctx, const float* cur, int c1, const float* cand, int c2, int D, float* out) {
  k_consensus<<<cdiv64((uint64_t)D, 128), 128, 0, ctx->stream>>>(cur, c1, cand, c2, D, out);
  KLAUNCH(ctx);
  return KLSH_OK;
}

int launch_sum_counts(klsh_ctx* ctx, const uint32_t* rows, uint64_t n, unsigned long long* total_dev) {
  KCUDA(ctx, cudaMemsetAsync(total_dev, 0, sizeof(unsigned long long), ctx->stream));
  if (!n) return KLSH_OK;
  const uint32_t grid = std::min<uint32_t>(cdiv64(n, 256), (uint32_t)ctx->sm_count * 8);
  k_sum_counts<<<grid, 256, 0, ctx->stream>>>(ctx->cur.cnt(), rows, n, total_dev);
  KLAUNCH(ctx);
  return KLSH_OK;
}
