// Signing on the 5th-generation tensor cores: tcgen05.mma kind::tf32, row operand and accumulators in tensor memory.
//
// Same arithmetic contract as the mma.sync kernel for wide rows (k_sign_tc_wide, kernels.cu): every (row, plane) sum
// is evaluated as a 3xTF32 product (lo*hi + hi*lo + hi*hi, fp32 accumulate) against the unit-length planes; a sum
// outside the margin c*|x| has the sign of the reference's mul-then-add chain (hash/lshash.cc:44-51), a sum inside
// it is re-evaluated with that chain.  Who does the work:
//   * a CTA of 128 threads owns tiles of 128 rows; thread m owns row m = tensor-memory lane m.  The rows are gathered
//     with cp.async into a two-stage ring of raw tiles (16-byte chunks XOR-swizzled by row so that a thread walks its
//     row without bank conflicts); a stage is refilled as soon as its rows have been split, so two tiles are in
//     flight while one is multiplied (the rare exact re-evaluation reads its row from global memory again);
//   * thread m splits its row once (hi = x & 0xffffe000, lo = x - hi: one LOP and one FADD per element) and writes both
//     halves straight from registers into tensor-memory lane m (tcgen05.st, column k = element k of the row);
//   * ONE thread issues 3*KW/8 tcgen05.mma (M = 128, N = 32, K = 8; the "TS" form: A from tensor memory, B from shared
//     memory) per tile; the planes sit in shared memory pre-split in the K-major, no-swizzle canonical layout — chunk c
//     (4 columns) of plane h at c*512 + h*16, i.e. core matrices of 8 planes x 16 bytes, 128 bytes apart along N (SBO)
//     and 512 bytes apart along K (LBO); the 128 x 32 sums land in 32 columns of tensor memory;
//   * tcgen05.commit signals an mbarrier; every thread reads ITS row's 32 sums with one tcgen05.ld.32x32b.x32 and packs
//     the key bits — no fragment shuffles, no per-lane transposition, no operand tiles in shared memory.
// A CTA needs the planes (KW*256 bytes), the gather ring and 2*KW + 64 columns of tensor memory (two allocations: A hi
// and lo, two buffers of sums), so four CTAs fit an SM at KW = 32 (two at KW = 64) and cover each other's gather /
// split / products / epilogue latencies.  Per 32 rows this is about 600 warp instructions against 1250 for the mma.sync kernel.
// Measured (tools/microbench/sign_umma_test.cu, profiles/r02d_sign_umma_microbench.txt): 8 M x 32 rows 4.24 TB/s
// algorithmic (65 % of the measured HBM copy peak), 4 M x 64 3.63 TB/s.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sign_umma {

constexpr int kThreads = 128;  // = rows per tile = tensor-memory lanes
constexpr int kStages = 2;
constexpr bool kPipe = true;   // which variant the library launches (see PIPE below)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t tf32_rna_bits(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (sm_100); offsets in 16-byte units
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at bit 17,
// M >> 4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// Tensor memory per CTA: A hi and lo, KW columns each (one allocation of 2*KW), and the 32 columns of sums (a second
// allocation): 96 columns at KW = 32, 160 at KW = 64, of the SM's 512.
template <int KW>
constexpr size_t smem_bytes() {
  return (size_t)2 * KW * 32 * 4 + (size_t)kStages * 128 * KW * 4 + 128;
}
// PIPE: two accumulator buffers, so the key packing of tile t-1 runs while the tensor core multiplies tile t.
template <int KW, bool PIPE>
constexpr int tmem_cols() { return 2 * KW + (PIPE ? 64 : 32); }
template <int KW, bool PIPE>
constexpr int ctas_per_sm() { return KW == 32 ? (PIPE ? 4 : 5) : 2; }  // registers, shared and tensor memory

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(kIdesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

template <int KW, bool PIPE>
__global__ void __launch_bounds__(kThreads, ctas_per_sm<KW, PIPE>())
k_sign_umma(const float* __restrict__ vals, int D, int ld, const uint32_t* __restrict__ rows, uint64_t n,
               const float* __restrict__ planes, int H, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ rows_out,
               unsigned long long* eps_rows, uint32_t key_or, unsigned long long* prof) {
  constexpr int KS = KW / 8;
  extern __shared__ __align__(16) uint8_t smem_raw_[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem[2];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw_ + 127) & ~(uintptr_t)127);
  float* b_hi = reinterpret_cast<float*>(base);  // [KW/4][32][4]
  float* b_lo = b_hi + KW * 32;
  float* raw = b_lo + KW * 32;                   // [kStages][128][KW], chunk c of row m at (c ^ (m & 7))
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < kStages * 128 * KW; i += kThreads) raw[i] = 0.f;
  if (tid < 32) {
    const int h = tid;
    float m = 0.f;
    if (h < H)
      for (int i = 0; i < D; ++i) {
        const float w = __ldg(planes + h * ld + i);
        m = __fmaf_rn(w, w, m);
      }
    const float nrm = sqrtf(m);
    const float inv = (h < H && nrm > 1e-30f && nrm < 1e30f) ? 1.f / nrm : 0.f;
    for (int k = 0; k < KW; ++k) {
      float w = (h < H && k < ld && inv != 0.f) ? __ldg(planes + h * ld + k) * inv : 0.f;
      const uint32_t hi = tf32_rna_bits(w);
      const uint32_t lo = tf32_rna_bits(w - __uint_as_float(hi));
      const int at = ((k >> 2) * 32 + h) * 4 + (k & 3);
      b_hi[at] = __uint_as_float(hi);
      b_lo[at] = __uint_as_float(lo);
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem[0])), "r"((uint32_t)(2 * KW)) : "memory");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem[1])), "r"(PIPE ? 64u : 32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the planes were written with ordinary stores
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm_a = s_tmem[0], tm_d = s_tmem[1];
  const uint32_t bar = smem_u32(&s_bar);
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

  const float kc = ((float)D + 34.f + 30.f * (float)((D + 7) / 8)) * 5.9604645e-8f;
  const uint32_t hmask = H >= 32 ? 0xFFFFFFFFu : ((1u << H) - 1u);
  const int vpr = ld >> 2;
  const uint64_t ntiles = (n + 127) / 128;
  uint32_t my_eps = 0;
  // (row, chunk) of the first 16-byte piece this lane copies, and the step to its next one (32 pieces on)
  const int rr0 = lane / vpr, cc0 = lane - rr0 * vpr, rr_step = 32 / vpr, cc_step = 32 - rr_step * vpr;

  auto issue = [&](uint64_t tile, int st) -> uint32_t {
    uint32_t r = 0u;
    if (tile < ntiles) {
      const uint64_t t0 = tile * 128 + (uint64_t)warp * 32;
      if (t0 < n) {
        const uint64_t t = t0 + lane;
        r = (t < n) ? (rows ? rows[t] : (uint32_t)t) : 0u;
        const int nrow = (int)min((uint64_t)32, n - t0);
        const int total = nrow * vpr;
        float* dst = raw + ((size_t)st * 128 + warp * 32) * KW;
        if (vpr == KW / 4 && nrow == 32) {  // full-width rows, full tile: (row, chunk) of piece i are compile-time strides
          constexpr int NCH = KW / 4, RPI = 32 / NCH;  // chunks per row, rows covered per iteration of the warp
          const int rr_l = lane / NCH, cc_l = lane % NCH;
#pragma unroll
          for (int i = 0; i < NCH; ++i) {
            const int rr = i * RPI + rr_l;
            const uint32_t ri = __shfl_sync(0xffffffffu, r, rr);
            cp_async16(dst + rr * KW + ((cc_l ^ (rr & 7)) << 2), vals + (uint64_t)ri * ld + cc_l * 4);
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          return r;
        }
        int rr = rr0, cc = cc0;
        for (int v0 = 0; v0 < total; v0 += 32) {
          const uint32_t ri = __shfl_sync(0xffffffffu, r, min(rr, 31));
          if (v0 + lane < total) cp_async16(dst + rr * KW + ((cc ^ (rr & 7)) << 2), vals + (uint64_t)ri * ld + cc * 4);
          rr += rr_step;
          cc += cc_step;
          if (cc >= vpr) {
            cc -= vpr;
            ++rr;
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    return r;
  };

  // this thread's 32 sums of one tile (accumulator columns at d_addr) -> key; r_e: the row's index, xn_e: its norm
  auto epilogue = [&](uint64_t tile_e, uint32_t r_e, float xn_e, uint32_t d_addr) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t c[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]), "=r"(c[8]), "=r"(c[9]),
          "=r"(c[10]), "=r"(c[11]), "=r"(c[12]), "=r"(c[13]), "=r"(c[14]), "=r"(c[15]), "=r"(c[16]), "=r"(c[17]), "=r"(c[18]),
          "=r"(c[19]), "=r"(c[20]), "=r"(c[21]), "=r"(c[22]), "=r"(c[23]), "=r"(c[24]), "=r"(c[25]), "=r"(c[26]), "=r"(c[27]),
          "=r"(c[28]), "=r"(c[29]), "=r"(c[30]), "=r"(c[31])
        : "r"(d_addr + lane_sel)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // Sign bits: a sum outside the margin is finite and non-zero, so its sign bit IS the answer; every other sum
    // (inside the margin, zero, NaN, or any sum of a row whose |x| is not finite) is flagged below and recomputed.
    const float thr = kc * xn_e;
    uint32_t sgn = 0u;  // after the loop: bit 31-h = sign bit of plane h's sum (one funnel shift per plane)
    float mn = 3.4e38f;
#pragma unroll
    for (int h = 0; h < 32; ++h) {
      sgn = __funnelshift_l(c[h], sgn, 1);  // (sgn << 1) | (c[h] >> 31)
      if (h < H) mn = fminf(mn, fabsf(__uint_as_float(c[h])));
    }
    uint32_t pos = ~sgn;  // bit 31-h set <=> plane h non-negative: already in key order (plane 0 most significant)
    uint32_t slow = 0u;
    if (!(mn > thr && thr <= 3.0e38f)) {  // rare; also taken for NaN or infinite |x|
      uint32_t flag = 0u;
#pragma unroll
      for (int h = 0; h < 32; ++h) {
        const float a = fabsf(__uint_as_float(c[h]));
        flag |= ((a > thr && a <= 3.0e38f) ? 0u : 1u) << h;
      }
      flag &= hmask;
      slow = flag ? 1u : 0u;
      while (flag) {  // reference arithmetic: sum = fl(sum + fl(w_i * x_i)), i ascending (hash/lshash.cc:44-51)
        const int h = __ffs(flag) - 1;
        flag &= flag - 1;
        const float* w = planes + h * ld;
        const float* x = vals + (uint64_t)r_e * ld;
        float sum = 0.f;
        for (int j = 0; j < D; ++j) sum = __fadd_rn(sum, __fmul_rn(__ldg(w + j), __ldg(x + j)));
        const uint32_t bit = 0x80000000u >> h;
        pos = (pos & ~bit) | (sum >= 0.f ? bit : 0u);
      }
    }
    const uint32_t key = H ? (pos >> (32 - H)) : 0u;  // planes 0..H-1 -> key bits H-1..0; planes past H drop out
    const uint64_t t = tile_e * 128 + tid;
    if (t < n) {
      keys_out[t] = key | key_or;
      rows_out[t] = r_e;
      my_eps += slow;
    }
  };

  uint64_t tile = blockIdx.x;
  uint32_t r_q[2];
  r_q[0] = issue(tile, 0);
  r_q[1] = issue(tile + gridDim.x, 1);
  uint32_t it = 0;
  uint64_t tile_prev = 0;
  uint32_t r_prev = 0u;
  float xn_prev = 0.f;
  for (; tile < ntiles; tile += gridDim.x, ++it) {
    const int st = (int)(it & 1u);
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    if (prof) c0 = clock64();
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    if (PIPE && it > 0) mbar_wait(bar, (it - 1u) & 1u);  // the products of the previous tile are done: A may be overwritten
    if (prof) c1 = clock64();
    // ---- this thread's row: norm, split, both halves into tensor-memory lane `tid` ----------------------
    const float* myrow = raw + ((size_t)st * 128 + tid) * KW;
    float xx = 0.f;
#pragma unroll
    for (int grp = 0; grp < KW / 16; ++grp) {  // 16 columns at a time keeps the live registers low
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 x = *reinterpret_cast<const float4*>(myrow + (((grp * 4 + q) ^ (tid & 7)) << 2));
        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          xx = __fmaf_rn(xv[e], xv[e], xx);
          hi[q * 4 + e] = __float_as_uint(xv[e]) & 0xFFFFE000u;
          lo[q * 4 + e] = __float_as_uint(xv[e] - __uint_as_float(hi[q * 4 + e]));
        }
      }
      tmem_st16(tm_a + lane_sel + (uint32_t)(grp * 16), hi);
      tmem_st16(tm_a + lane_sel + (uint32_t)(KW + grp * 16), lo);
    }
    const float xn = sqrtf(xx);
    __syncwarp();  // the warp's 32 rows of this stage have been read: refill it with the tile after next
    const uint32_t r_new = issue(tile + 2ull * gridDim.x, st);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (prof) c2 = clock64();
    const uint32_t d_cur = tm_d + (PIPE ? (it & 1u) * 32u : 0u);
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint64_t dbh = make_desc(bh + ks * 2 * 512, 512, 128), dbl = make_desc(bl + ks * 2 * 512, 512, 128);
        mma_tf32_ts(d_cur, tm_a + (uint32_t)(KW + ks * 8), dbh, ks > 0 ? 1u : 0u);  // lo * hi
        mma_tf32_ts(d_cur, tm_a + (uint32_t)(ks * 8), dbl, 1u);                      // hi * lo
        mma_tf32_ts(d_cur, tm_a + (uint32_t)(ks * 8), dbh, 1u);                      // hi * hi
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    if (PIPE) {
      if (prof) c3 = clock64();
      if (it > 0) epilogue(tile_prev, r_prev, xn_prev, tm_d + ((it - 1u) & 1u) * 32u);  // while this tile is multiplied
      tile_prev = tile;
      r_prev = r_q[0];
      xn_prev = xn;
    } else {
      mbar_wait(bar, it & 1u);
      if (prof) c3 = clock64();
      epilogue(tile, r_q[0], xn, d_cur);
    }
    r_q[0] = r_q[1];
    r_q[1] = r_new;
    if (prof && tid == 0) {  // cycles: waiting (gather, previous products), split + barrier, products, epilogue; tiles
      c4 = clock64();
      atomicAdd(prof + 0, (unsigned long long)(c1 - c0));
      atomicAdd(prof + 1, (unsigned long long)(c2 - c1));
      atomicAdd(prof + 2, (unsigned long long)(c3 - c2));
      atomicAdd(prof + 3, (unsigned long long)(c4 - c3));
      atomicAdd(prof + 4, 1ull);
    }
  }
  if (PIPE && it > 0) {
    mbar_wait(bar, (it - 1u) & 1u);
    epilogue(tile_prev, r_prev, xn_prev, tm_d + ((it - 1u) & 1u) * 32u);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (eps_rows) {
    const uint32_t tot = __reduce_add_sync(0xffffffffu, my_eps);
    if (lane == 0 && tot) atomicAdd(eps_rows, (unsigned long long)tot);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_a), "r"((uint32_t)(2 * KW)) : "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_d), "r"(PIPE ? 64u : 32u) : "memory");
  }
}

}  // namespace sign_umma
