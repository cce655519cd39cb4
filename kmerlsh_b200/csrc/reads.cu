// Read extraction votes: which reads carry enough differential k-mers.
//
// Reference: IOFQ::CheckRead (io/ioFastQ.cc:5-75) slides a Kmer over every read of at least k+10 characters,
// takes the canonical form rep = (km < tw) ? km : tw (tw = km.twin(), operator< = memcmp over the MAX_K/4 = 8
// bytes, kmer/Kmer.cc:98-100, :160-185), counts the reps found in the unordered_set of differential k-mers and
// marks the read when float(count) / float(len - k + 1) > kmer_vote.  The FASTQ parsing and the writing of the
// marked reads (IOFQ::ReadExtract, :77-158) stay host code.
//
// Here the set is an open-addressing hash table of the 8-byte records in device memory (klsh_kmer_set_load; the
// records are what klsh_select_kmers emits and kmer_set.hex holds) and one thread walks one read with the k-mer
// and its reverse complement rolling in two 64-bit registers.  The record of a k-mer is the little-endian
// image of v = sum base_i << 2i (base i in byte i/4 at bit 2*(i%4), kmer/Kmer.cc:131-150; characters other than
// A, C, G, T count as A there and here), so memcmp order is the order of the byte-swapped words.
// Everything is integer work except the final ratio, one IEEE float division: results are bit-identical.
#include <algorithm>

#include "klsh_internal.cuh"

namespace {

#define RLAUNCH(ctx)                                                                                     \
  do {                                                                                                   \
    (ctx)->launches++;                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                                \
    if (e__ != cudaSuccess)                                                                              \
      return klsh_fail((ctx), KLSH_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                       __FILE__, __LINE__);                                                              \
  } while (0)

constexpr unsigned long long kEmpty = 0xFFFFFFFFFFFFFFFFull;  // also a valid k-mer (32 x T): kept in a side flag

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

__global__ void k_set_insert(const unsigned long long* __restrict__ rec, uint64_t n, unsigned long long* table, uint64_t mask,
                             uint32_t* has_ones) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = rec[i];
  if (key == kEmpty) {
    *has_ones = 1u;
    return;
  }
  uint64_t slot = mix64(key) & mask;
  for (;;) {
    const unsigned long long old = atomicCAS(table + slot, kEmpty, key);
    if (old == kEmpty || old == key) return;
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ bool set_contains(const unsigned long long* __restrict__ table, uint64_t mask, bool has_ones,
                                             unsigned long long key) {
  if (key == kEmpty) return has_ones;
  uint64_t slot = mix64(key) & mask;
  for (;;) {
    const unsigned long long v = __ldg(table + slot);
    if (v == key) return true;
    if (v == kEmpty) return false;
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ unsigned base_code(char ch) {  // kmer/Kmer.cc:139-144, :222-227
  return ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 0u;
}

__device__ __forceinline__ unsigned long long bswap64(unsigned long long x) {
  const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  return ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | (unsigned long long)__byte_perm(hi, 0, 0x0123);
}

__global__ void k_check_reads(const char* __restrict__ seq, unsigned long long seq_base, const unsigned long long* __restrict__ offs, uint64_t n_reads, int k,
                              const unsigned long long* __restrict__ table, uint64_t mask, const uint32_t* __restrict__ has_ones_p,
                              float kmer_vote, uint8_t* record, uint32_t* votes) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const bool has_ones = *has_ones_p != 0u;
  const unsigned long long o0 = offs[r], len = offs[r + 1] - o0;
  const char* s = seq + (o0 - seq_base);
  uint32_t count = 0;
  uint8_t rec = 0;
  if (len >= (unsigned long long)k + 10ull && s[0] != '\0') {  // io/ioFastQ.cc:21-26
    const unsigned long long kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    const int top = 2 * (k - 1);
    unsigned long long fw = 0ull, rc = 0ull;
    for (int i = 0; i < k; ++i) {
      const unsigned long long c = base_code(s[i]);
      fw |= c << (2 * i);
      rc = ((rc << 2) | (3ull - c)) & kmask;
    }
    for (unsigned long long j = 0;; ++j) {
      const unsigned long long rep = bswap64(fw) < bswap64(rc) ? fw : rc;  // (km < tw) ? km : tw, memcmp order
      if (set_contains(table, mask, has_ones, rep)) ++count;
      if (j + (unsigned long long)k >= len) break;
      const unsigned long long c = base_code(s[j + (unsigned long long)k]);
      fw = (fw >> 2) | (c << top);               // Kmer::forwardBase, kmer/Kmer.cc:213-230
      rc = ((rc << 2) | (3ull - c)) & kmask;     // its twin
    }
    const float ratio = __fdiv_rn((float)count, (float)(len - (unsigned long long)k + 1ull));  // :57
    rec = ratio > kmer_vote ? 1 : 0;
  }
  record[r] = rec;
  if (votes) votes[r] = count;
}

}  // namespace

extern "C" int klsh_kmer_set_load(klsh_ctx* ctx, const uint8_t* records, uint64_t n_kmers, int record_bytes) {
  if (!ctx || (n_kmers && !records)) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_kmer_set_load: bad argument");
  if (record_bytes != 8)
    return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_kmer_set_load: records of %d bytes (only Kmer::MAX_K = 32, 8 bytes, is supported)", record_bytes);
  KCUDA(ctx, cudaSetDevice(ctx->device));
  uint64_t cap = 1024;
  while (cap < 2 * n_kmers) cap <<= 1;
  KTRY(dev_reserve(ctx, ctx->rd_table, sizeof(unsigned long long) * cap + 16));
  unsigned long long* table = ctx->rd_table.as<unsigned long long>();
  uint32_t* has_ones = reinterpret_cast<uint32_t*>(table + cap);
  KCUDA(ctx, cudaMemsetAsync(table, 0xFF, sizeof(unsigned long long) * cap, ctx->stream));
  KCUDA(ctx, cudaMemsetAsync(has_ones, 0, 16, ctx->stream));
  ctx->rd_cap = cap;
  ctx->rd_n = n_kmers;
  const uint64_t chunk = 32ull << 20;
  for (uint64_t base = 0; base < n_kmers; base += chunk) {
    const uint64_t c = std::min<uint64_t>(chunk, n_kmers - base);
    KTRY(dev_reserve(ctx, ctx->rd_seq, c * 8));
    KCUDA(ctx, cudaMemcpyAsync(ctx->rd_seq.p, records + base * 8, c * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_set_insert<<<(unsigned)((c + 255) / 256), 256, 0, ctx->stream>>>(ctx->rd_seq.as<unsigned long long>(), c, table, cap - 1, has_ones);
    RLAUNCH(ctx);
    KCUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host buffer may go away, and rd_seq is reused
  }
  KCUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KLSH_OK;
}

extern "C" int klsh_check_reads(klsh_ctx* ctx, int k, const char* seq, const uint64_t* seq_offsets, uint64_t n_reads, float kmer_vote,
                                uint8_t* record, uint32_t* votes) {
  if (!ctx || !seq_offsets || (n_reads && (!record || !seq))) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_check_reads: bad argument");
  if (k < 1 || k > 32) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_check_reads: k = %d (1..32, Kmer::MAX_K)", k);
  if (!ctx->rd_cap) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_check_reads: no k-mer set loaded (klsh_kmer_set_load)");
  for (uint64_t r = 0; r < n_reads; ++r)
    if (seq_offsets[r + 1] < seq_offsets[r]) return klsh_fail(ctx, KLSH_ERR_ARG, "klsh_check_reads: offsets decrease at read %llu", (unsigned long long)r);
  if (!n_reads) return KLSH_OK;
  KCUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t base = seq_offsets[0], bytes = seq_offsets[n_reads] - base;
  KTRY(dev_reserve(ctx, ctx->rd_seq, bytes + 16));
  KTRY(dev_reserve(ctx, ctx->rd_offs, sizeof(uint64_t) * (n_reads + 1)));
  KTRY(dev_reserve(ctx, ctx->rd_rec, n_reads));
  if (votes) KTRY(dev_reserve(ctx, ctx->rd_votes, sizeof(uint32_t) * n_reads));
  cudaStream_t st = ctx->stream;
  if (bytes) KCUDA(ctx, cudaMemcpyAsync(ctx->rd_seq.p, seq + base, bytes, cudaMemcpyHostToDevice, st));
  KCUDA(ctx, cudaMemcpyAsync(ctx->rd_offs.p, seq_offsets, sizeof(uint64_t) * (n_reads + 1), cudaMemcpyHostToDevice, st));
  unsigned long long* table = ctx->rd_table.as<unsigned long long>();
  k_check_reads<<<(unsigned)((n_reads + 127) / 128), 128, 0, st>>>(
      ctx->rd_seq.as<char>(), base, ctx->rd_offs.as<unsigned long long>(), n_reads, k, table, ctx->rd_cap - 1,
      reinterpret_cast<const uint32_t*>(table + ctx->rd_cap), kmer_vote, ctx->rd_rec.as<uint8_t>(), votes ? ctx->rd_votes.as<uint32_t>() : nullptr);
  RLAUNCH(ctx);
  KCUDA(ctx, cudaMemcpyAsync(record, ctx->rd_rec.p, n_reads, cudaMemcpyDeviceToHost, st));
  if (votes) KCUDA(ctx, cudaMemcpyAsync(votes, ctx->rd_votes.p, sizeof(uint32_t) * n_reads, cudaMemcpyDeviceToHost, st));
  KCUDA(ctx, cudaStreamSynchronize(st));
  return KLSH_OK;
}
