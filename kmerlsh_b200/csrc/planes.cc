// Hyperplane sources for the LSH tables.
//
// Built-in source: the reference draws each hash function as
//     random_device rd; mt19937 gen(rd()); normal_distribution<> dis(0,1); f[i] = dis(gen);
// (reference hash/lshash.cc:3-17) and a table as H such functions (:36-42).  Here the
// random_device is replaced by a seeded std::mt19937_64 whose successive outputs, truncated to
// 32 bits, seed the per-function engines; everything else is the same libstdc++ machinery, so a
// seeded reference build (oracle/seeded_rd.h) and this library produce the same floats.
// Callback source: the caller supplies the tables (e.g. the reference's own generator).
#include <random>

#include "klsh_internal.cuh"

struct PlaneSource {
  std::mt19937_64 master;
  uint64_t seed = 12345ULL;
  uint64_t drawn = 0;  // master outputs consumed since the seed (one per hash function)
  klsh_plane_fn fn = nullptr;
  void* user = nullptr;
};

PlaneSource* planes_new() {
  PlaneSource* p = new PlaneSource();
  p->master.seed(12345ULL);
  return p;
}
void planes_free(PlaneSource* p) { delete p; }
void planes_seed(PlaneSource* p, uint64_t seed) {
  p->master.seed(seed);
  p->seed = seed;
  p->drawn = 0;
  p->fn = nullptr;
  p->user = nullptr;
}
// position of the built-in stream: (seed, hash functions drawn so far); seeking replays the seed and discards
void planes_tell(const PlaneSource* p, uint64_t* seed, uint64_t* drawn) {
  *seed = p->seed;
  *drawn = p->drawn;
}
void planes_seek(PlaneSource* p, uint64_t seed, uint64_t drawn) {
  planes_seed(p, seed);
  p->master.discard(drawn);
  p->drawn = drawn;
}
void planes_callback(PlaneSource* p, klsh_plane_fn fn, void* user) {
  p->fn = fn;
  p->user = user;
}

void planes_draw(PlaneSource* p, int H, int D, float* out) {
  if (p->fn) {
    p->fn(p->user, H, D, out);
    return;
  }
  p->drawn += static_cast<uint64_t>(H);
  for (int h = 0; h < H; ++h) {
    std::mt19937 gen(static_cast<unsigned int>(p->master()));
    std::normal_distribution<> dis(0, 1);
    float* f = out + static_cast<size_t>(h) * D;
    for (int i = 0; i < D; ++i) f[i] = dis(gen);
  }
}
