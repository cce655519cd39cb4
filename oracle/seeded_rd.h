// TEST INFRASTRUCTURE — not product code.
//
// Seeding shim for the reference build under oracle/_ref (see oracle/Makefile).
// The reference draws every LSH hash function from a fresh, unseeded
// std::random_device (hash/lshash.cc:3-17, :36-42), so two runs never agree.
// This header is force-included (-include) ONLY when compiling the reference's
// hash/lshash.cc: it replaces `random_device` with a deterministic source whose
// successive 32-bit outputs are the successive outputs of one process-wide
// std::mt19937_64(seed).  Everything downstream (mt19937 gen(rd()),
// normal_distribution<double>, narrowing to float) remains the reference's code.
//
// Seed: klsh_oracle::reseed(s) (used by the harness), else env KLSH_SEED, else 12345.
#ifndef KLSH_ORACLE_SEEDED_RD_H
#define KLSH_ORACLE_SEEDED_RD_H
#include <algorithm>
#include <random>
#include <string>
#include <vector>
#include <cstdlib>

namespace klsh_oracle {

struct master_state {
  std::mt19937_64 eng;
  unsigned long long draws;
  master_state() : draws(0) {
    const char* e = std::getenv("KLSH_SEED");
    eng.seed(e ? std::strtoull(e, 0, 10) : 12345ULL);
  }
};

// One instance per process (inline function-local static is shared across TUs).
inline master_state& master() {
  static master_state m;
  return m;
}

inline void reseed(unsigned long long s) {
  master().eng.seed(s);
  master().draws = 0;
}

struct seeded_random_device {
  typedef unsigned int result_type;
  result_type operator()() {
    master_state& m = master();
    ++m.draws;
    return (result_type)m.eng();
  }
};

}  // namespace klsh_oracle

#define random_device klsh_oracle::seeded_random_device
#endif
