// TEST INFRASTRUCTURE — the reference-side binding of INTEGRATION.md section B, compiled and linked
// into the reference's own program so that the drop-in claim is executed, not just written down.
//
// This translation unit takes the place of the reference's function/cluster.cc in the link of
// app/kmerLSH.cc (oracle/Makefile target `shim` -> oracle/_ref/kmerLSH_shim): same signature
// (function/cluster.h:42), same ownership rules (the caller's vector is rewritten in place, the
// caller deletes the survivors), and — through the plane callback — the reference's own
// LSH::generateHashTable (hash/lshash.cc:36-42), so the hyperplane stream is the reference's draw
// for draw.  Everything between is libklsh (include/klsh.h).
#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "function/cluster.h"
#include "klsh.h"

static void planes_from_reference(void*, int H, int D, float* out) {
  hashTable t = LSH::generateHashTable(H, D);
  for (int h = 0; h < H; ++h) std::copy(t[h].begin(), t[h].end(), out + (size_t)h * D);
}

void Cluster(vector<Abundance*>* rows, float min_similarity, int cluster_iteration, unsigned int /*threads_to_use*/, int dim,
             int bucket_size_threshold, bool verbose) {
  static klsh_ctx* ctx = nullptr;
  if (!ctx && klsh_create(0, &ctx) != KLSH_OK) {
    std::cerr << klsh_last_error(nullptr) << std::endl;
    exit(-1);
  }
  klsh_set_plane_source(ctx, planes_from_reference, nullptr);

  // vector<Abundance*>  ->  values[n][dim], id_offsets[n+1], ids[]
  const size_t n = rows->size();
  std::vector<float> values(n * (size_t)dim);
  std::vector<uint64_t> offs(n + 1, 0), ids;
  for (size_t r = 0; r < n; ++r) {
    std::copy((*rows)[r]->_values.begin(), (*rows)[r]->_values.end(), values.begin() + r * (size_t)dim);
    ids.insert(ids.end(), (*rows)[r]->_ids.begin(), (*rows)[r]->_ids.end());
    offs[r + 1] = ids.size();
    delete (*rows)[r];
  }
  std::vector<klsh_iter_stats> st((size_t)cluster_iteration);
  if (klsh_set_rows(ctx, values.data(), offs.data(), ids.data(), n, dim) != KLSH_OK ||
      klsh_cluster(ctx, min_similarity, cluster_iteration, bucket_size_threshold, st.data()) != KLSH_OK) {
    std::cerr << klsh_last_error(ctx) << std::endl;
    exit(-1);
  }
  uint64_t m = 0, n_ids = 0;
  klsh_row_count(ctx, &m, &n_ids);
  values.resize(m * (size_t)dim + 1);
  offs.resize(m + 1);
  ids.resize(n_ids + 1);
  if (klsh_get_rows(ctx, values.data(), offs.data(), ids.data()) != KLSH_OK) {
    std::cerr << klsh_last_error(ctx) << std::endl;
    exit(-1);
  }
  rows->clear();
  for (uint64_t r = 0; r < m; ++r) {
    Abundance* ab = new Abundance();
    ab->_values.assign(values.begin() + r * (size_t)dim, values.begin() + (r + 1) * (size_t)dim);
    ab->_ids.assign(ids.begin() + offs[r], ids.begin() + offs[r + 1]);
    rows->push_back(ab);
  }
  if (verbose)
    for (int k = 0; k < cluster_iteration && st[k].rows_in; ++k)
      std::cout << "Iteration:\t" << k + 1 << ", cos sim threshold:\t" << st[k].threshold << " dimension : " << dim << std::endl
                << "Size of profilings : " << st[k].rows_in << std::endl
                << "#k-mers after clustering:\t" << st[k].rows_out << std::endl;
}
