// kmerLSH_b200 — mode-C command line of kmerLSH on the B200 library.
//
// Same flags, cwd-relative inputs and outputs as `kmerLSH -M C` (reference app/kmerLSH.cc:147-276,
// :432-521): reads kmer_count.bin / kmer_count.log and the two sample lists (only their line
// counts matter), spills batch results to <tmp_dir>/<n>.bin(.clust) exactly like init_clustering
// (:278-430), runs the -I iterations and writes <F> + <F>.clust with clusters of more than 5
// members.  All clustering goes through the C ABI in include/klsh.h; there is no CPU path.
// Additions: --seed=N (seeded hyperplanes; default draws a seed from std::random_device like the
// reference), --device=N, --batch=N (rows per phase-1 batch, reference constant 100000000).
// When the whole input is ONE batch, the phase-1 survivors stay resident on the device for the -I
// iterations instead of being re-read from the spill files (the rows, their order and their member
// lists are the same either way, so the results are too — SURVEY.md section 8f item 3); the spill
// files are still written for compatibility.  --reload-tmp forces the reference's file round trip,
// --no-tmp-files skips writing the spill when it is not needed.
// Modes K, B and E are outside this tool.
#include <getopt.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "klsh.h"

namespace {

struct Params {
  int cluster_iteration = 100;
  float min_similarity = 0.80f;
  unsigned threads_to_use = 12;
  bool verbose = false;
  std::string tmp_dir = "tmp/";
  std::string input1, input2, output1, output2;
  std::string clust_file_name = "clustering_result.txt";
  std::string mode;
  bool only = false;
  bool have_seed = false;
  uint64_t seed = 0;
  int device = 0;
  uint64_t batch = 100000000ull;
  bool reload_tmp = false, no_tmp_files = false;
};

int count_lines(const std::string& path) {  // GetInput, reference io/ioHT.cc:3-19
  std::ifstream in(path.c_str());
  if (!in.is_open()) {
    std::cerr << "Unable to open info file";
    return 0;
  }
  int n = 0;
  std::string line;
  while (std::getline(in, line)) ++n;
  return n;
}

#define CK(ctx, call)                                                        \
  do {                                                                       \
    int rc__ = (call);                                                       \
    if (rc__ != KLSH_OK) {                                                   \
      std::cerr << #call << " failed: " << klsh_last_error(ctx) << std::endl; \
      return 1;                                                              \
    }                                                                        \
  } while (0)

void print_iterations(const std::vector<klsh_iter_stats>& st, int dim) {
  for (size_t k = 0; k < st.size(); ++k) {
    if (st[k].rows_in == 0) break;
    std::cout << "Iteration:\t" << (k + 1) << ", cos sim threshold:\t" << st[k].threshold << " dimension : " << dim << std::endl;
    std::cout << "Size of profilings : " << st[k].rows_in << std::endl;
    std::cout << "hashing takes secs:\t" << (st[k].ms_sign + st[k].ms_group) * 1e-3f << std::endl;
    std::cout << "clustering takes secs:\t" << st[k].ms_merge * 1e-3f << std::endl << std::endl;
    std::cout << "merging takes secs:\t" << st[k].ms_compact * 1e-3f << std::endl;
    std::cout << "#k-mers after clustering:\t" << st[k].rows_out << std::endl << std::endl;
  }
}

int cluster_logged(klsh_ctx* ctx, float sim, int iters, int64_t thr, int dim, bool verbose) {
  std::vector<klsh_iter_stats> st((size_t)iters);
  auto t0 = std::chrono::high_resolution_clock::now();
  CK(ctx, klsh_cluster(ctx, sim, iters, thr, st.data()));
  if (verbose) {
    print_iterations(st, dim);
    float secs = std::chrono::duration_cast<std::chrono::duration<float>>(std::chrono::high_resolution_clock::now() - t0).count();
    std::cout << "kmerLSH algorithm hash+cluster takes (secs): " << secs << std::endl;
  }
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  Params p;
  int verbose_flag = 0, only_flag = 0;
  const char* opt_string = "o:p:a:b:H:I:N:X:C:T:K:S:P:V:F:M:";
  static struct option long_options[] = {{"verbose", no_argument, &verbose_flag, 1},
                                         {"only", no_argument, &only_flag, 1},
                                         {"output1", required_argument, 0, 'o'},
                                         {"output2", required_argument, 0, 'p'},
                                         {"input1", required_argument, 0, 'a'},
                                         {"input2", required_argument, 0, 'b'},
                                         {"cluster_iteration", optional_argument, 0, 'I'},
                                         {"min_similarity", optional_argument, 0, 'N'},
                                         {"max-memory", optional_argument, 0, 'X'},
                                         {"count-min", optional_argument, 0, 'C'},
                                         {"threads_to_use", optional_argument, 0, 'T'},
                                         {"kmer_size", optional_argument, 0, 'K'},
                                         {"size_thresh", optional_argument, 0, 'S'},
                                         {"pval_thresh", optional_argument, 0, 'P'},
                                         {"kmer_vote", optional_argument, 0, 'V'},
                                         {"tmp_dir", optional_argument, 0, 'D'},
                                         {"clust_file_name", optional_argument, 0, 'F'},
                                         {"mode", optional_argument, 0, 'M'},
                                         {"seed", required_argument, 0, 1000},
                                         {"device", required_argument, 0, 1001},
                                         {"batch", required_argument, 0, 1002},
                                         {"reload-tmp", no_argument, 0, 1003},
                                         {"no-tmp-files", no_argument, 0, 1004},
                                         {0, 0, 0, 0}};
  for (;;) {
    int idx = 0;
    int c = getopt_long(argc, argv, opt_string, long_options, &idx);
    if (c == -1) break;
    switch (c) {
      case 'o': p.output1 = optarg; break;
      case 'p': p.output2 = optarg; break;
      case 'a': p.input1 = optarg; break;
      case 'b': p.input2 = optarg; break;
      case 'I': if (optarg) p.cluster_iteration = atoi(optarg); break;
      case 'N': if (optarg) p.min_similarity = (float)atof(optarg); break;
      case 'T': if (optarg) p.threads_to_use = (unsigned)atoi(optarg); break;
      case 'D': if (optarg) p.tmp_dir = optarg; break;
      case 'F': if (optarg) p.clust_file_name = optarg; break;
      case 'M': if (optarg) p.mode = optarg; break;
      case 1000: p.have_seed = true; p.seed = strtoull(optarg, 0, 10); break;
      case 1001: p.device = atoi(optarg); break;
      case 1002: p.batch = strtoull(optarg, 0, 10); break;
      case 1003: p.reload_tmp = true; break;
      case 1004: p.no_tmp_files = true; break;
      default: break;  // -H -X -C -K -S -P -V: accepted, meaningless for mode C
    }
  }
  p.verbose = verbose_flag != 0;
  p.only = only_flag != 0;
  if (p.mode != "C") {
    std::cerr << "kmerLSH_b200 implements mode C only (-M C [--only]); modes K, B and E are the reference's." << std::endl;
    return 2;
  }
  if (p.batch < 1000) {
    std::cerr << "--batch must be at least 1000" << std::endl;
    return 2;
  }

  const int num_sample1 = count_lines(p.input1), num_sample2 = count_lines(p.input2);
  const int tot_sample = num_sample1 + num_sample2;
  if (p.verbose)
    std::cout << std::endl << "# samples in group 1: " << num_sample1 << std::endl << "# samples in group 2: " << num_sample2 << std::endl;
  if (tot_sample <= 0) {
    std::cerr << "no samples listed in -a/-b files" << std::endl;
    return 1;
  }

  // reference app/kmerLSH.cc:473-481
  size_t kmap_size = 0;
  std::vector<float> v_kmers;
  {
    std::ifstream logStream("kmer_count.log");
    if (!logStream.is_open()) {
      std::cerr << "cannot open kmer_count.log" << std::endl;
      return 1;
    }
    std::string line;
    std::getline(logStream, line);
    std::istringstream ss(line);
    ss >> kmap_size;
    for (int i = 0; i < tot_sample; i++) {
      float kmer_coverage = 0;
      ss >> kmer_coverage;
      v_kmers.push_back(kmer_coverage / kmap_size);
    }
  }

  klsh_ctx* ctx = nullptr;
  if (klsh_create(p.device, &ctx) != KLSH_OK) {
    std::cerr << "klsh_create failed: " << klsh_last_error(nullptr) << std::endl;
    return 1;
  }
  if (!p.have_seed) {
    std::random_device rd;
    p.seed = ((uint64_t)rd() << 32) | rd();
  }
  CK(ctx, klsh_set_seed(ctx, p.seed));

  // ---- init_clustering, reference app/kmerLSH.cc:278-430 -------------------------------------
  const uint64_t batch_thresh = p.batch;
  const int64_t phase1_bucket_thr = (int64_t)(batch_thresh / 1000);
  std::ifstream inStream("kmer_count.bin", std::ios::binary);
  if (!inStream.is_open()) {
    std::cerr << "cannot open kmer_count.bin" << std::endl;
    return 1;
  }
  uint64_t batch_offset = 0, total_size = 0;
  int tmp = 0, batches_run = 0;
  int iter = (int)(kmap_size / batch_thresh);
  std::cout << "iteration : " << iter << " kmap_size : " << kmap_size << std::endl;
  std::string write_tmp = p.tmp_dir + std::to_string(tmp++) + ".bin";
  std::vector<uint16_t> counts;
  for (int i = 0; i < iter + 1; i++) {
    const uint64_t batch_size = (i == iter) ? kmap_size - batch_offset : batch_thresh;
    std::cout << "i: " << i << " batch_size : " << batch_size << " batch_offset : " << batch_offset << std::endl;
    if (batch_size == 0) continue;  // the reference would call Cluster on an empty set here (undefined)
    counts.resize((size_t)tot_sample * batch_size);
    for (int j = 0; j < tot_sample; j++) {  // ReadHT, reference io/ioHT.cc:59-81
      inStream.seekg((std::streamoff)(((uint64_t)j * kmap_size + batch_offset) * sizeof(uint16_t)), std::ios::beg);
      inStream.read(reinterpret_cast<char*>(&counts[(size_t)j * batch_size]), (std::streamsize)(sizeof(uint16_t) * batch_size));
    }
    CK(ctx, klsh_load_counts(ctx, counts.data(), v_kmers.data(), tot_sample, batch_size, batch_offset));
    if (cluster_logged(ctx, p.min_similarity, 1, phase1_bucket_thr, tot_sample, p.verbose)) return 1;
    uint64_t rows = 0;
    CK(ctx, klsh_row_count(ctx, &rows, nullptr));
    total_size += rows;
    ++batches_run;
    // a single batch that needs no re-batching stays on the device; its spill is optional then
    const bool resident_ok = !p.reload_tmp && iter == 0 && rows <= batch_thresh;
    if (!(resident_ok && p.no_tmp_files)) CK(ctx, klsh_save(ctx, write_tmp.c_str(), i == 0, 0));
    batch_offset += batch_size;
    if (p.verbose) std::cout << "# loaded kmers: " << batch_offset << std::endl;
  }
  const bool resident = !p.reload_tmp && iter == 0 && batches_run == 1 && total_size <= batch_thresh;
  counts.clear();
  counts.shrink_to_fit();
  inStream.close();

  float similarity = p.min_similarity;
  while (total_size > batch_thresh) {
    similarity -= 0.001;
    batch_offset = 0;
    const std::string read_tmp = write_tmp;
    write_tmp = p.tmp_dir + std::to_string(tmp++) + ".bin";
    iter = (int)(total_size / batch_thresh);
    const uint64_t kcnt_rem = total_size;
    total_size = 0;
    for (int i = 0; i < iter + 1; i++) {
      const uint64_t batch_size = (i == iter) ? kcnt_rem - batch_offset : batch_thresh;
      std::cout << "i: " << i << " batch_size : " << batch_size << " batch_offset : " << batch_offset << std::endl;
      if (batch_size == 0) continue;
      CK(ctx, klsh_load_cluster_file(ctx, read_tmp.c_str(), tot_sample, batch_offset, batch_size));
      if (cluster_logged(ctx, similarity, 1 + 4, phase1_bucket_thr, tot_sample, p.verbose)) return 1;
      uint64_t rows = 0;
      CK(ctx, klsh_row_count(ctx, &rows, nullptr));
      total_size += rows;
      CK(ctx, klsh_save(ctx, write_tmp.c_str(), i == 0, 0));
      batch_offset += batch_size;
      if (p.verbose) std::cout << "# loaded kmers: " << batch_offset << std::endl;
    }
    if (std::remove(read_tmp.c_str()) != 0) perror("The temporary file deletion failed");
    else std::cout << read_tmp << "file are removed" << std::endl;
    const std::string rc = read_tmp + ".clust";
    if (std::remove(rc.c_str()) != 0) perror("The temporary file deletion failed");
    else std::cout << rc << "file are removed" << std::endl;
  }
  if (!resident) CK(ctx, klsh_load_cluster_file(ctx, write_tmp.c_str(), tot_sample, 0, 0));

  // ---- the -I iterations, reference app/kmerLSH.cc:490 -----------------------------------------
  if (cluster_logged(ctx, p.min_similarity, p.cluster_iteration, 1000000, tot_sample, p.verbose)) return 1;

  // ---- reference app/kmerLSH.cc:498-499 ---------------------------------------------------------
  if (p.verbose) std::cout << "Saving cluster results starts: " << std::endl;
  CK(ctx, klsh_save(ctx, p.clust_file_name.c_str(), 1, 5));
  klsh_destroy(ctx);
  return 0;
}
