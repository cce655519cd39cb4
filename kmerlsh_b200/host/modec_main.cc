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
// --binary-tmp-ids writes the spill files' member lists as <n>.bin.clust.bin (uint64 count + ids) instead of text.
// --no-tmp-files skips writing the spill when it is not needed.  --resident does the same for an input of
// SEVERAL batches on one GPU: every batch's survivors are appended to a device-resident stash.
// --gpus=N runs the batches on N GPUs (see below), --stats-json=FILE writes one record per LSH iteration.
// Modes K, B and E are outside this tool.
#include <getopt.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "klsh.h"
#include "modee.h"

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
  int gpus = 1;
  uint64_t batch = 100000000ull;
  bool reload_tmp = false, no_tmp_files = false, resident = false;
  int kmer_size = 23, size_thresh = 500000;       // -K, -S, -P, -V: mode E (reference defaults, app/kmerLSH.cc:139-142)
  float pval_thresh = 0.01f, kmer_vote = 0.5f;
  bool binary_tmp_ids = false;  // --binary-tmp-ids: member lists of the spill files as <n>.bin.clust.bin (klsh_set_id_format)
  std::string stats_json;
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

// ---- several GPUs, one hyperplane stream ---------------------------------------------------------------
// The reference draws every hash table of a run from one generator, Cluster() call after Cluster() call
// (app/kmerLSH.cc:311-345: batch after batch).  With several GPUs the batches run on different contexts,
// but the tables still come from ONE seeded stream in the reference's order: a context may draw only when
// it is its call's turn, and hands the stream on as soon as its call has drawn its last table (the
// draws-done callback) — i.e. while it is still merging.  Phase-1 calls (one iteration each) therefore
// overlap almost completely; results are byte-identical to --gpus=1 whatever the number of GPUs.
struct Shared {
  klsh_ctx* gen = nullptr;  // holds the run's seeded stream; never computes
  std::mutex mu;
  std::condition_variable cv;
  uint64_t draw_turn = 0;   // index of the Cluster() call that may draw
  uint64_t save_turn = 0;   // index of the batch that may append to the spill file
  std::atomic<bool> failed{false};
  std::string error;
  std::ostringstream json;  // --stats-json: one record per Cluster() iteration
  void fail(const std::string& msg) {
    std::lock_guard<std::mutex> lk(mu);
    if (!failed.exchange(true)) error = msg;
    cv.notify_all();
  }
};

struct Worker {
  klsh_ctx* ctx = nullptr;
  Shared* sh = nullptr;
  int index = 0, device = 0;
  uint64_t my_call = 0;
};

void plane_tramp(void* user, int H, int D, float* out) {
  Worker* w = static_cast<Worker*>(user);
  std::unique_lock<std::mutex> lk(w->sh->mu);
  w->sh->cv.wait(lk, [&] { return w->sh->draw_turn == w->my_call || w->sh->failed.load(); });
  klsh_draw_table(w->sh->gen, H, D, out);
}

void done_tramp(void* user) {
  Worker* w = static_cast<Worker*>(user);
  std::lock_guard<std::mutex> lk(w->sh->mu);
  if (w->sh->draw_turn == w->my_call) w->sh->draw_turn = w->my_call + 1;
  w->sh->cv.notify_all();
}

std::string iteration_lines(const std::vector<klsh_iter_stats>& st, int dim, float secs) {
  std::ostringstream o;
  for (size_t k = 0; k < st.size(); ++k) {
    if (st[k].rows_in == 0) break;
    o << "Iteration:\t" << (k + 1) << ", cos sim threshold:\t" << st[k].threshold << " dimension : " << dim << std::endl;
    o << "Size of profilings : " << st[k].rows_in << std::endl;
    o << "hashing takes secs:\t" << (st[k].ms_sign + st[k].ms_group) * 1e-3f << std::endl;
    o << "clustering takes secs:\t" << st[k].ms_merge * 1e-3f << std::endl << std::endl;
    o << "merging takes secs:\t" << st[k].ms_compact * 1e-3f << std::endl;
    o << "#k-mers after clustering:\t" << st[k].rows_out << std::endl << std::endl;
  }
  o << "kmerLSH algorithm hash+cluster takes (secs): " << secs << std::endl;
  return o.str();
}

void json_records(Shared& sh, const char* phase, uint64_t call, int gpu, const std::vector<klsh_iter_stats>& st) {
  std::lock_guard<std::mutex> lk(sh.mu);
  for (size_t k = 0; k < st.size(); ++k) {
    if (st[k].rows_in == 0) break;
    sh.json << "{\"phase\": \"" << phase << "\", \"call\": " << call << ", \"gpu\": " << gpu << ", \"iteration\": " << (k + 1)
            << ", \"rows_in\": " << st[k].rows_in << ", \"rows_out\": " << st[k].rows_out << ", \"H\": " << st[k].H
            << ", \"threshold\": " << st[k].threshold << ", \"buckets\": " << st[k].buckets << ", \"bucket_max\": " << st[k].bucket_max
            << ", \"nested_calls\": " << st[k].nested_calls << ", \"eps_margin_rows\": " << st[k].eps_margin_rows
            << ", \"ms_sign\": " << st[k].ms_sign << ", \"ms_group\": " << st[k].ms_group << ", \"ms_merge\": " << st[k].ms_merge
            << ", \"ms_compact\": " << st[k].ms_compact << ", \"ms_total\": " << st[k].ms_total << "}\n";
  }
}

// One Cluster() call on a worker, in stream order `call`.  Returns the reference-style log lines.
bool cluster_call(Worker& w, const char* phase, uint64_t call, float sim, int iters, int64_t thr, int dim, std::string* log) {
  w.my_call = call;
  std::vector<klsh_iter_stats> st((size_t)iters);
  auto t0 = std::chrono::high_resolution_clock::now();
  if (klsh_cluster(w.ctx, sim, iters, thr, st.data()) != KLSH_OK) {
    w.sh->fail(std::string("klsh_cluster failed: ") + klsh_last_error(w.ctx));
    done_tramp(&w);
    return false;
  }
  const float secs = std::chrono::duration_cast<std::chrono::duration<float>>(std::chrono::high_resolution_clock::now() - t0).count();
  if (log) *log = iteration_lines(st, dim, secs);
  json_records(*w.sh, phase, call, w.index, st);
  return true;
}

}  // namespace

static int run_mode_c(Params p);

static ModeEParams mode_e_params(const Params& p) {
  ModeEParams e;
  e.input1 = p.input1;
  e.input2 = p.input2;
  e.output1 = p.output1;
  e.output2 = p.output2;
  e.clust_file_name = p.clust_file_name;
  e.k = p.kmer_size;
  e.pval_thresh = p.pval_thresh;
  e.size_thresh = p.size_thresh;
  e.kmer_vote = p.kmer_vote;
  e.threads_to_use = p.threads_to_use;
  e.verbose = p.verbose;
  e.device = p.device;
  return e;
}

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
                                         {"gpus", required_argument, 0, 1005},
                                         {"stats-json", required_argument, 0, 1006},
                                         {"resident", no_argument, 0, 1007},
                                         {"binary-tmp-ids", no_argument, 0, 1008},
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
      case 'K': if (optarg) p.kmer_size = atoi(optarg); break;
      case 'S': if (optarg) p.size_thresh = atoi(optarg); break;
      case 'P': if (optarg) p.pval_thresh = (float)atof(optarg); break;
      case 'V': if (optarg) p.kmer_vote = (float)atof(optarg); break;
      case 1000: p.have_seed = true; p.seed = strtoull(optarg, 0, 10); break;
      case 1001: p.device = atoi(optarg); break;
      case 1002: p.batch = strtoull(optarg, 0, 10); break;
      case 1003: p.reload_tmp = true; break;
      case 1004: p.no_tmp_files = true; break;
      case 1005: p.gpus = atoi(optarg); break;
      case 1006: p.stats_json = optarg; break;
      case 1007: p.resident = true; break;
      case 1008: p.binary_tmp_ids = true; break;
      default: break;  // -H -X -C: accepted, meaningless for modes C and E
    }
  }
  p.verbose = verbose_flag != 0;
  p.only = only_flag != 0;
  // reference app/kmerLSH.cc:236-271: -M E runs the extraction alone, -M C the clustering and — without --only — the
  // extraction after it; modes K and B (KMC counting, bin-file construction) are the reference's own.
  if (p.mode == "E") return run_mode_e(mode_e_params(p));
  if (p.mode != "C") {
    std::cerr << "kmerLSH_b200 implements modes C and E (-M C [--only], -M E); modes K and B are the reference's." << std::endl;
    return 2;
  }
  const int rc = run_mode_c(p);
  if (rc != 0 || p.only) return rc;
  return run_mode_e(mode_e_params(p));
}

static int run_mode_c(Params p) {
  if (p.batch < 1000) {
    std::cerr << "--batch must be at least 1000" << std::endl;
    return 2;
  }
  if (p.gpus < 1 || p.gpus > 63) {
    std::cerr << "--gpus must be between 1 and 63" << std::endl;
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

  // ---- contexts: one worker per requested GPU (workers share physical GPUs when there are fewer), plus
  // the generator context that holds the run's hyperplane stream
  Shared sh;
  if (klsh_create(p.device, &sh.gen) != KLSH_OK) {
    std::cerr << "klsh_create failed: " << klsh_last_error(nullptr) << std::endl;
    return 1;
  }
  if (!p.have_seed) {
    std::random_device rd;
    p.seed = ((uint64_t)rd() << 32) | rd();
  }
  klsh_set_seed(sh.gen, p.seed);
  int physical = 1;
  for (int d = p.device + 1; d < p.device + p.gpus; ++d) {
    klsh_ctx* probe = nullptr;
    if (klsh_create(d, &probe) != KLSH_OK) break;
    klsh_destroy(probe);
    ++physical;
  }
  std::vector<Worker> workers((size_t)p.gpus);
  for (int g = 0; g < p.gpus; ++g) {
    Worker& w = workers[(size_t)g];
    w.sh = &sh;
    w.index = g;
    w.device = p.device + g % physical;
    if (klsh_create(w.device, &w.ctx) != KLSH_OK) {
      std::cerr << "klsh_create(device " << w.device << ") failed: " << klsh_last_error(nullptr) << std::endl;
      return 1;
    }
    klsh_set_plane_source(w.ctx, plane_tramp, &w);
    klsh_set_draws_done_callback(w.ctx, done_tramp, &w);
    if (p.binary_tmp_ids) klsh_set_id_format(w.ctx, 1);  // spill files only; the result below is written as text
  }
  if (p.gpus > 1)
    std::cout << "kmerLSH_b200: " << p.gpus << " workers on " << physical << " GPU(s); phase-1 batches round-robin, one hyperplane stream"
              << std::endl;

  // ---- init_clustering, reference app/kmerLSH.cc:278-430 -------------------------------------
  const uint64_t batch_thresh = p.batch;
  const int64_t phase1_bucket_thr = (int64_t)(batch_thresh / 1000);
  {
    std::ifstream probe("kmer_count.bin", std::ios::binary);
    if (!probe.is_open()) {
      std::cerr << "cannot open kmer_count.bin" << std::endl;
      return 1;
    }
  }
  uint64_t total_size = 0, next_call = 0;
  int tmp = 0;
  int iter = (int)(kmap_size / batch_thresh);
  std::cout << "iteration : " << iter << " kmap_size : " << kmap_size << std::endl;
  std::string write_tmp = p.tmp_dir + std::to_string(tmp++) + ".bin";

  // One round of batches (phase 1 from kmer_count.bin, or a re-batch round from the previous spill):
  // batch i runs on worker i % gpus; tables are drawn in batch order, spills are appended in batch order.
  struct Batch { uint64_t offset, size, call; int index; };
  auto run_round = [&](const std::vector<Batch>& batches, bool from_counts, const std::string& read_tmp, float sim, int iters,
                       bool keep_resident, bool stash, uint64_t* rows_total) -> bool {
    sh.save_turn = 0;
    std::atomic<uint64_t> total{0};
    auto body = [&](Worker& w) {
      std::vector<uint16_t> counts;
      std::ifstream in;
      if (from_counts) in.open("kmer_count.bin", std::ios::binary);
      for (size_t bi = (size_t)w.index; bi < batches.size(); bi += (size_t)p.gpus) {
        const Batch& b = batches[bi];
        if (sh.failed.load()) return;
        bool ok = true;
        if (from_counts) {
          counts.resize((size_t)tot_sample * b.size);
          for (int j = 0; j < tot_sample; j++) {  // ReadHT, reference io/ioHT.cc:59-81
            in.seekg((std::streamoff)(((uint64_t)j * kmap_size + b.offset) * sizeof(uint16_t)), std::ios::beg);
            in.read(reinterpret_cast<char*>(&counts[(size_t)j * b.size]), (std::streamsize)(sizeof(uint16_t) * b.size));
          }
          ok = klsh_load_counts(w.ctx, counts.data(), v_kmers.data(), tot_sample, b.size, b.offset) == KLSH_OK;
        } else {
          ok = klsh_load_cluster_file(w.ctx, read_tmp.c_str(), tot_sample, b.offset, b.size) == KLSH_OK;
        }
        if (!ok) {
          sh.fail(std::string("loading a batch failed: ") + klsh_last_error(w.ctx));
          w.my_call = b.call;
          done_tramp(&w);
          return;
        }
        std::string log;
        if (!cluster_call(w, from_counts ? "phase1" : "rebatch", b.call, sim, iters, phase1_bucket_thr, tot_sample, p.verbose ? &log : nullptr)) return;
        uint64_t rows = 0;
        klsh_row_count(w.ctx, &rows, nullptr);
        total += rows;
        if (stash && klsh_stash_rows(w.ctx) != KLSH_OK) {  // one worker: batches arrive in order
          sh.fail(std::string("klsh_stash_rows failed: ") + klsh_last_error(w.ctx));
          return;
        }
        {  // spill and log in batch order
          std::unique_lock<std::mutex> lk(sh.mu);
          sh.cv.wait(lk, [&] { return sh.save_turn == (uint64_t)b.index || sh.failed.load(); });
        }
        if (sh.failed.load()) return;
        std::cout << "i: " << b.index << " batch_size : " << b.size << " batch_offset : " << b.offset << std::endl;
        if (p.verbose) std::cout << log;
        if (!keep_resident && klsh_save(w.ctx, write_tmp.c_str(), b.index == 0, 0) != KLSH_OK) {
          sh.fail(std::string("klsh_save failed: ") + klsh_last_error(w.ctx));
          return;
        }
        if (p.verbose) std::cout << "# loaded kmers: " << (b.offset + b.size) << std::endl;
        {
          std::lock_guard<std::mutex> lk(sh.mu);
          sh.save_turn = (uint64_t)b.index + 1;
          sh.cv.notify_all();
        }
      }
    };
    std::vector<std::thread> pool;
    for (int g = 1; g < p.gpus; ++g) pool.emplace_back(body, std::ref(workers[(size_t)g]));
    body(workers[0]);
    for (auto& t : pool) t.join();
    *rows_total = total.load();
    return !sh.failed.load();
  };
  auto make_batches = [&](uint64_t rows, int n_iter) {
    std::vector<Batch> out;
    uint64_t off = 0;
    int index = 0;
    for (int i = 0; i < n_iter + 1; i++) {
      const uint64_t size = (i == n_iter) ? rows - off : batch_thresh;
      if (size == 0) {  // the reference would call Cluster on an empty set here (undefined)
        std::cout << "i: " << i << " batch_size : 0 batch_offset : " << off << std::endl;
        continue;
      }
      out.push_back(Batch{off, size, next_call++, index++});
      off += size;
    }
    return out;
  };

  std::vector<Batch> batches = make_batches(kmap_size, iter);
  // a single batch that needs no re-batching stays on the device (one GPU); its spill is optional then
  const bool single = batches.size() == 1 && !p.reload_tmp && p.gpus == 1;
  // --resident: several batches on one GPU — the survivors of every batch are appended to a device-resident
  // stash (klsh_stash_rows) instead of being re-read from the spill files afterwards
  const bool stashing = p.resident && !single && !p.reload_tmp && p.gpus == 1;
  if (!run_round(batches, true, "", p.min_similarity, 1, (single || stashing) && p.no_tmp_files, stashing, &total_size)) {
    std::cerr << sh.error << std::endl;
    return 1;
  }
  bool resident = single && total_size <= batch_thresh;
  if (stashing) {
    if (klsh_unstash_rows(workers[0].ctx) != KLSH_OK) {
      std::cerr << "klsh_unstash_rows failed: " << klsh_last_error(workers[0].ctx) << std::endl;
      return 1;
    }
    resident = total_size <= batch_thresh;
  }
  if ((single || stashing) && p.no_tmp_files && !resident) {  // the optional spill turned out to be needed after all
    if (klsh_save(workers[0].ctx, write_tmp.c_str(), 1, 0) != KLSH_OK) {
      std::cerr << "klsh_save failed: " << klsh_last_error(workers[0].ctx) << std::endl;
      return 1;
    }
  }

  float similarity = p.min_similarity;
  while (total_size > batch_thresh) {
    similarity -= 0.001;
    const std::string read_tmp = write_tmp;
    write_tmp = p.tmp_dir + std::to_string(tmp++) + ".bin";
    iter = (int)(total_size / batch_thresh);
    batches = make_batches(total_size, iter);
    if (!run_round(batches, false, read_tmp, similarity, 1 + 4, false, false, &total_size)) {
      std::cerr << sh.error << std::endl;
      return 1;
    }
    if (std::remove(read_tmp.c_str()) != 0) perror("The temporary file deletion failed");
    else std::cout << read_tmp << "file are removed" << std::endl;
    const std::string rc = read_tmp + (p.binary_tmp_ids ? ".clust.bin" : ".clust");
    if (std::remove(rc.c_str()) != 0) perror("The temporary file deletion failed");
    else std::cout << rc << "file are removed" << std::endl;
  }

  // ---- the -I iterations, reference app/kmerLSH.cc:490 -----------------------------------------
  // With several physical GPUs the single Cluster() call is sharded over them (klsh_mg_cluster: rows
  // replicated, merge work partitioned, NCCL exchange inside the library); every rank needs the same rows
  // and an identical copy of the hyperplane stream at its current position.
  uint64_t seed0 = 0, drawn0 = 0;
  klsh_plane_tell(sh.gen, &seed0, &drawn0);
  const bool sharded = p.gpus > 1 && physical >= p.gpus;
  klsh_ctx* ctx = workers[0].ctx;
  if (!sharded) {
    if (!resident && klsh_load_cluster_file(ctx, write_tmp.c_str(), tot_sample, 0, 0) != KLSH_OK) {
      std::cerr << "klsh_load_cluster_file failed: " << klsh_last_error(ctx) << std::endl;
      return 1;
    }
    std::string log;
    if (!cluster_call(workers[0], "phase2", next_call++, p.min_similarity, p.cluster_iteration, 1000000, tot_sample, p.verbose ? &log : nullptr)) {
      std::cerr << sh.error << std::endl;
      return 1;
    }
    if (p.verbose) std::cout << log;
  } else {
    unsigned char uid[128];
    if (klsh_nccl_unique_id(uid, sizeof uid) != KLSH_OK) {
      std::cerr << "klsh_nccl_unique_id failed: " << klsh_last_error(nullptr) << std::endl;
      return 1;
    }
    std::vector<std::vector<klsh_iter_stats>> st((size_t)p.gpus, std::vector<klsh_iter_stats>((size_t)p.cluster_iteration));
    auto body = [&](Worker& w) {
      klsh_set_draws_done_callback(w.ctx, nullptr, nullptr);
      if (klsh_mg_init(w.ctx, w.index, p.gpus, uid, sizeof uid) != KLSH_OK ||
          klsh_load_cluster_file(w.ctx, write_tmp.c_str(), tot_sample, 0, 0) != KLSH_OK ||
          klsh_plane_seek(w.ctx, seed0, drawn0) != KLSH_OK ||
          klsh_mg_cluster(w.ctx, p.min_similarity, p.cluster_iteration, 1000000, st[(size_t)w.index].data()) != KLSH_OK)
        sh.fail(std::string("sharded phase 2 failed on worker ") + std::to_string(w.index) + ": " + klsh_last_error(w.ctx));
    };
    std::vector<std::thread> pool;
    for (int g = 1; g < p.gpus; ++g) pool.emplace_back(body, std::ref(workers[(size_t)g]));
    body(workers[0]);
    for (auto& t : pool) t.join();
    if (sh.failed.load()) {
      std::cerr << sh.error << std::endl;
      return 1;
    }
    if (p.verbose) std::cout << iteration_lines(st[0], tot_sample, 0.f);
    json_records(sh, "phase2-sharded", next_call++, 0, st[0]);
  }

  // ---- reference app/kmerLSH.cc:498-499 ---------------------------------------------------------
  if (p.verbose) std::cout << "Saving cluster results starts: " << std::endl;
  klsh_set_id_format(ctx, 0);  // <F>.clust is always the reference's text format
  if (klsh_save(ctx, p.clust_file_name.c_str(), 1, 5) != KLSH_OK) {
    std::cerr << "klsh_save failed: " << klsh_last_error(ctx) << std::endl;
    return 1;
  }
  if (!p.stats_json.empty()) {
    std::ofstream js(p.stats_json.c_str());
    js << sh.json.str();
  }
  for (auto& w : workers) klsh_destroy(w.ctx);
  klsh_destroy(sh.gen);
  return 0;
}
