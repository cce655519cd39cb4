// Mode E of the command line (kmerlsh_b200/host/modee.cc): the reference's statistics + read-extraction branch
// (app/kmerLSH.cc:525-596) on the B200 library.
#pragma once
#include <string>

struct ModeEParams {
  std::string input1, input2, output1, output2;
  std::string clust_file_name = "clustering_result.txt";
  int k = 23;                 // -K, reference default (app/kmerLSH.cc:139)
  float pval_thresh = 0.01f;  // -P
  int size_thresh = 500000;   // -S
  float kmer_vote = 0.5f;     // -V
  unsigned threads_to_use = 12;
  bool verbose = false;
  int device = 0;
};

int run_mode_e(const ModeEParams& p);
