// kmerLSH_b200 — mode E (statistics + read extraction) on the B200 library.
//
// Mirrors the reference's `kmerLSH -M E` branch (app/kmerLSH.cc:525-596) and IOFQ::Extracting / ReadExtract
// (io/ioFastQ.cc:77-195): read the clustering result (<F>, <F>.clust), run the t-test on every cluster, label every
// k-mer id, pick the differential k-mers out of kmer_set.hex, and write — per input FASTQ file of each group — the
// reads whose share of differential k-mers exceeds --kmer_vote to <output>_<basename>.  The computing steps are
// library calls (klsh_differential_ids, klsh_select_kmers, klsh_kmer_set_load, klsh_check_reads); FASTQ parsing and
// writing stay here on the host.
//
// FASTQ records are read the way the reference's (modified) kseq.h reads them (kmer/kseq.h:153-212): the name is the
// whole header line after '@' (or '>'), the sequence is every printable character up to the next '+', '>' or '@', the
// quality string is as long as the sequence; a record with a truncated quality string ends the file's extraction as it
// does in the reference (FastqFile::read stops, ReadExtract leaves its loop on the short part).  Plain and gzip
// files are both read through zlib, as there.  Reads are tested in parts of 65 536 (FastqFile::part_size).
#include <zlib.h>

#include <cctype>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "klsh.h"
#include "modee.h"

namespace {

// GetInput, reference io/ioHT.cc:3-19: first token of every line is a sample's FASTQ path
std::vector<std::string> sample_paths(const std::string& list) {
  std::vector<std::string> out;
  std::ifstream in(list.c_str());
  if (!in.is_open()) {
    std::cerr << "Unable to open info file";
    return out;
  }
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::string sample, kmc_name;
    ss >> sample >> kmc_name;
    out.push_back(sample);
  }
  return out;
}

// kseq_read over a gzFile, reference kmer/kseq.h:60-212
class FastqReader {
 public:
  explicit FastqReader(const std::string& path) : f_(gzopen(path.c_str(), "r")) {}
  ~FastqReader() {
    if (f_) gzclose(f_);
  }
  bool ok() const { return f_ != nullptr; }
  // >= 0: sequence length; -1: end of file; -2: truncated quality string
  int next(std::string& name, std::string& seq, std::string& qual) {
    int c;
    if (last_char_ == 0) {
      while ((c = getc()) != -1 && c != '>' && c != '@') {
      }
      if (c == -1) return -1;
      last_char_ = c;
    }
    seq.clear();
    qual.clear();
    if (!get_line(name, &c)) return -1;
    while ((c = getc()) != -1 && c != '>' && c != '+' && c != '@')
      if (isgraph(c)) seq.push_back((char)c);
    if (c == '>' || c == '@') last_char_ = c;
    if (c != '+') return (int)seq.size();
    while ((c = getc()) != -1 && c != '\n') {
    }
    if (c == -1) return -2;
    while ((c = getc()) != -1 && qual.size() < seq.size())
      if (c >= 33 && c <= 127) qual.push_back((char)c);
    last_char_ = 0;
    if (seq.size() != qual.size()) return -2;
    return (int)seq.size();
  }

 private:
  int getc() {
    if (eof_ && begin_ >= end_) return -1;
    if (begin_ >= end_) {
      begin_ = 0;
      end_ = gzread(f_, buf_, sizeof buf_);
      if (end_ < (int)sizeof buf_) eof_ = true;
      if (end_ <= 0) {
        end_ = 0;
        return -1;
      }
    }
    return (int)(unsigned char)buf_[begin_++];
  }
  // ks_getuntil(ks, '\n', &name, &c): false at end of input before anything was read
  bool get_line(std::string& out, int* dret) {
    out.clear();
    *dret = 0;
    if (begin_ >= end_ && eof_) return false;
    for (;;) {
      if (begin_ >= end_) {
        if (eof_) break;
        begin_ = 0;
        end_ = gzread(f_, buf_, sizeof buf_);
        if (end_ < (int)sizeof buf_) eof_ = true;
        if (end_ <= 0) {
          end_ = 0;
          break;
        }
      }
      int i = begin_;
      while (i < end_ && buf_[i] != '\n') ++i;
      out.append(buf_ + begin_, (size_t)(i - begin_));
      begin_ = i + 1;
      if (i < end_) {
        *dret = '\n';
        break;
      }
    }
    return true;
  }
  gzFile f_;
  char buf_[4096];
  int begin_ = 0, end_ = 0;
  bool eof_ = false;
  int last_char_ = 0;
};

std::string base_name(const std::string& path) {
  const size_t p = path.find_last_of('/');
  return p == std::string::npos ? path : path.substr(p + 1);
}

// IOFQ::ReadExtract (io/ioFastQ.cc:77-158) for one file; the k-mer set is already loaded in ctx
bool extract_file(klsh_ctx* ctx, const std::string& fastq, const std::string& out_path, int k, float kmer_vote) {
  FastqReader fq(fastq);
  FILE* of = std::fopen(out_path.c_str(), "w");
  if (of == nullptr) {
    std::cerr << "Could not open file for writing, " << out_path << std::endl;
    return false;
  }
  if (!fq.ok()) {
    std::cerr << "Could not open " << fastq << std::endl;
    std::fclose(of);
    return false;
  }
  const size_t part_size = 1u << 16;
  std::vector<std::string> names, quals;
  std::string seqs, name, seq, qual, out;
  std::vector<uint64_t> offs;
  std::vector<uint8_t> rec;
  for (;;) {
    names.clear();
    quals.clear();
    seqs.clear();
    offs.assign(1, 0);
    while (names.size() < part_size && fq.next(name, seq, qual) >= 0) {
      names.push_back(name);
      quals.push_back(qual);
      seqs += seq;
      offs.push_back(seqs.size());
    }
    const size_t n = names.size();
    if (n == 0) break;
    rec.assign(n, 0);
    if (klsh_check_reads(ctx, k, seqs.data(), offs.data(), n, kmer_vote, rec.data(), nullptr) != KLSH_OK) {
      std::cerr << "klsh_check_reads failed: " << klsh_last_error(ctx) << std::endl;
      std::fclose(of);
      return false;
    }
    out.clear();
    for (size_t i = 0; i < n; ++i)
      if (rec[i]) {
        out.push_back('@');
        out += names[i];
        out.push_back('\n');
        out.append(seqs, (size_t)offs[i], (size_t)(offs[i + 1] - offs[i]));
        out += "\n+\n";
        out += quals[i];
        out.push_back('\n');
      }
    if (!out.empty()) std::fwrite(out.data(), 1, out.size(), of);
    if (n < part_size) break;
  }
  std::fclose(of);
  return true;
}

}  // namespace

int run_mode_e(const ModeEParams& p) {
  const std::vector<std::string> samples1 = sample_paths(p.input1), samples2 = sample_paths(p.input2);
  const int num_sample1 = (int)samples1.size(), num_sample2 = (int)samples2.size();
  const int tot_sample = num_sample1 + num_sample2;
  if (tot_sample <= 0) {
    std::cerr << "no samples listed in -a/-b files" << std::endl;
    return 1;
  }
  if (p.k < 1 || p.k > 32) {
    std::cerr << "kmer_size must be between 1 and 32 (Kmer::MAX_K)" << std::endl;
    return 2;
  }
  if (p.verbose) std::cout << "Start to extract the differential reads from raw data" << std::endl;
  klsh_ctx* ctx = nullptr;
  if (klsh_create(p.device, &ctx) != KLSH_OK) {
    std::cerr << "klsh_create failed: " << klsh_last_error(nullptr) << std::endl;
    return 1;
  }
  auto fail = [&](const std::string& what) {
    std::cerr << what << ": " << klsh_last_error(ctx) << std::endl;
    klsh_destroy(ctx);
    return 1;
  };
  // app/kmerLSH.cc:541: IOMat::ReadClusterAll(clusteredab_ptr, tot_sample, params.clust_file_name, ...)
  if (klsh_load_cluster_file(ctx, p.clust_file_name.c_str(), tot_sample, 0, 0) != KLSH_OK) return fail("klsh_load_cluster_file failed");
  // :556-561: kmap_size is the first number of kmer_count.log
  uint64_t kmap_size = 0;
  {
    std::ifstream logStream("kmer_count.log");
    if (!logStream.is_open()) {
      std::cerr << "cannot open kmer_count.log" << std::endl;
      klsh_destroy(ctx);
      return 1;
    }
    std::string line;
    std::getline(logStream, line);
    std::istringstream ss(line);
    ss >> kmap_size;
  }
  // :543-545: AB::WRS on every cluster -> the two id sets, here one label per k-mer id
  std::vector<uint8_t> label((size_t)kmap_size + 1);
  klsh_ttest_stats st;
  if (klsh_differential_ids(ctx, num_sample1, num_sample2, p.pval_thresh, p.size_thresh, kmap_size, label.data(), &st) != KLSH_OK)
    return fail("klsh_differential_ids failed");
  if (p.verbose) {
    std::cout << "# of differential kmers in group A : " << st.ids_a << std::endl;
    std::cout << "# of differential kmers in group B : " << st.ids_b << std::endl;
  }
  // :563-579: the k-mers of kmer_set.hex whose id is in either set
  const int rb = 8;  // Kmer::MAX_K / 4
  std::vector<uint8_t> hex((size_t)kmap_size * rb + 1), a((size_t)kmap_size * rb + 1), b((size_t)kmap_size * rb + 1);
  {
    FILE* kf = std::fopen("kmer_set.hex", "rb");
    if (kf == nullptr) {
      std::cerr << "cannot open kmer_set.hex" << std::endl;
      klsh_destroy(ctx);
      return 1;
    }
    const size_t got = std::fread(hex.data(), rb, (size_t)kmap_size, kf);
    std::fclose(kf);
    if (got != kmap_size) {
      std::cerr << "kmer_set.hex holds " << got << " k-mers, kmer_count.log says " << kmap_size << std::endl;
      klsh_destroy(ctx);
      return 1;
    }
  }
  uint64_t na = 0, nb = 0;
  if (klsh_select_kmers(ctx, hex.data(), kmap_size, rb, label.data(), a.data(), &na, b.data(), &nb) != KLSH_OK)
    return fail("klsh_select_kmers failed");
  // :583-584: IOFQ::Extracting for both groups (io/ioFastQ.cc:161-193)
  for (int g = 0; g < 2; ++g) {
    const std::vector<std::string>& samples = g == 0 ? samples1 : samples2;
    const std::string& prefix = g == 0 ? p.output1 : p.output2;
    if (klsh_kmer_set_load(ctx, g == 0 ? a.data() : b.data(), g == 0 ? na : nb, rb) != KLSH_OK) return fail("klsh_kmer_set_load failed");
    if (p.verbose) std::cout << "start " << p.threads_to_use << " threads" << std::endl;
    for (const std::string& s : samples) {
      const std::string filename = prefix + "_" + base_name(s);
      if (p.verbose) std::cout << "writing to " << filename << std::endl;
      if (!extract_file(ctx, s, filename, p.k, p.kmer_vote)) {
        klsh_destroy(ctx);
        return 1;
      }
    }
  }
  klsh_destroy(ctx);
  return 0;
}
