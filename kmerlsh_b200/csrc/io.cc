// File formats either side of the hot path (the drop-in boundary of mode C).
//
//   <F>.clust : one text line per cluster, "<count>\t<id>\t<id>...\n"   (reference
//               IOMat::SaveResult, io/ioMatrix.cc:265-294)
//   <F>       : D raw float32 per cluster, same order, no header          (reference
//               IOMat::SaveBinary, io/ioMatrix.cc:322-351)
// Only clusters with more than ignore_small members are written.  delfile removes both files
// first; both are opened in append mode (the reference appends batch after batch into tmp/0.bin).
// Reading back: row k of <F> <-> line k of <F>.clust, first token = member count
// (IOMat::ReadCluster / ReadClusterAll, io/ioMatrix.cc:121-196, :48-119).
//
// Opt-in binary member lists (klsh_set_id_format(ctx, 1); no reference counterpart):
//   <F>.clust.bin : per cluster uint64 count, then count uint64 ids, host endianness, same order as <F>.
// Formatting and parsing decimal text is what the spill of 10^8 ids per batch costs; the binary file is read and
// written with plain fread/fwrite.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "klsh_internal.cuh"

int io_save(const char* bin_path, int delfile, int64_t ignore_small, const float* values, int D,
            const uint64_t* id_offsets, const uint64_t* ids, uint64_t n, int id_format) {
  std::string clust = std::string(bin_path) + (id_format == 1 ? ".clust.bin" : ".clust");
  if (delfile) {
    std::remove(clust.c_str());
    std::remove(bin_path);
  }
  if (id_format == 1) {
    FILE* fi = std::fopen(clust.c_str(), "ab");
    FILE* fv = std::fopen(bin_path, "ab");
    if (!fi || !fv) {
      if (fi) std::fclose(fi);
      if (fv) std::fclose(fv);
      return KLSH_ERR_IO;
    }
    for (uint64_t r = 0; r < n; ++r) {
      const uint64_t b = id_offsets[r], e = id_offsets[r + 1], cnt = e - b;
      if ((int64_t)cnt <= ignore_small) continue;
      std::fwrite(&cnt, sizeof cnt, 1, fi);
      std::fwrite(ids + b, sizeof(uint64_t), (size_t)cnt, fi);
      std::fwrite(values + r * (uint64_t)D, sizeof(float), (size_t)D, fv);
    }
    const int bad = std::ferror(fi) | std::ferror(fv);
    std::fclose(fi);
    std::fclose(fv);
    return bad ? KLSH_ERR_IO : KLSH_OK;
  }
  FILE* ft = std::fopen(clust.c_str(), "a");
  FILE* fb = std::fopen(bin_path, "ab");
  if (!ft || !fb) {
    if (ft) std::fclose(ft);
    if (fb) std::fclose(fb);
    return KLSH_ERR_IO;
  }
  std::vector<char> line;
  char num[24];
  for (uint64_t r = 0; r < n; ++r) {
    uint64_t b = id_offsets[r], e = id_offsets[r + 1];
    if ((int64_t)(e - b) <= ignore_small) continue;
    line.clear();
    int k = std::snprintf(num, sizeof num, "%llu", (unsigned long long)(e - b));
    line.insert(line.end(), num, num + k);
    for (uint64_t s = b; s < e; ++s) {
      line.push_back('\t');
      k = std::snprintf(num, sizeof num, "%llu", (unsigned long long)ids[s]);
      line.insert(line.end(), num, num + k);
    }
    line.push_back('\n');
    std::fwrite(line.data(), 1, line.size(), ft);
    std::fwrite(values + r * (uint64_t)D, sizeof(float), (size_t)D, fb);
  }
  int bad = std::ferror(ft) | std::ferror(fb);
  std::fclose(ft);
  std::fclose(fb);
  return bad ? KLSH_ERR_IO : KLSH_OK;
}

int io_read_cluster(const char* bin_path, int D, uint64_t start_line, uint64_t num_lines,
                    std::vector<float>& values, std::vector<uint64_t>& id_offsets, std::vector<uint64_t>& ids, int id_format) {
  FILE* fb = std::fopen(bin_path, "rb");
  if (!fb) return KLSH_ERR_IO;
  std::fseek(fb, 0, SEEK_END);
  uint64_t total = (uint64_t)std::ftell(fb) / (sizeof(float) * (uint64_t)D);
  if (num_lines == 0) {
    start_line = 0;
    num_lines = total;
  }
  if (start_line > total) start_line = total;
  if (start_line + num_lines > total) num_lines = total - start_line;
  values.resize(num_lines * (uint64_t)D);
  std::fseek(fb, (long)(start_line * sizeof(float) * (uint64_t)D), SEEK_SET);
  size_t got = num_lines ? std::fread(values.data(), sizeof(float) * (size_t)D, num_lines, fb) : 0;
  std::fclose(fb);
  if (got != num_lines) return KLSH_ERR_IO;

  id_offsets.assign(1, 0);
  ids.clear();
  if (id_format == 1) {
    const std::string cb = std::string(bin_path) + ".clust.bin";
    FILE* fi = std::fopen(cb.c_str(), "rb");
    if (!fi) return KLSH_ERR_IO;
    uint64_t rec = 0, loc = 0, cnt = 0;
    bool ok = true;
    while (loc < num_lines && std::fread(&cnt, sizeof cnt, 1, fi) == 1) {
      if (rec++ < start_line) {
        if (std::fseek(fi, (long)(cnt * sizeof(uint64_t)), SEEK_CUR) != 0) { ok = false; break; }
        continue;
      }
      const size_t at = ids.size();
      ids.resize(at + cnt);
      if (cnt && std::fread(ids.data() + at, sizeof(uint64_t), (size_t)cnt, fi) != cnt) { ok = false; break; }
      id_offsets.push_back(ids.size());
      ++loc;
    }
    std::fclose(fi);
    return (ok && loc == num_lines) ? KLSH_OK : KLSH_ERR_IO;
  }
  std::string clust = std::string(bin_path) + ".clust";
  FILE* ft = std::fopen(clust.c_str(), "r");
  if (!ft) return KLSH_ERR_IO;
  char* line = nullptr;
  size_t cap = 0;
  uint64_t lineno = 0, loc = 0;
  while (loc < num_lines && getline(&line, &cap, ft) >= 0) {
    if (lineno++ < start_line) continue;
    char* end;
    const char* p = line;
    uint64_t cnt = (uint64_t)std::strtol(p, &end, 10);
    for (uint64_t t = 0; t < cnt && p != end; ++t) {
      p = end;
      ids.push_back((uint64_t)std::strtol(p, &end, 10));
    }
    id_offsets.push_back(ids.size());
    ++loc;
  }
  std::free(line);
  std::fclose(ft);
  if (loc != num_lines) return KLSH_ERR_IO;
  return KLSH_OK;
}
