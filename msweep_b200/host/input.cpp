// input.cpp — see input.hpp.  Semantics follow the reference line by line; the implementation does not:
// the reference tokenises with getline/stringstream/stoul on one thread into a compressed bit vector,
// here every strand is read into memory once, cut at line boundaries into one slice per thread, scanned
// with a hand-rolled integer parser into (row, column) pairs, bucketed by row and merged as sorted lists.
#include "input.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <fstream>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#include <sstream>
#include <unordered_map>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace b200 {

Grouping read_grouping(const std::string &path, char delimiter, size_t column) {
  std::ifstream in(path);
  if (!in.good()) throw std::runtime_error("Could not read cluster indicators.");     // src/Reference.cpp:40-42
  Grouping g;
  std::unordered_map<std::string, uint32_t> ids;
  std::string line, field;
  while (std::getline(in, line)) {
    std::stringstream ss(line);
    size_t col = 0;
    bool found = false;
    while (std::getline(ss, field, delimiter)) {
      if (col == column) found = true;
      if (found && col == column) break;
      ++col;
    }
    {   // number of groupings = max number of columns on a line (include/Reference.hpp:75-80)
      size_t ncol = (size_t)std::count(line.begin(), line.end(), delimiter) + (line.empty() ? 0 : 1);
      g.n_groupings = std::max(g.n_groupings, ncol);
    }
    if (!found) continue;
    auto it = ids.find(field);
    if (it == ids.end()) {
      it = ids.emplace(field, (uint32_t)g.names.size()).first;
      g.names.push_back(field);
      g.sizes.push_back(0);
    }
    g.sizes[it->second] += 1;
    g.group_of_target.push_back(it->second);
  }
  if (g.group_of_target.empty()) throw std::runtime_error("The grouping contains 0 reference sequences");   // Reference.hpp:90-92
  return g;
}

namespace {

struct StageClock {
  const bool on = getenv("MSWB_PARSE_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char *what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "  [parse] %-18s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  }
};

struct Strand {
  uint64_t n_lines = 0;
  std::vector<uint64_t> row_ptr;       // rows 0 .. n_rows-1, ascending unique columns
  raw_vector<uint32_t> cols;
};

// Read-only view of a whole file: mmap when possible (no copy, pages faulted in by the tokeniser threads), else a
// plain read into memory (pipes, /dev/stdin).
struct FileView {
  const char *data = nullptr;
  size_t size = 0;
  void *map = nullptr;
  std::string fallback;
  FileView() = default;
  FileView(const FileView &) = delete;
  FileView &operator=(const FileView &) = delete;
  FileView(FileView &&o) noexcept { *this = std::move(o); }
  FileView &operator=(FileView &&o) noexcept {
    release();
    map = o.map; size = o.size; fallback = std::move(o.fallback);
    data = map ? (const char *)map : fallback.data();
    o.map = nullptr; o.data = nullptr; o.size = 0;
    return *this;
  }
  ~FileView() { release(); }
  void release() { if (map) munmap(map, size); map = nullptr; data = nullptr; size = 0; std::string().swap(fallback); }
  void open(const std::string &path) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("Could not open " + path);
    unsigned char magic[2] = {0, 0};
    const bool gz = ::pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    if (gz) {   // gzip (the reference reads compressed alignments transparently through cxxio/bxzstr): inflate into memory
      ::close(fd);
      gzFile f = gzopen(path.c_str(), "rb");
      if (!f) throw std::runtime_error("Could not open " + path);
      gzbuffer(f, 1 << 20);
      std::vector<char> chunk(1 << 22);
      for (;;) {
        const int got = gzread(f, chunk.data(), (unsigned)chunk.size());
        if (got < 0) { gzclose(f); throw std::runtime_error("Could not decompress " + path); }
        if (got == 0) break;
        fallback.append(chunk.data(), (size_t)got);
      }
      gzclose(f);
      data = fallback.data(); size = fallback.size();
      return;
    }
    struct stat st;
    if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
      void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m != MAP_FAILED) {
        madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL | MADV_WILLNEED);
        map = m; size = (size_t)st.st_size; data = (const char *)m;
        ::close(fd);
        return;
      }
    }
    ::close(fd);
    std::ifstream in(path, std::ios::binary);
    if (!in.good()) throw std::runtime_error("Could not open " + path);
    fallback.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
    data = fallback.data(); size = fallback.size();
  }
};

// std::stoul semantics on a space-delimited token: optional leading whitespace, then digits; junk after
// the digits is ignored; no digits at all is an error.
inline bool parse_token(const char *&p, const char *end, uint64_t &out) {
  while (p < end && (*p == '\t' || *p == '\r' || *p == '\v' || *p == '\f')) ++p;
  if (p < end && *p == '+') ++p;
  if (p >= end || *p < '0' || *p > '9') return false;
  uint64_t v = 0;
  while (p < end && *p >= '0' && *p <= '9') v = v * 10 + (uint64_t)(*p++ - '0');
  while (p < end && *p != ' ') ++p;     // trailing junk inside the token
  out = v;
  return true;
}

Strand parse_strand(const FileView &buf, uint64_t T, int n_threads) {
  const char *first_nl = (const char *)memchr(buf.data, '\n', buf.size);
  const size_t first_len = first_nl ? (size_t)(first_nl - buf.data) : buf.size;
  if (buf.size && memchr(buf.data, ',', first_len))
    throw std::runtime_error("compact (alignment-writer) pseudoalignments are not supported by this backend; "
                             "convert to Themisto plaintext");
  const char *base = buf.data;
  const size_t n = buf.size;
  std::vector<size_t> cut(n_threads + 1, n);
  cut[0] = 0;
  for (int t = 1; t < n_threads; ++t) {
    size_t p = n * (size_t)t / n_threads;
    const void *nl = p < n ? memchr(base + p, '\n', n - p) : nullptr;
    cut[t] = nl ? (size_t)((const char *)nl - base) + 1 : n;
  }
  for (int t = 1; t <= n_threads; ++t) cut[t] = std::max(cut[t], cut[t - 1]);

  StageClock clk;
  // Per thread: the columns of its lines back to back, plus (read id, first column, count) per line.  A target id
  // >= T belongs to a later row (flat bit index read_id*T + target, mSWEEP_alignment.hpp:64): those rare hits go to
  // a side list as explicit (row, column) pairs.
  struct Line { uint64_t read; uint64_t off; uint32_t n; };
  std::vector<std::unique_ptr<uint32_t[]>> cols(n_threads);     // uninitialised; a token takes at least two bytes of text
  std::vector<std::vector<Line>> line_tab(n_threads);
  std::vector<std::vector<std::pair<uint64_t, uint32_t>>> spill(n_threads);
  std::vector<uint64_t> lines(n_threads, 0), bad_line(n_threads, 0);
  std::vector<std::string> bad_text(n_threads);
#pragma omp parallel for schedule(static, 1) num_threads(n_threads)
  for (int t = 0; t < n_threads; ++t) {
    const char *p = base + cut[t], *end = base + cut[t + 1];
    cols[t].reset(new uint32_t[(size_t)(end - p) / 2 + 2]);
    uint32_t *const out0 = cols[t].get();
    uint32_t *w = out0;
    line_tab[t].reserve((size_t)(end - p) / 128 + 16);
    while (p < end) {
      const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
      if (!eol) eol = end;
      ++lines[t];
      uint32_t *const line_w = w;
      const size_t spill0 = spill[t].size();
      // Fast path: decimal numbers separated by single spaces, which is what Themisto writes.  Anything else
      // (signs, tabs, junk after the digits) rewinds the line and goes through the stoul-compatible tokeniser below.
      const char *q = p;
      uint64_t read_id = 0;
      unsigned d;
      bool clean = q < eol && (d = (unsigned)(*q - '0')) <= 9;
      if (clean) {
        do { read_id = read_id * 10 + d; ++q; } while (q < eol && (d = (unsigned)(*q - '0')) <= 9);
        while (q < eol) {
          if (*q != ' ') { clean = false; break; }
          if (++q >= eol) break;                              // trailing delimiter: getline yields no further token
          if ((d = (unsigned)(*q - '0')) > 9) { clean = false; break; }
          uint64_t v = 0;
          do { v = v * 10 + d; ++q; } while (q < eol && (d = (unsigned)(*q - '0')) <= 9);
          if (v < T) *w++ = (uint32_t)v; else spill[t].emplace_back(read_id + v / T, (uint32_t)(v % T));
        }
      }
      bool ok = clean;
      if (!clean) {
        w = line_w;
        spill[t].resize(spill0);
        q = p;
        uint64_t tgt = 0;
        ok = parse_token(q, eol, read_id);
        while (ok && q < eol) {
          ++q;                                       // the single ' ' delimiter
          if (q >= eol) break;
          ok = parse_token(q, eol, tgt);
          if (ok) { if (tgt < T) *w++ = (uint32_t)tgt; else spill[t].emplace_back(read_id + tgt / T, (uint32_t)(tgt % T)); }
        }
      }
      if (!ok && !bad_line[t]) { bad_line[t] = lines[t]; bad_text[t].assign(p, eol); }
      if (ok) line_tab[t].push_back(Line{read_id, (uint64_t)(line_w - out0), (uint32_t)(w - line_w)}); else w = line_w;
      p = eol + 1;
    }
  }
  clk.lap("tokenise");
  uint64_t before = 0;
  for (int t = 0; t < n_threads; ++t) {
    if (bad_line[t]) throw std::runtime_error("File format not supported on line " + std::to_string(before + bad_line[t]) +
                                              " with content: " + bad_text[t]);
    before += lines[t];
  }
  Strand s;
  s.n_lines = before;
  // bucket by row, one atomic per LINE (not per hit); rows at or beyond the line count are never visited by collapse()
  const uint64_t R = s.n_lines;
  s.row_ptr.assign(R + 1, 0);
#pragma omp parallel for schedule(static, 1) num_threads(n_threads)
  for (int t = 0; t < n_threads; ++t) {
    for (const Line &l : line_tab[t]) if (l.read < R && l.n) __atomic_fetch_add(&s.row_ptr[l.read + 1], (uint64_t)l.n, __ATOMIC_RELAXED);
    for (const auto &rc : spill[t]) if (rc.first < R) __atomic_fetch_add(&s.row_ptr[rc.first + 1], 1, __ATOMIC_RELAXED);
  }
  for (uint64_t r = 0; r < R; ++r) s.row_ptr[r + 1] += s.row_ptr[r];
  s.cols.resize(s.row_ptr[R]);
  std::vector<uint64_t> fill(s.row_ptr.begin(), s.row_ptr.end() - 1);
#pragma omp parallel for schedule(static, 1) num_threads(n_threads)
  for (int t = 0; t < n_threads; ++t) {
    for (const Line &l : line_tab[t]) {
      if (l.read >= R || !l.n) continue;
      const uint64_t pos = __atomic_fetch_add(&fill[l.read], (uint64_t)l.n, __ATOMIC_RELAXED);
      std::memcpy(s.cols.data() + pos, cols[t].get() + l.off, (size_t)l.n * sizeof(uint32_t));
    }
    for (const auto &rc : spill[t]) if (rc.first < R) s.cols[__atomic_fetch_add(&fill[rc.first], 1, __ATOMIC_RELAXED)] = rc.second;
    cols[t].reset();
    std::vector<Line>().swap(line_tab[t]);
  }
  clk.lap("bucket by row");
  // ascending + unique inside each row (a bit can only be set once)
  std::vector<uint64_t> len(R);
#pragma omp parallel for schedule(static) num_threads(n_threads)
  for (uint64_t r = 0; r < R; ++r) {
    uint32_t *a = s.cols.data() + s.row_ptr[r], *b = s.cols.data() + s.row_ptr[r + 1];
    if (!std::is_sorted(a, b)) std::sort(a, b);
    len[r] = (uint64_t)(std::unique(a, b) - a);
  }
  uint64_t w = 0;
  for (uint64_t r = 0; r < R; ++r) {          // compact away duplicates
    const uint64_t a = s.row_ptr[r];
    if (w != a) std::memmove(s.cols.data() + w, s.cols.data() + a, len[r] * sizeof(uint32_t));
    s.row_ptr[r] = w;
    w += len[r];
  }
  s.row_ptr[R] = w;
  s.cols.resize(w);
  clk.lap("sort + compact");
  return s;
}

} // namespace

ReadTable read_themisto(const std::vector<std::string> &paths, uint64_t n_targets, const std::string &merge_mode, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  ReadTable out;
  out.n_targets = n_targets;
  // the files are read from disk concurrently (one reader thread each); tokenising uses all threads per strand
  std::vector<FileView> bufs(paths.size());
  for (size_t i = 0; i < paths.size(); ++i) bufs[i].open(paths[i]);
  Strand acc;
  for (size_t i = 0; i < paths.size(); ++i) {
    Strand s = parse_strand(bufs[i], n_targets, n_threads);
    bufs[i].release();
    if (i == 0) { acc = std::move(s); continue; }
    const bool isect = merge_mode == "intersection";
    if (!isect && merge_mode != "union")
      throw std::runtime_error("Unrecognized option `" + merge_mode + "` for --themisto-mode");   // mSWEEP_alignment.hpp:130-132
    // bit_and / bit_or of the two strands; n_queries follows the LAST strand (:121)
    const uint64_t R = s.n_lines, Ra = acc.n_lines;
    Strand m;
    m.n_lines = R;
    m.row_ptr.assign(R + 1, 0);
    std::vector<uint64_t> len(R, 0);
    auto row = [](const Strand &x, uint64_t r, const uint32_t *&a, const uint32_t *&b) {
      if (r < x.n_lines) { a = x.cols.data() + x.row_ptr[r]; b = x.cols.data() + x.row_ptr[r + 1]; } else { a = b = nullptr; }
    };
    if (isect) {
      // the intersection of row r fits inside the first strand's row r: one pass writes it there, a second one packs
#pragma omp parallel for schedule(static) num_threads(n_threads)
      for (uint64_t r = 0; r < R; ++r) {
        const uint32_t *a0, *a1, *b0, *b1;
        row(acc, r, a0, a1); row(s, r, b0, b1);
        uint32_t *o = const_cast<uint32_t *>(a0), *const o0 = o;
        while (a0 < a1 && b0 < b1) { if (*a0 < *b0) ++a0; else if (*b0 < *a0) ++b0; else { *o++ = *a0; ++a0; ++b0; } }
        len[r] = (uint64_t)(o - o0);
      }
      for (uint64_t r = 0; r < R; ++r) m.row_ptr[r + 1] = m.row_ptr[r] + len[r];
      m.cols.resize(m.row_ptr[R]);
#pragma omp parallel for schedule(static) num_threads(n_threads)
      for (uint64_t r = 0; r < R; ++r)
        if (len[r]) std::memcpy(m.cols.data() + m.row_ptr[r], acc.cols.data() + acc.row_ptr[r], len[r] * sizeof(uint32_t));
    } else {
#pragma omp parallel for schedule(static) num_threads(n_threads)
      for (uint64_t r = 0; r < R; ++r) {
        const uint32_t *a0, *a1, *b0, *b1;
        row(acc, r, a0, a1); row(s, r, b0, b1);
        uint64_t c = 0;
        while (a0 < a1 && b0 < b1) { if (*a0 < *b0) ++a0; else if (*b0 < *a0) ++b0; else { ++a0; ++b0; } ++c; }
        len[r] = c + (uint64_t)(a1 - a0) + (uint64_t)(b1 - b0);
      }
      for (uint64_t r = 0; r < R; ++r) m.row_ptr[r + 1] = m.row_ptr[r] + len[r];
      m.cols.resize(m.row_ptr[R]);
#pragma omp parallel for schedule(static) num_threads(n_threads)
      for (uint64_t r = 0; r < R; ++r) {
        const uint32_t *a0, *a1, *b0, *b1;
        row(acc, r, a0, a1); row(s, r, b0, b1);
        std::set_union(a0, a1, b0, b1, m.cols.data() + m.row_ptr[r]);
      }
    }
    (void)Ra;
    acc = std::move(m);
  }
  out.n_reads = acc.n_lines;
  out.row_ptr = std::move(acc.row_ptr);
  out.targets = std::move(acc.cols);
  if (out.row_ptr.empty()) out.row_ptr.assign(1, 0);
  return out;
}

} // namespace b200
