// Ingestion fast path: the per-gene keys of many `data_<g>.R` files, in parallel, natively.
//
// Every data_<g>.R of splotch_generate_input_files repeats the gene-shared covariates (E, W_sparse, U, V,
// eig_values, ...: tens of MB of text at 10^5 spots) and differs only in `counts` and, for the compositional
// model, `beta_prior_mean` / `beta_prior_std` (/root/reference/bin/splotch_generate_input_files:329-346).
// Those keys are assigned after generate_dictionary() filled the dict, so to_rdump
// (/root/reference/splotch/utils.py:113-125) writes them LAST.  The reference parses each whole file token by
// token in Python (utils.py:78-111), once per chain process.  Here a pool of host threads reads only the tail
// of each file (growing the window until the keys are found, so any key order still works), and parses the
// numbers straight into the caller's [G][N] / [G][K] arrays.  Host code only: no CUDA in this file.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/csplotch.h"

extern "C" int csplotch_internal_fail(int code, const char *msg);  // csplotch.cu: sets csplotch_last_error()

namespace {

// position of "<key> <-" at the start of a line inside [buf, buf+n), searching backwards; -1 if absent.
// `complete`: buf starts at the beginning of the file (then offset 0 counts as a line start).
long find_key(const char *buf, long n, const char *key, bool complete) {
  const long kl = (long)strlen(key);
  for (long i = n - kl - 2; i >= 0; --i) {
    if (buf[i] != key[0] || memcmp(buf + i, key, kl) != 0) continue;
    long j = i + kl;
    while (j < n && (buf[j] == ' ' || buf[j] == '\t')) ++j;
    if (j + 1 >= n || buf[j] != '<' || buf[j + 1] != '-') continue;
    long b = i - 1;
    while (b >= 0 && (buf[b] == ' ' || buf[b] == '\t')) --b;
    if (b >= 0 ? buf[b] == '\n' : complete) return j + 2;  // offset just after "<-"
  }
  return -1;
}

// the number list of the value that starts at p: `c(...)`, `structure(c(...), .Dim=...)` or a bare scalar.
// Returns [begin, end) of the comma-separated numbers.
bool value_span(const char *buf, long n, long p, long *begin, long *end) {
  while (p < n && isspace((unsigned char)buf[p])) ++p;
  if (p < n && buf[p] == 's') {  // structure(
    while (p < n && buf[p] != '(') ++p;
    ++p;
    while (p < n && isspace((unsigned char)buf[p])) ++p;
  }
  if (p < n && buf[p] == 'c') {
    ++p;
    while (p < n && isspace((unsigned char)buf[p])) ++p;
    if (p >= n || buf[p] != '(') return false;
    ++p;
    long q = p;
    while (q < n && buf[q] != ')') ++q;
    if (q >= n) return false;
    *begin = p;
    *end = q;
    return true;
  }
  long q = p;
  while (q < n && buf[q] != '\n') ++q;
  *begin = p;
  *end = q;
  return true;
}

// integers as to_rdump writes them ("3", also "3.0" / "3L" from other writers); -1 on a malformed token
long parse_ints(const char *s, const char *e, int32_t *out, long cap) {
  long cnt = 0;
  while (s < e) {
    while (s < e && (isspace((unsigned char)*s) || *s == ',')) ++s;
    if (s >= e) break;
    bool neg = false;
    if (*s == '-' || *s == '+') neg = *s++ == '-';
    if (s >= e || *s < '0' || *s > '9') return -1;
    long long v = 0;
    while (s < e && *s >= '0' && *s <= '9') {
      v = v * 10 + (*s++ - '0');
      if (v > 2147483647LL) return -1;
    }
    if (s < e && *s == '.') {  // "12.0": only an integral value is a count
      ++s;
      while (s < e && *s >= '0' && *s <= '9')
        if (*s++ != '0') return -1;
    }
    if (s < e && *s == 'L') ++s;
    if (s < e && !(isspace((unsigned char)*s) || *s == ',')) return -1;
    if (cnt < cap) out[cnt] = (int32_t)(neg ? -v : v);
    ++cnt;
  }
  return cnt;
}

long parse_doubles(const char *s, const char *e, double *out, long cap) {
  long cnt = 0;
  std::string tok;
  while (s < e) {
    while (s < e && (isspace((unsigned char)*s) || *s == ',')) ++s;
    if (s >= e) break;
    const char *t = s;
    while (t < e && !(isspace((unsigned char)*t) || *t == ',')) ++t;
    tok.assign(s, t - s);
    char *endp = nullptr;
    const double v = strtod(tok.c_str(), &endp);  // accepts nan / inf like float() in read_rdump
    if (endp == tok.c_str() || *endp != '\0') return -1;
    if (cnt < cap) out[cnt] = v;
    ++cnt;
    s = t;
  }
  return cnt;
}

struct Job {
  const char *const *paths;
  int32_t G, N, K;
  int32_t *counts;
  double *pm, *ps;
  std::atomic<int32_t> next{0};
  std::mutex mu;
  std::string err;
};

bool read_one(Job &J, int32_t g, std::string &err) {
  const char *path = J.paths[g];
  const int fd = open(path, O_RDONLY);
  if (fd < 0) {
    err = std::string(path) + ": " + strerror(errno);
    return false;
  }
  struct stat st;
  if (fstat(fd, &st) != 0) {
    err = std::string(path) + ": " + strerror(errno);
    close(fd);
    return false;
  }
  const long fsz = (long)st.st_size;
  const bool want_prior = J.pm != nullptr;
  // counts are mostly one or two digits + a comma; priors ~20 characters each
  long win = 4L * J.N + 64L * (want_prior ? J.K : 0) + 4096;
  std::vector<char> buf;
  bool ok = false;
  for (;;) {
    if (win > fsz) win = fsz;
    buf.resize(win);
    long got = 0;
    while (got < win) {
      const ssize_t r = pread(fd, buf.data() + got, win - got, fsz - win + got);
      if (r <= 0) break;
      got += r;
    }
    if (got != win) {
      err = std::string(path) + ": short read";
      break;
    }
    const bool complete = win == fsz;
    const long pc = find_key(buf.data(), win, "counts", complete);
    const long pm = want_prior ? find_key(buf.data(), win, "beta_prior_mean", complete) : 0;
    const long ps = want_prior ? find_key(buf.data(), win, "beta_prior_std", complete) : 0;
    if (pc < 0 || pm < 0 || ps < 0) {
      if (complete) {
        err = std::string(path) + ": key `" + (pc < 0 ? "counts" : (pm < 0 ? "beta_prior_mean" : "beta_prior_std")) +
              "` not found";
        break;
      }
      win *= 4;
      continue;
    }
    long b, e;
    if (!value_span(buf.data(), win, pc, &b, &e)) {
      err = std::string(path) + ": malformed `counts`";
      break;
    }
    const long nc = parse_ints(buf.data() + b, buf.data() + e, J.counts + (size_t)g * J.N, J.N);
    if (nc != J.N) {
      err = std::string(path) + (nc < 0 ? ": `counts` holds a non-integer token"
                                        : ": `counts` has " + std::to_string(nc) + " entries, expected " +
                                              std::to_string(J.N));
      break;
    }
    if (want_prior) {
      bool bad = false;
      for (int which = 0; which < 2 && !bad; ++which) {
        if (!value_span(buf.data(), win, which ? ps : pm, &b, &e)) {
          bad = true;
          break;
        }
        const long nk = parse_doubles(buf.data() + b, buf.data() + e, (which ? J.ps : J.pm) + (size_t)g * J.K, J.K);
        if (nk != J.K) bad = true;
      }
      if (bad) {
        err = std::string(path) + ": beta_prior_mean / beta_prior_std must hold N_celltypes = " + std::to_string(J.K) +
              " numbers";
        break;
      }
    }
    ok = true;
    break;
  }
  close(fd);
  return ok;
}

void worker(Job *J) {
  for (;;) {
    const int32_t g = J->next.fetch_add(1);
    if (g >= J->G) return;
    std::string err;
    if (!read_one(*J, g, err)) {
      std::lock_guard<std::mutex> lk(J->mu);
      if (J->err.empty()) J->err = err;
      J->next.store(J->G);  // stop handing out files
      return;
    }
  }
}

}  // namespace

extern "C" int32_t csplotch_rdump_read_genes(const char *const *paths, int32_t G, int32_t N, int32_t K,
                                             int32_t *counts, double *beta_prior_mean, double *beta_prior_std,
                                             int32_t n_threads) {
  if (G < 0 || N <= 0 || (G > 0 && (!paths || !counts)))
    return csplotch_internal_fail(CSPLOTCH_E_INVALID, "csplotch_rdump_read_genes: bad arguments");
  if ((beta_prior_mean == nullptr) != (beta_prior_std == nullptr) || (beta_prior_mean && K <= 0))
    return csplotch_internal_fail(CSPLOTCH_E_INVALID,
                                  "csplotch_rdump_read_genes: beta_prior_mean and beta_prior_std go together (K > 0)");
  Job J;
  J.paths = paths;
  J.G = G;
  J.N = N;
  J.K = K;
  J.counts = counts;
  J.pm = beta_prior_mean;
  J.ps = beta_prior_std;
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > G) nt = G > 0 ? G : 1;
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(worker, &J);
  worker(&J);
  for (auto &t : pool) t.join();
  if (!J.err.empty()) return csplotch_internal_fail(CSPLOTCH_E_INVALID, J.err.c_str());
  return CSPLOTCH_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Output fast path: `combined_<g>.csv` of many genes, in parallel, natively.
//
// What CmdStan's CSV writer plus the wrapper's grep / sed concatenation leave behind
// (/root/reference/bin/splotch:131,138-140): ONE header line (the one containing `lp__`), then one row per draw,
// chains one after the other, no `#` comment lines.  A row is 7 sampler columns + every write_array column
// (c2: 7.9 k numbers, c3: 3.6e5) and a gene has NUM_CHAINS x NUM_SAMPLES rows, so formatting dominates the CSV mode;
// a pool of host threads formats one file each.  Numbers: integral values print without a decimal point
// ("3", like treedepth__ / n_leapfrog__ / divergent__ in CmdStan's files); everything else in the shortest form
// that reads back to the same double (sig_figs = 0) or with `sig_figs` significant digits (CmdStan's `sig_figs=`,
// whose default is 6).  Each file is written to `<path>.tmp` and renamed.  Host code only.
#include <charconv>
#include <cstdio>

namespace {

struct CsvJob {
  const char *const *paths;
  int32_t n_files;
  const char *header;
  const double *lead, *values;
  int32_t n_lead, n_values, sig_figs;
  int64_t n_rows;
  std::atomic<int32_t> next{0};
  std::mutex mu;
  std::string err;
};

inline char *put_number(char *p, char *end, double v, int sig_figs) {
  if (v == std::floor(v) && std::fabs(v) < 1e15) {
    auto r = std::to_chars(p, end, (long long)v);
    return r.ptr;
  }
  auto r = sig_figs > 0 ? std::to_chars(p, end, v, std::chars_format::general, sig_figs) : std::to_chars(p, end, v);
  return r.ptr;
}

bool write_one(CsvJob &J, int32_t f, std::string &err) {
  const std::string path = J.paths[f];
  const std::string tmp = path + ".tmp";
  FILE *fp = fopen(tmp.c_str(), "wb");
  if (!fp) {
    err = tmp + ": " + strerror(errno);
    return false;
  }
  const size_t ncol = (size_t)J.n_lead + (size_t)J.n_values;
  std::vector<char> line(ncol * 32 + 16);  // a double needs at most 24 characters
  bool ok = fputs(J.header, fp) >= 0 && fputc('\n', fp) != EOF;
  for (int64_t r = 0; ok && r < J.n_rows; ++r) {
    char *p = line.data(), *end = line.data() + line.size();
    const double *a = J.lead + ((size_t)f * J.n_rows + r) * J.n_lead;
    const double *b = J.values + ((size_t)f * J.n_rows + r) * J.n_values;
    for (int32_t c = 0; c < J.n_lead; ++c) {
      p = put_number(p, end, a[c], J.sig_figs);
      *p++ = ',';
    }
    for (int32_t c = 0; c < J.n_values; ++c) {
      p = put_number(p, end, b[c], J.sig_figs);
      *p++ = ',';
    }
    p[-1] = '\n';
    ok = fwrite(line.data(), 1, (size_t)(p - line.data()), fp) == (size_t)(p - line.data());
  }
  if (fclose(fp) != 0) ok = false;
  if (ok && rename(tmp.c_str(), path.c_str()) != 0) ok = false;
  if (!ok) {
    err = path + ": " + strerror(errno);
    remove(tmp.c_str());
  }
  return ok;
}

void csv_worker(CsvJob *J) {
  for (;;) {
    const int32_t f = J->next.fetch_add(1);
    if (f >= J->n_files) return;
    std::string err;
    if (!write_one(*J, f, err)) {
      std::lock_guard<std::mutex> lk(J->mu);
      if (J->err.empty()) J->err = err;
      J->next.store(J->n_files);
      return;
    }
  }
}

}  // namespace

extern "C" int32_t csplotch_csv_write(const char *const *paths, int32_t n_files, const char *header,
                                      const double *lead, int32_t n_lead, const double *values, int32_t n_values,
                                      int64_t n_rows, int32_t sig_figs, int32_t n_threads) {
  if (n_files < 0 || n_rows < 0 || n_lead < 0 || n_values < 0 || n_lead + (int64_t)n_values < 1 || sig_figs < 0 ||
      sig_figs > 18 || !header || (n_files > 0 && (!paths || (n_rows > 0 && ((n_lead && !lead) || (n_values && !values))))))
    return csplotch_internal_fail(CSPLOTCH_E_INVALID, "csplotch_csv_write: bad arguments");
  CsvJob J;
  J.paths = paths;
  J.n_files = n_files;
  J.header = header;
  J.lead = lead;
  J.values = values;
  J.n_lead = n_lead;
  J.n_values = n_values;
  J.sig_figs = sig_figs;
  J.n_rows = n_rows;
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > n_files) nt = n_files > 0 ? n_files : 1;
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(csv_worker, &J);
  csv_worker(&J);
  for (auto &t : pool) t.join();
  if (!J.err.empty()) return csplotch_internal_fail(CSPLOTCH_E_INVALID, J.err.c_str());
  return CSPLOTCH_OK;
}
