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
