// car_precompute.cu — the CAR prior's pre-computation on the GPU.
//
// What it replaces in the reference (all host numpy / scipy, run once per input set by
// bin/splotch_generate_input_files through generate_dictionary):
//   /root/reference/splotch/utils.py:749-759  per tissue  eigvalsh(D^-1/2 W D^-1/2)  of the dense n x n matrix
//   /root/reference/splotch/utils.py:629-670  eigs_square_chunking: tissues with >= 10 000 spots are cut into
//                                             sqrt(chunksize)-wide coordinate squares, each solved on its own
//                                             (degrees taken INSIDE the chunk), eigenvalues concatenated
//   /root/reference/splotch/utils.py:592-614  sparse_to_csr_indptr: U, V of the symmetric adjacency (1-based)
//   /root/reference/splotch/utils.py:739      D_sparse = row sums of W
// Here: the dense matrices are assembled by a kernel and solved by cuSOLVER's syevd (a library eigensolver is the
// right tool for a dense symmetric spectrum; loaded lazily with dlopen so that the sampling library itself has no
// link-time dependency on it), identical problems are solved ONCE (c3: 24 identical Visium arrays -> one 4 992^2
// solve; the chunks of a regular grid repeat as well), the CSR arrays are built by counting kernels, and
// `exact_max_n` lets a caller with HBM to spare solve tissues beyond the reference's 10 000-spot limit exactly
// instead of by chunks.
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/csplotch.h"

extern "C" int csplotch_internal_fail(int code, const char *msg);
static int cfail(int code, const std::string &m) { return csplotch_internal_fail(code, m.c_str()); }
#define CAR_TRY(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess) return cfail(CSPLOTCH_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

namespace {

struct Solver {
  void *lib = nullptr;
  cusolverDnHandle_t h = nullptr;
  cusolverStatus_t (*create)(cusolverDnHandle_t *) = nullptr;
  cusolverStatus_t (*destroy)(cusolverDnHandle_t) = nullptr;
  cusolverStatus_t (*bufsize)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, const double *, int,
                              const double *, int *) = nullptr;
  cusolverStatus_t (*syevd)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, double *, int, double *, double *,
                            int, int *) = nullptr;
  bool open(std::string &err) {
    const char *names[] = {"libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11",
                           "/usr/local/cuda/lib64/libcusolver.so"};
    for (const char *n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (lib) break;
    }
    if (!lib) { err = "cuSOLVER (libcusolver.so.11) not found: needed for csplotch_car_precompute only"; return false; }
    create = (decltype(create))dlsym(lib, "cusolverDnCreate");
    destroy = (decltype(destroy))dlsym(lib, "cusolverDnDestroy");
    bufsize = (decltype(bufsize))dlsym(lib, "cusolverDnDsyevd_bufferSize");
    syevd = (decltype(syevd))dlsym(lib, "cusolverDnDsyevd");
    if (!create || !destroy || !bufsize || !syevd) { err = "cuSOLVER symbols missing"; return false; }
    if (create(&h) != CUSOLVER_STATUS_SUCCESS) { err = "cusolverDnCreate failed"; return false; }
    return true;
  }
  ~Solver() {
    if (h && destroy) destroy(h);
    if (lib) dlclose(lib);
  }
};

// one dense eigenproblem: n nodes, local pairs (a < b), degrees inside the problem
struct Problem {
  int n = 0;
  std::vector<int> pairs;   // 2 per pair, local 0-based
  std::vector<double> eig;  // filled by the solve
  int copies = 0;           // how many tissues / chunks share it
};

__global__ void k_fill_sym(double *A, int n, const int *__restrict__ pairs, int n_pairs, const double *__restrict__ isd) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_pairs; e += gridDim.x * blockDim.x) {
    const int a = pairs[2 * e], b = pairs[2 * e + 1];
    const double v = isd[a] * isd[b];
    A[(size_t)a * n + b] = v;
    A[(size_t)b * n + a] = v;
  }
}
__global__ void k_degree(const int *__restrict__ pairs, int n_pairs, int *deg) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_pairs; e += gridDim.x * blockDim.x) {
    atomicAdd(deg + pairs[2 * e], 1);
    atomicAdd(deg + pairs[2 * e + 1], 1);
  }
}
__global__ void k_csr_fill(const int *__restrict__ pairs, int n_pairs, const int *__restrict__ rowptr, int *cursor, int *col) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_pairs; e += gridDim.x * blockDim.x) {
    const int a = pairs[2 * e], b = pairs[2 * e + 1];
    col[rowptr[a] + atomicAdd(cursor + a, 1)] = b;
    col[rowptr[b] + atomicAdd(cursor + b, 1)] = a;
  }
}
__global__ void k_csr_sort_rows(const int *__restrict__ rowptr, int N, int *col) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const int r0 = rowptr[i], r1 = rowptr[i + 1];
    for (int a = r0 + 1; a < r1; ++a) {   // rows hold a handful of neighbours: insertion sort
      const int v = col[a];
      int b = a - 1;
      while (b >= r0 && col[b] > v) { col[b + 1] = col[b]; --b; }
      col[b + 1] = v;
    }
  }
}

struct DevMem {
  std::vector<void *> p;
  ~DevMem() { for (void *q : p) cudaFree(q); }
  template <class T> cudaError_t get(T **out, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(1, n) * sizeof(T));
    if (e == cudaSuccess) p.push_back(q);
    *out = (T *)q;
    return e;
  }
};

int solve(Solver &sv, Problem &pb) {
  const int n = pb.n, np = (int)pb.pairs.size() / 2;
  pb.eig.assign((size_t)n, 0.0);
  if (n == 0) return CSPLOTCH_OK;
  std::vector<int> deg((size_t)n, 0);
  for (int e = 0; e < np; ++e) { ++deg[pb.pairs[2 * e]]; ++deg[pb.pairs[2 * e + 1]]; }
  std::vector<double> isd((size_t)n);
  for (int i = 0; i < n; ++i) {
    if (deg[i] == 0) return cfail(CSPLOTCH_E_INVALID, "CAR pre-compute: a spot has no neighbour inside its tissue (chunk): D^-1/2 is undefined");
    isd[i] = 1.0 / std::sqrt((double)deg[i]);
  }
  DevMem dm;
  double *A = nullptr, *W = nullptr, *work = nullptr, *d_isd = nullptr;
  int *d_pairs = nullptr, *d_info = nullptr;
  if (dm.get(&A, (size_t)n * n) != cudaSuccess) { cudaGetLastError(); return cfail(CSPLOTCH_E_NOMEM, "CAR pre-compute: the dense n x n matrix does not fit in device memory; lower exact_max_n (chunking)"); }
  CAR_TRY(dm.get(&W, (size_t)n));
  CAR_TRY(dm.get(&d_isd, (size_t)n));
  CAR_TRY(dm.get(&d_pairs, (size_t)2 * np));
  CAR_TRY(dm.get(&d_info, 1));
  CAR_TRY(cudaMemset(A, 0, (size_t)n * n * sizeof(double)));
  CAR_TRY(cudaMemcpy(d_isd, isd.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  if (np) {
    CAR_TRY(cudaMemcpy(d_pairs, pb.pairs.data(), (size_t)2 * np * sizeof(int), cudaMemcpyHostToDevice));
    k_fill_sym<<<std::min(1184, (np + 255) / 256), 256>>>(A, n, d_pairs, np, d_isd);
    CAR_TRY(cudaGetLastError());
  }
  int lwork = 0;
  if (sv.bufsize(sv.h, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, W, &lwork) != CUSOLVER_STATUS_SUCCESS)
    return cfail(CSPLOTCH_E_CUDA, "cusolverDnDsyevd_bufferSize failed");
  if (dm.get(&work, (size_t)std::max(1, lwork)) != cudaSuccess) { cudaGetLastError(); return cfail(CSPLOTCH_E_NOMEM, "CAR pre-compute: eigensolver workspace does not fit in device memory"); }
  if (sv.syevd(sv.h, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, W, work, lwork, d_info) != CUSOLVER_STATUS_SUCCESS)
    return cfail(CSPLOTCH_E_CUDA, "cusolverDnDsyevd failed");
  int info = 0;
  CAR_TRY(cudaMemcpy(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost));
  if (info != 0) return cfail(CSPLOTCH_E_CUDA, "cusolverDnDsyevd did not converge (info = " + std::to_string(info) + ")");
  CAR_TRY(cudaMemcpy(pb.eig.data(), W, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  return CSPLOTCH_OK;
}

}  // namespace

extern "C" int32_t csplotch_car_precompute(int32_t N_tissues, const int32_t *N_spots, int32_t W_n, const int32_t *W_sparse,
                                           const int32_t *coords, int32_t chunk_spots, int32_t exact_max_n,
                                           double *eig_values, int32_t *U, int32_t *V, double *D_sparse,
                                           int32_t *n_solves) {
  if (N_tissues < 1 || !N_spots || W_n < 0 || (W_n > 0 && !W_sparse)) return cfail(CSPLOTCH_E_INVALID, "bad argument");
  if (chunk_spots <= 0) chunk_spots = 10000;      // eigs_square_chunking(coords, 10000), utils.py:757
  if (exact_max_n <= 0) exact_max_n = 9999;       // `if W_tissue.shape[0] < 10000`, utils.py:750
  if (exact_max_n > 40000) return cfail(CSPLOTCH_E_UNSUPPORTED, "exact_max_n above 40 000 needs the 64-bit eigensolver interface (not built)");
  std::vector<long long> off((size_t)N_tissues + 1, 0);
  for (int t = 0; t < N_tissues; ++t) {
    if (N_spots[t] < 1) return cfail(CSPLOTCH_E_INVALID, "N_spots must be >= 1");
    off[t + 1] = off[t] + N_spots[t];
  }
  if (off[N_tissues] > (1ll << 30)) return cfail(CSPLOTCH_E_INVALID, "too many spots");
  const int N = (int)off[N_tissues];
  std::vector<int> tissue_of((size_t)N);
  for (int t = 0; t < N_tissues; ++t)
    for (long long j = off[t]; j < off[t + 1]; ++j) tissue_of[j] = t;
  // pairs, 0-based; the adjacency is block diagonal over tissues (block_diag(W_list), utils.py:737)
  std::vector<int> pairs((size_t)2 * W_n);
  std::vector<std::vector<int>> tp((size_t)N_tissues);   // local pairs of every tissue
  for (int e = 0; e < W_n; ++e) {
    const int a = W_sparse[2 * e] - 1, b = W_sparse[2 * e + 1] - 1;
    if (a < 0 || b < 0 || a >= N || b >= N || a == b) return cfail(CSPLOTCH_E_INVALID, "W_sparse out of range");
    if (tissue_of[a] != tissue_of[b]) return cfail(CSPLOTCH_E_INVALID, "W_sparse links spots of two tissues");
    pairs[2 * e] = a; pairs[2 * e + 1] = b;
    const int t = tissue_of[a];
    tp[t].push_back(std::min(a, b) - (int)off[t]);
    tp[t].push_back(std::max(a, b) - (int)off[t]);
  }
  // ---- CSR of the symmetric adjacency + degrees, on the device ------------------------------------------
  {
    DevMem dm;
    int *d_pairs = nullptr, *d_deg = nullptr, *d_rowptr = nullptr, *d_cursor = nullptr, *d_col = nullptr;
    CAR_TRY(dm.get(&d_pairs, (size_t)2 * W_n));
    CAR_TRY(dm.get(&d_deg, (size_t)N));
    CAR_TRY(dm.get(&d_rowptr, (size_t)N + 1));
    CAR_TRY(dm.get(&d_cursor, (size_t)N));
    CAR_TRY(dm.get(&d_col, (size_t)2 * W_n));
    CAR_TRY(cudaMemcpy(d_pairs, pairs.data(), (size_t)2 * W_n * sizeof(int), cudaMemcpyHostToDevice));
    CAR_TRY(cudaMemset(d_deg, 0, (size_t)N * sizeof(int)));
    CAR_TRY(cudaMemset(d_cursor, 0, (size_t)N * sizeof(int)));
    const int grid = std::max(1, std::min(1184, (W_n + 255) / 256));
    if (W_n) k_degree<<<grid, 256>>>(d_pairs, W_n, d_deg);
    std::vector<int> deg((size_t)N), rowptr((size_t)N + 1, 0);
    CAR_TRY(cudaMemcpy(deg.data(), d_deg, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < N; ++i) rowptr[i + 1] = rowptr[i] + deg[i];
    CAR_TRY(cudaMemcpy(d_rowptr, rowptr.data(), ((size_t)N + 1) * sizeof(int), cudaMemcpyHostToDevice));
    if (W_n) {
      k_csr_fill<<<grid, 256>>>(d_pairs, W_n, d_rowptr, d_cursor, d_col);
      k_csr_sort_rows<<<std::max(1, std::min(1184, (N + 255) / 256)), 256>>>(d_rowptr, N, d_col);
    }
    CAR_TRY(cudaGetLastError());
    if (D_sparse)
      for (int i = 0; i < N; ++i) D_sparse[i] = (double)deg[i];
    if (U)
      for (int i = 0; i <= N; ++i) U[i] = rowptr[i] + 1;
    if (V && W_n) {
      std::vector<int> col((size_t)2 * W_n);
      CAR_TRY(cudaMemcpy(col.data(), d_col, (size_t)2 * W_n * sizeof(int), cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < col.size(); ++i) V[i] = col[i] + 1;
    }
  }
  if (!eig_values) {
    if (n_solves) *n_solves = 0;
    return CSPLOTCH_OK;
  }
  // ---- the eigenproblems: one per tissue, or one per coordinate chunk of a tissue beyond exact_max_n ---------
  std::vector<Problem> probs;
  std::map<std::string, int> seen;   // serialised (n, pairs) -> index into probs
  auto add_problem = [&](int n, std::vector<int> &pr) {
    // canonical order so that identical tissues hash alike whatever the order of their pairs
    const size_t np = pr.size() / 2;
    std::vector<std::pair<int, int>> ps(np);
    for (size_t e = 0; e < np; ++e) ps[e] = {pr[2 * e], pr[2 * e + 1]};
    std::sort(ps.begin(), ps.end());
    std::string key((size_t)(2 * np + 1) * sizeof(int), '\0');
    int *kp = reinterpret_cast<int *>(&key[0]);
    kp[0] = n;
    for (size_t e = 0; e < np; ++e) { kp[1 + 2 * e] = ps[e].first; kp[2 + 2 * e] = ps[e].second; }
    auto it = seen.find(key);
    if (it != seen.end()) { ++probs[it->second].copies; return; }
    Problem p;
    p.n = n;
    p.pairs.assign(kp + 1, kp + 1 + 2 * np);
    p.copies = 1;
    seen.emplace(std::move(key), (int)probs.size());
    probs.push_back(std::move(p));
  };
  for (int t = 0; t < N_tissues; ++t) {
    const int n = N_spots[t];
    if (n <= exact_max_n) { add_problem(n, tp[t]); continue; }
    if (!coords) return cfail(CSPLOTCH_E_INVALID, "tissue " + std::to_string(t) + " has more than exact_max_n spots: integer coordinates are needed for the chunked solution");
    // eigs_square_chunking (utils.py:629-670): squares of wc = rint(sqrt(chunksize)) coordinate units, origin 0
    const int wc = (int)std::lrint(std::sqrt((double)chunk_spots));
    const int *cx = coords + 2 * off[t];
    int xdim = 0, ydim = 0;
    for (int j = 0; j < n; ++j) {
      if (cx[2 * j] < 0 || cx[2 * j + 1] < 0) return cfail(CSPLOTCH_E_INVALID, "coordinates must be >= 0");
      xdim = std::max(xdim, cx[2 * j]); ydim = std::max(ydim, cx[2 * j + 1]);
    }
    // the reference's loops `range(0, xdim, wc)` skip a last square that starts exactly at the maximum coordinate
    // and then fail its length assert; cover the whole range instead
    const int nx = xdim / wc + 1, ny = ydim / wc + 1;
    std::vector<int> chunk_of((size_t)n), local_id((size_t)n), count((size_t)nx * ny, 0);
    for (int j = 0; j < n; ++j) {
      chunk_of[j] = (cx[2 * j] / wc) * ny + cx[2 * j + 1] / wc;
      local_id[j] = count[chunk_of[j]]++;
    }
    std::vector<std::vector<int>> cp((size_t)nx * ny);
    for (size_t e = 0; e < tp[t].size() / 2; ++e) {
      const int a = tp[t][2 * e], b = tp[t][2 * e + 1];
      if (chunk_of[a] != chunk_of[b]) continue;     // the chunk's own adjacency: links that leave it are dropped
      cp[chunk_of[a]].push_back(std::min(local_id[a], local_id[b]));
      cp[chunk_of[a]].push_back(std::max(local_id[a], local_id[b]));
    }
    for (int c = 0; c < nx * ny; ++c)
      if (count[c] > 0) {
        if (count[c] > 40000) return cfail(CSPLOTCH_E_UNSUPPORTED, "a coordinate chunk holds more than 40 000 spots");
        add_problem(count[c], cp[c]);
      }
  }
  Solver sv;
  std::string err;
  if (!sv.open(err)) return cfail(CSPLOTCH_E_UNSUPPORTED, err);
  size_t pos = 0;
  for (Problem &p : probs) {
    const int rc = solve(sv, p);
    if (rc) return rc;
    for (int c = 0; c < p.copies; ++c) {
      if (pos + p.eig.size() > (size_t)N) return cfail(CSPLOTCH_E_INVALID, "internal: eigenvalue count mismatch");
      std::memcpy(eig_values + pos, p.eig.data(), p.eig.size() * sizeof(double));
      pos += p.eig.size();
    }
  }
  if (pos != (size_t)N) return cfail(CSPLOTCH_E_INVALID, "internal: eigenvalue count mismatch");
  std::sort(eig_values, eig_values + N);     // numpy.sort(numpy.concatenate(eigval_list)), utils.py:759
  if (n_solves) *n_solves = (int)probs.size();
  return CSPLOTCH_OK;
}
