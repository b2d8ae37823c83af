// csplotch.cu — kernels + C ABI (include/csplotch.h) of the B200-native cSplotch sampling path.
//
// Kernels (sm_100a):
//   k_nuts_advance   persistent; one CTA per gene-chain pulled from an atomic queue; runs the whole
//                    Stan-compatible adaptive NUTS of nuts_core.h with the fused fp64
//                    log-density+gradient / leapfrog of device_ops.cuh.
//   k_lpgrad         stand-alone log_prob_grad (tier-1 parity, roofline B_grad)
//   k_leapfrog_fixed fixed-step trajectories (tier-2 parity)
//   small helpers    layout permutation, count repacking, per-gene constants, summaries
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/csplotch.h"
#include "device_ops.cuh"
#include "warp_ops.cuh"

using namespace csp;

// ------------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
// host-only translation units of this library (ingest.cpp) report errors through the same thread-local message
extern "C" int csplotch_internal_fail(int code, const char *msg) { return fail(code, msg ? msg : ""); }
#define CU_TRY(expr)                                                                              \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return fail(CSPLOTCH_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));         \
  } while (0)

struct DevBuf {
  std::vector<void *> ptrs;
  ~DevBuf() {
    for (void *p : ptrs) cudaFree(p);
  }
  template <class T>
  cudaError_t alloc(T **out, size_t n) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) ptrs.push_back(p);
    *out = (T *)p;
    return e;
  }
  template <class T>
  cudaError_t upload(T **out, const std::vector<T> &h) {
    cudaError_t e = alloc(out, h.size());
    if (e != cudaSuccess) return e;
    if (!h.empty()) e = cudaMemcpy(*out, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
  }
};

// ------------------------------------------------------------------------------------------------
// opaque handles
// ------------------------------------------------------------------------------------------------
struct csplotch_model {
  DevModel dm;
  DevBuf buf;
  int device = 0;
  int T = 0;
  std::vector<int> imap;  // internal -> Stan
  std::vector<int> perm;  // internal spot -> original spot
  int kt = 1;             // template instantiation (max K)
  size_t smem_bytes = 0;
  int n_outputs = 0;
  // thread-block cluster variants of the design (DevModel::cl): tile partition + the ELL table with halo positions
  struct ClusterVariant { int cl; int t0[kMaxCluster + 1]; unsigned short *ell_dev; };
  std::vector<ClusterVariant> clv;
  int max_cl = 1;        // the largest cluster size worth using on this design (by its tile count)
  int forced_cl = 0;     // CSPLOTCH_CLUSTER at model creation (tests / A-B measurements): always use this size
  bool warp_ok = false;  // small design: a gene-chain can run on one warp (warp_ops.cuh)
  int forced_scope = -1; // CSPLOTCH_SCOPE=cta|warp at model creation: 0 / 1, -1 = by the number of gene-chains
  size_t warp_smem_bytes = 0;
};

// the model descriptor of a launch with clusters of `cl` CTAs (cl == 1: the plain one)
static DevModel with_cluster(const csplotch_model *m, int cl) {
  DevModel d = m->dm;
  d.cl = 1;
  d.ell_cl = d.ell;
  for (int r = 0; r <= kMaxCluster; ++r) d.cl_t0[r] = r == 0 ? 0 : d.NT;
  for (const auto &v : m->clv)
    if (v.cl == cl) {
      d.cl = cl;
      d.ell_cl = v.ell_dev;
      for (int r = 0; r <= kMaxCluster; ++r) d.cl_t0[r] = v.t0[r];
    }
  return d;
}
static bool has_cluster(const csplotch_model *m, int cl) {
  for (const auto &v : m->clv)
    if (v.cl == cl) return true;
  return cl == 1;
}

struct csplotch_batch {
  csplotch_model *m = nullptr;
  DevBuf buf;
  int G = 0;
  uint32_t gene_id0 = 0;
  uint32_t *gene_ids = nullptr;  // [G] device: the id that keys the RNG stream of every gene (gene_id0 + g unless given)
  int *counts = nullptr;     // [G][Nc] internal spot order
  double *pm = nullptr;      // [G][K]
  double *ps = nullptr;
  double *lgam = nullptr;    // [G]
  double *nnz = nullptr;     // [G]
};

struct SamplerDev {
  ChainScalars *cs;
  double *q_cur, *minv, *wm, *wm2;  // [W][Di]
  double *ws;                       // [slots][pool_vecs + 6][Di]
  double *diag, *small, *theta, *summ;
  int *queue;
  int W, num_chains, n_slots, have_init;
  int pool_vecs;                    // 3 * max_depth + 8 D-vectors of momenta / rho sums per slot
  const uint32_t *gene_ids;         // [G]
  const int *order;                 // [W] gene-chains in the order the queue hands them out (longest first)
  long long *slot_ns;               // [n_slots] nanoseconds from a CTA's start to its exit in the last launch (launch tail)
  NutsConfig cfg;
  const int *counts;
  const double *pm, *ps, *lgam, *nnz;
};

struct csplotch_sampler {
  csplotch_batch *b = nullptr;
  DevBuf buf;
  SamplerDev sd;
  csplotch_nuts_cfg cfg;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  int grid = 0;
  int64_t launches = 0;
  bool warp_scope = false;           // one gene-chain per warp (small designs, many chains)
  int cl = 1;                        // CTAs per gene-chain of this sampler's launches (choose_cluster)
  DevModel dml;                      // the model descriptor with that cluster partition
  int *order_dev = nullptr;          // [W] queue order of the next launch
  std::vector<double> nnz_host;      // [G] non-zero counts per gene: the cost estimate before the first transition
  std::vector<ChainScalars> cs_host;
  std::vector<int> order_host;
  ~csplotch_sampler() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }
};

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_repack_counts(const int *__restrict__ raw, int *__restrict__ dst, const int *__restrict__ perm,
                                int G, int N, int Nc) {
  const size_t total = (size_t)G * Nc;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i / Nc), j = (int)(i - (size_t)g * Nc);
    dst[i] = j < N ? raw[(size_t)g * N + perm[j]] : 0;
  }
}

__global__ void k_gene_consts(const int *__restrict__ counts, int G, int N, int Nc, double *lgam, double *nnz) {
  __shared__ double red[kWarps * kRedMax];
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    double a = 0, b = 0;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      const int y = counts[(size_t)g * Nc + j];
      if (y > 0) { a += lgamma((double)y + 1.0); b += 1.0; }
    }
    double r[2] = {a, b};
    block_sum<2>(r, red);
    if (threadIdx.x == 0) { lgam[g] = r[0]; nnz[g] = r[1]; }
  }
}

// Stan order [R][D] -> internal [R][Di]
__global__ void k_to_internal(const double *__restrict__ src, double *__restrict__ dst, const int *__restrict__ imap,
                              int R, int D, int Di, double pad_value) {
  const size_t total = (size_t)R * Di;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / Di;
    const int e = (int)(i - r * Di);
    const int si = imap[e];
    dst[i] = si >= 0 ? src[r * D + si] : pad_value;
  }
}
// internal [R][Di] -> Stan order [R][D]
__global__ void k_to_stan(const double *__restrict__ src, double *__restrict__ dst, const int *__restrict__ imap,
                          int R, int D, int Di) {
  const size_t total = (size_t)R * Di;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / Di;
    const int e = (int)(i - r * Di);
    const int si = imap[e];
    if (si >= 0) dst[r * D + si] = src[i];
  }
}

// pooled-over-chains mean / std (ddof=0) of the streaming summaries, original spot order
__global__ void k_summ_finalize(const double *__restrict__ summ, const ChainScalars *__restrict__ cs,
                                const int *__restrict__ perm, int G, int nch, int N, int num_warmup,
                                double *o0, double *o1, double *o2, double *o3, double *o4, double *o5) {
  const size_t total = (size_t)G * N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i / N), j = (int)(i - (size_t)g * N);
    const int jo = perm[j];
    double *outs[6] = {o0, o1, o2, o3, o4, o5};
    for (int q = 0; q < 3; ++q) {
      double n = 0, mean = 0, M2 = 0;
      for (int c = 0; c < nch; ++c) {
        const int w = g * nch + c;
        const double nc = (double)max(0, cs[w].iter - num_warmup);
        if (nc <= 0) continue;
        const double mc = summ[((size_t)w * 6 + 2 * q) * N + j], vc = summ[((size_t)w * 6 + 2 * q + 1) * N + j];
        const double d = mc - mean, nt = n + nc;
        mean += d * nc / nt;
        M2 += vc + d * d * n * nc / nt;
        n = nt;
      }
      if (outs[2 * q]) outs[2 * q][(size_t)g * N + jo] = mean;
      if (outs[2 * q + 1]) outs[2 * q + 1][(size_t)g * N + jo] = n > 0 ? sqrt(M2 / n) : 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// main kernels
// ------------------------------------------------------------------------------------------------
// resident CTAs per SM: two, at 128 registers per thread, for every kernel.  Round 1 ran the non-compositional models
// (KT <= 2) at three CTAs of 80 registers; with this round's tile loop that build spills inside it (555 LDL / 476 STL in
// the K = 1 leapfrog, and how many end up in the hot loop changed from build to build), and two CTAs of 128 registers
// measured +8 % on the c2 leapfrog, +4 % on the c2 sampler and +49 % on the c4 leapfrog (LOOK = 2, 283 MB of workspace
// per resident CTA) — GPU call 17, profiles/r02_kernel_experiments.md.  CSP_SMALL_CTAS=3 rebuilds the old setting.
#ifndef CSP_SMALL_CTAS
#define CSP_SMALL_CTAS 2
#endif
template <int KT>
constexpr int kCtasPerSm = KT <= 2 ? CSP_SMALL_CTAS : 2;
static int ctas_per_sm(int kt) { return kt <= 2 ? CSP_SMALL_CTAS : 2; }
__device__ __forceinline__ void stage_model(DevModel &dst, const DevModel &src) {
  const int n = (int)(sizeof(DevModel) / sizeof(int));
  for (int i = threadIdx.x; i < n; i += blockDim.x) reinterpret_cast<int *>(&dst)[i] = reinterpret_cast<const int *>(&src)[i];
  __syncthreads();
}

// dynamic shared memory: B[nbeta] | adjB[nbeta] | red | mbar[10] | tdesc | ring[4T or 8T] + zero cell | psi slots | rest slots
constexpr int kClScratch = 16 + 4 + 16 + 12 + 8;  // doubles: cl_part | cl_bcast | cl_red[2][8] | cl_sc | sraw
template <class Ops>
__device__ __forceinline__ void bind_smem(Ops &ops, const DevModel &m, double *smem) {
  ops.B_o = 0;
  ops.adjB_o = m.nbeta;
  ops.red_o = 2 * m.nbeta;
  double *p = smem + 2 * m.nbeta + kWarps * kRedMax;
  ops.mbar = reinterpret_cast<uint64_t *>(p);
  p += 10;
  ops.cl_part = p;
  ops.cl_bcast = p + 16;
  ops.cl_red = p + 20;
  ops.cl_sc = p + 36;
  ops.sraw_o = (int)(p + 48 - smem);
  p += kClScratch;
  ops.tdesc = reinterpret_cast<TmaDesc *>(p);
  p += (kMaxDesc * sizeof(TmaDesc) + 15) / 16 * 2;
  const int ring_tiles = m.look == 1 ? 4 : 8;
  ops.ring = p;
  p += ring_tiles * kTile + 2;  // + the always-zero cell missing neighbours point at
  ops.halo = p;                 // cluster kernels: psi of the tiles next to this rank's range (ell_cl positions)
  if (m.cl > 1) p += 2 * m.look * kTile;
  ops.psi_st = p;
  p += (m.look + 2) * 4 * kTile;
  ops.rest_st = reinterpret_cast<unsigned char *>(p);
  ops.tile_seq = 0;
  if (m.tiled) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < 6; ++i) mbar_init(ops.mbar + i, 1);
      for (int i = 6; i < 10; ++i) mbar_init(ops.mbar + i, kArriveCount);
      mbar_fence_init();
      ops.ring[ring_tiles * kTile] = 0.0;
      ops.ring[ring_tiles * kTile + 1] = 0.0;
    }
    __syncthreads();
  }
}

template <int KT>
__global__ void __launch_bounds__(kBlock, kCtasPerSm<KT>)
k_lpgrad(const __grid_constant__ DevModel m_in, const int *__restrict__ counts, const double *__restrict__ pm,
         const double *__restrict__ ps, const double *__restrict__ lgam, const double *__restrict__ nnz, int G,
         const double *__restrict__ theta_int, const int *__restrict__ gene_of, int Bn, double *lp_out,
         double *grad_int, int jacobian) {
  // the model descriptor is read all over the per-evaluation prologue / epilogue: a shared-memory copy turns
  // those dependent generic loads from the parameter bank into LDS
  __shared__ DevModel s_model;
  stage_model(s_model, m_in);
  const DevModel &m = s_model;
  DevOps<KT> ops(m);
  bind_smem(ops, m, smem_d);
  ops.jacobian = jacobian;
  ops.minv = nullptr;
  for (int b = blockIdx.x; b < Bn; b += gridDim.x) {
    const int g = gene_of ? gene_of[b] : (b % G);
    ops.gn.counts = counts + (size_t)g * m.Nc;
    ops.gn.pm = pm ? pm + (size_t)g * m.K : nullptr;
    ops.gn.ps = ps ? ps + (size_t)g * m.K : nullptr;
    ops.gn.lgam = lgam[g];
    ops.gn.nnz = nnz[g];
    double V;
    bool fin;
    const double *q = theta_int + (size_t)b * m.Di;
    ops.template eval<false>(q, nullptr, nullptr, const_cast<double *>(q), grad_int + (size_t)b * m.Di, nullptr,
                             0.0, &V, nullptr, &fin);
    if (threadIdx.x == 0) lp_out[b] = -V;
    __syncthreads();
  }
}

// grad_int of k_lpgrad holds g = -grad lp; flip the sign while converting to Stan order
__global__ void k_negate(double *v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) v[i] = -v[i];
}

template <int KT, bool CL>
__global__ void __launch_bounds__(kBlock, kCtasPerSm<KT>)
k_leapfrog_fixed(const __grid_constant__ DevModel m_in, const int *__restrict__ counts, const double *__restrict__ pm,
                 const double *__restrict__ ps, const double *__restrict__ lgam, const double *__restrict__ nnz,
                 int G, double *q, double *p, double *g, const double *__restrict__ minv,
                 const int *__restrict__ gene_of, int Bn, double eps, int n_steps, double *H_out) {
  // the model descriptor is read all over the per-evaluation prologue / epilogue: a shared-memory copy turns
  // those dependent generic loads from the parameter bank into LDS
  __shared__ DevModel s_model;
  stage_model(s_model, m_in);
  const DevModel &m = s_model;
  DevOps<KT, CL> ops(m);
  bind_smem(ops, m, smem_d);
  ops.bind_cluster();
  ops.jacobian = 1;
  // one trajectory per CTA, or per cluster of m.cl CTAs
  for (int b = blockIdx.x / ops.csize; b < Bn; b += gridDim.x / ops.csize) {
    const int gi = gene_of ? gene_of[b] : (b % G);
    ops.gn.counts = counts + (size_t)gi * m.Nc;
    ops.gn.pm = pm ? pm + (size_t)gi * m.K : nullptr;
    ops.gn.ps = ps ? ps + (size_t)gi * m.K : nullptr;
    ops.gn.lgam = lgam[gi];
    ops.gn.nnz = nnz[gi];
    ops.minv = const_cast<double *>(minv) + (size_t)b * m.Di;
    double *qb = q + (size_t)b * m.Di, *pb = p + (size_t)b * m.Di, *gb = g + (size_t)b * m.Di;
    double V, K = 0.0;
    bool fin;
    ops.template eval<false>(qb, nullptr, nullptr, qb, gb, nullptr, 0.0, &V, nullptr, &fin);
    for (int s = 0; s < n_steps; ++s) ops.template eval<true>(qb, gb, pb, qb, gb, pb, eps, &V, &K, nullptr);
    if (n_steps == 0) {
      double k = 0.0;
      for (int sgi = 0; sgi < ops.nseg; ++sgi)
        for (int e = ops.seg_lo[sgi] + threadIdx.x; e < ops.seg_hi[sgi]; e += kBlock) k += ops.minv[e] * pb[e] * pb[e];
      double r[1] = {k};
      ops.template chain_sum<1>(r);
      K = 0.5 * r[0];
    }
    if (threadIdx.x == 0 && ops.crank == 0) H_out[b] = V + K;
    __syncthreads();
  }
  ops.cluster_sync();   // no CTA of a cluster leaves while another may still read its shared memory
}

template <int KT, bool CL>
__global__ void __launch_bounds__(kBlock, kCtasPerSm<KT>)
k_nuts_advance(const __grid_constant__ DevModel m_in, const __grid_constant__ SamplerDev s, int n_iter) {
  __shared__ int s_work;
  const size_t Di = m_in.Di;
  long long t_start_ns = 0;
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start_ns));
  // the model descriptor is read all over the per-evaluation prologue / epilogue: a shared-memory copy turns
  // those dependent generic loads from the parameter bank into LDS
  __shared__ DevModel s_model;
  stage_model(s_model, m_in);
  const DevModel &m = s_model;
  DevOps<KT, CL> ops(m);
  bind_smem(ops, m, smem_d);
  ops.bind_cluster();
  // workspace slot of this CTA, or of its cluster (every rank touches only its own element ranges)
  double *ws = s.ws + (size_t)(blockIdx.x / ops.csize) * (size_t)(s.pool_vecs + 6) * Di;
  ops.jacobian = 1;
  for (;;) {
    if (threadIdx.x == 0 && ops.crank == 0) {
      const int t = atomicAdd(s.queue, 1);
      s_work = t < s.W ? s.order[t] : -1;
    }
    int w;
    if (CL && ops.csize > 1) {
      ops.cluster_sync();
      w = *ops.remote(&s_work, 0);
      ops.cluster_sync();   // rank 0 may pop again only after every rank has read this one
    } else {
      __syncthreads();
      w = s_work;
      __syncthreads();
    }
    if (w < 0) break;
    const int g = w / s.num_chains, c = w - g * s.num_chains;
    ops.gn.counts = s.counts + (size_t)g * m.Nc;
    ops.gn.pm = s.pm ? s.pm + (size_t)g * m.K : nullptr;
    ops.gn.ps = s.ps ? s.ps + (size_t)g * m.K : nullptr;
    ops.gn.lgam = s.lgam[g];
    ops.gn.nnz = s.nnz[g];
    ops.key = RngKey{s.cfg.seed, (uint32_t)c + 1u, s.gene_ids[g]};
    ops.pool = ws;
    ops.zq[0] = ws + (size_t)s.pool_vecs * Di;
    ops.zg[0] = ops.zq[0] + Di;
    ops.zq[1] = ops.zg[0] + Di;
    ops.zg[1] = ops.zq[1] + Di;
    ops.q_prop = ops.zg[1] + Di;
    ops.q_samp = ops.q_prop + Di;
    ops.q_cur = s.q_cur + (size_t)w * Di;
    ops.minv = s.minv + (size_t)w * Di;
    ops.wm = s.wm + (size_t)w * Di;
    ops.wm2 = s.wm2 + (size_t)w * Di;
    ops.sample_is_q0 = true;
    const size_t ns = (size_t)s.cfg.num_samples;
    ops.out_diag = s.diag ? s.diag + (size_t)w * ns * 7 : nullptr;
    ops.out_small = s.small ? s.small + (size_t)w * ns * m.nsmall_true : nullptr;
    ops.out_theta = s.theta ? s.theta + (size_t)w * ns * m.D : nullptr;
    ops.out_summ = s.summ ? s.summ + (size_t)w * 6 * m.N : nullptr;
    ChainScalars cs = s.cs[w];
    ops.cyc_eval = ops.cyc_merge = ops.cyc_copy = ops.cyc_other = ops.n_merge = ops.n_copy = ops.cyc_pro = ops.cyc_fin = 0;
    const long long t_chain = clock64();
    Nuts<DevOps<KT, CL>> nuts(ops, s.cfg, cs, ops.key);
    nuts.advance(n_iter, s.have_init != 0);
    // (cluster: every rank has read this chain's scalars before rank 0 overwrites them)
    if (CL && ops.csize > 1) ops.cluster_sync(); else __syncthreads();
    if (threadIdx.x == 0 && ops.crank == 0) {
      cs.cyc_eval += ops.cyc_eval; cs.cyc_merge += ops.cyc_merge; cs.cyc_copy += ops.cyc_copy;
      cs.n_merge += ops.n_merge; cs.n_copy += ops.n_copy; cs.cyc_pro += ops.cyc_pro; cs.cyc_fin += ops.cyc_fin;
      cs.cyc_other += (clock64() - t_chain) - ops.cyc_eval - ops.cyc_merge - ops.cyc_copy;
#ifdef CSPLOTCH_PROF_WAITS
      cs.cyc_merge = ops.w_issue; cs.cyc_copy = ops.w_rest; cs.cyc_other = ops.w_ring; cs.n_merge = ops.w_rel; cs.n_copy = ops.w_bins;
      cs.cyc_pro = ops.w_pre; cs.cyc_fin = ops.w_post; cs.cyc_eval = ops.cyc_eval; cs.n_grad_total = cs.n_grad_total;
#endif
      s.cs[w] = cs;
    }
  }
  if (threadIdx.x == 0) {
    long long t_end_ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end_ns));
    s.slot_ns[blockIdx.x] = t_end_ns - t_start_ns;
  }
}

// ---- small designs: one gene-chain per WARP (warp_ops.cuh) --------------------------------------------------------
constexpr int kWarpCtasPerSm = 3;
template <int KT>
__device__ __forceinline__ void bind_warp_smem(WarpOps<KT> &ops, const DevModel &m, double *smem) {
  const int warp = threadIdx.x >> 5;
  double *base = smem + (size_t)warp * (2 * m.nbeta + 8);
  ops.Bw = base;
  ops.Aw = base + m.nbeta;
  ops.Sw = base + 2 * m.nbeta;
}

template <int KT>
__global__ void __launch_bounds__(kBlock, kWarpCtasPerSm)
k_leapfrog_fixed_w(const __grid_constant__ DevModel m_in, const int *__restrict__ counts, const double *__restrict__ pm,
                   const double *__restrict__ ps, const double *__restrict__ lgam, const double *__restrict__ nnz, int G,
                   double *q, double *p, double *g, const double *__restrict__ minv, const int *__restrict__ gene_of, int Bn,
                   double eps, int n_steps, double *H_out) {
  __shared__ DevModel s_model;
  stage_model(s_model, m_in);
  const DevModel &m = s_model;
  WarpOps<KT> ops(m);
  bind_warp_smem(ops, m, smem_d);
  const int wid = blockIdx.x * kWarps + (threadIdx.x >> 5), nw = gridDim.x * kWarps;
  for (int b = wid; b < Bn; b += nw) {
    const int gi = gene_of ? gene_of[b] : (b % G);
    ops.gn.counts = counts + (size_t)gi * m.Nc;
    ops.gn.pm = pm ? pm + (size_t)gi * m.K : nullptr;
    ops.gn.ps = ps ? ps + (size_t)gi * m.K : nullptr;
    ops.gn.lgam = lgam[gi];
    ops.gn.nnz = nnz[gi];
    ops.minv = const_cast<double *>(minv) + (size_t)b * m.Di;
    double *qb = q + (size_t)b * m.Di, *pb = p + (size_t)b * m.Di, *gb = g + (size_t)b * m.Di;
    double V, K = 0.0;
    bool fin;
    ops.template eval<false>(qb, nullptr, nullptr, qb, gb, nullptr, 0.0, &V, nullptr, &fin);
    for (int s = 0; s < n_steps; ++s) ops.template eval<true>(qb, gb, pb, qb, gb, pb, eps, &V, &K, nullptr);
    if (n_steps == 0) {
      double k = 0.0;
      for (int e = ops.lane; e < m.Di; e += 32) k += ops.minv[e] * pb[e] * pb[e];
      K = 0.5 * warp_sum(k);
    }
    if (ops.lane == 0) H_out[b] = V + K;
    __syncwarp();
  }
}

template <int KT>
__global__ void __launch_bounds__(kBlock, kWarpCtasPerSm)
k_nuts_advance_w(const __grid_constant__ DevModel m_in, const __grid_constant__ SamplerDev s, int n_iter) {
  const size_t Di = m_in.Di;
  long long t_start_ns = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start_ns));
  __shared__ DevModel s_model;
  stage_model(s_model, m_in);
  const DevModel &m = s_model;
  WarpOps<KT> ops(m);
  bind_warp_smem(ops, m, smem_d);
  const int slot = blockIdx.x * kWarps + (threadIdx.x >> 5);
  double *ws = s.ws + (size_t)slot * (size_t)(s.pool_vecs + 6) * Di;
  for (;;) {
    int w = 0;
    if (ops.lane == 0) {
      const int t = atomicAdd(s.queue, 1);
      w = t < s.W ? s.order[t] : -1;
    }
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w < 0) break;
    const int g = w / s.num_chains, c = w - g * s.num_chains;
    ops.gn.counts = s.counts + (size_t)g * m.Nc;
    ops.gn.pm = s.pm ? s.pm + (size_t)g * m.K : nullptr;
    ops.gn.ps = s.ps ? s.ps + (size_t)g * m.K : nullptr;
    ops.gn.lgam = s.lgam[g];
    ops.gn.nnz = s.nnz[g];
    ops.key = RngKey{s.cfg.seed, (uint32_t)c + 1u, s.gene_ids[g]};
    ops.pool = ws;
    ops.zq[0] = ws + (size_t)s.pool_vecs * Di;
    ops.zg[0] = ops.zq[0] + Di;
    ops.zq[1] = ops.zg[0] + Di;
    ops.zg[1] = ops.zq[1] + Di;
    ops.q_prop = ops.zg[1] + Di;
    ops.q_samp = ops.q_prop + Di;
    ops.q_cur = s.q_cur + (size_t)w * Di;
    ops.minv = s.minv + (size_t)w * Di;
    ops.wm = s.wm + (size_t)w * Di;
    ops.wm2 = s.wm2 + (size_t)w * Di;
    ops.sample_is_q0 = true;
    const size_t ns = (size_t)s.cfg.num_samples;
    ops.out_diag = s.diag ? s.diag + (size_t)w * ns * 7 : nullptr;
    ops.out_small = s.small ? s.small + (size_t)w * ns * m.nsmall_true : nullptr;
    ops.out_theta = s.theta ? s.theta + (size_t)w * ns * m.D : nullptr;
    ops.out_summ = s.summ ? s.summ + (size_t)w * 6 * m.N : nullptr;
    ChainScalars cs = s.cs[w];
    ops.cyc_eval = ops.cyc_merge = ops.cyc_copy = ops.cyc_other = ops.n_merge = ops.n_copy = ops.cyc_pro = ops.cyc_fin = 0;
    const long long t_chain = clock64();
    Nuts<WarpOps<KT>> nuts(ops, s.cfg, cs, ops.key);
    nuts.advance(n_iter, s.have_init != 0);
    __syncwarp();
    if (ops.lane == 0) {
      cs.cyc_eval += ops.cyc_eval; cs.cyc_merge += ops.cyc_merge; cs.cyc_copy += ops.cyc_copy;
      cs.n_merge += ops.n_merge; cs.n_copy += ops.n_copy;
      cs.cyc_other += (clock64() - t_chain) - ops.cyc_eval - ops.cyc_merge - ops.cyc_copy;
      s.cs[w] = cs;
    }
    __syncwarp();
  }
  if (ops.lane == 0) {
    long long t_end_ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end_ns));
    s.slot_ns[slot] = t_end_ns - t_start_ns;
  }
}

// ------------------------------------------------------------------------------------------------
// host: model
// ------------------------------------------------------------------------------------------------
extern "C" const char *csplotch_last_error(void) { return g_err.c_str(); }

// page-locked host memory for the caller's input / result arrays: cudaMemcpy to or from such a buffer is one
// DMA at PCIe speed instead of a staged copy through the driver's bounce buffers (pageable numpy arrays
// arrive at ~2 GB/s: as long as sampling itself for the 1.4 GB of spot summaries of the c2 gene list)
extern "C" int32_t csplotch_host_alloc(uint64_t bytes, void **out) {
  if (!out) return fail(CSPLOTCH_E_INVALID, "null argument");
  *out = nullptr;
  if (bytes == 0) return CSPLOTCH_OK;
  if (cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    *out = nullptr;
    return fail(CSPLOTCH_E_NOMEM, "cudaHostAlloc failed (no device, or not enough lockable host memory)");
  }
  return CSPLOTCH_OK;
}
extern "C" int32_t csplotch_host_free(void *p) {
  if (p) CU_TRY(cudaFreeHost(p));
  return CSPLOTCH_OK;
}
extern "C" int32_t csplotch_abi_version(void) { return CSPLOTCH_ABI_VERSION; }
extern "C" int32_t csplotch_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
extern "C" int32_t csplotch_set_device(int32_t device) {
  CU_TRY(cudaSetDevice(device));
  return CSPLOTCH_OK;
}

// internal (padded) number of cell types = the template instantiation; padded entries are dead
// elements (E = 0, q = p = g = 0) exactly like padding spots
static int pick_kt(int K) {
  const int opts[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16};
  for (int o : opts)
    if (K <= o) return o;
  return 0;
}

extern "C" int32_t csplotch_model_create(const csplotch_shared_desc *d, csplotch_model **out) {
  if (!d || !out) return fail(CSPLOTCH_E_INVALID, "null argument");
  *out = nullptr;
  if (d->family < 0 || d->family > 2) return fail(CSPLOTCH_E_INVALID, "family must be 0,1,2");
  if (d->nb < 0 || d->nb > 2) return fail(CSPLOTCH_E_INVALID, "nb must be 0, 1 (reference semantics) or 2 (per-spot ZINB term)");
  if (d->N_tissues < 1 || !d->N_spots || !d->tissue_mapping || !d->D || !d->size_factors)
    return fail(CSPLOTCH_E_INVALID, "missing N_spots / tissue_mapping / D / size_factors");
  const int T = d->N_tissues, C = d->N_covariates;
  const int K = d->family == CSPLOTCH_COMP_SPLOTCH ? d->N_celltypes : 1;
  const int L1 = d->N_level_1, L2 = d->N_level_2, L3 = d->N_level_3, L = L1 + L2 + L3;
  if (C < 1 || K < 1 || L1 < 1 || L2 < 0 || L3 < 0) return fail(CSPLOTCH_E_INVALID, "bad N_covariates / N_celltypes / N_level_k");
  if (L3 > 0 && L2 == 0) return fail(CSPLOTCH_E_INVALID, "N_level_3 > 0 needs N_level_2 > 0");
  if (d->family == CSPLOTCH_COMP_SPLOTCH && !d->E) return fail(CSPLOTCH_E_INVALID, "comp model needs E");
  if ((L2 && !d->level_2_mapping) || (L3 && !d->level_3_mapping)) return fail(CSPLOTCH_E_INVALID, "missing level mapping");
  long long Nll = 0;
  for (int t = 0; t < T; ++t) {
    if (d->N_spots[t] < 1) return fail(CSPLOTCH_E_INVALID, "N_spots must be >= 1");
    Nll += d->N_spots[t];
  }
  if (Nll > (1ll << 30)) return fail(CSPLOTCH_E_INVALID, "too many spots");
  const int N = (int)Nll;
  const int car = d->car ? 1 : 0, zi = d->zi ? 1 : 0;
  const int Lbot = L3 ? L3 : (L2 ? L2 : L1);
  const int kt = pick_kt(K);
  if (!kt) return fail(CSPLOTCH_E_UNSUPPORTED, "N_celltypes > 16 is not built yet");

  auto *m = new csplotch_model();
  cudaGetDevice(&m->device);
  m->T = T;
  m->kt = kt;
  DevModel &dm = m->dm;
  memset(&dm, 0, sizeof(dm));
  dm.family = d->family; dm.N = N; dm.NT = (N + kTile - 1) / kTile; dm.Npad = dm.NT * kTile; dm.C = C; dm.K = kt; dm.Ktrue = K;
  dm.L1 = L1; dm.L2 = L2; dm.L3 = L3; dm.L = L;
  dm.bot0 = L3 ? L1 + L2 : (L2 ? L1 : 0);
  dm.nbeta = L * C * kt;
  dm.nbeta_true = L * C * K;
  dm.zi = zi; dm.car = car;
  dm.nb = d->family == CSPLOTCH_COMP_SPLOTCH ? d->nb : 0;
  int ns = 0;
  dm.s_tau = car ? ns++ : -1;
  dm.s_a = car ? ns++ : -1;
  dm.s_sigma = ns++;
  dm.s_s2 = L2 ? ns++ : -1;
  dm.s_s3 = L3 ? ns++ : -1;
  dm.s_theta = zi ? ns++ : -1;
  dm.s_phi = dm.nb ? ns++ : -1;
  dm.nscal = ns;
  dm.nsmall = dm.nbeta + ns;
  dm.nsmall_true = dm.nbeta_true + ns;
  dm.nsmall_pad = (dm.nsmall + 1) & ~1;
  dm.o_psi = 0;
  dm.o_noise = car ? dm.Npad : 0;
  dm.o_small = dm.o_noise + dm.Npad;
  dm.Di = dm.o_small + dm.nsmall_pad;
  dm.D = (car ? N : 0) + dm.nbeta_true + ns + N;
  dm.Nc = dm.Npad;
  m->n_outputs = dm.D + N + (dm.nb ? N : 0) + dm.nbeta_true;

  // ---- adjacency in original spot ids ---------------------------------------------------------------
  std::vector<std::vector<int>> adj0(car ? N : 0);
  if (car) {
    if (!d->eig_values) { delete m; return fail(CSPLOTCH_E_INVALID, "car=1 needs eig_values"); }
    if (d->U && d->V) {
      for (int i = 0; i < N; ++i)
        for (int e = d->U[i] - 1; e < d->U[i + 1] - 1; ++e) {
          const int nb = d->V[e] - 1;
          if (nb < 0 || nb >= N) { delete m; return fail(CSPLOTCH_E_INVALID, "V out of range"); }
          adj0[i].push_back(nb);
        }
    } else if (d->W_sparse) {
      for (int e = 0; e < d->W_n; ++e) {
        const int a = d->W_sparse[2 * e] - 1, b = d->W_sparse[2 * e + 1] - 1;
        if (a < 0 || a >= N || b < 0 || b >= N) { delete m; return fail(CSPLOTCH_E_INVALID, "W_sparse out of range"); }
        adj0[a].push_back(b);
        adj0[b].push_back(a);
      }
    } else {
      delete m;
      return fail(CSPLOTCH_E_INVALID, "car=1 needs W_sparse or U,V");
    }
    for (int i = 0; i < N; ++i) {
      std::sort(adj0[i].begin(), adj0[i].end());
      if (d->D_sparse && std::fabs(d->D_sparse[i] - (double)adj0[i].size()) > 0.5) {
        delete m;
        return fail(CSPLOTCH_E_INVALID, "D_sparse does not match the adjacency");
      }
    }
  }

  // bottom-level bin of every spot (original order): (tissue_mapping[t]-1)*C + D_j-1
  std::vector<int> bin0_of(N), tissue_idx(N);
  {
    int j = 0;
    for (int t = 0; t < T; ++t) {
      const int tm = d->tissue_mapping[t];
      if (tm < 1 || tm > Lbot) { delete m; return fail(CSPLOTCH_E_INVALID, "tissue_mapping out of range"); }
      for (int sidx = 0; sidx < d->N_spots[t]; ++sidx, ++j) {
        if (d->D[j] < 1 || d->D[j] > C) { delete m; return fail(CSPLOTCH_E_INVALID, "D out of range"); }
        bin0_of[j] = (tm - 1) * C + (d->D[j] - 1);
        tissue_idx[j] = t;
      }
    }
  }

  // ---- internal spot order ---------------------------------------------------------------------------
  // The model is invariant under any relabelling of spots.  The tiled kernel needs every neighbour in the
  // same or an adjacent tile of kTile spots.  Try the given order first (row-major arrays already satisfy
  // it); otherwise relabel each connected component breadth-first (Cuthill-McKee), which bounds the
  // bandwidth by the width of the tissue; if that still does not fit the generic kernel is used.
  // smallest neighbour reach in tiles (1 or 2) under which `perm` tiles, 0 if it does not
  auto reach = [&](const std::vector<int> &perm) {
    if (!car) return 1;
    std::vector<int> inv(N);
    for (int j = 0; j < N; ++j) inv[perm[j]] = j;
    int r = 1;
    for (int j = 0; j < N; ++j) {
      const auto &a = adj0[perm[j]];
      if ((int)a.size() > kMaxEll) return 0;
      for (int nb : a) {
        const int dt = std::abs(inv[nb] / kTile - j / kTile);
        if (dt > 2) return 0;
        r = std::max(r, dt);
      }
    }
    return r;
  };
  auto fits = [&](const std::vector<int> &perm) { return reach(perm) != 0; };
  m->perm.resize(N);
  for (int j = 0; j < N; ++j) m->perm[j] = j;
  bool tiled = fits(m->perm);
  const char *force = std::getenv("CSPLOTCH_FORCE_ORDER");  // "cm": always relabel, "generic": never tile (tests)
  if (car && (!tiled || (force && std::string(force) == "cm"))) {
    std::vector<int> cm;
    cm.reserve(N);
    std::vector<char> seen(N, 0);
    for (int s0 = 0; s0 < N; ++s0) {
      if (seen[s0]) continue;
      // start each component from a far-away node (two BFS sweeps ~ pseudo-peripheral)
      int start = s0;
      for (int sweep = 0; sweep < 2; ++sweep) {
        std::vector<int> q{start};
        std::vector<char> vis(N, 0);
        vis[start] = 1;
        for (size_t h = 0; h < q.size(); ++h)
          for (int nb : adj0[q[h]])
            if (!vis[nb]) { vis[nb] = 1; q.push_back(nb); }
        start = q.back();
      }
      std::vector<int> q{start};
      seen[start] = 1;
      for (size_t h = 0; h < q.size(); ++h) {
        std::vector<int> nbs;
        for (int nb : adj0[q[h]])
          if (!seen[nb]) { seen[nb] = 1; nbs.push_back(nb); }
        std::sort(nbs.begin(), nbs.end(), [&](int a, int b) {
          return adj0[a].size() != adj0[b].size() ? adj0[a].size() < adj0[b].size() : a < b;
        });
        q.insert(q.end(), nbs.begin(), nbs.end());
      }
      cm.insert(cm.end(), q.begin(), q.end());
    }
    if (fits(cm)) { m->perm = cm; tiled = true; }
  }
  // refinement: inside each tissue list the spots bin by bin (stable), when that still tiles.  A warp
  // then sees one bin per tile almost always, so its contribution to the beta gradient is one transposed
  // warp reduction instead of one per distinct bin (small ST tissues in file order mix all AARs in a warp).
  if (tiled && !(force && std::string(force) == "nosort")) {
    std::vector<int> cand = m->perm;
    std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) {
      if (tissue_idx[a] != tissue_idx[b]) return tissue_idx[a] < tissue_idx[b];
      return bin0_of[a] < bin0_of[b];
    });
    // tissues must stay contiguous for this to make sense: perm lists whole tissues one after another in
    // both the given and the breadth-first order (components never straddle tissues)
    if (cand != m->perm && reach(cand) != 0 && reach(cand) <= reach(m->perm)) m->perm = cand;
  }
  if (force && std::string(force) == "generic") tiled = false;
  if (dm.nb) tiled = false;  // the NB / ZINB likelihood lives in the generic kernel only
  std::vector<int> inv(N);
  for (int j = 0; j < N; ++j) inv[m->perm[j]] = j;
  int ell_w = 0;
  if (car)
    for (int i = 0; i < N; ++i) ell_w = std::max(ell_w, (int)adj0[i].size());
  dm.tiled = tiled ? 1 : 0;
  dm.look = tiled ? std::max(1, reach(m->perm)) : 1;
  dm.ell_w = (car && tiled) ? ell_w : 0;
  dm.stage_bytes = (5 * kTile) * (int)sizeof(double) + 2 * kTile * (int)sizeof(int) + ((dm.ell_w + 1) / 2 * 2) * kTile * (int)sizeof(unsigned short);
  // gene-chains of long designs can run on a thread-block cluster (DevModel::cl); the partitions are prepared here,
  // the size is chosen per launch from the number of gene-chains (choose_cluster)
  dm.cl = 1;
  for (int r = 0; r <= kMaxCluster; ++r) dm.cl_t0[r] = r == 0 ? 0 : dm.NT;
  if (tiled) {
    m->max_cl = dm.NT >= 160 ? 4 : (dm.NT >= 64 ? 2 : 1);
    if (const char *fc = std::getenv("CSPLOTCH_CLUSTER")) {
      const int f = std::atoi(fc);
      if (f == 1 || f == 2 || f == 4 || f == 8) m->forced_cl = f;
    }
    const int min_tiles = dm.look + 3;
    for (int cl = 2; cl <= kMaxCluster; cl *= 2) {
      if (dm.NT < cl * min_tiles) break;
      if (cl > m->max_cl && cl != m->forced_cl) continue;
      csplotch_model::ClusterVariant v;
      v.cl = cl;
      v.ell_dev = nullptr;
      // equal shares: rank 0's epilogue of the small block runs after every rank's tiles are done (it needs their
      // partial sums), so giving rank 0 fewer tiles would only lengthen the others' share (measured: 107 / 120 tiles
      // at c3 cost 2 % against 117 / 117)
      for (int r = 0; r <= cl; ++r) v.t0[r] = (int)((long long)dm.NT * r / cl);
      for (int r = cl + 1; r <= kMaxCluster; ++r) v.t0[r] = dm.NT;
      m->clv.push_back(v);
    }
    if (m->forced_cl > 1 && !has_cluster(m, m->forced_cl)) m->forced_cl = m->clv.empty() ? 1 : m->clv.back().cl;
  }
  m->warp_ok = dm.nbeta <= kWarpMaxBeta && dm.Di <= 32768;
  m->warp_smem_bytes = (size_t)kWarps * (2 * dm.nbeta + 8) * sizeof(double);
  if (const char *fs = std::getenv("CSPLOTCH_SCOPE")) {
    if (std::string(fs) == "warp") m->forced_scope = 1;
    else if (std::string(fs) == "cta") m->forced_scope = 0;
  }
  m->smem_bytes = (size_t)(2 * dm.nbeta + kWarps * kRedMax + 10 + kClScratch) * sizeof(double) + (kMaxDesc * sizeof(TmaDesc) + 15) / 16 * 16;
  if (tiled) m->smem_bytes += ((size_t)((dm.look == 1 ? 4 : 8) + 4 * (dm.look + 2)) * kTile + 2) * sizeof(double) + 2 * (size_t)dm.stage_bytes;
  if (!m->clv.empty()) m->smem_bytes += (size_t)2 * dm.look * kTile * sizeof(double);
  if (m->smem_bytes > 220 * 1024) {
    delete m;
    return fail(CSPLOTCH_E_UNSUPPORTED, "beta table (L*C*K) too large for shared memory");
  }

  // Stan offsets (declaration order)
  const int so_psi = 0, so_beta = car ? N : 0, so_scal = so_beta + dm.nbeta_true, so_noise = so_scal + ns;
  m->imap.assign(dm.Di, -1);
  for (int j = 0; j < N; ++j) {
    if (car) m->imap[dm.o_psi + j] = so_psi + m->perm[j];
    m->imap[dm.o_noise + j] = so_noise + m->perm[j];
  }
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < C; ++j)
      for (int k = 0; k < K; ++k) {
        const int internal = (i * C + j) * kt + k;
        const int stan = d->family == CSPLOTCH_COMP_SPLOTCH ? (i * C + j) * K + k : i + L * j;
        m->imap[dm.o_small + internal] = so_beta + stan;
      }
  for (int s = 0; s < ns; ++s) m->imap[dm.o_small + dm.nbeta + s] = so_scal + s;

  // per-spot tables
  std::vector<double> lsf(dm.Npad, 0.0);
  std::vector<int> key(dm.Npad, -1);
  for (int ji = 0; ji < N; ++ji) {
    const int jo = m->perm[ji];
    if (!(d->size_factors[jo] > 0)) { delete m; return fail(CSPLOTCH_E_INVALID, "size_factors must be > 0"); }
    lsf[ji] = std::log(d->size_factors[jo]);
    key[ji] = bin0_of[jo] | ((car ? (int)adj0[jo].size() : 0) << 24);
  }
  std::vector<int> map2(L2), map3(L3);
  for (int i = 0; i < L2; ++i) {
    map2[i] = d->level_2_mapping[i] - 1;
    if (map2[i] < 0 || map2[i] >= L1) { delete m; return fail(CSPLOTCH_E_INVALID, "level_2_mapping out of range"); }
  }
  for (int i = 0; i < L3; ++i) {
    map3[i] = d->level_3_mapping[i] - 1;
    if (map3[i] < 0 || map3[i] >= L2) { delete m; return fail(CSPLOTCH_E_INVALID, "level_3_mapping out of range"); }
  }
  std::vector<int> rowptr(N + 1, 0), col;
  std::vector<unsigned short> ell;
  std::vector<double> eigu, eigm, eigg, eiggm;
  if (car) {
    const int ring_n = (dm.look == 1 ? 4 : 8) * kTile;
    if (tiled) ell.assign((size_t)dm.NT * dm.ell_w * kTile, (unsigned short)ring_n);
    for (int i = 0; i < N; ++i) {
      std::vector<int> a;
      for (int nb : adj0[m->perm[i]]) a.push_back(inv[nb]);
      std::sort(a.begin(), a.end());
      rowptr[i + 1] = rowptr[i] + (int)a.size();
      col.insert(col.end(), a.begin(), a.end());
      if (tiled)
        for (size_t w = 0; w < a.size(); ++w)
          ell[((size_t)(i / kTile) * dm.ell_w + w) * kTile + (i % kTile)] = (unsigned short)(a[w] & (ring_n - 1));
    }
    // the clusters' tables: a neighbour outside the tile range of the spot's rank lives in that rank's halo cells
    for (auto &v : m->clv) {
      std::vector<unsigned short> ec = ell;
      std::vector<int> rank_of_tile(dm.NT, 0);
      for (int r = 0; r < v.cl; ++r)
        for (int t = v.t0[r]; t < v.t0[r + 1]; ++t) rank_of_tile[t] = r;
      for (int i = 0; i < N; ++i) {
        const int r = rank_of_tile[i / kTile], ta = v.t0[r], tb = v.t0[r + 1];
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
          const int a = col[e], t = a / kTile;
          int pos;
          if (t >= ta && t < tb) continue;
          if (t < ta) pos = ring_n + 2 + (a - (ta - dm.look) * kTile);
          else pos = ring_n + 2 + dm.look * kTile + (a - tb * kTile);
          ec[((size_t)(i / kTile) * dm.ell_w + (e - rowptr[i])) * kTile + (i % kTile)] = (unsigned short)pos;
        }
      }
      if (m->buf.upload(&v.ell_dev, ec) != cudaSuccess) {
        delete m;
        return fail(CSPLOTCH_E_CUDA, "model upload failed (cluster table)");
      }
    }
    // distinct eigenvalues with multiplicity (exact duplicates only; c3 has 24 identical arrays)
    std::map<double, double> mult;
    for (int i = 0; i < N; ++i) mult[d->eig_values[i]] += 1.0;
    for (auto &kv : mult) { eigu.push_back(kv.first); eigm.push_back(kv.second); }
    dm.n_eig = (int)eigu.size();
    // groups of kEigGroup eigenvalues with the same multiplicity (eig_group_terms): one log and one division each
    std::vector<std::pair<double, double>> by_mult;  // (multiplicity, eigenvalue)
    for (auto &kv : mult) by_mult.emplace_back(kv.second, kv.first);
    std::sort(by_mult.begin(), by_mult.end());
    for (size_t i = 0; i < by_mult.size();) {
      const double mlt = by_mult[i].first;
      int cnt = 0;
      for (; i < by_mult.size() && by_mult[i].first == mlt && cnt < kEigGroup; ++i, ++cnt) eigg.push_back(by_mult[i].second);
      for (; cnt < kEigGroup; ++cnt) eigg.push_back(0.0);  // u = 1: contributes nothing to either sum
      eiggm.push_back(mlt);
    }
    dm.n_eigg = (int)eiggm.size();
  }
  std::vector<double> E;
  if (d->family == CSPLOTCH_COMP_SPLOTCH) {
    E.assign((size_t)dm.NT * kt * kTile, 0.0);
    for (int ji = 0; ji < N; ++ji)
      for (int k = 0; k < K; ++k)
        E[((size_t)(ji / kTile) * kt + k) * kTile + (ji % kTile)] = d->E[(size_t)m->perm[ji] * K + k];
  }
  bool ok = true;
  double *p_lsf, *p_E, *p_eu, *p_em, *p_eg, *p_egm;
  int *p_key, *p_rp, *p_col, *p_m2, *p_m3, *p_imap;
  unsigned short *p_ell;
  ok &= m->buf.upload(&p_lsf, lsf) == cudaSuccess;
  ok &= m->buf.upload(&p_key, key) == cudaSuccess;
  ok &= m->buf.upload(&p_rp, rowptr) == cudaSuccess;
  ok &= m->buf.upload(&p_col, col) == cudaSuccess;
  ok &= m->buf.upload(&p_ell, ell) == cudaSuccess;
  ok &= m->buf.upload(&p_E, E) == cudaSuccess;
  ok &= m->buf.upload(&p_eu, eigu) == cudaSuccess;
  ok &= m->buf.upload(&p_em, eigm) == cudaSuccess;
  ok &= m->buf.upload(&p_eg, eigg) == cudaSuccess;
  ok &= m->buf.upload(&p_egm, eiggm) == cudaSuccess;
  ok &= m->buf.upload(&p_m2, map2) == cudaSuccess;
  ok &= m->buf.upload(&p_m3, map3) == cudaSuccess;
  ok &= m->buf.upload(&p_imap, m->imap) == cudaSuccess;
  if (!ok) {
    std::string msg = cudaGetErrorString(cudaGetLastError());
    delete m;
    return fail(CSPLOTCH_E_CUDA, "model upload failed: " + msg);
  }
  dm.lsf = p_lsf; dm.key = p_key; dm.rowptr = p_rp; dm.col = p_col; dm.ell = p_ell; dm.ell_cl = p_ell; dm.E = p_E; dm.eigu = p_eu; dm.eigm = p_em; dm.eigg = p_eg; dm.eiggm = p_egm;
  dm.map2 = p_m2; dm.map3 = p_m3; dm.imap = p_imap;
  *out = m;
  return CSPLOTCH_OK;
}

extern "C" void csplotch_model_destroy(csplotch_model *m) { delete m; }

extern "C" int32_t csplotch_model_dims(const csplotch_model *m, csplotch_dims *o) {
  if (!m || !o) return fail(CSPLOTCH_E_INVALID, "null argument");
  o->N = m->dm.N; o->dim = m->dm.D; o->n_beta = m->dm.nbeta_true; o->n_scalar = m->dm.nscal;
  o->n_small = m->dm.nsmall_true; o->n_outputs = m->n_outputs;
  return CSPLOTCH_OK;
}

// ------------------------------------------------------------------------------------------------
// host: batch
// ------------------------------------------------------------------------------------------------
static int batch_finish(csplotch_model *m, csplotch_batch *b, const int *counts_raw_dev, const double *pm,
                        const double *ps, cudaStream_t st) {
  const DevModel &dm = m->dm;
  int *perm_dev = nullptr;
  CU_TRY(b->buf.upload(&perm_dev, m->perm));
  CU_TRY(b->buf.alloc(&b->counts, (size_t)b->G * dm.Nc));
  CU_TRY(b->buf.alloc(&b->lgam, (size_t)b->G));
  CU_TRY(b->buf.alloc(&b->nnz, (size_t)b->G));
  {
    std::vector<uint32_t> ids((size_t)b->G);
    for (int g = 0; g < b->G; ++g) ids[g] = b->gene_id0 + (uint32_t)g;
    CU_TRY(b->buf.upload(&b->gene_ids, ids));
  }
  k_repack_counts<<<1184, 256, 0, st>>>(counts_raw_dev, b->counts, perm_dev, b->G, dm.N, dm.Nc);
  k_gene_consts<<<std::min(b->G, 1184), kBlock, 0, st>>>(b->counts, b->G, dm.N, dm.Nc, b->lgam, b->nnz);
  CU_TRY(cudaGetLastError());
  if (dm.family == CSPLOTCH_COMP_SPLOTCH) {
    if (!pm || !ps) return fail(CSPLOTCH_E_INVALID, "comp model needs beta_prior_mean / beta_prior_std");
    std::vector<double> hm((size_t)b->G * dm.K, 0.0), hs((size_t)b->G * dm.K, 1.0);
    for (int g = 0; g < b->G; ++g)
      for (int k = 0; k < dm.Ktrue; ++k) {
        hm[(size_t)g * dm.K + k] = pm[(size_t)g * dm.Ktrue + k];
        hs[(size_t)g * dm.K + k] = ps[(size_t)g * dm.Ktrue + k];
        if (!(ps[(size_t)g * dm.Ktrue + k] > 0)) return fail(CSPLOTCH_E_INVALID, "beta_prior_std must be > 0");
      }
    CU_TRY(b->buf.upload(&b->pm, hm));
    CU_TRY(b->buf.upload(&b->ps, hs));
  }
  CU_TRY(cudaStreamSynchronize(st));
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_batch_create_dev(csplotch_model *m, const int32_t *counts_dev, const double *pm,
                                             const double *ps, int32_t G, uint32_t gene_id0, void *stream,
                                             csplotch_batch **out) {
  if (!m || !counts_dev || !out || G < 1) return fail(CSPLOTCH_E_INVALID, "bad argument");
  auto *b = new csplotch_batch();
  b->m = m; b->G = G; b->gene_id0 = gene_id0;
  int rc = batch_finish(m, b, counts_dev, pm, ps, (cudaStream_t)stream);
  if (rc) { delete b; return rc; }
  *out = b;
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_batch_create(csplotch_model *m, const int32_t *counts, const double *pm,
                                         const double *ps, int32_t G, uint32_t gene_id0, csplotch_batch **out) {
  if (!m || !counts || !out || G < 1) return fail(CSPLOTCH_E_INVALID, "bad argument");
  const size_t n = (size_t)G * m->dm.N;
  for (size_t i = 0; i < n; ++i)
    if (counts[i] < 0) return fail(CSPLOTCH_E_INVALID, "counts must be >= 0");
  int *raw = nullptr;
  CU_TRY(cudaMalloc(&raw, n * sizeof(int)));
  cudaError_t e = cudaMemcpy(raw, counts, n * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(raw); return fail(CSPLOTCH_E_CUDA, cudaGetErrorString(e)); }
  int rc = csplotch_batch_create_dev(m, raw, pm, ps, G, gene_id0, nullptr, out);
  cudaFree(raw);
  return rc;
}

extern "C" void csplotch_batch_destroy(csplotch_batch *b) { delete b; }

// The RNG stream of a gene is keyed by its id, so a gene list may be handed over in ANY order (cost-sorted,
// dealt round-robin over ranks) without changing a single draw.
extern "C" int32_t csplotch_batch_set_gene_ids(csplotch_batch *b, const uint32_t *gene_ids) {
  if (!b || !gene_ids) return fail(CSPLOTCH_E_INVALID, "null argument");
  CU_TRY(cudaMemcpy(b->gene_ids, gene_ids, (size_t)b->G * sizeof(uint32_t), cudaMemcpyHostToDevice));
  return CSPLOTCH_OK;
}

// ------------------------------------------------------------------------------------------------
// host: launch helpers
// ------------------------------------------------------------------------------------------------
static int sm_count() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

#ifdef CSPLOTCH_DEV_KT
// development builds (python -m csplotch_b200.build --dev): only the two instantiations the named configs use,
// a quarter of the compile time; never shipped (separate output file, loaded only through CSPLOTCH_LIB)
#define DISPATCH_KT(kt, ...)                 \
  switch (kt) {                               \
    case 1: { constexpr int KT = 1; __VA_ARGS__; } break;   \
    case 10: { constexpr int KT = 10; __VA_ARGS__; } break; \
    default: return fail(CSPLOTCH_E_UNSUPPORTED, "development build: N_celltypes in {1, 10} only"); \
  }
#else
#define DISPATCH_KT(kt, ...)                 \
  switch (kt) {                               \
    case 1: { constexpr int KT = 1; __VA_ARGS__; } break;   \
    case 2: { constexpr int KT = 2; __VA_ARGS__; } break;   \
    case 3: { constexpr int KT = 3; __VA_ARGS__; } break;   \
    case 4: { constexpr int KT = 4; __VA_ARGS__; } break;   \
    case 5: { constexpr int KT = 5; __VA_ARGS__; } break;   \
    case 6: { constexpr int KT = 6; __VA_ARGS__; } break;   \
    case 8: { constexpr int KT = 8; __VA_ARGS__; } break;   \
    case 10: { constexpr int KT = 10; __VA_ARGS__; } break; \
    case 12: { constexpr int KT = 12; __VA_ARGS__; } break; \
    case 16: { constexpr int KT = 16; __VA_ARGS__; } break; \
    default: return fail(CSPLOTCH_E_UNSUPPORTED, "unsupported N_celltypes"); \
  }
#endif

template <class Kern>
static cudaError_t set_smem(Kern kern, size_t bytes) {
  // (no cudaFuncAttributePreferredSharedMemoryCarveout: the driver's own choice keeps two CTAs resident up to 112 kB each
  //  (tools/occ_probe.cu); a hand-sized carve-out was no faster in the one run that tried it, and that run is not
  //  conclusive — profiles/r02_kernel_experiments.md, "build non-determinism")
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Cluster size for `n_chains` gene-chains: the smallest size that leaves >= 8 work units per resident slot (the tail
// of a launch is about 0.8 / units of its length: measured 10 % at 8), at most what the design supports.  Few chains
// (one gene block of bin/splotch, the ESS job, a strong-scaling share) thus run on 4-CTA clusters, thousands of
// chains on plain CTAs, whose throughput is 5-10 % higher (no serial rank-0 epilogue).
static int choose_cluster(const csplotch_model *m, long long n_chains) {
  if (m->forced_cl) return m->forced_cl;
  const long long slots = (long long)sm_count() * ctas_per_sm(m->kt);
  int cl = 1;
  while (cl < m->max_cl && has_cluster(m, 2 * cl) && n_chains * cl < 8 * slots) cl *= 2;
  return cl;
}

// One warp per gene-chain (warp_ops.cuh) for small designs: selected with CSPLOTCH_SCOPE=warp at model creation.
static bool choose_warp_scope(const csplotch_model *m, long long n_chains) {
  if (!m->warp_ok) return false;
  if (m->forced_scope >= 0) return m->forced_scope == 1;
  (void)n_chains;
  return false;   // opt-in (CSPLOTCH_SCOPE=warp): bitwise reproducible, but its un-pipelined spot loop runs c2 at 0.65x the
                  // CTA path (profiles/r02_kernel_experiments.md); it becomes the default for small designs once that
                  // loop batches its loads
}
static int warp_resident_ctas() {
  return sm_count() * kWarpCtasPerSm;
}

// kernels that run one gene-chain per CTA (CLV = false) or per thread-block cluster of CTAs (CLV = true)
#define DISPATCH_KT_CL(kt, clustered, ...)                                      \
  if (clustered) { constexpr bool CLV = true; DISPATCH_KT(kt, __VA_ARGS__) }    \
  else { constexpr bool CLV = false; DISPATCH_KT(kt, __VA_ARGS__) }

// launch `grid` CTAs in clusters of `cl` (cl == 1: a plain launch)
template <class... KArgs, class... Args>
static cudaError_t launch_chain_kernel(void (*kern)(KArgs...), int grid, int cl, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kBlock);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = cl > 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// number of gene-chains resident at once: CTAs per SM x SMs, or the clusters of `cl` CTAs the device can co-schedule
template <class... KArgs>
static int resident_chains(void (*kern)(KArgs...), int kt, int cl, size_t smem) {
  const int ctas = sm_count() * ctas_per_sm(kt);
  if (cl <= 1) return ctas;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ctas / cl * cl));
  cfg.blockDim = dim3(kBlock);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = ctas / cl / 2;   // conservative: half of the slots
  }
  return std::max(1, std::min(n, ctas / cl));
}

extern "C" int32_t csplotch_log_prob_grad_dev(csplotch_batch *b, const double *theta_dev, const int32_t *gene_of_dev,
                                              int32_t Bn, double *lp_dev, double *grad_dev, int32_t jacobian,
                                              void *stream) {
  if (!b || !theta_dev || !lp_dev || Bn < 1) return fail(CSPLOTCH_E_INVALID, "bad argument");
  csplotch_model *m = b->m;
  const DevModel &dm = m->dm;
  cudaStream_t st = (cudaStream_t)stream;
  double *th_int = nullptr, *g_int = nullptr;
  CU_TRY(cudaMalloc(&th_int, (size_t)Bn * dm.Di * sizeof(double)));
  if (cudaMalloc(&g_int, (size_t)Bn * dm.Di * sizeof(double)) != cudaSuccess) {
    cudaFree(th_int);
    return fail(CSPLOTCH_E_NOMEM, "out of device memory");
  }
  const int nb = sm_count() * 8;
  k_to_internal<<<nb, 256, 0, st>>>(theta_dev, th_int, dm.imap, Bn, dm.D, dm.Di, 0.0);
  const int grid = std::min(Bn, sm_count() * ctas_per_sm(m->kt));
  cudaError_t e = cudaSuccess;
  DISPATCH_KT(m->kt, {
    e = set_smem(k_lpgrad<KT>, m->smem_bytes);
    if (e == cudaSuccess)
      k_lpgrad<KT><<<grid, kBlock, m->smem_bytes, st>>>(dm, b->counts, b->pm, b->ps, b->lgam, b->nnz, b->G, th_int,
                                                       gene_of_dev, Bn, lp_dev, g_int, jacobian);
  });
  if (e == cudaSuccess && grad_dev) {
    k_negate<<<nb, 256, 0, st>>>(g_int, (size_t)Bn * dm.Di);
    k_to_stan<<<nb, 256, 0, st>>>(g_int, grad_dev, dm.imap, Bn, dm.D, dm.Di);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(th_int);
  cudaFree(g_int);
  if (e != cudaSuccess) return fail(CSPLOTCH_E_CUDA, cudaGetErrorString(e));
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_log_prob_grad(csplotch_batch *b, const double *theta, const int32_t *gene_of, int32_t Bn,
                                          double *lp, double *grad, int32_t jacobian) {
  if (!b || !theta || !lp || Bn < 1) return fail(CSPLOTCH_E_INVALID, "bad argument");
  const DevModel &dm = b->m->dm;
  if (gene_of)
    for (int i = 0; i < Bn; ++i)
      if (gene_of[i] < 0 || gene_of[i] >= b->G) return fail(CSPLOTCH_E_INVALID, "gene_of out of range");
  DevBuf tmp;
  double *th = nullptr, *lpd = nullptr, *gd = nullptr;
  int *go = nullptr;
  CU_TRY(tmp.alloc(&th, (size_t)Bn * dm.D));
  CU_TRY(tmp.alloc(&lpd, (size_t)Bn));
  if (grad) CU_TRY(tmp.alloc(&gd, (size_t)Bn * dm.D));
  if (gene_of) {
    CU_TRY(tmp.alloc(&go, (size_t)Bn));
    CU_TRY(cudaMemcpy(go, gene_of, (size_t)Bn * sizeof(int), cudaMemcpyHostToDevice));
  }
  CU_TRY(cudaMemcpy(th, theta, (size_t)Bn * dm.D * sizeof(double), cudaMemcpyHostToDevice));
  int rc = csplotch_log_prob_grad_dev(b, th, go, Bn, lpd, gd, jacobian, nullptr);
  if (rc) return rc;
  CU_TRY(cudaMemcpy(lp, lpd, (size_t)Bn * sizeof(double), cudaMemcpyDeviceToHost));
  if (grad) CU_TRY(cudaMemcpy(grad, gd, (size_t)Bn * dm.D * sizeof(double), cudaMemcpyDeviceToHost));
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_leapfrog_fixed(csplotch_batch *b, const double *theta0, const double *p0,
                                           const double *inv_metric, const int32_t *gene_of, int32_t Bn, double eps,
                                           int32_t n_steps, double *theta_out, double *p_out, double *H_out) {
  if (!b || !theta0 || !p0 || !inv_metric || Bn < 1 || n_steps < 0) return fail(CSPLOTCH_E_INVALID, "bad argument");
  csplotch_model *m = b->m;
  const DevModel &dm = m->dm;
  DevBuf tmp;
  double *stan = nullptr, *q = nullptr, *p = nullptr, *g = nullptr, *mi = nullptr, *H = nullptr;
  int *go = nullptr;
  const size_t nS = (size_t)Bn * dm.D, nI = (size_t)Bn * dm.Di;
  CU_TRY(tmp.alloc(&stan, nS));
  CU_TRY(tmp.alloc(&q, nI));
  CU_TRY(tmp.alloc(&p, nI));
  CU_TRY(tmp.alloc(&g, nI));
  CU_TRY(tmp.alloc(&mi, nI));
  CU_TRY(tmp.alloc(&H, (size_t)Bn));
  if (gene_of) {
    CU_TRY(tmp.alloc(&go, (size_t)Bn));
    CU_TRY(cudaMemcpy(go, gene_of, (size_t)Bn * sizeof(int), cudaMemcpyHostToDevice));
  }
  const int nb = sm_count() * 8;
  CU_TRY(cudaMemcpy(stan, theta0, nS * sizeof(double), cudaMemcpyHostToDevice));
  k_to_internal<<<nb, 256>>>(stan, q, dm.imap, Bn, dm.D, dm.Di, 0.0);
  CU_TRY(cudaMemcpy(stan, p0, nS * sizeof(double), cudaMemcpyHostToDevice));
  k_to_internal<<<nb, 256>>>(stan, p, dm.imap, Bn, dm.D, dm.Di, 0.0);
  CU_TRY(cudaMemcpy(stan, inv_metric, nS * sizeof(double), cudaMemcpyHostToDevice));
  k_to_internal<<<nb, 256>>>(stan, mi, dm.imap, Bn, dm.D, dm.Di, 1.0);
  cudaError_t e = cudaSuccess;
  const int cl = choose_cluster(m, Bn);
  const DevModel dml = with_cluster(m, cl);
  if (m->warp_ok && m->forced_scope == 1) {
    DISPATCH_KT(m->kt, {
      e = set_smem(k_leapfrog_fixed_w<KT>, m->warp_smem_bytes);
      if (e == cudaSuccess)
        k_leapfrog_fixed_w<KT><<<std::min((Bn + kWarps - 1) / kWarps, warp_resident_ctas()), kBlock, m->warp_smem_bytes>>>(
            dm, b->counts, b->pm, b->ps, b->lgam, b->nnz, b->G, q, p, g, mi, go, Bn, eps, n_steps, H);
    });
  } else
  DISPATCH_KT_CL(m->kt, cl > 1, {
    e = set_smem(k_leapfrog_fixed<KT, CLV>, m->smem_bytes);
    if (e == cudaSuccess) {
      const int grid = std::min(Bn, resident_chains(k_leapfrog_fixed<KT, CLV>, m->kt, cl, m->smem_bytes)) * cl;
      e = launch_chain_kernel(k_leapfrog_fixed<KT, CLV>, grid, cl, m->smem_bytes, (cudaStream_t) nullptr, dml, b->counts,
                              b->pm, b->ps, b->lgam, b->nnz, b->G, q, p, g, mi, go, Bn, eps, n_steps, H);
    }
  });
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CSPLOTCH_E_CUDA, cudaGetErrorString(e));
  if (theta_out) {
    k_to_stan<<<nb, 256>>>(q, stan, dm.imap, Bn, dm.D, dm.Di);
    CU_TRY(cudaMemcpy(theta_out, stan, nS * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (p_out) {
    k_to_stan<<<nb, 256>>>(p, stan, dm.imap, Bn, dm.D, dm.Di);
    CU_TRY(cudaMemcpy(p_out, stan, nS * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (H_out) CU_TRY(cudaMemcpy(H_out, H, (size_t)Bn * sizeof(double), cudaMemcpyDeviceToHost));
  CU_TRY(cudaDeviceSynchronize());
  return CSPLOTCH_OK;
}

// Device-timed microbenchmark of the fused leapfrog (diagnostic: scripts/leap_bench.py, profiles/): Bn trajectories
// of n_steps fixed-size steps from deterministic pseudo-random states, `reps` timed launches after one warm-up;
// nothing crosses PCIe.  Work per launch = Bn * (1 gradient + n_steps leapfrogs).
__global__ void k_fill_hash(double *v, size_t n, const int *__restrict__ imap, int Di, uint32_t salt, double scale, double base) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + salt;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    const bool live = imap[i % (size_t)Di] >= 0;
    v[i] = live ? base + scale * ((double)h * (1.0 / 4294967296.0) - 0.5) : (base != 0.0 ? 1.0 : 0.0);
  }
}
extern "C" int32_t csplotch_leapfrog_bench(csplotch_batch *b, int32_t Bn, double eps, int32_t n_steps, int32_t reps,
                                           float *ms_avg) {
  if (!b || Bn < 1 || n_steps < 0 || reps < 1 || !ms_avg) return fail(CSPLOTCH_E_INVALID, "bad argument");
  csplotch_model *m = b->m;
  const DevModel &dm = m->dm;
  DevBuf tmp;
  double *q = nullptr, *p = nullptr, *g = nullptr, *mi = nullptr, *H = nullptr;
  const size_t nI = (size_t)Bn * dm.Di;
  CU_TRY(tmp.alloc(&q, nI));
  CU_TRY(tmp.alloc(&p, nI));
  CU_TRY(tmp.alloc(&g, nI));
  CU_TRY(tmp.alloc(&mi, nI));
  CU_TRY(tmp.alloc(&H, (size_t)Bn));
  const int nb = sm_count() * 8;
  k_fill_hash<<<nb, 256>>>(q, nI, dm.imap, dm.Di, 1u, 1.0, 0.0);
  k_fill_hash<<<nb, 256>>>(p, nI, dm.imap, dm.Di, 2u, 2.0, 0.0);
  k_fill_hash<<<nb, 256>>>(mi, nI, dm.imap, dm.Di, 3u, 0.5, 1.0);
  CU_TRY(cudaMemset(g, 0, nI * sizeof(double)));
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  cudaError_t e = cudaSuccess;
  const int cl = choose_cluster(m, Bn);
  const DevModel dml = with_cluster(m, cl);
  for (int r = 0; r <= reps && e == cudaSuccess; ++r) {
    if (r == 1) cudaEventRecord(e0);
    if (m->warp_ok && m->forced_scope == 1) {
      DISPATCH_KT(m->kt, {
        e = set_smem(k_leapfrog_fixed_w<KT>, m->warp_smem_bytes);
        if (e == cudaSuccess)
          k_leapfrog_fixed_w<KT><<<std::min((Bn + kWarps - 1) / kWarps, warp_resident_ctas()), kBlock, m->warp_smem_bytes>>>(
              dm, b->counts, b->pm, b->ps, b->lgam, b->nnz, b->G, q, p, g, mi, (const int *)nullptr, Bn, eps, n_steps, H);
      });
    } else
    DISPATCH_KT_CL(m->kt, cl > 1, {
      e = set_smem(k_leapfrog_fixed<KT, CLV>, m->smem_bytes);
      if (e == cudaSuccess) {
        const int grid = std::min(Bn, resident_chains(k_leapfrog_fixed<KT, CLV>, m->kt, cl, m->smem_bytes)) * cl;
        e = launch_chain_kernel(k_leapfrog_fixed<KT, CLV>, grid, cl, m->smem_bytes, (cudaStream_t) nullptr, dml,
                                b->counts, b->pm, b->ps, b->lgam, b->nnz, b->G, q, p, g, mi, (const int *)nullptr, Bn, eps,
                                n_steps, H);
      }
    });
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  cudaEventRecord(e1);
  if (e == cudaSuccess) e = cudaEventSynchronize(e1);
  float ms = 0.f;
  if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (e != cudaSuccess) return fail(CSPLOTCH_E_CUDA, cudaGetErrorString(e));
  *ms_avg = ms / (float)reps;
  return CSPLOTCH_OK;
}

// ------------------------------------------------------------------------------------------------
// host: sampler
// ------------------------------------------------------------------------------------------------
extern "C" void csplotch_nuts_cfg_default(csplotch_nuts_cfg *c) {
  if (!c) return;
  c->num_warmup = 1000; c->num_samples = 1000; c->num_chains = 4; c->max_depth = 10;
  c->delta = 0.8; c->gamma = 0.05; c->kappa = 0.75; c->t0 = 10;
  c->init_buffer = 75; c->term_buffer = 50; c->window = 25;
  c->init_radius = 2; c->stepsize = 1; c->seed = 0; c->keep_draws = 0;
}

extern "C" int32_t csplotch_sampler_create(csplotch_batch *b, const csplotch_nuts_cfg *cfg, csplotch_sampler **out) {
  if (!b || !cfg || !out) return fail(CSPLOTCH_E_INVALID, "null argument");
  if (cfg->num_chains < 1 || cfg->num_warmup < 0 || cfg->num_samples < 0) return fail(CSPLOTCH_E_INVALID, "bad chain / iteration counts");
  if (cfg->max_depth < 1 || cfg->max_depth > kMaxDepth) return fail(CSPLOTCH_E_INVALID, "max_depth must be in 1..12");
  csplotch_model *m = b->m;
  const DevModel &dm = m->dm;
  auto *s = new csplotch_sampler();
  s->b = b; s->cfg = *cfg;
  SamplerDev &sd = s->sd;
  memset(&sd, 0, sizeof(sd));
  const int W = b->G * cfg->num_chains;
  sd.W = W; sd.num_chains = cfg->num_chains; sd.gene_ids = b->gene_ids;
  sd.cfg = NutsConfig{cfg->num_warmup, cfg->num_samples, cfg->max_depth, cfg->delta, cfg->gamma, cfg->kappa, cfg->t0,
                      cfg->init_buffer, cfg->term_buffer, cfg->window, cfg->init_radius, cfg->stepsize, cfg->seed};
  sd.counts = b->counts; sd.pm = b->pm; sd.ps = b->ps; sd.lgam = b->lgam; sd.nnz = b->nnz;
  // gene-chains resident at once (warps, CTAs, or clusters of s->cl CTAs)
  s->warp_scope = choose_warp_scope(m, W);
  s->cl = s->warp_scope ? 1 : choose_cluster(m, W);
  s->dml = with_cluster(m, s->cl);
  if (s->warp_scope) {
    cudaError_t e0 = cudaSuccess;
    DISPATCH_KT(m->kt, { e0 = set_smem(k_nuts_advance_w<KT>, m->warp_smem_bytes); });
    if (e0 != cudaSuccess) { delete s; return fail(CSPLOTCH_E_CUDA, cudaGetErrorString(e0)); }
    s->grid = std::min(W, warp_resident_ctas() * kWarps);   // chains; the launch uses ceil(grid / 8) CTAs
  } else {
    int rc_chains = 0;
    DISPATCH_KT_CL(m->kt, s->cl > 1, {
      cudaError_t e0 = set_smem(k_nuts_advance<KT, CLV>, m->smem_bytes);
      if (e0 != cudaSuccess) { delete s; return fail(CSPLOTCH_E_CUDA, cudaGetErrorString(e0)); }
      rc_chains = resident_chains(k_nuts_advance<KT, CLV>, m->kt, s->cl, m->smem_bytes);
    });
    s->grid = std::min(W, rc_chains);
  }
  sd.pool_vecs = 3 * cfg->max_depth + 8;
  const size_t Di = dm.Di, ns = (size_t)cfg->num_samples;
  // resident CTAs are also bounded by HBM: every CTA slot owns pool_vecs + 6 D-vectors of workspace
  // (c4: 284 MB).  Per-chain state and outputs come first; the slots share 90 % of what is left.
  {
    size_t free_b = 0, total_b = 0;
    CU_TRY(cudaMemGetInfo(&free_b, &total_b));
    const size_t per_chain = sizeof(ChainScalars) + (4 * Di + ns * (7 + dm.nsmall_true) + 6 * (size_t)dm.N +
                                                     (cfg->keep_draws ? ns * (size_t)dm.D : 0)) * sizeof(double);
    const size_t need = (size_t)W * per_chain;
    const size_t per_slot = (size_t)(sd.pool_vecs + 6) * Di * sizeof(double);
    if (need + per_slot > free_b) {
      delete s;
      return fail(CSPLOTCH_E_NOMEM, "not enough device memory for this many gene-chains: sample fewer genes per batch");
    }
    const size_t max_slots = (size_t)(0.9 * (double)(free_b - need)) / per_slot;
    s->grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)s->grid, max_slots));
  }
  if (s->warp_scope) s->grid = s->grid >= kWarps ? s->grid / kWarps * kWarps : kWarps;   // whole CTAs of eight chain slots
  sd.n_slots = s->grid;
  cudaError_t e = cudaSuccess;
  auto A = [&](auto **p, size_t n) { if (e == cudaSuccess) e = s->buf.alloc(p, n); };
  A(&sd.cs, (size_t)W);
  A(&sd.q_cur, (size_t)W * Di);
  A(&sd.minv, (size_t)W * Di);
  A(&sd.wm, (size_t)W * Di);
  A(&sd.wm2, (size_t)W * Di);
  A(&sd.ws, (size_t)sd.n_slots * (sd.pool_vecs + 6) * Di);
  A(&sd.diag, (size_t)W * ns * 7);
  A(&sd.small, (size_t)W * ns * dm.nsmall_true);
  A(&sd.summ, (size_t)W * 6 * dm.N);
  if (cfg->keep_draws) A(&sd.theta, (size_t)W * ns * dm.D);
  A(&sd.queue, 1);
  A(&s->order_dev, (size_t)W);
  sd.order = s->order_dev;
  A(&sd.slot_ns, (size_t)sd.n_slots * s->cl);
  if (e != cudaSuccess) {
    delete s;
    return fail(e == cudaErrorMemoryAllocation ? CSPLOTCH_E_NOMEM : CSPLOTCH_E_CUDA,
                std::string("sampler allocation failed: ") + cudaGetErrorString(e));
  }
  CU_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreate(&s->ev0));
  CU_TRY(cudaEventCreate(&s->ev1));
  CU_TRY(cudaMemsetAsync(sd.cs, 0, (size_t)W * sizeof(ChainScalars), s->stream));
  CU_TRY(cudaMemsetAsync(sd.q_cur, 0, (size_t)W * Di * sizeof(double), s->stream));
  CU_TRY(cudaMemsetAsync(sd.ws, 0, (size_t)sd.n_slots * (sd.pool_vecs + 6) * Di * sizeof(double), s->stream));
  CU_TRY(cudaMemsetAsync(sd.summ, 0, (size_t)W * 6 * dm.N * sizeof(double), s->stream));
  CU_TRY(cudaMemsetAsync(sd.diag, 0, (size_t)W * ns * 7 * sizeof(double), s->stream));
  CU_TRY(cudaMemsetAsync(sd.small, 0, (size_t)W * ns * dm.nsmall_true * sizeof(double), s->stream));
  CU_TRY(cudaStreamSynchronize(s->stream));
  s->nnz_host.resize((size_t)b->G);
  CU_TRY(cudaMemcpy(s->nnz_host.data(), b->nnz, (size_t)b->G * sizeof(double), cudaMemcpyDeviceToHost));
  *out = s;
  return CSPLOTCH_OK;
}

extern "C" void csplotch_sampler_destroy(csplotch_sampler *s) { delete s; }

// Queue order of the next launch: longest expected gene-chain first (LPT), so that the deepest trees start at the
// beginning of the launch instead of bounding its tail.  Expected cost = n_leapfrog__ of the chain's most recent
// transition (the tree depth follows the adapted step size, so it persists), or, before the first transition, the
// number of non-zero counts of the gene.  The order never changes a result: every random number is keyed by
// (seed, gene id, chain, iteration, purpose).  CSPLOTCH_QUEUE_ORDER=index restores index order (A/B measurements).
static int build_order(csplotch_sampler *s) {
  const int W = s->sd.W, nch = s->sd.num_chains;
  s->order_host.resize((size_t)W);
  for (int w = 0; w < W; ++w) s->order_host[w] = w;
  const char *mode = std::getenv("CSPLOTCH_QUEUE_ORDER");
  if (!(mode && std::string(mode) == "index")) {
    s->cs_host.resize((size_t)W);
    CU_TRY(cudaMemcpyAsync(s->cs_host.data(), s->sd.cs, (size_t)W * sizeof(ChainScalars), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(cudaStreamSynchronize(s->stream));
    std::vector<double> cost((size_t)W);
    for (int w = 0; w < W; ++w) {
      const ChainScalars &c = s->cs_host[w];
      cost[w] = (c.initialized && c.iter > 0) ? 1e9 + (double)c.last_nleap : s->nnz_host[w / nch];
    }
    std::stable_sort(s->order_host.begin(), s->order_host.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  }
  CU_TRY(cudaMemcpyAsync(s->order_dev, s->order_host.data(), (size_t)W * sizeof(int), cudaMemcpyHostToDevice, s->stream));
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_sampler_init(csplotch_sampler *s, const double *theta_init) {
  if (!s) return fail(CSPLOTCH_E_INVALID, "null argument");
  const DevModel &dm = s->b->m->dm;
  if (theta_init) {
    DevBuf tmp;
    double *stan = nullptr;
    const size_t n = (size_t)s->sd.W * dm.D;
    CU_TRY(tmp.alloc(&stan, n));
    CU_TRY(cudaMemcpy(stan, theta_init, n * sizeof(double), cudaMemcpyHostToDevice));
    k_to_internal<<<sm_count() * 8, 256, 0, s->stream>>>(stan, s->sd.q_cur, dm.imap, s->sd.W, dm.D, dm.Di, 0.0);
    CU_TRY(cudaStreamSynchronize(s->stream));
    s->sd.have_init = 1;
  }
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_sampler_advance(csplotch_sampler *s, int32_t n_iter, int32_t sync) {
  if (!s || n_iter < 0) return fail(CSPLOTCH_E_INVALID, "bad argument");
  csplotch_model *m = s->b->m;
  {
    const int rc = build_order(s);
    if (rc) return rc;
  }
  CU_TRY(cudaMemsetAsync(s->sd.queue, 0, sizeof(int), s->stream));
  CU_TRY(cudaEventRecord(s->ev0, s->stream));
  cudaError_t e = cudaSuccess;
  if (s->warp_scope) {
    DISPATCH_KT(m->kt, {
      e = set_smem(k_nuts_advance_w<KT>, m->warp_smem_bytes);
      if (e == cudaSuccess)
        k_nuts_advance_w<KT><<<s->grid / kWarps, kBlock, m->warp_smem_bytes, s->stream>>>(s->dml, s->sd, n_iter);
    });
  } else
  DISPATCH_KT_CL(m->kt, s->cl > 1, {
    e = set_smem(k_nuts_advance<KT, CLV>, m->smem_bytes);
    if (e == cudaSuccess)
      e = launch_chain_kernel(k_nuts_advance<KT, CLV>, s->grid * s->cl, s->cl, m->smem_bytes, s->stream, s->dml, s->sd,
                              n_iter);
  });
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CSPLOTCH_E_CUDA, cudaGetErrorString(e));
  CU_TRY(cudaEventRecord(s->ev1, s->stream));
  s->timed = true;
  s->launches += 1;
  if (sync) CU_TRY(cudaStreamSynchronize(s->stream));
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_sampler_sync(csplotch_sampler *s) {
  if (!s) return fail(CSPLOTCH_E_INVALID, "null argument");
  CU_TRY(cudaStreamSynchronize(s->stream));
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_sampler_last_ms(csplotch_sampler *s, float *ms) {
  if (!s || !ms || !s->timed) return fail(CSPLOTCH_E_INVALID, "no advance has been timed");
  CU_TRY(cudaEventSynchronize(s->ev1));
  CU_TRY(cudaEventElapsedTime(ms, s->ev0, s->ev1));
  return CSPLOTCH_OK;
}

static int fetch_cs(csplotch_sampler *s, std::vector<ChainScalars> &h) {
  h.resize(s->sd.W);
  CU_TRY(cudaStreamSynchronize(s->stream));
  CU_TRY(cudaMemcpy(h.data(), s->sd.cs, h.size() * sizeof(ChainScalars), cudaMemcpyDeviceToHost));
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_sampler_counters(csplotch_sampler *s, int64_t *n_leapfrog, int64_t *n_grad,
                                             int64_t *n_divergent, int64_t *n_launches) {
  if (!s) return fail(CSPLOTCH_E_INVALID, "null argument");
  std::vector<ChainScalars> h;
  int rc = fetch_cs(s, h);
  if (rc) return rc;
  int64_t a = 0, b = 0, c = 0;
  for (auto &x : h) { a += x.n_leapfrog_total; b += x.n_grad_total; c += x.n_divergent; }
  if (n_leapfrog) *n_leapfrog = a;
  if (n_grad) *n_grad = b;
  if (n_divergent) *n_divergent = c;
  if (n_launches) *n_launches = s->launches;
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_sampler_profile(csplotch_sampler *s, int64_t *out8) {
  if (!s || !out8) return fail(CSPLOTCH_E_INVALID, "null argument");
  std::vector<ChainScalars> h;
  int rc = fetch_cs(s, h);
  if (rc) return rc;
  for (int i = 0; i < 8; ++i) out8[i] = 0;
  for (auto &x : h) {
    out8[0] += x.cyc_eval; out8[1] += x.cyc_merge; out8[2] += x.cyc_copy; out8[3] += x.cyc_other;
    out8[4] += x.n_merge; out8[5] += x.n_copy; out8[6] += x.cyc_pro; out8[7] += x.cyc_fin;
  }
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_sampler_layout(csplotch_sampler *s, int32_t *resident_chains_out, int32_t *cluster_size) {
  if (!s) return fail(CSPLOTCH_E_INVALID, "null argument");
  if (resident_chains_out) *resident_chains_out = s->sd.n_slots;
  if (cluster_size) *cluster_size = s->warp_scope ? 0 : s->cl;   // 0: one gene-chain per warp
  return CSPLOTCH_OK;
}

// busy time of every resident CTA in the last launch: the spread between the mean and the maximum is the launch tail
extern "C" int32_t csplotch_sampler_slot_ns(csplotch_sampler *s, int64_t *ns, int32_t cap, int32_t *n_slots) {
  if (!s || !n_slots) return fail(CSPLOTCH_E_INVALID, "null argument");
  const int n_ctas = s->sd.n_slots * s->cl;
  *n_slots = n_ctas;
  if (ns) {
    if (cap < n_ctas) return fail(CSPLOTCH_E_INVALID, "buffer too small");
    CU_TRY(cudaStreamSynchronize(s->stream));
    CU_TRY(cudaMemcpy(ns, s->sd.slot_ns, (size_t)n_ctas * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_sampler_status(csplotch_sampler *s, int32_t *status, int32_t *iters_done) {
  if (!s) return fail(CSPLOTCH_E_INVALID, "null argument");
  std::vector<ChainScalars> h;
  int rc = fetch_cs(s, h);
  if (rc) return rc;
  for (size_t i = 0; i < h.size(); ++i) {
    if (status) status[i] = h[i].status;
    if (iters_done) iters_done[i] = h[i].iter;
  }
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_sampler_get_draws(csplotch_sampler *s, double *diag, double *small, int32_t *n_draws) {
  if (!s) return fail(CSPLOTCH_E_INVALID, "null argument");
  const DevModel &dm = s->b->m->dm;
  const size_t ns = (size_t)s->cfg.num_samples, W = (size_t)s->sd.W;
  CU_TRY(cudaStreamSynchronize(s->stream));
  if (diag) CU_TRY(cudaMemcpy(diag, s->sd.diag, W * ns * 7 * sizeof(double), cudaMemcpyDeviceToHost));
  if (small) CU_TRY(cudaMemcpy(small, s->sd.small, W * ns * dm.nsmall_true * sizeof(double), cudaMemcpyDeviceToHost));
  if (n_draws) {
    std::vector<ChainScalars> h;
    int rc = fetch_cs(s, h);
    if (rc) return rc;
    for (size_t i = 0; i < W; ++i) n_draws[i] = std::max(0, h[i].iter - s->cfg.num_warmup);
  }
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_sampler_get_theta_draws(csplotch_sampler *s, double *theta) {
  if (!s || !theta) return fail(CSPLOTCH_E_INVALID, "null argument");
  if (!s->sd.theta) return fail(CSPLOTCH_E_INVALID, "sampler was created without keep_draws");
  const DevModel &dm = s->b->m->dm;
  CU_TRY(cudaStreamSynchronize(s->stream));
  CU_TRY(cudaMemcpy(theta, s->sd.theta, (size_t)s->sd.W * s->cfg.num_samples * dm.D * sizeof(double), cudaMemcpyDeviceToHost));
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_sampler_get_theta_draws_range(csplotch_sampler *s, int32_t g0, int32_t n_genes, double *theta) {
  if (!s || !theta) return fail(CSPLOTCH_E_INVALID, "null argument");
  if (!s->sd.theta) return fail(CSPLOTCH_E_INVALID, "sampler was created without keep_draws");
  if (g0 < 0 || n_genes < 0 || g0 + n_genes > s->b->G) return fail(CSPLOTCH_E_INVALID, "gene range outside the batch");
  const DevModel &dm = s->b->m->dm;
  const size_t per_gene = (size_t)s->cfg.num_chains * s->cfg.num_samples * dm.D;
  CU_TRY(cudaStreamSynchronize(s->stream));
  if (n_genes)
    CU_TRY(cudaMemcpy(theta, s->sd.theta + (size_t)g0 * per_gene, (size_t)n_genes * per_gene * sizeof(double), cudaMemcpyDeviceToHost));
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_sampler_get_state(csplotch_sampler *s, double *theta, double *stepsize, double *inv_metric) {
  if (!s) return fail(CSPLOTCH_E_INVALID, "null argument");
  const DevModel &dm = s->b->m->dm;
  const size_t W = (size_t)s->sd.W;
  CU_TRY(cudaStreamSynchronize(s->stream));
  DevBuf tmp;
  double *stan = nullptr;
  CU_TRY(tmp.alloc(&stan, W * dm.D));
  const int nb = sm_count() * 8;
  if (theta) {
    k_to_stan<<<nb, 256, 0, s->stream>>>(s->sd.q_cur, stan, dm.imap, (int)W, dm.D, dm.Di);
    CU_TRY(cudaStreamSynchronize(s->stream));
    CU_TRY(cudaMemcpy(theta, stan, W * dm.D * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (inv_metric) {
    k_to_stan<<<nb, 256, 0, s->stream>>>(s->sd.minv, stan, dm.imap, (int)W, dm.D, dm.Di);
    CU_TRY(cudaStreamSynchronize(s->stream));
    CU_TRY(cudaMemcpy(inv_metric, stan, W * dm.D * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (stepsize) {
    std::vector<ChainScalars> h;
    int rc = fetch_cs(s, h);
    if (rc) return rc;
    for (size_t i = 0; i < W; ++i) stepsize[i] = h[i].eps;
  }
  return CSPLOTCH_OK;
}

extern "C" int32_t csplotch_sampler_get_spot_summaries(csplotch_sampler *s, double *ll_mean, double *ll_std,
                                                       double *la_mean, double *la_std, double *psi_mean,
                                                       double *psi_std) {
  if (!s) return fail(CSPLOTCH_E_INVALID, "null argument");
  const DevModel &dm = s->b->m->dm;
  const int G = s->b->G;
  const size_t n = (size_t)G * dm.N;
  double *host[6] = {ll_mean, ll_std, la_mean, la_std, psi_mean, psi_std};
  double *dev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  DevBuf tmp;
  int *perm_dev = nullptr;
  CU_TRY(tmp.upload(&perm_dev, s->b->m->perm));
  for (int i = 0; i < 6; ++i)
    if (host[i]) CU_TRY(tmp.alloc(&dev[i], n));
  k_summ_finalize<<<sm_count() * 8, 256, 0, s->stream>>>(s->sd.summ, s->sd.cs, perm_dev, G, s->cfg.num_chains, dm.N,
                                                        s->cfg.num_warmup, dev[0], dev[1], dev[2], dev[3], dev[4], dev[5]);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(s->stream));
  for (int i = 0; i < 6; ++i)
    if (host[i]) CU_TRY(cudaMemcpy(host[i], dev[i], n * sizeof(double), cudaMemcpyDeviceToHost));
  return CSPLOTCH_OK;
}

// write_array is host-side glue over small per-draw blocks; the heavy part (log_lambda) is
// evaluated on the device through the same load_small + spot loop as the sampler.
template <int KT>
__global__ void __launch_bounds__(kBlock, kCtasPerSm<KT>)
k_write_array(const __grid_constant__ DevModel m_in, const double *__restrict__ pm, const double *__restrict__ ps, int G,
              const double *__restrict__ theta_int, const int *__restrict__ gene_of, const int *__restrict__ perm,
              int Bn, double *out, int n_out) {
  // the model descriptor is read all over the per-evaluation prologue / epilogue: a shared-memory copy turns
  // those dependent generic loads from the parameter bank into LDS
  __shared__ DevModel s_model;
  stage_model(s_model, m_in);
  const DevModel &m = s_model;
  DevOps<KT> ops(m);
  bind_smem(ops, m, smem_d);
  ops.jacobian = 0;
  const int CK = m.C * m.K;
  for (int b = blockIdx.x; b < Bn; b += gridDim.x) {
    const int g = gene_of ? gene_of[b] : (b % G);
    ops.gn.pm = pm ? pm + (size_t)g * m.K : nullptr;
    ops.gn.ps = ps ? ps + (size_t)g * m.K : nullptr;
    const double *q = theta_int + (size_t)b * m.Di;
    double *o = out + (size_t)b * n_out;
    Scalars sc;
    ops.load_small(q, sc);
    // constrained parameters in declaration order (Stan index space)
    const int so_beta = m.car ? m.N : 0, so_scal = so_beta + m.nbeta_true;
    for (int e = threadIdx.x; e < m.Di; e += kBlock) {
      const int si = m.imap[e];
      if (si < 0) continue;
      double v = q[e];
      if (si >= so_scal && si < so_scal + m.nscal) {
        const int sidx = si - so_scal;
        if (sidx == m.s_tau) v = sc.tau;
        else if (sidx == m.s_a) v = sc.a;
        else if (sidx == m.s_sigma) v = sc.sigma;
        else if (sidx == m.s_s2) v = sc.s2;
        else if (sidx == m.s_s3) v = sc.s3;
        else if (sidx == m.s_theta) v = sc.theta;
        else if (sidx == m.s_phi) v = sc.phi;
      }
      // beta_raw is written column-major over (i,j,k): first index fastest
      int oi = si;
      if (si >= so_beta && si < so_scal) {
        const int internal = e - m.o_small;
        const int i = internal / CK, rem = internal - i * CK, j = rem / m.K, k = rem - j * m.K;
        oi = so_beta + i + m.L * (j + m.C * k);   // k < Ktrue here: padded entries have imap = -1
      }
      o[oi] = v;
    }
    // transformed parameters: log_lambda (original spot order), beta_level_1..3 (column-major)
    double *oll = o + m.D;
    const double sig03 = 0.3 * sc.sigma;
    for (int j = threadIdx.x; j < m.N; j += kBlock) {
      const double *bj = ops.Bs() + (m.bot0 * m.C) * m.K + (m.key[j] & 0xFFFFFF) * m.K;
      double ll = 0.0;
      if (KT == 1) ll = bj[0];
      else {
        const double *Ej = E_ptr(m, j);
        for (int k = 0; k < m.K; ++k) ll += bj[k] * Ej[(size_t)k * kTile];
      }
      if (m.car) ll += q[m.o_psi + j];
      ll += sig03 * q[m.o_noise + j];
      oll[perm[j]] = ll;
    }
    if (m.nb)
      for (int j = threadIdx.x; j < m.N; j += kBlock) oll[m.N + j] = sc.phi;  // transformed parameter phi_arr
    double *ob = oll + m.N + (m.nb ? m.N : 0);
    for (int idx = threadIdx.x; idx < m.nbeta; idx += kBlock) {
      const int i = idx / CK, rem = idx - i * CK, j = rem / m.K, k = rem - j * m.K;
      if (k >= m.Ktrue) continue;
      const int CKt = m.C * m.Ktrue;
      int lev0, nr, off;
      if (i < m.L1) { lev0 = 0; nr = m.L1; off = 0; }
      else if (i < m.L1 + m.L2) { lev0 = m.L1; nr = m.L2; off = m.L1 * CKt; }
      else { lev0 = m.L1 + m.L2; nr = m.L3; off = (m.L1 + m.L2) * CKt; }
      ob[off + (i - lev0) + nr * (j + m.C * k)] = ops.Bs()[idx];
    }
    __syncthreads();
  }
}

extern "C" int32_t csplotch_write_array(csplotch_batch *b, const double *theta, const int32_t *gene_of, int32_t Bn,
                                        double *out) {
  if (!b || !theta || !out || Bn < 1) return fail(CSPLOTCH_E_INVALID, "bad argument");
  csplotch_model *m = b->m;
  const DevModel &dm = m->dm;
  DevBuf tmp;
  double *stan = nullptr, *q = nullptr, *o = nullptr;
  int *go = nullptr, *perm_dev = nullptr;
  CU_TRY(tmp.alloc(&stan, (size_t)Bn * dm.D));
  CU_TRY(tmp.alloc(&q, (size_t)Bn * dm.Di));
  CU_TRY(tmp.alloc(&o, (size_t)Bn * m->n_outputs));
  CU_TRY(tmp.upload(&perm_dev, m->perm));
  if (gene_of) {
    CU_TRY(tmp.alloc(&go, (size_t)Bn));
    CU_TRY(cudaMemcpy(go, gene_of, (size_t)Bn * sizeof(int), cudaMemcpyHostToDevice));
  }
  CU_TRY(cudaMemcpy(stan, theta, (size_t)Bn * dm.D * sizeof(double), cudaMemcpyHostToDevice));
  k_to_internal<<<sm_count() * 8, 256>>>(stan, q, dm.imap, Bn, dm.D, dm.Di, 0.0);
  cudaError_t e = cudaSuccess;
  DISPATCH_KT(m->kt, {
    e = set_smem(k_write_array<KT>, m->smem_bytes);
    if (e == cudaSuccess)
      k_write_array<KT><<<std::min(Bn, sm_count() * ctas_per_sm(m->kt)), kBlock, m->smem_bytes>>>(dm, b->pm, b->ps, b->G, q, go, perm_dev,
                                                                                Bn, o, m->n_outputs);
  });
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CSPLOTCH_E_CUDA, cudaGetErrorString(e));
  CU_TRY(cudaMemcpy(out, o, (size_t)Bn * m->n_outputs * sizeof(double), cudaMemcpyDeviceToHost));
  return CSPLOTCH_OK;
}
