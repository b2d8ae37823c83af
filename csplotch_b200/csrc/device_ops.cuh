// device_ops.cuh — CTA-collective vector operations of one gene-chain (sm_100a).
//
// One CTA owns one gene-chain.  Every function here is called by ALL threads of the CTA with
// identical (uniform) scalar arguments and returns identical scalars to all of them (block
// reductions are broadcast through shared memory in a fixed order), so the control flow in
// nuts_core.h stays CTA-uniform.
//
// The fused log-density + gradient restates (reference files under
// /root/reference/stan/):
//   transformed parameters{}  splotch_stan_model.stan:100-181, comp_…:110-215  -> load_small()
//   log_lambda + likelihood   splotch_stan_model.stan:135-180,209-222, comp_…:155-206,267-280,
//                             _csr.stan:254-271                                  -> spot loop
//   sparse_car_lpdf           splotch_stan_model.stan:4-17, _csr.stan:4-25       -> CSR gather + eig loop
//   priors                    splotch_stan_model.stan:184-207                    -> finish()
// Internal vector layout (all vectors of a chain, HBM):  [ psi(Npad) | noise(Npad) | small ]
// small = beta_raw[(i*C+j)*K+k] | tau | a | sigma | sigma_level_2 | sigma_level_3 | theta.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fastmath.h"
#include "nuts_core.h"

namespace csp {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;
constexpr int kRedMax = 16;  // doubles per block reduction
constexpr int kTile = kBlock;    // spots per tile of the staged pipeline: thread t <-> spot t of the tile
constexpr int kMaxEll = 8;       // widest ELL adjacency the tiled path takes (Visium hex: 6, ST / HD grid: 4)
// The CAR log-determinant sum_i m_i log(1 - a lambda_i) and its derivative in a, -sum_i m_i lambda_i / (1 - a lambda_i),
// over GROUPS of eight eigenvalues of equal multiplicity: with u_i = 1 - a lambda_i in (0, 2],
//   sum_i log u_i = log(prod u_i)          and          sum_i lambda_i / u_i = (sum_i lambda_i prod_{l != i} u_l) / prod u_i,
// so a group costs ONE log and ONE division (and a depth-3 product tree) instead of eight log1p and eight divisions
// -- fp64 log1p alone is ~250 instructions on this GPU, more than the rest of a spot's work, and ST / HD designs
// have about as many distinct eigenvalues as spots.  prod u_i stays within [(1-a)^8, 256]: no over/underflow for
// any a the sampler can reach short of a == 1 (then log 0 = -inf as before); relative error of the product 8 ulp,
// i.e. ~1e-15 absolute per group in the log.
constexpr int kEigGroup = 8;
__device__ __forceinline__ void eig_group_terms(const double *__restrict__ lam8, double mult, double a, double &ldet,
                                                double &dldet) {
  const double2 l01 = __ldg(reinterpret_cast<const double2 *>(lam8));
  const double2 l23 = __ldg(reinterpret_cast<const double2 *>(lam8) + 1);
  const double2 l45 = __ldg(reinterpret_cast<const double2 *>(lam8) + 2);
  const double2 l67 = __ldg(reinterpret_cast<const double2 *>(lam8) + 3);
  const double lam[8] = {l01.x, l01.y, l23.x, l23.y, l45.x, l45.y, l67.x, l67.y};
  double p[4], sn[4];  // per pair: product of u, sum_i lambda_i prod_{l != i} u_l
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double u0 = fma(-a, lam[2 * i], 1.0), u1 = fma(-a, lam[2 * i + 1], 1.0);
    p[i] = u0 * u1;
    sn[i] = fma(lam[2 * i], u1, lam[2 * i + 1] * u0);
  }
  const double pa = p[0] * p[1], sa = fma(sn[0], p[1], sn[1] * p[0]);
  const double pb = p[2] * p[3], sb = fma(sn[2], p[3], sn[3] * p[2]);
  const double P = pa * pb, S = fma(sa, pb, sb * pa);
  ldet += mult * log(P);
  dldet -= mult * (S / P);
}

struct DevModel {
  int family, N, Npad, C, K, L1, L2, L3, L, bot0, nbeta, nscal, nsmall, nsmall_pad;
  int Ktrue, nbeta_true, nsmall_true;  // K is the padded (template) cell-type count; *_true are Stan's
  int zi, car;
  int nb;  // comp only: 0 Poisson/ZIP, 1 NB/ZINB exactly as the reference writes it (ZINB term x N), 2 per-spot ZINB term
  int D;   // Stan dimension
  int Di;  // internal dimension (padded)
  int o_psi, o_noise, o_small;
  int s_tau, s_a, s_sigma, s_s2, s_s3, s_theta, s_phi;  // index inside the scalar part, -1 if absent
  int n_eig;
  int Nc;      // row stride of the counts matrix (int32 elements) = Npad
  int NT;      // tiles of kTile spots; Npad = NT * kTile
  int tiled;   // 1: every neighbour lies within `look` tiles and degree <= kMaxEll
  int look;    // 1 or 2: how many tiles a neighbour may be away (Visium / ST: 1; 317-wide HD rows: 2)
  int ell_w;   // ELL width (max degree), 0 without CAR
  int stage_bytes;
  const double *lsf;     // [Npad] log(size_factors), 0 in the padding
  const int *key;        // [Npad] bits 0-23 bottom-level bin (tissue_mapping[t]-1)*C + D_j-1, bits 24-27 degree; -1 = padding
  const int *rowptr;     // [N+1] CSR of the symmetric adjacency (internal spot ids)
  const int *col;        // [2 W_n]
  const unsigned short *ell;  // [NT][ell_w][kTile] ring positions of the neighbours (id & rmask); rmask+1 = none
  const double *E;       // [NT][K][kTile]  (tile-major, cell type, spot-in-tile): coalesced per k
  const double *eigu;    // [n_eig] distinct eigenvalues
  const double *eigm;    // [n_eig] multiplicities
  int n_eigg;            // groups of kEigGroup distinct eigenvalues of equal multiplicity (padded with lambda = 0)
  const double *eigg;    // [n_eigg][kEigGroup]
  const double *eiggm;   // [n_eigg] the multiplicity of each group
  const int *map2;       // [L2] 0-based level-1 row of each level-2 row
  const int *map3;       // [L3] 0-based level-2 row of each level-3 row
  const int *imap;       // [Di] internal index -> Stan index (-1 = padding)
  // ---- one gene-chain per thread-block CLUSTER (cl > 1: long designs; the sampler and the trajectory kernels) ----
  // Rank r of the cluster evaluates tiles [cl_t0[r], cl_t0[r+1]) and owns the same element ranges of every D-vector
  // (rank 0 also the small block, and fewer tiles for it); psi of the `look` tiles either side of a rank's range is
  // drifted redundantly into halo cells behind the ring, which the cluster's own ELL table addresses.
  int cl;                // CTAs per gene-chain: 1, 2, 4 or 8
  int cl_t0[9];          // tile boundaries of the ranks
  const unsigned short *ell_cl;  // [NT][ell_w][kTile] like ell, neighbours outside the owner's tile range -> halo cells
};
constexpr int kMaxCluster = 8;

struct GeneDev {
  const int *counts;   // [N] internal spot order
  const double *pm;    // [K] beta_prior_mean (comp) or nullptr
  const double *ps;    // [K] beta_prior_std
  double lgam;         // sum_{y>0} lgamma(y+1)   (kept only under `target +=`, i.e. zi=1)
  double nnz;          // #{y>0}
};

// the one dynamic shared-memory array of every kernel in this library; indexing it directly (instead of
// through generic pointers kept in structs) lets ptxas emit LDS/STS/ATOMS with 32-bit addresses
extern __shared__ __align__(16) double smem_d[];

struct Scalars {
  double tau, a, sigma, s2, s3, theta;
  double jac;          // sum of log-Jacobians
  double log_tau;      // = the unconstrained value itself
  double log1m_theta;  // log(1 - theta), shared with the log-Jacobian of theta
  double log_theta;
  double phi, log_phi; // NB overdispersion (comp_splotch_stan_model.stan:104), 0 unless nb
};

// digamma by upward recurrence to x >= 10 and the asymptotic series (abs. error < 1e-15)
__device__ __forceinline__ double digamma_d(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}

// x = inv_logit(u) together with log x and log(1 - x) from ONE exp and ONE log1p
// (log_inv_logit(u) = min(u,0) - log1p(e^{-|u|}), log1m_inv_logit(u) = min(-u,0) - log1p(e^{-|u|}))
__device__ __forceinline__ void logit_triple(double u, double &x, double &log_x, double &log1m_x) {
  const double e = exp(-fabs(u));
  const double l = log1p(e);
  const double r = 1.0 / (1.0 + e);
  x = u >= 0 ? r : e * r;
  log_x = fmin(u, 0.0) - l;
  log1m_x = fmin(-u, 0.0) - l;
}

__device__ __forceinline__ double inv_logit_d(double u) {
  if (u < 0) {
    const double e = exp(u);
    return e / (1.0 + e);
  }
  return 1.0 / (1.0 + exp(-u));
}
__device__ __forceinline__ double log_inv_logit_d(double u) { return u < 0 ? u - log1p(exp(u)) : -log1p(exp(-u)); }
__device__ __forceinline__ double log1m_inv_logit_d(double u) { return u > 0 ? -u - log1p(exp(-u)) : -log1p(exp(u)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of NV values; result identical in every thread
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *red) {
  static_assert(NV <= kRedMax, "raise kRedMax");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double s = warp_sum(v[i]);
    if (lane == 0) red[warp * kRedMax + i] = s;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w * kRedMax + i];
    v[i] = s;
  }
  __syncthreads();
}


// ---- sm_100a async-copy plumbing: 1-D bulk copies (TMA engine, SASS UBLKCP) completing on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
#ifdef CSP_WAIT_HINT
  // try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or ~the hint
  // elapses) instead of burning issue slots of the other CTA on the SM in a polling loop
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
  } while (!ok);
#else
  // plain try_wait (hardware-default suspend window) in a loop
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
#endif
}
// CTA-wide "everyone has done X" barriers: one arrival per WARP (lane 0 after __syncwarp, which orders the
// warp's shared-memory accesses before the release-arrive) unless CSP_ARRIVE_ALL restores one per thread
#ifdef CSP_ARRIVE_ALL
constexpr uint32_t kArriveCount = 256;
__device__ __forceinline__ void mbar_arrive_cta(uint64_t *bar) { mbar_arrive(bar); }
#else
constexpr uint32_t kArriveCount = 8;
__device__ __forceinline__ void mbar_arrive_cta(uint64_t *bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
#endif

// E[j][k] in the tile-major layout [NT][K][kTile]
__device__ __forceinline__ const double *E_ptr(const DevModel &m, int j) {
  return m.E + ((size_t)(j / kTile) * m.K) * kTile + (j % kTile);
}

// Sum KP per-lane values over the 32 lanes of a warp with KP/2 + KP/4 + ... + log2(32/KP)+... shuffles
// (transposed butterfly).  On return v[0] of lane l holds the total of value index bin_of_lane<KP>(l).
template <int KP>
__device__ __forceinline__ void warp_transpose_sum(double (&v)[KP]) {
  const int lane = threadIdx.x & 31;
  int off = 16;
#pragma unroll
  for (int n = KP; n > 1; n >>= 1) {
    const int half = n >> 1;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < KP / 2; ++i) {
      if (i < half) {
        const double send = upper ? v[i] : v[i + half];
        const double keep = upper ? v[i + half] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    off >>= 1;
  }
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}
template <int KP>
__device__ __forceinline__ int bin_of_lane(int lane) {
  // KP = 16: (lane >> 1) & 15; 8: (lane >> 2) & 7; 4: (lane >> 3) & 3; 2: lane >> 4; 1: 0
  int sh = 0, n = KP;
  while (n < 32) { n <<= 1; ++sh; }
  return KP == 1 ? 0 : ((lane >> sh) & (KP - 1));
}
template <int KP>
__device__ __forceinline__ bool lane_owns_bin(int lane) {
  int sh = 0, n = KP;
  while (n < 32) { n <<= 1; ++sh; }
  return (lane & ((1 << sh) - 1)) == 0;
}

// One TMA bulk copy per tile and input stream, described once per call in shared memory so that the
// lanes of warp 0 can issue all of a tile's copies in parallel (a single issuing thread made warp 0
// the straggler of every tile).
struct TmaDesc {
  const unsigned char *src;  // stream base (tile 0)
  uint32_t tile_bytes;       // bytes per tile = advance per tile
  uint32_t dst_off;          // byte offset inside the slot
  uint32_t group;            // 0: psi block (two tiles ahead), 1: the rest (one tile ahead)
  uint32_t pad_;
};
constexpr int kMaxDesc = 12;

namespace cg = cooperative_groups;

// CL: one gene-chain per thread-block cluster (every collective below then spans the cluster: element loops run
// over this rank's ranges, reductions are combined through distributed shared memory in rank order, so every CTA of
// the cluster holds bitwise identical scalars and executes the same control flow)
template <int KT, bool CL = false>
struct DevOps {
  const DevModel &m;
  // cluster geometry of this CTA (CL only; otherwise rank 0 of 1 with the whole tile range)
  int crank = 0, csize = 1, ta = 0, tb = 0;
  int nseg = 1, seg_lo[3] = {0, 0, 0}, seg_hi[3] = {0, 0, 0};   // element ranges of a D-vector this CTA owns
  double *cl_part = nullptr;   // [16] this CTA's partial sums of finish()            (shared memory, read by rank 0)
  double *cl_bcast = nullptr;  // [4]  rank 0's results of finish(): V, K, finite, pp  (read by the other ranks)
  double *cl_red = nullptr;    // [2][8] partial sums of the all-to-all reductions, double buffered
  double *cl_sc = nullptr;     // [12] the constrained scalars of the evaluation in flight (finish_small, warp 0)
  double *halo = nullptr;      // [2][look * kTile] drifted psi of the tiles below / above this rank's range
  int cl_par = 0;
  GeneDev gn;
  RngKey key;
  int jacobian;
  // shared memory: B[nbeta] | adjB[nbeta] | red[kWarps*kRedMax] | mbar[kStages] | ring[4 T] | phr[2 T] | stages
  // offsets (in doubles) into the one dynamic shared-memory array: going through smem_d + offset (instead of generic
  // pointers kept in this object, which lives in local memory) lets ptxas emit LDS / STS / ATOMS with 32-bit addresses
  int B_o = 0, adjB_o = 0, red_o = 0, sraw_o = 0;
  __device__ __forceinline__ double *Bs() const { return smem_d + B_o; }
  __device__ __forceinline__ double *As() const { return smem_d + adjB_o; }
  __device__ __forceinline__ double *Rs() const { return smem_d + red_o; }
  __device__ __forceinline__ double *Ss() const { return smem_d + sraw_o; }
  uint64_t *mbar;        // [0..3] psi slot landed (TMA), [4,5] rest slot landed (TMA), [6,7] psi of a tile is in the
                         // ring (256 arrivals), [8,9] every thread is done with a tile's staged inputs (256 arrivals)
  double *ring;          // drifted psi of tiles k-1, k, k+1 (+1 spare): ring[j & (4 T - 1)]
  double *psi_st;        // 3 slots x {q,p,g,minv}[T] of the psi block (TMA destination); tile x in slot x mod 3
  unsigned char *rest_st;  // 2 slots x stage_bytes: {q,p,g,minv}[T] of noise_raw, lsf[T], counts[T], key[T], ell[w][T]
  uint32_t tile_seq;     // tiles consumed so far by this CTA, mod 4 (look+2): psi slot seq % (look+2), parity
                         // (seq / (look+2)) & 1; rest slot seq & 1, parity (seq >> 1) & 1
  TmaDesc *tdesc;        // kMaxDesc TMA stream descriptors (shared memory)
  // HBM vectors of this CTA slot / chain
  double *pool;         // kPoolVecs x Di
  double *zq[2], *zg[2];
  double *q_prop, *q_samp;
  double *q_cur, *minv, *wm, *wm2;
  bool sample_is_q0;
  // outputs of this chain (may be null)
  double *out_diag;   // [num_samples][7]
  double *out_small;  // [num_samples][nsmall]
  double *out_theta;  // [num_samples][D] Stan order
  double *out_summ;   // [6][N]: mean,M2 of log_lambda | mean,M2 of lambda | mean,M2 of psi
  long long cyc_eval = 0, cyc_merge = 0, cyc_copy = 0, cyc_other = 0, n_merge = 0, n_copy = 0, cyc_pro = 0, cyc_fin = 0;
#ifdef CSPLOTCH_PROF_WAITS
  // development instrumentation: cycles thread 0 spends in each wait of the tile loop (reported through the
  // profile slots of merge / copy / other / n_merge / n_copy by k_nuts_advance)
  long long w_psi = 0, w_rest = 0, w_ring = 0, w_rel = 0, w_bins = 0, w_issue = 0, w_pre = 0, w_post = 0, w_t = 0;
#define PWMARK(var) { const long long n__ = clock64(); var += n__ - w_t; w_t = n__; }
#define PWSTART() { w_t = clock64(); }
#define PW(var, ...) { const long long t__ = clock64(); __VA_ARGS__; var += clock64() - t__; }
#else
#define PW(var, ...) { __VA_ARGS__; }
#define PWMARK(var) {}
#define PWSTART() {}
#endif

  __device__ __forceinline__ DevOps(const DevModel &mm) : m(mm) {
    tb = mm.NT;
    seg_hi[0] = mm.Di;
  }
  // bind this CTA to its rank of the cluster (CL kernels, after the model descriptor is staged)
  __device__ __forceinline__ void bind_cluster() {
    if (!CL) return;
    cg::cluster_group cluster = cg::this_cluster();
    crank = (int)cluster.block_rank();
    csize = (int)cluster.num_blocks();
    ta = m.cl_t0[crank];
    tb = m.cl_t0[crank + 1];
    nseg = 0;
    if (m.car) { seg_lo[nseg] = m.o_psi + ta * kTile; seg_hi[nseg] = m.o_psi + tb * kTile; ++nseg; }
    seg_lo[nseg] = m.o_noise + ta * kTile; seg_hi[nseg] = m.o_noise + tb * kTile; ++nseg;
    if (crank == 0) { seg_lo[nseg] = m.o_small; seg_hi[nseg] = m.Di; ++nseg; }
  }
  __device__ __forceinline__ void cluster_sync() {
    if (CL) cg::this_cluster().sync();
  }
  template <class T>
  __device__ __forceinline__ T *remote(T *p, int rank) const {
    return cg::this_cluster().map_shared_rank(p, (unsigned)rank);
  }
  // sum of NV values over the whole gene-chain (CTA, or every CTA of the cluster in rank order): identical in
  // every thread of every CTA
  template <int NV>
  __device__ __forceinline__ void chain_sum(double (&v)[NV]) {
    block_sum<NV>(v, Rs());
    if (!CL) return;
    static_assert(NV <= 8, "cl_red holds 8 values per buffer");
    double *buf = cl_red + 8 * cl_par;
    cl_par ^= 1;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) buf[i] = v[i];
    }
    cluster_sync();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 0.0;
    for (int r = 0; r < csize; ++r) {
      const double *rb = remote(buf, r);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] += rb[i];
    }
  }

  __device__ __forceinline__ double *vec(int id) const { return pool + (size_t)id * m.Di; }

  // ---- constrained scalars + beta levels into shared memory --------------------------------
  // Two steps: the unconstrained small block is first STAGED in shared memory (beta_raw in Bs()[], the scalars in Ss()[]:
  // by stage_small() from a vector in HBM, or directly by the leapfrog prologue, which has the drifted values in
  // registers), then finish_small() turns it in place into the constrained scalars and the beta levels.  No element of
  // the small block is fetched from HBM more than once per evaluation, and every fetch is one of a batch of
  // independent loads (a dependent global load per loop trip cost a full L2 round trip each: a third of a c2 evaluation).
  static constexpr int kUB = 4;   // elements per thread and batch
  __device__ __forceinline__ void stage_small(const double *q) {
    const double *sb = q + m.o_small;
    for (int i0 = threadIdx.x; i0 < m.nsmall; i0 += kUB * kBlock) {
      double v[kUB];
#pragma unroll
      for (int u = 0; u < kUB; ++u) {
        const int i = i0 + u * kBlock;
        v[u] = i < m.nsmall ? sb[i] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kUB; ++u) {
        const int i = i0 + u * kBlock;
        if (i < m.nbeta) Bs()[i] = v[u];
        else if (i < m.nsmall) Ss()[i - m.nbeta] = v[u];
      }
    }
  }
  __device__ __forceinline__ void finish_small(Scalars &sc) {
    // The constrained scalars: exp / logit of <= 7 numbers, ~700 fp64 instructions when one thread does them all.
    // Seven lanes of warp 0 take one scalar each and publish value, logs and log-Jacobian term through shared memory
    // (cl_sc) while the other warps start on the beta levels; everybody picks the results up after the first barrier.
    const int last = m.nscal - 1;
    if (threadIdx.x < 8) {
      const int t = threadIdx.x;
      double *o = cl_sc;   // [0..5] tau a sigma s2 s3 theta | [6] jac | [7] log_tau | [8] log1m_theta | [9] log_theta | [10] phi | [11] log_phi
      double val = 0.0, jt = 0.0, e1 = 0.0, e2 = 0.0;
      if (t == 0 && m.car) { const double u = Ss()[max(0, min(m.s_tau, last))]; val = exp(u); jt = u; e1 = u; }
      if (t == 1 && m.car) { const double u = Ss()[max(0, min(m.s_a, last))]; double la, l1a; logit_triple(u, val, la, l1a); jt = la + l1a; }
      if (t == 2) { const double u = Ss()[m.s_sigma]; val = exp(u); jt = u; }
      if (t == 3 && m.L2) { const double u = Ss()[max(0, min(m.s_s2, last))]; val = exp(u); jt = u; }
      if (t == 4 && m.L3) { const double u = Ss()[max(0, min(m.s_s3, last))]; val = exp(u); jt = u; }
      if (t == 5 && m.zi) { const double u = Ss()[max(0, min(m.s_theta, last))]; logit_triple(u, val, e2, e1); jt = e2 + e1; }
      if (t == 6 && m.nb) { const double u = Ss()[max(0, min(m.s_phi, last))]; val = 1e-6 + exp(u); e1 = log(val); jt = u; }
      // log-Jacobian in declaration order: tau, a, sigma, sigma_level_2, sigma_level_3, theta, phi
      double jac = 0.0;
#pragma unroll
      for (int k = 0; k < 7; ++k) jac += __shfl_sync(0xffu, jt, k);
      if (t < 6) o[t] = val;
      if (t == 0) { o[6] = jacobian ? jac : 0.0; o[7] = e1; }
      if (t == 5) { o[8] = e1; o[9] = e2; }
      if (t == 6) { o[10] = val; o[11] = e1; }
    }
    // beta levels in place: Bs()[idx] holds beta_raw on entry
    const int CK = m.C * m.K;
    for (int idx = threadIdx.x; idx < m.L1 * CK; idx += kBlock) {
      const double r = Bs()[idx];
      Bs()[idx] = (m.family == 1) ? r * gn.ps[idx % m.K] + gn.pm[idx % m.K] : 2.0 * r;
    }
    __syncthreads();
    sc.tau = cl_sc[0]; sc.a = cl_sc[1]; sc.sigma = cl_sc[2]; sc.s2 = cl_sc[3]; sc.s3 = cl_sc[4]; sc.theta = cl_sc[5];
    sc.jac = cl_sc[6]; sc.log_tau = cl_sc[7]; sc.log1m_theta = cl_sc[8]; sc.log_theta = cl_sc[9]; sc.phi = cl_sc[10];
    sc.log_phi = cl_sc[11];
    for (int idx = threadIdx.x; idx < m.L2 * CK; idx += kBlock) {
      const int i = idx / CK, rem = idx - i * CK;
      Bs()[(m.L1 + i) * CK + rem] = Bs()[m.map2[i] * CK + rem] + sc.s2 * Bs()[(m.L1 + i) * CK + rem];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < m.L3 * CK; idx += kBlock) {
      const int i = idx / CK, rem = idx - i * CK;
      Bs()[(m.L1 + m.L2 + i) * CK + rem] =
          Bs()[(m.L1 + m.map3[i]) * CK + rem] + sc.s3 * Bs()[(m.L1 + m.L2 + i) * CK + rem];
    }
    __syncthreads();
  }
  __device__ __forceinline__ void load_small(const double *q, Scalars &sc) {
    stage_small(q);
    __syncthreads();
    finish_small(sc);
  }

  // warp-collective flush of the per-lane bottom-level bins into adjB (shared)
  __device__ __forceinline__ void flush_bins(double (&acc)[KT], int cur_key) {
    const int lane = threadIdx.x & 31;
    unsigned remaining = __ballot_sync(0xffffffffu, cur_key >= 0);
    while (remaining) {
      const int leader = __ffs(remaining) - 1;
      const int k0 = __shfl_sync(0xffffffffu, cur_key, leader);
      const bool mine = (cur_key == k0);
      const unsigned same = __ballot_sync(0xffffffffu, mine);
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        if (k < m.K) {
          const double s = warp_sum(mine ? acc[k] : 0.0);
          if (lane == leader) atomicAdd(&As()[(m.bot0 * m.C) * m.K + k0 * m.K + k], s);
        }
      }
      remaining &= ~same;
    }
#pragma unroll
    for (int k = 0; k < KT; ++k) acc[k] = 0.0;
  }

  // ---- everything after the spot loop: CAR log-determinant, beta prior, block reductions, reverse of
  // the hierarchy, gradients (and second half-kick) of the small block, scalar chain rule --------------
  template <bool LEAP>
  __device__ void finish(const Scalars &sc, double a_lp, double a_ng, double a_qD, double a_qW, double a_dth,
                         double a_kin, double a_g2, double tails, double *dq, double *dg, double *dp, double eps,
                         double *V_out, double *K_out, bool *finite_out, const double *sp = nullptr,
                         double *rho_out = nullptr, double a_pp = 0.0, double *pp_out = nullptr,
                         bool eig_done = false, double a_ldet = 0.0, double a_dldet = 0.0, double a_dphi = 0.0) {
    const int tid = threadIdx.x;
    // momenta / metric entries of the (<= 6) scalars: one batch of independent loads issued now, so that
    // their latency overlaps the reductions below instead of stalling the scalar epilogue
    const int so = m.o_small + m.nbeta;
    const int sidx[7] = {m.s_tau, m.s_a, m.s_sigma, m.s_s2, m.s_s3, m.s_theta, m.s_phi};
    double ph_s[7], mi_s[7], po_s[7];
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      const int e = so + max(0, sidx[s]);
      const bool w0 = tid < 32;   // only warp 0 runs the scalar epilogue
      ph_s[s] = (LEAP && w0) ? dp[e] : 0.0;
      mi_s[s] = (LEAP && w0) ? minv[e] : 0.0;
      po_s[s] = (LEAP && w0 && rho_out) ? sp[e] : 0.0;
    }
    // log-determinant term of the CAR prior over the distinct eigenvalues (the tiled pass folds it into
    // its tile loop, where it is independent work that fills stall slots)
    if (m.car && !eig_done) {
      for (int i = tid; i < m.n_eigg; i += kBlock)
        eig_group_terms(m.eigg + (size_t)i * kEigGroup, __ldg(m.eiggm + i), sc.a, a_ldet, a_dldet);
    }
    __syncthreads();  // the shared atomics on adjB of every warp are done

    // reverse of the hierarchy (A.3): parents collect their children
    const int CK = m.C * m.K;
    if (m.L3) {
      for (int idx = tid; idx < m.L2 * CK; idx += kBlock) {
        const int i2 = idx / CK, rem = idx - i2 * CK;
        double s = As()[(m.L1 + i2) * CK + rem];
        for (int i3 = 0; i3 < m.L3; ++i3)
          if (m.map3[i3] == i2) s += As()[(m.L1 + m.L2 + i3) * CK + rem];
        As()[(m.L1 + i2) * CK + rem] = s;
      }
      __syncthreads();
    }
    if (m.L2) {
      for (int idx = tid; idx < m.L1 * CK; idx += kBlock) {
        const int i1 = idx / CK, rem = idx - i1 * CK;
        double s = As()[i1 * CK + rem];
        for (int i2 = 0; i2 < m.L2; ++i2)
          if (m.map2[i2] == i1) s += As()[(m.L1 + i2) * CK + rem];
        As()[i1 * CK + rem] = s;
      }
      __syncthreads();
    }
    // beta_raw: prior normal(0,1), gradient, second half-kick (and the fused level-0 merge) — batches of independent loads
    const double *sb = dq + m.o_small;
    double a_ds2 = 0, a_ds3 = 0;
    for (int i0 = tid; i0 < m.nbeta; i0 += kUB * kBlock) {
      double rv[kUB], pv[kUB], wv[kUB], ov[kUB];
#pragma unroll
      for (int u = 0; u < kUB; ++u) {
        const int idx = i0 + u * kBlock;
        const bool in = idx < m.nbeta;
        rv[u] = in ? sb[idx] : 0.0;
        pv[u] = (LEAP && in) ? dp[m.o_small + idx] : 0.0;
        wv[u] = (LEAP && in) ? minv[m.o_small + idx] : 0.0;
        ov[u] = (LEAP && in && rho_out) ? sp[m.o_small + idx] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kUB; ++u) {
        const int idx = i0 + u * kBlock;
        if (idx < m.nbeta) {
          const int i = idx / CK;
          const double ab = As()[idx], r = rv[u];
          a_lp -= 0.5 * r * r;
          double scale;
          if (i < m.L1) scale = (m.family == 1) ? gn.ps[idx % m.K] : 2.0;
          else if (i < m.L1 + m.L2) { scale = sc.s2; a_ds2 += ab * r; }
          else { scale = sc.s3; a_ds3 += ab * r; }
          const double gr = scale * ab - r;
          dg[m.o_small + idx] = -gr;
          a_g2 += gr * gr;
          if (LEAP) {
            const double pn = pv[u] + 0.5 * eps * gr;
            dp[m.o_small + idx] = pn;
            a_kin += wv[u] * pn * pn;
            if (rho_out) {
              a_pp += wv[u] * ov[u] * pn;
              rho_out[m.o_small + idx] = ov[u] + pn;
            }
          }
        }
      }
    }
    // ONE block reduction for everything the spot loop and this epilogue accumulated, finished by warp 0 alone:
    // the scalar epilogue below is a few hundred fp64 instructions (divisions included) with CTA-uniform operands —
    // executed by all eight warps it was a sixth of a c2 evaluation's instructions.  Warp 0 computes, the four results
    // go through shared memory.
    constexpr int NR = 13;
    double r1[NR] = {a_lp, a_ng, a_qD, a_qW, a_dth, a_kin, a_ldet, a_dldet, a_g2, a_dphi, a_ds2, a_ds3, a_pp};
    {
      const int lane = tid & 31, warp = tid >> 5;
      double *red = Rs();
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const double sw = warp_sum(r1[i]);
        if (lane == 0) red[warp * kRedMax + i] = sw;
      }
      __syncthreads();
      if (warp == 0) {
        double tot = 0.0;
        if (lane < NR) {
#pragma unroll
          for (int w = 0; w < kWarps; ++w) tot += red[w * kRedMax + lane];
        }
#pragma unroll
        for (int i = 0; i < NR; ++i) r1[i] = __shfl_sync(0xffffffffu, tot, i);
      }
    }
    if (tid < 32) {
      const double r2[5] = {r1[10], r1[11], 0.0, 0.0, r1[12]};
      double lp = r1[0] + sc.jac;
      double kin = r1[5];
      double g2 = r1[8];
      double gu_tau = 0, gu_a = 0, gu_sigma, gu_s2 = 0, gu_s3 = 0, gu_theta = 0, gu_phi = 0;
      if (m.car) {
        const double Q = r1[2] - sc.a * r1[3];
        const double inv_tau = 1.0 / sc.tau;
        lp += -2.0 * sc.log_tau - inv_tau;
        lp += 0.5 * ((double)m.N * sc.log_tau + r1[6] - sc.tau * Q);
        const double d_tau = -2.0 * inv_tau + inv_tau * inv_tau + 0.5 * ((double)m.N * inv_tau - Q);
        const double d_a = 0.5 * (r1[7] + sc.tau * r1[3]);
        gu_tau = sc.tau * d_tau + (jacobian ? 1.0 : 0.0);
        gu_a = sc.a * (1.0 - sc.a) * d_a + (jacobian ? (1.0 - 2.0 * sc.a) : 0.0);
      }
      lp += -0.5 * sc.sigma * sc.sigma;
      gu_sigma = sc.sigma * (-sc.sigma + 0.3 * r1[1]) + (jacobian ? 1.0 : 0.0);
      if (m.L2) { lp += -0.5 * sc.s2 * sc.s2; gu_s2 = sc.s2 * (-sc.s2 + r2[0]) + (jacobian ? 1.0 : 0.0); }
      if (m.L3) { lp += -0.5 * sc.s3 * sc.s3; gu_s3 = sc.s3 * (-sc.s3 + r2[1]) + (jacobian ? 1.0 : 0.0); }
      if (m.zi) {
        lp += tails;                      // theta ~ beta(1,2)
        lp += gn.nnz * tails - (m.nb ? 0.0 : gn.lgam);  // bernoulli(0|theta) and -lgamma(y+1) of the non-zero spots
        const double d_theta = r1[4] - (gn.nnz + 1.0) / (1.0 - sc.theta);
        gu_theta = sc.theta * (1.0 - sc.theta) * d_theta + (jacobian ? (1.0 - 2.0 * sc.theta) : 0.0);
      }
      if (m.nb) {
        lp += sc.log_phi - sc.phi;  // phi ~ gamma(2,1)
        gu_phi = (sc.phi - 1e-6) * (r1[9] + 1.0 / sc.phi - 1.0) + (jacobian ? 1.0 : 0.0);
      }
      double pn_s[7] = {0, 0, 0, 0, 0, 0, 0};
      double pp = r2[4];
      const double gu[7] = {gu_tau, gu_a, gu_sigma, gu_s2, gu_s3, gu_theta, gu_phi};
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        if (sidx[s] >= 0) {
          g2 += gu[s] * gu[s];
          if (LEAP) {
            pn_s[s] = ph_s[s] + 0.5 * eps * gu[s];
            kin += mi_s[s] * pn_s[s] * pn_s[s];
            if (rho_out) pp += mi_s[s] * po_s[s] * pn_s[s];
          }
        }
      }
      if (tid == 0) {
#pragma unroll
        for (int s = 0; s < 7; ++s) {
          if (sidx[s] >= 0) {
            dg[so + sidx[s]] = -gu[s];
            if (LEAP) dp[so + sidx[s]] = pn_s[s];
            if (LEAP && rho_out) rho_out[so + sidx[s]] = po_s[s] + pn_s[s];
          }
        }
        for (int s = m.nsmall; s < m.nsmall_pad; ++s) {
          dg[m.o_small + s] = 0.0;
          if (LEAP) dp[m.o_small + s] = 0.0;
          if (LEAP && rho_out) rho_out[m.o_small + s] = 0.0;
        }
        cl_bcast[0] = -lp;
        cl_bcast[1] = 0.5 * kin;
        cl_bcast[2] = (isfinite(lp) && isfinite(g2)) ? 1.0 : 0.0;
        cl_bcast[3] = pp;
      }
    }
    __syncthreads();
    if (pp_out) *pp_out = cl_bcast[3];
    if (V_out) *V_out = cl_bcast[0];
    if (K_out) *K_out = cl_bcast[1];
    if (finite_out) *finite_out = cl_bcast[2] != 0.0;
    __syncthreads();   // cl_bcast may be rewritten by the next evaluation only after everyone has read it
  }


  // ---- tiled pipeline (the fast path) -----------------------------------------------------------------
  // (__noinline__: the sampler's control state must not compete for registers with this loop; inlined into
  //  the persistent kernel ptxas spilled ~1 kB per thread inside the spot loop)
  // Tile k = spots [k T, (k+1) T), thread t <-> spot k T + t.  One elected thread streams every per-spot
  // input of tile k+2 into shared memory with 1-D TMA bulk copies while the CTA (i) drifts psi of tile k+1
  // into a 4-tile ring and (ii) evaluates likelihood + CAR + gradient + second half-kick of tile k reading
  // neighbours from the ring.  HBM traffic per leapfrog = the algorithmic 56 D + 4 N bytes (+ gene-shared
  // tables from L2); one __syncthreads per tile.
  // Everything the tile loop touches, hoisted into registers once per call: the object itself lives in
  // local memory (the functions are __noinline__), so member / model-field accesses inside the loop would
  // be dependent generic loads on the critical path.
  struct TileCtx {
    int NT, ta, ell_w, car, zi, o_psi, o_noise, stage_bytes, n_desc;   // NT: tiles of THIS CTA, the first is tile ta
    uint32_t seq0, psi_tx, rest_tx;
    const double *E;
    int ring_o, psi_o, rest_o, Bbot_o, bins_o;  // offsets (in doubles) into smem_d
    uint64_t *mbar;
    const TmaDesc *desc;
  };
  template <bool LEAP>
  __device__ __forceinline__ TileCtx make_ctx(const double *sq, const double *sg, const double *sp) {
    TileCtx c;
    c.NT = tb - ta; c.ta = ta; c.ell_w = m.ell_w; c.car = m.car; c.zi = m.zi;
    c.o_psi = m.o_psi; c.o_noise = m.o_noise; c.stage_bytes = m.stage_bytes;
    const int bin0 = (m.bot0 * m.C) * KT;
    c.seq0 = tile_seq;
    c.E = m.E;
    c.ring_o = (int)(ring - smem_d); c.psi_o = (int)(psi_st - smem_d);
    c.rest_o = (int)(reinterpret_cast<double *>(rest_st) - smem_d);
    c.Bbot_o = B_o + bin0; c.bins_o = adjB_o + bin0;
    c.mbar = mbar;
    c.desc = tdesc;
    constexpr uint32_t vb = kTile * sizeof(double), ib = kTile * sizeof(int);
    const int n_psi = m.car ? (LEAP ? 4 : 1) : 0;
    const int n_rest = (LEAP ? 4 : 1) + 3 + (m.car ? 1 : 0);
    c.n_desc = n_psi + n_rest;
    c.psi_tx = (uint32_t)n_psi * vb;
    c.rest_tx = (LEAP ? 4 * vb : vb) + vb + 2 * ib + (m.car ? (uint32_t)m.ell_w * (ib / 2) : 0u);
    // descriptor table (thread i writes descriptor i)
    const int t = threadIdx.x;
    if (t < c.n_desc) {
      TmaDesc d;
      d.pad_ = 0;
      if (t < n_psi) {
        const double *base = t == 0 ? sq : (t == 1 ? sp : (t == 2 ? sg : minv));
        d.src = reinterpret_cast<const unsigned char *>(base + m.o_psi);
        d.tile_bytes = vb; d.dst_off = (uint32_t)t * vb; d.group = 0;
      } else {
        const int r = t - n_psi;
        const int nv = LEAP ? 4 : 1;
        d.group = 1;
        if (r < nv) {
          const double *base = r == 0 ? sq : (r == 1 ? sp : (r == 2 ? sg : minv));
          d.src = reinterpret_cast<const unsigned char *>(base + m.o_noise);
          d.tile_bytes = vb; d.dst_off = (uint32_t)r * vb;
        } else if (r == nv) {
          d.src = reinterpret_cast<const unsigned char *>(m.lsf); d.tile_bytes = vb; d.dst_off = 4 * vb;
        } else if (r == nv + 1) {
          d.src = reinterpret_cast<const unsigned char *>(gn.counts); d.tile_bytes = ib; d.dst_off = 5 * vb;
        } else if (r == nv + 2) {
          d.src = reinterpret_cast<const unsigned char *>(m.key); d.tile_bytes = ib; d.dst_off = 5 * vb + ib;
        } else {
          d.src = reinterpret_cast<const unsigned char *>((CL && csize > 1) ? m.ell_cl : m.ell); d.tile_bytes = (uint32_t)m.ell_w * (ib / 2);
          d.dst_off = 5 * vb + 2 * ib;
        }
      }
      tdesc[t] = d;
    }
    return c;
  }
  // psi inputs run look+1 tiles ahead (the drift of tile k+look precedes the gradient of tile k, which
  // still reads p, g, minv of its own tile) in look+2 slots; the rest one tile ahead in two slots.  Called
  // by every lane of warp 0 at trip k: issues psi(k+look+1) and rest(k+1).
  // copy number d of the table: psi streams go to tile k+look+1, the rest to tile k+1
  template <int LOOK>
  static __device__ __forceinline__ void issue_desc(const TileCtx &c, int k, int di) {
    constexpr uint32_t NPS = LOOK + 2;
    const TmaDesc d = c.desc[di];
    const int tile = d.group == 0 ? k + LOOK + 1 : k + 1;
    if (tile >= 0 && tile < c.NT) {
      const uint32_t seq = c.seq0 + (uint32_t)tile;
      const uint32_t slot = d.group == 0 ? seq % NPS : (seq & 1u);
      unsigned char *dst = d.group == 0
                               ? reinterpret_cast<unsigned char *>(smem_d + c.psi_o) + slot * (4 * kTile * sizeof(double))
                               : reinterpret_cast<unsigned char *>(smem_d + c.rest_o) + (size_t)slot * c.stage_bytes;
      bulk_g2s(dst + d.dst_off, d.src + (size_t)(c.ta + tile) * d.tile_bytes, d.tile_bytes, c.mbar + 4 * d.group + slot);
    }
  }
  template <int LOOK>
  static __device__ __forceinline__ void expect_tiles(const TileCtx &c, int k, bool psi, bool rest) {
    constexpr uint32_t NPS = LOOK + 2;
    const int kp = k + LOOK + 1;  // psi tile issued at trip k
    if (psi && c.car && kp >= 0 && kp < c.NT) mbar_expect_tx(c.mbar + ((c.seq0 + (uint32_t)kp) % NPS), c.psi_tx);
    if (rest && k + 1 >= 0 && k + 1 < c.NT) mbar_expect_tx(c.mbar + 4 + ((c.seq0 + (uint32_t)(k + 1)) & 1u), c.rest_tx);
  }
  template <int LOOK>
  static __device__ __forceinline__ void issue_tiles(const TileCtx &c, int k) {
    const int lane = threadIdx.x;  // warp 0: lane == tid
    expect_tiles<LOOK>(c, k, lane == 0, lane == 1);
    __syncwarp();
    if (lane < c.n_desc) issue_desc<LOOK>(c, k, lane);
  }
  // Measured alternatives (B200, c2 / c3 sampler throughput against this scheme), kept out:
  //  * the copies spread over the eight warps (lane 0 of warp w issues descriptors w, w + 8), issued after the ring
  //    barrier, which already implies the release of the slots: +1 % / -4.5 % (they land too late for the next trip);
  //  * the same, issued at the top of the trip after every warp waits for the release: -7 % / -26 % (a second
  //    CTA-wide synchronisation per tile).
  // drift psi of tile k into the ring
  template <bool LEAP, int LOOK>
  static __device__ __forceinline__ void drift_tile(const TileCtx &c, int k, double eps) {
    constexpr uint32_t NPS = LOOK + 2;
    constexpr int RMASK = (LOOK == 1 ? 4 : 8) * kTile - 1;
    const int t = threadIdx.x;
    const uint32_t seq = c.seq0 + (uint32_t)k;
    const uint32_t ps = seq % NPS;
    mbar_wait(c.mbar + ps, (seq / NPS) & 1u);
    const double *d = smem_d + c.psi_o + ps * 4 * kTile;
    double qn = d[t];
    if (LEAP) qn += eps * d[3 * kTile + t] * (d[kTile + t] - 0.5 * eps * d[2 * kTile + t]);
    smem_d[c.ring_o + (((c.ta + k) * kTile + t) & RMASK)] = qn;
    mbar_arrive_cta(c.mbar + 6 + (seq & 1u));  // "psi of tile k is in the ring"
  }

  // per-tile contribution of one warp to the bottom-level bins: G[bin][k] += sum_lanes gj * E[lane][k]
  template <int KP>
  // (fp64 atomicAdd on shared memory is a compare-and-swap spin loop, SASS ATOMS.CAST.SPIN; accumulating these bins
  //  with fire-and-forget RED.ADD.F64 in L2 instead measured 2-3 % SLOWER on c2 and c3: the fence + L2 round trip
  //  that finish() then needs costs more than the contention it removes)
  static __device__ __forceinline__ void add_bins(int bins_o, double (&v)[KP], int bin_l) {
    double *bins = smem_d + bins_o;
    const int lane = threadIdx.x & 31;
    const int b0 = __shfl_sync(0xffffffffu, bin_l, 0);
    if (__all_sync(0xffffffffu, bin_l == b0)) {
      if (b0 < 0) return;
      warp_transpose_sum<KP>(v);
      const int b = bin_of_lane<KP>(lane);
      if (lane_owns_bin<KP>(lane) && b < KT) atomicAdd(&bins[b0 * KT + b], v[0]);
      return;
    }
    if (KP <= 2) {
      // several bins in this warp (small ST tissues: ~6 spots per bin): segmented inclusive scan over the
      // runs of equal bin; the last lane of every run adds the run total.  Exact for any sequence of bins
      // (a bin that re-appears simply contributes two partial sums).
      const int prev = __shfl_up_sync(0xffffffffu, bin_l, 1);
      const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || bin_l != prev);
      const int seg_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
      const bool is_tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
#pragma unroll
      for (int i = 0; i < KP; ++i) {
        if (i < KT) {
          double sacc = v[i];
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, sacc, off);
            if (lane - off >= seg_start) sacc += t;
          }
          if (is_tail && bin_l >= 0) atomicAdd(&bins[bin_l * KT + i], sacc);
        }
      }
    } else {
      // compositional models: a bin boundary inside a warp is rare (one tile in ~13 at c3); one transposed
      // reduction per distinct bin keeps this path small enough not to cost the main loop registers
      unsigned remaining = __ballot_sync(0xffffffffu, bin_l >= 0);
      while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const int kk = __shfl_sync(0xffffffffu, bin_l, leader);
        const bool mine = bin_l == kk;
        double w[KP];
#pragma unroll
        for (int i = 0; i < KP; ++i) w[i] = mine ? v[i] : 0.0;
        warp_transpose_sum<KP>(w);
        const int b = bin_of_lane<KP>(lane);
        if (lane_owns_bin<KP>(lane) && b < KT) atomicAdd(&bins[kk * KT + b], w[0]);
        remaining &= ~__ballot_sync(0xffffffffu, mine);
      }
    }
  }

  template <bool LEAP, int LOOK>
  __device__ __noinline__ void eval_tiled(const double *__restrict__ sq, const double *__restrict__ sg,
                                          const double *__restrict__ sp, double *dq, double *dg, double *dp,
                                          double eps, double *V_out, double *K_out, bool *finite_out,
                                          double *rho_out, double *pp_out) {
    constexpr int KP = KT <= 1 ? 1 : (KT <= 2 ? 2 : (KT <= 4 ? 4 : (KT <= 8 ? 8 : 16)));
    constexpr uint32_t NPS = LOOK + 2;
    constexpr int RMASK = (LOOK == 1 ? 4 : 8) * kTile - 1;
    const int tid = threadIdx.x;
    const long long t_begin = clock64();
    const TileCtx c = make_ctx<LEAP>(sq, sg, sp);
    const int NT = c.NT;
    const bool write_q = LEAP || dq != sq;
    // generic-proxy writes of earlier passes (q, p, g of this chain) must be visible to the async proxy
    // (TMA) that is about to read them; the barrier also publishes the descriptor table
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __syncthreads();
    // the first tiles start streaming while the small block is prepared
    if (tid < 32)
      for (int k0 = -(LOOK + 1); k0 < 0; ++k0) issue_tiles<LOOK>(c, k0);  // psi(0..look), rest(0)
    const bool clustered = CL && csize > 1;
    if (clustered && c.car) {
      // halo: psi of the LOOK tiles below and above this rank's range, drifted here from the neighbouring ranks'
      // elements (L2 loads: another CTA wrote them) into the cells behind the ring that ell_cl points at.  Nothing of
      // a tile is overwritten before the cluster barrier below, so in-place leapfrogs are safe.
      for (int h = 0; h < 2; ++h) {
        const int t0h = h == 0 ? ta - LOOK : tb;
        for (int e = tid; e < LOOK * kTile; e += kBlock) {
          const int tile = t0h + e / kTile;
          double qn = 0.0;
          if (tile >= 0 && tile < m.NT && (tile < ta || tile >= tb)) {
            const int el = m.o_psi + tile * kTile + (e % kTile);
            qn = __ldcg(sq + el);
            if (LEAP) qn += eps * __ldcg(minv + el) * (__ldcg(sp + el) - 0.5 * eps * __ldcg(sg + el));
          }
          halo[h * LOOK * kTile + e] = qn;
        }
      }
    }
    // small block: half-kick + drift (or copy) in batches of independent loads; the new values are staged in shared
    // memory for finish_small() (no read-back from HBM) and written to HBM.  Every rank of a cluster stages the small
    // block for itself; rank 0 alone writes it, in a second pass after the cluster barrier (an in-place leapfrog must
    // not overwrite what another rank is still reading).
    auto small_pass = [&](bool to_smem, bool to_hbm) {
      for (int i0 = tid; i0 < m.nsmall_pad; i0 += kUB * kBlock) {
        double qv[kUB], pv[kUB], gv[kUB], wv[kUB];
#pragma unroll
        for (int u = 0; u < kUB; ++u) {
          const int e = m.o_small + i0 + u * kBlock;
          const bool in = i0 + u * kBlock < m.nsmall_pad;
          qv[u] = in ? sq[e] : 0.0;
          pv[u] = (LEAP && in) ? sp[e] : 0.0;
          gv[u] = (LEAP && in) ? sg[e] : 0.0;
          wv[u] = (LEAP && in) ? minv[e] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kUB; ++u) {
          const int i = i0 + u * kBlock, e = m.o_small + i;
          if (i < m.nsmall_pad) {
            double qn = qv[u];
            if (LEAP) {
              const double ph = pv[u] - 0.5 * eps * gv[u];
              qn += eps * wv[u] * ph;
              if (to_hbm) dp[e] = ph;
            }
            if (to_hbm && write_q) dq[e] = qn;
            if (to_smem) {
              if (i < m.nbeta) Bs()[i] = qn;
              else if (i < m.nsmall) Ss()[i - m.nbeta] = qn;
            }
          }
        }
      }
    };
    Scalars sc;
    small_pass(true, !clustered);
    __syncthreads();
    finish_small(sc);
    if (clustered) {
      cluster_sync();   // every halo and every rank's copy of the small block is read: in-place writes may begin
      if (crank == 0) small_pass(false, true);
    }
    for (int idx = tid; idx < m.nbeta; idx += kBlock) As()[idx] = 0.0;
    const double tails = sc.log1m_theta;
    const double sig03 = 0.3 * sc.sigma;
    const double theta = sc.theta, om_theta = 1.0 - sc.theta, tau = sc.tau, a_car = sc.a;
    const double half_eps = 0.5 * eps;
    double a_lp = 0, a_ng = 0, a_qD = 0, a_qW = 0, a_dth = 0, a_kin = 0, a_g2 = 0, a_pp = 0;
    double a_ldet = 0, a_dldet = 0;
    // zero-inflated spots contribute log(theta + (1-theta) e^{-mu}): summed as the log of a running product
    LogProd zi_prod;
    zi_prod.init();
    const bool merge0 = LEAP && rho_out != nullptr;
    // the non-compositional kernels have registers to spare and few tiles per evaluation: they fold the
    // eigenvalue term into the tile loop; the compositional ones (register-critical loop) leave it to finish()
    constexpr bool kEigInLoop = KT <= 2;
    // (a cluster splits the eigenvalue groups evenly over its ranks)
    const int eig_per = (m.n_eigg + csize - 1) / csize;
    const int eig0 = kEigInLoop ? crank * eig_per : 0;
    const int n_eigg = kEigInLoop ? min(m.n_eigg, eig0 + eig_per) : 0;
    const double *eigg = m.eigg, *eiggm = m.eiggm;

    if (c.car)
      for (int k0 = 0; k0 < LOOK && k0 < NT; ++k0) {
        drift_tile<LEAP, LOOK>(c, k0, eps);
        const uint32_t sq0 = c.seq0 + (uint32_t)k0;
        mbar_wait(c.mbar + 6 + (sq0 & 1u), (sq0 >> 1) & 1u);  // ring(0..look-1); later tiles are awaited `look` ahead
      }
    const long long t_loop = clock64();
    for (int k = 0; k < NT; ++k) {
      // cell-type fractions of this tile straight from L2 (gene-shared, coalesced per k); issued before
      // the barrier so that their latency overlaps the drift of the next tile and the barrier wait
      double e[KP];
      if (KT > 1) {
        const double *Ej = c.E + ((size_t)(c.ta + k) * KT) * kTile + tid;
#pragma unroll
        for (int i = 0; i < KP; ++i) e[i] = (i < KT) ? __ldg(Ej + (size_t)i * kTile) : 0.0;
      } else {
        e[0] = 1.0;
      }
      if (kEigInLoop && c.car) {
        // one group of eight distinct eigenvalues of the CAR precision per thread and tile (the groups fill the
        // first tiles densely: whole warps either have one or skip the block)
        const int ei = eig0 + k * kTile + tid;
        if (ei < n_eigg) eig_group_terms(eigg + (size_t)ei * kEigGroup, __ldg(eiggm + ei), a_car, a_ldet, a_dldet);
      }
      PW(w_psi, if (c.car && k + LOOK < NT) drift_tile<LEAP, LOOK>(c, k + LOOK, eps));
      const uint32_t seq = c.seq0 + (uint32_t)k;
      if (tid < 32) {
        // psi(k+look+1) and rest(k+1) go into the slots of tile k-1: wait until every thread has released them
        PW(w_rel, if (k > 0) mbar_wait(c.mbar + 8 + ((seq - 1u) & 1u), ((seq - 1u) >> 1) & 1u));
        PW(w_issue, issue_tiles<LOOK>(c, k));
      }
      PW(w_rest, mbar_wait(c.mbar + 4 + (seq & 1u), (seq >> 1) & 1u));
      PWSTART();
      const double *pd = smem_d + c.psi_o + (seq % NPS) * 4 * kTile;  // q, p, g, minv of psi, tile k
      const double *rd = smem_d + c.rest_o + (seq & 1u) * (c.stage_bytes >> 3);
      const int *ri = reinterpret_cast<const int *>(rd + 5 * kTile);
      const double *ringp = smem_d + c.ring_o;
      const int j = (c.ta + k) * kTile + tid;
      const int kraw = ri[kTile + tid];  // -1 in the padding
      const int bin = kraw < 0 ? -1 : (kraw & 0xFFFFFF);
      double gj = 0.0;
      if (kraw >= 0) {
        // noise_raw: drift in registers
        double eps_j = rd[tid], ph_e = 0.0, m_e = 0.0;
        if (LEAP) {
          m_e = rd[3 * kTile + tid];
          ph_e = rd[kTile + tid] - half_eps * rd[2 * kTile + tid];
          eps_j += eps * m_e * ph_e;
        }
        const double *bj = smem_d + c.Bbot_o + bin * KT;
        double ll;
        if (KT == 1) {
          ll = bj[0];
        } else if (KT % 2 == 0) {
          // two accumulation chains, the beta row in 16-byte loads (its offset is even whenever KT is)
          double l0 = 0.0, l1 = 0.0;
#pragma unroll
          for (int i = 0; i < KT; i += 2) {
            const double2 b2 = *reinterpret_cast<const double2 *>(bj + i);
            l0 += b2.x * e[i];
            l1 += b2.y * e[i + 1];
          }
          ll = l0 + l1;
        } else {
          ll = 0.0;
#pragma unroll
          for (int i = 0; i < KT; ++i) ll += bj[i] * e[i];
        }
        double psi_j = 0.0, Wpsi = 0.0, deg = 0.0;
        if (c.car) {
          deg = (double)(kraw >> 24);
          // ELL entries are ring positions (neighbour id & (4T-1)); missing neighbours point at the
          // always-zero cell ring[4T], so the gather is branch-free
          const unsigned short *el = reinterpret_cast<const unsigned short *>(ri + 2 * kTile) + tid;
          // (the widths that occur: 4 = ST / HD grids, 6 = Visium hex; anything else takes the 8-wide form)
          unsigned short pos[kMaxEll];
          const int ew = c.ell_w;
          if (ew == 6) {
#pragma unroll
            for (int w = 0; w < 6; ++w) pos[w] = el[w * kTile];
          } else if (ew == 4) {
#pragma unroll
            for (int w = 0; w < 4; ++w) pos[w] = el[w * kTile];
          } else {
#pragma unroll
            for (int w = 0; w < kMaxEll; ++w) pos[w] = (w < ew) ? el[w * kTile] : (unsigned short)(RMASK + 1);
          }
          // split barrier: every thread announced its part of ring(k+1) right after drifting it; only now,
          // with the independent work above already issued, do we need all of them
          PWMARK(w_pre);
          if (k + LOOK < NT) {
            const uint32_t sl = seq + (uint32_t)LOOK;
            PW(w_ring, mbar_wait(c.mbar + 6 + (sl & 1u), (sl >> 1) & 1u));
          }
          PWSTART();
          psi_j = ringp[j & RMASK];
          if (ew == 6) {
            Wpsi = ((ringp[pos[0]] + ringp[pos[1]]) + (ringp[pos[2]] + ringp[pos[3]])) + (ringp[pos[4]] + ringp[pos[5]]);
          } else if (ew == 4) {
            Wpsi = (ringp[pos[0]] + ringp[pos[1]]) + (ringp[pos[2]] + ringp[pos[3]]);
          } else {
#pragma unroll
            for (int w = 0; w < kMaxEll; ++w) Wpsi += ringp[pos[w]];
          }
          ll += psi_j;
        }
        ll += sig03 * eps_j;
        const double eta = ll + rd[4 * kTile + tid];
        const double mu = exp_fast(eta);
        const int y = ri[tid];
        if (c.zi && y == 0) {
          // log(theta + (1-theta) e^{-mu}) = log_sum_exp(log theta, log1m theta - mu); the two softmax
          // weights are theta / s and z / s
          const double em = exp_fast(-mu);
          const double z = om_theta * em;
          const double s = theta + z;
          double inv_s;
          if (s > 1e-290) {  // always, short of theta underflowing
            inv_s = rcp_fast(s);
            zi_prod.mul(s);
          } else {
            inv_s = 1.0 / s;
            a_lp += log(s);
          }
          gj = -(mu * z) * inv_s;
          a_dth += (1.0 - em) * inv_s;
        } else {
          a_lp += (double)y * eta - mu;
          gj = (double)y - mu;
        }
        a_lp -= 0.5 * eps_j * eps_j;
        a_ng += eps_j * gj;
        const double gr_eps = sig03 * gj - eps_j;
        a_g2 += gr_eps * gr_eps;
        dg[c.o_noise + j] = -gr_eps;
        if (write_q) dq[c.o_noise + j] = eps_j;
        if (LEAP) {
          const double pn = ph_e + half_eps * gr_eps;
          dp[c.o_noise + j] = pn;
          a_kin += m_e * pn * pn;
          if (merge0) {
            const double po = rd[kTile + tid];
            a_pp += m_e * po * pn;
            rho_out[c.o_noise + j] = po + pn;
          }
        }
        if (c.car) {
          a_qD += (psi_j * deg) * psi_j;
          a_qW += Wpsi * psi_j;
          const double gr_psi = gj - tau * (deg * psi_j - a_car * Wpsi);
          a_g2 += gr_psi * gr_psi;
          dg[c.o_psi + j] = -gr_psi;
          if (write_q) dq[c.o_psi + j] = psi_j;
          if (LEAP) {
            const double po = pd[kTile + tid], mi = pd[3 * kTile + tid];
            const double pn = (po - half_eps * pd[2 * kTile + tid]) + half_eps * gr_psi;
            dp[c.o_psi + j] = pn;
            a_kin += mi * pn * pn;
            if (merge0) {
              a_pp += mi * po * pn;
              rho_out[c.o_psi + j] = po + pn;
            }
          }
        }
      } else {
        // padding spots carry q = p = g = 0
        dg[c.o_noise + j] = 0.0;
        if (write_q) dq[c.o_noise + j] = 0.0;
        if (LEAP) dp[c.o_noise + j] = 0.0;
        if (merge0) rho_out[c.o_noise + j] = 0.0;
        if (c.car) {
          dg[c.o_psi + j] = 0.0;
          if (write_q) dq[c.o_psi + j] = 0.0;
          if (LEAP) dp[c.o_psi + j] = 0.0;
          if (merge0) rho_out[c.o_psi + j] = 0.0;
        }
      }
      PWMARK(w_post);
      mbar_arrive_cta(c.mbar + 8 + (seq & 1u));  // this warp no longer reads the staged inputs of tile k
#pragma unroll
      for (int i = 0; i < KP; ++i) e[i] *= gj;
      PW(w_bins, add_bins<KP>(c.bins_o, e, bin));
    }
    tile_seq = (c.seq0 + (uint32_t)NT) % (4u * NPS);
    const long long t_fin = clock64();
    if (c.zi) a_lp += zi_prod.log_value();
    // the zi-only part of d/dtheta was accumulated as (1 - e^{-mu}) / s; finish() expects it in a_dth
    if (!clustered) {
      finish<LEAP>(sc, a_lp, a_ng, a_qD, a_qW, a_dth, a_kin, a_g2, tails, dq, dg, dp, eps, V_out, K_out, finite_out, sp, rho_out,
                   a_pp, pp_out, kEigInLoop, a_ldet, a_dldet);
    } else {
      // cluster: every rank publishes its partial sums (and holds its partial beta-gradient bins in adjB); rank 0
      // adds them up in rank order, runs the epilogue of the small block alone and broadcasts the four results
      double r[kRedMax] = {a_lp, a_ng, a_qD, a_qW, a_dth, a_kin, a_g2, a_pp, a_ldet, a_dldet, 0.0, 0.0};
      block_sum<kRedMax>(r, Rs());
      if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kRedMax; ++i) cl_part[i] = r[i];
      }
      cluster_sync();
      if (crank == 0) {
        double t[kRedMax];
#pragma unroll
        for (int i = 0; i < kRedMax; ++i) t[i] = 0.0;
        for (int rk = 0; rk < csize; ++rk) {
          const double *rp = remote(cl_part, rk);
#pragma unroll
          for (int i = 0; i < kRedMax; ++i) t[i] += rp[i];
        }
        const int b0i = (m.bot0 * m.C) * KT, b1i = m.nbeta;
        for (int rk = 1; rk < csize; ++rk) {
          const double *ra = remote(As(), rk);
          for (int idx = b0i + tid; idx < b1i; idx += kBlock) As()[idx] += ra[idx];
        }
        __syncthreads();
        const bool lead = tid == 0;   // the totals enter finish()'s own block reduction once
        double Vv = 0.0, Kv = 0.0, ppv = 0.0;
        bool fin = true;
        finish<LEAP>(sc, lead ? t[0] : 0.0, lead ? t[1] : 0.0, lead ? t[2] : 0.0, lead ? t[3] : 0.0, lead ? t[4] : 0.0,
                     lead ? t[5] : 0.0, lead ? t[6] : 0.0, tails, dq, dg, dp, eps, &Vv, &Kv, &fin, sp, rho_out,
                     lead ? t[7] : 0.0, &ppv, kEigInLoop, lead ? t[8] : 0.0, lead ? t[9] : 0.0);
        // (finish() has left V, K, finite, pp in cl_bcast for the other ranks)
        (void)Vv; (void)Kv; (void)fin; (void)ppv;
      }
      cluster_sync();
      const double *rb = remote(cl_bcast, 0);
      if (V_out) *V_out = rb[0];
      if (K_out) *K_out = rb[1];
      if (finite_out) *finite_out = rb[2] != 0.0;
      if (pp_out) *pp_out = rb[3];
    }
    cyc_pro += t_loop - t_begin;
    cyc_fin += clock64() - t_fin;
  }

  // ---- fused [half-kick, drift,] log-density + gradient [, half-kick] ------------------------
  // LEAP: (sq,sg,sp) --eps--> (dq,dg,dp).  !LEAP: dq := sq (if different), dg := -grad lp(sq).
  // g vectors hold the gradient of the potential V = -lp, like Stan's ps_point::g.
  template <bool LEAP>
  // rho_out != nullptr (LEAP only): also the level-0 U-turn merge of the previous leaf with this one, i.e.
  // rho_out = sp + p_new and *pp_out = sum minv * sp * p_new (nuts_core.h), fused into the same pass
  __device__ void eval(const double *sq, const double *sg, const double *sp, double *dq, double *dg,
                       double *dp, double eps, double *V_out, double *K_out, bool *finite_out,
                       double *rho_out = nullptr, double *pp_out = nullptr) {
    // neighbours up to two tiles away: 257- to 512-wide rows (Visium-HD bins), compositional kernels included
    if (m.tiled && m.look == 2) eval_tiled<LEAP, 2>(sq, sg, sp, dq, dg, dp, eps, V_out, K_out, finite_out, rho_out, pp_out);
    else if (m.tiled) eval_tiled<LEAP, 1>(sq, sg, sp, dq, dg, dp, eps, V_out, K_out, finite_out, rho_out, pp_out);
    else eval_generic<LEAP>(sq, sg, sp, dq, dg, dp, eps, V_out, K_out, finite_out, rho_out, pp_out);
  }

  // Fallback for layouts the tiled pipeline does not take (neighbours further than one tile away or
  // degree > kMaxEll): two phases through HBM/L2 with a CSR gather.
  template <bool LEAP>
  __device__ __noinline__ void eval_generic(const double *sq, const double *sg, const double *sp, double *dq, double *dg,
                               double *dp, double eps, double *V_out, double *K_out, bool *finite_out,
                               double *rho_out, double *pp_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Di = m.Di;
    // phase A: half-kick + drift for every element (the spot loop below gathers psi of
    // neighbours, so all of psi must have moved before any gradient is taken)
    if (LEAP) {
      for (int e = 2 * tid; e < Di; e += 2 * kBlock) {
        const double2 q = *reinterpret_cast<const double2 *>(sq + e);
        const double2 g = *reinterpret_cast<const double2 *>(sg + e);
        const double2 p = *reinterpret_cast<const double2 *>(sp + e);
        const double2 mi = *reinterpret_cast<const double2 *>(minv + e);
        double2 ph, qn;
        ph.x = p.x - 0.5 * eps * g.x;
        ph.y = p.y - 0.5 * eps * g.y;
        qn.x = q.x + eps * mi.x * ph.x;
        qn.y = q.y + eps * mi.y * ph.y;
        *reinterpret_cast<double2 *>(dq + e) = qn;
        *reinterpret_cast<double2 *>(dp + e) = ph;
      }
    } else if (dq != sq) {
      for (int e = 2 * tid; e < Di; e += 2 * kBlock)
        *reinterpret_cast<double2 *>(dq + e) = *reinterpret_cast<const double2 *>(sq + e);
    }
    __syncthreads();

    Scalars sc;
    load_small(dq, sc);
    for (int idx = tid; idx < m.nbeta; idx += kBlock) As()[idx] = 0.0;
    __syncthreads();

    const double heads = sc.log_theta;
    const double tails = sc.log1m_theta;
    const double sig03 = 0.3 * sc.sigma;
    const double nb_mult = (m.nb == 1 && m.zi) ? (double)m.N : 1.0;
    const double lgam_phi = m.nb ? lgamma(sc.phi) : 0.0, dig_phi = m.nb ? digamma_d(sc.phi) : 0.0;
    double a_dphi = 0.0;
    const double *psi = dq + m.o_psi;
    const double *noi = dq + m.o_noise;
    double a_lp = 0, a_ng = 0, a_qD = 0, a_qW = 0, a_dth = 0, a_kin = 0, a_g2 = 0, a_pp = 0;
    double acc[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) acc[k] = 0.0;
    int cur_key = -1;

    // each warp owns a contiguous chunk of spots, 32 consecutive spots per trip
    const int chunk = (((m.N + kWarps - 1) / kWarps) + 31) & ~31;
    const int j_end = min(m.N, (warp + 1) * chunk);
    for (int base = warp * chunk; base < j_end; base += 32) {
      const int j = base + lane;
      const bool act = j < j_end;
      int kj = -1;
      double gj = 0.0;
      double e[KT];
      if (act) {
        kj = m.key[j] & 0xFFFFFF;
        const double eps_j = noi[j];
        double ll;
        const double *bj = Bs() + (m.bot0 * m.C) * m.K + kj * m.K;
        if (KT == 1) {
          ll = bj[0];
          e[0] = 1.0;
        } else {
          ll = 0.0;
          const double *Ej = E_ptr(m, j);
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            e[k] = (k < m.K) ? Ej[(size_t)k * kTile] : 0.0;
            if (k < m.K) ll += bj[k] * e[k];
          }
        }
        double psi_j = 0.0, Wpsi = 0.0, deg = 0.0;
        if (m.car) {
          psi_j = psi[j];
          const int r0 = m.rowptr[j], r1 = m.rowptr[j + 1];
          for (int r = r0; r < r1; ++r) Wpsi += psi[m.col[r]];
          deg = (double)(r1 - r0);
          ll += psi_j;
        }
        ll += sig03 * eps_j;
        const double eta = ll + m.lsf[j];
        const double mu = exp(eta);
        const int y = gn.counts[j];
        if (m.nb) {
          // neg_binomial_2_log (comp_splotch_stan_model.stan:248-265); mult = N reproduces the phi_arr
          // broadcast of the ZINB branch as the reference writes it (nb == 1), 1 the per-spot term (nb == 2)
          const double yd = (double)y;
          const double dl = eta - sc.log_phi;
          const double L = dl > 0 ? dl + log1p(exp(-dl)) : log1p(exp(dl));
          const double sg = inv_logit_d(dl);
          const double nbv = lgamma(yd + sc.phi) - lgamma(yd + 1.0) - lgam_phi + yd * eta - sc.phi * L - yd * (sc.log_phi + L);
          const double dnb_deta = yd - (sc.phi + yd) * sg;
          const double dnb_dphi = digamma_d(yd + sc.phi) - dig_phi - L + (sc.phi + yd) * sg / sc.phi - yd / sc.phi;
          if (m.zi && y == 0) {
            const double Bq = tails + nb_mult * nbv;
            const double dd = heads - Bq;
            const double t = exp(-fabs(dd));
            const double l1p = log1p(t);
            const double wmax = 1.0 / (1.0 + t), wmin = t / (1.0 + t);
            const double wA = dd >= 0 ? wmax : wmin;
            const double wB = dd >= 0 ? wmin : wmax;
            a_lp += (dd >= 0 ? heads : Bq) + l1p;
            gj = wB * nb_mult * dnb_deta;
            a_dphi += wB * nb_mult * dnb_dphi;
            a_dth += wA / sc.theta - wB / (1.0 - sc.theta);
          } else {
            a_lp += nb_mult * nbv;   // (zi: log1m theta of the non-zero spots is added once in finish())
            gj = nb_mult * dnb_deta;
            a_dphi += nb_mult * dnb_dphi;
          }
        } else if (m.zi && y == 0) {
          // log_sum_exp(log theta, log1m theta - mu) and its softmax weights
          const double Bq = tails - mu;
          const double dd = heads - Bq;
          const double t = exp(-fabs(dd));
          const double l1p = log1p(t);
          const double wmax = 1.0 / (1.0 + t), wmin = t / (1.0 + t);
          const double wA = dd >= 0 ? wmax : wmin;
          const double wB = dd >= 0 ? wmin : wmax;
          a_lp += (dd >= 0 ? heads : Bq) + l1p;
          gj = -wB * mu;
          a_dth += wA / sc.theta - wB / (1.0 - sc.theta);
        } else {
          a_lp += (double)y * eta - mu;
          gj = (double)y - mu;
        }
        a_lp -= 0.5 * eps_j * eps_j;
        a_ng += eps_j * gj;
        const double gr_eps = sig03 * gj - eps_j;
        dg[m.o_noise + j] = -gr_eps;
        a_g2 += gr_eps * gr_eps;
        if (LEAP) {
          const double pn = dp[m.o_noise + j] + 0.5 * eps * gr_eps;
          dp[m.o_noise + j] = pn;
          a_kin += minv[m.o_noise + j] * pn * pn;
          if (rho_out) {
            const double po = sp[m.o_noise + j];
            a_pp += minv[m.o_noise + j] * po * pn;
            rho_out[m.o_noise + j] = po + pn;
          }
        }
        if (m.car) {
          a_qD += (psi_j * deg) * psi_j;
          a_qW += Wpsi * psi_j;
          const double gr_psi = gj - sc.tau * (deg * psi_j - sc.a * Wpsi);
          dg[m.o_psi + j] = -gr_psi;
          a_g2 += gr_psi * gr_psi;
          if (LEAP) {
            const double pn = dp[m.o_psi + j] + 0.5 * eps * gr_psi;
            dp[m.o_psi + j] = pn;
            a_kin += minv[m.o_psi + j] * pn * pn;
            if (rho_out) {
              const double po = sp[m.o_psi + j];
              a_pp += minv[m.o_psi + j] * po * pn;
              rho_out[m.o_psi + j] = po + pn;
            }
          }
        }
      }
      // bottom-level bins: per-lane run-length accumulation, warp-collective flush on change
      const bool changed = act && (kj != cur_key) && (cur_key >= 0);
      if (__any_sync(0xffffffffu, changed)) flush_bins(acc, cur_key);
      if (act) {
        cur_key = kj;
#pragma unroll
        for (int k = 0; k < KT; ++k) acc[k] += gj * e[k];
      }
    }
    flush_bins(acc, cur_key);

    // padding elements (Npad > N) carry q = p = g = 0
    for (int j = m.N + tid; j < m.Npad; j += kBlock) {
      dg[m.o_noise + j] = 0.0;
      if (LEAP) dp[m.o_noise + j] = 0.0;
      if (LEAP && rho_out) rho_out[m.o_noise + j] = 0.0;
      if (m.car) {
        dg[m.o_psi + j] = 0.0;
        if (LEAP) dp[m.o_psi + j] = 0.0;
        if (LEAP && rho_out) rho_out[m.o_psi + j] = 0.0;
      }
    }

    finish<LEAP>(sc, a_lp, a_ng, a_qD, a_qW, a_dth, a_kin, a_g2, tails, dq, dg, dp, eps, V_out, K_out, finite_out, sp, rho_out,
                 a_pp, pp_out, false, 0.0, 0.0, a_dphi);
  }

  // ---- Ops interface of nuts_core.h ---------------------------------------------------------------
  __device__ __noinline__ double sample_momentum(int pv, uint32_t purpose, uint32_t sub, uint32_t iter) {
    const long long t0 = clock64();
    double *p = vec(pv);
    double k = 0.0;
    for (int sgi = 0; sgi < nseg; ++sgi)
      for (int e = seg_lo[sgi] + threadIdx.x; e < seg_hi[sgi]; e += kBlock) {
        const int si = m.imap[e];
        double v = 0.0;
        if (si >= 0) {
          const double mi = minv[e];
          v = rng_normal(key, purpose, sub, iter, (uint32_t)si) / sqrt(mi);
          k += mi * v * v;
        }
        p[e] = v;
      }
    double r[1] = {k};
    chain_sum<1>(r);
    sample_is_q0 = true;
    cyc_other += clock64() - t0;
    return 0.5 * r[0];
  }
  __device__ double eval_at_current(int z, bool *finite) {
    double V;
    const long long t0 = clock64();
    eval<false>(q_cur, nullptr, nullptr, zq[z], zg[z], nullptr, 0.0, &V, nullptr, finite);
    cyc_eval += clock64() - t0;
    return V;
  }
  // rho0 >= 0: also rho0 := p(ps) + p(pd) and *pp = <p(ps), p(pd)>_M (level-0 U-turn merge fused in)
  __device__ void leapfrog(int zs, int ps, int zd, int pd, double eps, double *V, double *K, int rho0, double *pp) {
    const long long t0 = clock64();
    eval<true>(zq[zs], zg[zs], vec(ps), zq[zd], zg[zd], vec(pd), eps, V, K, nullptr, rho0 >= 0 ? vec(rho0) : nullptr, pp);
    cyc_eval += clock64() - t0;
  }
  // Stan's three generalised U-turn criteria for the subtree L ++ C (base_nuts.hpp
  // build_tree / transition) as six metric-weighted inner products in one pass.
  __device__ bool merge(SubTree L, SubTree C, int rho_out) {
    const long long t0 = clock64();
    const bool r = merge_pass(L, C, rho_out);
    cyc_merge += clock64() - t0;
    ++n_merge;
    return r;
  }
  __device__ __noinline__ bool merge_pass(SubTree L, SubTree C, int rho_out) {
    const double *Lpb = vec(L.pb), *Lpe = vec(L.pe), *Lrho = vec(L.rho);
    const double *Cpb = vec(C.pb), *Cpe = vec(C.pe), *Crho = vec(C.rho);
    double *out = vec(rho_out);
    double a1 = 0, a2 = 0, b1 = 0, b2 = 0, c1 = 0, c2 = 0;
    const double *mv = minv;
    // streaming pass over 7 vectors: three strips of 16 B per thread in flight (21 independent loads): a merging CTA
    // runs next to CTAs doing arithmetic, so its own bytes in flight set its speed (two strips: 22 GB/s per CTA at c3)
    auto body = [&](const double2 w, const double2 lpb, const double2 lpe, const double2 lrh, const double2 cpb,
                    const double2 cpe, const double2 crh, double *o) {
      double2 rs;
      rs.x = lrh.x + crh.x;
      rs.y = lrh.y + crh.y;
      a1 += (w.x * lpb.x) * rs.x + (w.y * lpb.y) * rs.y;
      a2 += (w.x * cpe.x) * rs.x + (w.y * cpe.y) * rs.y;
      const double rbx = lrh.x + cpb.x, rby = lrh.y + cpb.y;
      b1 += (w.x * lpb.x) * rbx + (w.y * lpb.y) * rby;
      b2 += (w.x * cpb.x) * rbx + (w.y * cpb.y) * rby;
      const double rcx = crh.x + lpe.x, rcy = crh.y + lpe.y;
      c1 += (w.x * lpe.x) * rcx + (w.y * lpe.y) * rcy;
      c2 += (w.x * cpe.x) * rcx + (w.y * cpe.y) * rcy;
      *reinterpret_cast<double2 *>(o) = rs;
    };
#define CSP_LD2(p, e) (*reinterpret_cast<const double2 *>((p) + (e)))
    for (int sgi = 0; sgi < nseg; ++sgi) {
      const int hi = seg_hi[sgi];
      int e = seg_lo[sgi] + 2 * threadIdx.x;
      for (; e + 4 * kBlock < hi; e += 6 * kBlock) {   // three strips: 21 independent 16-byte loads per thread
        const int f = e + 2 * kBlock, h = e + 4 * kBlock;
        const double2 w0 = CSP_LD2(mv, e), w1 = CSP_LD2(mv, f), w2 = CSP_LD2(mv, h);
        const double2 p0 = CSP_LD2(Lpb, e), p1 = CSP_LD2(Lpb, f), p2 = CSP_LD2(Lpb, h);
        const double2 q0 = CSP_LD2(Lpe, e), q1 = CSP_LD2(Lpe, f), q2 = CSP_LD2(Lpe, h);
        const double2 r0 = CSP_LD2(Lrho, e), r1 = CSP_LD2(Lrho, f), r2 = CSP_LD2(Lrho, h);
        const double2 s0 = CSP_LD2(Cpb, e), s1 = CSP_LD2(Cpb, f), s2 = CSP_LD2(Cpb, h);
        const double2 t0 = CSP_LD2(Cpe, e), t1 = CSP_LD2(Cpe, f), t2 = CSP_LD2(Cpe, h);
        const double2 u0 = CSP_LD2(Crho, e), u1 = CSP_LD2(Crho, f), u2 = CSP_LD2(Crho, h);
        body(w0, p0, q0, r0, s0, t0, u0, out + e);
        body(w1, p1, q1, r1, s1, t1, u1, out + f);
        body(w2, p2, q2, r2, s2, t2, u2, out + h);
      }
      for (; e < hi; e += 2 * kBlock)
        body(CSP_LD2(mv, e), CSP_LD2(Lpb, e), CSP_LD2(Lpe, e), CSP_LD2(Lrho, e), CSP_LD2(Cpb, e), CSP_LD2(Cpe, e),
             CSP_LD2(Crho, e), out + e);
    }
    double r[6] = {a1, a2, b1, b2, c1, c2};
    chain_sum<6>(r);
    return (r[1] > 0 && r[0] > 0) && (r[3] > 0 && r[2] > 0) && (r[5] > 0 && r[4] > 0);
  }
  // Two merges of a cascade in one pass (nuts_core.h: leaf index with >= 2 trailing ones above level 0): L1 ++ C, then
  // L2 ++ C' with C' = (L1.pb, C.pe, rho(L1) + rho(C)).  Reads Minv and the nine vectors once and writes only the outer
  // rho: 11 vector streams where two merge() calls move 16 (C.pe, L1.pb and the inner rho would be re-read, the inner
  // rho written).  CSPLOTCH_NO_MERGE2 (compile time) restores the one-by-one cascade for A/B runs.
#ifndef CSPLOTCH_NO_MERGE2
  static constexpr bool kHasMerge2 = true;
#else
  static constexpr bool kHasMerge2 = false;
#endif
  __device__ unsigned merge2(SubTree L1, SubTree L2, SubTree C, int rho_out) {
    const long long t0 = clock64();
    const unsigned r = merge2_pass(L1, L2, C, rho_out);
    cyc_merge += clock64() - t0;
    n_merge += 2;
    return r;
  }
  __device__ __noinline__ unsigned merge2_pass(SubTree L1, SubTree L2, SubTree C, int rho_out) {
    const double *Apb = vec(L1.pb), *Ape = vec(L1.pe), *Arho = vec(L1.rho);
    const double *Bpb = vec(L2.pb), *Bpe = vec(L2.pe), *Brho = vec(L2.rho);
    const double *Cpb = vec(C.pb), *Cpe = vec(C.pe), *Crho = vec(C.rho);
    double *out = vec(rho_out);
    double a1 = 0, a2 = 0, b1 = 0, b2 = 0, c1 = 0, c2 = 0;   // L1 ++ C
    double d1 = 0, d2 = 0, e1 = 0, e2 = 0, f1 = 0, f2 = 0;   // L2 ++ (L1 ++ C)
    const double *mv = minv;
    auto body = [&](const double2 w, const double2 apb, const double2 ape, const double2 arh, const double2 bpb,
                    const double2 bpe, const double2 brh, const double2 cpb, const double2 cpe, const double2 crh,
                    double *o) {
      // inner merge: left = L1, right = C (the same arithmetic as merge_pass)
      double2 rs;
      rs.x = arh.x + crh.x;
      rs.y = arh.y + crh.y;
      a1 += (w.x * apb.x) * rs.x + (w.y * apb.y) * rs.y;
      a2 += (w.x * cpe.x) * rs.x + (w.y * cpe.y) * rs.y;
      const double rbx = arh.x + cpb.x, rby = arh.y + cpb.y;
      b1 += (w.x * apb.x) * rbx + (w.y * apb.y) * rby;
      b2 += (w.x * cpb.x) * rbx + (w.y * cpb.y) * rby;
      const double rcx = crh.x + ape.x, rcy = crh.y + ape.y;
      c1 += (w.x * ape.x) * rcx + (w.y * ape.y) * rcy;
      c2 += (w.x * cpe.x) * rcx + (w.y * cpe.y) * rcy;
      // outer merge: left = L2, right = (pb = L1.pb, pe = C.pe, rho = rs)
      double2 ro;
      ro.x = brh.x + rs.x;
      ro.y = brh.y + rs.y;
      d1 += (w.x * bpb.x) * ro.x + (w.y * bpb.y) * ro.y;
      d2 += (w.x * cpe.x) * ro.x + (w.y * cpe.y) * ro.y;
      const double sbx = brh.x + apb.x, sby = brh.y + apb.y;
      e1 += (w.x * bpb.x) * sbx + (w.y * bpb.y) * sby;
      e2 += (w.x * apb.x) * sbx + (w.y * apb.y) * sby;
      const double scx = rs.x + bpe.x, scy = rs.y + bpe.y;
      f1 += (w.x * bpe.x) * scx + (w.y * bpe.y) * scy;
      f2 += (w.x * cpe.x) * scx + (w.y * cpe.y) * scy;
      *reinterpret_cast<double2 *>(o) = ro;
    };
    for (int sgi = 0; sgi < nseg; ++sgi) {
      const int hi = seg_hi[sgi];
      int e = seg_lo[sgi] + 2 * threadIdx.x;
      for (; e + 2 * kBlock < hi; e += 4 * kBlock) {   // two strips: 20 independent 16-byte loads per thread
        const int f = e + 2 * kBlock;
        const double2 w0 = CSP_LD2(mv, e), w1 = CSP_LD2(mv, f);
        const double2 g0 = CSP_LD2(Apb, e), g1 = CSP_LD2(Apb, f);
        const double2 h0 = CSP_LD2(Ape, e), h1 = CSP_LD2(Ape, f);
        const double2 i0 = CSP_LD2(Arho, e), i1 = CSP_LD2(Arho, f);
        const double2 j0 = CSP_LD2(Bpb, e), j1 = CSP_LD2(Bpb, f);
        const double2 k0 = CSP_LD2(Bpe, e), k1 = CSP_LD2(Bpe, f);
        const double2 l0 = CSP_LD2(Brho, e), l1 = CSP_LD2(Brho, f);
        const double2 s0 = CSP_LD2(Cpb, e), s1 = CSP_LD2(Cpb, f);
        const double2 t0 = CSP_LD2(Cpe, e), t1 = CSP_LD2(Cpe, f);
        const double2 u0 = CSP_LD2(Crho, e), u1 = CSP_LD2(Crho, f);
        body(w0, g0, h0, i0, j0, k0, l0, s0, t0, u0, out + e);
        body(w1, g1, h1, i1, j1, k1, l1, s1, t1, u1, out + f);
      }
      for (; e < hi; e += 2 * kBlock)
        body(CSP_LD2(mv, e), CSP_LD2(Apb, e), CSP_LD2(Ape, e), CSP_LD2(Arho, e), CSP_LD2(Bpb, e), CSP_LD2(Bpe, e),
             CSP_LD2(Brho, e), CSP_LD2(Cpb, e), CSP_LD2(Cpe, e), CSP_LD2(Crho, e), out + e);
    }
    double r[6] = {a1, a2, b1, b2, c1, c2};
    chain_sum<6>(r);
    double q[6] = {d1, d2, e1, e2, f1, f2};
    chain_sum<6>(q);
    const bool p1 = (r[1] > 0 && r[0] > 0) && (r[3] > 0 && r[2] > 0) && (r[5] > 0 && r[4] > 0);
    const bool p2 = (q[1] > 0 && q[0] > 0) && (q[3] > 0 && q[2] > 0) && (q[5] > 0 && q[4] > 0);
    return (p1 ? 1u : 0u) | (p2 ? 2u : 0u);
  }
  __device__ __noinline__ void copy_vec(double *dst, const double *src) {
    for (int sgi = 0; sgi < nseg; ++sgi) {
      const int hi = seg_hi[sgi];
      int e = seg_lo[sgi] + 2 * threadIdx.x;
      for (; e + 6 * kBlock < hi; e += 8 * kBlock) {  // four 16-byte loads in flight per thread
        const double2 a = CSP_LD2(src, e), b = CSP_LD2(src, e + 2 * kBlock), c = CSP_LD2(src, e + 4 * kBlock),
                      d = CSP_LD2(src, e + 6 * kBlock);
        *reinterpret_cast<double2 *>(dst + e) = a;
        *reinterpret_cast<double2 *>(dst + e + 2 * kBlock) = b;
        *reinterpret_cast<double2 *>(dst + e + 4 * kBlock) = c;
        *reinterpret_cast<double2 *>(dst + e + 6 * kBlock) = d;
      }
      for (; e < hi; e += 2 * kBlock) *reinterpret_cast<double2 *>(dst + e) = CSP_LD2(src, e);
    }
    __syncthreads();
  }
  __device__ void copy_q_to_prop(int z) {
    const long long t0 = clock64();
    copy_vec(q_prop, zq[z]);
    cyc_copy += clock64() - t0;
    ++n_copy;
  }
  __device__ void swap_prop_sample() {
    double *t = q_prop;
    q_prop = q_samp;
    q_samp = t;
    sample_is_q0 = false;
  }
  __device__ void commit_sample() {
    const long long t0 = clock64();
    if (!sample_is_q0) copy_vec(q_cur, q_samp);
    cyc_copy += clock64() - t0;
  }
  __device__ __noinline__ void init_uniform(uint32_t attempt, double radius) {
    for (int sgi = 0; sgi < nseg; ++sgi)
      for (int e = seg_lo[sgi] + threadIdx.x; e < seg_hi[sgi]; e += kBlock) {
        const int si = m.imap[e];
        q_cur[e] = si >= 0 ? rng_init(key, attempt, (uint32_t)si, radius) : 0.0;
      }
    __syncthreads();
  }
  __device__ __noinline__ void set_unit_metric() {
    for (int sgi = 0; sgi < nseg; ++sgi)
      for (int e = seg_lo[sgi] + threadIdx.x; e < seg_hi[sgi]; e += kBlock) { minv[e] = 1.0; wm[e] = 0.0; wm2[e] = 0.0; }
    __syncthreads();
  }
  __device__ __noinline__ void welford_add(double n) {
    for (int sgi = 0; sgi < nseg; ++sgi)
      for (int e = seg_lo[sgi] + threadIdx.x; e < seg_hi[sgi]; e += kBlock) {
        const double q = q_cur[e];
        const double delta = q - wm[e];
        const double mean = wm[e] + delta / n;
        wm[e] = mean;
        wm2[e] += (q - mean) * delta;
      }
    __syncthreads();
  }
  __device__ __noinline__ void metric_update(double n) {
    for (int sgi = 0; sgi < nseg; ++sgi)
      for (int e = seg_lo[sgi] + threadIdx.x; e < seg_hi[sgi]; e += kBlock) {
        double v = minv[e];
        if (n > 1) v = wm2[e] / (n - 1.0);
        minv[e] = (m.imap[e] >= 0) ? (n / (n + 5.0)) * v + 1e-3 * (5.0 / (n + 5.0)) : 1.0;
        wm[e] = 0.0;
        wm2[e] = 0.0;
      }
    __syncthreads();
  }
  // one sampling draw: sampler columns, the small block, optional full theta, and the streaming
  // summaries summarize_stan_hdf5 needs (/root/reference/splotch/utils.py:234-257)
  __device__ __noinline__ void record_draw(const DrawInfo &info, int s) {
    const int tid = threadIdx.x;
    const bool clustered = CL && csize > 1;
    // cluster: the ranks have just committed their parts of q_cur; rank 0 reads all of it (through L2) for the spot
    // summaries, every rank writes its own elements of the draw
    if (clustered) cluster_sync();
    const bool lead = !clustered || crank == 0;
    if (out_diag && tid == 0 && lead) {
      double *r = out_diag + (size_t)s * 7;
      r[0] = info.lp; r[1] = info.accept_stat; r[2] = info.stepsize; r[3] = info.treedepth;
      r[4] = info.n_leapfrog; r[5] = info.divergent; r[6] = info.energy;
    }
    if (out_small && lead)
      for (int i = tid; i < m.nsmall; i += kBlock) {
        int o;
        if (i < m.nbeta) {
          const int row = i / m.K, k = i - row * m.K;
          o = k < m.Ktrue ? row * m.Ktrue + k : -1;
        } else {
          o = m.nbeta_true + (i - m.nbeta);
        }
        if (o >= 0) out_small[(size_t)s * m.nsmall_true + o] = q_cur[m.o_small + i];
      }
    if (out_theta)
      for (int sgi = 0; sgi < nseg; ++sgi)
        for (int e = seg_lo[sgi] + tid; e < seg_hi[sgi]; e += kBlock) {
          const int si = m.imap[e];
          if (si >= 0) out_theta[(size_t)s * m.D + si] = q_cur[e];
        }
    if (out_summ && lead) {
      Scalars sc;
      load_small(q_cur, sc);
      const double n = (double)(s + 1);
      const double sig03 = 0.3 * sc.sigma;
      double *mll = out_summ, *vll = out_summ + m.N, *mla = out_summ + 2 * (size_t)m.N,
             *vla = out_summ + 3 * (size_t)m.N, *mps = out_summ + 4 * (size_t)m.N,
             *vps = out_summ + 5 * (size_t)m.N;
      for (int j = tid; j < m.N; j += kBlock) {
        const double *bj = Bs() + (m.bot0 * m.C) * m.K + (m.key[j] & 0xFFFFFF) * m.K;
        double ll;
        if (KT == 1) {
          ll = bj[0];
        } else {
          ll = 0.0;
          const double *Ej = E_ptr(m, j);
          for (int k = 0; k < m.K; ++k) ll += bj[k] * Ej[(size_t)k * kTile];
        }
        double psi_j = 0.0;
        if (m.car) { psi_j = clustered ? __ldcg(q_cur + m.o_psi + j) : q_cur[m.o_psi + j]; ll += psi_j; }
        ll += sig03 * (clustered ? __ldcg(q_cur + m.o_noise + j) : q_cur[m.o_noise + j]);
        double d = ll - mll[j], mean = mll[j] + d / n;
        mll[j] = mean; vll[j] += (ll - mean) * d;
        const double la = exp(ll);
        d = la - mla[j]; mean = mla[j] + d / n;
        mla[j] = mean; vla[j] += (la - mean) * d;
        if (m.car) {
          d = psi_j - mps[j]; mean = mps[j] + d / n;
          mps[j] = mean; vps[j] += (psi_j - mean) * d;
        }
      }
    }
    __syncthreads();
  }
};

}  // namespace csp
