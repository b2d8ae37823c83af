// TEST INFRASTRUCTURE: runs the PRODUCT's sampler control logic (csplotch_b200/csrc/nuts_core.h,
// the code every thread of the CUDA kernel executes) on the host with trivially simple vector
// ops, so that the iterative tree / pool / adaptation logic can be compared draw for draw with
// the oracle's recursive restatement without a GPU.  Never loaded by the product.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "nuts_core.h"

#ifndef EMUL_MERGE2
#define EMUL_MERGE2 1   // the product default: two merges of a cascade per pass (0: one by one)
#endif

typedef double (*density_fn)(void *ctx, const double *q, double *grad);

struct HostOps {
  int D;
  density_fn fn;
  void *ctx;
  csp::RngKey key;
  std::vector<double> pool;          // kPoolVecs x D
  std::vector<double> zq[2], zg[2];  // state buffers
  std::vector<double> q_cur, q_prop, q_samp, minv, wm, wm2, tmp;
  bool sample_is_q0 = true;
  double *draws = nullptr;  // [num_samples][7 + D]
  long merges = 0, copies = 0;

  double *vec(int id) { return pool.data() + (size_t)id * D; }

  double sample_momentum(int pv, uint32_t purpose, uint32_t sub, uint32_t iter) {
    double *p = vec(pv), k = 0;
    for (int i = 0; i < D; ++i) {
      p[i] = csp::rng_normal(key, purpose, sub, iter, (uint32_t)i) / std::sqrt(minv[i]);
      k += minv[i] * p[i] * p[i];
    }
    sample_is_q0 = true;
    return 0.5 * k;
  }
  double eval_at_current(int z, bool *finite) {
    zq[z] = q_cur;
    double lp = fn(ctx, q_cur.data(), tmp.data());
    bool ok = std::isfinite(lp);
    for (int i = 0; i < D; ++i) { zg[z][i] = -tmp[i]; ok = ok && std::isfinite(tmp[i]); }
    *finite = ok;
    return -lp;
  }
  void leapfrog(int zs, int ps, int zd, int pd, double eps, double *V, double *K, int rho0, double *pp) {
    const double *sq = zq[zs].data(), *sg = zg[zs].data(), *sp = vec(ps);
    double *dq = zq[zd].data(), *dg = zg[zd].data(), *dp = vec(pd);
    for (int i = 0; i < D; ++i) {
      const double ph = sp[i] - 0.5 * eps * sg[i];
      dp[i] = ph;
      dq[i] = sq[i] + eps * minv[i] * ph;
    }
    const double lp = fn(ctx, dq, tmp.data());
    double k = 0, pq = 0;
    double *rho = rho0 >= 0 ? vec(rho0) : nullptr;
    for (int i = 0; i < D; ++i) {
      dg[i] = -tmp[i];
      dp[i] -= 0.5 * eps * dg[i];
      k += minv[i] * dp[i] * dp[i];
      if (rho) {
        pq += minv[i] * sp[i] * dp[i];
        rho[i] = sp[i] + dp[i];
      }
    }
    *V = -lp;
    *K = 0.5 * k;
    *pp = pq;
  }
  bool merge(csp::SubTree L, csp::SubTree C, int rho_out) {
    const double *Lpb = vec(L.pb), *Lpe = vec(L.pe), *Lrho = vec(L.rho);
    const double *Cpb = vec(C.pb), *Cpe = vec(C.pe), *Crho = vec(C.rho);
    double *out = vec(rho_out);
    double a1 = 0, a2 = 0, b1 = 0, b2 = 0, c1 = 0, c2 = 0;
    for (int i = 0; i < D; ++i) {
      const double w = minv[i];
      const double rs = Lrho[i] + Crho[i];
      a1 += (w * Lpb[i]) * rs;
      a2 += (w * Cpe[i]) * rs;
      const double rb = Lrho[i] + Cpb[i];
      b1 += (w * Lpb[i]) * rb;
      b2 += (w * Cpb[i]) * rb;
      const double rc = Crho[i] + Lpe[i];
      c1 += (w * Lpe[i]) * rc;
      c2 += (w * Cpe[i]) * rc;
      out[i] = rs;
    }
    ++merges;
    return (a2 > 0 && a1 > 0) && (b2 > 0 && b1 > 0) && (c2 > 0 && c1 > 0);
  }
  // the two-merge cascade step of the device Ops (device_ops.cuh merge2): L1 ++ C, then L2 ++ (L1 ++ C) in one pass
  static constexpr bool kHasMerge2 = EMUL_MERGE2 != 0;
  unsigned merge2(csp::SubTree L1, csp::SubTree L2, csp::SubTree C, int rho_out) {
    const double *Apb = vec(L1.pb), *Ape = vec(L1.pe), *Arho = vec(L1.rho);
    const double *Bpb = vec(L2.pb), *Bpe = vec(L2.pe), *Brho = vec(L2.rho);
    const double *Cpb = vec(C.pb), *Cpe = vec(C.pe), *Crho = vec(C.rho);
    double *out = vec(rho_out);
    double a1 = 0, a2 = 0, b1 = 0, b2 = 0, c1 = 0, c2 = 0, d1 = 0, d2 = 0, e1 = 0, e2 = 0, f1 = 0, f2 = 0;
    for (int i = 0; i < D; ++i) {
      const double w = minv[i];
      const double rs = Arho[i] + Crho[i];
      a1 += (w * Apb[i]) * rs;
      a2 += (w * Cpe[i]) * rs;
      const double rb = Arho[i] + Cpb[i];
      b1 += (w * Apb[i]) * rb;
      b2 += (w * Cpb[i]) * rb;
      const double rc = Crho[i] + Ape[i];
      c1 += (w * Ape[i]) * rc;
      c2 += (w * Cpe[i]) * rc;
      const double ro = Brho[i] + rs;
      d1 += (w * Bpb[i]) * ro;
      d2 += (w * Cpe[i]) * ro;
      const double sb = Brho[i] + Apb[i];
      e1 += (w * Bpb[i]) * sb;
      e2 += (w * Apb[i]) * sb;
      const double sc = rs + Bpe[i];
      f1 += (w * Bpe[i]) * sc;
      f2 += (w * Cpe[i]) * sc;
      out[i] = ro;
    }
    merges += 2;
    const bool p1 = (a2 > 0 && a1 > 0) && (b2 > 0 && b1 > 0) && (c2 > 0 && c1 > 0);
    const bool p2 = (d2 > 0 && d1 > 0) && (e2 > 0 && e1 > 0) && (f2 > 0 && f1 > 0);
    return (p1 ? 1u : 0u) | (p2 ? 2u : 0u);
  }
  void copy_q_to_prop(int z) { q_prop = zq[z]; ++copies; }
  void swap_prop_sample() { q_prop.swap(q_samp); sample_is_q0 = false; }
  void commit_sample() { if (!sample_is_q0) q_cur = q_samp; }
  void init_uniform(uint32_t attempt, double radius) {
    for (int i = 0; i < D; ++i) q_cur[i] = csp::rng_init(key, attempt, (uint32_t)i, radius);
  }
  void set_unit_metric() { std::fill(minv.begin(), minv.end(), 1.0); std::fill(wm.begin(), wm.end(), 0.0); std::fill(wm2.begin(), wm2.end(), 0.0); }
  void welford_add(double n) {
    for (int i = 0; i < D; ++i) {
      const double delta = q_cur[i] - wm[i];
      wm[i] += delta / n;
      wm2[i] += (q_cur[i] - wm[i]) * delta;
    }
  }
  void metric_update(double n) {
    for (int i = 0; i < D; ++i) {
      double v = minv[i];
      if (n > 1) v = wm2[i] / (n - 1.0);
      minv[i] = (n / (n + 5.0)) * v + 1e-3 * (5.0 / (n + 5.0));
      wm[i] = 0; wm2[i] = 0;
    }
  }
  void record_draw(const csp::DrawInfo &info, int s) {
    if (!draws) return;
    double *r = draws + (size_t)s * (7 + D);
    r[0] = info.lp; r[1] = info.accept_stat; r[2] = info.stepsize; r[3] = info.treedepth;
    r[4] = info.n_leapfrog; r[5] = info.divergent; r[6] = info.energy;
    std::memcpy(r + 7, q_cur.data(), sizeof(double) * D);
  }
};

extern "C" int emul_nuts_run(density_fn fn, void *ctx, int D, const csp::NutsConfig *cfg, uint32_t gene,
                             uint32_t chain, const double *q_init, int chunk, double *draws,
                             double *inv_metric_out, double *eps_out, long long *n_leapfrog_out,
                             long long *n_grad_out) {
  HostOps ops;
  ops.D = D; ops.fn = fn; ops.ctx = ctx;
  ops.key = csp::RngKey{cfg->seed, chain, gene};
  ops.pool.assign((size_t)csp::kPoolVecs * D, 0.0);
  for (int z = 0; z < 2; ++z) { ops.zq[z].assign(D, 0.0); ops.zg[z].assign(D, 0.0); }
  ops.q_cur.assign(D, 0.0); ops.q_prop.assign(D, 0.0); ops.q_samp.assign(D, 0.0);
  ops.minv.assign(D, 1.0); ops.wm.assign(D, 0.0); ops.wm2.assign(D, 0.0); ops.tmp.assign(D, 0.0);
  ops.draws = draws;
  if (q_init) std::memcpy(ops.q_cur.data(), q_init, sizeof(double) * D);
  csp::ChainScalars cs;
  std::memset(&cs, 0, sizeof(cs));
  csp::Nuts<HostOps> nuts(ops, *cfg, cs, ops.key);
  const int total = cfg->num_warmup + cfg->num_samples;
  // advance in chunks exactly as the host drives the kernel
  do {
    nuts.advance(chunk > 0 ? chunk : total, q_init != nullptr);
  } while (cs.status == 0 && cs.iter < total);
  if (inv_metric_out) std::memcpy(inv_metric_out, ops.minv.data(), sizeof(double) * D);
  if (eps_out) *eps_out = cs.eps;
  if (n_leapfrog_out) *n_leapfrog_out = cs.n_leapfrog_total;
  if (n_grad_out) *n_grad_out = cs.n_grad_total;
  return cs.status;
}
