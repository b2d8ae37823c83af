// warp_ops.cuh — WARP-collective vector operations of one gene-chain (sm_100a): the small-design path.
//
// For designs of a few thousand spots (the reference's ST examples: N = 2 594, D = 5 258; examples_csplotch: N = 10) a CTA
// of 256 threads per gene-chain spends its time in CTA-wide barriers, in the restart of the TMA pipeline at every
// evaluation and in prologue / epilogue sections executed by eight warps in lock step: the staged pipeline of
// device_ops.cuh runs such designs at a quarter of the HBM roof whatever is done to its instruction count
// (profiles/r01_kernel_experiments.md, profiles/r02_kernel_experiments.md).  Here ONE WARP owns a gene-chain: 24 chains
// per SM instead of 3, no barrier wider than __syncwarp, no shared state between chains, every reduction five shuffles,
// and every warp free to be in a different phase of a different tree — the "warp-synchronous tree building" of
// BASELINE.json's north star.  nuts_core.h is unchanged: it only sees an Ops whose collectives happen to span 32 lanes.
//
// Same mathematics as device_ops.cuh (/root/reference/stan/splotch_stan_model.stan:74-222,
// comp_splotch_stan_model.stan:81-280, splotch_stan_model_csr.stan:4-271), two phases per evaluation like its
// eval_generic(): (A) half-kick + drift of every element, (B) spot loop with a CSR gather of the drifted psi (the
// chain's 42 kB vector is L1 / L2 resident), then the warp-scope epilogue.  Vectors keep the internal layout
// [ psi(Npad) | noise(Npad) | small ] so that every host-side entry point is shared with the CTA path.
#pragma once
#include "device_ops.cuh"

namespace csp {

constexpr int kWarpMaxBeta = 320;   // beta table entries (L*C*K) a warp keeps twice in shared memory

template <int KT>
struct WarpOps {
  const DevModel &m;
  GeneDev gn;
  RngKey key;
  int jacobian = 1;
  int lane;
  double *Bw, *Aw, *Sw;   // this warp's shared memory: beta levels [nbeta], their adjoints [nbeta], scalars [8]
  double *pool;
  double *zq[2], *zg[2];
  double *q_prop, *q_samp;
  double *q_cur, *minv, *wm, *wm2;
  bool sample_is_q0 = true;
  double *out_diag = nullptr, *out_small = nullptr, *out_theta = nullptr, *out_summ = nullptr;
  long long cyc_eval = 0, cyc_merge = 0, cyc_copy = 0, cyc_other = 0, n_merge = 0, n_copy = 0, cyc_pro = 0, cyc_fin = 0;

  __device__ __forceinline__ WarpOps(const DevModel &mm) : m(mm) { lane = threadIdx.x & 31; }
  __device__ __forceinline__ double *vec(int id) const { return pool + (size_t)id * m.Di; }
  template <int NV>
  __device__ __forceinline__ void wsum(double (&v)[NV]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  }

  // ---- constrained scalars + beta levels (this warp's shared memory) ----------------------------------------
  __device__ __forceinline__ void load_small(const double *q, Scalars &sc) {
    const double *sb = q + m.o_small;
    for (int i = lane; i < m.nsmall; i += 32) {
      const double v = sb[i];
      if (i < m.nbeta) Bw[i] = v;
      else Sw[i - m.nbeta] = v;
    }
    __syncwarp();
    const int last = m.nscal - 1;
    const double u_tau = Sw[max(0, min(m.s_tau, last))], u_a = Sw[max(0, min(m.s_a, last))];
    const double u_sigma = Sw[m.s_sigma], u_s2 = Sw[max(0, min(m.s_s2, last))];
    const double u_s3 = Sw[max(0, min(m.s_s3, last))], u_theta = Sw[max(0, min(m.s_theta, last))];
    const double u_phi = Sw[max(0, min(m.s_phi, last))];
    sc.tau = sc.a = sc.s2 = sc.s3 = sc.theta = 0.0;
    sc.log_tau = sc.log1m_theta = sc.log_theta = sc.phi = sc.log_phi = 0.0;
    double jac = 0.0;
    if (m.car) {
      sc.tau = exp(u_tau);
      sc.log_tau = u_tau;
      jac += u_tau;
      double la, l1a;
      logit_triple(u_a, sc.a, la, l1a);
      jac += la + l1a;
    }
    sc.sigma = exp(u_sigma);
    jac += u_sigma;
    if (m.L2) { sc.s2 = exp(u_s2); jac += u_s2; }
    if (m.L3) { sc.s3 = exp(u_s3); jac += u_s3; }
    if (m.zi) {
      logit_triple(u_theta, sc.theta, sc.log_theta, sc.log1m_theta);
      jac += sc.log_theta + sc.log1m_theta;
    }
    if (m.nb) {
      sc.phi = 1e-6 + exp(u_phi);
      sc.log_phi = log(sc.phi);
      jac += u_phi;
    }
    sc.jac = jacobian ? jac : 0.0;
    const int CK = m.C * m.K;
    for (int idx = lane; idx < m.L1 * CK; idx += 32) {
      const double r = Bw[idx];
      Bw[idx] = (m.family == 1) ? r * gn.ps[idx % m.K] + gn.pm[idx % m.K] : 2.0 * r;
    }
    __syncwarp();
    for (int idx = lane; idx < m.L2 * CK; idx += 32) {
      const int i = idx / CK, rem = idx - i * CK;
      Bw[(m.L1 + i) * CK + rem] = Bw[m.map2[i] * CK + rem] + sc.s2 * Bw[(m.L1 + i) * CK + rem];
    }
    __syncwarp();
    for (int idx = lane; idx < m.L3 * CK; idx += 32) {
      const int i = idx / CK, rem = idx - i * CK;
      Bw[(m.L1 + m.L2 + i) * CK + rem] = Bw[(m.L1 + m.map3[i]) * CK + rem] + sc.s3 * Bw[(m.L1 + m.L2 + i) * CK + rem];
    }
    __syncwarp();
  }

  // bottom-level bins of one run of equal keys: the warp is alone on its adjoint table, so the leader lane adds
  // without atomics and the order of the additions is fixed (spot order): bitwise reproducible
  __device__ __forceinline__ void flush_bins(double (&acc)[KT], int cur_key) {
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
          if (lane == leader) Aw[(m.bot0 * m.C) * m.K + k0 * m.K + k] += s;
        }
      }
      __syncwarp();
      remaining &= ~same;
    }
#pragma unroll
    for (int k = 0; k < KT; ++k) acc[k] = 0.0;
  }

  // ---- fused [half-kick, drift,] log-density + gradient [, half-kick] -------------------------------------------
  // LEAP: (sq,sg,sp) --eps--> (dq,dg,dp); !LEAP: dq := sq (if different), dg := -grad lp(sq).  rho_out != nullptr
  // (LEAP): also rho_out = sp + p_new and *pp_out = <sp, p_new>_M (level-0 U-turn merge, nuts_core.h).
  template <bool LEAP>
  __device__ __noinline__ void eval(const double *sq, const double *sg, const double *sp, double *dq, double *dg, double *dp,
                                    double eps, double *V_out, double *K_out, bool *finite_out, double *rho_out = nullptr,
                                    double *pp_out = nullptr) {
    const int Di = m.Di;
    // phase A: every element moves before any gradient is taken (the spot loop gathers psi of neighbours)
    if (LEAP) {
      for (int e = 2 * lane; e < Di; e += 64) {
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
      for (int e = 2 * lane; e < Di; e += 64) *reinterpret_cast<double2 *>(dq + e) = *reinterpret_cast<const double2 *>(sq + e);
    }
    __syncwarp();
    Scalars sc;
    load_small(dq, sc);
    for (int idx = lane; idx < m.nbeta; idx += 32) Aw[idx] = 0.0;
    __syncwarp();

    const double heads = sc.log_theta, tails = sc.log1m_theta;
    const double sig03 = 0.3 * sc.sigma;
    const double nb_mult = (m.nb == 1 && m.zi) ? (double)m.N : 1.0;
    const double lgam_phi = m.nb ? lgamma(sc.phi) : 0.0, dig_phi = m.nb ? digamma_d(sc.phi) : 0.0;
    const double theta = sc.theta, om_theta = 1.0 - sc.theta;
    const double *psi = dq + m.o_psi, *noi = dq + m.o_noise;
    double a_lp = 0, a_ng = 0, a_qD = 0, a_qW = 0, a_dth = 0, a_kin = 0, a_g2 = 0, a_pp = 0, a_dphi = 0;
    double a_ldet = 0, a_dldet = 0;
    LogProd zi_prod;
    zi_prod.init();
    double acc[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) acc[k] = 0.0;
    int cur_key = -1;
    const bool merge0 = LEAP && rho_out != nullptr;

    for (int base = 0; base < m.N; base += 32) {
      const int j = base + lane;
      const bool act = j < m.N;
      int kj = -1;
      double gj = 0.0;
      double e[KT];
      if (act) {
        const int kraw = m.key[j];
        kj = kraw & 0xFFFFFF;
        const double eps_j = noi[j];
        const double *bj = Bw + (m.bot0 * m.C) * m.K + kj * m.K;
        double ll;
        if (KT == 1) {
          ll = bj[0];
          e[0] = 1.0;
        } else {
          ll = 0.0;
          const double *Ej = E_ptr(m, j);
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            e[k] = (k < m.K) ? __ldg(Ej + (size_t)k * kTile) : 0.0;
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
        const int y = gn.counts[j];
        if (m.nb) {
          // neg_binomial_2_log (comp_splotch_stan_model.stan:248-265), as in DevOps::eval_generic
          const double yd = (double)y;
          const double dl = eta - sc.log_phi;
          const double L = dl > 0 ? dl + log1p(exp(-dl)) : log1p(exp(dl));
          const double sgm = inv_logit_d(dl);
          const double nbv = lgamma(yd + sc.phi) - lgamma(yd + 1.0) - lgam_phi + yd * eta - sc.phi * L - yd * (sc.log_phi + L);
          const double dnb_deta = yd - (sc.phi + yd) * sgm;
          const double dnb_dphi = digamma_d(yd + sc.phi) - dig_phi - L + (sc.phi + yd) * sgm / sc.phi - yd / sc.phi;
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
            a_lp += nb_mult * nbv;
            gj = nb_mult * dnb_deta;
            a_dphi += nb_mult * dnb_dphi;
          }
        } else {
          const double mu = exp_fast(eta);
          if (m.zi && y == 0) {
            // log(theta + (1-theta) e^{-mu}) summed as the log of a running product (device_ops.cuh); the d/dtheta part
            // accumulated here is only the zero-count term, finish adds -(#{y>0} + 1) / (1 - theta)
            const double em = exp_fast(-mu);
            const double z = om_theta * em;
            const double s = theta + z;
            double inv_s;
            if (s > 1e-290) {
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
        }
        a_lp -= 0.5 * eps_j * eps_j;
        a_ng += eps_j * gj;
        const double gr_eps = sig03 * gj - eps_j;
        dg[m.o_noise + j] = -gr_eps;
        a_g2 += gr_eps * gr_eps;
        if (LEAP) {
          const double mi = minv[m.o_noise + j];
          const double pn = dp[m.o_noise + j] + 0.5 * eps * gr_eps;
          dp[m.o_noise + j] = pn;
          a_kin += mi * pn * pn;
          if (merge0) {
            const double po = sp[m.o_noise + j];
            a_pp += mi * po * pn;
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
            const double mi = minv[m.o_psi + j];
            const double pn = dp[m.o_psi + j] + 0.5 * eps * gr_psi;
            dp[m.o_psi + j] = pn;
            a_kin += mi * pn * pn;
            if (merge0) {
              const double po = sp[m.o_psi + j];
              a_pp += mi * po * pn;
              rho_out[m.o_psi + j] = po + pn;
            }
          }
        }
      }
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
    for (int j = m.N + lane; j < m.Npad; j += 32) {
      dg[m.o_noise + j] = 0.0;
      if (LEAP) dp[m.o_noise + j] = 0.0;
      if (merge0) rho_out[m.o_noise + j] = 0.0;
      if (m.car) {
        dg[m.o_psi + j] = 0.0;
        if (LEAP) dp[m.o_psi + j] = 0.0;
        if (merge0) rho_out[m.o_psi + j] = 0.0;
      }
    }
    if (m.zi && !m.nb) a_lp += zi_prod.log_value();
    // in the NB branch a_dth already holds the full softmax form; the Poisson / ZIP branch follows device_ops.cuh
    // (finish adds -(nnz + 1) / (1 - theta)); keep the two conventions apart
    const bool dth_is_partial = !m.nb;

    // ---- epilogue: log-determinant, prior of beta_raw, reverse of the hierarchy, small-block gradient ----------
    if (m.car)
      for (int i = lane; i < m.n_eigg; i += 32)
        eig_group_terms(m.eigg + (size_t)i * kEigGroup, __ldg(m.eiggm + i), sc.a, a_ldet, a_dldet);
    __syncwarp();
    const int CK = m.C * m.K;
    if (m.L3) {
      for (int idx = lane; idx < m.L2 * CK; idx += 32) {
        const int i2 = idx / CK, rem = idx - i2 * CK;
        double s = Aw[(m.L1 + i2) * CK + rem];
        for (int i3 = 0; i3 < m.L3; ++i3)
          if (m.map3[i3] == i2) s += Aw[(m.L1 + m.L2 + i3) * CK + rem];
        Aw[(m.L1 + i2) * CK + rem] = s;
      }
      __syncwarp();
    }
    if (m.L2) {
      for (int idx = lane; idx < m.L1 * CK; idx += 32) {
        const int i1 = idx / CK, rem = idx - i1 * CK;
        double s = Aw[i1 * CK + rem];
        for (int i2 = 0; i2 < m.L2; ++i2)
          if (m.map2[i2] == i1) s += Aw[(m.L1 + i2) * CK + rem];
        Aw[i1 * CK + rem] = s;
      }
      __syncwarp();
    }
    const double *sb = dq + m.o_small;
    double a_ds2 = 0, a_ds3 = 0;
    for (int idx = lane; idx < m.nbeta; idx += 32) {
      const int i = idx / CK;
      const double ab = Aw[idx], r = sb[idx];
      a_lp -= 0.5 * r * r;
      double scale;
      if (i < m.L1) scale = (m.family == 1) ? gn.ps[idx % m.K] : 2.0;
      else if (i < m.L1 + m.L2) { scale = sc.s2; a_ds2 += ab * r; }
      else { scale = sc.s3; a_ds3 += ab * r; }
      const double gr = scale * ab - r;
      dg[m.o_small + idx] = -gr;
      a_g2 += gr * gr;
      if (LEAP) {
        const double mi = minv[m.o_small + idx];
        const double pn = dp[m.o_small + idx] + 0.5 * eps * gr;
        dp[m.o_small + idx] = pn;
        a_kin += mi * pn * pn;
        if (merge0) {
          const double po = sp[m.o_small + idx];
          a_pp += mi * po * pn;
          rho_out[m.o_small + idx] = po + pn;
        }
      }
    }
    double r1[13] = {a_lp, a_ng, a_qD, a_qW, a_dth, a_kin, a_ldet, a_dldet, a_g2, a_dphi, a_ds2, a_ds3, a_pp};
    wsum<13>(r1);
    double lp = r1[0] + sc.jac;
    double kin = r1[5], g2 = r1[8], pp = r1[12];
    double gu[7] = {0, 0, 0, 0, 0, 0, 0};
    if (m.car) {
      const double Q = r1[2] - sc.a * r1[3];
      const double inv_tau = 1.0 / sc.tau;
      lp += -2.0 * sc.log_tau - inv_tau;
      lp += 0.5 * ((double)m.N * sc.log_tau + r1[6] - sc.tau * Q);
      const double d_tau = -2.0 * inv_tau + inv_tau * inv_tau + 0.5 * ((double)m.N * inv_tau - Q);
      const double d_a = 0.5 * (r1[7] + sc.tau * r1[3]);
      gu[0] = sc.tau * d_tau + (jacobian ? 1.0 : 0.0);
      gu[1] = sc.a * (1.0 - sc.a) * d_a + (jacobian ? (1.0 - 2.0 * sc.a) : 0.0);
    }
    lp += -0.5 * sc.sigma * sc.sigma;
    gu[2] = sc.sigma * (-sc.sigma + 0.3 * r1[1]) + (jacobian ? 1.0 : 0.0);
    if (m.L2) { lp += -0.5 * sc.s2 * sc.s2; gu[3] = sc.s2 * (-sc.s2 + r1[10]) + (jacobian ? 1.0 : 0.0); }
    if (m.L3) { lp += -0.5 * sc.s3 * sc.s3; gu[4] = sc.s3 * (-sc.s3 + r1[11]) + (jacobian ? 1.0 : 0.0); }
    if (m.zi) {
      lp += tails;   // theta ~ beta(1,2)
      double d_theta;
      if (dth_is_partial) {
        lp += gn.nnz * tails - gn.lgam;   // bernoulli(0|theta) and -lgamma(y+1) of the non-zero spots
        d_theta = r1[4] - (gn.nnz + 1.0) / (1.0 - sc.theta);
      } else {
        lp += gn.nnz * tails;
        d_theta = r1[4] - (gn.nnz + 1.0) / (1.0 - sc.theta);
      }
      gu[5] = sc.theta * (1.0 - sc.theta) * d_theta + (jacobian ? (1.0 - 2.0 * sc.theta) : 0.0);
    }
    if (m.nb) {
      lp += sc.log_phi - sc.phi;  // phi ~ gamma(2,1)
      gu[6] = (sc.phi - 1e-6) * (r1[9] + 1.0 / sc.phi - 1.0) + (jacobian ? 1.0 : 0.0);
    }
    const int so = m.o_small + m.nbeta;
    const int sidx[7] = {m.s_tau, m.s_a, m.s_sigma, m.s_s2, m.s_s3, m.s_theta, m.s_phi};
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      if (sidx[s] >= 0) {
        const int e = so + sidx[s];
        g2 += gu[s] * gu[s];
        double pn = 0.0, po = 0.0, mi = 0.0;
        if (LEAP) {
          mi = minv[e];
          pn = dp[e] + 0.5 * eps * gu[s];
          kin += mi * pn * pn;
          if (merge0) { po = sp[e]; pp += mi * po * pn; }
        }
        __syncwarp();   // every lane has read dp[e] / sp[e] before lane 0 overwrites them
        if (lane == 0) {
          dg[e] = -gu[s];
          if (LEAP) dp[e] = pn;
          if (merge0) rho_out[e] = po + pn;
        }
      }
    }
    if (lane == 0)
      for (int s = m.nsmall; s < m.nsmall_pad; ++s) {
        dg[m.o_small + s] = 0.0;
        if (LEAP) dp[m.o_small + s] = 0.0;
        if (merge0) rho_out[m.o_small + s] = 0.0;
      }
    __syncwarp();
    if (pp_out) *pp_out = pp;
    if (V_out) *V_out = -lp;
    if (K_out) *K_out = 0.5 * kin;
    if (finite_out) *finite_out = isfinite(lp) && isfinite(g2);
  }

  // ---- Ops interface of nuts_core.h ---------------------------------------------------------------------------
  __device__ __noinline__ double sample_momentum(int pv, uint32_t purpose, uint32_t sub, uint32_t iter) {
    const long long t0 = clock64();
    double *p = vec(pv);
    double k = 0.0;
    for (int e = lane; e < m.Di; e += 32) {
      const int si = m.imap[e];
      double v = 0.0;
      if (si >= 0) {
        const double mi = minv[e];
        v = rng_normal(key, purpose, sub, iter, (uint32_t)si) / sqrt(mi);
        k += mi * v * v;
      }
      p[e] = v;
    }
    __syncwarp();
    sample_is_q0 = true;
    cyc_other += clock64() - t0;
    return 0.5 * warp_sum(k);
  }
  __device__ double eval_at_current(int z, bool *finite) {
    double V;
    const long long t0 = clock64();
    eval<false>(q_cur, nullptr, nullptr, zq[z], zg[z], nullptr, 0.0, &V, nullptr, finite);
    cyc_eval += clock64() - t0;
    return V;
  }
  __device__ void leapfrog(int zs, int ps, int zd, int pd, double eps, double *V, double *K, int rho0, double *pp) {
    const long long t0 = clock64();
    eval<true>(zq[zs], zg[zs], vec(ps), zq[zd], zg[zd], vec(pd), eps, V, K, nullptr, rho0 >= 0 ? vec(rho0) : nullptr, pp);
    cyc_eval += clock64() - t0;
  }
  __device__ __noinline__ bool merge(SubTree L, SubTree C, int rho_out) {
    const long long t0 = clock64();
    const double *Lpb = vec(L.pb), *Lpe = vec(L.pe), *Lrho = vec(L.rho);
    const double *Cpb = vec(C.pb), *Cpe = vec(C.pe), *Crho = vec(C.rho);
    double *out = vec(rho_out);
    double a1 = 0, a2 = 0, b1 = 0, b2 = 0, c1 = 0, c2 = 0;
    for (int e = 2 * lane; e < m.Di; e += 64) {
      const double2 w = CSP_LD2(minv, e), lpb = CSP_LD2(Lpb, e), lpe = CSP_LD2(Lpe, e), lrh = CSP_LD2(Lrho, e);
      const double2 cpb = CSP_LD2(Cpb, e), cpe = CSP_LD2(Cpe, e), crh = CSP_LD2(Crho, e);
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
      *reinterpret_cast<double2 *>(out + e) = rs;
    }
    double r[6] = {a1, a2, b1, b2, c1, c2};
    wsum<6>(r);
    __syncwarp();
    cyc_merge += clock64() - t0;
    ++n_merge;
    return (r[1] > 0 && r[0] > 0) && (r[3] > 0 && r[2] > 0) && (r[5] > 0 && r[4] > 0);
  }
  __device__ __forceinline__ void copy_vec(double *dst, const double *src) {
    for (int e = 2 * lane; e < m.Di; e += 64) *reinterpret_cast<double2 *>(dst + e) = CSP_LD2(src, e);
    __syncwarp();
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
    for (int e = lane; e < m.Di; e += 32) {
      const int si = m.imap[e];
      q_cur[e] = si >= 0 ? rng_init(key, attempt, (uint32_t)si, radius) : 0.0;
    }
    __syncwarp();
  }
  __device__ __noinline__ void set_unit_metric() {
    for (int e = lane; e < m.Di; e += 32) { minv[e] = 1.0; wm[e] = 0.0; wm2[e] = 0.0; }
    __syncwarp();
  }
  __device__ __noinline__ void welford_add(double n) {
    for (int e = lane; e < m.Di; e += 32) {
      const double q = q_cur[e];
      const double delta = q - wm[e];
      const double mean = wm[e] + delta / n;
      wm[e] = mean;
      wm2[e] += (q - mean) * delta;
    }
    __syncwarp();
  }
  __device__ __noinline__ void metric_update(double n) {
    for (int e = lane; e < m.Di; e += 32) {
      double v = minv[e];
      if (n > 1) v = wm2[e] / (n - 1.0);
      minv[e] = (m.imap[e] >= 0) ? (n / (n + 5.0)) * v + 1e-3 * (5.0 / (n + 5.0)) : 1.0;
      wm[e] = 0.0;
      wm2[e] = 0.0;
    }
    __syncwarp();
  }
  __device__ __noinline__ void record_draw(const DrawInfo &info, int s) {
    if (out_diag && lane == 0) {
      double *r = out_diag + (size_t)s * 7;
      r[0] = info.lp; r[1] = info.accept_stat; r[2] = info.stepsize; r[3] = info.treedepth;
      r[4] = info.n_leapfrog; r[5] = info.divergent; r[6] = info.energy;
    }
    if (out_small)
      for (int i = lane; i < m.nsmall; i += 32) {
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
      for (int e = lane; e < m.Di; e += 32) {
        const int si = m.imap[e];
        if (si >= 0) out_theta[(size_t)s * m.D + si] = q_cur[e];
      }
    if (out_summ) {
      Scalars sc;
      load_small(q_cur, sc);
      const double n = (double)(s + 1);
      const double sig03 = 0.3 * sc.sigma;
      double *mll = out_summ, *vll = out_summ + m.N, *mla = out_summ + 2 * (size_t)m.N, *vla = out_summ + 3 * (size_t)m.N,
             *mps = out_summ + 4 * (size_t)m.N, *vps = out_summ + 5 * (size_t)m.N;
      for (int j = lane; j < m.N; j += 32) {
        const double *bj = Bw + (m.bot0 * m.C) * m.K + (m.key[j] & 0xFFFFFF) * m.K;
        double ll;
        if (KT == 1) {
          ll = bj[0];
        } else {
          ll = 0.0;
          const double *Ej = E_ptr(m, j);
          for (int k = 0; k < m.K; ++k) ll += bj[k] * Ej[(size_t)k * kTile];
        }
        double psi_j = 0.0;
        if (m.car) { psi_j = q_cur[m.o_psi + j]; ll += psi_j; }
        ll += sig03 * q_cur[m.o_noise + j];
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
    __syncwarp();
  }
};

}  // namespace csp
