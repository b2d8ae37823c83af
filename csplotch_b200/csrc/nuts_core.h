// nuts_core.h — control logic of Stan's `hmc_nuts_diag_e_adapt` for ONE gene-chain,
// written once for the CTA-per-chain CUDA kernel (every thread of the CTA executes this
// code with identical scalars; the vector work is delegated to CTA-collective `Ops`).
//
// What it restates (upstream stan::mcmc, invoked by /root/reference/bin/splotch:131 with
// CmdStan defaults):
//   base_nuts::transition / build_tree      -> Nuts::transition (iterative, see below)
//   base_hmc::init_stepsize                 -> Nuts::init_stepsize
//   stepsize_adaptation::learn_stepsize     -> Nuts::da_learn
//   windowed_adaptation / var_adaptation    -> Nuts::wa_*
//   run_adaptive_sampler / initialize       -> Nuts::init_chain / Nuts::step
//
// B200-first restructuring (not a port of the recursion):
//  * build_tree is ITERATIVE.  The 2^depth leaves of a new subtree are visited in
//    integration order; a binary-counter stack keeps, per level, the pending left
//    subtree as three vector ids (p at its first leaf, p at its last leaf, rho = sum of
//    p) in a pool of D-vectors in HBM.  When leaf i completes levels 0..m-1 (m =
//    trailing ones of i) the merges run bottom-up; each is ONE fused pass computing the
//    six inner products of Stan's three U-turn criteria and writing rho of the merged
//    subtree.  p-sharp is never stored (diag metric: p_sharp . rho = sum Minv p rho).
//  * the uniform-within-subtree multinomial proposal is drawn by weighted reservoir
//    sampling in leaf order (P(leaf) = w_leaf / w_subtree, as upstream's pairwise
//    tournament) so only ONE candidate q-vector exists instead of one per level.
//  * top-level z_sample <- z_propose is a pointer swap.
//  * all randomness is counter-based (philox.h): no RNG state, no ordering dependence.
#pragma once
#include "philox.h"

namespace csp {

constexpr int kMaxDepth = 12;                 // cfg.max_depth <= kMaxDepth (CmdStan default 10)
constexpr int kPoolVecs = 3 * kMaxDepth + 8;  // D-vectors in the p/rho pool

struct NutsConfig {
  int32_t num_warmup, num_samples, max_depth;
  double delta, gamma, kappa, t0;
  int32_t init_buffer, term_buffer, window;
  double init_radius, stepsize;
  uint32_t seed;
};

// persisted per gene-chain between kernel launches
struct ChainScalars {
  int32_t status;       // 0 ok, 1 initialisation failed, 2 step-size search failed
  int32_t initialized;
  int32_t iter;         // completed transitions (warm-up + sampling)
  int32_t n_divergent;  // divergent sampling transitions
  double eps;
  double da_mu, da_counter, da_sbar, da_xbar;
  uint32_t wa_num_warmup, wa_init_buffer, wa_term_buffer, wa_base_window;
  uint32_t wa_counter, wa_size, wa_next;
  uint32_t last_nleap;  // n_leapfrog__ of the most recent transition: the cost estimate that orders the work queue
  double wel_n;
  long long n_leapfrog_total;  // sum of n_leapfrog__ over transitions
  long long n_grad_total;      // every gradient evaluation actually performed
  // device-side profile (SM clock cycles of this chain's CTA spent inside each collective op)
  long long cyc_eval, cyc_merge, cyc_copy, cyc_other, n_merge, n_copy, cyc_pro, cyc_fin;
};

struct DrawInfo {
  double lp, accept_stat, stepsize, energy;
  int32_t treedepth, n_leapfrog, divergent;
};

CSP_HD double log1p_exp(double x) { return x > 0 ? x + log1p(exp(-x)) : log1p(exp(x)); }
CSP_HD double log_sum_exp2(double a, double b) {
  if (a == -INFINITY) return b;
  if (a == INFINITY && b == INFINITY) return INFINITY;
  if (a > b) return a + log1p_exp(b - a);
  return b + log1p_exp(a - b);
}

struct SubTree {
  int pb, pe, rho;  // pool ids: p at first leaf, p at last leaf, sum of p over the leaves
};

// Ops (CTA-collective; every result is uniform over the CTA):
//   double sample_momentum(int pvec, uint32_t purpose, uint32_t sub, uint32_t iter)   -> K = p'Minv p / 2
//   double eval_at_current(int z, bool *finite)       z.q = q_cur, z.g = -grad lp(q_cur); returns V = -lp
//   void   leapfrog(int zs, int ps, int zd, int pd, double eps, double *V, double *K, int rho0, double *pp)
//          rho0 >= 0: also vec(rho0) = p(ps) + p(pd) and *pp = <p(ps), p(pd)>_M (fused level-0 merge)
//   bool   merge(SubTree L, SubTree C, int rho_out)    three U-turn criteria of L ++ C; rho_out = rho(L)+rho(C)
//   void   copy_q_to_prop(int z); void swap_prop_sample(); void commit_sample()
//   void   init_uniform(uint32_t attempt, double radius); void set_unit_metric()
//   void   welford_add(double n); void metric_update(double n)
//   void   record_draw(const DrawInfo&, int sample_index)
// Ops may offer `unsigned merge2(SubTree L1, SubTree L2, SubTree C, int rho_out)`: the merges L1 ++ C and then
// L2 ++ (L1 ++ C) in one pass; bit 0 / bit 1 of the result = the first / second subtree persists, vec(rho_out) = the sum
// of all three rho (meaningful only when both persist)
template <class Ops, class = void>
struct ops_has_merge2 { static constexpr bool value = false; };
template <class Ops>
struct ops_has_merge2<Ops, decltype((void)Ops::kHasMerge2)> { static constexpr bool value = Ops::kHasMerge2; };

template <class Ops>
struct Nuts {
  Ops &ops;
  const NutsConfig &cfg;
  ChainScalars &cs;
  RngKey key;

  CSP_HD Nuts(Ops &o, const NutsConfig &c, ChainScalars &s, RngKey k) : ops(o), cfg(c), cs(s), key(k) {}

  // ---- pool: a vector is free when nothing references it --------------------------------
  struct Refs {
    unsigned long long mask;
    CSP_HD void add(int id) { mask |= 1ull << id; }
  };
  CSP_HD static int pool_alloc(Refs &r) {
    unsigned long long free_mask = ~r.mask;
    int id = 0;
    while (!((free_mask >> id) & 1ull)) ++id;  // kPoolVecs > max live references, always terminates
    r.add(id);
    return id;
  }

  // ---- base_nuts::transition ---------------------------------------------------------------
  CSP_HD void transition(uint32_t iter, DrawInfo &info) {
    const int p0 = 0;
    const double K0 = ops.sample_momentum(p0, PURP_MOMENTUM, 0, iter);
    bool fin;
    const double V0 = ops.eval_at_current(0, &fin);  // hamiltonian.init
    cs.n_grad_total += 1;
    const double H0 = K0 + V0;

    int fwd_z = 0, bck_z = 0;
    int p_fwd = p0, p_bck = p0;
    int rho_traj = p0;
    double log_sum_weight = 0.0;
    double V_sample = V0, h_sample = H0;
    int n_leapfrog = 0, depth = 0;
    double sum_metro = 0.0;
    bool divergent = false;

    SubTree L[kMaxDepth];
    while (depth < cfg.max_depth) {
      const double u = rng_u(key, PURP_DIRECTION, 0, iter, (uint32_t)depth);
      const int sign = u > 0.5 ? 1 : -1;
      const int zsrc = sign > 0 ? fwd_z : bck_z;
      const int psrc = sign > 0 ? p_fwd : p_bck;
      const int Y = (fwd_z == bck_z) ? (1 - zsrc) : zsrc;  // the end that moves is integrated in place
      unsigned lvalid = 0;
      double lsw_run = -INFINITY, V_prop = 0.0, h_prop = 0.0;
      bool valid = true;
      SubTree cur{p0, p0, p0};
      int z_from = zsrc, p_from = psrc;
      const int nleaf = 1 << depth;
      double K_prev = 0.0;
      for (int i = 0; i < nleaf; ++i) {
        Refs refs{0ull};
        refs.add(p_fwd); refs.add(p_bck); refs.add(rho_traj); refs.add(p_from);
        for (int k = 0; k < depth; ++k)
          if ((lvalid >> k) & 1u) { refs.add(L[k].pb); refs.add(L[k].pe); refs.add(L[k].rho); }
        const int pnew = pool_alloc(refs);
        // an odd leaf closes a two-leaf subtree with its predecessor: that level-0 merge (rho = p_prev + p_new
        // and <p_prev, p_new>_M) rides along in the leapfrog pass, where both momenta are in registers
        const bool merge0 = (i & 1) != 0;
        const int rho0 = merge0 ? pool_alloc(refs) : -1;
        double Vn, Kn, pp_on = 0.0;
        ops.leapfrog(z_from, p_from, Y, pnew, sign * cs.eps, &Vn, &Kn, rho0, &pp_on);
        ++n_leapfrog;
        double h = Vn + Kn;
        if (isnan(h)) h = INFINITY;
        if ((h - H0) > 1000.0) divergent = true;
        const double w = H0 - h;
        lsw_run = log_sum_exp2(lsw_run, w);
        sum_metro += (w > 0) ? 1.0 : exp(w);
        const double ur = rng_u(key, PURP_RESERVOIR, 0, iter, (uint32_t)n_leapfrog);
        if (!divergent && ur < exp(w - lsw_run)) {
          ops.copy_q_to_prop(Y);
          V_prop = Vn;
          h_prop = h;
        }
        if (divergent) { valid = false; break; }
        cur = SubTree{pnew, pnew, pnew};
        int k = 0;
        if (merge0) {
          // L = {p_prev}, C = {p_new}: all three criteria reduce to <p_prev, rho> > 0 and <p_new, rho> > 0
          // with rho = p_prev + p_new, <p,p>_M = 2 K
          const bool persist = (2.0 * Kn + pp_on > 0) && (2.0 * K_prev + pp_on > 0);
          cur = SubTree{L[0].pb, pnew, rho0};
          lvalid &= ~1u;
          k = 1;
          if (!persist) { valid = false; break; }
        }
        while ((i >> k) & 1) {
          refs.mask = 0ull;
          refs.add(p_fwd); refs.add(p_bck); refs.add(rho_traj);
          refs.add(cur.pb); refs.add(cur.pe); refs.add(cur.rho);
          for (int kk = 0; kk < depth; ++kk)
            if ((lvalid >> kk) & 1u) { refs.add(L[kk].pb); refs.add(L[kk].pe); refs.add(L[kk].rho); }
          const int rho_out = pool_alloc(refs);
          if constexpr (ops_has_merge2<Ops>::value) {
            // the next level closes at this leaf too: both merges in ONE pass over the vectors (the inner rho sum and
            // the momenta the two share never go back to memory: 11 vector streams instead of 16)
            if ((i >> (k + 1)) & 1) {
              const unsigned pr = ops.merge2(L[k], L[k + 1], cur, rho_out);
              cur = SubTree{L[k + 1].pb, cur.pe, rho_out};
              lvalid &= ~(3u << k);
              k += 2;
              if (pr != 3u) { valid = false; break; }
              continue;
            }
          }
          const bool persist = ops.merge(L[k], cur, rho_out);
          cur = SubTree{L[k].pb, cur.pe, rho_out};
          lvalid &= ~(1u << k);
          ++k;
          if (!persist) { valid = false; break; }
        }
        if (!valid) break;
        if (k < depth) { L[k] = cur; lvalid |= 1u << k; }
        z_from = Y;
        p_from = pnew;
        K_prev = Kn;
      }
      if (!valid) break;
      ++depth;
      bool accept;
      if (lsw_run > log_sum_weight) {
        accept = true;
      } else {
        const double ua = rng_u(key, PURP_ACCEPT, 0, iter, (uint32_t)depth);
        accept = ua < exp(lsw_run - log_sum_weight);
      }
      if (accept) {
        ops.swap_prop_sample();
        V_sample = V_prop;
        h_sample = h_prop;
      }
      log_sum_weight = log_sum_exp2(log_sum_weight, lsw_run);
      // top level: the old trajectory is the "init" half in integration order
      const SubTree Ltop = sign > 0 ? SubTree{p_bck, p_fwd, rho_traj} : SubTree{p_fwd, p_bck, rho_traj};
      Refs refs{0ull};
      refs.add(p_fwd); refs.add(p_bck); refs.add(rho_traj);
      refs.add(cur.pb); refs.add(cur.pe); refs.add(cur.rho);
      const int rho_out = pool_alloc(refs);
      const bool persist = ops.merge(Ltop, cur, rho_out);
      rho_traj = rho_out;
      if (sign > 0) { p_fwd = cur.pe; fwd_z = Y; } else { p_bck = cur.pe; bck_z = Y; }
      if (!persist) break;
    }
    ops.commit_sample();
    info.n_leapfrog = n_leapfrog;
    info.treedepth = depth;
    info.divergent = divergent ? 1 : 0;
    info.accept_stat = sum_metro / (double)n_leapfrog;
    info.lp = -V_sample;
    info.energy = h_sample;
    info.stepsize = cs.eps;
    cs.n_leapfrog_total += n_leapfrog;
    cs.n_grad_total += n_leapfrog;
    cs.last_nleap = (uint32_t)n_leapfrog;
  }

  // ---- base_hmc::init_stepsize ----------------------------------------------------------------
  // q_cur is never modified here (upstream restores z_init after every attempt); the
  // gradient at q_cur is evaluated once and reused by every attempt (same value upstream).
  CSP_HD void init_stepsize(uint32_t tag_iter) {
    if (cs.eps == 0 || cs.eps > 1e7 || isnan(cs.eps)) return;
    bool fin;
    const double V0 = ops.eval_at_current(0, &fin);
    cs.n_grad_total += 1;
    const double log08 = log(0.8);
    int direction = 0;
    for (uint32_t attempt = 0;; ++attempt) {
      const double K0 = ops.sample_momentum(0, PURP_SS_MOMENTUM, attempt, tag_iter);
      double Vn, Kn;
      double pp_unused;
      ops.leapfrog(0, 0, 1, 1, cs.eps, &Vn, &Kn, -1, &pp_unused);
      cs.n_grad_total += 1;
      double h = Vn + Kn;
      if (isnan(h)) h = INFINITY;
      const double delta_H = (V0 + K0) - h;
      if (attempt == 0) {
        direction = delta_H > log08 ? 1 : -1;
        continue;
      }
      if (direction == 1 && !(delta_H > log08)) break;
      else if (direction == -1 && !(delta_H < log08)) break;
      else cs.eps = direction == 1 ? 2 * cs.eps : 0.5 * cs.eps;
      if (cs.eps > 1e7 || cs.eps == 0) { cs.status = 2; break; }
    }
  }

  // ---- stepsize_adaptation ----------------------------------------------------------------------
  CSP_HD void da_restart() { cs.da_counter = 0; cs.da_sbar = 0; cs.da_xbar = 0; }
  CSP_HD void da_learn(double adapt_stat) {
    cs.da_counter += 1;
    adapt_stat = adapt_stat > 1 ? 1 : adapt_stat;
    const double eta = 1.0 / (cs.da_counter + cfg.t0);
    cs.da_sbar = (1.0 - eta) * cs.da_sbar + eta * (cfg.delta - adapt_stat);
    const double x = cs.da_mu - cs.da_sbar * sqrt(cs.da_counter) / cfg.gamma;
    const double x_eta = pow(cs.da_counter, -cfg.kappa);
    cs.da_xbar = (1.0 - x_eta) * cs.da_xbar + x_eta * x;
    cs.eps = exp(x);
  }

  // ---- windowed_adaptation + var_adaptation ------------------------------------------------------
  CSP_HD void wa_restart() {
    cs.wa_counter = 0;
    cs.wa_size = cs.wa_base_window;
    cs.wa_next = cs.wa_init_buffer + cs.wa_size - 1;
  }
  CSP_HD void wa_setup() {
    cs.wa_num_warmup = 0; cs.wa_init_buffer = 0; cs.wa_term_buffer = 0; cs.wa_base_window = 0;
    wa_restart();
    uint32_t nw = (uint32_t)cfg.num_warmup, ib = (uint32_t)cfg.init_buffer, tb = (uint32_t)cfg.term_buffer,
             bw = (uint32_t)cfg.window;
    if (nw < 20) return;  // no variance estimation for num_warmup < 20
    if (ib + bw + tb > nw) {
      ib = (uint32_t)(0.15 * nw);
      tb = (uint32_t)(0.1 * nw);
      bw = nw - (ib + tb);
    }
    cs.wa_num_warmup = nw; cs.wa_init_buffer = ib; cs.wa_term_buffer = tb; cs.wa_base_window = bw;
    wa_restart();
  }
  CSP_HD bool wa_in_window() const {
    return (cs.wa_counter >= cs.wa_init_buffer) && (cs.wa_counter < cs.wa_num_warmup - cs.wa_term_buffer) &&
           (cs.wa_counter != cs.wa_num_warmup);
  }
  CSP_HD bool wa_end_window() const {
    return (cs.wa_counter == cs.wa_next) && (cs.wa_counter != cs.wa_num_warmup);
  }
  CSP_HD void wa_compute_next() {
    if (cs.wa_next == cs.wa_num_warmup - cs.wa_term_buffer - 1) return;
    cs.wa_size *= 2;
    cs.wa_next = cs.wa_counter + cs.wa_size;
    if (cs.wa_next == cs.wa_num_warmup - cs.wa_term_buffer - 1) return;
    const uint32_t boundary = cs.wa_next + 2 * cs.wa_size;
    if (boundary >= cs.wa_num_warmup - cs.wa_term_buffer) cs.wa_next = cs.wa_num_warmup - cs.wa_term_buffer - 1;
  }
  CSP_HD bool wa_learn_variance() {
    if (wa_in_window()) {
      cs.wel_n += 1;
      ops.welford_add(cs.wel_n);
    }
    if (wa_end_window()) {
      wa_compute_next();
      ops.metric_update(cs.wel_n);
      cs.wel_n = 0;
      ++cs.wa_counter;
      return true;
    }
    ++cs.wa_counter;
    return false;
  }

  // ---- services: initialize + run_adaptive_sampler -------------------------------------------------
  CSP_HD void init_chain(bool have_init) {
    cs.status = 0; cs.iter = 0; cs.n_divergent = 0;
    cs.n_leapfrog_total = 0; cs.n_grad_total = 0;
    cs.wel_n = 0;
    ops.set_unit_metric();
    bool ok = false;
    if (have_init) {
      ops.eval_at_current(0, &ok);
      cs.n_grad_total += 1;
    } else {
      for (uint32_t attempt = 0; attempt < 100 && !ok; ++attempt) {
        ops.init_uniform(attempt, cfg.init_radius);
        ops.eval_at_current(0, &ok);
        cs.n_grad_total += 1;
      }
    }
    cs.initialized = 1;
    if (!ok) { cs.status = 1; return; }
    cs.eps = cfg.stepsize;
    cs.da_mu = log(10 * cfg.stepsize);
    da_restart();
    wa_setup();
    init_stepsize(0);
  }

  // one iteration of generate_transitions (+ adapt_diag_e_nuts::transition during warm-up)
  CSP_HD void step() {
    const int it = cs.iter;
    const bool warm = it < cfg.num_warmup;
    if (it == cfg.num_warmup) cs.eps = exp(cs.da_xbar);  // complete_adaptation
    DrawInfo info;
    transition((uint32_t)it, info);
    if (warm) {
      da_learn(info.accept_stat);
      const bool update = wa_learn_variance();
      if (update) {
        init_stepsize((uint32_t)it + 1);
        cs.da_mu = log(10 * cs.eps);
        da_restart();
      }
    } else {
      cs.n_divergent += info.divergent;
      ops.record_draw(info, it - cfg.num_warmup);
    }
    cs.iter = it + 1;
    if (cs.iter == cfg.num_warmup && cfg.num_samples == 0) cs.eps = exp(cs.da_xbar);
  }

  CSP_HD void advance(int n_iter, bool have_init) {
    if (!cs.initialized) init_chain(have_init);
    const int total = cfg.num_warmup + cfg.num_samples;
    for (int k = 0; k < n_iter && cs.status == 0 && cs.iter < total; ++k) step();
  }
};

}  // namespace csp
