/*
 * csplotch_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar fp64) of cSplotch's per-gene posterior
 * sampling path: the log-density + gradient of the three Stan programs
 *   /root/reference/stan/splotch_stan_model.stan        (family 0)
 *   /root/reference/stan/comp_splotch_stan_model.stan   (family 1)
 *   /root/reference/stan/splotch_stan_model_csr.stan    (family 2)
 * and Stan's `hmc_nuts_diag_e_adapt` sampler (upstream stan::mcmc, CmdStan
 * defaults as invoked by /root/reference/bin/splotch:131).
 *
 * PARITY UNPINNED: CmdStan / stanc / Stan Math are third-party, un-vendored
 * and absent from this image and the reference ships no test,
 * golden vector or known-answer fixture for this path.  This oracle is pinned
 * only by (i) finite differences, (ii) 50-digit mpmath evaluation, (iii) an
 * independent torch-autograd transcription of the .stan files and (iv)
 * agreement of the three model code paths (tests/test_oracle_*.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library.  The product never does.
 */
#ifndef CSPLOTCH_ORACLE_H
#define CSPLOTCH_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Data block of the Stan programs (R-dump keys, utils.py:691-772).  All index
 * arrays are 1-based exactly as written in data_<g>.R. */
typedef struct {
  int32_t family;          /* 0 splotch_stan_model, 1 comp_splotch_stan_model, 2 splotch_stan_model_csr */
  int32_t N_tissues;
  const int32_t *N_spots;  /* [N_tissues] */
  int32_t N_covariates;
  int32_t N_celltypes;     /* family 1 only; 1 otherwise */
  int32_t N_level_1, N_level_2, N_level_3;
  int32_t zi, car, nb;
  const int32_t *level_2_mapping;  /* [N_level_2] */
  const int32_t *level_3_mapping;  /* [N_level_3] */
  const int32_t *tissue_mapping;   /* [N_tissues] */
  const double *size_factors;      /* [N] */
  const int32_t *D;                /* [N] 1..N_covariates */
  const double *E;                 /* [N][K] row-major (family 1) */
  int32_t W_n;
  const int32_t *W_sparse;         /* [W_n][2] row-major (families 0,1) */
  const int32_t *U;                /* [N+1] CSR indptr, 1-based (family 2) */
  const int32_t *V;                /* [2 W_n] CSR indices, 1-based (family 2) */
  const double *D_sparse;          /* [N] */
  const double *eig_values;        /* [N] */
  /* gene-specific */
  const int32_t *counts;           /* [N] */
  const double *beta_prior_mean;   /* [K] family 1 */
  const double *beta_prior_std;    /* [K] family 1 */
} orc_data;

typedef struct {
  int32_t num_warmup, num_samples;
  int32_t max_depth;       /* 10 */
  double delta;            /* 0.8 */
  double gamma;            /* 0.05 */
  double kappa;            /* 0.75 */
  double t0;               /* 10 */
  int32_t init_buffer;     /* 75 */
  int32_t term_buffer;     /* 50 */
  int32_t window;          /* 25 */
  double init_radius;      /* 2 */
  double stepsize;         /* 1 */
  uint32_t seed;
} orc_nuts_cfg;

/* generic density callback: returns lp, writes d lp / d q */
typedef double (*orc_density_fn)(void *ctx, const double *q, double *grad);

int32_t orc_num_spots(const orc_data *d);
int32_t orc_dim(const orc_data *d);
/* number of doubles written by orc_write_array (constrained params + transformed params) */
int32_t orc_num_outputs(const orc_data *d);

/* lp (propto=true) and gradient at unconstrained theta; returns lp. grad may be NULL. */
double orc_log_prob_grad(const orc_data *d, const double *theta, double *grad, int32_t jacobian);
/* density callbacks for orc_nuts_generic / test harnesses */
double orc_splotch_density(void *ctx, const double *q, double *grad);
double orc_gauss_density(void *ctx, const double *q, double *grad);

/* constrained parameters + transformed parameters in Stan CSV column order */
void orc_write_array(const orc_data *d, const double *theta, double *out);
/* inverse: constrained parameter block (first orc_dim entries of write_array order) -> theta */
void orc_unconstrain(const orc_data *d, const double *params, double *theta);

/* n leapfrog steps of size eps from (theta0,p0) with diagonal inverse metric;
 * traj (may be NULL) receives (n+1) rows of [theta(D), p(D), H] */
void orc_leapfrog_fixed(const orc_data *d, const double *theta0, const double *p0,
                        const double *inv_metric, double eps, int32_t n_steps,
                        double *theta_out, double *p_out, double *H_out, double *traj);

/* Philox4x32-10 (shared stream definition with the product) */
void orc_philox(uint32_t key0, uint32_t key1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                uint32_t out[4]);

/* NUTS over a generic density.  draws: [num_samples][7 + D] =
 * lp__,accept_stat__,stepsize__,treedepth__,n_leapfrog__,divergent__,energy__, q[D].
 * If save_warmup != 0 draws has num_warmup+num_samples rows.
 * inv_metric_out[D], stepsize_out: adapted values (may be NULL).
 * q_init: explicit initial point or NULL for uniform(-R,R).
 * returns 0 on success, 1 if initialisation failed. n_grad_out counts gradient evaluations. */
int32_t orc_nuts_generic(orc_density_fn fn, void *ctx, int32_t D, const orc_nuts_cfg *cfg,
                         uint32_t gene_id, uint32_t chain_id, const double *q_init,
                         int32_t save_warmup, double *draws, double *inv_metric_out,
                         double *stepsize_out, int64_t *n_grad_out);

/* NUTS on the Splotch density */
int32_t orc_nuts_run(const orc_data *d, const orc_nuts_cfg *cfg, uint32_t gene_id, uint32_t chain_id,
                     const double *q_init, int32_t save_warmup, double *draws,
                     double *inv_metric_out, double *stepsize_out, int64_t *n_grad_out);

/* gene-parallel CPU baseline: G genes (counts [G][N], priors [G][K] or NULL) x n_chains,
 * OpenMP over gene-chains (one thread per gene-chain, like CmdStan's one process per chain).
 * draws: [G][n_chains][num_samples][7 + D] or NULL (discard).  Returns total gradient evals. */
int64_t orc_nuts_batch(const orc_data *shared, const int32_t *counts, const double *prior_mean,
                       const double *prior_std, int32_t G, int32_t n_chains, const orc_nuts_cfg *cfg,
                       uint32_t gene_id0, int32_t n_threads, double *draws, int32_t *status);

/* wall-clock budget for the chains of this process from now on (bench.py's bounded CPU samples): once it has run out
 * every chain stops at its next leapfrog; seconds <= 0 removes the budget (default) */
void orc_set_time_budget(double seconds);

/* thread-busy seconds of the last orc_nuts_batch call, summed over its worker threads (0 without OpenMP) */
double orc_last_batch_busy_seconds(void);

/* batch of stand-alone gradient evaluations (CPU baseline for the lp+grad kernel) */
void orc_log_prob_grad_batch(const orc_data *shared, const int32_t *counts, const double *prior_mean,
                             const double *prior_std, int32_t G, const double *theta, int32_t B,
                             const int32_t *gene_of, double *lp, double *grad, int32_t jacobian,
                             int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif
