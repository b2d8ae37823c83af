/*
 * csplotch.h — C ABI of the B200-native cSplotch sampling path.
 *
 * This library replaces ONE call of the reference: the per-gene, per-chain
 * CmdStan process that /root/reference/bin/splotch:129-133 forks
 *   "$BINARY" sample num_samples=N num_warmup=N random id=CHAIN \
 *        data file=DATA/<g/100>/data_<g>.R output file=OUT/.../output_<g>_<c>.csv
 * (second call site: pystan .sampling(), /root/reference/simulation/fitting.py:35-48)
 * with batched NUTS over many genes x chains on one GPU.  Every entry point
 * cites the reference interface it stands in for.  Plain pointers and sizes
 * only; no torch types.  All functions return 0 on success, a CSPLOTCH_E_*
 * code otherwise; csplotch_last_error() holds the message (thread-local).
 *
 * Memory spaces: every array argument is a HOST pointer unless its name ends
 * in `_dev` or the function name ends in `_dev` (then it is a CUDA device
 * pointer on the library's current device and `stream` is a cudaStream_t
 * passed as void*; NULL = default stream).
 *
 * Index conventions at this boundary are the R-dump's: 1-based mappings, D,
 * W_sparse, U, V exactly as `generate_dictionary` writes them
 * (/root/reference/splotch/utils.py:684-774).  The unconstrained vector theta
 * follows the parameters{} declaration order of the .stan files (SURVEY.md
 * A.1): psi[N] | beta_raw | tau | a | sigma | sigma_level_2 | sigma_level_3 |
 * theta | noise_raw[N], with beta_raw column-major (i + L*j) for
 * splotch_stan_model / _csr and (i*C + j)*K + k for comp_splotch_stan_model.
 */
#ifndef CSPLOTCH_H
#define CSPLOTCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSPLOTCH_ABI_VERSION 1

enum {
  CSPLOTCH_OK = 0,
  CSPLOTCH_E_INVALID = 1,     /* bad argument / inconsistent data block */
  CSPLOTCH_E_UNSUPPORTED = 2, /* e.g. N_celltypes > 16 */
  CSPLOTCH_E_CUDA = 3,        /* CUDA runtime error; message has the cause */
  CSPLOTCH_E_NOMEM = 4
};

/* `-b BINARY` of bin/splotch selects the model family by basename */
enum {
  CSPLOTCH_SPLOTCH = 0,      /* stan/splotch_stan_model.stan */
  CSPLOTCH_COMP_SPLOTCH = 1, /* stan/comp_splotch_stan_model.stan */
  CSPLOTCH_SPLOTCH_CSR = 2   /* stan/splotch_stan_model_csr.stan */
};

typedef struct csplotch_model csplotch_model;     /* gene-shared covariates, resident in HBM */
typedef struct csplotch_batch csplotch_batch;     /* per-gene counts (+ beta priors) of G genes */
typedef struct csplotch_sampler csplotch_sampler; /* NUTS state of G genes x n_chains chains */

/* The gene-independent part of the Stan data{} block
 * (stan/splotch_stan_model.stan:20-50, comp_…:20-57, _csr:36-70). */
typedef struct {
  int32_t family;
  int32_t N_tissues;
  const int32_t *N_spots;          /* [N_tissues] */
  int32_t N_covariates;
  int32_t N_celltypes;             /* comp only, else 1 */
  int32_t N_level_1, N_level_2, N_level_3;
  int32_t zi, car;
  int32_t nb;                      /* comp only: 0 Poisson/ZIP; 1 NB/ZINB exactly as comp_...stan:248-265 evaluates it (the
                                      ZINB branch passes the vector phi_arr to a scalar-count lpmf, so every spot's NB
                                      term counts N times); 2 = ZINB with the per-spot term counted once */
  const int32_t *level_2_mapping;  /* [N_level_2], 1-based */
  const int32_t *level_3_mapping;  /* [N_level_3], 1-based */
  const int32_t *tissue_mapping;   /* [N_tissues], 1-based */
  const double *size_factors;      /* [N] */
  const int32_t *D;                /* [N], 1..N_covariates */
  const double *E;                 /* [N][K] row-major, comp only */
  int32_t W_n;                     /* adjacent pairs (0 if car == 0) */
  const int32_t *W_sparse;         /* [W_n][2] row-major, 1-based, or NULL if U,V given */
  const int32_t *U;                /* [N+1] CSR indptr, 1-based, or NULL (derived from W_sparse) */
  const int32_t *V;                /* [2 W_n] CSR indices, 1-based, or NULL */
  const double *D_sparse;          /* [N] */
  const double *eig_values;        /* [N] */
} csplotch_shared_desc;

/* `sample` arguments of the CmdStan executable; defaults = CmdStan's
 * (bin/splotch:131 passes only num_samples, num_warmup, id). */
typedef struct {
  int32_t num_warmup;
  int32_t num_samples;
  int32_t num_chains;   /* -c NUM_CHAINS; chain ids are 1..num_chains like `id=CHAIN` */
  int32_t max_depth;    /* 10 */
  double delta;         /* 0.8 */
  double gamma;         /* 0.05 */
  double kappa;         /* 0.75 */
  double t0;            /* 10 */
  int32_t init_buffer;  /* 75 */
  int32_t term_buffer;  /* 50 */
  int32_t window;       /* 25 */
  double init_radius;   /* 2 */
  double stepsize;      /* 1 */
  uint32_t seed;
  int32_t keep_draws;   /* 1: keep every sampling draw of the full unconstrained vector (CSV output) */
} csplotch_nuts_cfg;

typedef struct {
  int32_t N;            /* sum(N_spots) */
  int32_t dim;          /* length of theta */
  int32_t n_beta;       /* L*C*K */
  int32_t n_scalar;     /* tau,a,sigma,... present in this model */
  int32_t n_small;      /* n_beta + n_scalar: the non-spot block written per draw */
  int32_t n_outputs;    /* doubles per CSV row after the 7 sampler columns (write_array) */
} csplotch_dims;

const char *csplotch_last_error(void);
int32_t csplotch_abi_version(void);
/* number of CUDA devices visible; selects the device used by this thread's later calls */
int32_t csplotch_device_count(void);
int32_t csplotch_set_device(int32_t device);

/* (no reference counterpart: /root/reference/bin/splotch moves data through files and pipes)
 * Page-locked host buffers for arrays handed to / filled by the calls below (optional: every host pointer may
 * also be ordinary pageable memory).  The reference moves these data through files and pipes; here they cross
 * PCIe, and a pinned buffer makes each transfer a single DMA.  csplotch_host_free(NULL) is a no-op. */
int32_t csplotch_host_alloc(uint64_t bytes, void **out);
int32_t csplotch_host_free(void *p);

/* ---- data{} + transformed data{} (stan/splotch_stan_model.stan:20-72) ---------------- */
int32_t csplotch_model_create(const csplotch_shared_desc *desc, csplotch_model **out);
void csplotch_model_destroy(csplotch_model *m);
int32_t csplotch_model_dims(const csplotch_model *m, csplotch_dims *out);

/* counts[G][N] (int32 as Stan reads them); beta_prior_mean/std [G][K] for comp, else NULL.
 * gene_id0 = global index of the first gene (keys the RNG stream so that results do not
 * depend on how genes are sharded). */
int32_t csplotch_batch_create(csplotch_model *m, const int32_t *counts, const double *beta_prior_mean,
                              const double *beta_prior_std, int32_t G, uint32_t gene_id0,
                              csplotch_batch **out);
/* same, counts already in HBM */
int32_t csplotch_batch_create_dev(csplotch_model *m, const int32_t *counts_dev,
                                  const double *beta_prior_mean, const double *beta_prior_std, int32_t G,
                                  uint32_t gene_id0, void *stream, csplotch_batch **out);
void csplotch_batch_destroy(csplotch_batch *b);
/* explicit ids gene_ids[G] of the genes of a batch instead of gene_id0 + g (call before csplotch_sampler_create): the
 * id keys the RNG stream, so one gene list can be cost-sorted or dealt round-robin over the GPUs (the job array
 * around /root/reference/bin/splotch:129-133 hands out arbitrary gene indices too) and every gene still gets the
 * draws it would get alone */
int32_t csplotch_batch_set_gene_ids(csplotch_batch *b, const uint32_t *gene_ids);

/* ---- model.log_prob_grad<propto=true, jacobian>() of the stanc-generated class ------- */
/* theta[B][dim], gene_of[B] (index into the batch, or NULL for b % G); lp[B]; grad[B][dim] or NULL */
int32_t csplotch_log_prob_grad(csplotch_batch *b, const double *theta, const int32_t *gene_of, int32_t B,
                               double *lp, double *grad, int32_t jacobian);
int32_t csplotch_log_prob_grad_dev(csplotch_batch *b, const double *theta_dev, const int32_t *gene_of_dev,
                                   int32_t B, double *lp_dev, double *grad_dev, int32_t jacobian,
                                   void *stream);

/* ---- model.write_array(): constrained params + transformed params, CSV column order ---
 * (parameters{} / transformed parameters{} of stan/splotch_stan_model.stan:74-130, comp_…:81-150; the columns that
 * read_stan_csv maps back by dotted name, /root/reference/splotch/utils.py:196-224) */
int32_t csplotch_write_array(csplotch_batch *b, const double *theta, const int32_t *gene_of, int32_t B,
                             double *out /* [B][n_outputs] */);

/* ---- expl_leapfrog with a diagonal metric, fixed step (tier-2 parity) ------------------
 * (upstream stan::mcmc::expl_leapfrog + diag_e_metric, the integrator behind /root/reference/bin/splotch:131;
 * not in the reference tree — a test hook, no reference call site binds it) */
/* theta0,p0,inv_metric: [B][dim]; outputs [B][dim], H_out[B] */
int32_t csplotch_leapfrog_fixed(csplotch_batch *b, const double *theta0, const double *p0,
                                const double *inv_metric, const int32_t *gene_of, int32_t B, double eps,
                                int32_t n_steps, double *theta_out, double *p_out, double *H_out);

/* diagnostic (no reference call site): device-timed microbenchmark of the fused leapfrog pass — B trajectories of
 * n_steps fixed-size steps from deterministic pseudo-random states, average CUDA-event milliseconds of `reps` launches
 * after one warm-up launch; per launch B * (1 gradient + n_steps leapfrogs) are evaluated, nothing crosses PCIe */
int32_t csplotch_leapfrog_bench(csplotch_batch *b, int32_t B, double eps, int32_t n_steps, int32_t reps, float *ms_avg);

/* ---- `BINARY sample ...`: hmc_nuts_diag_e_adapt over all genes x chains ----------------
 * (/root/reference/bin/splotch:129-133: one process per chain with num_samples, num_warmup, id;
 * /root/reference/bin/splotch_vinit:194 adds init=FILE.json -> csplotch_sampler_init;
 * /root/reference/simulation/fitting.py:35-48: pystan .sampling(iter, chains)) */
void csplotch_nuts_cfg_default(csplotch_nuts_cfg *cfg);
int32_t csplotch_sampler_create(csplotch_batch *b, const csplotch_nuts_cfg *cfg, csplotch_sampler **out);
void csplotch_sampler_destroy(csplotch_sampler *s);
/* optional explicit initial points theta_init[G][num_chains][dim] (CmdStan `init=`); NULL = uniform(-R,R) */
int32_t csplotch_sampler_init(csplotch_sampler *s, const double *theta_init);
/* run up to n_iter further transitions per chain (warm-up first, then sampling); asynchronous on
 * the sampler's stream unless sync != 0 */
int32_t csplotch_sampler_advance(csplotch_sampler *s, int32_t n_iter, int32_t sync);
int32_t csplotch_sampler_sync(csplotch_sampler *s);
/* device time (ms) of the last completed advance, measured with CUDA events on the sampler's stream */
int32_t csplotch_sampler_last_ms(csplotch_sampler *s, float *ms);
/* totals over all chains so far: leapfrogs (= n_leapfrog__ summed), gradient evaluations
 * (leapfrogs + per-transition re-initialisations + step-size searches), divergences, kernel launches */
int32_t csplotch_sampler_counters(csplotch_sampler *s, int64_t *n_leapfrog, int64_t *n_grad,
                                  int64_t *n_divergent, int64_t *n_launches);
/* device-side profile summed over chains: SM cycles inside {log-density/leapfrog, U-turn merge passes,
 * candidate copies, everything else}, the number of merge passes and candidate copies, then the cycles of
 * the log-density pass spent before its tile loop (prologue) and after it (reductions, small block) */
int32_t csplotch_sampler_profile(csplotch_sampler *s, int64_t *out8);
/* how the sampler's launches are laid out: gene-chains resident at once, and the CTAs of the thread-block cluster that
 * works on each (1 for short designs or thousands of chains; 2 or 4 when few chains of a long design would otherwise
 * leave the launch waiting for its deepest tree; CSPLOTCH_CLUSTER=1|2|4 at csplotch_model_create forces a size) */
int32_t csplotch_sampler_layout(csplotch_sampler *s, int32_t *resident_chains, int32_t *cluster_size);
/* diagnostic: nanoseconds every resident CTA of the last launch was busy (ns[n_slots]; pass ns = NULL to query n_slots):
 * max - mean is the tail of the launch that the longest-first queue order is there to shorten */
int32_t csplotch_sampler_slot_ns(csplotch_sampler *s, int64_t *ns, int32_t cap, int32_t *n_slots);
/* per-chain status [G][num_chains]: 0 ok, 1 initialisation failed, 2 step size search failed */
int32_t csplotch_sampler_status(csplotch_sampler *s, int32_t *status, int32_t *iters_done);

/* sampling draws so far: the rows CmdStan writes to `output file=…csv` and the wrapper concatenates
 * (/root/reference/bin/splotch:131,138-140).
 * diag  [G][num_chains][num_samples][7]  lp__,accept_stat__,stepsize__,treedepth__,n_leapfrog__,divergent__,energy__
 * small [G][num_chains][num_samples][n_small]  the UNCONSTRAINED non-spot block: beta_raw[(i*C + j)*K + k]
 *        (level-major rows, the same for all three families) followed by tau, a, sigma, sigma_level_2,
 *        sigma_level_3, theta as present; beta_level_1..3 follow from it (csplotch_b200/api.py);
 * n_draws [G][num_chains] = sampling iterations completed. Any pointer may be NULL. */
int32_t csplotch_sampler_get_draws(csplotch_sampler *s, double *diag, double *small, int32_t *n_draws);
/* full unconstrained draws [G][num_chains][num_samples][dim] (needs cfg.keep_draws) */
int32_t csplotch_sampler_get_theta_draws(csplotch_sampler *s, double *theta);
/* the same for genes g0 .. g0 + n_genes - 1 of the batch only: theta [n_genes][num_chains][num_samples][dim] (the CSV
 * writer walks over the batch a few genes at a time; all draws of a c3-sized batch exceed host memory) */
int32_t csplotch_sampler_get_theta_draws_range(csplotch_sampler *s, int32_t g0, int32_t n_genes, double *theta);
/* current position, adapted step size and inverse metric: [G][num_chains][dim], [G][num_chains] */
int32_t csplotch_sampler_get_state(csplotch_sampler *s, double *theta, double *stepsize, double *inv_metric);
/* posterior summaries pooled over chains, the layout summarize_stan_hdf5 writes
 * (/root/reference/splotch/utils.py:234-257): mean and std (ddof=0) of log_lambda, lambda=exp(log_lambda)
 * and psi; each [G][N]; any pointer may be NULL */
int32_t csplotch_sampler_get_spot_summaries(csplotch_sampler *s, double *log_lambda_mean,
                                            double *log_lambda_std, double *lambda_mean, double *lambda_std,
                                            double *psi_mean, double *psi_std);

/* ---- CAR pre-compute on the GPU ---------------------------------------------------
 * Stands in for the CAR part of generate_dictionary (/root/reference/splotch/utils.py:735-765): per tissue the
 * eigenvalues of D^-1/2 W D^-1/2 (utils.py:749-755), for tissues with more than `exact_max_n` spots (reference: >= 10 000)
 * the chunked approximation eigs_square_chunking (utils.py:629-670: coordinate squares of rint(sqrt(chunk_spots)) units,
 * degrees inside the chunk), all eigenvalues concatenated and sorted (utils.py:759); D_sparse = degrees (utils.py:739);
 * U, V = 1-based CSR of the symmetric adjacency (sparse_to_csr_indptr, utils.py:592-614).  Identical tissues / chunks
 * are solved once (n_solves = dense eigenproblems actually solved).  Dense matrices are assembled on the device and
 * solved by cuSOLVER (loaded on first use).
 * N_spots[N_tissues]; W_sparse[W_n][2] 1-based global spot ids, every pair once, both ends in one tissue;
 * coords[N][2] non-negative integer x, y (only read for tissues beyond exact_max_n; may be NULL otherwise);
 * chunk_spots <= 0: 10 000; exact_max_n <= 0: 9 999 (the reference's rule), at most 40 000.
 * Outputs (any may be NULL): eig_values[N] ascending, U[N+1], V[2 W_n], D_sparse[N]. */
int32_t csplotch_car_precompute(int32_t N_tissues, const int32_t *N_spots, int32_t W_n, const int32_t *W_sparse,
                                const int32_t *coords, int32_t chunk_spots, int32_t exact_max_n, double *eig_values,
                                int32_t *U, int32_t *V, double *D_sparse, int32_t *n_solves);

/* ---- ingestion fast path ------------------------------------------------------
 * Stands in for read_rdump (/root/reference/splotch/utils.py:78-111) on the per-gene part of the
 * `data_<g>.R` files that bin/splotch_generate_input_files:329-346 writes: only `counts` (and, for
 * comp_splotch_stan_model, `beta_prior_mean` / `beta_prior_std`) differ between genes and to_rdump writes them
 * last, so only the tail of every file is read and parsed, n_threads files at a time (0 = all host cores).
 * paths[G]; counts[G][N]; beta_prior_mean/std [G][K] or both NULL.  Fails (CSPLOTCH_E_INVALID, message names the
 * file) on a missing file or key, a non-integer count, or a length other than N / K.  Host only, no device. */
int32_t csplotch_rdump_read_genes(const char *const *paths, int32_t G, int32_t N, int32_t K, int32_t *counts,
                                  double *beta_prior_mean, double *beta_prior_std, int32_t n_threads);

/* ---- output fast path (host only) ------------------------------------------------------
 * Stands in for CmdStan's CSV writer and the wrapper's concatenation of the chain files
 * (/root/reference/bin/splotch:131,138-140): file f = `header` + '\n', then n_rows rows, each the n_lead numbers
 * lead[f][r][:] (the sampler columns lp__ … energy__, or lp_approx__, lp__, path__ of the pathfinder method)
 * followed by the n_values numbers values[f][r][:] (csplotch_write_array), comma separated, no `#` lines — what
 * read_stan_csv (/root/reference/splotch/utils.py:139-232) parses.  Integral values print without a decimal point;
 * others in the shortest form that reads back to the same double (sig_figs = 0) or with sig_figs significant digits
 * (CmdStan's `sig_figs=`; its default is 6).  n_threads files at a time (0 = all host cores); each file is written
 * to `<path>.tmp` and renamed.  Fails (CSPLOTCH_E_INVALID, message names the file) when a file cannot be written. */
int32_t csplotch_csv_write(const char *const *paths, int32_t n_files, const char *header, const double *lead,
                           int32_t n_lead, const double *values, int32_t n_values, int64_t n_rows, int32_t sig_figs,
                           int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif
