/* ci_b200.h -- C ABI of the B200-native state-space engine for CausalImpact.
 *
 * This is the drop-in boundary for ONE path of google/tfp-causalimpact: the
 * model-fit + posterior-predictive path that the reference runs through
 * TensorFlow Probability.  Each entry point names the reference interface it
 * replaces (file:line relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 (CI_OK) or a negative ci_status; none throws or
 *     aborts; ci_last_error() gives the thread-local message of the last failure.
 *   - buffers are CALLER-OWNED, contiguous, row-major.  Pointers are HOST
 *     pointers unless the function name ends in _d, in which case they are
 *     DEVICE pointers on the context's device and `stream` is a cudaStream_t
 *     (passed as void*) the work is enqueued on; _d calls do not synchronise.
 *   - `dtype` (ci_problem.dtype) selects the element type of every `void*`
 *     buffer: 0 = float32, 1 = float64  (reference: DataOptions.dtype,
 *     causalimpact/causalimpact_lib.py:155-159).
 *   - theta layout, one row per chain / draw, all unconstrained:
 *       [ w_0 .. w_{p-1}, log sigma_obs^2, log sigma_level^2 (, log sigma_slope^2) ]
 *     dim = p + 1 + d,  d = 1 (local level) or 2 (local linear trend).
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef CI_B200_H_
#define CI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CI_B200_VERSION 100 /* major*10000 + minor*100 + patch */

typedef enum {
  CI_OK = 0,
  CI_ERR_INVALID = -1,     /* bad argument / shape / unsupported size        */
  CI_ERR_CUDA = -2,        /* CUDA runtime error (message has the details)   */
  CI_ERR_NO_DEVICE = -3,   /* no usable CUDA device                          */
  CI_ERR_STATE = -4,       /* call order (e.g. compute before ci_set_data)   */
  CI_ERR_UNSUPPORTED = -5  /* feature outside this path (seasonal, ...)      */
} ci_status;

enum { CI_MODEL_LOCAL_LEVEL = 0, CI_MODEL_LOCAL_LINEAR_TREND = 1 };
enum { CI_F32 = 0, CI_F64 = 1 };
/* ci_logprob*: which filter kernel runs */
enum { CI_VARIANT_SEQ = 0,   /* per-timestep recursion, state in registers   */
       CI_VARIANT_SCAN = 1   /* time-parallel associative scan (warp shuffles) */ };
/* ci_logprob* flags */
enum { CI_WITH_PRIOR = 1     /* add log prior + Jacobian: value is the HMC target */ };

/* The model + priors.  Replaces the TFP StructuralTimeSeries object built by
 * _build_default_gibbs_model (causalimpact/causalimpact_lib.py:398-500):
 *   obs_*   InverseGamma(conc, scale) on sigma_obs^2; obs_ub is the `upper_bound`
 *           attribute the reference hangs on that prior (lib.py:434-443)
 *   lvl_*   same for the local-level random walk        (lib.py:424-432)
 *   slope_* same for the trend slope (extension; the reference has no slope,
 *           lib.py:496) -- ignored when model == 0
 *   m0, P0  initial level ~ N(m0, P0)                   (lib.py:467-469)
 *   ub_on_scale  what the *_ub bounds (and ci_seasonal.drift_ub) limit.  The reference sets
 *           `prior.upper_bound = sd` on InverseGamma priors OVER VARIANCES (lib.py:432,
 *           442-443, 474) and TFP's Gibbs sampler clips the sampled variance:
 *           sample_parameters.sample_with_optional_upper_bound and the spike-and-slab
 *           sampler's observation_noise_variance_upper_bound both do min(variance, bound).
 *             0 (default)  bound on the VARIANCE: sigma^2 <= ub           (TFP semantics)
 *             1            bound on the SCALE:    sigma   <= ub           (round-1 behaviour)
 *           Gibbs kernels clip the draw, the HMC target truncates the prior at the bound.
 */
typedef struct {
  int32_t model;    /* CI_MODEL_*                                            */
  int32_t dtype;    /* CI_F32 / CI_F64                                       */
  int32_t T;        /* length of the extended series: pre + after-pre
                       (lib.py:548-562)                                      */
  int32_t p;        /* covariates + intercept (data.py:129-135); 0 = none    */
  double obs_conc, obs_scale, obs_ub;
  double lvl_conc, lvl_scale, lvl_ub;
  double slope_conc, slope_scale, slope_ub;
  double m0, P0;
  double m0_slope, P0_slope;
  int32_t ub_on_scale;  /* 0: *_ub limit variances (TFP), 1: they limit scales         */
  int32_t reserved;
} ci_problem;

/* HMC driver options (replaces num_results / num_warmup_steps handed to
 * gibbs_sampler.fit_with_gibbs_sampling, lib.py:365-388). */
typedef struct {
  int32_t n_warmup;       /* adaptation iterations (discarded)               */
  int32_t n_results;      /* kept iterations per chain                       */
  int32_t max_leapfrog;   /* integration length is U{1..max_leapfrog} per
                             iteration (same for every chain)                */
  int32_t adapt_mass;     /* 1 = adapt a diagonal mass matrix in warm-up     */
  double init_step;       /* initial leapfrog step size                      */
  double target_accept;   /* dual-averaging target (0.8)                     */
} ci_hmc_opts;

typedef struct {
  float accept_rate;      /* mean acceptance probability over kept iterations */
  float step_size;        /* final adapted step size                          */
  int32_t n_divergent;    /* kept iterations with non-finite / exploding H    */
  int32_t n_leapfrog;     /* total gradient evaluations of this chain         */
} ci_hmc_stats;

typedef struct ci_ctx ci_ctx;

/* ---- library / context ------------------------------------------------- */
int ci_version(void);
const char* ci_last_error(void);
int ci_device_count(void);                 /* < 0 on error                   */
int ci_ctx_create(int device, ci_ctx** out);
int ci_ctx_destroy(ci_ctx* ctx);
/* Number of kernel launches this context has enqueued so far. */
int64_t ci_launch_count(const ci_ctx* ctx);

/* Upload one problem: replaces the tensors _train_causalimpact_sts hands to the
 * sampler (lib.py:545-581).
 *   y     [T]    standardized outcome, NaN = missing (pre-period NaNs and the
 *                whole after-pre range, lib.py:548-562)
 *   X     [T,p]  standardized design matrix incl. intercept (NULL when p == 0)
 *   Omega [p,p]  slab precision (lib.py:451-453)          (NULL when p == 0)
 */
int ci_set_data(ci_ctx* ctx, const ci_problem* prob, const void* y,
                const void* X, const void* Omega);

/* ---- K1/K2/K3: Kalman log-prob (+ gradient) for a batch of chains -------
 * Replaces tfd.LinearGaussianStateSpaceModel(...).log_prob as it would be
 * evaluated under tfp.sts.fit_with_hmc (north_star); the reference call site
 * being replaced is the sampler loop at lib.py:365-388.
 *   theta [C,dim] in;  value [C] out;  grad [C,dim] out (may be NULL).
 */
int ci_logprob(ci_ctx* ctx, const void* theta, int n_chains, void* value,
               int variant, int flags);
int ci_logprob_grad(ci_ctx* ctx, const void* theta, int n_chains, void* value,
                    void* grad, int variant, int flags);
int ci_logprob_grad_d(ci_ctx* ctx, const void* theta_d, int n_chains,
                      void* value_d, void* grad_d, int variant, int flags,
                      void* stream);

/* ---- K6: batched-chain HMC over the filter kernel ------------------------
 * Replaces gibbs_sampler.fit_with_gibbs_sampling (lib.py:365-388).
 *   theta0 [C,dim]  initial points
 *   draws  [n_results, C, dim] out;  stats [C] out.
 * RNG is Philox4x32-10 keyed by (seed, chain_id0 + c): results do not depend
 * on how chains are split across devices (causalimpact_lib_test.py:493-502).
 */
int ci_hmc_run(ci_ctx* ctx, const ci_hmc_opts* opts, uint64_t seed,
               uint64_t chain_id0, const void* theta0, int n_chains,
               void* draws, ci_hmc_stats* stats);
int ci_hmc_run_d(ci_ctx* ctx, const ci_hmc_opts* opts, uint64_t seed,
                 uint64_t chain_id0, const void* theta0_d, int n_chains,
                 void* draws_d, ci_hmc_stats* stats_d, void* stream);

/* ---- Gibbs sampler with spike-and-slab regression (the reference's sampler) -
 * Replaces gibbs_sampler.fit_with_gibbs_sampling exactly as the reference calls
 * it (lib.py:365-388): one sweep = spike-and-slab (sigma_obs^2, weights) draw,
 * FFBS level draw, sigma_level^2 draw; local level model only.  The initial
 * state is the reference's (lib.py:566-581).
 *   draws [n_results, C, dim]  (w, log sigma_obs^2, log sigma_level^2) per sweep
 *   level [n_results, C, T]    level path of each kept sweep        (may be NULL)
 *   traj  [n_results, C, T]    level + X.w + sigma_obs N(0,1)       (may be NULL)
 *   incl  [C, p] float         inclusion frequency per feature      (may be NULL)
 * sparse = 1: inclusion probability nonzero_prob (lib.py:449-450: min(1, 3/p));
 * sparse = 0: every feature always included.
 */
typedef struct {
  int32_t n_warmup;
  int32_t n_results;
  int32_t sparse;
  int32_t chain_major;    /* 0: outputs [n_results, C, .] (sweep-major, as documented above)
                             1: outputs [C, n_results, .] -- every chain's draws contiguous,
                                the layout the multi-GPU all-gather wants             */
  double nonzero_prob;
  int32_t ssvs_order;     /* order in which a sweep visits the features' inclusion indicators:
                             0 = a fresh random permutation per sweep (Philox; TFP's sampler
                                 permutes the features), 1 = index order 0..p-1             */
  int32_t reserved;
  uint64_t series_stride; /* batched runs only: series s uses the global chain ids
                             chain_id0 + s * series_stride + c.  0 = every series consumes the
                             SAME random streams (bit-identical to a single-series run, but the
                             Monte-Carlo errors of the series are then perfectly correlated);
                             >= n_chains gives every series its own streams                */
} ci_gibbs_opts;

int ci_gibbs_run(ci_ctx* ctx, const ci_gibbs_opts* opts, uint64_t seed,
                 uint64_t chain_id0, int n_chains, void* draws, void* level,
                 void* traj, float* incl);
int ci_gibbs_run_d(ci_ctx* ctx, const ci_gibbs_opts* opts, uint64_t seed,
                   uint64_t chain_id0, int n_chains, void* draws_d,
                   void* level_d, void* traj_d, float* incl_d, void* stream);

/* ---- K4: simulation smoother + one-step predictive draw ------------------
 * Replaces _resample_latents' LGSSM posterior_sample (inside the sampler,
 * lib.py:365-388) and _get_posterior_means_and_trajectories (lib.py:609-632).
 *   theta_draws [S,dim] in
 *   level [S,T] out  (may be NULL) posterior sample of the level path
 *   traj  [S,T] out  level + X.w + sigma_obs * N(0,1)      (lib.py:629-631)
 *   mean  [T]   out  average over draws of level + X.w     (lib.py:627)
 * RNG keyed by (seed, draw_id0 + s).  With model = CI_MODEL_LOCAL_LINEAR_TREND (extension,
 * BASELINE configs[2]) the state is (level, slope): the d = 2 simulation smoother runs, `level`
 * receives the level component, theta rows have p + 3 entries.
 */
int ci_posterior_predict(ci_ctx* ctx, const void* theta_draws, int S,
                         uint64_t seed, uint64_t draw_id0, void* level,
                         void* traj, void* mean);
int ci_posterior_predict_d(ci_ctx* ctx, const void* theta_draws_d, int S,
                           uint64_t seed, uint64_t draw_id0, void* level_d,
                           void* traj_d, void* mean_d, void* stream);

/* ---- K5: per-time quantiles across draws ---------------------------------
 * Replaces posterior_processing.calculate_trajectory_quantiles
 * (causalimpact/posterior_processing.py:25-60): pandas
 * DataFrame.quantile(axis=1), i.e. linear interpolation at q*(S-1), NaNs
 * skipped.   a [S,T] (draw-major, as _get_posterior_means_and_trajectories
 * returns it) -> out [T,nq].  dtype: CI_F32 / CI_F64.
 */
int ci_row_quantiles(ci_ctx* ctx, const void* a, int S, int T, int dtype,
                     const double* q, int nq, void* out);
int ci_row_quantiles_d(ci_ctx* ctx, const void* a_d, int S, int T, int dtype,
                       const double* q, int nq, void* out_d, void* stream);

/* ---- seasonal components --------------------------------
 * Replaces the tfp.sts.Seasonal components the reference adds to its Gibbs model for
 * every ModelOptions.seasons entry (lib.py:471-489: allow_drift, mean effect constrained
 * to zero, drift variance ~ InverseGamma(drift_conc, drift_scale) bounded by drift_ub,
 * initial effects ~ N(0, init_sd)) and the sampler's joint (level, seasonal) latent draw
 * + drift-scale draws (lib.py:365-388; initial drift scale 0.01 sd, :573-574).
 *   active [K][T] uint8 HOST: season index (0 .. num_seasons-1) active at step t
 *   ends   [K][T] uint8 HOST: 1 when that season is over after step t
 * (the host derives both from Seasons.num_steps_per_season, lib.py:162-180).
 * Limits: K <= CI_MAX_SEASONAL, 1 + sum(num_seasons) <= 192 (up to six state elements per warp
 * lane: week-of-year and hour-of-week calendars fit) and the d x d state covariance of a chain must
 * fit in shared memory (float64 stops near d = 160); beyond that CI_ERR_UNSUPPORTED.  Call after ci_set_data; n_components = 0
 * (or seas == NULL) removes the components; ci_set_data also removes them.
 */
#define CI_MAX_SEASONAL 7
typedef struct {
  int32_t n_components;
  int32_t num_seasons[CI_MAX_SEASONAL];
  const uint8_t* active;
  const uint8_t* ends;
  double init_sd;
  double drift_conc, drift_scale, drift_ub;
} ci_seasonal;

int ci_set_seasonal(ci_ctx* ctx, const ci_seasonal* seas);

/* Gibbs run with the seasonal components of ci_set_seasonal.  Outputs as ci_gibbs_run plus
 *   latent   [rows, T]     level + sum of the seasonal contributions       (may be NULL)
 *   seasonal [rows, T, K]  contribution of each component at every step    (may be NULL)
 *   drift    [rows, K]     log drift VARIANCE of each component            (may be NULL)
 * (rows ordered like draws: opts->chain_major).  traj = latent + X.w + sigma_obs N(0,1). */
int ci_gibbs_seasonal_run(ci_ctx* ctx, const ci_gibbs_opts* opts, uint64_t seed,
                          uint64_t chain_id0, int n_chains, void* draws, void* level,
                          void* traj, float* incl, void* latent, void* seasonal, void* drift);
int ci_gibbs_seasonal_run_d(ci_ctx* ctx, const ci_gibbs_opts* opts, uint64_t seed,
                            uint64_t chain_id0, int n_chains, void* draws_d, void* level_d,
                            void* traj_d, float* incl_d, void* latent_d, void* seasonal_d,
                            void* drift_d, void* stream);

/* ---- batches of independent series -------------------------
 * The reference fits ONE series per fit_causalimpact call (single process, single chain;
 * a user with thousands of geographies loops).  A batch holds N series that share their
 * shape (T, p, dtype): data.py:77-137 output for each of them, uploaded together.
 *   probs [N]      the model + priors of each series (m0, P0 and the prior scales differ)
 *   y     [N,T]    X [N,T,p]    Omega [N,p,p]
 * ci_gibbs_run_batch_d runs n_chains chains of EVERY series in one launch (grid.y = series).
 * Every series uses the chain ids chain_id0 .. chain_id0 + n_chains - 1, so its draws are
 * bit-identical to ci_gibbs_run on that series alone; outputs are series-major:
 *   draws [N, rows, dim]   level / traj [N, rows, T]   incl [N, n_chains, p]
 * with rows = n_chains * n_results ordered as opts->chain_major says.
 * ci_batch_select makes one series of the batch the context's current problem, so that
 * every single-series entry point (ci_predictive_mean_d, ci_logprob*, ci_hmc_run*,
 * ci_posterior_predict*, ci_gibbs_run*) works on it without another upload.
 */
int ci_set_data_batch(ci_ctx* ctx, const ci_problem* probs, int n_series, const void* y,
                      const void* X, const void* Omega);
int ci_batch_select(ci_ctx* ctx, int series);
int ci_gibbs_run_batch_d(ci_ctx* ctx, const ci_gibbs_opts* opts, uint64_t seed,
                         uint64_t chain_id0, int n_chains, void* draws_d, void* level_d,
                         void* traj_d, float* incl_d, void* stream);
/* The same with seasonal components.  ci_set_seasonal_batch (after ci_set_data_batch) gives the
 * season calendar of the WHOLE panel (every series has the same T) and, per series, the priors
 * that scale with the series' own outcome sd (lib.py:472-489): init_sd [N], drift_scale [N],
 * drift_ub [N] (HOST arrays; the scalar fields of `seas` are ignored by the batched kernel).
 * Outputs as ci_gibbs_seasonal_run, series-major; bit-identical per series to the single run. */
int ci_set_seasonal_batch(ci_ctx* ctx, const ci_seasonal* seas, const double* init_sd,
                          const double* drift_scale, const double* drift_ub);
int ci_gibbs_seasonal_run_batch_d(ci_ctx* ctx, const ci_gibbs_opts* opts, uint64_t seed,
                                  uint64_t chain_id0, int n_chains, void* draws_d,
                                  void* level_d, void* traj_d, float* incl_d, void* latent_d,
                                  void* seasonal_d, void* drift_d, void* stream);

/* ---- panel data preparation on the device ----------------------------------
 * Replaces, for N series at once, the per-series pandas work of CausalImpactData (data.py:77-137:
 * split into pre / after-pre rows, nan-aware standardisation with the pre-period statistics of
 * standardize.py:42-64, intercept column, masked outcome), the priors / initial state of
 * lib.py:398-500, 563-572 and the host half of ci_set_data_batch (tiles, X'X / X'y over observed
 * rows, slab precision over the full design): ONE kernel, one CTA per series, float64 arithmetic.
 *   values [N, T_total, n_cols]  HOST float64, column 0 = outcome (NaN = missing), the others are
 *                                covariates (no NaN); rows [row0, row0 + n_pre) are the
 *                                pre-period, rows [row0, T_total) the modelled span
 *   stats  [N, CI_PANEL_STATS]   HOST out: y_scale, y_offset (original = standardized * scale +
 *                                offset), outcome_sd, n_obs, y'y, m0, error code, reserved
 * Afterwards the context holds the batch exactly as after ci_set_data_batch (T = T_total - row0,
 * p = n_cols if n_cols > 1 else 0): every batched and -- through ci_batch_select -- every
 * single-series entry point works on it.  Errors carry the reference's messages (data.py:140-190).
 */
#define CI_PANEL_STATS 8
typedef struct {
  int32_t n_series, T_total, n_cols;
  int32_t row0, n_pre;
  int32_t standardize;    /* DataOptions.standardize_data                                   */
  int32_t dtype;          /* element type of the ENGINE buffers (tiles, draws): CI_F32 / F64 */
  int32_t ub_on_scale;    /* see ci_problem                                                  */
  double prior_level_sd;  /* ModelOptions.prior_level_sd (lib.py:202)                        */
} ci_panel_args;
int ci_set_panel(ci_ctx* ctx, const ci_panel_args* args, const double* values, double* stats);

/* Mean of the predictive mixture alone (lib.py:627): mean_t = avg_s level[s,t] + x_t . avg_s w_s,
 * for draws whose level paths are already on the device (the Gibbs kernel's output).
 * Deterministic: fixed summation order, float64 accumulation. */
int ci_predictive_mean_d(ci_ctx* ctx, const void* theta_draws_d, const void* level_d,
                         int S, void* mean_d, void* stream);

/* ---- impact series + summary ----------------------------
 * Replaces the O(S*T) part of _compute_impact (causalimpact_lib.py:635-705):
 * un-standardising (posterior_processing.py:88-89 -> standardize.py:60-64),
 * point / cumulative effect paths (lib.py:793-837), the three per-time quantile
 * passes (posterior_processing.py:25-60 at lib.py:760, 886, 888) and the
 * post-period summary statistics (lib.py:934-1093).  Float64 arithmetic.
 *   traj     [S,T]  predictive draws on the standardized scale (args.dtype)
 *   mean     [T]    predictive mean on the standardized scale  (args.dtype)
 *   observed [T]    float64 HOST array: the outcome on the ORIGINAL scale over the
 *                   modelled span; NaN where it does not count (gap between the
 *                   periods, after the post-period, missing)
 *   period   [T]    uint8 HOST array: 0 = before the post-period starts,
 *                   1 = inside the post-period, 2 = after it; must be non-decreasing
 *   series   [T,9]  float64 out: posterior_mean, posterior_lower, posterior_upper,
 *                   point_effects_{mean,lower,upper}, cumulative_effects_{mean,lower,upper}
 *                   (rows not yet NaN-ed by the gap / after-post rules of lib.py:899-915,
 *                   which are O(T) host work)
 *   summary  [CI_IMPACT_SUMMARY_LEN] float64 out:
 *     [0..9]   (lower, upper) quantiles over draws of: post-period mean prediction, summed
 *              prediction, mean effect, summed effect, relative effect
 *     [10..14] their sample standard deviations (ddof = 1)
 *     [15]     mean relative effect          [16],[17]  #draws with obs_sum <= / >= summed prediction
 *     [18],[19] post-period mean / sum of the predictive mean
 * Columns of up to ~25 000 draws are selected in shared memory, longer ones from global memory.
 */
#define CI_IMPACT_SERIES_COLS 9
#define CI_IMPACT_SUMMARY_LEN 20
typedef struct {
  int32_t S, T;
  int32_t dtype;          /* CI_F32 / CI_F64: element type of traj and mean          */
  int32_t reserved;
  double scale, offset;   /* original = standardized * scale + offset (1, 0 = none)  */
  double q_lo, q_hi;      /* alpha/2 and 1 - alpha/2                                 */
  double obs_sum;         /* sum of the observed outcome over the post-period        */
} ci_impact_args;

int ci_impact(ci_ctx* ctx, const ci_impact_args* args, const void* traj, const void* mean,
              const double* observed, const uint8_t* period, double* series,
              double* summary);
/* traj_d / mean_d / series_d / summary_d are device pointers; observed / period stay HOST. */
int ci_impact_d(ci_ctx* ctx, const ci_impact_args* args, const void* traj_d,
                const void* mean_d, const double* observed, const uint8_t* period,
                double* series_d, double* summary_d, void* stream);

/* The two halves of ci_impact_d, for draws sharded over GPUs.  The quantile columns
 * need ALL draws of a time step, so a sharded fit exchanges the TRANSPOSED paths by time block
 * (an all-to-all of S*T/world elements per rank) instead of gathering every draw on every GPU:
 *   1. ci_impact_rows_d on each rank's own draws (args->S = its draw count):
 *        trT [T,S] (args.dtype), cumT [T - t_c0, S] float64, stats [5,S] float64 -- the paths and
 *        the per-draw post-period statistics, draws contiguous per time step.  With mean_d (the
 *        predictive mean over ALL draws; one rank passes it, the others NULL) the mean-derived
 *        series columns 0, 3, 6 and summary[18..19] are written too.
 *   2. the caller's collective moves rows [t_begin, t_begin + t_count) of every rank's trT (and
 *        [c_begin, ...) of cumT) to the rank that owns that time block, and stats to one rank,
 *        and lays the received blocks side by side as [t_count, S_total].
 *   3. ci_impact_cols_d (args->S = S_total): the quantile columns of this rank's time block; with
 *        stats_d (all draws, [5,S_total]) also summary[0..17].  Entries this call does not own
 *        are left untouched: zero series / summary first and sum the buffers over ranks.
 * t_c0 = the first step with period != 0.  Results equal ci_impact_d on the gathered draws bit
 * for bit (the select is exact); reference: the same lines as ci_impact. */
int ci_impact_rows_d(ci_ctx* ctx, const ci_impact_args* args, const void* traj_d,
                     const void* mean_d, const double* observed, const uint8_t* period,
                     void* trT_d, double* cumT_d, double* stats_d, double* series_d,
                     double* summary_d, void* stream);
int ci_impact_cols_d(ci_ctx* ctx, const ci_impact_args* args, const void* trT_d, int t_begin,
                     int t_count, const double* cumT_d, int c_begin, int c_count,
                     const double* stats_d, const double* observed, const uint8_t* period,
                     double* series_d, double* summary_d, void* stream);

/* Batched (grid.y = series) versions of ci_predictive_mean_d and ci_impact_d for the batch in the
 * context: theta [N,S,dim], level [N,S,T] -> mean [N,T];
 * traj [N,S,T], mean [N,T], observed HOST [N,T], period HOST [T] (shared), per-series scale /
 * offset / obs_sum HOST [N] (args->scale / offset / obs_sum are ignored)
 * -> series [N,T,9], summary [N,CI_IMPACT_SUMMARY_LEN].  Two launches for the whole panel. */
int ci_predictive_mean_batch_d(ci_ctx* ctx, const void* theta_d, const void* level_d, int S,
                               void* mean_d, void* stream);
int ci_impact_batch_d(ci_ctx* ctx, const ci_impact_args* args, int n_series, const double* scale,
                      const double* offset, const double* obs_sum, const void* traj_d,
                      const void* mean_d, const double* observed, const uint8_t* period,
                      double* series_d, double* summary_d, void* stream);

/* ---- the one collective of the path -----------------------------------------
 * Chains and posterior draws are sharded over GPUs by GLOBAL id (chain_id0 / draw_id0 above), one
 * process per GPU; the only exchange is an all-gather of the per-draw result rows at the end of a
 * fit.  ci_comm wraps an NCCL communicator (resolved at run time with dlopen("libnccl.so.2"); no
 * link-time dependency, shared with a torch-loaded NCCL when there is one).  The reference has no
 * counterpart: it is single process, single device (lib.py:342-345).
 *   ci_comm_get_unique_id   rank 0 creates the 128-byte ncclUniqueId and hands it to the other
 *                           ranks out of band (file, socket, MPI, torch.distributed store, ...)
 *   ci_comm_create          every rank, with the same id; collective (returns when all joined)
 *   ci_allgather            recv_d [nranks * bytes_per_rank] <- every rank's send_d
 *                           [bytes_per_rank], in rank order; enqueued on `stream`, no host sync
 */
#define CI_COMM_ID_BYTES 128
typedef struct ci_comm ci_comm;
int ci_comm_get_unique_id(uint8_t* id /* [CI_COMM_ID_BYTES] out */);
int ci_comm_create(ci_ctx* ctx, const uint8_t* id /* [CI_COMM_ID_BYTES] */, int rank, int nranks,
                   ci_comm** out);
int ci_allgather(ci_comm* comm, const void* send_d, void* recv_d, size_t bytes_per_rank,
                 void* stream);
int ci_comm_destroy(ci_comm* comm);

/* The impact stage of a fit whose draws are sharded over the ranks of `comm`, in ONE call: steps
 * 1-3 of ci_impact_rows_d / ci_impact_cols_d above with the exchange inside.  Everything is
 * enqueued on `stream` (kernels and NCCL alike; no host synchronisation except when the exchange
 * windows first grow):
 *   k_impact_rows on the rank's own draws, storing every transposed tile straight into the window
 *   of the rank that owns its time block -- the ranks' windows are mapped into each other through
 *   CUDA IPC (peer memory over NVLink / NVSwitch), so compute and transfer are one kernel; the
 *   per-draw statistics go to rank 0's window -> an ncclAllGather of the ranks' predictive-mean
 *   parts, which is also the barrier of the exchange -> k_impact_jobs on T/nranks time steps over
 *   all draws, read as they arrived -> ncclAllReduce of the [T*9 + 20] result (each entry written
 *   by one rank).  Where peer mapping is unavailable (agreed collectively; CI_B200_NO_PEER=1
 *   forces it) the tiles go to local arrays and ONE grouped ncclSend/ncclRecv exchange moves the
 *   time blocks; results are the same.  At most 16 ranks (one NVSwitch domain).
 * Time blocks: rank r owns steps start..start+count-1 with base = T / nranks, extra = T % nranks,
 * start = r*base + min(r, extra), count = base + (r < extra); likewise for the T - t_c0
 * cumulative columns.
 *   args->S       this rank's draw count (may be 0 on ranks other than 0)
 *   counts        [nranks] HOST int32: the draw count of every rank (global draw order = rank order)
 *   traj_d        [args->S, T] this rank's predictive draws
 *   mean_part_d   [T] the predictive mean over THIS rank's draws (ci_predictive_mean_d); ignored
 *                 where counts[rank] == 0
 *   mean_d        [T] out: the predictive mean over all draws (draw-count-weighted float64 sum of
 *                 the parts in rank order)
 *   out_d         [T*9 + CI_IMPACT_SUMMARY_LEN] float64 out, on every rank: series then summary of
 *                 ci_impact; the quantile entries are bit-identical to ci_impact_d on the gathered
 *                 draws, the mean-derived ones agree to the rounding of the parts to args->dtype.
 * comm with nranks == 1 is allowed (the window is the rank's own). */
int ci_impact_sharded_d(ci_ctx* ctx, ci_comm* comm, const ci_impact_args* args,
                        const int32_t* counts, const void* traj_d, const void* mean_part_d,
                        const double* observed, const uint8_t* period, void* mean_d,
                        double* out_d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CI_B200_H_ */
