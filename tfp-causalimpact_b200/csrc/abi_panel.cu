// abi_panel.cu -- SURVEY section 8 row f4 on the device: data preparation of a whole PANEL of
// independent series in one kernel, and the batched (grid.y = series) predictive-mean and impact
// entry points that follow the batched sampler.
//
// ci_set_panel replaces, for N series at once, what the reference does per series in pandas before
// the sampler sees anything (relative to /root/reference):
//   causalimpact/data.py:77-137        split into pre / after-pre, nan-aware standardisation with
//                                      the pre-period statistics, intercept column, masked outcome
//   causalimpact/standardize.py:42-64  mean / std (ddof = 1) per column, (x - mean) / std
//   causalimpact/causalimpact_lib.py:398-500, 563-572  priors and initial state from outcome_sd
// and what ci_set_data_batch does on the host (tile layout, Gram matrix / X'y over observed rows,
// slab precision over the full-length design).  The raw float64 panel crosses PCIe once; tiles,
// priors, sufficient statistics and the per-series kernel descriptors are written where the
// sampler reads them.  One CTA per series, float64 arithmetic, fixed reduction order.
#include "ci_host.cuh"
#include "ci_gibbs.cuh"
#include "ci_impact.cuh"

namespace {

using namespace ci;

constexpr int PREP_THREADS = 256;

struct PanelDev {
  int N, T_total, ncol, row0, n_pre, Tm, p, NB, ld, standardize, ub_on_scale;
  double prior_level_sd;
  size_t tile_stride, om_stride;      // elements of R per series
};

// fixed-order block sum (deterministic): warp shuffles, then thread 0 adds the warp totals
__device__ __forceinline__ double prep_block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < PREP_THREADS / 32; ++w) t += red[w];
    red[32] = t;
  }
  __syncthreads();
  return red[32];
}

template <typename R>
__global__ void __launch_bounds__(PREP_THREADS)
k_panel_prep(const double* __restrict__ vals, PanelDev a, R* __restrict__ tiles, R* __restrict__ omega,
             R* __restrict__ gram, R* __restrict__ xty, BatchDev<R>* __restrict__ dev,
             double* __restrict__ stats) {
  __shared__ double red[33];
  __shared__ double col_mean[MAX_DIM], col_sd[MAX_DIM];
  __shared__ int s_first;
  const int s = blockIdx.x, tid = threadIdx.x;
  const int ncol = a.ncol, p = a.p, k = ncol - 1, Tm = a.Tm;
  const double* v0 = vals + (size_t)s * a.T_total * ncol;
  const double* vm = v0 + (size_t)a.row0 * ncol;                  // first modelled row
  double err = 0.0;
  // ---- validation over the WHOLE input (data.py:140-190) ----
  {
    double cnt = 0.0, sum = 0.0, bad = 0.0;
    for (int t = tid; t < a.T_total; t += PREP_THREADS) {
      const double y = v0[(size_t)t * ncol];
      if (y == y) { cnt += 1.0; sum += y; }
      for (int j = 1; j < ncol; ++j) { const double x = v0[(size_t)t * ncol + j]; if (!(x == x)) bad += 1.0; }
    }
    const double n_all = prep_block_sum(cnt, red), s_all = prep_block_sum(sum, red);
    const double n_bad = prep_block_sum(bad, red);
    const double m_all = n_all > 0 ? s_all / n_all : 0.0;
    double ss = 0.0;
    for (int t = tid; t < a.T_total; t += PREP_THREADS) {
      const double y = v0[(size_t)t * ncol];
      if (y == y) ss += (y - m_all) * (y - m_all);
    }
    const double ss_all = prep_block_sum(ss, red);
    if (n_all < 3.0) err = 2.0;                 // "Input data must have at least 3 observations."
    else if (!(ss_all > 0.0)) err = 1.0;        // "Input response cannot be constant."
    else if (n_bad > 0.0) err = 3.0;            // "Input data cannot have any missing values."
  }
  // ---- standardize.py:42-47 on the pre-period rows: nan-aware mean, std (ddof = 1) ----
  for (int c = 0; c < ncol; ++c) {
    double cnt = 0.0, sum = 0.0;
    for (int t = tid; t < a.n_pre; t += PREP_THREADS) {
      const double x = vm[(size_t)t * ncol + c];
      if (x == x) { cnt += 1.0; sum += x; }
    }
    const double n = prep_block_sum(cnt, red), tot = prep_block_sum(sum, red);
    const double m = n > 0 ? tot / n : 0.0;
    double ss = 0.0;
    for (int t = tid; t < a.n_pre; t += PREP_THREADS) {
      const double x = vm[(size_t)t * ncol + c];
      if (x == x) ss += (x - m) * (x - m);
    }
    const double sst = prep_block_sum(ss, red);
    if (tid == 0) { col_mean[c] = m; col_sd[c] = n > 1 ? sqrt(sst / (n - 1.0)) : CUDART_NAN; }
  }
  __syncthreads();
  double y_scale = 1.0, y_offset = 0.0;
  if (a.standardize) { y_scale = col_sd[0]; y_offset = col_mean[0]; }
  auto scaled = [&](double x, int c) -> double {       // data.py:114-119: columns with std 0 pass through
    if (!a.standardize) return x;
    const double sd = col_sd[c];
    return sd > 0.0 ? (x - col_mean[c]) / sd : x;
  };
  // ---- tiles: [x_1 .. x_k, 1 | y] rounded to the engine's dtype; y = NaN where masked ----
  R* tl0 = tiles + (size_t)s * a.tile_stride;
  const size_t te = (size_t)tile_elems(p);
  for (size_t i = tid; i < (size_t)a.NB * te; i += PREP_THREADS) tl0[i] = (R)0;
  __syncthreads();
  for (int t = tid; t < a.NB * TB; t += PREP_THREADS) {
    const int b = t / TB, tl = t - b * TB;
    R* row = tl0 + (size_t)b * te + tile_off(tl, a.ld);
    if (t < Tm) {
      const double* src = vm + (size_t)t * ncol;
      for (int j = 0; j < k; ++j) row[j] = (R)scaled(src[1 + j], 1 + j);
      if (p > 0) row[k] = (R)1;                                  // intercept (data.py:129-135)
      row[p] = t < a.n_pre ? (R)scaled(src[0], 0) : Num<R>::nan();   // lib.py:548-562
    } else {
      row[p] = Num<R>::nan();                                    // padded step == masked step
    }
  }
  __syncthreads();
  auto y_at = [&](int t) -> double {
    const int b = t / TB, tl = t - b * TB;
    return (double)tl0[(size_t)b * te + tile_off(tl, a.ld) + p];
  };
  auto x_at = [&](int t, int j) -> double {
    const int b = t / TB, tl = t - b * TB;
    return (double)tl0[(size_t)b * te + tile_off(tl, a.ld) + j];
  };
  // ---- outcome_sd = nanstd(pre y, ddof = 1) of the ROUNDED series (lib.py:563-564), y'y, n_obs,
  //      first observed value (the prior mean of the initial level, lib.py:467-469) ----
  double cnt = 0.0, sum = 0.0, sq = 0.0;
  int first = Tm;
  for (int t = tid; t < a.n_pre; t += PREP_THREADS) {
    const double y = y_at(t);
    if (y == y) { cnt += 1.0; sum += y; sq += y * y; first = min(first, t); }
  }
  const double n_obs = prep_block_sum(cnt, red), ysum = prep_block_sum(sum, red);
  const double yty = prep_block_sum(sq, red);
  if (tid == 0) s_first = Tm;
  __syncthreads();
  atomicMin(&s_first, first);
  __syncthreads();
  const double ymean = n_obs > 0 ? ysum / n_obs : 0.0;
  double ss = 0.0;
  for (int t = tid; t < a.n_pre; t += PREP_THREADS) {
    const double y = y_at(t);
    if (y == y) ss += (y - ymean) * (y - ymean);
  }
  const double yss = prep_block_sum(ss, red);
  const double sd = n_obs > 1 ? (double)(R)sqrt(yss / (n_obs - 1.0)) : CUDART_NAN;
  if (n_obs < 2.0 && err == 0.0) err = 4.0;
  // ---- X'X and X'y over observed rows, X'X over ALL modelled rows (slab precision) ----
  R* gr = gram + (size_t)s * p * p;
  R* xt = xty + (size_t)s * (p > 0 ? p : 1);
  R* om = omega + (size_t)s * a.om_stride;
  const int npairs = p * (p + 1) / 2;
  auto put_pair = [&](int i, int j, double g_obs, double g_all) {
    gr[i * p + j] = (R)g_obs; gr[j * p + i] = (R)g_obs;
    // lib.py:451-453: 0.01 * set_diag(0.5 X'X, diag(X'X)) / T
    const double o = 0.01 * (i == j ? g_all : 0.5 * g_all) / (double)Tm;
    om[i * p + j] = (R)o; om[j * p + i] = (R)o;
  };
  if (npairs <= 64) {
    // few covariates (the panel use case): the whole CTA shares the rows of one pair
    for (int i = 0; i < p; ++i) {
      for (int j = 0; j <= i; ++j) {
        double g_obs = 0.0, g_all = 0.0;
        for (int t = tid; t < Tm; t += PREP_THREADS) {
          const double prod = x_at(t, i) * x_at(t, j), y = y_at(t);
          g_all += prod;
          if (y == y) g_obs += prod;
        }
        g_obs = prep_block_sum(g_obs, red); g_all = prep_block_sum(g_all, red);
        if (tid == 0) put_pair(i, j, g_obs, g_all);
      }
      double b = 0.0;
      for (int t = tid; t < a.n_pre; t += PREP_THREADS) {
        const double y = y_at(t);
        if (y == y) b += x_at(t, i) * y;
      }
      b = prep_block_sum(b, red);
      if (tid == 0) xt[i] = (R)b;
    }
  } else {
    // many covariates: one thread per (i >= j) pair walks the rows
    for (int pair = tid; pair < npairs; pair += PREP_THREADS) {
      int i = 0, rem = pair;
      while (rem > i) { rem -= i + 1; ++i; }
      const int j = rem;
      double g_obs = 0.0, g_all = 0.0;
      for (int t = 0; t < Tm; ++t) {
        const double prod = x_at(t, i) * x_at(t, j), y = y_at(t);
        g_all += prod;
        if (y == y) g_obs += prod;
      }
      put_pair(i, j, g_obs, g_all);
    }
    for (int i = tid; i < p; i += PREP_THREADS) {
      double b = 0.0;
      for (int t = 0; t < a.n_pre; ++t) {
        const double y = y_at(t);
        if (y == y) b += x_at(t, i) * y;
      }
      xt[i] = (R)b;
    }
  }
  __syncthreads();
  // ---- the model + priors of this series (lib.py:398-500) and its kernel descriptor ----
  if (tid == 0) {
    const double level0 = a.prior_level_sd * sd;
    const double m0 = s_first < Tm ? y_at(s_first) : 0.0;
    auto ubv = [&](double ub) -> double { return a.ub_on_scale ? ub * ub : ub; };
    BatchDev<R> d;
    d.pr.tiles = tl0; d.pr.omega = om;
    d.pr.T = Tm; d.pr.p = p; d.pr.ld = a.ld; d.pr.NB = a.NB; d.pr.dim = p + 2; d.pr.model = 0;
    d.pr.m0 = (R)m0; d.pr.P0 = (R)(sd * sd);
    d.pr.obs_conc = (R)(p > 0 ? 25.0 : 0.005);
    d.pr.obs_scale = (R)((p > 0 ? 5.0 : 0.005) * sd * sd);
    d.pr.obs_ub = (R)ubv(1.2 * sd);
    d.pr.lvl_conc = (R)16.0; d.pr.lvl_scale = (R)(16.0 * level0 * level0); d.pr.lvl_ub = (R)ubv(sd);
    d.gd.gram = gr; d.gd.xty0 = xt; d.gd.yty0 = (R)yty;
    d.n_obs = (int)n_obs;
    dev[s] = d;
    double* st = stats + (size_t)s * CI_PANEL_STATS;
    st[0] = y_scale; st[1] = y_offset; st[2] = sd; st[3] = n_obs; st[4] = yty; st[5] = m0;
    st[6] = err; st[7] = 0.0;
  }
}

}  // namespace

extern "C" {

int ci_set_panel(ci_ctx* c, const ci_panel_args* a, const double* values, double* stats) {
  if (!c || !a || !values || !stats) return fail(CI_ERR_INVALID, "null argument");
  if (a->n_series < 1) return fail(CI_ERR_INVALID, "n_series must be >= 1");
  if (a->n_cols < 1) return fail(CI_ERR_INVALID, "values must have at least the outcome column");
  if (a->dtype != CI_F32 && a->dtype != CI_F64) return fail(CI_ERR_INVALID, "dtype must be 0 or 1");
  if (a->row0 < 0 || a->n_pre < 3 || a->row0 + a->n_pre > a->T_total)
    return fail(CI_ERR_INVALID, "bad pre-period: row0=%d n_pre=%d T_total=%d", a->row0, a->n_pre,
                a->T_total);
  const int N = a->n_series, Tm = a->T_total - a->row0;
  const int p = a->n_cols > 1 ? a->n_cols : 0;
  if (p + 2 > ci::MAX_DIM) return fail(CI_ERR_UNSUPPORTED, "p=%d exceeds the supported maximum", p);
  if (!(a->prior_level_sd > 0)) return fail(CI_ERR_INVALID, "prior_level_sd must be positive");
  CU_TRY(cudaSetDevice(c->device));
  c->free_retired();
  c->has_data = false; c->batch_n = 0;
  c->seas = ci::SeasDev{};
  c->esz = a->dtype == CI_F64 ? 8 : 4;
  c->NB = (Tm + ci::TB - 1) / ci::TB;
  c->ld = ci::tile_ld(p);
  c->dim = p + 2;
  ci_problem p0{};
  p0.model = CI_MODEL_LOCAL_LEVEL; p0.dtype = a->dtype; p0.T = Tm; p0.p = p;
  p0.ub_on_scale = a->ub_on_scale; p0.P0 = 1.0;
  c->prob = p0;
  {  // the pipeline must fit before anything is uploaded
    ci::SmemCfg cfg;
    int rc = plan_smem(c, 1, 0, &cfg);
    if (rc) return rc;
  }
  PanelDev d{};
  d.N = N; d.T_total = a->T_total; d.ncol = a->n_cols; d.row0 = a->row0; d.n_pre = a->n_pre;
  d.Tm = Tm; d.p = p; d.NB = c->NB; d.ld = c->ld; d.standardize = a->standardize ? 1 : 0;
  d.ub_on_scale = a->ub_on_scale ? 1 : 0; d.prior_level_sd = a->prior_level_sd;
  const size_t te = (size_t)ci::tile_elems(p);
  d.tile_stride = (size_t)c->NB * te;
  d.om_stride = (((size_t)p * p * c->esz + 15) & ~(size_t)15) / c->esz + 16 / c->esz;
  c->b_tile_stride = d.tile_stride * c->esz; c->b_omega_stride = d.om_stride * c->esz;
  c->b_gram_stride = (size_t)p * p * c->esz; c->b_xty_stride = (size_t)(p > 0 ? p : 1) * c->esz;
  const size_t vbytes = (size_t)N * a->T_total * a->n_cols * sizeof(double);
  const size_t dev_bytes = (size_t)N * (a->dtype == CI_F64 ? sizeof(ci::BatchDev<double>)
                                                           : sizeof(ci::BatchDev<float>));
  CU_TRY(c->w_raw.reserve(vbytes));
  CU_TRY(c->w_pstats.reserve((size_t)N * CI_PANEL_STATS * sizeof(double)));
  CU_TRY(c->b_tiles.reserve((size_t)N * c->b_tile_stride));
  CU_TRY(c->b_omega.reserve((size_t)N * c->b_omega_stride));
  CU_TRY(c->b_gram.reserve((size_t)N * c->b_gram_stride + 16));
  CU_TRY(c->b_xty.reserve((size_t)N * c->b_xty_stride + 16));
  CU_TRY(c->b_dev.reserve(dev_bytes));
  CU_TRY(cudaMemcpyAsync(c->w_raw.p, values, vbytes, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemsetAsync(c->b_omega.p, 0, (size_t)N * c->b_omega_stride, c->stream));
  if (a->dtype == CI_F64)
    k_panel_prep<double><<<N, PREP_THREADS, 0, c->stream>>>(
        static_cast<const double*>(c->w_raw.p), d, static_cast<double*>(c->b_tiles.p),
        static_cast<double*>(c->b_omega.p), static_cast<double*>(c->b_gram.p),
        static_cast<double*>(c->b_xty.p), static_cast<ci::BatchDev<double>*>(c->b_dev.p),
        static_cast<double*>(c->w_pstats.p));
  else
    k_panel_prep<float><<<N, PREP_THREADS, 0, c->stream>>>(
        static_cast<const double*>(c->w_raw.p), d, static_cast<float*>(c->b_tiles.p),
        static_cast<float*>(c->b_omega.p), static_cast<float*>(c->b_gram.p),
        static_cast<float*>(c->b_xty.p), static_cast<ci::BatchDev<float>*>(c->b_dev.p),
        static_cast<double*>(c->w_pstats.p));
  CU_TRY(cudaGetLastError());
  c->launches++;
  CU_TRY(cudaMemcpyAsync(stats, c->w_pstats.p, (size_t)N * CI_PANEL_STATS * sizeof(double),
                         cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  // host mirror of the per-series problems (ci_batch_select, validation of the batched runs)
  c->b_prob.assign(N, p0); c->b_yty.assign(N, 0.0); c->b_nobs.assign(N, 0);
  for (int s = 0; s < N; ++s) {
    const double* st = stats + (size_t)s * CI_PANEL_STATS;
    const int e = (int)st[6];
    if (e == 1) return fail(CI_ERR_INVALID, "Input response cannot be constant. (series %d)", s);
    if (e == 2) return fail(CI_ERR_INVALID, "Input data must have at least 3 observations. (series %d)", s);
    if (e == 3) return fail(CI_ERR_INVALID, "Input data cannot have any missing values. (series %d)", s);
    if (e == 4) return fail(CI_ERR_INVALID, "series %d has fewer than 2 observed pre-period points", s);
    const double sd = st[2], lvl0 = a->prior_level_sd * sd;
    ci_problem& q = c->b_prob[s];
    q.m0 = st[5]; q.P0 = sd * sd;
    q.obs_conc = p > 0 ? 25.0 : 0.005; q.obs_scale = (p > 0 ? 5.0 : 0.005) * sd * sd;
    q.obs_ub = 1.2 * sd;
    q.lvl_conc = 16.0; q.lvl_scale = 16.0 * lvl0 * lvl0; q.lvl_ub = sd;
    q.slope_conc = 16.0; q.slope_scale = q.lvl_scale; q.slope_ub = sd; q.P0_slope = sd * sd;
    c->b_yty[s] = st[4]; c->b_nobs[s] = (int)st[3];
  }
  c->batch_n = N;
  return ci_batch_select(c, 0);
}

int ci_predictive_mean_batch_d(ci_ctx* c, const void* theta_d, const void* level_d, int S,
                               void* mean_d, void* stream) {
  if (!c || !theta_d || !level_d || !mean_d) return fail(CI_ERR_INVALID, "null argument");
  if (c->batch_n < 1) return fail(CI_ERR_STATE, "no batch: call ci_set_data_batch or ci_set_panel");
  if (S < 1) return fail(CI_ERR_INVALID, "S must be >= 1");
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 blk(ci::MEAN_COLS, ci::MEAN_ROWS);
  const dim3 grid((c->prob.T + ci::MEAN_COLS - 1) / ci::MEAN_COLS, c->batch_n);
  // series 0's view + the stride to the next series' tiles
  const void* keep_tiles = c->v_tiles;
  c->v_tiles = c->b_tiles.p;
  if (c->prob.dtype == CI_F64)
    ci::k_predict_mean<double><<<grid, blk, 0, st>>>(
        make_probdev<double>(c), static_cast<const double*>(theta_d),
        static_cast<const double*>(level_d), S, static_cast<double*>(mean_d),
        c->b_tile_stride / sizeof(double));
  else
    ci::k_predict_mean<float><<<grid, blk, 0, st>>>(
        make_probdev<float>(c), static_cast<const float*>(theta_d),
        static_cast<const float*>(level_d), S, static_cast<float*>(mean_d),
        c->b_tile_stride / sizeof(float));
  c->v_tiles = keep_tiles;
  CU_TRY(cudaGetLastError());
  c->launches++;
  return CI_OK;
}

int ci_impact_batch_d(ci_ctx* c, const ci_impact_args* a, int n_series, const double* scale,
                      const double* offset, const double* obs_sum, const void* traj_d,
                      const void* mean_d, const double* observed, const uint8_t* period,
                      double* series_d, double* summ_d, void* stream) {
  if (!c || !a || !scale || !offset || !obs_sum || !traj_d || !mean_d || !observed || !period ||
      !series_d || !summ_d)
    return fail(CI_ERR_INVALID, "null argument");
  if (n_series < 1 || a->S < 1 || a->T < 1) return fail(CI_ERR_INVALID, "n_series, S and T must be >= 1");
  if (a->dtype != CI_F32 && a->dtype != CI_F64) return fail(CI_ERR_INVALID, "dtype must be 0 or 1");
  if (!(a->q_lo >= 0.0 && a->q_lo <= 1.0 && a->q_hi >= 0.0 && a->q_hi <= 1.0))
    return fail(CI_ERR_INVALID, "quantiles must be in [0,1]");
  const int S = a->S, T = a->T, N = n_series;
  ci::ImpactDev d{};
  d.S = S; d.T = T; d.scale = 1.0; d.offset = 0.0; d.q_lo = a->q_lo; d.q_hi = a->q_hi;
  d.t_c0 = T; d.n_post = 0;
  for (int t = 0; t < T; ++t) {
    if (period[t] > 2 || (t > 0 && period[t] < period[t - 1]))
      return fail(CI_ERR_INVALID, "period[] must be non-decreasing values in {0,1,2}");
    if (period[t] != 0 && d.t_c0 == T) d.t_c0 = t;
    d.n_post += period[t] == 1;
  }
  if (d.n_post < 1) return fail(CI_ERR_INVALID, "the post-period is empty");
  std::vector<ci::ImpactSeries> per(N);
  for (int s = 0; s < N; ++s) {
    if (!(scale[s] > 0.0)) return fail(CI_ERR_INVALID, "scale must be positive (series %d)", s);
    per[s].scale = scale[s]; per[s].offset = offset[s]; per[s].obs_sum = obs_sum[s];
  }
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Tc = T - d.t_c0;
  const size_t es = a->dtype == CI_F64 ? 8 : 4;
  CU_TRY(c->i_cum.reserve((size_t)N * S * (Tc > 0 ? Tc : 1) * sizeof(double)));
  CU_TRY(c->i_stats.reserve((size_t)N * S * ci::IMP_STATS * sizeof(double)));
  CU_TRY(c->i_trT.reserve((size_t)N * S * T * es));
  const size_t ob = (size_t)N * T * sizeof(double), pb = ((size_t)T + 7) & ~(size_t)7;
  const size_t sb = (size_t)N * sizeof(ci::ImpactSeries);
  CU_TRY(c->i_meta.reserve(ob + pb + sb));
  char* meta = static_cast<char*>(c->i_meta.p);
  {  // one asynchronous copy through the pinned ring (pageable sources would wait for the stream)
    const void* srcs[3] = {observed, period, per.data()};
    const size_t sizes[3] = {ob, (size_t)T, sb}, offs[3] = {0, ob, ob + pb};
    CU_TRY(c->ring.upload(meta, srcs, sizes, offs, 3, ob + pb + sb, st));
  }
  const double* obs_d = reinterpret_cast<const double*>(meta);
  const uint8_t* per_d = reinterpret_cast<const uint8_t*>(meta + ob);
  const ci::ImpactSeries* ps_d = reinterpret_cast<const ci::ImpactSeries*>(meta + ob + pb);
  int row_ctas, nseg;
  ci::impact_rows_grid(S, T, d.t_c0, true, &row_ctas, &nseg);
  size_t bytes; int in_smem, nt;
  select_launch_cfg(c, S, sizeof(double), &nt, &bytes, &in_smem);
  if (a->dtype == CI_F64) {
    ci::k_impact_rows<double><<<dim3(row_ctas * nseg, N), 32 * ci::IMP_WARPS, 0, st>>>(
        static_cast<const double*>(traj_d), static_cast<const double*>(mean_d), obs_d, per_d, d,
        static_cast<double*>(c->i_trT.p), static_cast<double*>(c->i_cum.p),
        static_cast<double*>(c->i_stats.p), series_d, summ_d, ps_d, row_ctas, ci::impact_seg_len(), ci::PeerDest{});
    CU_TRY(cudaGetLastError());
    auto kern = ci::k_impact_jobs<double>;
    CU_TRY(set_smem(kern, (uint32_t)bytes));
    kern<<<dim3(Tc + ci::IMP_STATS + T + 1, N), nt, bytes, st>>>(
        static_cast<const double*>(c->i_trT.p), static_cast<const double*>(c->i_cum.p),
        static_cast<const double*>(c->i_stats.p), obs_d, d, series_d, summ_d, in_smem, ps_d,
        ci::ImpactCols{0, T, 0, Tc, 1}, ci::ColBlocks{});
  } else {
    ci::k_impact_rows<float><<<dim3(row_ctas * nseg, N), 32 * ci::IMP_WARPS, 0, st>>>(
        static_cast<const float*>(traj_d), static_cast<const float*>(mean_d), obs_d, per_d, d,
        static_cast<float*>(c->i_trT.p), static_cast<double*>(c->i_cum.p),
        static_cast<double*>(c->i_stats.p), series_d, summ_d, ps_d, row_ctas, ci::impact_seg_len(), ci::PeerDest{});
    CU_TRY(cudaGetLastError());
    auto kern = ci::k_impact_jobs<float>;
    CU_TRY(set_smem(kern, (uint32_t)bytes));
    kern<<<dim3(Tc + ci::IMP_STATS + T + 1, N), nt, bytes, st>>>(
        static_cast<const float*>(c->i_trT.p), static_cast<const double*>(c->i_cum.p),
        static_cast<const double*>(c->i_stats.p), obs_d, d, series_d, summ_d, in_smem, ps_d,
        ci::ImpactCols{0, T, 0, Tc, 1}, ci::ColBlocks{});
  }
  CU_TRY(cudaGetLastError());
  c->launches += 2;
  return CI_OK;
}

}  // extern "C"
