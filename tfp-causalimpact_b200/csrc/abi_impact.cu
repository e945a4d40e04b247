// abi_impact.cu -- impact series + summary entry points.
#include "ci_host.cuh"
#include "ci_impact.cuh"

namespace {

using namespace ci;

template <typename R>
int launch_impact(ci_ctx* c, const ImpactDev& a, const void* traj_d, const void* mean_d,
                  const double* obs_d, const uint8_t* period_d, double* series_d, double* summ_d,
                  cudaStream_t st) {
  const int S = a.S, T = a.T, Tc = T - a.t_c0;
  R* trT = static_cast<R*>(c->i_trT.p);
  double* cumT = static_cast<double*>(c->i_cum.p);
  double* statsT = static_cast<double*>(c->i_stats.p);
  int row_ctas, nseg;
  impact_rows_grid(S, T, a.t_c0, true, &row_ctas, &nseg);             // + the predictive mean
  k_impact_rows<R><<<row_ctas * nseg, 32 * IMP_WARPS, 0, st>>>(
      static_cast<const R*>(traj_d), static_cast<const R*>(mean_d), obs_d, period_d, a, trT, cumT,
      statsT, series_d, summ_d, nullptr, row_ctas, impact_seg_len(), PeerDest{});
  CU_TRY(cudaGetLastError());
  c->launches++;
  size_t bytes; int in_smem, nt;                                      // float64 jobs size the staging
  select_launch_cfg(c, S, sizeof(double), &nt, &bytes, &in_smem);
  auto kern = k_impact_jobs<R>;
  CU_TRY(set_smem(kern, (uint32_t)bytes));
  kern<<<Tc + IMP_STATS + T + 1, nt, bytes, st>>>(trT, cumT, statsT, obs_d, a, series_d, summ_d,
                                                   in_smem, nullptr, ImpactCols{0, T, 0, Tc, 1}, ColBlocks{});
  CU_TRY(cudaGetLastError());
  c->launches++;
  return CI_OK;
}

}  // namespace

// Shared by the device entry points: argument checks, the period scan and the
// asynchronous upload of observed / period through the pinned ring.
int cih_impact_prepare(ci_ctx* c, const ci_impact_args* a, const double* observed,
                          const uint8_t* period, cudaStream_t st, ci::ImpactDev* d,
                          const double** obs_d, const uint8_t** per_d) {
  if (a->S < 1 || a->T < 1) return fail(CI_ERR_INVALID, "S and T must be >= 1");
  if (a->dtype != CI_F32 && a->dtype != CI_F64) return fail(CI_ERR_INVALID, "dtype must be 0 or 1");
  if (!(a->q_lo >= 0.0 && a->q_lo <= 1.0 && a->q_hi >= 0.0 && a->q_hi <= 1.0))
    return fail(CI_ERR_INVALID, "quantiles must be in [0,1]");
  if (!(a->scale > 0.0)) return fail(CI_ERR_INVALID, "scale must be positive");
  const int T = a->T;
  d->S = a->S; d->T = T; d->scale = a->scale; d->offset = a->offset; d->q_lo = a->q_lo;
  d->q_hi = a->q_hi; d->obs_sum = a->obs_sum;
  d->t_c0 = T; d->n_post = 0;
  for (int t = 0; t < T; ++t) {
    if (period[t] > 2 || (t > 0 && period[t] < period[t - 1]))
      return fail(CI_ERR_INVALID, "period[] must be non-decreasing values in {0,1,2}");
    if (period[t] != 0 && d->t_c0 == T) d->t_c0 = t;
    d->n_post += period[t] == 1;
  }
  if (d->n_post < 1) return fail(CI_ERR_INVALID, "the post-period is empty");
  CU_TRY(cudaSetDevice(c->device));
  const size_t ob = (size_t)T * sizeof(double);
  CU_TRY(c->i_meta.reserve(ob + (size_t)T));
  {  // one asynchronous copy through the pinned ring (pageable sources would wait for the stream)
    const void* srcs[2] = {observed, period};
    const size_t sizes[2] = {ob, (size_t)T}, offs[2] = {0, ob};
    CU_TRY(c->ring.upload(c->i_meta.p, srcs, sizes, offs, 2, ob + (size_t)T, st));
  }
  *obs_d = static_cast<const double*>(c->i_meta.p);
  *per_d = reinterpret_cast<const uint8_t*>(static_cast<char*>(c->i_meta.p) + ob);
  return CI_OK;
}

extern "C" {

int ci_impact_d(ci_ctx* c, const ci_impact_args* a, const void* traj_d, const void* mean_d,
                const double* observed, const uint8_t* period, double* series_d, double* summ_d,
                void* stream) {
  if (!c || !a || !traj_d || !mean_d || !observed || !period || !series_d || !summ_d)
    return fail(CI_ERR_INVALID, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ci::ImpactDev d{};
  const double* obs_d; const uint8_t* per_d;
  if (int rc = cih_impact_prepare(c, a, observed, period, st, &d, &obs_d, &per_d)) return rc;
  const int S = a->S, T = a->T, Tc = T - d.t_c0;
  CU_TRY(c->i_cum.reserve((size_t)S * (Tc > 0 ? Tc : 1) * sizeof(double)));
  CU_TRY(c->i_stats.reserve((size_t)S * ci::IMP_STATS * sizeof(double)));
  CU_TRY(c->i_trT.reserve((size_t)S * T * (a->dtype == CI_F64 ? 8 : 4)));
  if (a->dtype == CI_F64)
    return launch_impact<double>(c, d, traj_d, mean_d, obs_d, per_d, series_d, summ_d, st);
  return launch_impact<float>(c, d, traj_d, mean_d, obs_d, per_d, series_d, summ_d, st);
}

int ci_impact_rows_d(ci_ctx* c, const ci_impact_args* a, const void* traj_d, const void* mean_d,
                     const double* observed, const uint8_t* period, void* trT_d, double* cumT_d,
                     double* stats_d, double* series_d, double* summ_d, void* stream) {
  if (!c || !a || !traj_d || !observed || !period || !trT_d || !cumT_d || !stats_d)
    return fail(CI_ERR_INVALID, "null argument");
  if (mean_d && (!series_d || !summ_d))
    return fail(CI_ERR_INVALID, "series_d / summary_d are required with mean_d");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ci::ImpactDev d{};
  const double* obs_d; const uint8_t* per_d;
  if (int rc = cih_impact_prepare(c, a, observed, period, st, &d, &obs_d, &per_d)) return rc;
  int row_ctas, nseg;
  ci::impact_rows_grid(a->S, a->T, d.t_c0, mean_d != nullptr, &row_ctas, &nseg);
  if (a->dtype == CI_F64)
    ci::k_impact_rows<double><<<row_ctas * nseg, 32 * ci::IMP_WARPS, 0, st>>>(
        static_cast<const double*>(traj_d), static_cast<const double*>(mean_d), obs_d, per_d, d,
        static_cast<double*>(trT_d), cumT_d, stats_d, series_d, summ_d, nullptr, row_ctas,
        ci::impact_seg_len(), ci::PeerDest{});
  else
    ci::k_impact_rows<float><<<row_ctas * nseg, 32 * ci::IMP_WARPS, 0, st>>>(
        static_cast<const float*>(traj_d), static_cast<const float*>(mean_d), obs_d, per_d, d,
        static_cast<float*>(trT_d), cumT_d, stats_d, series_d, summ_d, nullptr, row_ctas,
        ci::impact_seg_len(), ci::PeerDest{});
  CU_TRY(cudaGetLastError());
  c->launches++;
  return CI_OK;
}

int ci_impact_cols_d(ci_ctx* c, const ci_impact_args* a, const void* trT_d, int t_begin,
                     int t_count, const double* cumT_d, int c_begin, int c_count,
                     const double* stats_d, const double* observed, const uint8_t* period,
                     double* series_d, double* summ_d, void* stream) {
  if (!c || !a || !observed || !period || !series_d || !summ_d)
    return fail(CI_ERR_INVALID, "null argument");
  if ((t_count > 0 && !trT_d) || (c_count > 0 && !cumT_d))
    return fail(CI_ERR_INVALID, "a column block is missing");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ci::ImpactDev d{};
  const double* obs_d; const uint8_t* per_d;
  if (int rc = cih_impact_prepare(c, a, observed, period, st, &d, &obs_d, &per_d)) return rc;
  const int Tc = a->T - d.t_c0;
  if (t_begin < 0 || t_count < 0 || t_begin + t_count > a->T || c_begin < 0 || c_count < 0 ||
      c_begin + c_count > Tc)
    return fail(CI_ERR_INVALID, "column block outside [0,T) / [0,T - t_c0)");
  const ci::ImpactCols jc{t_begin, t_count, c_begin, c_count, stats_d ? 1 : 0};
  const int jobs = c_count + t_count + (stats_d ? ci::IMP_STATS + 1 : 0);
  if (jobs == 0) return CI_OK;
  size_t bytes; int in_smem, nt;
  select_launch_cfg(c, a->S, sizeof(double), &nt, &bytes, &in_smem);
  if (a->dtype == CI_F64) {
    auto kern = ci::k_impact_jobs<double>;
    CU_TRY(set_smem(kern, (uint32_t)bytes));
    kern<<<jobs, nt, bytes, st>>>(static_cast<const double*>(trT_d), cumT_d, stats_d, obs_d, d,
                                  series_d, summ_d, in_smem, nullptr, jc, ci::ColBlocks{});
  } else {
    auto kern = ci::k_impact_jobs<float>;
    CU_TRY(set_smem(kern, (uint32_t)bytes));
    kern<<<jobs, nt, bytes, st>>>(static_cast<const float*>(trT_d), cumT_d, stats_d, obs_d, d,
                                  series_d, summ_d, in_smem, nullptr, jc, ci::ColBlocks{});
  }
  CU_TRY(cudaGetLastError());
  c->launches++;
  return CI_OK;
}

int ci_impact(ci_ctx* c, const ci_impact_args* a, const void* traj, const void* mean,
              const double* observed, const uint8_t* period, double* series, double* summary) {
  if (!c || !a || !traj || !mean || !observed || !period || !series || !summary)
    return fail(CI_ERR_INVALID, "null argument");
  if (a->S < 1 || a->T < 1) return fail(CI_ERR_INVALID, "S and T must be >= 1");
  if (a->dtype != CI_F32 && a->dtype != CI_F64) return fail(CI_ERR_INVALID, "dtype must be 0 or 1");
  CU_TRY(cudaSetDevice(c->device));
  const size_t es = a->dtype == CI_F64 ? 8 : 4;
  const size_t tb = (size_t)a->S * a->T * es, mb = (size_t)a->T * es;
  const size_t sb = (size_t)a->T * CI_IMPACT_SERIES_COLS * sizeof(double);
  const size_t ub = CI_IMPACT_SUMMARY_LEN * sizeof(double);
  CU_TRY(c->w_traj.reserve(tb));
  CU_TRY(c->w_mean.reserve(mb));
  CU_TRY(c->i_series.reserve(sb));
  CU_TRY(c->i_summ.reserve(ub));
  CU_TRY(cudaMemcpyAsync(c->w_traj.p, traj, tb, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemcpyAsync(c->w_mean.p, mean, mb, cudaMemcpyHostToDevice, c->stream));
  int rc = ci_impact_d(c, a, c->w_traj.p, c->w_mean.p, observed, period,
                       static_cast<double*>(c->i_series.p), static_cast<double*>(c->i_summ.p),
                       c->stream);
  if (rc) return rc;
  CU_TRY(cudaMemcpyAsync(series, c->i_series.p, sb, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaMemcpyAsync(summary, c->i_summ.p, ub, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return CI_OK;
}

}  // extern "C"
