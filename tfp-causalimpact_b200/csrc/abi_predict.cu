// abi_predict.cu -- K4 / K5 entry points: simulation smoother, predictive mean, per-time quantiles.
#include "ci_host.cuh"
#include "ci_predict.cuh"
#include "ci_llt_predict.cuh"
#include "ci_team_kernels.cuh"

namespace {

using namespace ci;

template <typename R>
int launch_predict(ci_ctx* c, const void* theta_d, int S, uint64_t seed, uint64_t draw_id0,
                   void* level_d, void* traj_d, void* mean_d, cudaStream_t st) {
  SmemCfg cfg;
  const ProbDev<R> prt = make_probdev<R>(c);
  int GT = 0;
  if (c->prob.model == CI_MODEL_LOCAL_LINEAR_TREND) {
    // d = 2 simulation smoother (ci_llt_predict.cuh): one warp per draw, any T
    const int G = pick_G(c, S);
    int rc = plan_smem(c, G, 0, &cfg);
    if (rc) return rc;
    auto lk = k_predict_llt<R>;
    CU_TRY(set_smem(lk, (uint32_t)cfg.total_bytes));
    lk<<<(S + G - 1) / G, 32 * (G + 1), cfg.total_bytes, st>>>(
        prt, make_lltdev<R>(c), cfg, static_cast<const R*>(theta_d), S, seed, draw_id0,
        static_cast<R*>(level_d), nullptr, static_cast<R*>(traj_d));
    CU_TRY(cudaGetLastError());
    c->launches++;
    if (mean_d) {
      k_predict_mean<R><<<(c->prob.T + MEAN_COLS - 1) / MEAN_COLS, dim3(MEAN_COLS, MEAN_ROWS), 0, st>>>(
          prt, static_cast<const R*>(theta_d), static_cast<const R*>(level_d), S,
          static_cast<R*>(mean_d));
      CU_TRY(cudaGetLastError());
      c->launches++;
    }
    return CI_OK;
  }
  // The kernel choice must NOT depend on S: a draw has to come out bit-identical however
  // the batch is split over calls / GPUs.
  // Team kernel (one warp per tile of a draw) whenever the series is resident: the choice
  // depends on the SHAPE only; the number of teams per CTA may follow S (it does not change a
  // draw's arithmetic).
  bool team_ok = false;
  if (c->predict_team && c->team_mode && c->NB >= 1 && c->NB <= MAXW &&
      c->prob.model == CI_MODEL_LOCAL_LEVEL) {
    const int W = c->NB;
    int gt = c->force_G > 0 ? c->force_G : (S + c->sm_count - 1) / c->sm_count;
    if (gt < 1) gt = 1;
    if (gt * W > PT_MAXWARPS) gt = PT_MAXWARPS / W;
    std::string keep = cih_err();
    for (; gt >= 1 && !team_ok; --gt) {
      const uint32_t tail = (uint32_t)gt * (uint32_t)sizeof(TeamShared<R>) + 16u;
      if (plan_smem(c, gt * W, 0, &cfg, tail, 0) == CI_OK && cfg.resident) { team_ok = true; GT = gt; }
    }
    cih_err() = keep;
  }
  if (team_ok) {
    auto tk = k_predict_team<R>;
    CU_TRY(set_smem(tk, (uint32_t)cfg.total_bytes));
    tk<<<(S + GT - 1) / GT, 32 * GT * c->NB, cfg.total_bytes, st>>>(
        prt, cfg, c->NB, static_cast<const R*>(theta_d), S, seed, draw_id0,
        static_cast<R*>(level_d), static_cast<R*>(traj_d));
    CU_TRY(cudaGetLastError());
    c->launches++;
    if (mean_d) {
      k_predict_mean<R><<<(c->prob.T + MEAN_COLS - 1) / MEAN_COLS, dim3(MEAN_COLS, MEAN_ROWS), 0, st>>>(
          prt, static_cast<const R*>(theta_d), static_cast<const R*>(level_d), S,
          static_cast<R*>(mean_d));
      CU_TRY(cudaGetLastError());
      c->launches++;
    }
    return CI_OK;
  }
  const int G = pick_G(c, S);
  int rc = plan_smem(c, G, 0, &cfg);
  if (rc) return rc;
  auto kern = k_predict<R>;
  CU_TRY(set_smem(kern, (uint32_t)cfg.total_bytes));
  const int grid = (S + G - 1) / G;
  const ProbDev<R> pr = make_probdev<R>(c);
  kern<<<grid, 32 * (G + 1), cfg.total_bytes, st>>>(pr, cfg, static_cast<const R*>(theta_d), S,
                                                     seed, draw_id0, static_cast<R*>(level_d),
                                                     static_cast<R*>(traj_d));
  CU_TRY(cudaGetLastError());
  c->launches++;
  if (mean_d) {
    k_predict_mean<R><<<(c->prob.T + MEAN_COLS - 1) / MEAN_COLS, dim3(MEAN_COLS, MEAN_ROWS), 0, st>>>(
        pr, static_cast<const R*>(theta_d), static_cast<const R*>(level_d), S,
        static_cast<R*>(mean_d));
    CU_TRY(cudaGetLastError());
    c->launches++;
  }
  return CI_OK;
}

template <typename R>
int launch_quantiles(ci_ctx* c, const void* a_d, int S, int T, const double* q, int nq,
                     void* out_d, cudaStream_t st, int out_ld = 0) {
  // the whole column lives in shared memory as integer keys when it fits; longer columns are
  // selected straight from global memory (every sweep re-reads them through L2)
  size_t bytes; int in_smem, nt;
  select_launch_cfg(c, S, sizeof(R), &nt, &bytes, &in_smem);
  QuantArgs qa;
  qa.nq = nq;
  for (int i = 0; i < nq; ++i) qa.q[i] = q[i];
  auto kern = k_row_quantiles<R>;
  CU_TRY(set_smem(kern, (uint32_t)bytes));
  kern<<<T, nt, bytes, st>>>(static_cast<const R*>(a_d), S, T, qa, static_cast<R*>(out_d),
                             out_ld > 0 ? out_ld : nq, in_smem);
  CU_TRY(cudaGetLastError());
  c->launches++;
  return CI_OK;
}

}  // namespace

extern "C" {

int ci_posterior_predict_d(ci_ctx* c, const void* theta_d, int S, uint64_t seed, uint64_t draw_id0,
                           void* level_d, void* traj_d, void* mean_d, void* stream) {
  if (!c || !theta_d || !traj_d) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  if (S < 1) return fail(CI_ERR_INVALID, "S must be >= 1");
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!level_d && mean_d) {   // the mean needs the level paths: use the workspace
    CU_TRY(c->w_level.reserve((size_t)S * c->prob.T * c->esz));
    level_d = c->w_level.p;
  }
  if (c->prob.dtype == CI_F64)
    return launch_predict<double>(c, theta_d, S, seed, draw_id0, level_d, traj_d, mean_d, st);
  return launch_predict<float>(c, theta_d, S, seed, draw_id0, level_d, traj_d, mean_d, st);
}

int ci_posterior_predict(ci_ctx* c, const void* theta, int S, uint64_t seed, uint64_t draw_id0,
                         void* level, void* traj, void* mean) {
  if (!c || !theta || !traj) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  if (S < 1) return fail(CI_ERR_INVALID, "S must be >= 1");
  CU_TRY(cudaSetDevice(c->device));
  const size_t tb = (size_t)S * c->dim * c->esz, st_b = (size_t)S * c->prob.T * c->esz;
  const size_t mb = (size_t)c->prob.T * c->esz;
  CU_TRY(c->w_theta.reserve(tb));
  CU_TRY(c->w_level.reserve(st_b));
  CU_TRY(c->w_traj.reserve(st_b));
  CU_TRY(c->w_mean.reserve(mb));
  CU_TRY(cudaMemcpyAsync(c->w_theta.p, theta, tb, cudaMemcpyHostToDevice, c->stream));
  int rc = ci_posterior_predict_d(c, c->w_theta.p, S, seed, draw_id0, c->w_level.p, c->w_traj.p,
                                  mean ? c->w_mean.p : nullptr, c->stream);
  if (rc) return rc;
  if (level) CU_TRY(cudaMemcpyAsync(level, c->w_level.p, st_b, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaMemcpyAsync(traj, c->w_traj.p, st_b, cudaMemcpyDeviceToHost, c->stream));
  if (mean) CU_TRY(cudaMemcpyAsync(mean, c->w_mean.p, mb, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return CI_OK;
}

int ci_row_quantiles_d(ci_ctx* c, const void* a_d, int S, int T, int dtype, const double* q,
                       int nq, void* out_d, void* stream) {
  if (!c || !a_d || !q || !out_d) return fail(CI_ERR_INVALID, "null argument");
  if (S < 1 || T < 1) return fail(CI_ERR_INVALID, "S and T must be >= 1");
  if (nq < 1 || nq > 8) return fail(CI_ERR_INVALID, "nq must be in [1,8]");
  for (int i = 0; i < nq; ++i)
    if (!(q[i] >= 0.0 && q[i] <= 1.0)) return fail(CI_ERR_INVALID, "quantile %d out of [0,1]", i);
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == CI_F64) return launch_quantiles<double>(c, a_d, S, T, q, nq, out_d, st);
  if (dtype == CI_F32) return launch_quantiles<float>(c, a_d, S, T, q, nq, out_d, st);
  return fail(CI_ERR_INVALID, "dtype must be 0 or 1");
}

int ci_row_quantiles(ci_ctx* c, const void* a, int S, int T, int dtype, const double* q, int nq,
                     void* out) {
  if (!c || !a || !q || !out) return fail(CI_ERR_INVALID, "null argument");
  if (S < 1 || T < 1) return fail(CI_ERR_INVALID, "S and T must be >= 1");
  if (dtype != CI_F32 && dtype != CI_F64) return fail(CI_ERR_INVALID, "dtype must be 0 or 1");
  CU_TRY(cudaSetDevice(c->device));
  const size_t es = dtype == CI_F64 ? 8 : 4;
  const size_t ab = (size_t)S * T * es, ob = (size_t)T * nq * es;
  CU_TRY(c->w_traj.reserve(ab));
  CU_TRY(c->w_q.reserve(ob > 0 ? ob : 16));
  CU_TRY(cudaMemcpyAsync(c->w_traj.p, a, ab, cudaMemcpyHostToDevice, c->stream));
  int rc = ci_row_quantiles_d(c, c->w_traj.p, S, T, dtype, q, nq, c->w_q.p, c->stream);
  if (rc) return rc;
  CU_TRY(cudaMemcpyAsync(out, c->w_q.p, ob, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return CI_OK;
}

int ci_predictive_mean_d(ci_ctx* c, const void* theta_d, const void* level_d, int S, void* mean_d,
                         void* stream) {
  if (!c || !theta_d || !level_d || !mean_d) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  if (S < 1) return fail(CI_ERR_INVALID, "S must be >= 1");
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int T = c->prob.T, p = c->prob.p, dim = c->dim;
  // enough draw ranges for ~4 CTAs per SM, at least 64 draws each
  const int cx = (T + ci::MEANP_COLS - 1) / ci::MEANP_COLS;
  int SY = (4 * c->sm_count + cx - 1) / cx;
  if (SY > ci::MEANP_MAXY) SY = ci::MEANP_MAXY;
  if (SY > (S + 63) / 64) SY = (S + 63) / 64;
  if (SY < 1) SY = 1;
  CU_TRY(c->w_mpart.reserve(((size_t)SY * T + (size_t)SY * (p > 0 ? p : 1)) * sizeof(double)));
  double* colpart = static_cast<double*>(c->w_mpart.p);
  double* wpart = colpart + (size_t)SY * T;
  const dim3 blk(ci::MEANP_COLS, ci::MEANP_ROWS), grid(cx, SY);
  if (c->prob.dtype == CI_F64) {
    ci::k_mean_partial<double><<<grid, blk, 0, st>>>(
        static_cast<const double*>(theta_d), static_cast<const double*>(level_d), S, T, dim, p,
        colpart, wpart);
    ci::k_mean_final<double><<<(T + 255) / 256, 256, 0, st>>>(
        make_probdev<double>(c), colpart, wpart, S, SY, static_cast<double*>(mean_d));
  } else {
    ci::k_mean_partial<float><<<grid, blk, 0, st>>>(
        static_cast<const float*>(theta_d), static_cast<const float*>(level_d), S, T, dim, p,
        colpart, wpart);
    ci::k_mean_final<float><<<(T + 255) / 256, 256, 0, st>>>(
        make_probdev<float>(c), colpart, wpart, S, SY, static_cast<float*>(mean_d));
  }
  c->launches++;
  CU_TRY(cudaGetLastError());
  c->launches++;
  return CI_OK;
}

}  // extern "C"
