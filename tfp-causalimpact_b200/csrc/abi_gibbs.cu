// abi_gibbs.cu -- Gibbs (spike-and-slab) entry points, single series and batched.
#include "ci_host.cuh"
#include "ci_gibbs.cuh"
#include "ci_gibbs_team.cuh"

namespace {

using namespace ci;

template <typename R>
int launch_gibbs(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0, int C,
                 void* draws_d, void* level_d, void* traj_d, float* incl_d, cudaStream_t st,
                 bool batch = false) {
  const int p = c->prob.p;
  const uint32_t extra = (uint32_t)(2 * p * p + 6 * p + 8);
  const uint32_t tail = (uint32_t)(p * p) * (uint32_t)c->esz + 16u;
  SmemCfg cfg;
  // a batch runs C chains of EVERY series (grid.y = series): size CTAs for the whole grid
  int G = batch ? pick_G(c, C * c->batch_n) : pick_G(c, C), rc = CI_OK;
  if (G > C) G = C;
  for (; G >= 1; --G) {            // wide problems: fewer chains per CTA
    rc = plan_smem(c, G, extra, &cfg, tail);
    if (rc == CI_OK) break;
  }
  if (rc) return rc;
  GibbsPlan plan;
  plan.n_warmup = o->n_warmup; plan.n_results = o->n_results; plan.sparse = o->sparse ? 1 : 0;
  plan.n_obs = c->n_obs; plan.chain_major = o->chain_major ? 1 : 0;
  plan.ssvs_random = o->ssvs_order == 0 ? 1 : 0; plan.series_stride = o->series_stride;
  const double pi = o->nonzero_prob;
  plan.logit_pi = (plan.sparse && pi < 1.0) ? std::log(pi) - std::log1p(-pi) : 1e30;
  if (!(pi < 1.0)) plan.sparse = 0;
  GibbsDev<R> gd;
  gd.gram = static_cast<const R*>(c->v_gram); gd.xty0 = static_cast<const R*>(c->v_xty);
  gd.yty0 = (R)c->yty0;
  // team mode (ci_gibbs_team.cuh): one warp per tile when the series is resident in shared memory
  if (c->team_mode && c->gibbs_team && c->NB >= 1 && c->NB <= MAXW) {
    const int W = c->NB;
    const int total = batch ? C * c->batch_n : C;
    int gt = c->force_G > 0 ? c->force_G : (total + c->sm_count - 1) / c->sm_count;
    if (gt < 1) gt = 1;
    if (gt * W > GT_MAXWARPS) gt = GT_MAXWARPS / W;
    if (gt > C) gt = C;
    if (gt > 8) gt = 8;
    std::string keep = cih_err();
    for (; gt >= 1; --gt) {
      const uint32_t ttail = (uint32_t)gt * (uint32_t)sizeof(GibbsTeamShared<R>) + tail + 32u;
      SmemCfg tcfg;
      if (plan_smem(c, gt * W, extra, &tcfg, ttail, 0) == CI_OK && tcfg.resident) {
        cih_err() = keep;
        auto tk = k_gibbs_team<R>;
        CU_TRY(set_smem(tk, (uint32_t)tcfg.total_bytes));
        const dim3 tgrid((C + gt - 1) / gt, batch ? c->batch_n : 1);
        tk<<<tgrid, 32 * gt * W, tcfg.total_bytes, st>>>(
            make_probdev<R>(c), gd, tcfg, plan, W, seed, chain_id0, C, static_cast<R*>(draws_d),
            static_cast<R*>(level_d), static_cast<R*>(traj_d), incl_d,
            batch ? static_cast<const BatchDev<R>*>(c->b_dev.p) : nullptr);
        CU_TRY(cudaGetLastError());
        c->launches++;
        return CI_OK;
      }
    }
    cih_err() = keep;
  }
  auto kern = k_gibbs<R>;
  CU_TRY(set_smem(kern, (uint32_t)cfg.total_bytes));
  const dim3 grid((C + G - 1) / G, batch ? c->batch_n : 1);
  kern<<<grid, 32 * (G + 1), cfg.total_bytes, st>>>(
      make_probdev<R>(c), gd, cfg, plan, seed, chain_id0, C, static_cast<R*>(draws_d),
      static_cast<R*>(level_d), static_cast<R*>(traj_d), incl_d,
      batch ? static_cast<const BatchDev<R>*>(c->b_dev.p) : nullptr);
  CU_TRY(cudaGetLastError());
  c->launches++;
  return CI_OK;
}

}  // namespace

extern "C" {

int ci_gibbs_run_batch_d(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0,
                         int Cs, void* draws_d, void* level_d, void* traj_d, float* incl_d,
                         void* stream) {
  if (!c || !o || !draws_d) return fail(CI_ERR_INVALID, "null argument");
  if (c->batch_n < 1) return fail(CI_ERR_STATE, "ci_set_data_batch has not been called");
  if (Cs < 1 || o->n_results < 1 || o->n_warmup < 0)
    return fail(CI_ERR_INVALID, "n_chains >= 1, n_results >= 1, n_warmup >= 0 required");
  if (o->sparse && !(o->nonzero_prob > 0.0 && o->nonzero_prob <= 1.0))
    return fail(CI_ERR_INVALID, "nonzero_prob must be in (0, 1]");
  if (o->ssvs_order != 0 && o->ssvs_order != 1) return fail(CI_ERR_INVALID, "ssvs_order must be 0 or 1");
  for (int s = 0; s < c->batch_n; ++s)
    if (c->b_nobs[s] < 2) return fail(CI_ERR_INVALID, "series %d has fewer than 2 observed points", s);
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->prob.dtype == CI_F64)
    return launch_gibbs<double>(c, o, seed, chain_id0, Cs, draws_d, level_d, traj_d, incl_d, st, true);
  return launch_gibbs<float>(c, o, seed, chain_id0, Cs, draws_d, level_d, traj_d, incl_d, st, true);
}

int ci_gibbs_run_d(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0, int C,
                   void* draws_d, void* level_d, void* traj_d, float* incl_d, void* stream) {
  if (!c || !o || !draws_d) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  if (c->prob.model != CI_MODEL_LOCAL_LEVEL)
    return fail(CI_ERR_UNSUPPORTED, "ci_gibbs_run: local level model only (as the reference)");
  if (C < 1 || o->n_results < 1 || o->n_warmup < 0)
    return fail(CI_ERR_INVALID, "n_chains >= 1, n_results >= 1, n_warmup >= 0 required");
  if (o->sparse && !(o->nonzero_prob > 0.0 && o->nonzero_prob <= 1.0))
    return fail(CI_ERR_INVALID, "nonzero_prob must be in (0, 1]");
  if (o->ssvs_order != 0 && o->ssvs_order != 1) return fail(CI_ERR_INVALID, "ssvs_order must be 0 or 1");
  if (c->n_obs < 2) return fail(CI_ERR_INVALID, "need at least 2 observed points");
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->prob.dtype == CI_F64)
    return launch_gibbs<double>(c, o, seed, chain_id0, C, draws_d, level_d, traj_d, incl_d, st);
  return launch_gibbs<float>(c, o, seed, chain_id0, C, draws_d, level_d, traj_d, incl_d, st);
}

int ci_gibbs_run(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0, int C,
                 void* draws, void* level, void* traj, float* incl) {
  if (!c || !o || !draws) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  if (C < 1 || o->n_results < 1) return fail(CI_ERR_INVALID, "n_chains and n_results must be >= 1");
  CU_TRY(cudaSetDevice(c->device));
  const size_t rows = (size_t)o->n_results * C;
  const size_t db = rows * c->dim * c->esz, tb = rows * c->prob.T * c->esz;
  const size_t ib = (size_t)C * (c->prob.p > 0 ? c->prob.p : 1) * sizeof(float);
  CU_TRY(c->w_draws.reserve(db));
  if (level) CU_TRY(c->w_level.reserve(tb));
  if (traj) CU_TRY(c->w_traj.reserve(tb));
  if (incl) CU_TRY(c->w_incl.reserve(ib));
  int rc = ci_gibbs_run_d(c, o, seed, chain_id0, C, c->w_draws.p, level ? c->w_level.p : nullptr,
                          traj ? c->w_traj.p : nullptr,
                          incl ? static_cast<float*>(c->w_incl.p) : nullptr, c->stream);
  if (rc) return rc;
  CU_TRY(cudaMemcpyAsync(draws, c->w_draws.p, db, cudaMemcpyDeviceToHost, c->stream));
  if (level) CU_TRY(cudaMemcpyAsync(level, c->w_level.p, tb, cudaMemcpyDeviceToHost, c->stream));
  if (traj) CU_TRY(cudaMemcpyAsync(traj, c->w_traj.p, tb, cudaMemcpyDeviceToHost, c->stream));
  if (incl && c->prob.p > 0)
    CU_TRY(cudaMemcpyAsync(incl, c->w_incl.p, (size_t)C * c->prob.p * sizeof(float),
                           cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return CI_OK;
}

}  // extern "C"
