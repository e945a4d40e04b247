// abi_seasonal.cu -- seasonal Gibbs entry points.
#include "ci_host.cuh"
#include "ci_seasonal.cuh"

namespace {

using namespace ci;

}  // namespace

// CI_ONLY (set by the build, _build.py): 0 = this object holds the float32 kernels AND the entry
// points, 1 = the float64 kernels only; undefined = everything in one object.
namespace cih {
using namespace ci;

template <typename R>
int launch_gibbs_seasonal(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0, int C,
                          void* draws_d, void* level_d, void* traj_d, float* incl_d, void* latent_d,
                          void* seas_d, void* drift_d, cudaStream_t st, bool batch = false);

template <typename R>
int launch_gibbs_seasonal(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0, int C,
                          void* draws_d, void* level_d, void* traj_d, float* incl_d, void* latent_d,
                          void* seas_d, void* drift_d, cudaStream_t st, bool batch) {
  const int p = c->prob.p, d = c->seas.d;
  const uint32_t base_extra = (uint32_t)(2 * p * p + 6 * p + 8) + (uint32_t)(d * (d | 1) + d) +
                              (3u + (uint32_t)c->seas.K) * (uint32_t)ci::TB;
  const uint32_t scr_elems = (uint32_t)c->prob.T * (uint32_t)(d + 1);
  const uint32_t tail = (uint32_t)(p * p) * (uint32_t)c->esz + 16u;
  SmemCfg cfg;
  int G = batch ? pick_G(c, C * c->batch_n) : pick_G(c, C), rc = CI_ERR_UNSUPPORTED;
  if (G > C) G = C;
  bool scr_smem = false;
  {  // first choice: the per-step scratch in shared memory with every tile resident
    std::string keep = cih_err();
    for (int g = G; g >= 1 && rc != CI_OK; --g) {
      SmemCfg t;
      if (plan_smem(c, g, base_extra + scr_elems, &t, tail) == CI_OK && t.resident) {
        cfg = t; G = g; rc = CI_OK; scr_smem = true;
      }
    }
    cih_err() = keep;
  }
  for (; rc != CI_OK && G >= 1; --G) {
    rc = plan_smem(c, G, base_extra, &cfg, tail);
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
  SeasDev sz = c->seas;
  sz.scratch = nullptr;                       // nullptr: the kernel's scratch is in shared memory
  if (!scr_smem) {
    CU_TRY(c->s_scratch.reserve((size_t)C * (batch ? c->batch_n : 1) * c->prob.T * (d + 1) * sizeof(R)));
    sz.scratch = c->s_scratch.p;
  }
  // state elements per lane: 1 (d <= 32), 2 (d <= 64: week-of-year), 6 (d <= 192: hour-of-week)
  auto kern = d <= 32 ? k_gibbs_seasonal<R, 1> : (d <= 64 ? k_gibbs_seasonal<R, 2> : k_gibbs_seasonal<R, 6>);
  CU_TRY(set_smem(kern, (uint32_t)cfg.total_bytes));
  const dim3 grid((C + G - 1) / G, batch ? c->batch_n : 1);
  kern<<<grid, 32 * (G + 1), cfg.total_bytes, st>>>(
      make_probdev<R>(c), gd, sz, cfg, plan, seed, chain_id0, C, static_cast<R*>(draws_d),
      static_cast<R*>(level_d), static_cast<R*>(traj_d), static_cast<R*>(latent_d),
      static_cast<R*>(seas_d), static_cast<R*>(drift_d), incl_d,
      batch ? static_cast<const BatchDev<R>*>(c->b_dev.p) : nullptr);
  CU_TRY(cudaGetLastError());
  c->launches++;
  return CI_OK;
}


#if !defined(CI_ONLY) || CI_ONLY == 0
template int launch_gibbs_seasonal<float>(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0, int C,
                          void* draws_d, void* level_d, void* traj_d, float* incl_d, void* latent_d,
                          void* seas_d, void* drift_d, cudaStream_t st, bool batch);
#endif
#if !defined(CI_ONLY) || CI_ONLY == 1
template int launch_gibbs_seasonal<double>(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0, int C,
                          void* draws_d, void* level_d, void* traj_d, float* incl_d, void* latent_d,
                          void* seas_d, void* drift_d, cudaStream_t st, bool batch);
#endif
#if defined(CI_ONLY) && CI_ONLY == 0
extern template int launch_gibbs_seasonal<double>(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0, int C,
                          void* draws_d, void* level_d, void* traj_d, float* incl_d, void* latent_d,
                          void* seas_d, void* drift_d, cudaStream_t st, bool batch);
#endif
}  // namespace cih

#if !defined(CI_ONLY) || CI_ONLY == 0
using cih::launch_gibbs_seasonal;

extern "C" {

int ci_gibbs_seasonal_run_batch_d(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed,
                                  uint64_t chain_id0, int Cs, void* draws_d, void* level_d,
                                  void* traj_d, float* incl_d, void* latent_d, void* seas_d,
                                  void* drift_d, void* stream) {
  if (!c || !o || !draws_d) return fail(CI_ERR_INVALID, "null argument");
  if (c->batch_n < 1) return fail(CI_ERR_STATE, "ci_set_data_batch has not been called");
  if (c->seas.K < 1) return fail(CI_ERR_STATE, "ci_set_seasonal has not been called");
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
    return launch_gibbs_seasonal<double>(c, o, seed, chain_id0, Cs, draws_d, level_d, traj_d, incl_d,
                                         latent_d, seas_d, drift_d, st, true);
  return launch_gibbs_seasonal<float>(c, o, seed, chain_id0, Cs, draws_d, level_d, traj_d, incl_d,
                                      latent_d, seas_d, drift_d, st, true);
}

int ci_gibbs_seasonal_run_d(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0,
                            int C, void* draws_d, void* level_d, void* traj_d, float* incl_d,
                            void* latent_d, void* seas_d, void* drift_d, void* stream) {
  if (!c || !o || !draws_d) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  if (c->seas.K < 1) return fail(CI_ERR_STATE, "ci_set_seasonal has not been called");
  if (C < 1 || o->n_results < 1 || o->n_warmup < 0)
    return fail(CI_ERR_INVALID, "n_chains >= 1, n_results >= 1, n_warmup >= 0 required");
  if (o->sparse && !(o->nonzero_prob > 0.0 && o->nonzero_prob <= 1.0))
    return fail(CI_ERR_INVALID, "nonzero_prob must be in (0, 1]");
  if (o->ssvs_order != 0 && o->ssvs_order != 1) return fail(CI_ERR_INVALID, "ssvs_order must be 0 or 1");
  if (c->n_obs < 2) return fail(CI_ERR_INVALID, "need at least 2 observed points");
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->prob.dtype == CI_F64)
    return launch_gibbs_seasonal<double>(c, o, seed, chain_id0, C, draws_d, level_d, traj_d, incl_d,
                                         latent_d, seas_d, drift_d, st);
  return launch_gibbs_seasonal<float>(c, o, seed, chain_id0, C, draws_d, level_d, traj_d, incl_d,
                                      latent_d, seas_d, drift_d, st);
}

int ci_gibbs_seasonal_run(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0,
                          int C, void* draws, void* level, void* traj, float* incl, void* latent,
                          void* seasonal, void* drift) {
  if (!c || !o || !draws) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  if (c->seas.K < 1) return fail(CI_ERR_STATE, "ci_set_seasonal has not been called");
  if (C < 1 || o->n_results < 1) return fail(CI_ERR_INVALID, "n_chains and n_results must be >= 1");
  CU_TRY(cudaSetDevice(c->device));
  const int K = c->seas.K;
  const size_t rows = (size_t)o->n_results * C;
  const size_t db = rows * c->dim * c->esz, tb = rows * c->prob.T * c->esz;
  const size_t ib = (size_t)C * (c->prob.p > 0 ? c->prob.p : 1) * sizeof(float);
  CU_TRY(c->w_draws.reserve(db));
  if (level) CU_TRY(c->w_level.reserve(tb));
  if (traj) CU_TRY(c->w_traj.reserve(tb));
  if (incl) CU_TRY(c->w_incl.reserve(ib));
  if (latent) CU_TRY(c->w_latent.reserve(tb));
  if (seasonal) CU_TRY(c->w_seas.reserve(tb * K));
  if (drift) CU_TRY(c->w_drift.reserve(rows * K * c->esz));
  int rc = ci_gibbs_seasonal_run_d(c, o, seed, chain_id0, C, c->w_draws.p,
                                   level ? c->w_level.p : nullptr, traj ? c->w_traj.p : nullptr,
                                   incl ? static_cast<float*>(c->w_incl.p) : nullptr,
                                   latent ? c->w_latent.p : nullptr,
                                   seasonal ? c->w_seas.p : nullptr, drift ? c->w_drift.p : nullptr,
                                   c->stream);
  if (rc) return rc;
  CU_TRY(cudaMemcpyAsync(draws, c->w_draws.p, db, cudaMemcpyDeviceToHost, c->stream));
  if (level) CU_TRY(cudaMemcpyAsync(level, c->w_level.p, tb, cudaMemcpyDeviceToHost, c->stream));
  if (traj) CU_TRY(cudaMemcpyAsync(traj, c->w_traj.p, tb, cudaMemcpyDeviceToHost, c->stream));
  if (latent) CU_TRY(cudaMemcpyAsync(latent, c->w_latent.p, tb, cudaMemcpyDeviceToHost, c->stream));
  if (seasonal) CU_TRY(cudaMemcpyAsync(seasonal, c->w_seas.p, tb * K, cudaMemcpyDeviceToHost, c->stream));
  if (drift) CU_TRY(cudaMemcpyAsync(drift, c->w_drift.p, rows * K * c->esz, cudaMemcpyDeviceToHost, c->stream));
  if (incl && c->prob.p > 0)
    CU_TRY(cudaMemcpyAsync(incl, c->w_incl.p, (size_t)C * c->prob.p * sizeof(float),
                           cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return CI_OK;
}

}  // extern "C"
#endif
