// abi_hmc.cu -- K6 entry points: persistent HMC kernels.
#include "ci_host.cuh"
#include "ci_team_kernels.cuh"
#include "ci_llt_kernels.cuh"

namespace {

using namespace ci;

// Stan-style windowed adaptation schedule (oracle/hmc_np.py:adapt_schedule).
void make_hmc_plan(const ci_hmc_opts* o, uint64_t seed, HmcPlan* pl) {
  HmcPlan h{};
  h.n_warmup = o->n_warmup; h.n_results = o->n_results; h.max_leapfrog = o->max_leapfrog;
  h.adapt_mass = o->adapt_mass; h.init_step = o->init_step;
  h.target_accept = o->target_accept;
  const int W = o->n_warmup;
  if (W < 20) { h.init_buf = W; h.slow_end = W; h.n_ends = 0; }
  else {
    int init = 75, term = 50, base = 25;
    if (init + base + term > W) { init = (int)(0.15 * W); term = (int)(0.1 * W); base = W - init - term; }
    const int last = W - term - 1;
    int size = base, nxt = init + base - 1, n = 0;
    while (n < 16) {
      h.ends[n++] = nxt;
      if (nxt == last) break;
      size *= 2;
      int n2 = nxt + size;
      if (n2 != last && n2 + 2 * size >= W - term) n2 = last;
      if (n2 > last) n2 = last;
      nxt = n2;
    }
    h.init_buf = init; h.slow_end = W - term; h.n_ends = n;
  }
  long long ev = 1;
  for (int it = 0; it < o->n_warmup + o->n_results; ++it)
    ev += hmc_leapfrog_count(seed, it, o->max_leapfrog);
  h.n_evals = ev;
  *pl = h;
}

}  // namespace

// CI_ONLY (set by the build, _build.py): 0 = this object holds the float32 kernels AND the entry
// points, 1 = the float64 kernels only; undefined = everything in one object.
namespace cih {
using namespace ci;

template <typename R>
int launch_hmc(ci_ctx* c, const ci_hmc_opts* o, uint64_t seed, uint64_t chain_id0,
               const void* theta0_d, int C, void* draws_d, ci_hmc_stats* stats_d,
               cudaStream_t st);

template <typename R>
int launch_hmc(ci_ctx* c, const ci_hmc_opts* o, uint64_t seed, uint64_t chain_id0,
               const void* theta0_d, int C, void* draws_d, ci_hmc_stats* stats_d,
               cudaStream_t st) {
  SmemCfg cfg;
  HmcPlan plan;
  make_hmc_plan(o, seed, &plan);
  if (c->prob.model == CI_MODEL_LOCAL_LINEAR_TREND) {
    const int G = pick_G(c, C);
    int rc = plan_smem(c, G, 0, &cfg);
    if (rc) return rc;
    auto lk = k_hmc_llt<R>;
    CU_TRY(set_smem(lk, (uint32_t)cfg.total_bytes));
    lk<<<(C + G - 1) / G, 32 * (G + 1), cfg.total_bytes, st>>>(
        make_probdev<R>(c), make_lltdev<R>(c), cfg, plan, seed, chain_id0,
        static_cast<const R*>(theta0_d), C, static_cast<R*>(draws_d), stats_d);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return CI_OK;
  }
  int GT = 0;
  if (plan_team<R>(c, C, &GT, &cfg)) {
    const int W = c->NB;
    auto tk = k_hmc_team<R>;
    CU_TRY(set_smem(tk, (uint32_t)cfg.total_bytes));
    tk<<<(C + GT - 1) / GT, 32 * (GT * W + 1), cfg.total_bytes, st>>>(
        make_probdev<R>(c), cfg, W, plan, seed, chain_id0, static_cast<const R*>(theta0_d), C,
        static_cast<R*>(draws_d), stats_d);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return CI_OK;
  }
  int TW = 0;
  if (plan_tstream<R>(c, C, &GT, &TW, &cfg)) {
    auto sk = TW != TS_W ? k_hmc_tstream<R, 0, 0>
              : (c->prob.p == 2 ? k_hmc_tstream<R, 2, TS_W> : k_hmc_tstream<R, 0, TS_W>);
    CU_TRY(set_smem(sk, (uint32_t)cfg.total_bytes));
    sk<<<(C + GT - 1) / GT, 32 * GT * TW, cfg.total_bytes, st>>>(
        make_probdev<R>(c), cfg, TW, plan, seed, chain_id0, static_cast<const R*>(theta0_d), C,
        static_cast<R*>(draws_d), stats_d);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return CI_OK;
  }
  const int G = pick_G(c, C);
  int rc = plan_smem(c, G, 0, &cfg);
  if (rc) return rc;
  auto kern = k_hmc<R>;
  CU_TRY(set_smem(kern, (uint32_t)cfg.total_bytes));
  const int grid = (C + G - 1) / G;
  kern<<<grid, 32 * (G + 1), cfg.total_bytes, st>>>(
      make_probdev<R>(c), cfg, plan, seed, chain_id0, static_cast<const R*>(theta0_d), C,
      static_cast<R*>(draws_d), stats_d);
  CU_TRY(cudaGetLastError());
  c->launches++;
  return CI_OK;
}


#if !defined(CI_ONLY) || CI_ONLY == 0
template int launch_hmc<float>(ci_ctx* c, const ci_hmc_opts* o, uint64_t seed, uint64_t chain_id0,
               const void* theta0_d, int C, void* draws_d, ci_hmc_stats* stats_d,
               cudaStream_t st);
#endif
#if !defined(CI_ONLY) || CI_ONLY == 1
template int launch_hmc<double>(ci_ctx* c, const ci_hmc_opts* o, uint64_t seed, uint64_t chain_id0,
               const void* theta0_d, int C, void* draws_d, ci_hmc_stats* stats_d,
               cudaStream_t st);
#endif
#if defined(CI_ONLY) && CI_ONLY == 0
extern template int launch_hmc<double>(ci_ctx* c, const ci_hmc_opts* o, uint64_t seed, uint64_t chain_id0,
               const void* theta0_d, int C, void* draws_d, ci_hmc_stats* stats_d,
               cudaStream_t st);
#endif
}  // namespace cih

#if !defined(CI_ONLY) || CI_ONLY == 0
using cih::launch_hmc;

extern "C" {

int ci_hmc_run_d(ci_ctx* c, const ci_hmc_opts* o, uint64_t seed, uint64_t chain_id0,
                 const void* theta0_d, int C, void* draws_d, ci_hmc_stats* stats_d, void* stream) {
  if (!c || !o || !theta0_d || !draws_d || !stats_d) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  if (C < 1) return fail(CI_ERR_INVALID, "n_chains must be >= 1");
  if (o->n_warmup < 0 || o->n_results < 1) return fail(CI_ERR_INVALID, "n_warmup >= 0 and n_results >= 1 required");
  if (o->max_leapfrog < 1 || o->max_leapfrog > 1024) return fail(CI_ERR_INVALID, "max_leapfrog must be in [1,1024]");
  if (!(o->init_step > 0)) return fail(CI_ERR_INVALID, "init_step must be positive");
  if (!(o->target_accept > 0 && o->target_accept < 1)) return fail(CI_ERR_INVALID, "target_accept must be in (0,1)");
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->prob.dtype == CI_F64)
    return launch_hmc<double>(c, o, seed, chain_id0, theta0_d, C, draws_d, stats_d, st);
  return launch_hmc<float>(c, o, seed, chain_id0, theta0_d, C, draws_d, stats_d, st);
}

int ci_hmc_run(ci_ctx* c, const ci_hmc_opts* o, uint64_t seed, uint64_t chain_id0,
               const void* theta0, int C, void* draws, ci_hmc_stats* stats) {
  if (!c || !o || !theta0 || !draws || !stats) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  if (C < 1 || o->n_results < 1) return fail(CI_ERR_INVALID, "n_chains and n_results must be >= 1");
  CU_TRY(cudaSetDevice(c->device));
  const size_t tb = (size_t)C * c->dim * c->esz;
  const size_t db = (size_t)o->n_results * tb, sb = (size_t)C * sizeof(ci_hmc_stats);
  CU_TRY(c->w_theta.reserve(tb));
  CU_TRY(c->w_draws.reserve(db));
  CU_TRY(c->w_stats.reserve(sb));
  CU_TRY(cudaMemcpyAsync(c->w_theta.p, theta0, tb, cudaMemcpyHostToDevice, c->stream));
  int rc = ci_hmc_run_d(c, o, seed, chain_id0, c->w_theta.p, C, c->w_draws.p,
                        static_cast<ci_hmc_stats*>(c->w_stats.p), c->stream);
  if (rc) return rc;
  CU_TRY(cudaMemcpyAsync(draws, c->w_draws.p, db, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaMemcpyAsync(stats, c->w_stats.p, sb, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return CI_OK;
}

}  // extern "C"
#endif
