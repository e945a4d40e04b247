// ci_abi.cu -- the C ABI declared in include/ci_b200.h.
// Plain pointers and sizes only; no torch types.  There is no CPU fallback:
// every compute entry point fails with CI_ERR_NO_DEVICE / CI_ERR_CUDA when no
// sm_100 device is usable.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ci_b200.h"
#include "ci_kernels.cuh"
#include "ci_predict.cuh"
#include "ci_hmc.cuh"
#include "ci_team_kernels.cuh"
#include "ci_seq.cuh"
#include "ci_llt_kernels.cuh"
#include "ci_gibbs.cuh"
#include "ci_impact.cuh"
#include "ci_seasonal.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU_TRY(expr)                                                                   \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return fail(CI_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                  __FILE__, __LINE__);                                                 \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct ci_ctx {
  int device = -1;
  int sm_count = 0;
  int smem_optin = 0;
  cudaStream_t stream = nullptr;
  bool has_data = false;
  ci_problem prob{};
  int NB = 0, ld = 0, dim = 0;
  size_t esz = 4;
  DevBuf tiles, omega;
  DevBuf w_theta, w_value, w_grad;   // workspaces of the host-pointer entry points
  DevBuf w_level, w_traj, w_mean, w_q, w_draws, w_stats, w_incl;
  DevBuf gram, xty0;                 // X'X, X'y over observed rows (Gibbs regression step)
  DevBuf i_cum, i_stats, i_meta, i_series, i_summ, i_trT;   // ci_impact workspaces
  DevBuf s_sched, s_scratch, s_series, w_latent, w_seas, w_drift;   // seasonal components
  ci::SeasDev seas{};                // seas.K == 0: no seasonal components
  // views of the CURRENT series: the context's own buffers after ci_set_data, a slice of the
  // batch buffers after ci_batch_select
  const void* v_tiles = nullptr; const void* v_omega = nullptr;
  const void* v_gram = nullptr; const void* v_xty = nullptr;
  // batch of independent series (ci_set_data_batch)
  int batch_n = 0;
  DevBuf b_tiles, b_omega, b_gram, b_xty, b_dev;
  size_t b_tile_stride = 0, b_omega_stride = 0, b_gram_stride = 0, b_xty_stride = 0;   // bytes
  std::vector<ci_problem> b_prob;
  std::vector<double> b_yty;
  std::vector<int> b_nobs;
  double yty0 = 0.0;
  int n_obs = 0;
  int64_t launches = 0;
  int force_G = 0;                   // CI_B200_G env override (tuning)
  int team_mode = 1;                 // CI_B200_TEAM=0 disables the warp-team kernels
  int predict_team = 0;              // CI_B200_PREDICT_TEAM=1: team kernel for ci_posterior_predict
  int team_sub = 0;                  // CI_B200_TEAM_SUB=1: sub-tile (4 steps / lane) team kernels (experiment)
};

namespace {

using namespace ci;

template <typename R> ProbDev<R> make_probdev(const ci_ctx* c) {
  ProbDev<R> pr;
  pr.tiles = static_cast<const R*>(c->v_tiles);
  pr.omega = static_cast<const R*>(c->v_omega);
  pr.T = c->prob.T; pr.p = c->prob.p; pr.ld = c->ld; pr.NB = c->NB; pr.dim = c->dim;
  pr.model = c->prob.model;
  pr.m0 = (R)c->prob.m0; pr.P0 = (R)c->prob.P0;
  pr.obs_conc = (R)c->prob.obs_conc; pr.obs_scale = (R)c->prob.obs_scale;
  pr.obs_ub = (R)c->prob.obs_ub;
  pr.lvl_conc = (R)c->prob.lvl_conc; pr.lvl_scale = (R)c->prob.lvl_scale;
  pr.lvl_ub = (R)c->prob.lvl_ub;
  return pr;
}

template <typename R> LltDev<R> make_lltdev(const ci_ctx* c) {
  LltDev<R> d;
  d.q_conc = (R)c->prob.slope_conc; d.q_scale = (R)c->prob.slope_scale;
  d.q_ub = (R)(c->prob.slope_ub > 1e30 ? 1e30 : c->prob.slope_ub);
  d.m0s = (R)c->prob.m0_slope; d.P0s = (R)c->prob.P0_slope;
  return d;
}

// static shared memory of the select kernels (16 x 256 histograms + bookkeeping), rounded up
constexpr size_t QSTATIC = 28 * 1024;

inline uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

// Opt in to `bytes` of dynamic shared memory AND ask for the largest shared-memory
// carveout: without the second attribute the driver may pick a carveout that fits a
// single CTA per SM (ncu, round 1 run 7: occupancy_limit_shared_mem = 1 at 73 KB/CTA).
template <typename Kern> cudaError_t set_smem(Kern kern, uint32_t bytes) {
  // Two driver calls per launch cost ~2 us of host time on the e2e path: remember what
  // was last set per (kernel ADDRESS, device) -- kernels with equal signatures share this
  // template instantiation, so the key must be the pointer -- and skip when unchanged.
  struct Slot { const void* fn; int dev; uint32_t bytes; };
  static thread_local Slot cache[64] = {};
  const void* fn = reinterpret_cast<const void*>(kern);
  int dev = 0;
  cudaGetDevice(&dev);
  Slot* slot = nullptr;
  for (auto& sl : cache) {
    if (sl.fn == fn && sl.dev == dev) { slot = &sl; break; }
    if (sl.fn == nullptr) { slot = &sl; break; }
  }
  if (slot && slot->fn == fn && slot->bytes == bytes) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                           (int)cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess && slot) { slot->fn = fn; slot->dev = dev; slot->bytes = bytes; }
  return e;
}

// Shared-memory plan for a kernel with G consumer warps and `extra_elems`
// kernel-specific per-warp scratch elements.
int plan_smem(const ci_ctx* c, int G, uint32_t extra_elems, SmemCfg* out,
              uint32_t tail_bytes = 0) {
  const uint32_t esz = (uint32_t)c->esz;
  const int p = c->prob.p, NB = c->NB;
  SmemCfg cfg{};
  cfg.stage_elems = (uint32_t)tile_elems(p);
  const uint32_t stage_bytes = cfg.stage_elems * esz;
  // per-warp scratch
  uint32_t e = 0;
  cfg.w_off = e;     e += align_up((uint32_t)(p + 4), 4);   // holds the full theta in the HMC kernel
  cfg.rbuf_off = e;  e += TB + 8;
  cfg.ckpt_off = e;  e += align_up(6u * (uint32_t)NB, 4);    // (a,P) or the 5-value trend state
  cfg.extra_off = e; e += align_up(extra_elems, 4);
  cfg.warp_bytes = align_up(e * esz, 16);
  const uint32_t omega_bytes = align_up((uint32_t)(p * p) * esz, 16);
  const uint32_t fixed = omega_bytes + (uint32_t)G * cfg.warp_bytes + tail_bytes + 16u;
  const uint32_t budget = (uint32_t)c->smem_optin;
  // stages: as many as fit (each stage also needs 16 bytes of barriers)
  // one stage is enough to be correct (no copy/compute overlap); two or more overlap
  if (fixed + 1u * (stage_bytes + 16u) + 128u > budget)
    return fail(CI_ERR_UNSUPPORTED,
                "problem too wide for the tile pipeline: p=%d needs %u B per stage", p,
                stage_bytes);
  uint32_t nst = (budget - fixed - 128u) / (stage_bytes + 16u);
  if (nst >= (uint32_t)NB) { nst = (uint32_t)NB; cfg.resident = 1; }
  else { cfg.resident = 0; if (nst > 8) nst = 8; }
  cfg.nstage = nst;
  uint32_t off = align_up(nst * stage_bytes, 128);
  cfg.off_full = off;   off += nst * 8;
  cfg.off_empty = off;  off += nst * 8 + 8;   // + the Omega barrier
  off = align_up(off, 16);
  cfg.off_omega = off;  off += omega_bytes;
  cfg.off_warp = off;   off += (uint32_t)G * cfg.warp_bytes;
  off = align_up(off, 16) + tail_bytes;
  cfg.total_bytes = off;
  *out = cfg;
  return CI_OK;
}

int pick_G(const ci_ctx* c, int C) {
  if (c->force_G > 0) return c->force_G > MAXG ? MAXG : c->force_G;
  int G = (C + c->sm_count - 1) / c->sm_count;
  if (G < 1) G = 1;
  if (G > MAXG) G = MAXG;
  return G;
}

// Team mode (ci_team.cuh): one warp per tile, W = NB warps per chain.  Used when
// the whole series is resident in shared memory and has 2..MAXW tiles.
// `sub` (may be NULL): out, warps per tile -- 2 when the sub-tile kernels (ci_team4.cuh, 4 steps
// per lane) apply: small p and at most MAXW/2 tiles; callers without sub-tile kernels pass NULL.
template <typename R>
bool plan_team(const ci_ctx* c, int C, int* GT, SmemCfg* cfg, int* sub = nullptr) {
  int W = c->NB;
  if (sub) {
    *sub = 1;
    if (c->team_sub && 2 * W <= MAXW && c->prob.p <= PSMALL) { *sub = 2; W *= 2; }
  }
  // (W = 1, a single tile, is a team of one: same kernel, no checkpoint replay between the
  // forward and the adjoint sweep)
  if (!c->team_mode || W < 1 || W > MAXW || c->prob.model != CI_MODEL_LOCAL_LEVEL) return false;
  int gt = (C >= 4 * c->sm_count) ? MAXW / W : 1;
  if (gt < 1) gt = 1;
  if (c->force_G > 0) gt = c->force_G * W <= MAXW ? c->force_G : 1;
  const uint32_t tail = (uint32_t)gt * (uint32_t)sizeof(TeamShared<R>);
  std::string keep = g_err;
  if (plan_smem(c, gt * W, 0, cfg, tail) != CI_OK || !cfg->resident) { g_err = keep; return false; }
  *GT = gt;
  return true;
}

template <typename R>
int launch_logpost(ci_ctx* c, const void* theta_d, int C, void* value_d, void* grad_d, int variant,
                   int flags, cudaStream_t st) {
  if (variant != CI_VARIANT_SCAN && variant != CI_VARIANT_SEQ)
    return fail(CI_ERR_INVALID, "unknown variant %d", variant);
  SmemCfg cfg;
  if (c->prob.model == CI_MODEL_LOCAL_LINEAR_TREND) {
    if (variant != CI_VARIANT_SCAN)
      return fail(CI_ERR_UNSUPPORTED, "the local linear trend model has only the scan variant");
    const int G = pick_G(c, C);
    int rc = plan_smem(c, G, 0, &cfg);
    if (rc) return rc;
    auto lk = k_logpost_llt<R>;
    CU_TRY(set_smem(lk, (uint32_t)cfg.total_bytes));
    lk<<<(C + G - 1) / G, 32 * (G + 1), cfg.total_bytes, st>>>(
        make_probdev<R>(c), make_lltdev<R>(c), cfg, static_cast<const R*>(theta_d), C,
        static_cast<R*>(value_d), static_cast<R*>(grad_d), flags);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return CI_OK;
  }
  if (variant == CI_VARIANT_SEQ) {
    const int G = pick_G(c, C);
    int rc = plan_smem(c, G, 2u * (uint32_t)c->NB * (uint32_t)GROUPS_PER_TILE, &cfg);
    if (rc) return rc;
    auto sk = k_logpost_seq<R>;
    CU_TRY(set_smem(sk, (uint32_t)cfg.total_bytes));
    sk<<<(C + G - 1) / G, 32 * (G + 1), cfg.total_bytes, st>>>(
        make_probdev<R>(c), cfg, static_cast<const R*>(theta_d), C, static_cast<R*>(value_d),
        static_cast<R*>(grad_d), flags);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return CI_OK;
  }
  int GT = 0, sub = 1;
  if (plan_team<R>(c, C, &GT, &cfg, &sub)) {
    const int W = c->NB * sub;
    auto tk = sub == 2 ? k_logpost_teamq<R, 4> : k_logpost_team<R>;
    CU_TRY(set_smem(tk, (uint32_t)cfg.total_bytes));
    tk<<<(C + GT - 1) / GT, 32 * (GT * W + 1), cfg.total_bytes, st>>>(
        make_probdev<R>(c), cfg, W, static_cast<const R*>(theta_d), C,
        static_cast<R*>(value_d), static_cast<R*>(grad_d), flags);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return CI_OK;
  }
  const int G = pick_G(c, C);
  int rc = plan_smem(c, G, 0, &cfg);
  if (rc) return rc;
  auto kern = k_logpost_scan<R>;
  CU_TRY(set_smem(kern, (uint32_t)cfg.total_bytes));
  const int grid = (C + G - 1) / G;
  kern<<<grid, 32 * (G + 1), cfg.total_bytes, st>>>(
      make_probdev<R>(c), cfg, static_cast<const R*>(theta_d), C, static_cast<R*>(value_d),
      static_cast<R*>(grad_d), flags);
  CU_TRY(cudaGetLastError());
  c->launches++;
  return CI_OK;
}

template <typename R>
void build_tiles(const ci_problem* pb, const void* y_, const void* X_, int NB, int ld,
                 std::vector<R>& out) {
  const R* y = static_cast<const R*>(y_);
  const R* X = static_cast<const R*>(X_);
  const int T = pb->T, p = pb->p;
  const size_t te = (size_t)tile_elems(p);
  out.assign((size_t)NB * te, (R)0);
  const R qnan = std::numeric_limits<R>::quiet_NaN();
  for (int b = 0; b < NB; ++b) {
    R* tile = out.data() + (size_t)b * te;
    for (int tl = 0; tl < TB; ++tl) {
      const int t = b * TB + tl;
      R* row = tile + tile_off(tl, ld);
      if (t < T) {
        for (int j = 0; j < p; ++j) row[j] = X[(size_t)t * p + j];
        row[p] = y[t];
      } else {
        row[p] = qnan;   // padded step == masked step
      }
    }
  }
}

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
  int GT = 0, sub = 1;
  if (plan_team<R>(c, C, &GT, &cfg, &sub)) {
    const int W = c->NB * sub;
    auto tk = sub == 2 ? k_hmc_teamq<R, 4> : k_hmc_team<R>;
    CU_TRY(set_smem(tk, (uint32_t)cfg.total_bytes));
    tk<<<(C + GT - 1) / GT, 32 * (GT * W + 1), cfg.total_bytes, st>>>(
        make_probdev<R>(c), cfg, W, plan, seed, chain_id0, static_cast<const R*>(theta0_d), C,
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

template <typename R>
int launch_gibbs(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0, int C,
                 void* draws_d, void* level_d, void* traj_d, float* incl_d, cudaStream_t st,
                 bool batch = false) {
  const int p = c->prob.p;
  const uint32_t extra = (uint32_t)(2 * p * p + 5 * p + 8);
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
  const double pi = o->nonzero_prob;
  plan.logit_pi = (plan.sparse && pi < 1.0) ? std::log(pi) - std::log1p(-pi) : 1e30;
  if (!(pi < 1.0)) plan.sparse = 0;
  GibbsDev<R> gd;
  gd.gram = static_cast<const R*>(c->v_gram); gd.xty0 = static_cast<const R*>(c->v_xty);
  gd.yty0 = (R)c->yty0;
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

template <typename R>
int launch_gibbs_seasonal(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0, int C,
                          void* draws_d, void* level_d, void* traj_d, float* incl_d, void* latent_d,
                          void* seas_d, void* drift_d, cudaStream_t st, bool batch = false) {
  const int p = c->prob.p, d = c->seas.d;
  const uint32_t base_extra = (uint32_t)(2 * p * p + 5 * p + 8) + (uint32_t)(d * (d | 1) + d) +
                              (3u + (uint32_t)c->seas.K) * (uint32_t)ci::TB;
  const uint32_t scr_elems = (uint32_t)c->prob.T * (uint32_t)(d + 1);
  const uint32_t tail = (uint32_t)(p * p) * (uint32_t)c->esz + 16u;
  SmemCfg cfg;
  int G = batch ? pick_G(c, C * c->batch_n) : pick_G(c, C), rc = CI_ERR_UNSUPPORTED;
  if (G > C) G = C;
  bool scr_smem = false;
  {  // first choice: the per-step scratch in shared memory with every tile resident
    std::string keep = g_err;
    for (int g = G; g >= 1 && rc != CI_OK; --g) {
      SmemCfg t;
      if (plan_smem(c, g, base_extra + scr_elems, &t, tail) == CI_OK && t.resident) {
        cfg = t; G = g; rc = CI_OK; scr_smem = true;
      }
    }
    g_err = keep;
  }
  for (; rc != CI_OK && G >= 1; --G) {
    rc = plan_smem(c, G, base_extra, &cfg, tail);
    if (rc == CI_OK) break;
  }
  if (rc) return rc;
  GibbsPlan plan;
  plan.n_warmup = o->n_warmup; plan.n_results = o->n_results; plan.sparse = o->sparse ? 1 : 0;
  plan.n_obs = c->n_obs; plan.chain_major = o->chain_major ? 1 : 0;
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
  auto kern = k_gibbs_seasonal<R>;
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

template <typename R>
int launch_predict(ci_ctx* c, const void* theta_d, int S, uint64_t seed, uint64_t draw_id0,
                   void* level_d, void* traj_d, void* mean_d, cudaStream_t st) {
  SmemCfg cfg;
  const ProbDev<R> prt = make_probdev<R>(c);
  int GT = 0;
  // The kernel choice must NOT depend on S: a draw has to come out bit-identical however
  // the batch is split over calls / GPUs.  The one-warp kernel is the default (better
  // occupancy at thousands of draws: 21.5 vs 20.0 M draws/s at S=4096, run 8); the team
  // kernel (lower latency for a handful of draws) is opt-in via CI_B200_PREDICT_TEAM=1.
  if (c->predict_team && plan_team<R>(c, S, &GT, &cfg)) {
    auto tk = k_predict_team<R>;
    CU_TRY(set_smem(tk, (uint32_t)cfg.total_bytes));
    tk<<<(S + GT - 1) / GT, 32 * (GT * c->NB + 1), cfg.total_bytes, st>>>(
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
  size_t bytes = (((size_t)S * sizeof(R)) + 15) & ~(size_t)15;
  const int in_smem = bytes + QSTATIC <= (size_t)c->smem_optin;
  if (!in_smem) bytes = 0;
  QuantArgs qa;
  qa.nq = nq;
  for (int i = 0; i < nq; ++i) qa.q[i] = q[i];
  auto kern = k_row_quantiles<R>;
  CU_TRY(set_smem(kern, (uint32_t)bytes));
  int nt = 1024;
  while (nt > 64 && nt / 2 >= S) nt >>= 1;
  kern<<<T, nt, bytes, st>>>(static_cast<const R*>(a_d), S, T, qa, static_cast<R*>(out_d),
                             out_ld > 0 ? out_ld : nq, in_smem);
  CU_TRY(cudaGetLastError());
  c->launches++;
  return CI_OK;
}

template <typename R>
int launch_impact(ci_ctx* c, const ImpactDev& a, const void* traj_d, const void* mean_d,
                  const double* obs_d, const uint8_t* period_d, double* series_d, double* summ_d,
                  cudaStream_t st) {
  const int S = a.S, T = a.T, Tc = T - a.t_c0;
  R* trT = static_cast<R*>(c->i_trT.p);
  double* cumT = static_cast<double*>(c->i_cum.p);
  double* statsT = static_cast<double*>(c->i_stats.p);
  const int row_ctas = (S + IMP_TILE - 1) / IMP_TILE + 1;            // + the predictive mean
  k_impact_rows<R><<<row_ctas, 32 * IMP_TILE, 0, st>>>(
      static_cast<const R*>(traj_d), static_cast<const R*>(mean_d), obs_d, period_d, a, trT, cumT,
      statsT, series_d, summ_d);
  CU_TRY(cudaGetLastError());
  c->launches++;
  size_t bytes = (((size_t)S * sizeof(double)) + 15) & ~(size_t)15;   // float64 jobs
  const int in_smem = bytes + QSTATIC <= (size_t)c->smem_optin;       // else: select from global memory
  if (!in_smem) bytes = 0;
  auto kern = k_impact_jobs<R>;
  CU_TRY(set_smem(kern, (uint32_t)bytes));
  int nt = 1024;
  while (nt > 64 && nt / 2 >= S) nt >>= 1;
  kern<<<Tc + IMP_STATS + T + 1, nt, bytes, st>>>(trT, cumT, statsT, obs_d, a, series_d, summ_d,
                                                   in_smem);
  CU_TRY(cudaGetLastError());
  c->launches++;
  return CI_OK;
}

}  // namespace

namespace {
template <typename R>
int upload_batch(ci_ctx* c, const ci_problem* probs, int N, const void* y_, const void* X_,
                 const void* Om_) {
  const int T = probs[0].T, p = probs[0].p;
  const size_t te = (size_t)ci::tile_elems(p);
  const size_t tile_stride = (size_t)c->NB * te;                        // elements per series
  const size_t om_stride = (((size_t)p * p * sizeof(R) + 15) & ~(size_t)15) / sizeof(R) + 16 / sizeof(R);
  std::vector<R> tiles(tile_stride * N), om(om_stride * N, (R)0), gram((size_t)p * p * N + 4),
      xty((size_t)(p > 0 ? p : 1) * N + 4);
  c->b_prob.assign(probs, probs + N);
  c->b_yty.assign(N, 0.0); c->b_nobs.assign(N, 0);
  const R* y = static_cast<const R*>(y_);
  const R* X = static_cast<const R*>(X_);
  const R* Om = static_cast<const R*>(Om_);
  std::vector<R> one;
  for (int s = 0; s < N; ++s) {
    build_tiles<R>(&probs[s], y + (size_t)s * T, p ? X + (size_t)s * T * p : nullptr, c->NB, c->ld, one);
    std::copy(one.begin(), one.end(), tiles.begin() + tile_stride * s);
    if (p) std::copy(Om + (size_t)s * p * p, Om + (size_t)(s + 1) * p * p, om.begin() + om_stride * s);
    // sufficient statistics over observed rows, float64 on the host (as ci_set_data)
    std::vector<double> g((size_t)p * p, 0.0), b((size_t)(p > 0 ? p : 1), 0.0);
    double yty = 0.0; int nobs = 0;
    for (int t = 0; t < T; ++t) {
      const double yt = (double)y[(size_t)s * T + t];
      if (!(yt == yt)) continue;
      ++nobs; yty += yt * yt;
      const R* xr = X + ((size_t)s * T + t) * p;
      for (int i = 0; i < p; ++i) {
        b[i] += (double)xr[i] * yt;
        for (int j = 0; j <= i; ++j) g[(size_t)i * p + j] += (double)xr[i] * (double)xr[j];
      }
    }
    for (int i = 0; i < p; ++i)
      for (int j = i + 1; j < p; ++j) g[(size_t)i * p + j] = g[(size_t)j * p + i];
    for (int i = 0; i < p * p; ++i) gram[(size_t)s * p * p + i] = (R)g[i];
    for (int i = 0; i < p; ++i) xty[(size_t)s * p + i] = (R)b[i];
    c->b_yty[s] = yty; c->b_nobs[s] = nobs;
  }
  c->b_tile_stride = tile_stride * sizeof(R); c->b_omega_stride = om_stride * sizeof(R);
  c->b_gram_stride = (size_t)p * p * sizeof(R); c->b_xty_stride = (size_t)p * sizeof(R);
  CU_TRY(c->b_tiles.reserve(tiles.size() * sizeof(R)));
  CU_TRY(c->b_omega.reserve(om.size() * sizeof(R)));
  CU_TRY(c->b_gram.reserve(gram.size() * sizeof(R)));
  CU_TRY(c->b_xty.reserve(xty.size() * sizeof(R)));
  CU_TRY(cudaMemcpyAsync(c->b_tiles.p, tiles.data(), tiles.size() * sizeof(R), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemcpyAsync(c->b_omega.p, om.data(), om.size() * sizeof(R), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemcpyAsync(c->b_gram.p, gram.data(), gram.size() * sizeof(R), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemcpyAsync(c->b_xty.p, xty.data(), xty.size() * sizeof(R), cudaMemcpyHostToDevice, c->stream));
  // per-series device descriptors of the batched kernels
  std::vector<ci::BatchDev<R>> dev(N);
  for (int s = 0; s < N; ++s) {
    c->prob = probs[s];
    c->v_tiles = static_cast<char*>(c->b_tiles.p) + c->b_tile_stride * s;
    c->v_omega = static_cast<char*>(c->b_omega.p) + c->b_omega_stride * s;
    dev[s].pr = make_probdev<R>(c);
    dev[s].gd.gram = reinterpret_cast<const R*>(static_cast<char*>(c->b_gram.p) + c->b_gram_stride * s);
    dev[s].gd.xty0 = reinterpret_cast<const R*>(static_cast<char*>(c->b_xty.p) + c->b_xty_stride * s);
    dev[s].gd.yty0 = (R)c->b_yty[s];
    dev[s].n_obs = c->b_nobs[s];
  }
  CU_TRY(c->b_dev.reserve(dev.size() * sizeof(ci::BatchDev<R>)));
  CU_TRY(cudaMemcpyAsync(c->b_dev.p, dev.data(), dev.size() * sizeof(ci::BatchDev<R>),
                         cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return CI_OK;
}
}  // namespace

// ===========================================================================
extern "C" {

int ci_version(void) { return CI_B200_VERSION; }
const char* ci_last_error(void) { return g_err.c_str(); }

int ci_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { fail(CI_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); return CI_ERR_NO_DEVICE; }
  return n;
}

int ci_ctx_create(int device, ci_ctx** out) {
  if (!out) return fail(CI_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(CI_ERR_NO_DEVICE, "no CUDA device (%s); this engine has no CPU fallback",
                e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(CI_ERR_INVALID, "device %d out of range [0,%d)", device, n);
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(CI_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  ci_ctx* c = new ci_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = (int)prop.sharedMemPerBlockOptin;
  cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (se != cudaSuccess) { delete c; return fail(CI_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(se)); }
  if (const char* g = getenv("CI_B200_G")) c->force_G = atoi(g);
  if (const char* g = getenv("CI_B200_TEAM")) c->team_mode = atoi(g);
  if (const char* g = getenv("CI_B200_PREDICT_TEAM")) c->predict_team = atoi(g);
  if (const char* g = getenv("CI_B200_TEAM_SUB")) c->team_sub = atoi(g);
  *out = c;
  return CI_OK;
}

int ci_ctx_destroy(ci_ctx* c) {
  if (!c) return CI_OK;
  cudaSetDevice(c->device);
  if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
  c->tiles.release(); c->omega.release();
  c->w_theta.release(); c->w_value.release(); c->w_grad.release();
  c->w_level.release(); c->w_traj.release(); c->w_mean.release(); c->w_q.release(); c->w_draws.release(); c->w_stats.release(); c->w_incl.release();
  c->gram.release(); c->xty0.release();
  c->b_tiles.release(); c->b_omega.release(); c->b_gram.release(); c->b_xty.release(); c->b_dev.release();
  c->s_sched.release(); c->s_scratch.release(); c->s_series.release(); c->w_latent.release(); c->w_seas.release(); c->w_drift.release();
  c->i_trT.release(); c->i_cum.release(); c->i_stats.release(); c->i_meta.release(); c->i_series.release(); c->i_summ.release();
  delete c;
  return CI_OK;
}

int64_t ci_launch_count(const ci_ctx* c) { return c ? c->launches : 0; }

int ci_set_data(ci_ctx* c, const ci_problem* pb, const void* y, const void* X, const void* Omega) {
  if (!c || !pb || !y) return fail(CI_ERR_INVALID, "null argument");
  if (pb->T < 1) return fail(CI_ERR_INVALID, "T must be >= 1 (got %d)", pb->T);
  if (pb->p < 0) return fail(CI_ERR_INVALID, "p must be >= 0 (got %d)", pb->p);
  if (pb->p > 0 && (!X || !Omega)) return fail(CI_ERR_INVALID, "X and Omega are required when p > 0");
  if (pb->dtype != CI_F32 && pb->dtype != CI_F64) return fail(CI_ERR_INVALID, "dtype must be 0 or 1");
  if (pb->model != CI_MODEL_LOCAL_LEVEL && pb->model != CI_MODEL_LOCAL_LINEAR_TREND)
    return fail(CI_ERR_INVALID, "unknown model %d", pb->model);
  if (pb->model == CI_MODEL_LOCAL_LINEAR_TREND && !(pb->P0_slope > 0))
    return fail(CI_ERR_INVALID, "P0_slope must be positive");
  const int d = pb->model == CI_MODEL_LOCAL_LINEAR_TREND ? 2 : 1;
  if (pb->p + 1 + d > ci::MAX_DIM)
    return fail(CI_ERR_UNSUPPORTED, "p=%d exceeds the supported maximum %d", pb->p, ci::MAX_DIM - 1 - d);
  if (!(pb->P0 > 0)) return fail(CI_ERR_INVALID, "P0 must be positive");
  CU_TRY(cudaSetDevice(c->device));
  c->has_data = false;
  c->seas = ci::SeasDev{};
  c->prob = *pb;
  c->esz = pb->dtype == CI_F64 ? 8 : 4;
  c->NB = (pb->T + ci::TB - 1) / ci::TB;
  c->ld = ci::tile_ld(pb->p);
  c->dim = pb->p + 1 + d;
  const size_t te = (size_t)ci::tile_elems(pb->p);
  const size_t tile_bytes = (size_t)c->NB * te * c->esz;
  CU_TRY(c->tiles.reserve(tile_bytes));
  if (pb->dtype == CI_F64) {
    std::vector<double> h; build_tiles<double>(pb, y, X, c->NB, c->ld, h);
    CU_TRY(cudaMemcpyAsync(c->tiles.p, h.data(), tile_bytes, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
  } else {
    std::vector<float> h; build_tiles<float>(pb, y, X, c->NB, c->ld, h);
    CU_TRY(cudaMemcpyAsync(c->tiles.p, h.data(), tile_bytes, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
  }
  const size_t ob = (size_t)pb->p * pb->p * c->esz;
  CU_TRY(c->omega.reserve(ob + 16));          // bulk copies move whole 16-byte units
  if (ob) {
    CU_TRY(cudaMemcpyAsync(c->omega.p, Omega, ob, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
  }
  {  // sufficient statistics of the regression step over OBSERVED rows (float64 on the host)
    const int T = pb->T, p = pb->p;
    std::vector<double> gram((size_t)p * p, 0.0), xty((size_t)(p > 0 ? p : 1), 0.0);
    double yty = 0.0; int nobs = 0;
    for (int t = 0; t < T; ++t) {
      const double yt = pb->dtype == CI_F64 ? static_cast<const double*>(y)[t]
                                            : (double)static_cast<const float*>(y)[t];
      if (!(yt == yt)) continue;
      ++nobs; yty += yt * yt;
      for (int i = 0; i < p; ++i) {
        const double xi = pb->dtype == CI_F64 ? static_cast<const double*>(X)[(size_t)t * p + i]
                                              : (double)static_cast<const float*>(X)[(size_t)t * p + i];
        xty[i] += xi * yt;
        for (int j = 0; j <= i; ++j) {
          const double xj = pb->dtype == CI_F64 ? static_cast<const double*>(X)[(size_t)t * p + j]
                                                : (double)static_cast<const float*>(X)[(size_t)t * p + j];
          gram[(size_t)i * p + j] += xi * xj;
        }
      }
    }
    for (int i = 0; i < p; ++i)
      for (int j = i + 1; j < p; ++j) gram[(size_t)i * p + j] = gram[(size_t)j * p + i];
    c->yty0 = yty; c->n_obs = nobs;
    CU_TRY(c->gram.reserve((size_t)p * p * c->esz + 16));
    CU_TRY(c->xty0.reserve((size_t)p * c->esz + 16));
    if (p > 0) {
      if (pb->dtype == CI_F64) {
        CU_TRY(cudaMemcpyAsync(c->gram.p, gram.data(), (size_t)p * p * 8, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaMemcpyAsync(c->xty0.p, xty.data(), (size_t)p * 8, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
      } else {
        std::vector<float> gf(gram.begin(), gram.end()), xf(xty.begin(), xty.end());
        CU_TRY(cudaMemcpyAsync(c->gram.p, gf.data(), (size_t)p * p * 4, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaMemcpyAsync(c->xty0.p, xf.data(), (size_t)p * 4, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
      }
    }
  }
  c->v_tiles = c->tiles.p; c->v_omega = c->omega.p; c->v_gram = c->gram.p; c->v_xty = c->xty0.p;
  c->batch_n = 0;
  // validate that the pipeline fits before accepting the problem
  ci::SmemCfg cfg;
  int rc = plan_smem(c, 1, 0, &cfg);
  if (rc) return rc;
  c->has_data = true;
  return CI_OK;
}

// ---- batches of independent series -------------------------------------
int ci_batch_select(ci_ctx* c, int s) {
  if (!c) return fail(CI_ERR_INVALID, "null argument");
  if (c->batch_n < 1) return fail(CI_ERR_STATE, "ci_set_data_batch has not been called");
  if (s < 0 || s >= c->batch_n) return fail(CI_ERR_INVALID, "series %d out of range [0,%d)", s, c->batch_n);
  c->prob = c->b_prob[s];
  c->v_tiles = static_cast<char*>(c->b_tiles.p) + c->b_tile_stride * s;
  c->v_omega = static_cast<char*>(c->b_omega.p) + c->b_omega_stride * s;
  c->v_gram = static_cast<char*>(c->b_gram.p) + c->b_gram_stride * s;
  c->v_xty = static_cast<char*>(c->b_xty.p) + c->b_xty_stride * s;
  c->yty0 = c->b_yty[s]; c->n_obs = c->b_nobs[s];
  // (a season calendar set with ci_set_seasonal after ci_set_data_batch belongs to the whole
  // panel -- same T for every series -- and survives the selection)
  c->has_data = true;
  return CI_OK;
}

int ci_set_data_batch(ci_ctx* c, const ci_problem* probs, int N, const void* y, const void* X,
                      const void* Omega) {
  if (!c || !probs || !y) return fail(CI_ERR_INVALID, "null argument");
  if (N < 1) return fail(CI_ERR_INVALID, "n_series must be >= 1");
  const ci_problem& p0 = probs[0];
  if (p0.T < 1 || p0.p < 0) return fail(CI_ERR_INVALID, "bad T / p");
  if (p0.p > 0 && (!X || !Omega)) return fail(CI_ERR_INVALID, "X and Omega are required when p > 0");
  if (p0.dtype != CI_F32 && p0.dtype != CI_F64) return fail(CI_ERR_INVALID, "dtype must be 0 or 1");
  if (p0.model != CI_MODEL_LOCAL_LEVEL)
    return fail(CI_ERR_UNSUPPORTED, "batches are local-level only (as the reference's model)");
  if (p0.p + 2 > ci::MAX_DIM) return fail(CI_ERR_UNSUPPORTED, "p=%d exceeds the supported maximum", p0.p);
  for (int s = 0; s < N; ++s) {
    if (probs[s].T != p0.T || probs[s].p != p0.p || probs[s].dtype != p0.dtype || probs[s].model != p0.model)
      return fail(CI_ERR_INVALID, "series %d differs in T / p / dtype / model: a batch shares its shape", s);
    if (!(probs[s].P0 > 0)) return fail(CI_ERR_INVALID, "P0 must be positive (series %d)", s);
  }
  CU_TRY(cudaSetDevice(c->device));
  c->has_data = false;
  c->seas = ci::SeasDev{};
  c->prob = p0;
  c->esz = p0.dtype == CI_F64 ? 8 : 4;
  c->NB = (p0.T + ci::TB - 1) / ci::TB;
  c->ld = ci::tile_ld(p0.p);
  c->dim = p0.p + 2;
  int rc = p0.dtype == CI_F64 ? upload_batch<double>(c, probs, N, y, X, Omega)
                              : upload_batch<float>(c, probs, N, y, X, Omega);
  if (rc) return rc;
  c->batch_n = N;
  rc = ci_batch_select(c, 0);
  if (rc) return rc;
  ci::SmemCfg cfg;
  rc = plan_smem(c, 1, 0, &cfg);
  if (rc) { c->has_data = false; c->batch_n = 0; return rc; }
  return CI_OK;
}

int ci_set_seasonal_batch(ci_ctx* c, const ci_seasonal* sp, const double* init_sd,
                          const double* drift_scale, const double* drift_ub) {
  if (!c || !sp || !init_sd || !drift_scale || !drift_ub) return fail(CI_ERR_INVALID, "null argument");
  if (c->batch_n < 1) return fail(CI_ERR_STATE, "ci_set_data_batch has not been called");
  int rc = ci_set_seasonal(c, sp);
  if (rc) return rc;
  const int N = c->batch_n;
  std::vector<double> h((size_t)3 * N);
  for (int s = 0; s < N; ++s) {
    if (!(init_sd[s] > 0) || !(drift_scale[s] > 0) || !(drift_ub[s] > 0)) {
      c->seas = ci::SeasDev{};
      return fail(CI_ERR_INVALID, "init_sd, drift_scale, drift_ub must be positive (series %d)", s);
    }
    h[3 * s] = init_sd[s] * init_sd[s]; h[3 * s + 1] = drift_scale[s]; h[3 * s + 2] = drift_ub[s];
  }
  CU_TRY(c->s_series.reserve(h.size() * sizeof(double)));
  CU_TRY(cudaMemcpyAsync(c->s_series.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  c->seas.per_series = static_cast<const double*>(c->s_series.p);
  return CI_OK;
}

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

int ci_gibbs_run_batch_d(ci_ctx* c, const ci_gibbs_opts* o, uint64_t seed, uint64_t chain_id0,
                         int Cs, void* draws_d, void* level_d, void* traj_d, float* incl_d,
                         void* stream) {
  if (!c || !o || !draws_d) return fail(CI_ERR_INVALID, "null argument");
  if (c->batch_n < 1) return fail(CI_ERR_STATE, "ci_set_data_batch has not been called");
  if (Cs < 1 || o->n_results < 1 || o->n_warmup < 0)
    return fail(CI_ERR_INVALID, "n_chains >= 1, n_results >= 1, n_warmup >= 0 required");
  if (o->sparse && !(o->nonzero_prob > 0.0 && o->nonzero_prob <= 1.0))
    return fail(CI_ERR_INVALID, "nonzero_prob must be in (0, 1]");
  for (int s = 0; s < c->batch_n; ++s)
    if (c->b_nobs[s] < 2) return fail(CI_ERR_INVALID, "series %d has fewer than 2 observed points", s);
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->prob.dtype == CI_F64)
    return launch_gibbs<double>(c, o, seed, chain_id0, Cs, draws_d, level_d, traj_d, incl_d, st, true);
  return launch_gibbs<float>(c, o, seed, chain_id0, Cs, draws_d, level_d, traj_d, incl_d, st, true);
}

int ci_logprob_grad_d(ci_ctx* c, const void* theta_d, int C, void* value_d, void* grad_d,
                      int variant, int flags, void* stream) {
  if (!c || !theta_d || !value_d) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  if (C < 1) return fail(CI_ERR_INVALID, "n_chains must be >= 1");
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->prob.dtype == CI_F64)
    return launch_logpost<double>(c, theta_d, C, value_d, grad_d, variant, flags, st);
  return launch_logpost<float>(c, theta_d, C, value_d, grad_d, variant, flags, st);
}

// Device-visible alias of a PINNED host buffer (cudaHostAlloc / cudaHostRegister),
// or nullptr for pageable memory.
static void* pinned_alias(const void* host) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
  return at.devicePointer;
}

int ci_logprob_grad(ci_ctx* c, const void* theta, int C, void* value, void* grad, int variant,
                    int flags) {
  if (!c || !theta || !value) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  if (C < 1) return fail(CI_ERR_INVALID, "n_chains must be >= 1");
  CU_TRY(cudaSetDevice(c->device));
  const size_t tb = (size_t)C * c->dim * c->esz, vb = (size_t)C * c->esz;
  // Pinned caller buffers: the kernel reads theta and writes value / grad straight
  // over PCIe (zero-copy) -- one launch + one sync instead of three staged copies.
  void* th_a = pinned_alias(theta);
  void* va_a = th_a ? pinned_alias(value) : nullptr;
  void* gr_a = (va_a && grad) ? pinned_alias(grad) : nullptr;
  if (th_a && va_a && (!grad || gr_a) && tb <= (1u << 20)) {
    int rc = ci_logprob_grad_d(c, th_a, C, va_a, gr_a, variant, flags, c->stream);
    if (rc) return rc;
    CU_TRY(cudaStreamSynchronize(c->stream));
    return CI_OK;
  }
  CU_TRY(c->w_theta.reserve(tb));
  CU_TRY(c->w_value.reserve(vb));
  if (grad) CU_TRY(c->w_grad.reserve(tb));
  CU_TRY(cudaMemcpyAsync(c->w_theta.p, theta, tb, cudaMemcpyHostToDevice, c->stream));
  int rc = ci_logprob_grad_d(c, c->w_theta.p, C, c->w_value.p, grad ? c->w_grad.p : nullptr,
                             variant, flags, c->stream);
  if (rc) return rc;
  CU_TRY(cudaMemcpyAsync(value, c->w_value.p, vb, cudaMemcpyDeviceToHost, c->stream));
  if (grad) CU_TRY(cudaMemcpyAsync(grad, c->w_grad.p, tb, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return CI_OK;
}

int ci_logprob(ci_ctx* c, const void* theta, int C, void* value, int variant, int flags) {
  return ci_logprob_grad(c, theta, C, value, nullptr, variant, flags);
}


int ci_posterior_predict_d(ci_ctx* c, const void* theta_d, int S, uint64_t seed, uint64_t draw_id0,
                           void* level_d, void* traj_d, void* mean_d, void* stream) {
  if (!c || !theta_d || !traj_d) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  if (S < 1) return fail(CI_ERR_INVALID, "S must be >= 1");
  if (c->prob.model != CI_MODEL_LOCAL_LEVEL)
    return fail(CI_ERR_UNSUPPORTED, "ci_posterior_predict: local level only (the reference has "
                "no slope component, causalimpact_lib.py:496)");
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

int ci_set_seasonal(ci_ctx* c, const ci_seasonal* sp) {
  if (!c) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  c->seas = ci::SeasDev{};
  if (!sp || sp->n_components == 0) return CI_OK;
  const int K = sp->n_components, T = c->prob.T;
  if (K < 0 || K > CI_MAX_SEASONAL)
    return fail(CI_ERR_UNSUPPORTED, "at most %d seasonal components (got %d)", CI_MAX_SEASONAL, K);
  if (c->prob.model != CI_MODEL_LOCAL_LEVEL)
    return fail(CI_ERR_UNSUPPORTED, "seasonal components need the local level model");
  if (!sp->active || !sp->ends) return fail(CI_ERR_INVALID, "null schedule");
  if (!(sp->init_sd > 0) || !(sp->drift_conc > 0) || !(sp->drift_scale > 0) || !(sp->drift_ub > 0))
    return fail(CI_ERR_INVALID, "init_sd, drift_conc, drift_scale, drift_ub must be positive");
  ci::SeasDev sz{};
  sz.K = K;
  int d = 1;
  for (int k = 0; k < K; ++k) {
    if (sp->num_seasons[k] < 2) return fail(CI_ERR_INVALID, "num_seasons must be >= 2");
    sz.n[k] = sp->num_seasons[k]; sz.off[k] = d; d += sz.n[k];
  }
  if (d > ci::SEAS_MAXD)
    return fail(CI_ERR_UNSUPPORTED, "1 + sum(num_seasons) = %d exceeds the supported state "
                "dimension %d", d, ci::SEAS_MAXD);
  sz.d = d;
  std::vector<uint8_t> sched((size_t)T * (K + 1));
  for (int t = 0; t < T; ++t) {
    uint8_t em = 0;
    for (int k = 0; k < K; ++k) {
      const uint8_t a = sp->active[(size_t)k * T + t];
      if (a >= sz.n[k]) return fail(CI_ERR_INVALID, "active[%d][%d] = %d out of range", k, t, (int)a);
      sched[(size_t)t * (K + 1) + k] = a;
      if (sp->ends[(size_t)k * T + t]) { em |= (uint8_t)(1u << k); if (t < T - 1) sz.n_ends[k]++; }
    }
    sched[(size_t)t * (K + 1) + K] = em;
  }
  sz.init_var = sp->init_sd * sp->init_sd;
  sz.drift_conc = sp->drift_conc; sz.drift_scale = sp->drift_scale; sz.drift_ub = sp->drift_ub;
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(c->s_sched.reserve(sched.size()));
  CU_TRY(cudaMemcpyAsync(c->s_sched.p, sched.data(), sched.size(), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  sz.sched = static_cast<const uint8_t*>(c->s_sched.p);
  c->seas = sz;
  return CI_OK;
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

int ci_predictive_mean_d(ci_ctx* c, const void* theta_d, const void* level_d, int S, void* mean_d,
                         void* stream) {
  if (!c || !theta_d || !level_d || !mean_d) return fail(CI_ERR_INVALID, "null argument");
  if (!c->has_data) return fail(CI_ERR_STATE, "ci_set_data has not been called");
  if (S < 1) return fail(CI_ERR_INVALID, "S must be >= 1");
  if (c->prob.model != CI_MODEL_LOCAL_LEVEL)
    return fail(CI_ERR_UNSUPPORTED, "ci_predictive_mean: local level only");
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 blk(ci::MEAN_COLS, ci::MEAN_ROWS);
  const int grid = (c->prob.T + ci::MEAN_COLS - 1) / ci::MEAN_COLS;
  if (c->prob.dtype == CI_F64)
    ci::k_predict_mean<double><<<grid, blk, 0, st>>>(
        make_probdev<double>(c), static_cast<const double*>(theta_d),
        static_cast<const double*>(level_d), S, static_cast<double*>(mean_d));
  else
    ci::k_predict_mean<float><<<grid, blk, 0, st>>>(
        make_probdev<float>(c), static_cast<const float*>(theta_d),
        static_cast<const float*>(level_d), S, static_cast<float*>(mean_d));
  CU_TRY(cudaGetLastError());
  c->launches++;
  return CI_OK;
}

int ci_impact_d(ci_ctx* c, const ci_impact_args* a, const void* traj_d, const void* mean_d,
                const double* observed, const uint8_t* period, double* series_d, double* summ_d,
                void* stream) {
  if (!c || !a || !traj_d || !mean_d || !observed || !period || !series_d || !summ_d)
    return fail(CI_ERR_INVALID, "null argument");
  if (a->S < 1 || a->T < 1) return fail(CI_ERR_INVALID, "S and T must be >= 1");
  if (a->dtype != CI_F32 && a->dtype != CI_F64) return fail(CI_ERR_INVALID, "dtype must be 0 or 1");
  if (!(a->q_lo >= 0.0 && a->q_lo <= 1.0 && a->q_hi >= 0.0 && a->q_hi <= 1.0))
    return fail(CI_ERR_INVALID, "quantiles must be in [0,1]");
  if (!(a->scale > 0.0)) return fail(CI_ERR_INVALID, "scale must be positive");
  const int S = a->S, T = a->T;
  ci::ImpactDev d{};
  d.S = S; d.T = T; d.scale = a->scale; d.offset = a->offset; d.q_lo = a->q_lo; d.q_hi = a->q_hi;
  d.obs_sum = a->obs_sum;
  d.t_c0 = T; d.n_post = 0;
  for (int t = 0; t < T; ++t) {
    if (period[t] > 2 || (t > 0 && period[t] < period[t - 1]))
      return fail(CI_ERR_INVALID, "period[] must be non-decreasing values in {0,1,2}");
    if (period[t] != 0 && d.t_c0 == T) d.t_c0 = t;
    d.n_post += period[t] == 1;
  }
  if (d.n_post < 1) return fail(CI_ERR_INVALID, "the post-period is empty");
  CU_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Tc = T - d.t_c0;
  CU_TRY(c->i_cum.reserve((size_t)S * (Tc > 0 ? Tc : 1) * sizeof(double)));
  CU_TRY(c->i_stats.reserve((size_t)S * ci::IMP_STATS * sizeof(double)));
  CU_TRY(c->i_trT.reserve((size_t)S * T * (a->dtype == CI_F64 ? 8 : 4)));
  const size_t ob = (size_t)T * sizeof(double);
  CU_TRY(c->i_meta.reserve(ob + (size_t)T));
  CU_TRY(cudaMemcpyAsync(c->i_meta.p, observed, ob, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(static_cast<char*>(c->i_meta.p) + ob, period, (size_t)T,
                         cudaMemcpyHostToDevice, st));
  const double* obs_d = static_cast<const double*>(c->i_meta.p);
  const uint8_t* per_d = reinterpret_cast<const uint8_t*>(static_cast<char*>(c->i_meta.p) + ob);
  if (a->dtype == CI_F64)
    return launch_impact<double>(c, d, traj_d, mean_d, obs_d, per_d, series_d, summ_d, st);
  return launch_impact<float>(c, d, traj_d, mean_d, obs_d, per_d, series_d, summ_d, st);
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

#ifdef CI_CLK
// developer build only (-DCI_CLK): phase clocks of the last k_logpost_team launch
int ci_debug_clocks(long long* out32) {
  CU_TRY(cudaDeviceSynchronize());
  CU_TRY(cudaMemcpyFromSymbol(out32, ci::g_clk, sizeof(long long) * 32));
  return CI_OK;
}
#endif

}  // extern "C"
