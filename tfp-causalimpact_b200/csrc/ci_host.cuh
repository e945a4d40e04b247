// ci_host.cuh -- host-side state shared by the translation units of the C ABI
// (include/ci_b200.h): the context, error reporting, workspaces and the shared-memory
// planner.  Every abi_*.cu includes this; kernels are instantiated only in the unit that
// launches them, so the units compile in parallel (causalimpact_b200/_build.py).
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/ci_b200.h"
#include "ci_kernels.cuh"
#include "ci_team.cuh"
#include "ci_team_stream.cuh"
#include "ci_llt.cuh"

// thread-local message of the last failure (defined in abi_core.cu; ONE instance for the library)
std::string& cih_err();

namespace {

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  cih_err() = buf;
  return code;
}

#define CU_TRY(expr)                                                                   \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return fail(CI_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                  __FILE__, __LINE__);                                                 \
  } while (0)

// A grow-only device workspace.  Growing NEVER frees inside an entry point: cudaFree is a
// device-wide synchronisation, which the _d entry points promise not to do.  The outgrown
// block is parked on the context's retired list and released by ci_ctx_destroy /
// ci_set_data (both synchronising calls); capacities grow geometrically so a workspace is
// re-allocated O(log size) times over its life.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  std::vector<void*>* retired = nullptr;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    size_t want = cap + cap / 2;
    if (want < bytes) want = bytes;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, want);
    if (e != cudaSuccess && want > bytes) { cudaGetLastError(); want = bytes; e = cudaMalloc(&q, want); }
    if (e != cudaSuccess) return e;
    if (p) { if (retired) retired->push_back(p); else cudaFree(p); }
    p = q; cap = want;
    return cudaSuccess;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Host -> device staging for the small per-call arrays of the _d entry points (observed / period of
// ci_impact_d).  cudaMemcpyAsync from PAGEABLE memory first waits for the stream's earlier work;
// copying through a ring of pinned slots keeps the call asynchronous: a slot is reused only after
// the copy that read it has completed (its event), which blocks only when RING calls are in flight.
struct PinnedRing {
  static constexpr int RING = 8;
  struct Slot { void* p = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; bool used = false; };
  Slot slots[RING];
  int next = 0;
  cudaError_t upload(void* dst_d, const void* const* srcs, const size_t* sizes, const size_t* offs,
                     int n, size_t total, cudaStream_t st) {
    Slot& s = slots[next];
    next = (next + 1) % RING;
    cudaError_t e;
    if (s.used) { e = cudaEventSynchronize(s.ev); if (e != cudaSuccess) return e; }
    if (s.cap < total) {
      if (s.p) cudaFreeHost(s.p);
      s.p = nullptr; s.cap = 0;
      size_t want = total < 4096 ? 4096 : total + total / 2;
      e = cudaHostAlloc(&s.p, want, cudaHostAllocDefault);
      if (e != cudaSuccess) return e;
      s.cap = want;
    }
    if (!s.ev) { e = cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming); if (e != cudaSuccess) return e; }
    for (int i = 0; i < n; ++i) memcpy(static_cast<char*>(s.p) + offs[i], srcs[i], sizes[i]);
    e = cudaMemcpyAsync(dst_d, s.p, total, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    s.used = true;
    return cudaEventRecord(s.ev, st);
  }
  void release() {
    for (Slot& s : slots) {
      if (s.ev) { cudaEventSynchronize(s.ev); cudaEventDestroy(s.ev); }
      if (s.p) cudaFreeHost(s.p);
      s = Slot{};
    }
  }
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
  DevBuf w_raw, w_pstats;            // ci_set_panel: the raw panel and the per-series statistics
  DevBuf w_mpart;                    // ci_predictive_mean_d: partial column / weight sums
  DevBuf x_pack, x_recvT, x_recvC, x_allT, x_allC, x_parts;   // ci_impact_sharded_d exchange buffers
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
  PinnedRing ring;                   // pinned staging of the _d entry points' small host arrays
  std::vector<void*> retired;        // outgrown workspaces, freed by the next synchronising call
  std::vector<DevBuf*> bufs() {
    return {&tiles, &omega, &w_theta, &w_value, &w_grad, &w_level, &w_traj, &w_mean, &w_q, &w_draws,
            &w_stats, &w_incl, &gram, &xty0, &i_cum, &i_stats, &i_meta, &i_series, &i_summ, &i_trT,
            &s_sched, &s_scratch, &s_series, &w_latent, &w_seas, &w_drift, &b_tiles, &b_omega,
            &b_gram, &b_xty, &b_dev, &w_raw, &w_pstats, &x_pack, &x_recvT, &x_recvC, &x_allT, &x_allC,
            &x_parts, &w_mpart};
  }
  // only from entry points that synchronise anyway (ci_set_data*, ci_ctx_destroy, host-pointer calls)
  void free_retired() {
    for (void* q : retired) cudaFree(q);
    retired.clear();
  }
  int force_G = 0;                   // CI_B200_G env override (tuning)
  int team_mode = 1;                 // CI_B200_TEAM=0 disables the warp-team kernels
  int predict_team = 1;              // CI_B200_PREDICT_TEAM=0: one warp per draw in ci_posterior_predict
  int sel_nt = 0, sel_smem = -1;     // CI_B200_SEL_NT / CI_B200_SEL_SMEM: select-kernel launch shape (tuning)
  int gibbs_team = 1;                // CI_B200_GIBBS_TEAM=0: one warp per chain in the Gibbs kernel
  int tstream_mode = 1;              // CI_B200_TSTREAM=0 disables the long-series team kernels
  int tstream_W = 0;                 // CI_B200_TSW: warps per chain of the long-series team kernels (tuning)
};

// abi_impact.cu: argument checks, period scan and the asynchronous upload of observed / period
// (shared by ci_impact_d, ci_impact_rows_d, ci_impact_cols_d and ci_impact_sharded_d)
namespace ci { struct ImpactDev; }
int cih_impact_prepare(ci_ctx* c, const ci_impact_args* a, const double* observed,
                       const uint8_t* period, cudaStream_t st, ci::ImpactDev* d,
                       const double** obs_d, const uint8_t** per_d);

namespace {

using namespace ci;

// the kernels compare VARIANCES with their bound: a bound on the scale is squared here
// (ci_problem.ub_on_scale, include/ci_b200.h)
inline double ub_variance(double ub, int on_scale) {
  if (!on_scale) return ub;
  return ub > 1e150 ? ub : ub * ub;
}

template <typename R> ProbDev<R> make_probdev(const ci_ctx* c) {
  ProbDev<R> pr;
  pr.tiles = static_cast<const R*>(c->v_tiles);
  pr.omega = static_cast<const R*>(c->v_omega);
  pr.T = c->prob.T; pr.p = c->prob.p; pr.ld = c->ld; pr.NB = c->NB; pr.dim = c->dim;
  pr.model = c->prob.model;
  pr.m0 = (R)c->prob.m0; pr.P0 = (R)c->prob.P0;
  pr.obs_conc = (R)c->prob.obs_conc; pr.obs_scale = (R)c->prob.obs_scale;
  pr.obs_ub = (R)ub_variance(c->prob.obs_ub, c->prob.ub_on_scale);
  pr.lvl_conc = (R)c->prob.lvl_conc; pr.lvl_scale = (R)c->prob.lvl_scale;
  pr.lvl_ub = (R)ub_variance(c->prob.lvl_ub, c->prob.ub_on_scale);
  return pr;
}

template <typename R> LltDev<R> make_lltdev(const ci_ctx* c) {
  LltDev<R> d;
  d.q_conc = (R)c->prob.slope_conc; d.q_scale = (R)c->prob.slope_scale;
  const double qv = ub_variance(c->prob.slope_ub, c->prob.ub_on_scale);
  d.q_ub = (R)(qv > 1e30 ? 1e30 : qv);
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
// `ckpt_elems` < 0: the per-tile checkpoints of the one-warp kernels (6 NB); `max_stages`:
// ring depth when the series is streamed.
int plan_smem(const ci_ctx* c, int G, uint32_t extra_elems, SmemCfg* out,
              uint32_t tail_bytes = 0, int ckpt_elems = -1, uint32_t max_stages = 8) {
  const uint32_t esz = (uint32_t)c->esz;
  const int p = c->prob.p, NB = c->NB;
  SmemCfg cfg{};
  cfg.stage_elems = (uint32_t)tile_elems(p);
  const uint32_t stage_bytes = cfg.stage_elems * esz;
  // per-warp scratch
  uint32_t e = 0;
  cfg.w_off = e;     e += align_up((uint32_t)(p + 4), 4);   // holds the full theta in the HMC kernel
  cfg.rbuf_off = e;  e += TB + 8;
  cfg.ckpt_off = e;  e += align_up(ckpt_elems < 0 ? 6u * (uint32_t)NB : (uint32_t)ckpt_elems, 4);    // (a,P) or the 5-value trend state
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
  else { cfg.resident = 0; if (nst > max_stages) nst = max_stages; }
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

// Launch shape of the per-column select kernels (k_row_quantiles, k_impact_jobs): threads per
// column CTA and whether the column's keys are staged in shared memory (else every sweep reads
// them through L2).  elem_bytes = size of one staged key.  CI_B200_SEL_NT / CI_B200_SEL_SMEM
// override (tuning).
inline void select_launch_cfg(const ci_ctx* c, int S, size_t elem_bytes, int* nt, size_t* bytes,
                              int* in_smem) {
  size_t b = (((size_t)S * elem_bytes) + 15) & ~(size_t)15;
  int sm = b + QSTATIC <= (size_t)c->smem_optin;
  if (c->sel_smem >= 0 && !c->sel_smem) sm = 0;
  if (!sm) b = 0;
  // threads per column CTA, measured (run r2_15, tools/tune_impact.py): 10 000 draws per column
  // 0.38 ms with 512 threads vs 0.48 with 1024 (two key-staging CTAs per SM either way); 400 draws
  // per column (the batched panel call, ~400 k column CTAs) 7.1 ms with 128 threads vs 16.0 with 512
  int n = S <= 128 ? 64 : (S <= 2048 ? 128 : (S <= 4096 ? 256 : 512));
  if (c->sel_nt > 0) n = c->sel_nt;
  *nt = n; *bytes = b; *in_smem = sm;
}

int pick_G(const ci_ctx* c, int C) {
  if (c->force_G > 0) return c->force_G > MAXG ? MAXG : c->force_G;
  int G = (C + c->sm_count - 1) / c->sm_count;
  if (G < 1) G = 1;
  if (G > MAXG) G = MAXG;
  return G;
}

// Team mode (ci_team.cuh): one warp per tile, W = NB warps per chain.  Used when
// the whole series is resident in shared memory and has 1..MAXW tiles.
template <typename R>
bool plan_team(const ci_ctx* c, int C, int* GT, SmemCfg* cfg) {
  const int W = c->NB;
  // (W = 1, a single tile, is a team of one: same kernel, no checkpoint replay between the
  // forward and the adjoint sweep)
  if (!c->team_mode || W < 1 || W > MAXW || c->prob.model != CI_MODEL_LOCAL_LEVEL) return false;
  int gt = (C >= 4 * c->sm_count) ? MAXW / W : 1;
  if (gt < 1) gt = 1;
  if (c->force_G > 0) gt = c->force_G * W <= MAXW ? c->force_G : 1;
  const uint32_t tail = (uint32_t)gt * (uint32_t)sizeof(TeamShared<R>);
  std::string keep = cih_err();
  if (plan_smem(c, gt * W, 0, cfg, tail) != CI_OK || !cfg->resident) { cih_err() = keep; return false; }
  *GT = gt;
  return true;
}


// Long-series team mode (ci_team_stream.cuh): W warps per chain walking the series in rounds
// of W tiles.  Used for the local level model when the series has more tiles than a resident
// team can hold (NB > MAXW, or the tiles do not fit in shared memory).  GT teams per CTA.
template <typename R>
bool plan_tstream(const ci_ctx* c, int C, int* GT, int* Wout, SmemCfg* cfg) {
  if (!c->team_mode || !c->tstream_mode || c->prob.model != CI_MODEL_LOCAL_LEVEL) return false;
  const int NB = c->NB;
  if (NB < 2) return false;
  int W = c->tstream_W > 0 ? c->tstream_W : TS_W;
  if (W > MAXW) W = MAXW;
  if (W > NB) W = NB;
  if (W < 2) return false;
  // teams per CTA: enough warps per SM to hide the scan latencies (up to TS_MAXWARPS), but no
  // more CTAs' worth of chains than there are
  int gt = c->force_G > 0 ? c->force_G : (C + c->sm_count - 1) / c->sm_count;
  if (gt < 1) gt = 1;
  if (gt * W > TS_MAXWARPS) gt = TS_MAXWARPS / W;
  if (gt > 8) gt = 8;
  std::string keep = cih_err();
  for (; gt >= 1; --gt) {
    const uint32_t tail = (uint32_t)gt * (uint32_t)tstream_team_bytes<R>(NB) + 16u;
    // streaming needs a ring of at least two rounds: finishing a tile triggers the copy of the
    // tile nstage - W positions ahead, which must reach into the next round
    if (plan_smem(c, gt * W, 0, cfg, tail, 0, 3u * (uint32_t)W) == CI_OK &&
        (cfg->resident || cfg->nstage >= 2u * (uint32_t)W)) {
      cih_err() = keep;
      *GT = gt; *Wout = W;
      return true;
    }
  }
  cih_err() = keep;
  return false;
}

}  // namespace
