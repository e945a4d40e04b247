// ci_team_kernels.cuh -- team-mode kernels (one warp per tile of a chain).
// CTA = GT teams x W warps + 1 producer warp; tiles are resident in shared
// memory and shared by every team of the CTA.
#pragma once
#include "ci_hmc.cuh"
#include "ci_team.cuh"

namespace ci {

// byte offset of the team exchange areas = cfg.off_warp + n_warps*cfg.warp_bytes
template <typename R>
__device__ __forceinline__ TeamShared<R>* team_area(unsigned char* smem, const SmemCfg& cfg,
                                                    int n_warps, int team) {
  size_t off = (size_t)cfg.off_warp + (size_t)n_warps * cfg.warp_bytes;
  off = (off + 15) & ~(size_t)15;
  return reinterpret_cast<TeamShared<R>*>(smem + off) + team;
}

template <typename R> struct TeamEval {
  const R* tile; const ProbDev<R>& pr; TeamShared<R>* ts; const WarpScratch<R>& ws;
  const R* omega; int lane, wt, W, bar_id;
  __device__ __forceinline__ bool writer() const { return wt == 0; }
  __device__ __forceinline__ void publish(const R (&t)[DSLOTS]) {
    __syncwarp();
#pragma unroll
    for (int s = 0; s < DSLOTS; ++s) {
      const int i = lane + 32 * s;
      if (i < pr.dim) ws.w[i] = t[s];       // every warp keeps its own copy
    }
    __syncwarp();
  }
  __device__ __forceinline__ void eval(double& lp, R (&g)[DSLOTS]) {
    const int p = pr.p;
    const R u = ws.w[p], l = ws.w[p + 1];
    const R s_e = Num<R>::exp(u), s_h = Num<R>::exp(l);
    double ll, g_se, g_sh;
    R gw[JS];
    team_eval(tile, pr, ts, ws.w, ws.rbuf, s_e, s_h, true, lane, wt, W, bar_id, ll, g_se, g_sh,
              gw);
    double g_u = g_se * (double)s_e, g_l = g_sh * (double)s_h;
    lp = ll + chain_prior(pr, omega, ws.w, u, l, s_e, s_h, lane, gw, g_u, g_l);
#pragma unroll
    for (int s = 0; s < DSLOTS; ++s) {
      const int i = lane + 32 * s;
      g[s] = i < p ? gw[s] : (i == p ? (R)g_u : (i == p + 1 ? (R)g_l : (R)0));
    }
  }
};

// shared prologue of the team kernels: returns false for warps that have no work
template <typename R>
__device__ __forceinline__ bool team_prologue(unsigned char* smem, const SmemCfg& cfg,
                                              const ProbDev<R>& pr, int W, int C,
                                              CtaShared<R>& cs, int& team, int& wt, int& c) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_warps = (blockDim.x >> 5) - 1;
  const int GT = n_warps / W;
  cs = cta_prologue(smem, cfg, pr, 1);
  if (warp == n_warps) {
    if (lane == 0) omega_fetch(cs, pr);
    if (lane == 0)
      tile_producer(pr.tiles, cs.stage0, cs.full, cs.empty, cfg.stage_elems, cfg.nstage, pr.NB,
                    true, 0LL, [](long long) { return true; });
    return false;
  }
  team = warp / W; wt = warp - team * W;
  c = blockIdx.x * GT + team;
  return c < C;
}

template <typename R>
__global__ void __launch_bounds__(32 * (MAXW + 1), 1)
k_logpost_team(ProbDev<R> pr, SmemCfg cfg, int W, const R* __restrict__ theta, int C,
               R* __restrict__ value, R* __restrict__ grad, int flags) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = (blockDim.x >> 5) - 1;
  const int GT = n_warps / W;
  const int p = pr.p, dim = pr.dim;
#ifdef CI_CLK
  const long long clk_entry = clock64();
#endif
  const CtaShared<R> cs = cta_prologue(smem, cfg, pr, 1);
  if (warp == n_warps) {
    // producer warp: start the bulk copies, then -- instead of idling -- evaluate the log
    // prior + Jacobian of every chain of the CTA (it needs theta and Omega, not the series)
    // and hand it to the team's warp 0 through a named barrier (run 31: the prior was a
    // 1 800-cycle serial tail of the evaluation)
    if (lane == 0) {
      omega_fetch(cs, pr);
      tile_producer(pr.tiles, cs.stage0, cs.full, cs.empty, cfg.stage_elems, cfg.nstage, pr.NB,
                    true, 0LL, [](long long) { return true; });
    }
    __syncwarp();
    if (flags & 1) {
      omega_wait(cs);
      for (int t = 0; t < GT; ++t) {
        const int cc = blockIdx.x * GT + t;
        if (cc >= C) break;
        const R* th = theta + (size_t)cc * dim;
        const R u = th[p], l = th[p + 1];
        const R s_e = Num<R>::exp(u), s_h = Num<R>::exp(l);
        R gw[JS];
#pragma unroll
        for (int s = 0; s < JS; ++s) gw[s] = 0;
        double g_u = 0.0, g_l = 0.0;
        const double lp = chain_prior(pr, cs.omega, th, u, l, s_e, s_h, lane, gw, g_u, g_l);
        TeamShared<R>* ts = team_area<R>(smem, cfg, n_warps, t);
        if (lane == 0) { ts->prior[0] = lp; ts->prior[1] = g_u; ts->prior[2] = g_l; }
#pragma unroll
        for (int s = 0; s < JS; ++s) {
          const int j = lane + 32 * s;
          if (j < p) ts->gwprior[j] = gw[s];
        }
        asm volatile("bar.arrive %0, 64;" ::"r"(8 + t) : "memory");
      }
    }
    return;
  }
  const int team = warp / W, wt = warp - team * W;
  const int c = blockIdx.x * GT + team;
  if (c >= C) return;
#ifdef CI_CLK
  if (blockIdx.x == 0 && lane == 0 && (wt == 0 || wt == W - 1)) g_clk[wt == 0 ? 0 : 1][0] = clk_entry;
#endif
  const R* th = theta + (size_t)c * dim;
  const WarpScratch<R> ws = warp_scratch<R>(smem, cfg, warp);
  for (int j = lane; j < dim; j += 32) ws.w[j] = th[j];
  __syncwarp();
  const R u = ws.w[p], l = ws.w[p + 1];
  const R s_e = Num<R>::exp(u), s_h = Num<R>::exp(l);
  mbar_wait(&cs.full[wt], 0u);
  const R* tile = cs.stage0 + (size_t)wt * cfg.stage_elems;
  const bool want_grad = grad != nullptr;
  double ll, g_se, g_sh;
  R gw[JS];
  TeamShared<R>* ts = team_area<R>(smem, cfg, n_warps, team);
  team_eval(tile, pr, ts, ws.w, ws.rbuf, s_e, s_h, want_grad, lane, wt, W, team + 1, ll, g_se,
            g_sh, gw);
  if (wt != 0) return;
  double val = ll;
  double g_u = g_se * (double)s_e, g_l = g_sh * (double)s_h;
  if (flags & 1) {
    asm volatile("bar.sync %0, 64;" ::"r"(8 + team) : "memory");
    val += ts->prior[0]; g_u += ts->prior[1]; g_l += ts->prior[2];
#pragma unroll
    for (int s = 0; s < JS; ++s) {
      const int j = lane + 32 * s;
      if (j < p) gw[s] += ts->gwprior[j];
    }
  }
  if (lane == 0) value[c] = (R)val;
  if (want_grad) {
    R* g = grad + (size_t)c * dim;
#pragma unroll
    for (int s = 0; s < JS; ++s) {
      const int j = lane + 32 * s;
      if (j < p) g[j] = gw[s];
    }
    if (lane == 0) { g[p] = (R)g_u; g[p + 1] = (R)g_l; }
  }
  CI_CLK_MARK(15);
}

template <typename R>
__global__ void __launch_bounds__(32 * (MAXW + 1), 1)
k_hmc_team(ProbDev<R> pr, SmemCfg cfg, int W, HmcPlan plan, uint64_t seed, uint64_t chain_id0,
           const R* __restrict__ theta0, int C, R* __restrict__ draws,
           ci_hmc_stats* __restrict__ stats) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int n_warps = (blockDim.x >> 5) - 1;
  CtaShared<R> cs; int team, wt, c;
  if (!team_prologue(smem, cfg, pr, W, C, cs, team, wt, c)) return;
  const WarpScratch<R> ws = warp_scratch<R>(smem, cfg, threadIdx.x >> 5);
  mbar_wait(&cs.full[wt], 0u);
  const R* tile = cs.stage0 + (size_t)wt * cfg.stage_elems;
  omega_wait(cs);
  TeamEval<R> ev{tile, pr, team_area<R>(smem, cfg, n_warps, team), ws, cs.omega, lane, wt, W,
                 team + 1};
  hmc_chain<R>(ev, plan, seed, chain_id0 + (uint64_t)c, theta0 + (size_t)c * pr.dim, pr.dim, lane,
               c, C, draws, stats);
}

// CTA = GT teams x W warps, NO producer warp (thread 0 starts the resident tile copies): up to
// PT_MAXWARPS warps share one resident copy of the series -- at BASELINE configs[4] (T = 2000:
// 106 KB of tiles per CTA) the one-warp kernel fits a single 8-warp CTA per SM (ncu r2_09:
// 10.8 % warps active); 16 warps per CTA double that with ~45 % fewer instructions per draw.
constexpr int PT_MAXWARPS = 16;

template <typename R>
__global__ void __launch_bounds__(32 * PT_MAXWARPS, 1)
k_predict_team(ProbDev<R> pr, SmemCfg cfg, int W, const R* __restrict__ theta, int S,
               uint64_t seed, uint64_t draw_id0, R* __restrict__ level, R* __restrict__ traj) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = blockDim.x >> 5;
  const int GT = n_warps / W;
  const CtaShared<R> cs = cta_prologue(smem, cfg, pr, 1);
  if (threadIdx.x == 0)
    tile_producer(pr.tiles, cs.stage0, cs.full, cs.empty, cfg.stage_elems, cfg.nstage, pr.NB, true,
                  0LL, [](long long) { return true; });
  const int team = warp / W, wt = warp - team * W;
  const int s = blockIdx.x * GT + team;
  if (s >= S) return;
  const int p = pr.p, dim = pr.dim;
  const R* th = theta + (size_t)s * dim;
  const WarpScratch<R> ws = warp_scratch<R>(smem, cfg, warp);
  for (int j = lane; j < dim; j += 32) ws.w[j] = th[j];
  __syncwarp();
  const R s_e = Num<R>::exp(ws.w[p]), s_h = Num<R>::exp(ws.w[p + 1]);
  mbar_wait(&cs.full[wt], 0u);
  const R* tile = cs.stage0 + (size_t)wt * cfg.stage_elems;
  const size_t row = (size_t)s * pr.T;
  team_predict(tile, pr, team_area<R>(smem, cfg, n_warps, team), ws.w, s_e, s_h,
               Num<R>::sqrt(s_e), seed, draw_id0 + (uint64_t)s, lane, wt, W, team + 1,
               level ? level + row : nullptr, traj + row);
}

}  // namespace ci
