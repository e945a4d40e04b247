// ci_llt_kernels.cuh -- __global__ entry points of the local-linear-trend model
// (one warp per chain, tiles walked sequentially; resident or streamed).
#pragma once
#include "ci_hmc.cuh"
#include "ci_llt.cuh"

namespace ci {

template <typename R>
__global__ void __launch_bounds__(32 * (MAXG + 1), 1)
k_logpost_llt(ProbDev<R> pr, LltDev<R> ld2, SmemCfg cfg, const R* __restrict__ theta, int C,
              R* __restrict__ value, R* __restrict__ grad, int flags) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (blockDim.x >> 5) - 1;
  const int chain0 = blockIdx.x * G;
  const int nactive = min(G, C - chain0);
  const CtaShared<R> cs = cta_prologue(smem, cfg, pr, nactive);
  const bool want_grad = grad != nullptr;
  if (warp == G) {
    if (lane == 0) omega_fetch(cs, pr);
    if (lane == 0)
      tile_producer(pr.tiles, cs.stage0, cs.full, cs.empty, cfg.stage_elems, cfg.nstage, pr.NB,
                    cfg.resident != 0, want_grad ? 2LL : 1LL,
                    [](long long s) { return (s & 1) == 0; });
    return;
  }
  if (warp >= nactive) return;
  const int c = chain0 + warp;
  const int p = pr.p, dim = pr.dim;
  const R* th = theta + (size_t)c * dim;
  const WarpScratch<R> ws = warp_scratch<R>(smem, cfg, warp);
  for (int j = lane; j < p; j += 32) ws.w[j] = th[j];
  const R u = th[p], l = th[p + 1], s = th[p + 2];
  const R s_e = Num<R>::exp(u), q1 = Num<R>::exp(l), q2 = Num<R>::exp(s);
  __syncwarp();
  TilePipe<R> pipe = make_pipe(cs, cfg);
  double ll, g_se, g_q1, g_q2;
  R gw[JS];
  chain_eval_llt(pipe, pr, ld2, ws, s_e, q1, q2, want_grad, lane, ll, g_se, g_q1, g_q2, gw);
  double val = ll;
  double g_u = g_se * (double)s_e, g_l = g_q1 * (double)q1, g_s = g_q2 * (double)q2;
  if (flags & 1) {
    omega_wait(cs);
    val += chain_prior(pr, cs.omega, ws.w, u, l, s_e, q1, lane, gw, g_u, g_l);
    val += llt_slope_prior(ld2, s, q2, g_s);
  }
  if (lane == 0) value[c] = (R)val;
  if (want_grad) {
    R* g = grad + (size_t)c * dim;
#pragma unroll
    for (int sl = 0; sl < JS; ++sl) {
      const int j = lane + 32 * sl;
      if (j < p) g[j] = gw[sl];
    }
    if (lane == 0) { g[p] = (R)g_u; g[p + 1] = (R)g_l; g[p + 2] = (R)g_s; }
  }
}

template <typename R> struct WarpEvalLlt {
  TilePipe<R>& pipe; const ProbDev<R>& pr; const LltDev<R>& ld2; const WarpScratch<R>& ws;
  const R* omega; int lane;
  __device__ __forceinline__ bool writer() const { return true; }
  __device__ __forceinline__ void publish(const R (&t)[DSLOTS]) {
    __syncwarp();
#pragma unroll
    for (int s = 0; s < DSLOTS; ++s) {
      const int i = lane + 32 * s;
      if (i < pr.dim) ws.w[i] = t[s];
    }
    __syncwarp();
  }
  __device__ __forceinline__ void eval(double& lp, R (&g)[DSLOTS]) {
    const int p = pr.p;
    const R u = ws.w[p], l = ws.w[p + 1], s = ws.w[p + 2];
    const R s_e = Num<R>::exp(u), q1 = Num<R>::exp(l), q2 = Num<R>::exp(s);
    double ll, g_se, g_q1, g_q2;
    R gw[JS];
    chain_eval_llt(pipe, pr, ld2, ws, s_e, q1, q2, true, lane, ll, g_se, g_q1, g_q2, gw);
    double g_u = g_se * (double)s_e, g_l = g_q1 * (double)q1, g_s = g_q2 * (double)q2;
    lp = ll + chain_prior(pr, omega, ws.w, u, l, s_e, q1, lane, gw, g_u, g_l);
    lp += llt_slope_prior(ld2, s, q2, g_s);
#pragma unroll
    for (int sl = 0; sl < DSLOTS; ++sl) {
      const int i = lane + 32 * sl;
      g[sl] = i < p ? gw[sl]
                    : (i == p ? (R)g_u : (i == p + 1 ? (R)g_l : (i == p + 2 ? (R)g_s : (R)0)));
    }
  }
};

template <typename R>
__global__ void __launch_bounds__(32 * (MAXG + 1), 1)
k_hmc_llt(ProbDev<R> pr, LltDev<R> ld2, SmemCfg cfg, HmcPlan plan, uint64_t seed,
          uint64_t chain_id0, const R* __restrict__ theta0, int C, R* __restrict__ draws,
          ci_hmc_stats* __restrict__ stats) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (blockDim.x >> 5) - 1;
  const int chain0 = blockIdx.x * G;
  const int nactive = min(G, C - chain0);
  const CtaShared<R> cs = cta_prologue(smem, cfg, pr, nactive);
  if (warp == G) {
    if (lane == 0) omega_fetch(cs, pr);
    if (lane == 0)
      tile_producer(pr.tiles, cs.stage0, cs.full, cs.empty, cfg.stage_elems, cfg.nstage, pr.NB,
                    cfg.resident != 0, 2LL * plan.n_evals,
                    [](long long s) { return (s & 1) == 0; });
    return;
  }
  if (warp >= nactive) return;
  const int c = chain0 + warp;
  const WarpScratch<R> ws = warp_scratch<R>(smem, cfg, warp);
  TilePipe<R> pipe = make_pipe(cs, cfg);
  omega_wait(cs);
  WarpEvalLlt<R> ev{pipe, pr, ld2, ws, cs.omega, lane};
  hmc_chain<R>(ev, plan, seed, chain_id0 + (uint64_t)c, theta0 + (size_t)c * pr.dim, pr.dim, lane,
               c, C, draws, stats);
}

}  // namespace ci
