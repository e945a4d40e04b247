// ci_hmc.cuh -- K6: the whole HMC run of a chain inside ONE persistent kernel.
//
// Replaces gibbs_sampler.fit_with_gibbs_sampling as called by the reference
// (causalimpact/causalimpact_lib.py:365-388) with the batched-chain HMC that
// BASELINE.json's north_star specifies.  One warp per chain; chains never
// communicate, so there is no grid-wide synchronisation: warm-up adaptation,
// leapfrog integration, Metropolis test and draw output all happen in
// registers / shared memory of the owning warp, and every leapfrog step calls
// chain_eval() (ci_device.cuh) on [X|y] tiles that stay resident in shared
// memory (or stream through the mbarrier ring for long / wide problems).
// theta component i lives in lane i%32, slot i/32.
// oracle/hmc_np.py restates this algorithm line by line.
#pragma once
#include "../../include/ci_b200.h"
#include "ci_kernels.cuh"

namespace ci {

struct HmcPlan {
  int n_warmup, n_results, max_leapfrog, adapt_mass;
  double init_step, target_accept;
  int init_buf, slow_end, n_ends;
  int ends[16];
  long long n_evals;       // 1 + sum_it L_it  (drives the tile producer)
};

__host__ __device__ inline int hmc_leapfrog_count(uint64_t seed, int it, int max_leapfrog) {
  const uint4 x = Philox::gen(seed, 0u, RNG_LEAPFROG, (uint32_t)it, 0u);
  return 1 + (int)(((uint64_t)x.x * (uint64_t)max_leapfrog) >> 32);
}

// The HMC loop of ONE chain.  `Ev` supplies the target:
//   ev.publish(theta)   make theta visible to the evaluator (shared memory)
//   ev.eval(lp, g)      log posterior + gradient slots at the published point
//   ev.writer()         true for the warp that writes draws / stats
// In team mode every warp of the team runs this loop redundantly on identical
// values (same RNG keys, same totals), so no extra synchronisation is needed.
template <typename R, typename Ev>
__device__ __forceinline__ void hmc_chain(Ev& ev, const HmcPlan& plan, uint64_t seed, uint64_t gid,
                                          const R* __restrict__ theta0_row, int dim, int lane,
                                          int c, int C, R* __restrict__ draws,
                                          ci_hmc_stats* __restrict__ stats) {
  const uint32_t id_lo = (uint32_t)gid, id_hi8 = (uint32_t)(gid >> 32) << 8;
  R th[DSLOTS], g[DSLOTS], minv[DSLOTS], wmean[DSLOTS], wm2[DSLOTS];
#pragma unroll
  for (int s = 0; s < DSLOTS; ++s) {
    const int i = lane + 32 * s;
    th[s] = i < dim ? theta0_row[i] : (R)0;
    minv[s] = 1; wmean[s] = 0; wm2[s] = 0;
  }
  double lp;
  ev.publish(th);
  ev.eval(lp, g);

  double eps = plan.init_step;
  double mu = log(10.0 * eps), hbar = 0.0, leb = 0.0, dac = 0.0, wn = 0.0;
  double acc_sum = 0.0;
  int n_div = 0, n_leap = 1, next_end = 0;
  const int n_iter = plan.n_warmup + plan.n_results;

  for (int it = 0; it < n_iter; ++it) {
    const int L = hmc_leapfrog_count(seed, it, plan.max_leapfrog);
    // ---- momentum ----
    R rho[DSLOTS], thn[DSLOTS], gn[DSLOTS];
    double kin = 0.0;
#pragma unroll
    for (int s = 0; s < DSLOTS; ++s) {
      const int i = lane + 32 * s;
      R z = 0;
      if (i < dim) {
        const uint4 x = Philox::gen(seed, id_lo, RNG_MOMENTUM | id_hi8, (uint32_t)it,
                                    (uint32_t)(i >> 2));
        R z0, z1, z2, z3;
        box_muller<R>(x.x, x.y, z0, z1);
        box_muller<R>(x.z, x.w, z2, z3);
        const int sel = i & 3;
        z = sel == 0 ? z0 : (sel == 1 ? z1 : (sel == 2 ? z2 : z3));
      }
      rho[s] = i < dim ? z / Num<R>::sqrt(minv[s]) : (R)0;
      kin += (double)(minv[s] * rho[s] * rho[s]);
      thn[s] = th[s]; gn[s] = g[s];
    }
    const double H0 = -lp + 0.5 * warp_sum(kin);
    // ---- leapfrog ----
    const R e = (R)eps;
    double lpn = lp;
#pragma unroll
    for (int s = 0; s < DSLOTS; ++s) rho[s] = fma((R)0.5 * e, gn[s], rho[s]);
    for (int i = 0; i < L; ++i) {
#pragma unroll
      for (int s = 0; s < DSLOTS; ++s) thn[s] = fma(e * minv[s], rho[s], thn[s]);
      ev.publish(thn);
      ev.eval(lpn, gn);
      const R f = (i < L - 1) ? e : (R)0.5 * e;
#pragma unroll
      for (int s = 0; s < DSLOTS; ++s) rho[s] = fma(f, gn[s], rho[s]);
    }
    n_leap += L;
    kin = 0.0;
#pragma unroll
    for (int s = 0; s < DSLOTS; ++s) kin += (double)(minv[s] * rho[s] * rho[s]);
    const double H1 = -lpn + 0.5 * warp_sum(kin);
    const double dH = H0 - H1;
    const bool fin = isfinite(dH);
    const double alpha = fin ? fmin(1.0, exp(fmin(dH, 0.0))) : 0.0;
    const bool div = !fin || dH < -1000.0;
    const uint4 xa = Philox::gen(seed, id_lo, RNG_ACCEPT | id_hi8, (uint32_t)it, 0u);
    const double ua = u01<double>(xa.x);
    if (ua < alpha) {
#pragma unroll
      for (int s = 0; s < DSLOTS; ++s) { th[s] = thn[s]; g[s] = gn[s]; }
      lp = lpn;
    }
    if (it < plan.n_warmup) {
      // ---- dual averaging ----
      dac += 1.0;
      const double eta = 1.0 / (dac + 10.0);
      hbar = (1.0 - eta) * hbar + eta * (plan.target_accept - alpha);
      const double le = mu - hbar * sqrt(dac) / 0.05;
      const double ex = pow(dac, -0.75);
      leb = (1.0 - ex) * leb + ex * le;
      eps = exp(le);
      // ---- windowed diagonal mass ----
      if (plan.adapt_mass && it >= plan.init_buf && it < plan.slow_end) {
        wn += 1.0;
        const R rn = (R)(1.0 / wn);
#pragma unroll
        for (int s = 0; s < DSLOTS; ++s) {
          const R d = th[s] - wmean[s];
          wmean[s] = fma(d, rn, wmean[s]);
          wm2[s] = fma(d, th[s] - wmean[s], wm2[s]);
        }
        if (next_end < plan.n_ends && it == plan.ends[next_end]) {
          ++next_end;
          const R a = (R)(wn / (wn + 5.0)), b = (R)(1e-3 * 5.0 / (wn + 5.0));
          const R rd = (R)(1.0 / (wn - 1.0));
#pragma unroll
          for (int s = 0; s < DSLOTS; ++s) {
            minv[s] = fma(a, wm2[s] * rd, b);
            wmean[s] = 0; wm2[s] = 0;
          }
          wn = 0.0;
          mu = log(10.0 * eps); hbar = 0.0; leb = 0.0; dac = 0.0;
        }
      }
      if (it == plan.n_warmup - 1) eps = exp(leb);
    } else {
      if (ev.writer()) {
        const size_t row = ((size_t)(it - plan.n_warmup) * C + c) * dim;
#pragma unroll
        for (int s = 0; s < DSLOTS; ++s) {
          const int i = lane + 32 * s;
          if (i < dim) draws[row + i] = th[s];
        }
      }
      acc_sum += alpha;
      n_div += div ? 1 : 0;
    }
  }
  if (ev.writer() && lane == 0) {
    ci_hmc_stats st;
    st.accept_rate = (float)(acc_sum / (plan.n_results > 0 ? plan.n_results : 1));
    st.step_size = (float)eps;
    st.n_divergent = n_div;
    st.n_leapfrog = n_leap;
    stats[c] = st;
  }
}

// ---- evaluator: one warp per chain, tiles walked sequentially (any T) ----
template <typename R> struct WarpEval {
  TilePipe<R>& pipe; const ProbDev<R>& pr; const WarpScratch<R>& ws; const R* omega; int lane;
  // (the caller waits for Omega once before the HMC loop starts)
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
    const R u = ws.w[p], l = ws.w[p + 1];
    const R s_e = Num<R>::exp(u), s_h = Num<R>::exp(l);
    double ll, g_se, g_sh;
    R gw[JS];
    chain_eval(pipe, pr, ws, s_e, s_h, true, lane, ll, g_se, g_sh, gw);
    double g_u = g_se * (double)s_e, g_l = g_sh * (double)s_h;
    lp = ll + chain_prior(pr, omega, ws.w, u, l, s_e, s_h, lane, gw, g_u, g_l);
#pragma unroll
    for (int s = 0; s < DSLOTS; ++s) {
      const int i = lane + 32 * s;
      g[s] = i < p ? gw[s] : (i == p ? (R)g_u : (i == p + 1 ? (R)g_l : (R)0));
    }
  }
};

template <typename R>
__global__ void __launch_bounds__(32 * (MAXG + 1), 1)
k_hmc(ProbDev<R> pr, SmemCfg cfg, HmcPlan plan, uint64_t seed, uint64_t chain_id0,
      const R* __restrict__ theta0, int C, R* __restrict__ draws,
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
  WarpEval<R> ev{pipe, pr, ws, cs.omega, lane};
  hmc_chain<R>(ev, plan, seed, chain_id0 + (uint64_t)c, theta0 + (size_t)c * pr.dim, pr.dim, lane,
               c, C, draws, stats);
}

}  // namespace ci
