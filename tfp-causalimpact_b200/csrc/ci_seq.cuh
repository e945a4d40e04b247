// ci_seq.cuh -- CI_VARIANT_SEQ: the per-timestep Kalman recursion, strictly
// sequential in time, filter state in registers, ONE WARP PER CHAIN.
//
// This is the "plain" variant BASELINE.json's north_star names next to the
// associative-scan one: the 32 lanes cooperate only on the data-parallel part
// (lane L computes the residual r_t = y_t - x_t.w of step t = 32 g + L from the
// shared-memory tile), then every lane runs the same 32-step predict/update
// recursion redundantly, fetching r_t with a shuffle broadcast; the values of
// step t that the adjoint sweep needs (P_t, K_t, 1/F_t, v_t) stay in lane t's
// registers.  The backward sweep replays each 32-step group from an (a, P)
// checkpoint and runs the adjoint recursion of ci_filter.cuh step by step.
// It replaces the same reference arithmetic as the scan kernels (TFP LGSSM
// log_prob, call site causalimpact/causalimpact_lib.py:365-388) and serves as
// the in-repo baseline the scan / team kernels are measured against, and as an
// independent implementation for cross-checks (tests/test_gpu_logprob.py).
#pragma once
#include "ci_kernels.cuh"

namespace ci {

constexpr int SEQ_G = 32;                 // steps per group (one per lane)
constexpr int GROUPS_PER_TILE = TB / SEQ_G;

// residual of step tl (tile-local) for the calling lane; NaN if masked
template <typename R>
__device__ __forceinline__ R seq_residual(const R* __restrict__ tile, const R* __restrict__ w_s,
                                          int p, int ld, int tl) {
  const R* row = tile + tile_off(tl, ld);
  R acc = row[p];
#pragma unroll 4
  for (int j = 0; j < p; ++j) acc = fma(-row[j], w_s[j], acc);
  return acc;
}

template <typename R>
__global__ void __launch_bounds__(32 * (MAXG + 1), 1)
k_logpost_seq(ProbDev<R> pr, SmemCfg cfg, const R* __restrict__ theta, int C,
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
  const int p = pr.p, dim = pr.dim, ld = pr.ld, NB = pr.NB;
  const R* th = theta + (size_t)c * dim;
  const WarpScratch<R> ws = warp_scratch<R>(smem, cfg, warp);
  for (int j = lane; j < p; j += 32) ws.w[j] = th[j];
  const R u = th[p], l = th[p + 1];
  const R s_e = Num<R>::exp(u), s_h = Num<R>::exp(l);
  __syncwarp();
  TilePipe<R> pipe = make_pipe(cs, cfg);
  R* ckpt = ws.extra;                      // [2 * NB * GROUPS_PER_TILE]

  // ---------------- forward ----------------
  R a = pr.m0, P = pr.P0;
  double acc = 0.0;
  int n_obs = 0;
  for (int b = 0; b < NB; ++b) {
    const R* tile = pipe.acquire(b);
    for (int g = 0; g < GROUPS_PER_TILE; ++g) {
      const R r_mine = seq_residual(tile, ws.w, p, ld, g * SEQ_G + lane);
      const unsigned obs = __ballot_sync(FULL, r_mine == r_mine);
      if (want_grad && lane == 0) {
        ckpt[2 * (b * GROUPS_PER_TILE + g)] = a; ckpt[2 * (b * GROUPS_PER_TILE + g) + 1] = P;
      }
      R part = 0;
#pragma unroll 8
      for (int k = 0; k < SEQ_G; ++k) {
        const R r = __shfl_sync(FULL, r_mine, k);
        if ((obs >> k) & 1u) {
          const R F = P + s_e, rF = Num<R>::rcp(F), K = P * rF, v = r - a;
          if (lane == k) part = Num<R>::log(F) + v * v * rF;    // one lane pays for the log
          a = fma(K, v, a);
          P = fma(-K, P, P);
        }
        P += s_h;
      }
      acc += (double)part;
      n_obs += __popc(obs);
    }
    pipe.release(lane);
  }
  const double ll = -0.5 * (warp_sum(acc) + 1.8378770664093453 * (double)n_obs);
  double g_se = 0.0, g_sh = 0.0;
  R gw[JS];
#pragma unroll
  for (int s = 0; s < JS; ++s) gw[s] = 0;

  // ---------------- backward (adjoint), group by group in reverse ----------------
  if (want_grad) {
    __syncwarp();
    R ab = 0, pb = 0;
    double ge = 0.0, gh = 0.0;
    for (int b = NB - 1; b >= 0; --b) {
      const R* tile = pipe.acquire(b);
      for (int g = GROUPS_PER_TILE - 1; g >= 0; --g) {
        const int tl = g * SEQ_G + lane;
        const R r_mine = seq_residual(tile, ws.w, p, ld, tl);
        const unsigned obs = __ballot_sync(FULL, r_mine == r_mine);
        a = ckpt[2 * (b * GROUPS_PER_TILE + g)]; P = ckpt[2 * (b * GROUPS_PER_TILE + g) + 1];
        // replay: lane k keeps the values of its own step
        R mP = 0, mK = 0, mrF = 0, mv = 0;
#pragma unroll 8
        for (int k = 0; k < SEQ_G; ++k) {
          const R r = __shfl_sync(FULL, r_mine, k);
          R K = 0, rF = 0, v = 0;
          const R Pk = P;
          if ((obs >> k) & 1u) {
            rF = Num<R>::rcp(P + s_e); K = P * rF; v = r - a;
            a = fma(K, v, a);
            P = fma(-K, P, P);
          }
          P += s_h;
          if (lane == k) { mP = Pk; mK = K; mrF = rF; mv = v; }
        }
        // adjoint recursion, step 31 down to 0
        R rbar_mine = 0;
        R lge = 0, lgh = 0;
#pragma unroll 8
        for (int k = SEQ_G - 1; k >= 0; --k) {
          const R K = __shfl_sync(FULL, mK, k), rF = __shfl_sync(FULL, mrF, k);
          const R v = __shfl_sync(FULL, mv, k), Pk = __shfl_sync(FULL, mP, k);
          lgh += pb;
          const R dF = (R)-0.5 * (rF - v * v * rF * rF);
          const R omk = (R)1 - K;
          const R rb = fma(K, ab, -v * rF);
          lge += fma(K * K, pb, dF) - ab * v * Pk * rF * rF;
          pb = fma(omk * omk, pb, fma(ab * v * s_e, rF * rF, dF));
          ab = fma(omk, ab, v * rF);
          if (lane == k) rbar_mine = rb;
        }
        ge += (double)lge; gh += (double)lgh;     // identical in every lane
        // d ll / d w_j  -=  sum_t rbar_t x_tj : lane t contributes its step
        const R* row = tile + tile_off(tl, ld);
        for (int j = 0; j < p; ++j) {
          const R tot = warp_sum(rbar_mine * row[j]);
#pragma unroll
          for (int s = 0; s < JS; ++s)
            if ((j >> 5) == s && lane == (j & 31)) gw[s] -= tot;
        }
      }
      pipe.release(lane);
    }
    g_se = ge; g_sh = gh;
  }

  double val = ll;
  double g_u = g_se * (double)s_e, g_l = g_sh * (double)s_h;
  if (flags & 1) {
    omega_wait(cs);
    val += chain_prior(pr, cs.omega, ws.w, u, l, s_e, s_h, lane, gw, g_u, g_l);
  }
  if (lane == 0) value[c] = (R)val;
  if (want_grad) {
    R* gout = grad + (size_t)c * dim;
#pragma unroll
    for (int s = 0; s < JS; ++s) {
      const int j = lane + 32 * s;
      if (j < p) gout[j] = gw[s];
    }
    if (lane == 0) { gout[p] = (R)g_u; gout[p + 1] = (R)g_l; }
  }
}

}  // namespace ci
