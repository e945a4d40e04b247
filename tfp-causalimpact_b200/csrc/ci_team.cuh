// ci_team.cuh -- TEAM MODE: one chain is evaluated by a team of W = NB warps,
// ONE TILE (256 time steps) PER WARP, for series that fit in shared memory
// (NB <= 8, i.e. T <= 2048).  This is the second level of the associative
// scan: each warp scans its tile with shuffles exactly as in ci_filter.cuh,
// publishes the tile's aggregate map (2x2 Moebius matrix for the variance, an
// affine pair for the mean / abar / Pbar recursions) to shared memory, and
// after a named barrier every warp composes the aggregates of the tiles before
// (or, for the adjoints, after) its own to obtain its carry-in.  Nothing is
// recomputed: a tile's filter state stays in registers from the forward to the
// backward sweep, so a value+gradient evaluation costs one tile's worth of
// latency instead of 2 NB, and 35 % fewer instructions than the sequential-
// tile path (no checkpoint replay).
//
// Replaces the same reference arithmetic as ci_filter.cuh (TFP LGSSM log_prob,
// call site causalimpact/causalimpact_lib.py:365-388).
#pragma once
#include "ci_device.cuh"

namespace ci {

constexpr int MAXW = 8;   // warps per team == max tiles in team mode

// Developer instrumentation (-DCI_CLK): SM clock at the phase boundaries of team_eval, for
// CTA 0, warps 0 and W-1, read back through ci_debug_clocks().  Compiled out by default.
#ifdef CI_CLK
__device__ long long g_clk[2][16];
#define CI_CLK_MARK(i)                                                              \
  do {                                                                              \
    if (blockIdx.x == 0 && lane == 0 && (wt == 0 || wt == W - 1))                   \
      g_clk[wt == 0 ? 0 : 1][i] = clock64();                                        \
  } while (0)
#else
#define CI_CLK_MARK(i) do {} while (0)
#endif

// per-team exchange area in shared memory
template <typename R> struct TeamShared {
  R aggM[MAXW][4];     // Moebius aggregate of each tile
  R aggA[MAXW][2];     // mean recursion      a_out = m a_in + c
  R aggAB[MAXW][2];    // abar recursion (reverse)
  R aggPB[MAXW][2];    // Pbar recursion (reverse)
  double red[MAXW][4]; // ll-terms, ge, gh, n_obs partials
  R gwpart[MAXW][MAX_DIM];
  // log prior + Jacobian and its gradient, computed by the otherwise idle producer warp
  // while the team filters (k_logpost_team): lp, d/du, d/dl, then d/dw
  double prior[4];
  R gwprior[MAX_DIM];
};

__device__ __forceinline__ void team_sync(int bar_id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
}

// Evaluate the chain whose weights are in w_s.  wt = this warp's index in the
// team (== its tile).  On return every warp of the team holds identical
// results.  gw[s] = d ll / d w_j for j = lane + 32 s.
template <typename R>
__device__ __forceinline__ void team_eval(const R* __restrict__ tile, const ProbDev<R>& pr,
                                          TeamShared<R>* ts, const R* __restrict__ w_s, R* rbuf,
                                          R s_e, R s_h, bool want_grad, int lane, int wt, int W,
                                          int bar_id, double& ll, double& g_se, double& g_sh,
                                          R (&gw)[JS]) {
  const int p = pr.p, ld = pr.ld;
  const int nthreads = 32 * W;
  CI_CLK_MARK(1);
  Blk<R> B;
  blk_residuals(B, tile, w_s, p, ld, lane);
  CI_CLK_MARK(2);

  // ---------------- F1: variance path, tile aggregate ----------------
  const R alpha = s_e + s_h, beta = s_e * s_h;
  Mob<R> M{(R)1, (R)0, (R)0, (R)1};
  Mob<R> Mx[KS];                       // lane-local map of the steps BEFORE step k
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    Mx[k] = M;
    const bool o = (B.obs >> k) & 1u;
    const R e1 = o ? alpha : (R)1, e2 = o ? beta : s_h;
    const R f1 = o ? (R)1 : (R)0, f2 = o ? s_e : (R)1;
    Mob<R> N;
    N.a = fma(e1, M.a, e2 * M.c); N.b = fma(e1, M.b, e2 * M.d);
    N.c = fma(f1, M.a, f2 * M.c); N.d = fma(f1, M.b, f2 * M.d);
    M = N;
  }
  {
    const R s = Num<R>::rcp_fast(M.a + M.b + M.c + M.d);
    M.a *= s; M.b *= s; M.c *= s; M.d *= s;
  }
mob_scan_up(M, lane);
  if (lane == 31) { ts->aggM[wt][0] = M.a; ts->aggM[wt][1] = M.b; ts->aggM[wt][2] = M.c; ts->aggM[wt][3] = M.d; }
  Mob<R> E = mob_shfl_up(M, 1);
  if (lane == 0) { E.a = 1; E.b = 0; E.c = 0; E.d = 1; }
  CI_CLK_MARK(3);
  team_sync(bar_id, nthreads);
  CI_CLK_MARK(4);

  // ---------------- F2: carry-in P, sequential P; mean aggregate ----------------
  {
    Mob<R> Pre{(R)1, (R)0, (R)0, (R)1};
    for (int t = 0; t < wt; ++t) {
      const Mob<R> A{ts->aggM[t][0], ts->aggM[t][1], ts->aggM[t][2], ts->aggM[t][3]};
      Pre = mob_mul(A, Pre);
    }
    E = mob_mul(E, Pre);
  }
  // The KS predicted variances of the lane, ALL AT ONCE: P_k is the Moebius image of P0 under
  // (lane-local prefix before step k) o (everything before the lane) -- 8 independent
  // compositions and reciprocals instead of a dependent chain of 8 (run 31: the sequential
  // variance / gain pass was ~900 of the 14 000 cycles of an evaluation).
  const R n0 = fma(E.a, pr.P0, E.b), d0 = fma(E.c, pr.P0, E.d);    // P at the lane's first step = n0/d0
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const R num = fma(Mx[k].a, n0, Mx[k].b * d0), den = fma(Mx[k].c, n0, Mx[k].d * d0);
    const R Pk = num * Num<R>::rcp(den);
    B.P[k] = Pk;
    const bool o = (B.obs >> k) & 1u;
    const R rF = o ? Num<R>::rcp(Pk + s_e) : (R)0;
    B.rF[k] = rF; B.K[k] = Pk * rF;
  }
  R m = 1, c = 0;
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const R omk = (R)1 - B.K[k];
    c = fma(omk, c, B.K[k] * B.r[k]);
    m = omk * m;
  }
affine_scan_up(m, c, lane);
  if (lane == 31) { ts->aggA[wt][0] = m; ts->aggA[wt][1] = c; }
  R me = __shfl_up_sync(FULL, m, 1), ce = __shfl_up_sync(FULL, c, 1);
  if (lane == 0) { me = 1; ce = 0; }
  CI_CLK_MARK(5);
  team_sync(bar_id, nthreads);
  CI_CLK_MARK(6);

  // ---------------- F3: carry-in a, innovations, log-lik terms ----------------
  R a_in = pr.m0;
  for (int t = 0; t < wt; ++t) a_in = fma(ts->aggA[t][0], a_in, ts->aggA[t][1]);
  R ac = fma(me, a_in, ce);
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const R v = ((B.obs >> k) & 1u) ? (B.r[k] - ac) : (R)0;
    ac = fma(B.K[k], v, ac);
    B.v[k] = v;
  }
  const double ll_terms = warp_sum((double)blk_loglik_terms(B, s_e));
  const int n_obs = __reduce_add_sync(FULL, __popc(B.obs));

  CI_CLK_MARK(7);
  R abn[KS], q[KS], dF[KS], rbar[KS];
  R lge = 0, lgh = 0;
  if (want_grad) {
    // ---------------- B1: abar reverse scan, tile aggregate ----------------
    m = 1; c = 0;
#pragma unroll
    for (int k = KS - 1; k >= 0; --k) {
      const R omk = (R)1 - B.K[k];
      c = fma(omk, c, B.v[k] * B.rF[k]);
      m = omk * m;
    }
affine_scan_down(m, c, lane);
    if (lane == 0) { ts->aggAB[wt][0] = m; ts->aggAB[wt][1] = c; }
    me = __shfl_down_sync(FULL, m, 1); ce = __shfl_down_sync(FULL, c, 1);
    if (lane == 31) { me = 1; ce = 0; }
    CI_CLK_MARK(8);
    team_sync(bar_id, nthreads);
    CI_CLK_MARK(9);
    R ab_in = 0;
    for (int t = W - 1; t > wt; --t) ab_in = fma(ts->aggAB[t][0], ab_in, ts->aggAB[t][1]);
    R ab = fma(me, ab_in, ce);
#pragma unroll
    for (int k = KS - 1; k >= 0; --k) {
      abn[k] = ab;
      ab = fma((R)1 - B.K[k], ab, B.v[k] * B.rF[k]);
    }
    // ---------------- B2: Pbar reverse scan ----------------
    m = 1; c = 0;
#pragma unroll
    for (int k = KS - 1; k >= 0; --k) {
      const R omk = (R)1 - B.K[k];
      const R mult = omk * omk;
      const R rF = B.rF[k], v = B.v[k];
      dF[k] = (R)-0.5 * (rF - v * v * rF * rF);
      q[k] = fma(abn[k] * v * s_e, rF * rF, dF[k]);
      c = fma(mult, c, q[k]);
      m = mult * m;
    }
affine_scan_down(m, c, lane);
    if (lane == 0) { ts->aggPB[wt][0] = m; ts->aggPB[wt][1] = c; }
    me = __shfl_down_sync(FULL, m, 1); ce = __shfl_down_sync(FULL, c, 1);
    if (lane == 31) { me = 1; ce = 0; }
    CI_CLK_MARK(10);
    team_sync(bar_id, nthreads);
    CI_CLK_MARK(11);
    R pb_in = 0;
    for (int t = W - 1; t > wt; --t) pb_in = fma(ts->aggPB[t][0], pb_in, ts->aggPB[t][1]);
    R pb = fma(me, pb_in, ce);
#pragma unroll
    for (int k = KS - 1; k >= 0; --k) {
      const R K = B.K[k], rF = B.rF[k], v = B.v[k];
      const R omk = (R)1 - K;
      lgh += pb;
      lge += fma(K * K, pb, dF[k]) - abn[k] * v * B.P[k] * rF * rF;
      rbar[k] = fma(K, abn[k], -v * rF);
      pb = fma(omk * omk, pb, q[k]);
    }
    CI_CLK_MARK(12);
    // ---------------- X^T rbar of this tile ----------------
    if (p > 0) {
      if (p <= PSMALL) {
        R accw[PSMALL];
#pragma unroll
        for (int j = 0; j < PSMALL; ++j) accw[j] = 0;
        blk_xt_rbar_small(tile, rbar, p, ld, lane, accw);
        static_assert(PSMALL == 16, "warp_multi_sum16");
        warp_multi_sum16(accw, lane);            // 16 shuffles for all covariates (was 5 each)
        if (!(lane & 1) && (lane >> 1) < p) ts->gwpart[wt][lane >> 1] = accw[0];
      } else {
        const XtMap xm = xt_map(p, lane);
        R acc[JS];
#pragma unroll
        for (int s = 0; s < JS; ++s) acc[s] = 0;
#pragma unroll
        for (int k = 0; k < KS; ++k) rbuf[lane * KS + k + (lane >> 2)] = rbar[k];
        __syncwarp();
        if (p <= 32) {
          R g1[1] = {0};
          blk_xt_rbar<R, 1>(tile, rbuf, p, ld, xm.jj, xm.part, xm.nparts, g1);
          acc[0] = g1[0];
        } else {
          blk_xt_rbar<R, JS>(tile, rbuf, p, ld, xm.jj, xm.part, xm.nparts, acc);
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < JS; ++s) {
          R a = acc[s];
          for (int o = xm.PJ; o < 32; o <<= 1) a += __shfl_xor_sync(FULL, a, o);
          const int j = lane + 32 * s;
          if (j < p) ts->gwpart[wt][j] = a;
        }
      }
    }
  }
  CI_CLK_MARK(13);
  const double ge_w = warp_sum((double)lge), gh_w = warp_sum((double)lgh);
  if (lane == 0) {
    ts->red[wt][0] = ll_terms; ts->red[wt][1] = ge_w; ts->red[wt][2] = gh_w;
    ts->red[wt][3] = (double)n_obs;
  }
  team_sync(bar_id, nthreads);

  // ---------------- final: fixed-order sums, identical in every warp ----------------
  double s_ll = 0.0, s_ge = 0.0, s_gh = 0.0, s_n = 0.0;
  for (int t = 0; t < W; ++t) {
    s_ll += ts->red[t][0]; s_ge += ts->red[t][1]; s_gh += ts->red[t][2]; s_n += ts->red[t][3];
  }
  ll = -0.5 * (s_ll + 1.8378770664093453 * s_n);
  g_se = s_ge; g_sh = s_gh;
#pragma unroll
  for (int s = 0; s < JS; ++s) {
    const int j = lane + 32 * s;
    R a = 0;
    if (want_grad && j < p)
      for (int t = 0; t < W; ++t) a += ts->gwpart[t][j];
    gw[s] = -a;
  }
  // the exchange area is reused by the next evaluation
  team_sync(bar_id, nthreads);
  CI_CLK_MARK(14);
}

// ---------------------------------------------------------------------------
// TEAM-mode simulation smoother (K4): one warp per tile of a posterior draw.
// Forward exactly as team_eval (F1/F2/F3) but keeping the FILTERED means; the
// backward sampling recursion x_t = J_t x_{t+1} + (1-J_t) m_t + sqrt(V_t) z_t is
// a reverse affine scan whose tile aggregates are exchanged once more.  No
// checkpoint replay: ~45 % fewer instructions per draw than k_predict.
// ---------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void team_predict(const R* __restrict__ tile, const ProbDev<R>& pr,
                                             TeamShared<R>* ts, const R* __restrict__ w_s,
                                             R s_e, R s_h, R sig_e, uint64_t seed, uint64_t gid,
                                             int lane, int wt, int W, int bar_id,
                                             R* __restrict__ level_row, R* __restrict__ traj_row) {
  const int p = pr.p, ld = pr.ld, T = pr.T;
  const int nthreads = 32 * W;
  Blk<R> B;
  R xw[KS];
  blk_residuals_xw(B, xw, tile, w_s, p, ld, lane);
  // ---- F1: variance aggregate ----
  const R alpha = s_e + s_h, beta = s_e * s_h;
  Mob<R> M{(R)1, (R)0, (R)0, (R)1};
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const bool o = (B.obs >> k) & 1u;
    const R e1 = o ? alpha : (R)1, e2 = o ? beta : s_h;
    const R f1 = o ? (R)1 : (R)0, f2 = o ? s_e : (R)1;
    Mob<R> N;
    N.a = fma(e1, M.a, e2 * M.c); N.b = fma(e1, M.b, e2 * M.d);
    N.c = fma(f1, M.a, f2 * M.c); N.d = fma(f1, M.b, f2 * M.d);
    M = N;
  }
  {
    const R s = Num<R>::rcp_fast(M.a + M.b + M.c + M.d);
    M.a *= s; M.b *= s; M.c *= s; M.d *= s;
  }
mob_scan_up(M, lane);
  if (lane == 31) { ts->aggM[wt][0] = M.a; ts->aggM[wt][1] = M.b; ts->aggM[wt][2] = M.c; ts->aggM[wt][3] = M.d; }
  Mob<R> E = mob_shfl_up(M, 1);
  if (lane == 0) { E.a = 1; E.b = 0; E.c = 0; E.d = 1; }
  team_sync(bar_id, nthreads);
  // ---- F2: P path, mean aggregate ----
  {
    Mob<R> Pre{(R)1, (R)0, (R)0, (R)1};
    for (int t = 0; t < wt; ++t) {
      const Mob<R> A{ts->aggM[t][0], ts->aggM[t][1], ts->aggM[t][2], ts->aggM[t][3]};
      Pre = mob_mul(A, Pre);
    }
    E = mob_mul(E, Pre);
  }
  R Pc = fma(E.a, pr.P0, E.b) * Num<R>::rcp(fma(E.c, pr.P0, E.d));
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    B.P[k] = Pc;
    const bool o = (B.obs >> k) & 1u;
    const R rF = o ? Num<R>::rcp(Pc + s_e) : (R)0;
    const R K = Pc * rF;
    B.K[k] = K;
    Pc = fma(-K, Pc, Pc) + s_h;
  }
  R m = 1, c = 0;
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const R omk = (R)1 - B.K[k];
    c = fma(omk, c, B.K[k] * B.r[k]);
    m = omk * m;
  }
affine_scan_up(m, c, lane);
  if (lane == 31) { ts->aggA[wt][0] = m; ts->aggA[wt][1] = c; }
  R me = __shfl_up_sync(FULL, m, 1), ce = __shfl_up_sync(FULL, c, 1);
  if (lane == 0) { me = 1; ce = 0; }
  team_sync(bar_id, nthreads);
  // ---- F3: filtered means; sampling elements; reverse scan ----
  R a_in = pr.m0;
  for (int t = 0; t < wt; ++t) a_in = fma(ts->aggA[t][0], a_in, ts->aggA[t][1]);
  R ac = fma(me, a_in, ce);
  const int t0 = wt * TB + lane * KS;
  const uint32_t c0 = (uint32_t)gid, c1 = RNG_SMOOTH | ((uint32_t)(gid >> 32) << 8);
  R zs[KS], zp[KS];
#pragma unroll
  for (int k = 0; k < KS; k += 2) {
    const uint4 x = Philox::gen(seed, c0, c1, (uint32_t)((t0 + k) >> 1), 0u);
    box_muller<R>(x.x, x.y, zs[k], zp[k]);
    box_muller<R>(x.z, x.w, zs[k + 1], zp[k + 1]);
  }
  R J[KS], off_[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const R v = ((B.obs >> k) & 1u) ? (B.r[k] - ac) : (R)0;
    ac = fma(B.K[k], v, ac);                       // filtered mean m_k
    const R Cf = B.P[k] * ((R)1 - B.K[k]);
    const R Jk = (t0 + k < T - 1) ? Cf * Num<R>::rcp(Cf + s_h) : (R)0;
    const R Vk = Cf * ((R)1 - Jk);
    J[k] = Jk;
    off_[k] = fma((R)1 - Jk, ac, Num<R>::sqrt(Vk) * zs[k]);
  }
  m = 1; c = 0;
#pragma unroll
  for (int k = KS - 1; k >= 0; --k) { c = fma(J[k], c, off_[k]); m = J[k] * m; }
affine_scan_down(m, c, lane);
  if (lane == 0) { ts->aggAB[wt][0] = m; ts->aggAB[wt][1] = c; }
  me = __shfl_down_sync(FULL, m, 1); ce = __shfl_down_sync(FULL, c, 1);
  if (lane == 31) { me = 1; ce = 0; }
  team_sync(bar_id, nthreads);
  R x_in = 0;
  for (int t = W - 1; t > wt; --t) x_in = fma(ts->aggAB[t][0], x_in, ts->aggAB[t][1]);
  R x = fma(me, x_in, ce);
  R lv[KS], tr[KS];
#pragma unroll
  for (int k = KS - 1; k >= 0; --k) {
    x = fma(J[k], x, off_[k]);
    lv[k] = x;
    tr[k] = x + xw[k] + sig_e * zp[k];
  }
  if (level_row) store_run(level_row, t0, T, lv);
  store_run(traj_row, t0, T, tr);
}

}  // namespace ci
