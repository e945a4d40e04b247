// ci_team4.cuh -- TEAM MODE with SUB-TILES: the evaluation of ci_team.cuh with KQ
// consecutive steps per lane instead of 8, i.e. 32 KQ steps per warp and 8 / KQ warps per
// 256-step tile.  KQ = 4 halves every per-lane serial recursion (the Moebius composition,
// the sequential variance / gain pass with its reciprocals, the mean and adjoint
// recursions, the residual and X' rbar dot products) at the price of twice the warps per
// chain; the warp-shuffle scans and the cross-warp exchange keep their depth.  The kernel
// is bound by that dependent chain, not by issue slots (ncu: 10 % warp occupancy, IPC
// 0.66), so the extra warps are free.  Same tiles in shared memory: a lane's KQ rows
// never straddle a 32-row pad boundary, and the row stride stays odd, so the residual
// reads remain bank-conflict free.  Small-p only (p <= PSMALL): one register per covariate.
//
// Replaces the same reference arithmetic as ci_filter.cuh (TFP LGSSM log_prob, call site
// causalimpact/causalimpact_lib.py:365-388).
#pragma once
#include "ci_team.cuh"

namespace ci {

template <typename R, int KQ> struct BlkQ {
  R r[KQ], P[KQ], K[KQ], rF[KQ], v[KQ];
  uint32_t obs;
};

template <typename R> __device__ __forceinline__ Mob<R> mob_mul_raw(const Mob<R>& L, const Mob<R>& E) {
  Mob<R> o;
  o.a = L.a * E.a + L.b * E.c; o.b = L.a * E.b + L.b * E.d;
  o.c = L.c * E.a + L.d * E.c; o.d = L.c * E.b + L.d * E.d;
  return o;
}

// r_k = y_k - sum_j x_kj w_j for rows first_row .. first_row + KQ - 1 of the tile
template <typename R, int KQ>
__device__ __forceinline__ void blkq_residuals(BlkQ<R, KQ>& B, const R* __restrict__ tile,
                                               const R* __restrict__ w_s, int p, int ld,
                                               int first_row) {
  const R* row0 = tile + tile_off(first_row, ld);
  R acc[KQ];
#pragma unroll
  for (int k = 0; k < KQ; ++k) acc[k] = row0[k * ld + p];
#pragma unroll 4
  for (int j = 0; j < p; ++j) {
    const R wj = w_s[j];
#pragma unroll
    for (int k = 0; k < KQ; ++k) acc[k] = fma(-row0[k * ld + j], wj, acc[k]);
  }
  uint32_t obs = 0;
#pragma unroll
  for (int k = 0; k < KQ; ++k) {
    const bool o = (acc[k] == acc[k]);
    obs |= (o ? 1u : 0u) << k;
    B.r[k] = o ? acc[k] : (R)0;
  }
  B.obs = obs;
}

template <typename R, int KQ>
__device__ __forceinline__ R blkq_loglik_terms(const BlkQ<R, KQ>& B, R s_e) {
  static_assert(KQ % 4 == 0, "logs are taken of products of 4 innovation variances");
  R s = 0;
  R prod[KQ / 4];
#pragma unroll
  for (int h = 0; h < KQ / 4; ++h) prod[h] = 1;
#pragma unroll
  for (int k = 0; k < KQ; ++k) {
    const bool o = (B.obs >> k) & 1u;
    prod[k >> 2] *= o ? (B.P[k] + s_e) : (R)1;
    s = fma(B.v[k] * B.v[k], B.rF[k], s);
  }
#pragma unroll
  for (int h = 0; h < KQ / 4; ++h) s += Num<R>::log(prod[h]);
  return s;
}

// Evaluate the chain whose weights are in w_s.  wt = this warp's index in the team: it owns
// rows row_base .. row_base + 32 KQ - 1 of `tile`.  Results as team_eval.
template <typename R, int KQ>
__device__ __forceinline__ void team_eval_q(const R* __restrict__ tile, const ProbDev<R>& pr,
                                          TeamShared<R>* ts, const R* __restrict__ w_s, R* rbuf,
                                          R s_e, R s_h, bool want_grad, int lane, int wt, int W,
                                          int bar_id, int row_base, double& ll, double& g_se, double& g_sh,
                                          R (&gw)[JS]) {
  const int p = pr.p, ld = pr.ld;
  const int nthreads = 32 * W;
  BlkQ<R, KQ> B;
  blkq_residuals<R, KQ>(B, tile, w_s, p, ld, row_base + lane * KQ);

  // ---------------- F1: variance path, tile aggregate ----------------
  const R alpha = s_e + s_h, beta = s_e * s_h;
  Mob<R> M{(R)1, (R)0, (R)0, (R)1};
#pragma unroll
  for (int k = 0; k < KQ; ++k) {
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

  // ---------------- F2: carry-in P, sequential P; mean aggregate ----------------
  {
    Mob<R> Pre{(R)1, (R)0, (R)0, (R)1};
    for (int t = 0; t < wt; ++t) {
      const Mob<R> A{ts->aggM[t][0], ts->aggM[t][1], ts->aggM[t][2], ts->aggM[t][3]};
      Pre = mob_mul_raw(A, Pre);               // projective: scale-free, normalise once below
    }
    E = mob_mul(E, Pre);
  }
  R Pc = fma(E.a, pr.P0, E.b) * Num<R>::rcp(fma(E.c, pr.P0, E.d));
#pragma unroll
  for (int k = 0; k < KQ; ++k) {
    B.P[k] = Pc;
    const bool o = (B.obs >> k) & 1u;
    const R rF = o ? Num<R>::rcp(Pc + s_e) : (R)0;
    const R K = Pc * rF;
    B.rF[k] = rF; B.K[k] = K;
    Pc = fma(-K, Pc, Pc) + s_h;
  }
  R m = 1, c = 0;
#pragma unroll
  for (int k = 0; k < KQ; ++k) {
    const R omk = (R)1 - B.K[k];
    c = fma(omk, c, B.K[k] * B.r[k]);
    m = omk * m;
  }
affine_scan_up(m, c, lane);
  if (lane == 31) { ts->aggA[wt][0] = m; ts->aggA[wt][1] = c; }
  R me = __shfl_up_sync(FULL, m, 1), ce = __shfl_up_sync(FULL, c, 1);
  if (lane == 0) { me = 1; ce = 0; }
  team_sync(bar_id, nthreads);

  // ---------------- F3: carry-in a, innovations, log-lik terms ----------------
  R a_in = pr.m0;
  for (int t = 0; t < wt; ++t) a_in = fma(ts->aggA[t][0], a_in, ts->aggA[t][1]);
  R ac = fma(me, a_in, ce);
#pragma unroll
  for (int k = 0; k < KQ; ++k) {
    const R v = ((B.obs >> k) & 1u) ? (B.r[k] - ac) : (R)0;
    ac = fma(B.K[k], v, ac);
    B.v[k] = v;
  }
  const double ll_terms = warp_sum((double)blkq_loglik_terms<R, KQ>(B, s_e));
  const int n_obs = __reduce_add_sync(FULL, __popc(B.obs));

  R abn[KQ], q[KQ], dF[KQ], rbar[KQ];
  R lge = 0, lgh = 0;
  if (want_grad) {
    // ---------------- B1: abar reverse scan, tile aggregate ----------------
    m = 1; c = 0;
#pragma unroll
    for (int k = KQ - 1; k >= 0; --k) {
      const R omk = (R)1 - B.K[k];
      c = fma(omk, c, B.v[k] * B.rF[k]);
      m = omk * m;
    }
affine_scan_down(m, c, lane);
    if (lane == 0) { ts->aggAB[wt][0] = m; ts->aggAB[wt][1] = c; }
    me = __shfl_down_sync(FULL, m, 1); ce = __shfl_down_sync(FULL, c, 1);
    if (lane == 31) { me = 1; ce = 0; }
    team_sync(bar_id, nthreads);
    R ab_in = 0;
    for (int t = W - 1; t > wt; --t) ab_in = fma(ts->aggAB[t][0], ab_in, ts->aggAB[t][1]);
    R ab = fma(me, ab_in, ce);
#pragma unroll
    for (int k = KQ - 1; k >= 0; --k) {
      abn[k] = ab;
      ab = fma((R)1 - B.K[k], ab, B.v[k] * B.rF[k]);
    }
    // ---------------- B2: Pbar reverse scan ----------------
    m = 1; c = 0;
#pragma unroll
    for (int k = KQ - 1; k >= 0; --k) {
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
    team_sync(bar_id, nthreads);
    R pb_in = 0;
    for (int t = W - 1; t > wt; --t) pb_in = fma(ts->aggPB[t][0], pb_in, ts->aggPB[t][1]);
    R pb = fma(me, pb_in, ce);
#pragma unroll
    for (int k = KQ - 1; k >= 0; --k) {
      const R K = B.K[k], rF = B.rF[k], v = B.v[k];
      const R omk = (R)1 - K;
      lgh += pb;
      lge += fma(K * K, pb, dF[k]) - abn[k] * v * B.P[k] * rF * rF;
      rbar[k] = fma(K, abn[k], -v * rF);
      pb = fma(omk * omk, pb, q[k]);
    }
    // ---------------- X^T rbar of this sub-tile (p <= PSMALL only) ----------------
    if (p > 0) {
      R accw[PSMALL];
#pragma unroll
      for (int j = 0; j < PSMALL; ++j) accw[j] = 0;
      const R* row0 = tile + tile_off(row_base + lane * KQ, ld);
#pragma unroll
      for (int j = 0; j < PSMALL; ++j) {
        if (j < p) {
#pragma unroll
          for (int k = 0; k < KQ; ++k) accw[j] = fma(rbar[k], row0[k * ld + j], accw[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < PSMALL; ++j) {
        if (j < p) {
          const R tot = warp_sum(accw[j]);
          if (lane == j) ts->gwpart[wt][j] = tot;
        }
      }
    }
  }
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
}


}  // namespace ci
