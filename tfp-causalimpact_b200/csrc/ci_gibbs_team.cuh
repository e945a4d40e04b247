// ci_gibbs_team.cuh -- TEAM-MODE Gibbs sweep: the reference's sampler (spike-and-slab regression
// + FFBS level draw + InverseGamma draws; causalimpact/causalimpact_lib.py:365-388) with ONE WARP
// PER TILE of a chain instead of one warp per chain, for series whose tiles are resident in
// shared memory (T <= 2048: the reference's own use cases and BASELINE configs[1], [4]).
//
// What the W = NB warps of a chain's team do in a sweep:
//   A. spike-and-slab step.  The inclusion draws are sequential in law (feature j's conditional
//      depends on the indicators drawn before it), but a draw rarely CHANGES its indicator once
//      the chain has found the support, so the team SPECULATES: warp t evaluates the log-marginal
//      of "flip feature j0 + t" under the current configuration, all W at once (each a pair of
//      warp-cooperative Cholesky factorisations in the warp's own scratch); the draws are then
//      taken in order, identically by every warp, up to and including the first one that flips --
//      later candidates were evaluated under a configuration that no longer holds and are simply
//      re-evaluated in the next round.  Same Philox keys, same log-marginals, same decisions as
//      the one-warp kernel: p / W + (number of flips) evaluations deep instead of p.
//      sigma_obs^2 and the active weights are drawn by warp 0 and published to the team.
//   B. FFBS = the team simulation smoother of ci_team.cuh (forward filter with tile aggregates
//      exchanged through shared memory, reverse affine sampling scan exchanged once more): no
//      per-tile checkpoints, no forward replay -- one tile of latency per direction instead of
//      NB.  Every warp accumulates its tile's share of the next sweep's sufficient statistics.
//   C. sigma_level^2: every warp draws it redundantly from the team totals (same key, same value).
// Results differ from k_gibbs only by the summation order of the statistics; which kernel runs
// depends on the series' shape alone, never on the batch, so draws stay bit-identical under any
// split of the chains / series over launches and GPUs.
#pragma once
#include "ci_gibbs.cuh"
#include "ci_team.cuh"

namespace ci {

constexpr int GT_MAXWARPS = 16;        // warps per CTA (teams x W); no producer warp

template <typename R> struct GibbsTeamShared {
  TeamShared<R> ts;                    // tile aggregates, per-warp partials (red, gwpart)
  R w[MAX_DIM];                        // the sweep's weights (warp 0 -> team)
  double lm[2][MAXW];                  // log-marginals of a round's candidates (double-buffered)
  double s_e;                          // the sweep's sigma_obs^2 (warp 0 -> team)
};

// Step A for a team.  Every warp holds identical (gam, yty) and its own copy of bvec in
// reg.gs.bvec; returns with s_e and gt->w published (after a team barrier).
template <typename R>
__device__ __forceinline__ void gibbs_reg_step_team(const GibbsReg<R>& reg, GibbsTeamShared<R>* gt,
                                                    int it, const GibbsPlan& plan,
                                                    uint32_t (&gam)[4], double yty, double& s_e,
                                                    int wt, int W, int bar_id) {
  const int p = reg.p, lane = reg.lane, nthreads = 32 * W;
  if (p > 0 && plan.sparse) {
    reg.make_order(it, plan);                       // every warp its own copy: no barrier
    int k0; double ss0;
    double lm_cur = reg.log_marginal(gam, yty, k0, ss0);     // redundant in every warp: no barrier
    int j0 = 0, par = 0;
    while (j0 < p) {
      const int jj = j0 + wt;
      double lm_new = 0.0;
      if (jj < p) {
        const int j = reg.order(jj, plan);
        uint32_t gf[4] = {gam[0], gam[1], gam[2], gam[3]};
        gf[j >> 5] ^= 1u << (j & 31);
        int kf; double ssf;
        lm_new = reg.log_marginal(gf, yty, kf, ssf);
      }
      if (lane == 0) gt->lm[par][wt] = lm_new;
      team_sync(bar_id, nthreads);
      const int adv = min(W, p - j0);
      int taken = adv;
      for (int t = 0; t < adv; ++t) {
        const int j = reg.order(j0 + t, plan);
        const bool cur = (gam[j >> 5] >> (j & 31)) & 1u;
        const double lm_t = gt->lm[par][t];
        if (reg.flips(it, j, cur, lm_cur, lm_t, plan)) {
          gam[j >> 5] ^= 1u << (j & 31);
          lm_cur = lm_t;
          taken = t + 1;                            // the rest of the round is stale
          break;
        }
      }
      j0 += taken;
      par ^= 1;
    }
  }
  if (wt == 0) {
    double ss = yty;
    int k = 0;
    if (p > 0) (void)reg.log_marginal(gam, yty, k, ss);       // factor of the final configuration
    reg.draw(it, k, ss, s_e, gt->w);
    if (lane == 0) gt->s_e = s_e;
  }
  team_sync(bar_id, nthreads);
  s_e = gt->s_e;
}

// byte offset of the team areas = cfg.off_warp + n_warps * cfg.warp_bytes, then the CTA-shared
// Gram matrix
template <typename R>
__device__ __forceinline__ GibbsTeamShared<R>* gibbs_team_area(unsigned char* smem, const SmemCfg& cfg,
                                                               int n_warps, int team) {
  size_t off = (size_t)cfg.off_warp + (size_t)n_warps * cfg.warp_bytes;
  off = (off + 15) & ~(size_t)15;
  return reinterpret_cast<GibbsTeamShared<R>*>(smem + off) + team;
}

template <typename R>
__global__ void __launch_bounds__(32 * GT_MAXWARPS, 1)
k_gibbs_team(ProbDev<R> pr, GibbsDev<R> gd, SmemCfg cfg, GibbsPlan plan, int W, uint64_t seed,
             uint64_t chain_id0, int C, R* __restrict__ draws, R* __restrict__ level_out,
             R* __restrict__ traj_out, float* __restrict__ incl_out,
             const BatchDev<R>* __restrict__ batch) {
  extern __shared__ __align__(128) unsigned char smem[];
  const size_t series_row0 = batch ? (size_t)blockIdx.y * C * plan.n_results : 0;
  const size_t series_chain0 = batch ? (size_t)blockIdx.y * C : 0;
  if (batch) {
    pr = batch[blockIdx.y].pr; gd = batch[blockIdx.y].gd; plan.n_obs = batch[blockIdx.y].n_obs;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_warps = blockDim.x >> 5;
  const int GT = n_warps / W;
  const CtaShared<R> cs = cta_prologue(smem, cfg, pr, 1);
  const int p = pr.p, dim = pr.dim, ld = pr.ld, T = pr.T;
  // CTA-shared copy of the Gram matrix lives after the team areas
  size_t goff = (size_t)cfg.off_warp + (size_t)n_warps * cfg.warp_bytes;
  goff = ((goff + 15) & ~(size_t)15) + (size_t)GT * sizeof(GibbsTeamShared<R>);
  R* gram_s = reinterpret_cast<R*>(smem + ((goff + 15) & ~(size_t)15));
  for (int i = threadIdx.x; i < p * p; i += blockDim.x) gram_s[i] = gd.gram[i];
  if (threadIdx.x == 0) {
    omega_fetch(cs, pr);
    tile_producer(pr.tiles, cs.stage0, cs.full, cs.empty, cfg.stage_elems, cfg.nstage, pr.NB, true,
                  0LL, [](long long) { return true; });
  }
  __syncthreads();
  const int team = warp / W, wt = warp - team * W;
  const int c = blockIdx.x * GT + team;
  if (c >= C) return;
  omega_wait(cs);
  const R* om_s = cs.omega;
  const WarpScratch<R> ws = warp_scratch<R>(smem, cfg, warp);
  GibbsScratch<R> gs;
  gs.bvec = ws.extra; gs.La = gs.bvec + p; gs.Lo = gs.La + (p + 1) * (p + 1);
  gs.idx = gs.Lo + p * p; gs.vec = gs.idx + p; gs.perm = gs.vec + (p + 1);
  GibbsTeamShared<R>* gt = gibbs_team_area<R>(smem, cfg, n_warps, team);
  TeamShared<R>* ts = &gt->ts;
  const int bar_id = team + 1, nthreads = 32 * W;
  const uint64_t gid = chain_id0 + (uint64_t)c + (batch ? (uint64_t)blockIdx.y * plan.series_stride : 0ull);
  const uint32_t id_lo = (uint32_t)gid, id_hi8 = (uint32_t)(gid >> 32) << 8;
  mbar_wait(&cs.full[wt], 0u);
  const R* tile = cs.stage0 + (size_t)wt * cfg.stage_elems;
  const int t0 = wt * TB + lane * KS;

  // ---- initial state: the reference's (lib.py:566-581) ----
  double s_e = p > 0 ? 0.2 * (double)pr.P0 : (double)pr.P0;
  double s_h = (double)pr.lvl_scale / (double)pr.lvl_conc;
  uint32_t gam[4] = {0u, 0u, 0u, 0u};
  if (!plan.sparse)
    for (int j = 0; j < p; ++j) gam[j >> 5] |= 1u << (j & 31);
  for (int j = lane; j < p; j += 32) gs.bvec[j] = gd.xty0[j];
  double yty = (double)gd.yty0;
  double incl_cnt[DSLOTS] = {0.0, 0.0, 0.0, 0.0};
  const double conc_e = (double)pr.obs_conc + 0.5 * plan.n_obs;
  __syncwarp();
  const GibbsReg<R> reg{pr, gs, gram_s, om_s, p, lane, conc_e, seed, id_lo, id_hi8};
  const int n_iter = plan.n_warmup + plan.n_results;
  const XtMap xm = xt_map(p, lane);
  const bool small_p = p <= PSMALL;

  for (int it = 0; it < n_iter; ++it) {
    // =================== A. regression block ===================
    gibbs_reg_step_team(reg, gt, it, plan, gam, yty, s_e, wt, W, bar_id);
    const R* w_s = gt->w;
    // =================== B. level | rest : team FFBS ===================
    const R se = (R)s_e, sh = (R)s_h, sig_e = (R)sqrt(s_e);
    Blk<R> B;
    R xw[KS];
    blk_residuals_xw(B, xw, tile, w_s, p, ld, lane);
    // ---- F1: variance aggregate ----
    const R alpha = se + sh, beta = se * sh;
    Mob<R> M{(R)1, (R)0, (R)0, (R)1};
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const bool o = (B.obs >> k) & 1u;
      const R e1 = o ? alpha : (R)1, e2 = o ? beta : sh;
      const R f1 = o ? (R)1 : (R)0, f2 = o ? se : (R)1;
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
    // ---- F2: variances / gains, mean aggregate ----
    {
      R Pt = pr.P0;
      for (int t = 0; t < wt; ++t)
        Pt = fma(ts->aggM[t][0], Pt, ts->aggM[t][1]) * Num<R>::rcp(fma(ts->aggM[t][2], Pt, ts->aggM[t][3]));
      R Pc = fma(E.a, Pt, E.b) * Num<R>::rcp(fma(E.c, Pt, E.d));
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        B.P[k] = Pc;
        const bool o = (B.obs >> k) & 1u;
        const R rF = o ? Num<R>::rcp(Pc + se) : (R)0;
        const R K = Pc * rF;
        B.K[k] = K;
        Pc = fma(-K, Pc, Pc) + sh;
      }
    }
    R m = 1, cc = 0;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const R omk = (R)1 - B.K[k];
      cc = fma(omk, cc, B.K[k] * B.r[k]);
      m = omk * m;
    }
    affine_scan_up(m, cc, lane);
    if (lane == 31) { ts->aggA[wt][0] = m; ts->aggA[wt][1] = cc; }
    R me = __shfl_up_sync(FULL, m, 1), ce = __shfl_up_sync(FULL, cc, 1);
    if (lane == 0) { me = 1; ce = 0; }
    team_sync(bar_id, nthreads);
    // ---- F3: filtered means; sampling elements; reverse scan ----
    R a_in = pr.m0;
    for (int t = 0; t < wt; ++t) a_in = fma(ts->aggA[t][0], a_in, ts->aggA[t][1]);
    R ac = fma(me, a_in, ce);
    R zs[KS], zp[KS];
#pragma unroll
    for (int kk = 0; kk < KS; kk += 2) {
      const uint4 x = Philox::gen(seed, id_lo, RNG_SMOOTH | id_hi8, (uint32_t)((t0 + kk) >> 1),
                                  (uint32_t)it);
      box_muller<R>(x.x, x.y, zs[kk], zp[kk]);
      box_muller<R>(x.z, x.w, zs[kk + 1], zp[kk + 1]);
    }
    R J[KS], off[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const R v = ((B.obs >> k) & 1u) ? (B.r[k] - ac) : (R)0;
      ac = fma(B.K[k], v, ac);                       // filtered mean m_k
      const R Cf = B.P[k] * ((R)1 - B.K[k]);
      const R Jk = (t0 + k < T - 1) ? Cf * Num<R>::rcp(Cf + sh) : (R)0;
      const R Vk = Cf * ((R)1 - Jk);
      J[k] = Jk;
      off[k] = fma((R)1 - Jk, ac, Num<R>::sqrt(Vk) * zs[k]);
    }
    m = 1; cc = 0;
#pragma unroll
    for (int k = KS - 1; k >= 0; --k) { cc = fma(J[k], cc, off[k]); m = J[k] * m; }
    affine_scan_down(m, cc, lane);
    if (lane == 0) { ts->aggAB[wt][0] = m; ts->aggAB[wt][1] = cc; }
    me = __shfl_down_sync(FULL, m, 1); ce = __shfl_down_sync(FULL, cc, 1);
    if (lane == 31) { me = 1; ce = 0; }
    team_sync(bar_id, nthreads);
    R x_in = 0;
    for (int t = W - 1; t > wt; --t) x_in = fma(ts->aggAB[t][0], x_in, ts->aggAB[t][1]);
    R x = fma(me, x_in, ce);
    // ---- level path of the tile + this warp's share of the next sweep's statistics ----
    const bool keep = it >= plan.n_warmup;
    const size_t out_row = !keep ? 0
        : series_row0 + (plan.chain_major ? ((size_t)c * plan.n_results + (size_t)(it - plan.n_warmup))
                                          : ((size_t)(it - plan.n_warmup) * C + c));
    R lv[KS], tgt[KS];
    R ly = 0, ld2 = 0;
#pragma unroll
    for (int kk = KS - 1; kk >= 0; --kk) {
      const R xn = x;                       // level at t+1
      x = fma(J[kk], x, off[kk]);
      lv[kk] = x;
      const int t = t0 + kk;
      if (t + 1 < T) { const R dl = xn - x; ld2 = fma(dl, dl, ld2); }
      const bool o = (B.obs >> kk) & 1u;
      tgt[kk] = o ? (B.r[kk] + xw[kk] - x) : (R)0;       // y - level on observed steps
      ly = fma(tgt[kk], tgt[kk], ly);
    }
    if (keep) {
      R tr[KS];
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) tr[kk] = lv[kk] + xw[kk] + sig_e * zp[kk];
      if (level_out) store_run(level_out + out_row * T, t0, T, lv);
      if (traj_out) store_run(traj_out + out_row * T, t0, T, tr);
    }
    if (p > 0) {
      if (small_p) {
        R accw[PSMALL];
#pragma unroll
        for (int j = 0; j < PSMALL; ++j) accw[j] = 0;
        blk_xt_rbar_small(tile, tgt, p, ld, lane, accw);
        static_assert(PSMALL == 16, "warp_multi_sum16");
        warp_multi_sum16(accw, lane);
        if (!(lane & 1) && (lane >> 1) < p) ts->gwpart[wt][lane >> 1] = accw[0];
      } else {
        R accg[JS];
#pragma unroll
        for (int s = 0; s < JS; ++s) accg[s] = 0;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) ws.rbuf[lane * KS + kk + (lane >> 2)] = tgt[kk];
        __syncwarp();
        blk_xt_rbar<R, JS>(tile, ws.rbuf, p, ld, xm.jj, xm.part, xm.nparts, accg);
        __syncwarp();
#pragma unroll
        for (int s = 0; s < JS; ++s) {
          R a = accg[s];
          for (int o = xm.PJ; o < 32; o <<= 1) a += __shfl_xor_sync(FULL, a, o);
          const int j = lane + 32 * s;
          if (j < p) ts->gwpart[wt][j] = a;
        }
      }
    }
    {
      const double y_w = warp_sum((double)ly), d_w = warp_sum((double)ld2);
      if (lane == 0) { ts->red[wt][0] = y_w; ts->red[wt][1] = d_w; }
    }
    team_sync(bar_id, nthreads);
    // team totals in fixed order, identical in every warp
    double d2 = 0.0;
    yty = 0.0;
    for (int t = 0; t < W; ++t) { yty += ts->red[t][0]; d2 += ts->red[t][1]; }
    for (int j = lane; j < p; j += 32) {
      R a = 0;
      for (int t = 0; t < W; ++t) a += ts->gwpart[t][j];
      gs.bvec[j] = a;                       // this warp's own copy
    }
    __syncwarp();
    // =================== C. sigma_level^2 ===================
    {
      const double g = gamma_draw((double)pr.lvl_conc + 0.5 * (T - 1), seed, id_lo,
                                  RNG_G_GAMMA | id_hi8, (uint32_t)it, 1u);
      s_h = ((double)pr.lvl_scale + 0.5 * d2) / g;
      const double ub2 = (double)pr.lvl_ub;                            // variance bound
      if (s_h > ub2) s_h = ub2;                                        // lib.py:432
    }
    if (keep && wt == 0) {
      R* row = draws + out_row * dim;
      for (int j = lane; j < p; j += 32) row[j] = w_s[j];
      if (lane == 0) { row[p] = (R)log(s_e); row[p + 1] = (R)log(s_h); }
    }
    if (keep) {
#pragma unroll
      for (int wd = 0; wd < DSLOTS; ++wd) incl_cnt[wd] += (double)((gam[wd] >> lane) & 1u);
    }
    // (the next sweep's first write to the exchange area -- lm[] or w -- comes after its own
    // barrier-separated reads; red / gwpart are rewritten only after three more barriers)
  }
  if (incl_out && wt == 0) {
#pragma unroll
    for (int wd = 0; wd < DSLOTS; ++wd) {
      const int j = lane + 32 * wd;
      if (j < p) incl_out[(series_chain0 + (size_t)c) * p + j] = (float)(incl_cnt[wd] / (plan.n_results > 0 ? plan.n_results : 1));
    }
  }
}

}  // namespace ci
