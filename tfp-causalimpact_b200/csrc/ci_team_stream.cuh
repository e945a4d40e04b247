// ci_team_stream.cuh -- TEAM MODE FOR LONG SERIES: a chain is evaluated by W warps that
// walk the series in ROUNDS of W tiles (tile b = round * W + warp), for series with more
// tiles than a team has warps (T > 2048) and/or tiles that do not all fit in shared memory
// (BASELINE.json configs[3]: T = 20 000, 79 tiles, streamed through the mbarrier ring).
//
// It is the two-level associative scan of ci_team.cuh with a third, sequential level:
//   level 1  warp shuffles inside a tile (ci_filter.cuh)
//   level 2  tile aggregates (Moebius matrix / affine pairs) of the W tiles of a round,
//            exchanged through shared memory with a named barrier
//   level 3  the round's carry (a, P) / (abar, Pbar) in registers, identical in every warp
// Forward rounds run 0 .. NR-1, adjoint rounds NR-1 .. 0.  Nothing per-step is written to
// HBM and -- unlike the one-warp path (ci_device.cuh:chain_eval), whose backward sweep
// replays each tile's whole forward scan from a per-TILE checkpoint -- the forward sweep
// leaves a per-LANE checkpoint (P and a at the lane's first step: 64 values per tile in
// shared memory), so the adjoint sweep re-derives a tile's gains and innovations with 8
// dependent steps per lane: no Moebius scan, no mean scan, no barrier.  W warps per chain
// turn the 2 NB sequential tile latencies of the one-warp path into 2 NB / W round
// latencies and put 8-16 warps on every SM at 512 chains.
//
// Replaces the same reference arithmetic as ci_filter.cuh (TFP LGSSM log_prob inside the
// sampler loop, call site causalimpact/causalimpact_lib.py:365-388).
#pragma once
#include "ci_hmc.cuh"
#include "ci_team.cuh"

namespace ci {

constexpr int TS_MAXWARPS = 16;   // warps per CTA (teams x W); there is NO producer warp: 512 threads
                                  // keep the 128-register budget (a 17th warp cut it to 96 and the
                                  // spills were the top stall, ncu run r2_02)

constexpr int TS_W = 4;           // default warps per chain (tuned on configs[3], run r2_03)

template <typename R> struct TeamStreamShared {
  R aggM[2][MAXW][4];    // Moebius aggregate of each tile of the round (double-buffered by
  R aggA[2][MAXW][2];    // round parity: a round's writes never meet the previous round's
  R aggAB[2][MAXW][2];   // late readers, so two barriers per round suffice)
  R aggPB[2][MAXW][2];
  R carryP[2];           // predicted variance after the round's last step
  double red[MAXW][4];   // ll-terms, ge, gh, n_obs partials
  R gwpart[MAXW][MAX_DIM];
};

// A warp's view of the tile ring.  Unlike TilePipe (one consumer walks every tile) a team
// warp only takes the tiles of its own rounds, so the stage / parity of a tile come from
// its position `seq` in the producer's stream instead of a running counter.
template <typename R> struct RingView {
  R* stage0;
  uint64_t* full;
  uint64_t* empty;
  const R* gtiles;         // the series' tiles in global memory
  uint32_t stage_elems, nstage;
  uint32_t ahead;          // a finished tile at stream position q triggers the load of q + ahead
  uint32_t total;          // number of tiles in the whole stream (sweeps x NB)
  int NB;
  bool resident, fetcher;  // fetcher: this warp belongs to team 0, which issues the bulk copies
  bool alternate;          // sweeps alternate direction (value + gradient); else all forward
  // tile index at stream position seq (start-up only: one division)
  __device__ __forceinline__ int tile_at(uint32_t seq) const {
    const uint32_t sw = seq / (uint32_t)NB, pos = seq - sw * (uint32_t)NB;
    return (alternate && (sw & 1u)) ? NB - 1 - (int)pos : (int)pos;
  }
  // one thread: bulk copy of tile `tile` into stage st; `wait_par` < 2: first wait until every
  // team has released the stage's previous occupant (phase parity wait_par)
  __device__ __forceinline__ void fetch_to(uint32_t st, int tile, uint32_t wait_par) const {
    const uint32_t bytes = stage_elems * (uint32_t)sizeof(R);
    if (wait_par < 2u) mbar_wait(&empty[st], wait_par);
    mbar_expect_tx(&full[st], bytes);
    bulk_g2s(stage0 + (size_t)st * stage_elems, tile_src(gtiles, (uint32_t)tile, bytes), bytes, &full[st]);
  }
  __device__ __forceinline__ void fetch(uint32_t seq) const {      // start-up
    fetch_to(seq % nstage, tile_at(seq), seq >= nstage ? ((seq / nstage) - 1u) & 1u : 2u);
  }
  __device__ __forceinline__ const R* stage(uint32_t st) const { return stage0 + (size_t)st * stage_elems; }
};

// A warp's position in the stream: stage and phase parity of its current tile, advanced by W
// positions per round WITHOUT integer division (the first version spent 3 % of its instructions
// on seq % nstage, ncu r2_05); recomputed with one division at the start of a sweep.
struct RingPos {
  uint32_t st, par;
  __device__ __forceinline__ void set(uint32_t seq, uint32_t nstage) {
    const uint32_t w = seq / nstage;
    st = seq - w * nstage; par = w & 1u;
  }
  __device__ __forceinline__ void advance(uint32_t by, uint32_t nstage) {
    st += by;
    if (st >= nstage) { st -= nstage; par ^= 1u; }
  }
};

template <typename R>
__device__ __forceinline__ const R* ring_acquire(const RingView<R>& ring, const RingPos& pos, int tile) {
  if (ring.resident) { mbar_wait(&ring.full[tile], 0u); return ring.stage((uint32_t)tile); }
  mbar_wait(&ring.full[pos.st], pos.par);
  return ring.stage(pos.st);
}
// Done with the tile at stream position seq (stage pos.st).  Team 0 also keeps the ring full: it
// issues the copy of the tile `ahead` = nstage - W positions further down the stream, which lands
// in stage pos.st - W (mod nstage).  fwd / b: direction of the current sweep and the tile just
// finished, from which the tile at seq + ahead follows by comparison (at most one sweep boundary
// lies in between: the planner keeps ahead < NB when streaming).
template <typename R>
__device__ __forceinline__ void ring_release(const RingView<R>& ring, const RingPos& pos, uint32_t seq,
                                             bool fwd, int b, int W, int lane) {
  if (ring.resident) return;
  __syncwarp();
  if (lane != 0) return;
  mbar_arrive(&ring.empty[pos.st]);
  if (!ring.fetcher || seq + ring.ahead >= ring.total) return;
  const int NB = ring.NB, ahead = (int)ring.ahead;
  int tile;
  if (fwd) { const int q = b + ahead; tile = q < NB ? q : (ring.alternate ? 2 * NB - 1 - q : q - NB); }
  else { const int q = (NB - 1 - b) + ahead; tile = q < NB ? NB - 1 - q : q - NB; }
  // seq + ahead = seq + nstage - W: same phase count as seq when pos.st < W, one more otherwise
  const bool wrap = pos.st >= (uint32_t)W;
  const uint32_t st2 = wrap ? pos.st - (uint32_t)W : pos.st + ring.nstage - (uint32_t)W;
  const uint32_t par2 = wrap ? pos.par : pos.par ^ 1u;       // parity of (phase count of seq+ahead) - 1
  const bool first_use = !wrap && seq < ring.nstage;          // stage never used before: nothing to wait for
  ring.fetch_to(st2, tile, first_use ? 2u : par2);
}

// Carry-in of warp wt (x_in) and of the next round (x_tot) from the round's affine aggregates
// agg[t] = (m, c), x' = m x + c.  All 2 MAXW values are loaded first (independent shared-memory
// loads) and the fold runs on registers: the dependent chain is W FMAs, not W load latencies.
template <typename R, int WC>
__device__ __forceinline__ void fold_up(const R (*agg)[2], R x, int wt, int W, R& x_in, R& x_tot) {
  constexpr int NW = WC ? WC : MAXW;
  R m[NW], c[NW];
#pragma unroll
  for (int t = 0; t < NW; ++t) { m[t] = t < W ? agg[t][0] : (R)1; c[t] = t < W ? agg[t][1] : (R)0; }
  x_in = x;
#pragma unroll
  for (int t = 0; t < NW; ++t) {
    if (t == wt) x_in = x;
    x = fma(m[t], x, c[t]);
  }
  x_tot = x;
}
template <typename R, int WC>
__device__ __forceinline__ void fold_down(const R (*agg)[2], R x, int wt, int W, R& x_in, R& x_tot) {
  constexpr int NW = WC ? WC : MAXW;
  R m[NW], c[NW];
#pragma unroll
  for (int t = 0; t < NW; ++t) { m[t] = t < W ? agg[t][0] : (R)1; c[t] = t < W ? agg[t][1] : (R)0; }
  x_in = x;
#pragma unroll
  for (int t = NW - 1; t >= 0; --t) {
    if (t == wt) x_in = x;
    x = fma(m[t], x, c[t]);
  }
  x_tot = x;
}

// gains of a lane's KS steps from the predicted variance at its first step (B.P is not kept:
// P rF = K and F = 1 / rF are all the adjoint and the log-likelihood need)
template <typename R>
__device__ __forceinline__ R blk_gains_seq(Blk<R>& B, R Pc, R s_e, R s_h) {
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const bool o = (B.obs >> k) & 1u;
    const R rF = o ? Num<R>::rcp(Pc + s_e) : (R)0;
    const R K = Pc * rF;
    B.rF[k] = rF; B.K[k] = K;
    Pc = fma(-K, Pc, Pc) + s_h;
  }
  return Pc;
}

// innovations of a lane's steps from the predicted mean at its first step.  v of a masked step is
// a finite don't-care: every use multiplies it by that step's K = 0 or rF = 0.
template <typename R> __device__ __forceinline__ void blk_innov_seq(Blk<R>& B, R ac) {
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const R v = B.r[k] - ac;
    ac = fma(B.K[k], v, ac);
    B.v[k] = v;
  }
}
// sum over the lane's observed steps of  log F + v^2 / F  from rF = 1 / F (logs of 4-products)
template <typename R> __device__ __forceinline__ R blk_loglik_terms_rf(const Blk<R>& B) {
  R s = 0;
  R prod[KS / 4];
#pragma unroll
  for (int h = 0; h < KS / 4; ++h) prod[h] = 1;
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const bool o = (B.obs >> k) & 1u;
    prod[k >> 2] *= o ? B.rF[k] : (R)1;
    s = fma(B.v[k] * B.v[k], B.rF[k], s);
  }
  // (float: MUFU.LG2-based log, |error| ~ 1e-7 per 4-product, random over the series: ~1e-5
  // absolute on a T = 20 000 log-likelihood of ~1e4; logf was 3.4 % of the kernel, ncu r2_05)
#pragma unroll
  for (int h = 0; h < KS / 4; ++h) s -= Num<R>::log_fast(prod[h]);
  return s;
}
// residuals with the number of regression columns known at compile time (PC > 0): the row
// stride and every loop bound are constants
template <typename R, int PC>
__device__ __forceinline__ void blk_residuals_c(Blk<R>& B, const R* __restrict__ tile,
                                                const R* __restrict__ w_s, int p, int ld, int lane) {
  if (PC == 0) { blk_residuals(B, tile, w_s, p, ld, lane); return; }
  constexpr int LD = (PC + 1) | 1;
  const R* row0 = tile + tile_off(lane * KS, LD);
  R acc[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) acc[k] = row0[k * LD + PC];
#pragma unroll
  for (int j = 0; j < (PC ? PC : 1); ++j) {
    const R wj = w_s[j];
#pragma unroll
    for (int k = 0; k < KS; ++k) acc[k] = fma(-row0[k * LD + j], wj, acc[k]);
  }
  uint32_t obs = 0;
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const bool o = (acc[k] == acc[k]);
    obs |= (o ? 1u : 0u) << k;
    B.r[k] = o ? acc[k] : (R)0;
  }
  B.obs = obs;
}

// Evaluate the chain whose weights are in w_s.  `seq0` = position of this evaluation's first
// tile in the producer's stream (advanced on return).  lck: [NB][64] per-lane checkpoints of
// the team.  On return every warp of the team holds identical results.
// PC > 0: the number of regression columns is PC at compile time (small-p instantiations);
// WC > 0: the team has WC warps at compile time.  0 = run-time value.
template <typename R, int PC, int WC>
__device__ __forceinline__ void team_stream_eval(const RingView<R>& ring, uint32_t& seq0,
                                                 const ProbDev<R>& pr, TeamStreamShared<R>* ts,
                                                 R* __restrict__ lck, const R* __restrict__ w_s,
                                                 R* rbuf, R s_e, R s_h, bool want_grad, int lane,
                                                 int wt, int W_rt, int bar_id, double& ll,
                                                 double& g_se, double& g_sh, R (&gw)[JS]) {
  const int p = PC ? PC : pr.p, ld = PC ? ((PC + 1) | 1) : pr.ld, NB = pr.NB;
  const int W = WC ? WC : W_rt;
  const int NR = (NB + W - 1) / W;
  const int nthreads = 32 * W;
  const R alpha = s_e + s_h, beta = s_e * s_h;

  // ======================= forward rounds =======================
  R a_c = pr.m0, P_c = pr.P0;
  double acc_ll = 0.0;
  int n_obs = 0;
  RingPos pos;
  pos.set(seq0 + (uint32_t)wt, ring.nstage);
  for (int r = 0; r < NR; ++r, pos.advance((uint32_t)W, ring.nstage)) {
    const int par = r & 1;
    const int b = r * W + wt;
    const bool act = b < NB;
    Blk<R> B;
    B.obs = 0u;
    Mob<R> M{(R)1, (R)0, (R)0, (R)1};
    if (act) {
      const R* tile = ring_acquire(ring, pos, b);
      blk_residuals_c<R, PC>(B, tile, w_s, p, ld, lane);
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
      const R s = Num<R>::rcp_fast(M.a + M.b + M.c + M.d);
      M.a *= s; M.b *= s; M.c *= s; M.d *= s;
      mob_scan_up(M, lane);
    }
    if (lane == 31) {
      ts->aggM[par][wt][0] = M.a; ts->aggM[par][wt][1] = M.b;
      ts->aggM[par][wt][2] = M.c; ts->aggM[par][wt][3] = M.d;
    }
    Mob<R> E = mob_shfl_up(M, 1);
    if (lane == 0) { E.a = 1; E.b = 0; E.c = 0; E.d = 1; }
    team_sync(bar_id, nthreads);
    R m = 1, c = 0;
    if (act) {
      // predicted variance at the tile's first step: the earlier tiles' maps applied to the
      // round's carry one after the other (a scalar Moebius image each: cheaper than composing
      // the matrices), then this lane's prefix inside the tile
      R Pt = P_c;
      for (int t = 0; t < wt; ++t)
        Pt = fma(ts->aggM[par][t][0], Pt, ts->aggM[par][t][1]) *
             Num<R>::rcp(fma(ts->aggM[par][t][2], Pt, ts->aggM[par][t][3]));
      const R Pl = fma(E.a, Pt, E.b) * Num<R>::rcp(fma(E.c, Pt, E.d));
      if (want_grad) lck[b * 64 + lane] = Pl;
      const R Pend = blk_gains_seq(B, Pl, s_e, s_h);
      if (wt == W - 1 && lane == 31) ts->carryP[par] = Pend;
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const R omk = (R)1 - B.K[k];
        c = fma(omk, c, B.K[k] * B.r[k]);
        m = omk * m;
      }
      affine_scan_up(m, c, lane);
    }
    if (lane == 31) { ts->aggA[par][wt][0] = m; ts->aggA[par][wt][1] = c; }
    R me = __shfl_up_sync(FULL, m, 1), ce = __shfl_up_sync(FULL, c, 1);
    if (lane == 0) { me = 1; ce = 0; }
    team_sync(bar_id, nthreads);
    R a_in, a_tot;
    fold_up<R, WC>(ts->aggA[par], a_c, wt, W, a_in, a_tot);
    if (act) {
      const R ac = fma(me, a_in, ce);
      if (want_grad) lck[b * 64 + 32 + lane] = ac;
      blk_innov_seq(B, ac);
      acc_ll += (double)blk_loglik_terms_rf(B);
      n_obs += __popc(B.obs);
      ring_release(ring, pos, seq0 + (uint32_t)b, true, b, W, lane);
    }
    a_c = a_tot;
    if (r + 1 < NR) P_c = ts->carryP[par];     // (a round followed by another one is full)
  }
  seq0 += (uint32_t)NB;

  // ======================= adjoint rounds =======================
  double ge_d = 0.0, gh_d = 0.0;
  constexpr int NA = PC ? PC : PSMALL;       // covariate accumulators of the lane <-> time mapping
  R accw[NA];
#pragma unroll
  for (int j = 0; j < NA; ++j) accw[j] = 0;
  R accg[JS];
#pragma unroll
  for (int s = 0; s < JS; ++s) accg[s] = 0;
  const bool small_p = p <= PSMALL;
  const XtMap xm = xt_map(p, lane);
  if (want_grad) {
    __syncwarp();
    R ab_c = 0, pb_c = 0;
    // the adjoint sweep streams tiles NB-1 .. 0; this warp's first tile is b = (NR-1) W + wt
    // (possibly past the end: the position is still advanced by W per round)
    // (seq0 >= NB >= W after the forward sweep, so the position below is never negative)
    pos.set((uint32_t)((int)seq0 + NB - 1 - ((NR - 1) * W + wt)), ring.nstage);
    for (int r = NR - 1; r >= 0; --r, pos.advance((uint32_t)W, ring.nstage)) {
      const int par = r & 1;
      const int b = r * W + wt;
      const bool act = b < NB;
      Blk<R> B;
      B.obs = 0u;
      const R* tile = nullptr;
      R m = 1, c = 0;
      if (act) {
        tile = ring_acquire(ring, pos, b);
        blk_residuals_c<R, PC>(B, tile, w_s, p, ld, lane);
        blk_gains_seq(B, lck[b * 64 + lane], s_e, s_h);
        blk_innov_seq(B, lck[b * 64 + 32 + lane]);
#pragma unroll
        for (int k = KS - 1; k >= 0; --k) {
          const R omk = (R)1 - B.K[k];
          c = fma(omk, c, B.v[k] * B.rF[k]);
          m = omk * m;
        }
        affine_scan_down(m, c, lane);
      }
      if (lane == 0) { ts->aggAB[par][wt][0] = m; ts->aggAB[par][wt][1] = c; }
      R me = __shfl_down_sync(FULL, m, 1), ce = __shfl_down_sync(FULL, c, 1);
      if (lane == 31) { me = 1; ce = 0; }
      team_sync(bar_id, nthreads);
      R ab_in, ab_tot;
      fold_down<R, WC>(ts->aggAB[par], ab_c, wt, W, ab_in, ab_tot);
      R abn[KS], q[KS], dF[KS];
      m = 1; c = 0;
      if (act) {
        R ab = fma(me, ab_in, ce);
#pragma unroll
        for (int k = KS - 1; k >= 0; --k) {
          abn[k] = ab;
          ab = fma((R)1 - B.K[k], ab, B.v[k] * B.rF[k]);
        }
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
      }
      if (lane == 0) { ts->aggPB[par][wt][0] = m; ts->aggPB[par][wt][1] = c; }
      me = __shfl_down_sync(FULL, m, 1); ce = __shfl_down_sync(FULL, c, 1);
      if (lane == 31) { me = 1; ce = 0; }
      team_sync(bar_id, nthreads);
      R pb_in, pb_tot;
      fold_down<R, WC>(ts->aggPB[par], pb_c, wt, W, pb_in, pb_tot);
      if (act) {
        R pb = fma(me, pb_in, ce);
        R lge = 0, lgh = 0, rbar[KS];
#pragma unroll
        for (int k = KS - 1; k >= 0; --k) {
          const R K = B.K[k], rF = B.rF[k], v = B.v[k];
          const R omk = (R)1 - K;
          lgh += pb;
          lge += fma(K * K, pb, dF[k]) - abn[k] * v * K * rF;
          rbar[k] = fma(K, abn[k], -v * rF);
          pb = fma(omk * omk, pb, q[k]);
        }
        ge_d += (double)lge; gh_d += (double)lgh;
        if (p > 0) {
          if (PC) {
            constexpr int LD = (PC + 1) | 1;
            const R* row0 = tile + tile_off(lane * KS, LD);
#pragma unroll
            for (int j = 0; j < NA; ++j)
#pragma unroll
              for (int k = 0; k < KS; ++k) accw[j] = fma(rbar[k], row0[k * LD + j], accw[j]);
          } else if (small_p) {
            blk_xt_rbar_small(tile, rbar, p, ld, lane, *reinterpret_cast<R(*)[PSMALL]>(&accw[0]));
          } else {
#pragma unroll
            for (int k = 0; k < KS; ++k) rbuf[lane * KS + k + (lane >> 2)] = rbar[k];
            __syncwarp();
            blk_xt_rbar<R, JS>(tile, rbuf, p, ld, xm.jj, xm.part, xm.nparts, accg);
            __syncwarp();
          }
        }
        ring_release(ring, pos, seq0 + (uint32_t)(NB - 1 - b), false, b, W, lane);
      }
      ab_c = ab_tot; pb_c = pb_tot;
    }
    seq0 += (uint32_t)NB;
  }

  // ======================= team totals (fixed order) =======================
  const double ll_w = warp_sum(acc_ll), ge_w = warp_sum(ge_d), gh_w = warp_sum(gh_d);
  const int n_w = __reduce_add_sync(FULL, n_obs);
  if (lane == 0) {
    ts->red[wt][0] = ll_w; ts->red[wt][1] = ge_w; ts->red[wt][2] = gh_w; ts->red[wt][3] = (double)n_w;
  }
  if (want_grad && p > 0) {
    if (PC) {
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        const R tot = warp_sum(accw[j]);
        if (lane == j) ts->gwpart[wt][j] = tot;
      }
    } else if (small_p) {
      static_assert(PSMALL == 16, "warp_multi_sum16");
      warp_multi_sum16(*reinterpret_cast<R(*)[PSMALL]>(&accw[0]), lane);
      if (!(lane & 1) && (lane >> 1) < p) ts->gwpart[wt][lane >> 1] = accw[0];
    } else {
#pragma unroll
      for (int s = 0; s < JS; ++s) {
        R a = accg[s];
        for (int o = xm.PJ; o < 32; o <<= 1) a += __shfl_xor_sync(FULL, a, o);
        const int j = lane + 32 * s;
        if (j < p) ts->gwpart[wt][j] = a;
      }
    }
  }
  team_sync(bar_id, nthreads);
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
  // red / gwpart are rewritten by the next evaluation only after several barriers
  team_sync(bar_id, nthreads);
}

// byte offset of the team areas = cfg.off_warp + n_warps * cfg.warp_bytes; per team:
// TeamStreamShared then the lane checkpoints [NB][64]
template <typename R> __host__ __device__ inline size_t tstream_team_bytes(int NB) {
  return ((sizeof(TeamStreamShared<R>) + 15) & ~(size_t)15) + (size_t)NB * 64 * sizeof(R);
}
template <typename R>
__device__ __forceinline__ TeamStreamShared<R>* tstream_area(unsigned char* smem, const SmemCfg& cfg,
                                                             int n_warps, int team, int NB, R** lck) {
  size_t off = (size_t)cfg.off_warp + (size_t)n_warps * cfg.warp_bytes;
  off = (off + 15) & ~(size_t)15;
  unsigned char* base = smem + off + (size_t)team * tstream_team_bytes<R>(NB);
  *lck = reinterpret_cast<R*>(base + ((sizeof(TeamStreamShared<R>) + 15) & ~(size_t)15));
  return reinterpret_cast<TeamStreamShared<R>*>(base);
}

template <typename R, int PC, int WC> struct TeamStreamEval {
  RingView<R> ring; uint32_t seq0;
  const ProbDev<R>& pr; TeamStreamShared<R>* ts; R* lck; const WarpScratch<R>& ws;
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
    team_stream_eval<R, PC, WC>(ring, seq0, pr, ts, lck, ws.w, ws.rbuf, s_e, s_h, true, lane, wt, W,
                                bar_id, ll, g_se, g_sh, gw);
    double g_u = g_se * (double)s_e, g_l = g_sh * (double)s_h;
    lp = ll + chain_prior(pr, omega, ws.w, u, l, s_e, s_h, lane, gw, g_u, g_l);
#pragma unroll
    for (int s = 0; s < DSLOTS; ++s) {
      const int i = lane + 32 * s;
      g[s] = i < p ? gw[s] : (i == p ? (R)g_u : (i == p + 1 ? (R)g_l : (R)0));
    }
  }
};

// CTA = GT teams x W warps, no producer warp: thread 0 starts the ring (the first `ahead` tiles
// and Omega), after that the warp of team 0 that finishes a tile issues the bulk copy of the
// tile `ahead` positions further down the stream.  A tile of the ring is consumed by ONE warp of
// every active team, so the `empty` barriers count teams.
template <typename R>
__device__ __forceinline__ bool tstream_prologue(unsigned char* smem, const SmemCfg& cfg,
                                                 const ProbDev<R>& pr, int W, int C,
                                                 long long n_sweeps, bool alternate, CtaShared<R>& cs,
                                                 RingView<R>& ring, int& team, int& wt, int& c) {
  const int warp = threadIdx.x >> 5;
  const int n_warps = blockDim.x >> 5;
  const int GT = n_warps / W;
  const int teams_active = min(GT, C - (int)blockIdx.x * GT);
  cs = cta_prologue(smem, cfg, pr, teams_active);
  team = warp / W; wt = warp - team * W;
  c = blockIdx.x * GT + team;
  ring.stage0 = cs.stage0; ring.full = cs.full; ring.empty = cs.empty; ring.gtiles = pr.tiles;
  ring.stage_elems = cfg.stage_elems; ring.nstage = cfg.nstage; ring.NB = pr.NB;
  ring.resident = cfg.resident != 0; ring.fetcher = team == 0; ring.alternate = alternate;
  ring.total = (uint32_t)(n_sweeps * (long long)pr.NB);
  ring.ahead = cfg.nstage - (uint32_t)W;      // (planner: nstage >= 2 W when streaming)
  if (threadIdx.x == 0) {
    omega_fetch(cs, pr);
    if (ring.resident) {
      tile_producer(pr.tiles, cs.stage0, cs.full, cs.empty, cfg.stage_elems, cfg.nstage, pr.NB, true,
                    0LL, [](long long) { return true; });
    } else {
      const uint32_t n0 = ring.ahead < ring.total ? ring.ahead : ring.total;
      for (uint32_t q = 0; q < n0; ++q) ring.fetch(q);
    }
  }
  return c < C;
}

template <typename R, int PC, int WC>
__global__ void __launch_bounds__(32 * TS_MAXWARPS, 1)
k_logpost_tstream(ProbDev<R> pr, SmemCfg cfg, int W, const R* __restrict__ theta, int C,
                  R* __restrict__ value, R* __restrict__ grad, int flags) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int n_warps = blockDim.x >> 5;
  const bool want_grad = grad != nullptr;
  CtaShared<R> cs; RingView<R> ring; int team, wt, c;
  if (!tstream_prologue(smem, cfg, pr, W, C, want_grad ? 2LL : 1LL, want_grad, cs, ring, team, wt, c))
    return;
  const int p = pr.p, dim = pr.dim;
  const R* th = theta + (size_t)c * dim;
  const WarpScratch<R> ws = warp_scratch<R>(smem, cfg, threadIdx.x >> 5);
  for (int j = lane; j < dim; j += 32) ws.w[j] = th[j];
  __syncwarp();
  const R u = ws.w[p], l = ws.w[p + 1];
  const R s_e = Num<R>::exp(u), s_h = Num<R>::exp(l);
  R* lck;
  TeamStreamShared<R>* ts = tstream_area<R>(smem, cfg, n_warps, team, pr.NB, &lck);
  uint32_t seq0 = 0;
  double ll, g_se, g_sh;
  R gw[JS];
  team_stream_eval<R, PC, WC>(ring, seq0, pr, ts, lck, ws.w, ws.rbuf, s_e, s_h, want_grad, lane, wt, W,
                              team + 1, ll, g_se, g_sh, gw);
  if (wt != 0) return;
  double val = ll;
  double g_u = g_se * (double)s_e, g_l = g_sh * (double)s_h;
  if (flags & 1) {
    omega_wait(cs);
    val += chain_prior(pr, cs.omega, ws.w, u, l, s_e, s_h, lane, gw, g_u, g_l);
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
}

template <typename R, int PC, int WC>
__global__ void __launch_bounds__(32 * TS_MAXWARPS, 1)
k_hmc_tstream(ProbDev<R> pr, SmemCfg cfg, int W, HmcPlan plan, uint64_t seed, uint64_t chain_id0,
              const R* __restrict__ theta0, int C, R* __restrict__ draws,
              ci_hmc_stats* __restrict__ stats) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int n_warps = blockDim.x >> 5;
  CtaShared<R> cs; RingView<R> ring; int team, wt, c;
  if (!tstream_prologue(smem, cfg, pr, W, C, 2LL * plan.n_evals, true, cs, ring, team, wt, c)) return;
  const WarpScratch<R> ws = warp_scratch<R>(smem, cfg, threadIdx.x >> 5);
  omega_wait(cs);
  R* lck;
  TeamStreamShared<R>* ts = tstream_area<R>(smem, cfg, n_warps, team, pr.NB, &lck);
  TeamStreamEval<R, PC, WC> ev{ring, 0u, pr, ts, lck, ws, cs.omega, lane, wt, W, team + 1};
  hmc_chain<R>(ev, plan, seed, chain_id0 + (uint64_t)c, theta0 + (size_t)c * pr.dim, pr.dim, lane,
               c, C, draws, stats);
}

}  // namespace ci
