// ci_seasonal.cuh -- SURVEY section 8 row f3: the reference's Gibbs sweep for models WITH
// seasonal components (ModelOptions.seasons), one warp per chain.
//
// Replaces gibbs_sampler.fit_with_gibbs_sampling as the reference calls it when
// `seasons` is non-empty (causalimpact/causalimpact_lib.py:365-388; the components are
// tfp.sts.Seasonal(allow_drift, constrain_mean_effect_to_zero) built at :471-489, initial
// drift scale 0.01 sd at :573-574).  Per sweep:
//   A. spike-and-slab regression + sigma_obs^2      (GibbsReg, ci_gibbs.cuh)
//   B. (level, seasonal effects) | rest, drawn JOINTLY with Durbin & Koopman's
//      mean-correction simulation smoother + Koopman's fast state smoother
//   C. sigma_level^2 and one drift variance per component ~ InverseGamma.
//
// State ("effects space", oracle/seasonal_np.py proves it equal in law to TFP's rotating
// constrained construction): x = (level, effects of component 0, effects of component 1, ..),
// dimension d <= 32, LANE i OWNS ELEMENT i.  Transition = identity; the observation sums the
// level and each component's ACTIVE season; when a season ends its effect receives
// drift * N(0,1) * (e_j - 1/n) (zero-sum direction); prior of a block: sd^2 (I - 11'/n).
//   pass A (forward):  simulate x+, y+ from the prior; Kalman filter on y* = r - y+ with
//                      the d x d covariance in shared memory (row i in lane i): rank-1
//                      downdates; gains K_t and e_t = v_t / F_t go to a per-chain scratch
//                      (shared memory when T (d+1) fits beside the tiles, else L2-resident)
//   pass B (backward): r_{t-1} = r_t + h_t (e_t - K_t' r_t); r_t overwrites K_t
//   pass C (forward):  xhat_0 = P_0 r_{-1}, xhat_{t+1} = xhat_t + Q_t r_t; x = x+ + xhat with
//                      x+ REGENERATED from the same Philox counters (nothing stored);
//                      emits level, per-component contributions, the trajectory, and the
//                      sufficient statistics of the next sweep's steps A and C.
// [X|y] is streamed twice per sweep through the same tile pipeline as every other kernel.
#pragma once
#include "ci_gibbs.cuh"

namespace ci {

enum : uint32_t { RNG_S_PATH = 9, RNG_S_INIT = 10, RNG_S_DRIFT = 11, RNG_S_GAMMA_U = 12 };
// Gamma(shape, 1) for any shape > 0 (boost for shape < 1: G(a) = G(a+1) U^(1/a)).
__device__ __forceinline__ double gamma_draw_any(double shape, uint64_t seed, uint32_t c0,
                                                 uint32_t c1, uint32_t it, uint32_t site) {
  if (shape >= 1.0) return gamma_draw(shape, seed, c0, c1, it, site);
  const double g = gamma_draw(shape + 1.0, seed, c0, c1, it, site);
  const uint4 x = Philox::gen(seed, c0, (c1 & ~0xffu) | RNG_S_GAMMA_U, it, site);
  return g * pow(u01<double>(x.x), 1.0 / shape);
}

__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// value of element `idx` (uniform or per-lane) of a vector distributed as element i <-> lane i % 32,
// slot i / 32
template <typename R, int E>
__device__ __forceinline__ R seas_pick(const R (&v)[E], int idx) {
  R r = __shfl_sync(FULL, v[0], idx & 31);
#pragma unroll
  for (int e = 1; e < E; ++e) {
    const R t = __shfl_sync(FULL, v[e], idx & 31);
    r = (idx >> 5) == e ? t : r;
  }
  return r;
}

// E = state elements per lane: the state x = (level, effects of every component) has d <= 32 E
// elements, element i lives in lane i % 32, slot i / 32 (E = 1: day-of-week and the like; E = 2:
// week-of-year, d = 53; E = 6: hour-of-week, d = 169).
template <typename R, int E>
__global__ void __launch_bounds__(32 * (MAXG + 1), 1)
k_gibbs_seasonal(ProbDev<R> pr, GibbsDev<R> gd, SeasDev sz, SmemCfg cfg, GibbsPlan plan,
                 uint64_t seed, uint64_t chain_id0, int C, R* __restrict__ draws,
                 R* __restrict__ level_out, R* __restrict__ traj_out, R* __restrict__ latent_out,
                 R* __restrict__ seas_out, R* __restrict__ drift_out,
                 float* __restrict__ incl_out, const BatchDev<R>* __restrict__ batch) {
  extern __shared__ __align__(128) unsigned char smem[];
  // batched launch (grid.y = series; the season calendar is shared by the panel): as k_gibbs
  const size_t series_row0 = batch ? (size_t)blockIdx.y * C * plan.n_results : 0;
  const size_t series_chain0 = batch ? (size_t)blockIdx.y * C : 0;
  if (batch) {
    pr = batch[blockIdx.y].pr; gd = batch[blockIdx.y].gd; plan.n_obs = batch[blockIdx.y].n_obs;
    if (sz.per_series) {        // the priors scale with each series' own outcome sd (lib.py:472-489)
      const double* ps = sz.per_series + 3 * (size_t)blockIdx.y;
      sz.init_var = ps[0]; sz.drift_scale = ps[1]; sz.drift_ub = ps[2];
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (blockDim.x >> 5) - 1;
  const int chain0 = blockIdx.x * G;
  const int nactive = min(G, C - chain0);
  const CtaShared<R> cs = cta_prologue(smem, cfg, pr, nactive);
  const int p = pr.p, dim = pr.dim, ld = pr.ld, NB = pr.NB, T = pr.T;
  R* gram_s = reinterpret_cast<R*>(smem + (((size_t)cfg.off_warp + (size_t)G * cfg.warp_bytes + 15) & ~(size_t)15));
  for (int i = threadIdx.x; i < p * p; i += blockDim.x) gram_s[i] = gd.gram[i];
  __syncthreads();
  const int n_iter = plan.n_warmup + plan.n_results;
  if (warp == G) {
    if (lane == 0) {
      omega_fetch(cs, pr);
      tile_producer(pr.tiles, cs.stage0, cs.full, cs.empty, cfg.stage_elems, cfg.nstage, NB,
                    cfg.resident != 0, 2LL * n_iter, [](long long) { return true; });
    }
    return;
  }
  if (warp >= nactive) return;
  omega_wait(cs);
  const R* om_s = cs.omega;
  const int c = chain0 + warp;
  const WarpScratch<R> ws = warp_scratch<R>(smem, cfg, warp);
  GibbsScratch<R> gs;
  gs.bvec = ws.extra; gs.La = gs.bvec + p; gs.Lo = gs.La + (p + 1) * (p + 1);
  gs.idx = gs.Lo + p * p; gs.vec = gs.idx + p; gs.perm = gs.vec + (p + 1);
  const int d = sz.d, K = sz.K, LDP = d | 1;
  R* Ps = gs.perm + p;                  // [d][LDP] state covariance, row i <-> element i
  R* phs = Ps + d * LDP;                // [d]      P h of the current step
  R* st_a = phs + d;                    // [TB] x3: per-step staging of the current tile
  R* st_b = st_a + TB;
  R* st_c = st_b + TB;
  R* st_d = st_c + TB;                  // [K][TB] drift normals of the tile's steps
  // per-step scratch (gains / r_t): in shared memory when the launch found room, else L2
  const bool scr_smem = sz.scratch == nullptr;
  R* scr = scr_smem ? st_d + (size_t)K * TB : static_cast<R*>(sz.scratch) + (series_chain0 + (size_t)c) * T * (d + 1);
  TilePipe<R> pipe = make_pipe(cs, cfg);
  const uint64_t gid = chain_id0 + (uint64_t)c + (batch ? (uint64_t)blockIdx.y * plan.series_stride : 0ull);
  const uint32_t id_lo = (uint32_t)gid, id_hi8 = (uint32_t)(gid >> 32) << 8;

  // which block each of the lane's elements belongs to: -1 level, k component, -2 unused slot
  int comp[E], my_n[E], my_off[E], el[E];
  R inv_n[E];
  R* Prow[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    el[e] = lane + 32 * e;
    comp[e] = el[e] == 0 ? -1 : -2; my_n[e] = 1; my_off[e] = 0;
    for (int k = 0; k < K; ++k)
      if (el[e] >= sz.off[k] && el[e] < sz.off[k] + sz.n[k]) { comp[e] = k; my_n[e] = sz.n[k]; my_off[e] = sz.off[k]; }
    inv_n[e] = (R)1 / (R)my_n[e];
    Prow[e] = Ps + (el[e] < d ? el[e] : 0) * LDP;
  }
  const int my_coff = (lane >= 1 && lane <= K) ? sz.off[lane - 1] : 0;   // lane l in 1..K <-> component l-1

  // ---- initial state: the reference's (lib.py:566-581) ----
  double s_e = p > 0 ? 0.2 * (double)pr.P0 : (double)pr.P0;
  double s_h = (double)pr.lvl_scale / (double)pr.lvl_conc;
  double sd_l = 1e-4 * sz.init_var;                  // drift VARIANCE of component `lane` (0.01 sd)^2
  uint32_t gam[4] = {0u, 0u, 0u, 0u};
  if (!plan.sparse)
    for (int j = 0; j < p; ++j) gam[j >> 5] |= 1u << (j & 31);
  for (int j = lane; j < p; j += 32) { gs.bvec[j] = gd.xty0[j]; ws.w[j] = 0; }
  double yty = (double)gd.yty0;
  double incl_cnt[DSLOTS] = {0.0, 0.0, 0.0, 0.0};
  const double conc_e = (double)pr.obs_conc + 0.5 * plan.n_obs;
  __syncwarp();
  const GibbsReg<R> reg{pr, gs, gram_s, om_s, p, lane, conc_e, seed, id_lo, id_hi8};
  const R sd0 = Num<R>::sqrt(pr.P0), sdi = (R)sqrt(sz.init_var);

  // v minus the mean of v over its component's block (0 outside blocks)
  auto block_centered = [&](const R (&v)[E], R (&out)[E]) {
#pragma unroll
    for (int e = 0; e < E; ++e) out[e] = 0;
    for (int k = 0; k < K; ++k) {
      R part = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) part += comp[e] == k ? v[e] : (R)0;
      const R s = warp_sum(part);
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (comp[e] == k) out[e] = v[e] - s * inv_n[e];
    }
  };
  // the step's schedule: cols[k] (uniform) = state index of the k-th observed column
  // (k = 0: the level, k >= 1: component k-1's active season), per-slot masks of those columns,
  // ends flags
  auto step_sched = [&](int t, int (&cols)[MAX_SEAS + 1], unsigned (&cmask)[E], int& em) -> int {
    const uint8_t* sc = sz.sched + (size_t)t * (K + 1);
    em = sc[K];
    const int src = (lane >= 1 && lane <= K) ? my_coff + (int)sc[lane - 1] : 0;
#pragma unroll
    for (int e = 0; e < E; ++e) cmask[e] = 0u;
#pragma unroll
    for (int k = 0; k <= MAX_SEAS; ++k) {
      cols[k] = 0;
      if (k <= K) {                                   // K is uniform: no divergence
        cols[k] = __shfl_sync(FULL, src, k);
#pragma unroll
        for (int e = 0; e < E; ++e)
          if ((cols[k] >> 5) == e) cmask[e] |= 1u << (cols[k] & 31);
      }
    }
    return src;      // lane l in 1..K: state index of component l-1's active season
  };
  auto gather = [&](const R (&v)[E], const int (&cols)[MAX_SEAS + 1]) -> R {      // h' v
    R acc = 0;
#pragma unroll
    for (int k = 0; k <= MAX_SEAS; ++k)
      if (k <= K) acc += seas_pick<R, E>(v, cols[k]);
    return acc;
  };
  // All normals of step t: level noise, observation noise of y+, predictive noise, and one
  // drift normal per component.  Generated LANE-PARALLEL (a lane does its 8 steps of the
  // tile) and staged in shared memory: run 28 showed the per-step, warp-redundant
  // Philox + Box-Muller of the first version was 35 % of all instructions.
  auto step_normals = [&](int t, int it, R& ze, R& zo, R& zpred, R (&dr)[MAX_SEAS]) {
    const uint4 x = Philox::gen(seed, id_lo, RNG_S_PATH | id_hi8, (uint32_t)t, (uint32_t)it);
    box_muller<R>(x.x, x.y, ze, zo);
    box_muller<R>(x.z, x.w, zpred, dr[0]);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      if (K > 1 + 4 * m) {
        const uint4 y = Philox::gen(seed, id_lo, RNG_S_DRIFT | id_hi8, (uint32_t)t,
                                    (uint32_t)it * 2u + (uint32_t)m);
        R q0, q1, q2, q3;
        box_muller<R>(y.x, y.y, q0, q1);
        box_muller<R>(y.z, y.w, q2, q3);
        dr[1 + 4 * m] = q0; dr[2 + 4 * m] = q1;
        if (m == 0) { dr[3] = q2; dr[4] = q3; }
      }
    }
  };
  auto init_xplus = [&](int it, R (&out)[E]) {
    R z[E], zc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const uint4 x = Philox::gen(seed, id_lo, RNG_S_INIT | id_hi8, (uint32_t)el[e], (uint32_t)it);
      R z1;
      box_muller<R>(x.x, x.y, z[e], z1);
    }
    block_centered(z, zc);
#pragma unroll
    for (int e = 0; e < E; ++e)
      out[e] = el[e] == 0 ? fma(sd0, z[e], pr.m0) : (comp[e] >= 0 ? sdi * zc[e] : (R)0);
  };

  for (int it = 0; it < n_iter; ++it) {
    // =================== A. regression block ===================
    reg.step(it, plan, gam, yty, s_e, ws.w);

    // =================== B. (level, seasonal) | rest ===================
    const R se = (R)s_e, sh = (R)s_h, sig_e = (R)sqrt(s_e), sig_h = (R)sqrt(s_h);
    const R sdv = (R)sd_l;
    R sig_d[MAX_SEAS], inv_nk[MAX_SEAS];                    // uniform: drift scale, 1/n per component
#pragma unroll
    for (int k = 0; k < MAX_SEAS; ++k) {
      sig_d[k] = k < K ? Num<R>::sqrt(__shfl_sync(FULL, sdv, k)) : (R)0;
      inv_nk[k] = k < K ? (R)1 / (R)sz.n[k] : (R)0;
    }
    // ---- pass A ----
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (el[e] < d) {
        for (int j = 0; j < d; ++j) Prow[e][j] = 0;
        if (el[e] == 0) Prow[e][0] = pr.P0;
        if (comp[e] >= 0)
          for (int j = 0; j < my_n[e]; ++j)
            Prow[e][my_off[e] + j] = (R)sz.init_var * ((my_off[e] + j == el[e] ? (R)1 : (R)0) - inv_n[e]);
      }
    }
    __syncwarp();
    R xp[E], a[E];
    init_xplus(it, xp);
#pragma unroll
    for (int e = 0; e < E; ++e) a[e] = 0;
    for (int b = 0; b < NB; ++b) {
      const R* tile = pipe.acquire(b);
      Blk<R> B;
      R xw[KS];
      blk_residuals_xw(B, xw, tile, ws.w, p, ld, lane);
      const int t0 = b * TB + lane * KS;
      // stage the tile's per-step inputs (residual or NaN, level noise, observation noise) so
      // that the sequential loop below stays ROLLED (an 8x unrolled body overflowed the
      // instruction cache: 2200 cycles per step, run 22)
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        R ze, zo, zpred, dr[MAX_SEAS];
        step_normals(t0 + kk, it, ze, zo, zpred, dr);
        st_a[lane * KS + kk] = ((B.obs >> kk) & 1u) ? B.r[kk] : Num<R>::nan();
        st_b[lane * KS + kk] = ze;
        st_c[lane * KS + kk] = zo;
#pragma unroll
        for (int k = 0; k < MAX_SEAS; ++k)
          if (k < K) st_d[k * TB + lane * KS + kk] = dr[k];
      }
      __syncwarp();
      const int nstep = min(TB, T - b * TB);
#pragma unroll 1
      for (int tl = 0; tl < nstep; ++tl) {
        const int t = b * TB + tl;
        const R r_t = st_a[tl], eta = st_b[tl], eps = st_c[tl];
        int cols[MAX_SEAS + 1], em; unsigned cmask[E];
        step_sched(t, cols, cmask, em);
        R Kg[E], ee = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) Kg[e] = 0;
        if (r_t == r_t) {                                       // observed step
          R xa[E];
#pragma unroll
          for (int e = 0; e < E; ++e) xa[e] = xp[e] + a[e];
          const R hxa = gather(xa, cols);
          R Ph[E];
#pragma unroll
          for (int e = 0; e < E; ++e) {
            Ph[e] = 0;
            if (el[e] < d) {
#pragma unroll
              for (int k = 0; k <= MAX_SEAS; ++k)
                if (k <= K) Ph[e] += Prow[e][cols[k]];
            }
          }
          const R rF = Num<R>::rcp(gather(Ph, cols) + se);
          const R v = (r_t - sig_e * eps) - hxa;
          ee = v * rF;
#pragma unroll
          for (int e = 0; e < E; ++e) { Kg[e] = Ph[e] * rF; a[e] = fma(Kg[e], v, a[e]); }
          // P -= (P h)(P h)' / F: row i in the lane owning element i; (P h)_j is broadcast
          // by shuffle for one slot per lane (symmetric in fp: the product Ph_i Ph_j commutes),
          // through shared memory for wider states
          if (E == 1) {
#pragma unroll 4
            for (int j = 0; j < d; ++j) {
              const R pj = __shfl_sync(FULL, Ph[0], j);
              if (lane < d) Prow[0][j] = fma(-(Ph[0] * pj), rF, Prow[0][j]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < E; ++e)
              if (el[e] < d) phs[el[e]] = Ph[e];
            __syncwarp();
#pragma unroll 2
            for (int j = 0; j < d; ++j) {
              const R pj = phs[j];
#pragma unroll
              for (int e = 0; e < E; ++e)
                if (el[e] < d) Prow[e][j] = fma(-(Ph[e] * pj), rF, Prow[e][j]);
            }
          }
          __syncwarp();
        }
        R* srow = scr + (size_t)t * (d + 1);
#pragma unroll
        for (int e = 0; e < E; ++e)
          if (el[e] < d) srow[el[e]] = Kg[e];
        if (lane == 0) srow[d] = ee;
        // x_{t+1} = x_t + noise_t
        if (lane == 0) { Prow[0][0] += sh; xp[0] = fma(sig_h, eta, xp[0]); }
        if (em) {
#pragma unroll
          for (int k = 0; k < MAX_SEAS; ++k) {
            if (k >= K || !((em >> k) & 1)) continue;
            const int jk = cols[k + 1];
            const R sdk = __shfl_sync(FULL, sdv, k);
#pragma unroll
            for (int e = 0; e < E; ++e) {
              const R ci = comp[e] == k ? ((el[e] == jk ? (R)1 : (R)0) - inv_n[e]) : (R)0;
              if (comp[e] == k) {
                // P_block += sdk c c',  c = e_jk - 1/n:  row i gets  sdk c_i (e_jk - 1/n)'
                const R f = sdk * ci, g0 = -f * inv_n[e];
                for (int j = 0; j < my_n[e]; ++j) Prow[e][my_off[e] + j] += g0;
                Prow[e][jk] += f;
              }
              xp[e] = fma(sig_d[k] * st_d[k * TB + tl], ci, xp[e]);
            }
          }
          __syncwarp();
        }
      }
      __syncwarp();
      pipe.release(lane);
    }
    __syncwarp();
    // ---- pass B ----
    R rr[E];
#pragma unroll
    for (int e = 0; e < E; ++e) rr[e] = 0;
#pragma unroll 1
    for (int t = T - 1; t >= 0; --t) {
      int cols[MAX_SEAS + 1], em; unsigned cmask[E];
      step_sched(t, cols, cmask, em);
      R* srow = scr + (size_t)t * (d + 1);
      if (!scr_smem && t >= 4 && lane == 0) prefetch_l1(srow - 4 * (d + 1));
      const R ee = srow[d];
      R part = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const R Kg = el[e] < d ? srow[el[e]] : (R)0;
        if (el[e] < d) srow[el[e]] = rr[e];                // r_t, read back by pass C
        part = fma(Kg, rr[e], part);
      }
      const R dot = warp_sum(part);
#pragma unroll
      for (int e = 0; e < E; ++e)
        if ((cmask[e] >> lane) & 1u) rr[e] += ee - dot;
    }
    __syncwarp();
    // ---- pass C ----
    const bool keep = it >= plan.n_warmup;
    const size_t out_row = !keep ? 0
        : series_row0 + (plan.chain_major ? ((size_t)c * plan.n_results + (size_t)(it - plan.n_warmup))
                                          : ((size_t)(it - plan.n_warmup) * C + c));
    R xt[E];                                               // x = x+ + xhat
    init_xplus(it, xt);
    {
      R rc[E];
      block_centered(rr, rc);
#pragma unroll
      for (int e = 0; e < E; ++e)
        xt[e] += el[e] == 0 ? pr.P0 * rr[e] : (comp[e] >= 0 ? (R)sz.init_var * rc[e] : (R)0);
    }
    const XtMap xm = xt_map(p, lane);
    const bool small_p = p <= PSMALL;
    R accw[PSMALL];
#pragma unroll
    for (int j = 0; j < PSMALL; ++j) accw[j] = 0;
    R accg[JS];
#pragma unroll
    for (int s = 0; s < JS; ++s) accg[s] = 0;
    double n_yty = 0.0, d2 = 0.0, su2 = 0.0;
    R prev_level = 0;
    for (int b = 0; b < NB; ++b) {
      const R* tile = pipe.acquire(b);
      Blk<R> B;
      R xw[KS];
      blk_residuals_xw(B, xw, tile, ws.w, p, ld, lane);
      const int t0 = b * TB + lane * KS;
      R zp[KS];
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        R ze, zo, dr[MAX_SEAS];
        step_normals(t0 + kk, it, ze, zo, zp[kk], dr);      // the SAME normals as pass A
        st_a[lane * KS + kk] = ze;                          // level noise of the step
#pragma unroll
        for (int k = 0; k < MAX_SEAS; ++k)
          if (k < K) st_d[k * TB + lane * KS + kk] = dr[k];
      }
      __syncwarp();
      const int nstep = min(TB, T - b * TB);
#pragma unroll 1
      for (int tl = 0; tl < nstep; ++tl) {
        const int t = b * TB + tl;
        const R eta = st_a[tl];
        int cols[MAX_SEAS + 1], em; unsigned cmask[E];
        const int src = step_sched(t, cols, cmask, em);
        const R* srow = scr + (size_t)t * (d + 1);
        if (!scr_smem && t + 4 < T && lane == 0) prefetch_l1(srow + 4 * (d + 1));
        const R level_t = __shfl_sync(FULL, xt[0], 0);
        const R tot = gather(xt, cols);                     // level + all contributions
        const R mine = seas_pick<R, E>(xt, src);            // lane l in 1..K: component l-1's share
        if (lane == 0) { st_b[tl] = level_t; st_c[tl] = tot - level_t; }
        if (keep && seas_out && lane >= 1 && lane <= K)
          seas_out[(out_row * T + t) * K + (lane - 1)] = mine;
        if (t > 0) { const double dl = (double)(level_t - prev_level); d2 += dl * dl; }
        prev_level = level_t;
        if (t < T - 1) {
          R rt[E];
#pragma unroll
          for (int e = 0; e < E; ++e) rt[e] = el[e] < d ? srow[el[e]] : (R)0;
          if (lane == 0) xt[0] += fma(sh, rt[0], sig_h * eta);
          if (em) {
#pragma unroll
            for (int k = 0; k < MAX_SEAS; ++k) {
              if (k >= K || !((em >> k) & 1)) continue;
              const int jk = cols[k + 1];
              const R sdk = __shfl_sync(FULL, sdv, k);
              R part = 0;
#pragma unroll
              for (int e = 0; e < E; ++e) part += comp[e] == k ? rt[e] : (R)0;
              const R cr = seas_pick<R, E>(rt, jk) - warp_sum(part) * inv_nk[k];
              const R u = fma(sdk, cr, sig_d[k] * st_d[k * TB + tl]);
#pragma unroll
              for (int e = 0; e < E; ++e) {
                const R ci = comp[e] == k ? ((el[e] == jk ? (R)1 : (R)0) - inv_n[e]) : (R)0;
                xt[e] = fma(u, ci, xt[e]);
              }
              if (lane == k) su2 += (double)u * (double)u;
            }
          }
        }
      }
      __syncwarp();
      R lv[KS], sc[KS];
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) { lv[kk] = st_b[lane * KS + kk]; sc[kk] = st_c[lane * KS + kk]; }
      __syncwarp();
      // ---- per-lane epilogue for the tile's 8 owned steps (as k_gibbs) ----
      R tgt[KS];
      R ly = 0;
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        const bool o = (B.obs >> kk) & 1u;
        // targets of the next step A: y - level - seasonal on observed steps (B.r = y - x.w)
        tgt[kk] = o ? (B.r[kk] + xw[kk]) - lv[kk] - sc[kk] : (R)0;
        ly = fma(tgt[kk], tgt[kk], ly);
      }
      n_yty += (double)ly;
      if (p > 0) {
        if (small_p) {
          blk_xt_rbar_small(tile, tgt, p, ld, lane, accw);
        } else {
#pragma unroll
          for (int kk = 0; kk < KS; ++kk) ws.rbuf[lane * KS + kk + (lane >> 2)] = tgt[kk];
          __syncwarp();
          blk_xt_rbar<R, JS>(tile, ws.rbuf, p, ld, xm.jj, xm.part, xm.nparts, accg);
          __syncwarp();
        }
      }
      if (keep) {
        R la[KS], tr[KS];
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
          la[kk] = lv[kk] + sc[kk];
          tr[kk] = la[kk] + xw[kk] + sig_e * zp[kk];
        }
        if (level_out) store_run(level_out + out_row * T, t0, T, lv);
        if (latent_out) store_run(latent_out + out_row * T, t0, T, la);
        if (traj_out) store_run(traj_out + out_row * T, t0, T, tr);
      }
      pipe.release(lane);
    }
    // publish the statistics of the next sweep
    yty = warp_sum(n_yty);
    __syncwarp();
    if (small_p) {
#pragma unroll
      for (int j = 0; j < PSMALL; ++j) {
        if (j < p) {
          const R tot = warp_sum(accw[j]);
          if (lane == 0) gs.bvec[j] = tot;
        }
      }
    } else {
#pragma unroll
      for (int s = 0; s < JS; ++s) {
        R av = accg[s];
        for (int o = xm.PJ; o < 32; o <<= 1) av += __shfl_xor_sync(FULL, av, o);
        const int j = lane + 32 * s;
        if (j < p && lane < xm.PJ) gs.bvec[j] = av;
      }
    }
    __syncwarp();
    // =================== C. variances ===================
    {
      const double g = gamma_draw((double)pr.lvl_conc + 0.5 * (T - 1), seed, id_lo,
                                  RNG_G_GAMMA | id_hi8, (uint32_t)it, 1u);
      s_h = ((double)pr.lvl_scale + 0.5 * d2) / g;
      const double ub2 = (double)pr.lvl_ub;                            // variance bound
      if (s_h > ub2) s_h = ub2;                                        // lib.py:432
    }
    for (int k = 0; k < K; ++k) {
      const double u2 = __shfl_sync(FULL, su2, k);
      const double g = gamma_draw_any(sz.drift_conc + 0.5 * sz.n_ends[k], seed, id_lo,
                                      RNG_G_GAMMA | id_hi8, (uint32_t)it, 2u + (uint32_t)k);
      double v = (sz.drift_scale + 0.5 * u2) / g;
      const double ub2 = sz.drift_ub;                                  // variance bound
      if (v > ub2) v = ub2;                                            // lib.py:474
      if (lane == k) sd_l = v;
    }
    if (keep) {
      R* row = draws + out_row * dim;
      for (int j = lane; j < p; j += 32) row[j] = ws.w[j];
      if (lane == 0) { row[p] = (R)log(s_e); row[p + 1] = (R)log(s_h); }
      if (drift_out && lane < K) drift_out[out_row * K + lane] = (R)log(sd_l);
#pragma unroll
      for (int wd = 0; wd < DSLOTS; ++wd) incl_cnt[wd] += (double)((gam[wd] >> lane) & 1u);
    }
  }
  if (incl_out) {
#pragma unroll
    for (int wd = 0; wd < DSLOTS; ++wd) {
      const int j = lane + 32 * wd;
      if (j < p) incl_out[(series_chain0 + (size_t)c) * p + j] = (float)(incl_cnt[wd] / (plan.n_results > 0 ? plan.n_results : 1));
    }
  }
}

}  // namespace ci
