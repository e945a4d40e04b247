// ci_gibbs.cuh -- the reference's OWN sampler on the GPU: a persistent Gibbs
// kernel with the spike-and-slab regression step.
//
// Replaces gibbs_sampler.fit_with_gibbs_sampling as the reference calls it
// (causalimpact/causalimpact_lib.py:365-388, model :398-500, initial state
// :566-581): per sweep
//   A. targets = y - level on observed steps; for every feature, draw its
//      inclusion indicator from its conditional (stochastic search variable
//      selection with Scott & Varian's marginal, inclusion prior min(1, 3/p),
//      lib.py:449-450); sigma_obs^2 ~ InvGamma; active weights ~ Normal.
//   B. level ~ p(level | y - Xw, sigma's): forward filter + backward sampling
//      (the same reverse affine scan as K4), which also accumulates the
//      sufficient statistics of the next sweep (X'(y - level), |y - level|^2,
//      sum (d level)^2) so [X|y] is streamed exactly twice per sweep.
//   C. sigma_level^2 ~ InvGamma.
// One warp per chain, many chains per launch; the dense algebra of step A (a
// Cholesky of the (k+1)x(k+1) augmented active Gram matrix per candidate) runs
// warp-cooperatively in shared memory.  TFP is not available here, so feature
// visiting order (0..p-1) and the absence of TFP's experimental weight
// adjustment are documented deviations; oracle/gibbs_np.py restates the sweep
// and tests/test_gpu_gibbs.py compares the two statistically.
#pragma once
#include "../../include/ci_b200.h"
#include "ci_kernels.cuh"

namespace ci {

enum : uint32_t { RNG_G_INCL = 6, RNG_G_GAMMA = 7, RNG_G_W = 8, RNG_G_PERM = 13 };

struct GibbsPlan {
  int n_warmup, n_results, sparse, n_obs, chain_major;
  int ssvs_random;                     // visit the features in a fresh random order every sweep
  unsigned long long series_stride;    // batch: chain ids of series s start at chain_id0 + s * stride
  double logit_pi;
};

template <typename R> struct GibbsDev {
  const R* gram;   // [p,p]  X'X over observed rows
  const R* xty0;   // [p]    X'y over observed rows (targets of sweep 0: level = 0)
  R yty0;          //        y'y over observed rows
};

// One entry per series of a batch (ci_set_data_batch): the kernel of CTA
// row blockIdx.y works on series blockIdx.y -- its own tiles, priors and sufficient statistics.
template <typename R> struct BatchDev {
  ProbDev<R> pr;
  GibbsDev<R> gd;
  int n_obs;
};

// per-warp dense-algebra scratch (elements of R)
template <typename R> struct GibbsScratch {
  R* bvec;   // [p]        X'(y - level)
  R* La;     // [(p+1)^2]  Cholesky of the augmented active matrix
  R* Lo;     // [p^2]      Cholesky of Omega_gamma
  R* idx;    // [p]        active feature list (stored as R, exact for p <= 128)
  R* vec;    // [p+1]      work vector
  R* perm;   // [p]        visiting order of the sweep (ssvs_random)
};

// Gamma(shape, 1), shape >= 1 (Marsaglia & Tsang 2000); every lane computes the same draw.
__device__ __forceinline__ double gamma_draw(double shape, uint64_t seed, uint32_t c0, uint32_t c1,
                                             uint32_t it, uint32_t site) {
  const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (uint32_t attempt = 0; attempt < 64; ++attempt) {
    const uint4 x = Philox::gen(seed, c0, c1, it, site * 64u + attempt);
    double z0, z1;
    box_muller<double>(x.x, x.y, z0, z1);
    const double v1 = 1.0 + c * z0;
    if (v1 <= 0.0) continue;
    const double v = v1 * v1 * v1;
    const double u = u01<double>(x.z);
    if (log(u) < 0.5 * z0 * z0 + d - d * v + d * log(v)) return d * v;
  }
  return d;   // (probability ~ 1e-80) mode fallback
}

// Warp-cooperative Cholesky of the n x n matrix A(i,j) given by `elem`, into L
// (row-major, stride n).  Returns sum_{a < n_log} log L[a][a]; pivots are clamped.
template <typename R, typename Elem>
__device__ __forceinline__ double warp_cholesky(R* L, int n, int n_log, int lane, Elem elem) {
  double logdiag = 0.0;
  for (int c = 0; c < n; ++c) {
    R s = elem(c, c);
    if (c <= 32) {
      // small systems (the usual case: a handful of active features): every lane forms the
      // pivot redundantly from broadcast shared-memory reads -- no shuffle reduction
      for (int m = 0; m < c; ++m) s = fma(-L[c * n + m], L[c * n + m], s);
    } else {
      R part = 0;
      for (int m = lane; m < c; m += 32) part = fma(L[c * n + m], L[c * n + m], part);
      s -= warp_sum(part);
    }
    const R dd = Num<R>::sqrt(s > (R)1e-30 ? s : (R)1e-30);
    const R rd = (R)1 / dd;
    if (c < n_log) logdiag += (double)Num<R>::log(dd);
    __syncwarp();
    if (lane == 0) L[c * n + c] = dd;
    for (int i = c + 1 + lane; i < n; i += 32) {
      R t = elem(i, c);
      for (int m = 0; m < c; ++m) t = fma(-L[i * n + m], L[c * n + m], t);
      L[i * n + c] = t * rd;
    }
    __syncwarp();
  }
  return logdiag;
}

// ---------------------------------------------------------------------------
// Step A of the sweep (spike-and-slab regression + sigma_obs^2), shared by k_gibbs and the
// seasonal kernel (ci_seasonal.cuh).  Warp-cooperative; every lane holds the same scalars.
// ---------------------------------------------------------------------------
template <typename R> struct GibbsReg {
  const ProbDev<R>& pr;
  GibbsScratch<R> gs;
  const R* gram_s;
  const R* om_s;
  int p, lane;
  double conc_e;
  uint64_t seed;
  uint32_t id_lo, id_hi8;

  // number of active features and their list -> gs.idx; returns k
  __device__ __forceinline__ int build_idx(const uint32_t (&g)[4]) const {
    int k = 0;
#pragma unroll
    for (int wd = 0; wd < 4; ++wd) {
      const int j = lane + 32 * wd;
      const bool on = j < p && ((g[wd] >> lane) & 1u);
      const int pos = k + __popc(g[wd] & ((1u << lane) - 1u));
      if (on) gs.idx[pos] = (R)j;
      k += __popc(g[wd]);
    }
    __syncwarp();
    return k;
  }
  // log marginal of configuration g (Scott & Varian 2014, b = 0); leaves the
  // augmented factor in gs.La (stride k+1) and returns SS through `ss`
  __device__ __forceinline__ double log_marginal(const uint32_t (&g)[4], double yty, int& k_out,
                                                 double& ss) const {
    const int k = build_idx(g);
    k_out = k;
    const int n = k + 1;
    const R ytyR = (R)yty;
    const GibbsScratch<R>& s_ = gs;
    const R* gram = gram_s; const R* om = om_s; const int pp = p;
    const double ld_lam = warp_cholesky<R>(gs.La, n, k, lane, [&](int i, int j) -> R {
      if (i == k) return j == k ? ytyR : s_.bvec[(int)s_.idx[j]];
      const int a = (int)s_.idx[i], b = (int)s_.idx[j];
      return gram[a * pp + b] + om[a * pp + b];
    });
    const double ld_om = warp_cholesky<R>(gs.Lo, k, k, lane, [&](int i, int j) -> R {
      return om[(int)s_.idx[i] * pp + (int)s_.idx[j]];
    });
    const double piv = (double)gs.La[k * n + k];
    ss = piv * piv;
    return ld_om - ld_lam - conc_e * log((double)pr.obs_scale + 0.5 * ss);
  }

  // visiting order of the sweep's inclusion draws -> gs.perm (ssvs_random): one Fisher-Yates
  // shuffle per sweep, keyed like every other draw of the chain
  __device__ __forceinline__ void make_order(int it, const GibbsPlan& plan) const {
    if (!plan.ssvs_random) return;
    for (int j = lane; j < p; j += 32) gs.perm[j] = (R)j;
    __syncwarp();
    if (lane == 0) {
      for (int i = p - 1; i > 0; --i) {
        const uint4 x = Philox::gen(seed, id_lo, RNG_G_PERM | id_hi8, (uint32_t)it, (uint32_t)i);
        const int k2 = (int)(((uint64_t)x.x * (uint64_t)(i + 1)) >> 32);
        const R t = gs.perm[i]; gs.perm[i] = gs.perm[k2]; gs.perm[k2] = t;
      }
    }
    __syncwarp();
  }
  __device__ __forceinline__ int order(int jj, const GibbsPlan& plan) const {
    return plan.ssvs_random ? (int)gs.perm[jj] : jj;
  }
  // the inclusion draw of feature j given the log-marginals of the current configuration and of
  // the one with j flipped; returns true when the indicator changes
  __device__ __forceinline__ bool flips(int it, int j, bool cur, double lm_cur, double lm_new,
                                        const GibbsPlan& plan) const {
    const double d = (cur ? lm_cur - lm_new : lm_new - lm_cur) + plan.logit_pi;
    const uint4 x = Philox::gen(seed, id_lo, RNG_G_INCL | id_hi8, (uint32_t)it, (uint32_t)j);
    const bool take = u01<double>(x.x) < 1.0 / (1.0 + exp(-d));
    return take != cur;
  }

  // sigma_obs^2 | gam  and the active weights | sigma_obs^2, gam  -> s_e, w_s.  Needs the factor
  // of the final configuration in gs.La (log_marginal(gam, ...) was the last factorisation):
  // k active features, residual sum of squares ss.
  __device__ __forceinline__ void draw(int it, int k, double ss, double& s_e, R* w_s) const {
    {
      const double g = gamma_draw(conc_e, seed, id_lo, RNG_G_GAMMA | id_hi8, (uint32_t)it, 0u);
      s_e = ((double)pr.obs_scale + 0.5 * ss) / g;
      const double ub2 = (double)pr.obs_ub;                            // variance bound
      if (s_e > ub2) s_e = ub2;                                        // lib.py:442-443
    }
    if (p > 0) {
      // w_active = L^-T (z + sqrt(s_e) eps),  z = last row of the augmented factor
      const int n = k + 1;
      const R sig = (R)sqrt(s_e);
      for (int a = lane; a < k; a += 32) {
        const uint4 x = Philox::gen(seed, id_lo, RNG_G_W | id_hi8, (uint32_t)it, (uint32_t)(a >> 2));
        R z0, z1, z2, z3;
        box_muller<R>(x.x, x.y, z0, z1);
        box_muller<R>(x.z, x.w, z2, z3);
        const int sel = a & 3;
        const R e = sel == 0 ? z0 : (sel == 1 ? z1 : (sel == 2 ? z2 : z3));
        gs.vec[a] = fma(sig, e, gs.La[k * n + a]);
      }
      __syncwarp();
      for (int a = k - 1; a >= 0; --a) {
        R acc = gs.vec[a];
        if (k <= 32) {
          for (int m = a + 1; m < k; ++m) acc = fma(-gs.La[m * n + a], gs.vec[m], acc);
        } else {
          R part = 0;
          for (int m = a + 1 + lane; m < k; m += 32) part = fma(gs.La[m * n + a], gs.vec[m], part);
          acc -= warp_sum(part);
        }
        const R xa = acc / gs.La[a * n + a];
        __syncwarp();
        if (lane == 0) gs.vec[a] = xa;
        __syncwarp();
      }
      for (int j = lane; j < p; j += 32) w_s[j] = 0;
      __syncwarp();
      for (int a = lane; a < k; a += 32) w_s[(int)gs.idx[a]] = gs.vec[a];
      __syncwarp();
    }
  }

  // One regression step: inclusion indicators (gam), sigma_obs^2 (s_e), weights -> w_s.
  // Inputs: gs.bvec = X'(targets), yty = |targets|^2 over observed steps.
  __device__ __forceinline__ void step(int it, const GibbsPlan& plan, uint32_t (&gam)[4], double yty,
                                       double& s_e, R* w_s) const {
    double ss = yty;
    int k = 0;
    if (p > 0) {
      double lm_cur = log_marginal(gam, yty, k, ss);
      if (plan.sparse) {
        make_order(it, plan);
        for (int jj = 0; jj < p; ++jj) {
          const int j = order(jj, plan);
          uint32_t gf[4] = {gam[0], gam[1], gam[2], gam[3]};
          gf[j >> 5] ^= 1u << (j & 31);
          int kf; double ssf;
          const double lm_new = log_marginal(gf, yty, kf, ssf);
          const bool cur = (gam[j >> 5] >> (j & 31)) & 1u;
          if (flips(it, j, cur, lm_cur, lm_new, plan)) { gam[j >> 5] ^= 1u << (j & 31); lm_cur = lm_new; }
        }
        lm_cur = log_marginal(gam, yty, k, ss);      // factor of the final configuration
      }
      (void)lm_cur;
    }
    draw(it, k, ss, s_e, w_s);
  }
};

template <typename R>
__global__ void __launch_bounds__(32 * (MAXG + 1), 1)
k_gibbs(ProbDev<R> pr, GibbsDev<R> gd, SmemCfg cfg, GibbsPlan plan, uint64_t seed,
        uint64_t chain_id0, int C, R* __restrict__ draws, R* __restrict__ level_out,
        R* __restrict__ traj_out, float* __restrict__ incl_out,
        const BatchDev<R>* __restrict__ batch) {
  extern __shared__ __align__(128) unsigned char smem[];
  // batched launch: grid.y = series; every series has C chains with the SAME global chain
  // ids (so a series' draws equal those of a single-series run), outputs are series-major
  const size_t series_row0 = batch ? (size_t)blockIdx.y * C * plan.n_results : 0;
  const size_t series_chain0 = batch ? (size_t)blockIdx.y * C : 0;
  if (batch) {
    pr = batch[blockIdx.y].pr; gd = batch[blockIdx.y].gd; plan.n_obs = batch[blockIdx.y].n_obs;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (blockDim.x >> 5) - 1;
  const int chain0 = blockIdx.x * G;
  const int nactive = min(G, C - chain0);
  const CtaShared<R> cs = cta_prologue(smem, cfg, pr, nactive);
  const int p = pr.p, dim = pr.dim, ld = pr.ld, NB = pr.NB, T = pr.T;
  // CTA-shared copy of the Gram matrix lives after the per-warp scratch
  R* gram_s = reinterpret_cast<R*>(smem + (((size_t)cfg.off_warp + (size_t)G * cfg.warp_bytes + 15) & ~(size_t)15));
  for (int i = threadIdx.x; i < p * p; i += blockDim.x) gram_s[i] = gd.gram[i];
  __syncthreads();
  const int n_iter = plan.n_warmup + plan.n_results;
  if (warp == G) {
    if (lane == 0) {
      omega_fetch(cs, pr);
      tile_producer(pr.tiles, cs.stage0, cs.full, cs.empty, cfg.stage_elems, cfg.nstage, NB,
                    cfg.resident != 0, 2LL * n_iter, [](long long s) { return (s & 1) == 0; });
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
  TilePipe<R> pipe = make_pipe(cs, cfg);
  const uint64_t gid = chain_id0 + (uint64_t)c + (batch ? (uint64_t)blockIdx.y * plan.series_stride : 0ull);
  const uint32_t id_lo = (uint32_t)gid, id_hi8 = (uint32_t)(gid >> 32) << 8;

  // ---- initial state: the reference's (lib.py:566-581) ----
  const R sd = Num<R>::sqrt(pr.P0);
  double s_e = p > 0 ? 0.2 * (double)pr.P0 : (double)pr.P0;           // (sqrt(1-0.8) sd)^2 or sd^2
  double s_h = (double)pr.lvl_scale / (double)pr.lvl_conc;            // (prior_level_sd sd)^2
  (void)sd;
  uint32_t gam[4] = {0u, 0u, 0u, 0u};                                 // weights start at 0
  if (!plan.sparse)
    for (int j = 0; j < p; ++j) gam[j >> 5] |= 1u << (j & 31);
  for (int j = lane; j < p; j += 32) { gs.bvec[j] = gd.xty0[j]; ws.w[j] = 0; }
  double yty = (double)gd.yty0;
  double incl_cnt[DSLOTS] = {0.0, 0.0, 0.0, 0.0};
  const double conc_e = (double)pr.obs_conc + 0.5 * plan.n_obs;
  __syncwarp();

  const GibbsReg<R> reg{pr, gs, gram_s, om_s, p, lane, conc_e, seed, id_lo, id_hi8};

  for (int it = 0; it < n_iter; ++it) {
    // =================== A. regression block ===================
    reg.step(it, plan, gam, yty, s_e, ws.w);
    // =================== B. level | rest : FFBS ===================
    const R se = (R)s_e, sh = (R)s_h, sig_e = (R)sqrt(s_e);
    R a_c = pr.m0, P_c = pr.P0;
    for (int b = 0; b < NB; ++b) {
      const R* tile = pipe.acquire(b);
      Blk<R> B;
      blk_residuals(B, tile, ws.w, p, ld, lane);
      if (lane == 0) { ws.ckpt[2 * b] = a_c; ws.ckpt[2 * b + 1] = P_c; }
      blk_forward<R, false>(B, se, sh, a_c, P_c, lane);
      pipe.release(lane);
    }
    __syncwarp();
    const bool keep = it >= plan.n_warmup;
    const size_t out_row = !keep ? 0
        : series_row0 + (plan.chain_major ? ((size_t)c * plan.n_results + (size_t)(it - plan.n_warmup))
                                          : ((size_t)(it - plan.n_warmup) * C + c));
    const XtMap xm = xt_map(p, lane);
    const bool small_p = p <= PSMALL;
    R accw[PSMALL];
#pragma unroll
    for (int j = 0; j < PSMALL; ++j) accw[j] = 0;
    R accg[JS];
#pragma unroll
    for (int s = 0; s < JS; ++s) accg[s] = 0;
    double n_yty = 0.0, n_d2 = 0.0;
    R x_c = 0;
    for (int b = NB - 1; b >= 0; --b) {
      const R* tile = pipe.acquire(b);
      Blk<R> B;
      R xw[KS];
      blk_residuals_xw(B, xw, tile, ws.w, p, ld, lane);
      a_c = ws.ckpt[2 * b]; P_c = ws.ckpt[2 * b + 1];
      blk_forward<R, true>(B, se, sh, a_c, P_c, lane);
      const int t0 = b * TB + lane * KS;
      R zs[KS], zp[KS];
#pragma unroll
      for (int kk = 0; kk < KS; kk += 2) {
        const uint4 x = Philox::gen(seed, id_lo, RNG_SMOOTH | id_hi8, (uint32_t)((t0 + kk) >> 1),
                                    (uint32_t)it);
        box_muller<R>(x.x, x.y, zs[kk], zp[kk]);
        box_muller<R>(x.z, x.w, zs[kk + 1], zp[kk + 1]);
      }
      R J[KS], off[KS];
      R m = 1, cc = 0;
#pragma unroll
      for (int kk = KS - 1; kk >= 0; --kk) {
        const R Cf = B.P[kk] * ((R)1 - B.K[kk]);
        const R Jk = (t0 + kk < T - 1) ? Cf * Num<R>::rcp(Cf + sh) : (R)0;
        const R Vk = Cf * ((R)1 - Jk);
        J[kk] = Jk;
        off[kk] = fma((R)1 - Jk, B.v[kk], Num<R>::sqrt(Vk) * zs[kk]);
        cc = fma(Jk, cc, off[kk]);
        m = Jk * m;
      }
affine_scan_down(m, cc, lane);
      R me = __shfl_down_sync(FULL, m, 1), ce = __shfl_down_sync(FULL, cc, 1);
      if (lane == 31) { me = 1; ce = 0; }
      R x = fma(me, x_c, ce);
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
      x_c = __shfl_sync(FULL, x, 0);
      n_yty += (double)ly; n_d2 += (double)ld2;
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
        R tr[KS];
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) tr[kk] = lv[kk] + xw[kk] + sig_e * zp[kk];
        if (level_out) store_run(level_out + out_row * T, t0, T, lv);
        if (traj_out) store_run(traj_out + out_row * T, t0, T, tr);
      }
      pipe.release(lane);
    }
    // publish the statistics of the next sweep
    yty = warp_sum(n_yty);
    const double d2 = warp_sum(n_d2);
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
        R a = accg[s];
        for (int o = xm.PJ; o < 32; o <<= 1) a += __shfl_xor_sync(FULL, a, o);
        const int j = lane + 32 * s;
        if (j < p && lane < xm.PJ) gs.bvec[j] = a;
      }
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
    if (keep) {
      R* row = draws + out_row * dim;
      for (int j = lane; j < p; j += 32) row[j] = ws.w[j];
      if (lane == 0) { row[p] = (R)log(s_e); row[p + 1] = (R)log(s_h); }
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
