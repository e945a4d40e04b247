// ci_device.cuh -- device-side problem description, shared-memory carve-up and
// the per-chain evaluation routine shared by the log-prob and HMC kernels.
#pragma once
#include "ci_filter.cuh"

namespace ci {

constexpr int MAXG = 7;     // chains (consumer warps) per CTA; +1 producer warp = 256 threads
constexpr int JS = DSLOTS;  // covariate slots per lane in the X^T rbar product

// Problem constants as the kernels see them (built by ci_set_data).
template <typename R> struct ProbDev {
  const R* tiles;   // [NB][tile_elems]  swizzled [X|y] tiles
  const R* omega;   // [p,p] slab precision
  int T, p, ld, NB, dim, model;
  R m0, P0;
  R obs_conc, obs_scale, obs_ub;   // *_ub: bound on the VARIANCE (the host converts a scale bound)
  R lvl_conc, lvl_scale, lvl_ub;
};

// Seasonal components as the kernels see them (ci_set_seasonal; kernel in ci_seasonal.cuh).
constexpr int MAX_SEAS = 7;     // components (the "season ends" flags of a step are one byte)
constexpr int SEAS_MAXD = 192;  // 1 + sum of num_seasons: up to 6 state elements per lane (hour-of-week = 169)

struct SeasDev {
  int K, d;
  int n[MAX_SEAS], off[MAX_SEAS], n_ends[MAX_SEAS];
  const uint8_t* sched;         // [T][K+1]: active season of each component, then the ends mask
  double init_var;              // initial_effect_prior variance (lib.py:489: sd^2)
  double drift_conc, drift_scale, drift_ub;   // InverseGamma on the drift variance (lib.py:472-474); drift_ub bounds the VARIANCE
  void* scratch;                // [C][T][d+1] elements of R
  const double* per_series;     // batch only: [N][3] = init_var, drift_scale, drift_ub of every series
};

// Dynamic shared memory layout (byte offsets), computed on the host.
struct SmemCfg {
  uint32_t stage_elems;  // elements per stage (== tile_elems)
  uint32_t nstage;
  int resident;
  uint32_t off_full, off_empty, off_omega, off_warp;
  uint32_t warp_bytes;   // per-consumer-warp scratch
  // per-warp scratch sub-offsets (in elements of R)
  uint32_t w_off, rbuf_off, ckpt_off, extra_off;
  uint32_t total_bytes;
};

template <typename R> struct WarpScratch {
  R* w;      // [p]        regression weights of this chain (broadcast reads)
  R* rbuf;   // [TB + 8]   rbar of the current tile, padded like the tile rows
  R* ckpt;   // [2*NB]     (a, P) at every tile start
  R* extra;  // kernel-specific
};

template <typename R>
__device__ __forceinline__ WarpScratch<R> warp_scratch(unsigned char* smem, const SmemCfg& cfg,
                                                       int warp) {
  R* base = reinterpret_cast<R*>(smem + cfg.off_warp + (size_t)warp * cfg.warp_bytes);
  WarpScratch<R> ws;
  ws.w = base + cfg.w_off; ws.rbuf = base + cfg.rbuf_off; ws.ckpt = base + cfg.ckpt_off;
  ws.extra = base + cfg.extra_off;
  return ws;
}

// Lane <-> covariate mapping of the transposed product.
struct XtMap { int PJ, nparts, jj, part; };
__device__ __forceinline__ XtMap xt_map(int p, int lane) {
  XtMap m;
  int PJ = 1;
  while (PJ < p && PJ < 32) PJ <<= 1;
  m.PJ = PJ; m.nparts = 32 / PJ; m.jj = lane & (PJ - 1); m.part = lane / PJ;
  return m;
}

// ---------------------------------------------------------------------------
// Kalman log-likelihood of one chain (+ adjoint gradient), warp-cooperative.
//   ws.w must hold the chain's weights.  On return (all lanes):
//     ll           log p(y | theta)
//     g_se, g_sh   d ll / d sigma_obs^2, d ll / d sigma_level^2
//     gw[s]        d ll / d w_j   for j = lane + 32 s   (valid on lanes < PJ)
// ---------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void chain_eval(TilePipe<R>& pipe, const ProbDev<R>& pr,
                                           const WarpScratch<R>& ws, R s_e, R s_h, bool want_grad,
                                           int lane, double& ll, double& g_se, double& g_sh,
                                           R (&gw)[JS]) {
  const int p = pr.p, ld = pr.ld, NB = pr.NB;
  double acc_ll = 0.0;
  int n_obs = 0;
  R a_c = pr.m0, P_c = pr.P0;
  for (int b = 0; b < NB; ++b) {
    const R* tile = pipe.acquire(b);
    Blk<R> B;
    blk_residuals(B, tile, ws.w, p, ld, lane);
    if (want_grad && lane == 0) { ws.ckpt[2 * b] = a_c; ws.ckpt[2 * b + 1] = P_c; }
    blk_forward(B, s_e, s_h, a_c, P_c, lane);
    acc_ll += (double)blk_loglik_terms(B, s_e);
    n_obs += __popc(B.obs);
    pipe.release(lane);
  }
  acc_ll += 1.8378770664093453 * (double)n_obs;
  ll = -0.5 * warp_sum(acc_ll);
  g_se = 0.0; g_sh = 0.0;
#pragma unroll
  for (int s = 0; s < JS; ++s) gw[s] = 0;
  if (!want_grad) return;
  __syncwarp();
  const XtMap xm = xt_map(p, lane);
  const bool small_p = p <= PSMALL;
  R accw[PSMALL];
#pragma unroll
  for (int j = 0; j < PSMALL; ++j) accw[j] = 0;
  R g1[1] = {0};
  R ab_c = 0, pb_c = 0;
  double ge = 0.0, gh = 0.0;
  for (int b = NB - 1; b >= 0; --b) {
    const R* tile = pipe.acquire(b);
    Blk<R> B;
    blk_residuals(B, tile, ws.w, p, ld, lane);
    a_c = ws.ckpt[2 * b]; P_c = ws.ckpt[2 * b + 1];
    blk_forward(B, s_e, s_h, a_c, P_c, lane);
    R rbar[KS];
    R lge = 0, lgh = 0;
    blk_backward(B, s_e, ab_c, pb_c, lane, lge, lgh, rbar);
    ge += (double)lge; gh += (double)lgh;
    if (small_p) {
      blk_xt_rbar_small(tile, rbar, p, ld, lane, accw);
    } else {
#pragma unroll
      for (int k = 0; k < KS; ++k) ws.rbuf[lane * KS + k + (lane >> 2)] = rbar[k];
      __syncwarp();
      if (p <= 32) blk_xt_rbar<R, 1>(tile, ws.rbuf, p, ld, xm.jj, xm.part, xm.nparts, g1);
      else blk_xt_rbar<R, JS>(tile, ws.rbuf, p, ld, xm.jj, xm.part, xm.nparts, gw);
      __syncwarp();
    }
    pipe.release(lane);
  }
  g_se = warp_sum(ge); g_sh = warp_sum(gh);
  if (small_p) {
    // lane j ends up owning d ll / d w_j
#pragma unroll
    for (int j = 0; j < PSMALL; ++j) {
      if (j < p) {
        const R tot = warp_sum(accw[j]);
        if (lane == j) gw[0] = -tot;
      }
    }
  } else {
    if (p <= 32) gw[0] = g1[0];
#pragma unroll
    for (int s = 0; s < JS; ++s) {
      R a = gw[s];
      for (int o = xm.PJ; o < 32; o <<= 1) a += __shfl_xor_sync(FULL, a, o);
      gw[s] = -a;   // r = y - Xw  =>  d ll/d w = -X^T rbar
    }
  }
}

// log prior + log|Jacobian| in theta coordinates (oracle/kalman_np.py:log_prior;
// priors of causalimpact_lib.py:424-462).  Adds the prior gradient into
// (gw, g_u, g_l); returns lp, or -inf outside the reference's upper bounds.
template <typename R>
__device__ __forceinline__ double chain_prior(const ProbDev<R>& pr, const R* om_s, const R* w_s,
                                              R u, R l, R s_e, R s_h, int lane, R (&gw)[JS],
                                              double& g_u, double& g_l) {
  const int p = pr.p;
  const R rse = (R)1 / s_e, rsh = (R)1 / s_h;
  double lp = -((double)pr.obs_conc + 1.0) * u - (double)pr.obs_scale * rse + u
            - ((double)pr.lvl_conc + 1.0) * l - (double)pr.lvl_scale * rsh + l;
  g_u += -((double)pr.obs_conc + 1.0) + (double)pr.obs_scale * rse + 1.0;
  g_l += -((double)pr.lvl_conc + 1.0) + (double)pr.lvl_scale * rsh + 1.0;
  if (p > 0) {
    R q = 0;
#pragma unroll
    for (int s = 0; s < JS; ++s) {
      const int j = lane + 32 * s;
      if (j < p) {
        R ow = 0;
        for (int i = 0; i < p; ++i) ow = fma(om_s[i * p + j], w_s[i], ow);   // Omega symmetric
        q = fma(w_s[j], ow, q);
        gw[s] -= ow * rse;
      }
    }
    q = warp_sum(q);
    lp += -0.5 * p * (double)u - 0.5 * (double)q * rse;
    g_u += -0.5 * p + 0.5 * (double)q * rse;
  }
  const bool ok = (s_e <= pr.obs_ub) && (s_h <= pr.lvl_ub) &&
                  (u == u) && (l == l);
  return ok ? lp : -CUDART_INF;
}

// Store a lane's KS consecutive values dst[t0 .. t0+KS) (t < T guarded).  When the
// address is 16-byte aligned and the whole run is in range, two 128-bit stores
// replace eight scalar ones.
template <typename R>
__device__ __forceinline__ void store_run(R* __restrict__ dst, int t0, int T, const R (&v)[KS]) {
  R* p = dst + t0;
  if (sizeof(R) == 4 && t0 + KS <= T && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) {
    float4* q = reinterpret_cast<float4*>(p);
    q[0] = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
    q[1] = make_float4((float)v[4], (float)v[5], (float)v[6], (float)v[7]);
  } else {
#pragma unroll
    for (int k = 0; k < KS; ++k)
      if (t0 + k < T) p[k] = v[k];
  }
}

}  // namespace ci
