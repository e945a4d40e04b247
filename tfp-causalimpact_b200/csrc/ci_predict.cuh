// ci_predict.cuh -- K4: simulation smoother + one-step predictive draw, and
// K5: per-time quantiles across draws.
//
// K4 replaces (reference, relative to /root/reference) the latent resampling
// TFP performs in gibbs_sampler._resample_latents (LGSSM.posterior_sample; call
// site causalimpact/causalimpact_lib.py:365-388) and
// _get_posterior_means_and_trajectories (causalimpact_lib.py:609-632).
// One warp per posterior draw.  Forward sweep: filter, checkpoint (a,P) per
// tile.  Backward sweep: recompute the tile's filtered moments (m_t, C_t) and
// sample x_t = J_t x_{t+1} + (1-J_t) m_t + sqrt(V_t) z_t as a REVERSE AFFINE
// SCAN (J_t = C_t/(C_t+s_h), V_t = C_t (1-J_t); J = 0 at the last step), then
//   traj_t = x_t + x_t.w + sigma_obs * z'_t                 (lib.py:629-631).
// Normals come from Philox keyed by (seed, global draw id, t): the result does
// not depend on how draws are split over warps / CTAs / GPUs.
//
// K5 replaces posterior_processing.calculate_trajectory_quantiles
// (causalimpact/posterior_processing.py:25-60): pandas quantile(axis=1) =
// NaN-skipping linear interpolation at q*(n-1) (numpy's _lerp, bit for bit).
#pragma once
#include "ci_kernels.cuh"

namespace ci {

template <typename R>
__global__ void __launch_bounds__(32 * (MAXG + 1), 1)
k_predict(ProbDev<R> pr, SmemCfg cfg, const R* __restrict__ theta, int S, uint64_t seed,
          uint64_t draw_id0, R* __restrict__ level, R* __restrict__ traj) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (blockDim.x >> 5) - 1;
  const int s0 = blockIdx.x * G;
  const int nactive = min(G, S - s0);
  const CtaShared<R> cs = cta_prologue(smem, cfg, pr, nactive);
  if (warp == G) {
    if (lane == 0) omega_fetch(cs, pr);
    if (lane == 0)
      tile_producer(pr.tiles, cs.stage0, cs.full, cs.empty, cfg.stage_elems, cfg.nstage, pr.NB,
                    cfg.resident != 0, 2LL, [](long long s) { return (s & 1) == 0; });
    return;
  }
  if (warp >= nactive) return;

  const int s = s0 + warp;
  const int p = pr.p, dim = pr.dim, ld = pr.ld, NB = pr.NB, T = pr.T;
  const R* th = theta + (size_t)s * dim;
  const WarpScratch<R> ws = warp_scratch<R>(smem, cfg, warp);
  for (int j = lane; j < p; j += 32) ws.w[j] = th[j];
  const R s_e = Num<R>::exp(th[p]), s_h = Num<R>::exp(th[p + 1]);
  const R sig_e = Num<R>::sqrt(s_e);
  __syncwarp();
  TilePipe<R> pipe = make_pipe(cs, cfg);

  // ---- forward sweep: checkpoints only ----
  R a_c = pr.m0, P_c = pr.P0;
  for (int b = 0; b < NB; ++b) {
    const R* tile = pipe.acquire(b);
    Blk<R> B;
    blk_residuals(B, tile, ws.w, p, ld, lane);
    if (lane == 0) { ws.ckpt[2 * b] = a_c; ws.ckpt[2 * b + 1] = P_c; }
    blk_forward<R, false>(B, s_e, s_h, a_c, P_c, lane);
    pipe.release(lane);
  }
  __syncwarp();

  // ---- backward sweep: sample ----
  const uint64_t gid = draw_id0 + (uint64_t)s;
  const uint32_t c0 = (uint32_t)gid, c1 = RNG_SMOOTH | ((uint32_t)(gid >> 32) << 8);
  R x_c = 0;
  for (int b = NB - 1; b >= 0; --b) {
    const R* tile = pipe.acquire(b);
    Blk<R> B;
    R xw[KS];
    blk_residuals_xw(B, xw, tile, ws.w, p, ld, lane);
    a_c = ws.ckpt[2 * b]; P_c = ws.ckpt[2 * b + 1];
    blk_forward<R, true>(B, s_e, s_h, a_c, P_c, lane);   // B.v = filtered means
    const int t0 = b * TB + lane * KS;
    R zs[KS], zp[KS];
#pragma unroll
    for (int k = 0; k < KS; k += 2) {
      const uint4 x = Philox::gen(seed, c0, c1, (uint32_t)((t0 + k) >> 1), 0u);
      box_muller<R>(x.x, x.y, zs[k], zp[k]);
      box_muller<R>(x.z, x.w, zs[k + 1], zp[k + 1]);
    }
    R J[KS], off[KS];
    R m = 1, c = 0;
#pragma unroll
    for (int k = KS - 1; k >= 0; --k) {
      const R Cf = B.P[k] * ((R)1 - B.K[k]);
      const R Jk = (t0 + k < T - 1) ? Cf / (Cf + s_h) : (R)0;
      const R Vk = Cf * ((R)1 - Jk);
      J[k] = Jk;
      off[k] = fma((R)1 - Jk, B.v[k], Num<R>::sqrt(Vk) * zs[k]);
      c = fma(Jk, c, off[k]);
      m = Jk * m;
    }
affine_scan_down(m, c, lane);
    R me = __shfl_down_sync(FULL, m, 1), ce = __shfl_down_sync(FULL, c, 1);
    if (lane == 31) { me = 1; ce = 0; }
    R x = fma(me, x_c, ce);
    R lv[KS];
#pragma unroll
    for (int k = KS - 1; k >= 0; --k) { x = fma(J[k], x, off[k]); lv[k] = x; }
    x_c = __shfl_sync(FULL, x, 0);
    const size_t row = (size_t)s * T;
    R tr[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) tr[k] = lv[k] + xw[k] + sig_e * zp[k];
    if (level) store_run(level + row, t0, T, lv);
    store_run(traj + row, t0, T, tr);
    pipe.release(lane);
  }
}

// mean_t = (1/S) sum_s level[s,t] + x_t . wbar,  wbar = (1/S) sum_s w_s
// (lib.py:627: mixture mean == average of loc over draws).  Deterministic:
// fixed summation order, no atomics.  block = (32 columns, 32 row-groups).
template <typename R>
__global__ void k_predict_mean(ProbDev<R> pr, const R* __restrict__ theta,
                               const R* __restrict__ level, int S, R* __restrict__ mean) {
  __shared__ double part[32][33];
  __shared__ double wbar[MAX_DIM];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int p = pr.p, dim = pr.dim, T = pr.T;
  const int tid = ty * 32 + tx;
  for (int j = tid; j < p; j += 1024) {
    double a = 0.0;
    for (int s = 0; s < S; ++s) a += (double)theta[(size_t)s * dim + j];
    wbar[j] = a / S;
  }
  const int t = blockIdx.x * 32 + tx;
  double acc = 0.0;
  if (t < T)
    for (int s = ty; s < S; s += 32) acc += (double)level[(size_t)s * T + t];
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && t < T) {
    double tot = 0.0;
    for (int g = 0; g < 32; ++g) tot += part[g][tx];
    tot /= S;
    const int b = t / TB, tl = t - b * TB;
    const R* row = pr.tiles + (size_t)b * tile_elems(p) + tile_off(tl, pr.ld);
    for (int j = 0; j < p; ++j) tot += (double)row[j] * wbar[j];
    mean[t] = (R)tot;
  }
}

// ---------------------------------------------------------------------------
// K5: one CTA per time column; bitonic sort of the column in shared memory.
// ---------------------------------------------------------------------------
struct QuantArgs { double q[8]; int nq; };

template <typename R>
__global__ void k_row_quantiles(const R* __restrict__ a, int S, int T, int n_pad, QuantArgs qa,
                                R* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char qsmem[];
  R* buf = reinterpret_cast<R*>(qsmem);
  __shared__ int n_valid;
  const int t = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) n_valid = 0;
  __syncthreads();
  const R inf = sizeof(R) == 4 ? (R)CUDART_INF_F : (R)CUDART_INF;
  int cnt = 0;
  for (int i = tid; i < n_pad; i += nt) {
    R v = inf;
    if (i < S) {
      v = a[(size_t)i * T + t];
      if (v == v) ++cnt; else v = inf;      // pandas skips NaN
    }
    buf[i] = v;
  }
  if (cnt) atomicAdd(&n_valid, cnt);
  __syncthreads();
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int pid = tid; pid < (n_pad >> 1); pid += nt) {
        const int i = 2 * j * (pid / j) + (pid % j);
        const int ixj = i + j;
        const bool up = (i & k) == 0;
        const R x = buf[i], y = buf[ixj];
        if ((x > y) == up) { buf[i] = y; buf[ixj] = x; }
      }
      __syncthreads();
    }
  }
  if (tid < qa.nq) {
    const int n = n_valid;
    R res;
    if (n == 0) {
      res = Num<R>::nan();
    } else {
      const double pos = qa.q[tid] * (double)(n - 1);
      int lo = (int)floor(pos);
      if (lo < 0) lo = 0;
      if (lo > n - 1) lo = n - 1;
      const int hi = lo + 1 < n ? lo + 1 : n - 1;
      const R g = (R)(pos - (double)lo);
      const R va = buf[lo], vb = buf[hi];
      const R diff = vb - va;
      // numpy.lib._function_base_impl._lerp
      res = va + diff * g;
      if (g >= (R)0.5) res = vb - diff * ((R)1 - g);
      if (g == (R)0) res = va;   // guards inf - inf when hi is a +inf pad
    }
    out[(size_t)t * qa.nq + tid] = res;
  }
}

}  // namespace ci
