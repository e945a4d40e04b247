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

#ifndef CI_PREDICT_MINB
#define CI_PREDICT_MINB 1
#endif
template <typename R>
__global__ void __launch_bounds__(32 * (MAXG + 1), CI_PREDICT_MINB)
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
// block = (MEAN_COLS columns, MEAN_ROWS row-groups): 8 consecutive floats of a row are one
// 32-byte sector, so T/8 CTAs spread over the whole GPU while every load stays sector-exact.
constexpr int MEAN_COLS = 8, MEAN_ROWS = 128;

template <typename R>
__global__ void __launch_bounds__(MEAN_COLS * MEAN_ROWS)
k_predict_mean(ProbDev<R> pr, const R* __restrict__ theta, const R* __restrict__ level, int S,
               R* __restrict__ mean) {
  __shared__ double part[MEAN_ROWS][MEAN_COLS + 1];
  __shared__ double wpart[32][8];
  __shared__ double wbar[MAX_DIM];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int p = pr.p, dim = pr.dim, T = pr.T;
  const int tid = ty * MEAN_COLS + tx, lane = tid & 31, warp = tid >> 5;
  // wbar_j = mean_s w[s][j]: all 1024 threads stride over the draws, 8 covariates at a
  // time; warp sums, then a fixed-order sum over the 32 warps (deterministic).
  for (int j0 = 0; j0 < p; j0 += 8) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = tid; s < S; s += 1024) {
      const R* row = theta + (size_t)s * dim + j0;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
        if (j0 + jj < p) acc[jj] += (double)row[jj];
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const double v = warp_sum(acc[jj]);
      if (lane == 0) wpart[warp][jj] = v;
    }
    __syncthreads();
    if (tid < 8 && j0 + tid < p) {
      double tot = 0.0;
      for (int w = 0; w < 32; ++w) tot += wpart[w][tid];
      wbar[j0 + tid] = tot / S;
    }
    __syncthreads();
  }
  const int t = blockIdx.x * MEAN_COLS + tx;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;     // 4 loads in flight per thread
  if (t < T) {
    int s = ty;
    for (; s + 3 * MEAN_ROWS < S; s += 4 * MEAN_ROWS) {
      a0 += (double)level[(size_t)s * T + t];
      a1 += (double)level[(size_t)(s + MEAN_ROWS) * T + t];
      a2 += (double)level[(size_t)(s + 2 * MEAN_ROWS) * T + t];
      a3 += (double)level[(size_t)(s + 3 * MEAN_ROWS) * T + t];
    }
    for (; s < S; s += MEAN_ROWS) a0 += (double)level[(size_t)s * T + t];
  }
  part[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && t < T) {
    double tot = 0.0;
    for (int g = 0; g < MEAN_ROWS; ++g) tot += part[g][tx];
    tot /= S;
    const int b = t / TB, tl = t - b * TB;
    const R* row = pr.tiles + (size_t)b * tile_elems(p) + tile_off(tl, pr.ld);
    for (int j = 0; j < p; ++j) tot += (double)row[j] * wbar[j];
    mean[t] = (R)tot;
  }
}

// ---------------------------------------------------------------------------
// K5: one CTA per time column.  The column's S values are gathered into shared
// memory as order-preserving integer keys; every needed order statistic
// (floor / ceil rank of each quantile) is found by an exact MSB-first RADIX
// SELECT: per 11-bit digit one histogram sweep over the keys that still match
// the prefix (shared-memory atomics, integer => deterministic), a warp-level
// search for the bin that contains the rank, repeat.  3 sweeps per rank for
// float32, 6 for float64 -- O(S) work per rank instead of the O(S log^2 S)
// of a bitonic sort (round-1 run 12: the sort took 1.8 ms of a 2.8 ms
// 10000-draw forecast).  Interpolation is numpy's _lerp, bit for bit.
// ---------------------------------------------------------------------------
struct QuantArgs { double q[8]; int nq; };

template <typename R> struct KeyOf;
template <> struct KeyOf<float> {
  using type = uint32_t;
  static constexpr int NPASS = 3;
  static __device__ __forceinline__ uint32_t enc(float v) {
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  }
  static __device__ __forceinline__ float dec(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
  }
  static __device__ __forceinline__ int shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
  static __device__ __forceinline__ int bits(int pass) { return pass == 2 ? 10 : 11; }
  static __device__ __forceinline__ uint32_t nan_key() { return 0xffffffffu; }
};
template <> struct KeyOf<double> {
  using type = unsigned long long;
  static constexpr int NPASS = 6;
  static __device__ __forceinline__ unsigned long long enc(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
  }
  static __device__ __forceinline__ double dec(unsigned long long k) {
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
  }
  static __device__ __forceinline__ int shift(int pass) { return pass < 5 ? 53 - 11 * pass : 0; }
  static __device__ __forceinline__ int bits(int pass) { return pass < 5 ? 11 : 9; }
  static __device__ __forceinline__ unsigned long long nan_key() { return ~0ull; }
};

constexpr int QBINS = 2048;

// k-th smallest key (0-based) among keys[0..n); executed by the whole CTA.
template <typename R>
__device__ typename KeyOf<R>::type radix_select(const typename KeyOf<R>::type* keys, int n, int k,
                                                int* hist, int* res) {
  using Key = typename KeyOf<R>::type;
  Key prefix = 0, mask = 0;
  int remaining = k;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int pass = 0; pass < KeyOf<R>::NPASS; ++pass) {
    const int sh = KeyOf<R>::shift(pass), nb = 1 << KeyOf<R>::bits(pass);
    for (int b = tid; b < nb; b += nt) hist[b] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
      const Key key = keys[i];
      if ((key & mask) == prefix) atomicAdd(&hist[(int)((key >> sh) & (Key)(nb - 1))], 1);
    }
    __syncthreads();
    if (tid < 32) {                       // warp 0 locates the bin holding the rank
      const int per = nb / 32;
      int loc = 0;
      for (int b = 0; b < per; ++b) loc += hist[tid * per + b];
      int inc = loc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (tid >= o) inc += t;
      }
      const int before = inc - loc;
      if (remaining >= before && remaining < inc) {
        int acc = before, b = tid * per;
        while (acc + hist[b] <= remaining) { acc += hist[b]; ++b; }
        res[0] = b; res[1] = acc;
      }
    }
    __syncthreads();
    prefix |= (Key)res[0] << sh;
    mask |= (Key)(nb - 1) << sh;
    remaining -= res[1];
    __syncthreads();
  }
  return prefix;
}

template <typename R>
__global__ void k_row_quantiles(const R* __restrict__ a, int S, int T, QuantArgs qa,
                                R* __restrict__ out, int out_ld) {
  using Key = typename KeyOf<R>::type;
  extern __shared__ __align__(16) unsigned char qsmem[];
  Key* keys = reinterpret_cast<Key*>(qsmem);
  __shared__ int hist[QBINS];
  __shared__ int res[2];
  __shared__ int n_valid;
  const int t = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) n_valid = 0;
  __syncthreads();
  int cnt = 0;
  for (int i = tid; i < S; i += nt) {
    const R v = a[(size_t)i * T + t];
    const bool ok = (v == v);                      // pandas skips NaN
    keys[i] = ok ? KeyOf<R>::enc(v) : KeyOf<R>::nan_key();
    cnt += ok ? 1 : 0;
  }
  cnt = __reduce_add_sync(FULL, cnt);
  if ((tid & 31) == 0 && cnt) atomicAdd(&n_valid, cnt);
  __syncthreads();
  const int n = n_valid;
  for (int iq = 0; iq < qa.nq; ++iq) {
    R res_v;
    if (n == 0) {
      res_v = Num<R>::nan();
    } else {
      const double pos = qa.q[iq] * (double)(n - 1);
      int lo = (int)floor(pos);
      if (lo < 0) lo = 0;
      if (lo > n - 1) lo = n - 1;
      const int hi = lo + 1 < n ? lo + 1 : n - 1;
      const R g = (R)(pos - (double)lo);
      const R va = KeyOf<R>::dec(radix_select<R>(keys, S, lo, hist, res));
      const R vb = (hi == lo || g == (R)0) ? va
                                           : KeyOf<R>::dec(radix_select<R>(keys, S, hi, hist, res));
      const R diff = vb - va;
      // numpy.lib._function_base_impl._lerp
      res_v = va + diff * g;
      if (g >= (R)0.5) res_v = vb - diff * ((R)1 - g);
      if (g == (R)0) res_v = va;
    }
    if (tid == 0) out[(size_t)t * out_ld + iq] = res_v;
  }
}

}  // namespace ci
