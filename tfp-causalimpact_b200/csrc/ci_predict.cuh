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
#include <type_traits>
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

// Batched launch (grid.y = series): series_tiles / tile_stride give every series'
// tiles; theta [N,S,dim], level [N,S,T], mean [N,T].
template <typename R>
__global__ void __launch_bounds__(MEAN_COLS * MEAN_ROWS)
k_predict_mean(ProbDev<R> pr, const R* __restrict__ theta, const R* __restrict__ level, int S,
               R* __restrict__ mean, size_t tile_stride_elems = 0) {
  if (gridDim.y > 1 || tile_stride_elems) {
    const size_t sidx = blockIdx.y;
    pr.tiles += sidx * tile_stride_elems;
    theta += sidx * (size_t)S * pr.dim; level += sidx * (size_t)S * pr.T; mean += sidx * (size_t)pr.T;
  }
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

// The same mean for ONE series with many draws (ci_predictive_mean_d), split so that the whole
// GPU reads the [S,T] level paths: (1) k_mean_partial -- grid (T/32, SY): CTA (bx, by) sums columns
// 32 bx .. of draws [by S/SY, (by+1) S/SY) in float64 (a warp reads one 128-byte run of a row); the
// CTAs of column block 0 also sum their draws' weights; (2) k_mean_final adds the SY partial sums
// in order, divides, and adds x_t . wbar.  Deterministic.  (Run 25: the one-kernel version --
// T/8 CTAs of 1024 threads, each re-deriving wbar from all S draws -- took 55 us for 80 MB.)
constexpr int MEANP_COLS = 32, MEANP_ROWS = 8, MEANP_MAXY = 32;

template <typename R>
__global__ void __launch_bounds__(MEANP_COLS * MEANP_ROWS)
k_mean_partial(const R* __restrict__ theta, const R* __restrict__ level, int S, int T, int dim, int p,
               double* __restrict__ colpart /* [SY][T] */, double* __restrict__ wpart /* [SY][p] */) {
  __shared__ double part[MEANP_ROWS][MEANP_COLS + 1];
  const int tx = threadIdx.x, ty = threadIdx.y, by = blockIdx.y, SY = gridDim.y;
  const int s0 = (int)((long long)S * by / SY), s1 = (int)((long long)S * (by + 1) / SY);
  const int t = blockIdx.x * MEANP_COLS + tx;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;     // 4 loads in flight per thread
  if (t < T) {
    int s = s0 + ty;
    for (; s + 3 * MEANP_ROWS < s1; s += 4 * MEANP_ROWS) {
      a0 += (double)level[(size_t)s * T + t];
      a1 += (double)level[(size_t)(s + MEANP_ROWS) * T + t];
      a2 += (double)level[(size_t)(s + 2 * MEANP_ROWS) * T + t];
      a3 += (double)level[(size_t)(s + 3 * MEANP_ROWS) * T + t];
    }
    for (; s < s1; s += MEANP_ROWS) a0 += (double)level[(size_t)s * T + t];
  }
  part[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && t < T) {
    double tot = 0.0;
#pragma unroll
    for (int g = 0; g < MEANP_ROWS; ++g) tot += part[g][tx];
    colpart[(size_t)by * T + t] = tot;
  }
  if (blockIdx.x != 0) return;
  // weights of this draw range: thread (ty, tx) walks draws s0 + ty, + 8, ..; tx <-> covariate
  for (int j0 = 0; j0 < p; j0 += MEANP_COLS) {
    __syncthreads();
    const int j = j0 + tx;
    double acc = 0.0;
    if (j < p)
      for (int s = s0 + ty; s < s1; s += MEANP_ROWS) acc += (double)theta[(size_t)s * dim + j];
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && j < p) {
      double tot = 0.0;
#pragma unroll
      for (int g = 0; g < MEANP_ROWS; ++g) tot += part[g][tx];
      wpart[(size_t)by * p + j] = tot;
    }
  }
}

template <typename R>
__global__ void __launch_bounds__(256)
k_mean_final(ProbDev<R> pr, const double* __restrict__ colpart, const double* __restrict__ wpart,
             int S, int SY, R* __restrict__ mean) {
  __shared__ double wbar[MAX_DIM];
  const int p = pr.p, T = pr.T;
  for (int j = threadIdx.x; j < p; j += blockDim.x) {
    double tot = 0.0;
    for (int y = 0; y < SY; ++y) tot += wpart[(size_t)y * p + j];
    wbar[j] = tot / S;
  }
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  double tot = 0.0;
  for (int y = 0; y < SY; ++y) tot += colpart[(size_t)y * T + t];
  tot /= S;
  const int b = t / TB, tl = t - b * TB;
  const R* row = pr.tiles + (size_t)b * tile_elems(p) + tile_off(tl, pr.ld);
  for (int j = 0; j < p; ++j) tot += (double)row[j] * wbar[j];
  mean[t] = (R)tot;
}

// ---------------------------------------------------------------------------
// K5: one CTA per time column.  The column's S values are gathered into shared
// memory as order-preserving integer keys; ALL needed order statistics (floor /
// ceil rank of every quantile) are found together: one equal-width histogram of the
// VALUES locates the bin of each rank, the handful of keys in those bins is collected
// and each rank finished by counting (binned_select); columns that defeat the binning
// fall back to an exact bit-wise bisection on the key (bisect_select_impl).  History (B200, 10 000 draws per column):
// bitonic sort -> per-rank radix select with 11-bit shared-memory histograms (same-
// address atomics: 42 M bank conflicts per 10 000 x 2000 forecast, ~85 us per column
// CTA, ncu run 22) -> multi-rank 8-bit radix with warp-aggregated atomics (~75 us: the
// match.any and the serial regrouping cost what the saved sweeps gained, run 24) ->
// bisection alone (O(S bits ranks) compares: 1.1 ms, run 25) -> value binning.
// Interpolation is numpy's _lerp, bit for bit.
// ---------------------------------------------------------------------------
struct QuantArgs { double q[8]; int nq; };

template <typename R> struct KeyOf;
template <> struct KeyOf<float> {
  using type = uint32_t;
  static constexpr int NPASS = 4;
  static __device__ __forceinline__ uint32_t enc(float v) {
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  }
  static __device__ __forceinline__ float dec(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
  }
  static __device__ __forceinline__ uint32_t nan_key() { return 0xffffffffu; }
};
template <> struct KeyOf<double> {
  using type = unsigned long long;
  static constexpr int NPASS = 8;
  static __device__ __forceinline__ unsigned long long enc(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
  }
  static __device__ __forceinline__ double dec(unsigned long long k) {
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
  }
  static __device__ __forceinline__ unsigned long long nan_key() { return ~0ull; }
};

constexpr int QMAXR = 16;      // ranks selected together (8 quantiles x floor / ceil)

constexpr int LBINS = 2048;    // value bins of the fast path
constexpr int LCAP = 128;      // candidates kept per needed bin

template <typename R> struct SelectShared {
  using Key = typename KeyOf<R>::type;
  Key out[QMAXR];              // result: key of each rank
  int rank[QMAXR];             // in: 0-based ranks (any order, distinct)
  int cnt[3][QMAXR];           // bisection: CTA-wide counts, rotating buffers
  // fast path
  int hist[LBINS];
  Key cand[QMAXR][LCAP];       // keys that fell into a needed bin
  int ccount[QMAXR];
  int gbin[QMAXR], gbelow[QMAXR];   // per group: its bin, number of values in lower bins
  int rgrp[QMAXR];             // group of each rank
  int ngroups, fallback;
  int wsum[32];                // warp totals of the bin-count scan
  Key kmin, kmax;
};

// Where the select reads its keys from: shared memory (the column was encoded once), or --
// for columns too long for shared memory -- straight from global memory, encoding on the fly
// (every sweep re-reads the column through L2; same results, no size limit).
template <typename R> struct SmemKeys {
  const typename KeyOf<R>::type* k;
  __device__ __forceinline__ typename KeyOf<R>::type operator()(int i) const { return k[i]; }
};
template <typename R> struct GlobalKeys {
  const R* base;
  size_t stride;
  __device__ __forceinline__ typename KeyOf<R>::type operator()(int i) const {
    const R v = base[(size_t)i * stride];
    return (v == v) ? KeyOf<R>::enc(v) : KeyOf<R>::nan_key();
  }
};

// Keys of the sh.rank[0..nr) order statistics (0-based ranks, < n) of keys[0..n) ->
// sh.out[0..nr).  BIT-WISE BISECTION on the key value, all ranks at once: the result is
// built from the most significant bit down -- a bit is kept when the number of keys below
// the trial value does not exceed the rank.  One pass over the shared-memory keys and ONE
// barrier per bit; counting is compare + add in registers, warp totals by redux.sync, CTA
// totals by 32 integer atomics per rank on a rotating triple buffer -- no histogram, no
// same-address contention, deterministic.  Executed by the whole CTA.
template <typename R, int NRT, typename KeyAt>
__device__ __forceinline__ void bisect_select_impl(const KeyAt keys, int n, int nr,
                                                   SelectShared<R>& sh) {
  using Key = typename KeyOf<R>::type;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  constexpr int NB = (int)(8 * sizeof(Key));
  Key result[NRT];
  int rank[NRT];
#pragma unroll
  for (int r = 0; r < NRT; ++r) { result[r] = 0; rank[r] = r < nr ? sh.rank[r] : 0; }
  if (tid < 3 * QMAXR) sh.cnt[tid / QMAXR][tid % QMAXR] = 0;
  __syncthreads();
  for (int bit = NB - 1, itn = 0; bit >= 0; --bit, ++itn) {
    Key trial[NRT];
    int c[NRT];
#pragma unroll
    for (int r = 0; r < NRT; ++r) { trial[r] = result[r] | ((Key)1 << bit); c[r] = 0; }
    for (int i = tid; i < n; i += nt) {
      const Key key = keys(i);
#pragma unroll
      for (int r = 0; r < NRT; ++r) c[r] += (key < trial[r]) ? 1 : 0;
    }
    int* cnt = sh.cnt[itn % 3];
#pragma unroll
    for (int r = 0; r < NRT; ++r) {
      const int w = __reduce_add_sync(FULL, c[r]);
      if (lane == r && r < nr && w) atomicAdd(&cnt[r], w);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NRT; ++r)
      if (cnt[r] <= rank[r]) result[r] = trial[r];
    if (tid < QMAXR) sh.cnt[(itn + 2) % 3][tid] = 0;     // used again two iterations from now
  }
#pragma unroll
  for (int r = 0; r < NRT; ++r)
    if (tid == 0 && r < nr) sh.out[r] = result[r];
  __syncthreads();
}

// Fast path of the select: the order statistics of a column of (mostly) distinct real
// numbers.  (1) min / max; (2) ONE histogram of the values over LBINS equal-width bins --
// a monotone map, so bins are ordered like the values, and unlike the top bits of the key
// it spreads the column evenly (no same-address contention); (3) the bin of every wanted
// rank; (4) one more sweep collects the few keys of those bins; (5) each rank is finished
// by brute-force counting among its bin's candidates (exact, ties included).  ~25
// instructions per key instead of a pass per digit / bit.  Returns false -- nothing
// selected -- when the column defeats the binning (infinities, a bin with more than LCAP
// keys: heavy ties or extreme outliers); the caller then runs the bisection.
// `have_minmax`: the loader already left the smallest / largest valid key in sh.kmin / sh.kmax
// (load_contig_keys / load_block_keys of ci_impact.cuh fold that sweep into the load).
template <typename R, typename KeyAt>
__device__ __forceinline__ bool binned_select(const KeyAt keys, int n, int nr,
                                              SelectShared<R>& sh, bool have_minmax = false) {
  using Key = typename KeyOf<R>::type;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = nt >> 5;
  const Key NANK = KeyOf<R>::nan_key();
  if (tid == 0) {
    if (!have_minmax) { sh.kmin = NANK; sh.kmax = 0; }
    sh.fallback = 0; sh.ngroups = 0;
  }
  for (int b = tid; b < LBINS; b += nt) sh.hist[b] = 0;
  if (tid < QMAXR) sh.ccount[tid] = 0;
  __syncthreads();
  if (!have_minmax) {
    Key mn = NANK, mx = 0;
    for (int i = tid; i < n; i += nt) {
      const Key k = keys(i);
      if (k != NANK) { mn = k < mn ? k : mn; mx = k > mx ? k : mx; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const Key a = __shfl_xor_sync(FULL, mn, o), b2 = __shfl_xor_sync(FULL, mx, o);
      mn = a < mn ? a : mn; mx = b2 > mx ? b2 : mx;
    }
    if (lane == 0) { atomicMin(&sh.kmin, mn); atomicMax(&sh.kmax, mx); }
  }
  __syncthreads();
  const Key kmin = sh.kmin, kmax = sh.kmax;
  if (kmin == kmax) {                                 // a constant column
    if (tid < nr) sh.out[tid] = kmin;
    __syncthreads();
    return true;
  }
  const R lo = KeyOf<R>::dec(kmin), hi = KeyOf<R>::dec(kmax);
  const R scale = (R)LBINS / (hi - lo);
  if (!(scale > (R)0) || !(scale < (R)1e30) || !(lo - lo == (R)0) || !(hi - hi == (R)0))
    return false;                                     // infinities / overflowing range (uniform)
  auto bin_of = [&](Key k) -> int {
    const int b = (int)((KeyOf<R>::dec(k) - lo) * scale);
    return b < LBINS - 1 ? b : LBINS - 1;
  };
  for (int i = tid; i < n; i += nt) {
    const Key k = keys(i);
    if (k != NANK) atomicAdd(&sh.hist[bin_of(k)], 1);
  }
  __syncthreads();
  {
    // bins of the wanted ranks: every thread owns LBINS / nt consecutive bins (CTA-wide
    // exclusive scan of the bin counts: warp shuffles + one exchange of the warp totals).
    // (Run 21: warp 0 alone walking 64 bins per lane -- a 32-way bank conflict per read --
    // with the other warps parked at the barrier was 37% of the kernel's stall samples.)
    const int per = (LBINS + nt - 1) / nt, b0 = tid * per;
    int loc = 0;
    for (int b = b0; b < b0 + per && b < LBINS; ++b) loc += sh.hist[b];
    int inc = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sh.wsum[warp] = inc;
    __syncthreads();
    int before = inc - loc;
    for (int w = 0; w < warp; ++w) before += sh.wsum[w];
    if (loc > 0) {
      for (int r = 0; r < nr; ++r) {
        const int rk = sh.rank[r];
        if (rk >= before && rk < before + loc) {
          int acc = before, b = b0;
          while (acc + sh.hist[b] <= rk) { acc += sh.hist[b]; ++b; }
          sh.rgrp[r] = b;                             // (bin for now; group id below)
          sh.cnt[0][r] = acc;
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      int ng = 0;
      for (int r = 0; r < nr; ++r) {
        const int b = sh.rgrp[r];
        int g = -1;
        for (int gg = 0; gg < ng; ++gg)
          if (sh.gbin[gg] == b) g = gg;
        if (g < 0) {
          g = ng++;
          sh.gbin[g] = b; sh.gbelow[g] = sh.cnt[0][r];
          if (sh.hist[b] > LCAP) sh.fallback = 1;
        }
        sh.rgrp[r] = g;
      }
      sh.ngroups = ng;
    }
  }
  __syncthreads();
  if (sh.fallback) return false;
  const int ng = sh.ngroups;
  auto collect = [&](auto ngt) {                      // the needed bins, in registers
    constexpr int NG = decltype(ngt)::value;
    int gb[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) gb[g] = g < ng ? sh.gbin[g] : -1;
    for (int i = tid; i < n; i += nt) {
      const Key k = keys(i);
      if (k == NANK) continue;
      const int b = bin_of(k);
#pragma unroll
      for (int g = 0; g < NG; ++g)
        if (gb[g] == b) sh.cand[g][atomicAdd(&sh.ccount[g], 1)] = k;
    }
  };
  if (ng <= 2) collect(std::integral_constant<int, 2>{});
  else if (ng <= 4) collect(std::integral_constant<int, 4>{});
  else collect(std::integral_constant<int, QMAXR>{});
  __syncthreads();
  for (int r = warp; r < nr; r += nwarps) {           // one warp finishes one rank
    const int g = sh.rgrp[r], m = sh.ccount[g], kk = sh.rank[r] - sh.gbelow[g];
    for (int i = lane; i < m; i += 32) {
      const Key ci = sh.cand[g][i];
      int lt = 0, le = 0;
      for (int j = 0; j < m; ++j) {
        const Key cj = sh.cand[g][j];
        lt += cj < ci; le += cj <= ci;
      }
      if (lt <= kk && kk < le) sh.out[r] = ci;        // equal keys may race: same value
    }
  }
  __syncthreads();
  return true;
}

template <typename R, typename KeyAt>
__device__ void radix_select_multi(const KeyAt keys, int n, int nr, SelectShared<R>& sh,
                                   bool have_minmax = false) {
  if (binned_select<R>(keys, n, nr, sh, have_minmax)) return;
  if (nr <= 4) bisect_select_impl<R, 4>(keys, n, nr, sh);
  else if (nr <= 8) bisect_select_impl<R, 8>(keys, n, nr, sh);
  else bisect_select_impl<R, 16>(keys, n, nr, sh);
}

// Number of non-NaN keys of an accessor (the global-memory path has no load pass to count in).
template <typename R, typename KeyAt>
__device__ __forceinline__ int count_valid_keys(const KeyAt keys, int n, int* n_valid) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) *n_valid = 0;
  __syncthreads();
  int cnt = 0;
  for (int i = tid; i < n; i += nt) cnt += keys(i) != KeyOf<R>::nan_key() ? 1 : 0;
  cnt = __reduce_add_sync(FULL, cnt);
  if ((tid & 31) == 0 && cnt) atomicAdd(n_valid, cnt);
  __syncthreads();
  return *n_valid;
}

// Gather column t of a [S,T] array into shared-memory keys; returns the number of non-NaN.
template <typename R>
__device__ __forceinline__ int load_column_keys(const R* __restrict__ a, int S, int T, int t,
                                                typename KeyOf<R>::type* keys, int* n_valid) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) *n_valid = 0;
  __syncthreads();
  int cnt = 0;
  for (int i = tid; i < S; i += nt) {
    const R v = a[(size_t)i * T + t];
    const bool ok = (v == v);                      // pandas skips NaN
    keys[i] = ok ? KeyOf<R>::enc(v) : KeyOf<R>::nan_key();
    cnt += ok ? 1 : 0;
  }
  cnt = __reduce_add_sync(FULL, cnt);
  if ((tid & 31) == 0 && cnt) atomicAdd(n_valid, cnt);
  __syncthreads();
  return *n_valid;
}

// Register rank k in the shared rank list (deduplicated); returns its slot.  Thread 0 only.
template <typename R>
__device__ __forceinline__ int add_rank(SelectShared<R>& sh, int& nr, int k) {
  for (int i = 0; i < nr; ++i)
    if (sh.rank[i] == k) return i;
  sh.rank[nr] = k;
  return nr++;
}

template <typename R>
__global__ void __launch_bounds__(1024)
k_row_quantiles(const R* __restrict__ a, int S, int T, QuantArgs qa,
                                R* __restrict__ out, int out_ld, int in_smem) {
  using Key = typename KeyOf<R>::type;
  extern __shared__ __align__(16) unsigned char qsmem[];
  Key* keys = reinterpret_cast<Key*>(qsmem);
  __shared__ SelectShared<R> sh;
  __shared__ int n_valid, s_nr;
  __shared__ int slot_lo[8], slot_hi[8];
  const int t = blockIdx.x, tid = threadIdx.x;
  const GlobalKeys<R> gkeys{a + t, (size_t)T};
  const int n = in_smem ? load_column_keys<R>(a, S, T, t, keys, &n_valid)
                        : count_valid_keys<R>(gkeys, S, &n_valid);
  if (n == 0) {
    if (tid < qa.nq) out[(size_t)t * out_ld + tid] = Num<R>::nan();
    return;
  }
  if (tid == 0) {
    int nr = 0;
    for (int iq = 0; iq < qa.nq; ++iq) {
      const double pos = qa.q[iq] * (double)(n - 1);
      int lo = (int)floor(pos);
      lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
      const int hi = lo + 1 < n ? lo + 1 : n - 1;
      slot_lo[iq] = add_rank(sh, nr, lo);
      slot_hi[iq] = add_rank(sh, nr, hi);
    }
    s_nr = nr;
  }
  __syncthreads();
  if (in_smem) radix_select_multi<R>(SmemKeys<R>{keys}, S, s_nr, sh);
  else radix_select_multi<R>(gkeys, S, s_nr, sh);
  if (tid < qa.nq) {
    const double pos = qa.q[tid] * (double)(n - 1);
    int lo = (int)floor(pos);
    lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
    const R g = (R)(pos - (double)lo);
    const R va = KeyOf<R>::dec(sh.out[slot_lo[tid]]);
    const R vb = KeyOf<R>::dec(sh.out[slot_hi[tid]]);
    const R diff = vb - va;
    // numpy.lib._function_base_impl._lerp
    R res_v = va + diff * g;
    if (g >= (R)0.5) res_v = vb - diff * ((R)1 - g);
    if (g == (R)0) res_v = va;
    out[(size_t)t * out_ld + tid] = res_v;
  }
}

}  // namespace ci
