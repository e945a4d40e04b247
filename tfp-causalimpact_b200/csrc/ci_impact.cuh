// ci_impact.cuh -- SURVEY section 8 row f1: the impact series and summary on
// the device, straight from the predictive trajectories the smoother / Gibbs
// kernel left in HBM (no host round trip of the [S,T] draws).
//
// Replaces (reference, relative to /root/reference) the O(S*T) pandas work of
//   causalimpact/causalimpact_lib.py:793-837   point / cumulative effect paths
//   causalimpact/causalimpact_lib.py:840-931   per-time quantiles of the three families
//   causalimpact/causalimpact_lib.py:934-1093  post-period per-draw mean / sum,
//                                              their quantiles, sd, rel. effect, p-value
//   causalimpact/posterior_processing.py:63-98 un-standardising (standardize.py:60-64)
// All arithmetic is float64 (the reference converts the float32 draws to a
// float64 frame first).  Order statistics commute with the monotone
// un-standardising map, so the select runs on the raw keys and only the two
// selected values are un-scaled; the point-effect quantiles reuse the same
// column: the k-th smallest of (y_t - x) is y_t minus the k-th LARGEST x.
//
// Two launches:
//   k_impact_rows  one warp per draw, 32 draws per CTA (+1 CTA for the predictive mean):
//                  cumulative effect paths (warp scan over time), post-period per-draw
//                  statistics; the raw draws and the cumulative paths leave the CTA
//                  TRANSPOSED ([T,S], through a 32x32 shared-memory tile, coalesced on both
//                  sides) so that every column job below reads one contiguous run.  (Run 22:
//                  gathering a column from the [S,T] layout cost a 32-byte sector per 4-byte
//                  value and 0.6 ms of the 1.25 ms call.)
//   k_impact_jobs  one CTA per column job, all independent: T_c cumulative-effect columns
//                  (float64 keys), the 5 per-draw statistics, T prediction / point-effect
//                  columns (raw keys), and one CTA for sd / mean / tail counts.  Each job =
//                  contiguous key load + the multi-rank radix select of ci_predict.cuh.
#pragma once
#include "ci_predict.cuh"

namespace ci {

constexpr int IMP_SERIES_COLS = 9;   // mean, pred lo/hi, point mean/lo/hi, cum mean/lo/hi
constexpr int IMP_STATS = 5;         // per draw: pred mean, pred sum, effect mean, effect sum, rel
constexpr int IMP_SUMMARY_LEN = 20;

struct ImpactDev {
  int S, T, t_c0, n_post;            // t_c0: first step that is not before the post-period
  double scale, offset, q_lo, q_hi, obs_sum;
};
// per-series part of ImpactDev for batched launches (grid.y = series)
struct ImpactSeries { double scale, offset, obs_sum; };

// standardize.py:60-64: (x * stddev) + mean, two roundings like numpy (no FMA contraction)
__device__ __forceinline__ double imp_unscale(double x, double scale, double offset) {
  return __dadd_rn(__dmul_rn(x, scale), offset);
}

constexpr int IMP_TILE = 32;         // draws per CTA == time steps per chunk

// Row r < S: draw r of traj (CTA b holds draws 32b .. 32b+31, warp w <-> draw, lane <-> t);
// the LAST CTA's warp 0 handles the predictive mean (its cumulative path and post-period
// mean / sum are the *_mean series columns and `predicted`).
constexpr int IMP_CH = 2;            // 32-step chunks per loop iteration (loads overlap)

template <typename R>
__global__ void __launch_bounds__(32 * IMP_TILE)
k_impact_rows(const R* __restrict__ traj, const R* __restrict__ mean,
              const double* __restrict__ obs, const uint8_t* __restrict__ period, ImpactDev a,
              R* __restrict__ trT, double* __restrict__ cumT, double* __restrict__ statsT,
              double* __restrict__ series, double* __restrict__ summ,
              const ImpactSeries* __restrict__ per = nullptr) {
  if (per) {      // batched: this CTA row works on series blockIdx.y (obs is [N,T], period shared)
    const size_t sidx = blockIdx.y;
    a.scale = per[sidx].scale; a.offset = per[sidx].offset; a.obs_sum = per[sidx].obs_sum;
    traj += sidx * (size_t)a.S * a.T; mean += sidx * (size_t)a.T; obs += sidx * (size_t)a.T;
    trT += sidx * (size_t)a.S * a.T; cumT += sidx * (size_t)a.S * (a.T - a.t_c0);
    statsT += sidx * (size_t)a.S * IMP_STATS;
    series += sidx * (size_t)a.T * IMP_SERIES_COLS; summ += sidx * (size_t)IMP_SUMMARY_LEN;
  }
  __shared__ R tile_raw[IMP_TILE][IMP_CH * IMP_TILE + 1];
  __shared__ double tile_cum[IMP_TILE][IMP_CH * IMP_TILE + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool is_mean = blockIdx.x == gridDim.x - 1;
  const int r0 = blockIdx.x * IMP_TILE;
  const int r = is_mean ? a.S : r0 + warp;
  if (is_mean && warp != 0) return;                  // (no CTA-wide barrier on this path)
  const bool row_ok = is_mean || r < a.S;
  const R* src = is_mean ? mean : traj + (size_t)(row_ok ? r : 0) * a.T;
  double carry = 0.0, pred_sum = 0.0, eff_sum = 0.0;
  int eff_cnt = 0;
  for (int base = 0; base < a.T; base += IMP_CH * IMP_TILE) {
    R raw[IMP_CH];
    double ob[IMP_CH];
    int per[IMP_CH];
#pragma unroll
    for (int h = 0; h < IMP_CH; ++h) {               // all loads of the iteration first
      const int t = base + h * IMP_TILE + lane;
      const bool valid = row_ok && t < a.T;
      raw[h] = valid ? src[t] : (R)0;
      ob[h] = valid ? obs[t] : CUDART_NAN;
      per[h] = valid ? (int)period[t] : 0;
    }
    // nothing to accumulate before the post-period starts: cumulative effect is 0 there
    const bool need_cum = base + IMP_CH * IMP_TILE > a.t_c0;
#pragma unroll
    for (int h = 0; h < IMP_CH; ++h) {
      const int t = base + h * IMP_TILE + lane;
      const bool valid = row_ok && t < a.T;
      const double x = imp_unscale((double)raw[h], a.scale, a.offset);
      const double pt = valid ? ob[h] - x : CUDART_NAN;            // lib.py:822-823
      const bool isn = !(pt == pt);
      double out = 0.0;
      if (need_cum) {
        // lib.py:826-831: effects before the post-period count as 0; NaNs are skipped
        double inc = (valid && t >= a.t_c0 && !isn) ? pt : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_up_sync(FULL, inc, o);
          if (lane >= o) inc += up;
        }
        const double cv = carry + inc;
        carry += __shfl_sync(FULL, inc, 31);
        out = (t >= a.t_c0 && isn) ? CUDART_NAN : cv;
      }
      if (valid && per[h] == 1) {                    // inside the post-period (lib.py:966-1011)
        pred_sum += x;
        if (!isn) { eff_sum += pt; ++eff_cnt; }
      }
      if (is_mean) {
        if (valid) {
          double* row = series + (size_t)t * IMP_SERIES_COLS;
          row[0] = x; row[3] = pt; row[6] = out;
        }
      } else {
        tile_raw[warp][h * IMP_TILE + lane] = raw[h];
        if (need_cum) tile_cum[warp][h * IMP_TILE + lane] = out;
      }
    }
    if (is_mean) continue;
    // transpose the chunk: [draw][t] -> [t][draw]; warp <-> time step, lane <-> draw
    __syncthreads();
    const int rr = r0 + lane;
#pragma unroll
    for (int h = 0; h < IMP_CH; ++h) {
      const int tc = base + h * IMP_TILE + warp;
      if (tc < a.T && rr < a.S) {
        trT[(size_t)tc * a.S + rr] = tile_raw[lane][h * IMP_TILE + warp];
        if (tc >= a.t_c0) cumT[(size_t)(tc - a.t_c0) * a.S + rr] = tile_cum[lane][h * IMP_TILE + warp];
      }
    }
    __syncthreads();
  }
  pred_sum = warp_sum(pred_sum);
  eff_sum = warp_sum(eff_sum);
  eff_cnt = __reduce_add_sync(FULL, eff_cnt);
  if (lane == 0 && row_ok) {
    const double pm = pred_sum / (double)a.n_post;
    if (is_mean) {
      summ[18] = pm; summ[19] = pred_sum;
    } else {
      statsT[0 * (size_t)a.S + r] = pm;
      statsT[1 * (size_t)a.S + r] = pred_sum;
      statsT[2 * (size_t)a.S + r] = eff_cnt > 0 ? eff_sum / (double)eff_cnt : CUDART_NAN;
      statsT[3 * (size_t)a.S + r] = eff_sum;
      statsT[4 * (size_t)a.S + r] = a.obs_sum / pred_sum - 1.0;      // lib.py:1010-1011
    }
  }
}

// numpy.lib._function_base_impl._lerp in float64
__device__ __forceinline__ double imp_lerp(double va, double vb, double g) {
  const double diff = vb - va;
  double r = va + diff * g;
  if (g >= 0.5) r = vb - diff * (1.0 - g);
  if (g == 0.0) r = va;
  return r;
}

// Contiguous column -> shared-memory keys; returns the number of non-NaN values.
template <typename V>
__device__ __forceinline__ int load_contig_keys(const V* __restrict__ col, int S,
                                                typename KeyOf<V>::type* keys, int* n_valid) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) *n_valid = 0;
  __syncthreads();
  int cnt = 0;
  for (int i = tid; i < S; i += nt) {
    const V v = col[i];
    const bool ok = (v == v);
    keys[i] = ok ? KeyOf<V>::enc(v) : KeyOf<V>::nan_key();
    cnt += ok ? 1 : 0;
  }
  cnt = __reduce_add_sync(FULL, cnt);
  if ((tid & 31) == 0 && cnt) atomicAdd(n_valid, cnt);
  __syncthreads();
  return *n_valid;
}

// (q_lo, q_hi) quantiles of one contiguous column of V; MIRROR also selects the mirrored
// ranks: the k-th smallest of (o - x) is o minus the k-th LARGEST x.  Results (as V
// values, not yet un-scaled) in res[iq][0..3] = lo, hi, mirrored lo, mirrored hi; g[iq] =
// interpolation weight.  Returns n (0 => no valid value).  Whole CTA.
template <typename V, bool MIRROR>
__device__ __forceinline__ int column_quantiles(const V* __restrict__ col, int S, double q_lo,
                                                double q_hi, unsigned char* key_mem, int in_smem,
                                                SelectShared<V>& sh, int* ibuf, double (&res)[2][4],
                                                double (&g)[2]) {
  using Key = typename KeyOf<V>::type;
  Key* keys = reinterpret_cast<Key*>(key_mem);
  int* n_valid = ibuf;                // [0]; [1] = number of ranks; [2..9] = slots
  const GlobalKeys<V> gkeys{col, (size_t)1};
  const int n = in_smem ? load_contig_keys<V>(col, S, keys, n_valid)
                        : count_valid_keys<V>(gkeys, S, n_valid);
  if (n == 0) return 0;
  if (threadIdx.x == 0) {
    int nr = 0;
    for (int iq = 0; iq < 2; ++iq) {
      const double pos = (iq == 0 ? q_lo : q_hi) * (double)(n - 1);
      int lo = (int)floor(pos);
      lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
      const int hi = lo + 1 < n ? lo + 1 : n - 1;
      ibuf[2 + 4 * iq + 0] = add_rank(sh, nr, lo);
      ibuf[2 + 4 * iq + 1] = add_rank(sh, nr, hi);
      if (MIRROR) {
        ibuf[2 + 4 * iq + 2] = add_rank(sh, nr, n - 1 - lo);
        ibuf[2 + 4 * iq + 3] = add_rank(sh, nr, n - 1 - hi);
      }
    }
    ibuf[1] = nr;
  }
  __syncthreads();
  if (in_smem) radix_select_multi<V>(SmemKeys<V>{keys}, S, ibuf[1], sh);
  else radix_select_multi<V>(gkeys, S, ibuf[1], sh);
#pragma unroll
  for (int iq = 0; iq < 2; ++iq) {
    const double pos = (iq == 0 ? q_lo : q_hi) * (double)(n - 1);
    int lo = (int)floor(pos);
    lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
    g[iq] = pos - (double)lo;
#pragma unroll
    for (int j = 0; j < (MIRROR ? 4 : 2); ++j)
      res[iq][j] = (double)KeyOf<V>::dec(sh.out[ibuf[2 + 4 * iq + j]]);
  }
  return n;
}

// sd (ddof = 1), mean relative effect and the tail counts of the p-value (lib.py:1021-1090).
// One CTA, fixed reduction order: deterministic.
__device__ __forceinline__ void impact_summary_block(const double* __restrict__ statsT,
                                                     const ImpactDev& a, double* __restrict__ summ,
                                                     double* red /*[33] shared*/) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int S = a.S, nw = nt >> 5;
  auto block_sum = [&](double v) -> double {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < nw; ++w) tot += red[w];
      red[32] = tot;
    }
    __syncthreads();
    return red[32];
  };
  for (int j = 0; j < IMP_STATS; ++j) {
    const double* col = statsT + (size_t)j * S;
    double s = 0.0;
    for (int i = tid; i < S; i += nt) s += col[i];
    const double m = block_sum(s) / (double)S;
    double ss = 0.0;
    for (int i = tid; i < S; i += nt) {
      const double d = col[i] - m;
      ss += d * d;
    }
    const double tot = block_sum(ss);
    if (tid == 0) {
      summ[10 + j] = S > 1 ? sqrt(tot / (double)(S - 1)) : CUDART_NAN;
      if (j == IMP_STATS - 1) summ[15] = m;
    }
  }
  double le = 0.0, ge = 0.0;
  const double* ps = statsT + (size_t)S;
  for (int i = tid; i < S; i += nt) {
    le += (a.obs_sum <= ps[i]) ? 1.0 : 0.0;
    ge += (a.obs_sum >= ps[i]) ? 1.0 : 0.0;
  }
  const double tle = block_sum(le);
  const double tge = block_sum(ge);
  if (tid == 0) { summ[16] = tle; summ[17] = tge; }
}

// One CTA per column job (posterior_processing.py:25-60 called at lib.py:760, 886, 888 and
// the quantiles of lib.py:1021-1075).  Heavy float64 jobs first.
template <typename R>
__global__ void __launch_bounds__(1024)
k_impact_jobs(const R* __restrict__ trT, const double* __restrict__ cumT,
                              const double* __restrict__ statsT, const double* __restrict__ obs,
                              ImpactDev a, double* __restrict__ series, double* __restrict__ summ,
                              int in_smem, const ImpactSeries* __restrict__ per = nullptr) {
  if (per) {
    const size_t sidx = blockIdx.y;
    a.scale = per[sidx].scale; a.offset = per[sidx].offset; a.obs_sum = per[sidx].obs_sum;
    obs += sidx * (size_t)a.T;
    trT += sidx * (size_t)a.S * a.T; cumT += sidx * (size_t)a.S * (a.T - a.t_c0);
    statsT += sidx * (size_t)a.S * IMP_STATS;
    series += sidx * (size_t)a.T * IMP_SERIES_COLS; summ += sidx * (size_t)IMP_SUMMARY_LEN;
  }
  extern __shared__ __align__(16) unsigned char key_mem[];
  __shared__ __align__(16) unsigned char sel_raw[sizeof(SelectShared<double>)];
  __shared__ int ibuf[10];
  __shared__ double red[33];
  const int Tc = a.T - a.t_c0, S = a.S;
  const int b = blockIdx.x, tid = threadIdx.x;
  double res[2][4], g[2];
  if (b < Tc + IMP_STATS) {                         // cumulative-effect / per-draw statistic column
    const bool is_cum = b < Tc;
    const double* col = is_cum ? cumT + (size_t)b * S : statsT + (size_t)(b - Tc) * S;
    SelectShared<double>& sh = *reinterpret_cast<SelectShared<double>*>(sel_raw);
    const int n = column_quantiles<double, false>(col, S, a.q_lo, a.q_hi, key_mem, in_smem, sh, ibuf, res, g);
    if (tid == 0) {
      double* out = is_cum ? series + (size_t)(a.t_c0 + b) * IMP_SERIES_COLS + 7
                           : summ + 2 * (b - Tc);
      for (int iq = 0; iq < 2; ++iq)
        out[iq] = n ? imp_lerp(res[iq][0], res[iq][1], g[iq]) : CUDART_NAN;
    }
    return;
  }
  if (b < Tc + IMP_STATS + a.T) {                   // prediction + point-effect column
    const int t = b - Tc - IMP_STATS;
    SelectShared<R>& sh = *reinterpret_cast<SelectShared<R>*>(sel_raw);
    const int n = column_quantiles<R, true>(trT + (size_t)t * S, S, a.q_lo, a.q_hi, key_mem, in_smem,
                                            sh, ibuf, res, g);
    if (tid == 0) {
      double* row = series + (size_t)t * IMP_SERIES_COLS;
      if (t < a.t_c0) { row[7] = 0.0; row[8] = 0.0; }          // cumulative effect is 0 before post
      const double o = obs[t];
      for (int iq = 0; iq < 2; ++iq) {
        if (n == 0) { row[1 + iq] = CUDART_NAN; row[4 + iq] = CUDART_NAN; continue; }
        double v[4];
        for (int j = 0; j < 4; ++j) v[j] = imp_unscale(res[iq][j], a.scale, a.offset);
        row[1 + iq] = imp_lerp(v[0], v[1], g[iq]);
        row[4 + iq] = (o == o) ? imp_lerp(o - v[2], o - v[3], g[iq]) : CUDART_NAN;
      }
    }
    return;
  }
  impact_summary_block(statsT, a, summ, red);
}

}  // namespace ci
