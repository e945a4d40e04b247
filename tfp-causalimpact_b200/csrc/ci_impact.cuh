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
// Three kernels + two launches of K5 (ci_predict.cuh):
//   k_impact_rows   one warp per draw (+1 warp for the predictive mean): cumulative
//                   effect paths (warp scan over time), post-period per-draw statistics
//   k_impact_cols   one CTA per time step: prediction and point-effect quantiles
//   k_row_quantiles<double> on the cumulative paths and on the per-draw statistics
//   k_impact_summary  one CTA: sd (ddof=1), mean relative effect, tail counts
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

// standardize.py:60-64: (x * stddev) + mean, two roundings like numpy (no FMA contraction)
__device__ __forceinline__ double imp_unscale(double x, double scale, double offset) {
  return __dadd_rn(__dmul_rn(x, scale), offset);
}

constexpr int IMP_ROWS_PER_CTA = 8;

// Row r < S: draw r of traj; row S: the predictive mean (its cumulative path and
// post-period mean / sum are the *_mean series columns and `predicted`).
template <typename R>
__global__ void __launch_bounds__(32 * IMP_ROWS_PER_CTA)
k_impact_rows(const R* __restrict__ traj, const R* __restrict__ mean,
              const double* __restrict__ obs, const uint8_t* __restrict__ period, ImpactDev a,
              double* __restrict__ cum, double* __restrict__ stats, double* __restrict__ series,
              double* __restrict__ summ) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * IMP_ROWS_PER_CTA + (threadIdx.x >> 5);
  if (r > a.S) return;
  const bool is_mean = (r == a.S);
  const R* src = is_mean ? mean : traj + (size_t)r * a.T;
  const int Tc = a.T - a.t_c0;
  double carry = 0.0, pred_sum = 0.0, eff_sum = 0.0;
  int eff_cnt = 0;
  for (int base = is_mean ? 0 : a.t_c0; base < a.T; base += 32) {
    const int t = base + lane;
    const bool valid = t < a.T;
    double x = 0.0, pt = CUDART_NAN;
    int per = 0;
    if (valid) {
      x = imp_unscale((double)src[t], a.scale, a.offset);
      pt = obs[t] - x;                               // lib.py:822-823
      per = period[t];
    }
    const bool isn = !(pt == pt);
    // lib.py:826-831: effects before the post-period count as 0; NaNs are skipped, not spread
    double inc = (valid && t >= a.t_c0 && !isn) ? pt : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += up;
    }
    const double cv = carry + inc;
    carry += __shfl_sync(FULL, inc, 31);
    if (valid) {
      const double out = (t >= a.t_c0 && isn) ? CUDART_NAN : cv;
      if (is_mean) {
        double* row = series + (size_t)t * IMP_SERIES_COLS;
        row[0] = x; row[3] = pt; row[6] = out;
      } else {
        cum[(size_t)r * Tc + (t - a.t_c0)] = out;
      }
      if (per == 1) {                                // inside the post-period (lib.py:966-1011)
        pred_sum += x;
        if (!isn) { eff_sum += pt; ++eff_cnt; }
      }
    }
  }
  pred_sum = warp_sum(pred_sum);
  eff_sum = warp_sum(eff_sum);
  eff_cnt = __reduce_add_sync(FULL, eff_cnt);
  if (lane == 0) {
    const double pm = pred_sum / (double)a.n_post;
    if (is_mean) {
      summ[18] = pm; summ[19] = pred_sum;
    } else {
      double* st = stats + (size_t)r * IMP_STATS;
      st[0] = pm; st[1] = pred_sum;
      st[2] = eff_cnt > 0 ? eff_sum / (double)eff_cnt : CUDART_NAN;
      st[3] = eff_sum;
      st[4] = a.obs_sum / pred_sum - 1.0;            // lib.py:1010-1011
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

// One CTA per time step: prediction quantiles and point-effect quantiles from ONE read
// of the column (posterior_processing.py:25-60 called at lib.py:760 and :886).
template <typename R>
__global__ void k_impact_cols(const R* __restrict__ traj, const double* __restrict__ obs,
                              ImpactDev a, double* __restrict__ series) {
  using Key = typename KeyOf<R>::type;
  extern __shared__ __align__(16) unsigned char qsmem[];
  Key* keys = reinterpret_cast<Key*>(qsmem);
  __shared__ SelectShared<R> sh;
  __shared__ int n_valid, s_nr;
  __shared__ int slot[2][4];     // per quantile: lo, hi, mirrored lo, mirrored hi
  const int t = blockIdx.x, tid = threadIdx.x;
  const int S = a.S, T = a.T;
  const int n = load_column_keys<R>(traj, S, T, t, keys, &n_valid);
  double* row = series + (size_t)t * IMP_SERIES_COLS;
  if (t < a.t_c0 && tid == 0) { row[7] = 0.0; row[8] = 0.0; }   // cumulative effect is 0 before post
  if (n == 0) {
    if (tid == 0) { row[1] = row[2] = row[4] = row[5] = CUDART_NAN; }
    return;
  }
  if (tid == 0) {
    int nr = 0;
    for (int iq = 0; iq < 2; ++iq) {
      const double pos = (iq == 0 ? a.q_lo : a.q_hi) * (double)(n - 1);
      int lo = (int)floor(pos);
      lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
      const int hi = lo + 1 < n ? lo + 1 : n - 1;
      slot[iq][0] = add_rank(sh, nr, lo);
      slot[iq][1] = add_rank(sh, nr, hi);
      // k-th smallest of (o - x) = o - (k-th largest x)
      slot[iq][2] = add_rank(sh, nr, n - 1 - lo);
      slot[iq][3] = add_rank(sh, nr, n - 1 - hi);
    }
    s_nr = nr;
  }
  __syncthreads();
  radix_select_multi<R>(keys, S, s_nr, sh);
  if (tid < 2) {
    const int iq = tid;
    const double pos = (iq == 0 ? a.q_lo : a.q_hi) * (double)(n - 1);
    int lo = (int)floor(pos);
    lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
    const double g = pos - (double)lo;
    auto val = [&](int s_) { return imp_unscale((double)KeyOf<R>::dec(sh.out[slot[iq][s_]]), a.scale, a.offset); };
    row[1 + iq] = imp_lerp(val(0), val(1), g);
    const double o = obs[t];
    row[4 + iq] = (o == o) ? imp_lerp(o - val(2), o - val(3), g) : CUDART_NAN;
  }
}

// sd (ddof = 1), mean relative effect and the tail counts of the p-value (lib.py:1021-1090).
// One CTA, fixed reduction order: deterministic.
__global__ void __launch_bounds__(1024)
k_impact_summary(const double* __restrict__ stats, ImpactDev a, double* __restrict__ summ) {
  __shared__ double red[32];
  __shared__ double bc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = a.S;
  auto block_sum = [&](double v) -> double {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < 32; ++w) tot += red[w];
      bc = tot;
    }
    __syncthreads();
    return bc;
  };
  for (int j = 0; j < IMP_STATS; ++j) {
    double s = 0.0;
    for (int i = tid; i < S; i += 1024) s += stats[(size_t)i * IMP_STATS + j];
    const double m = block_sum(s) / (double)S;
    double ss = 0.0;
    for (int i = tid; i < S; i += 1024) {
      const double d = stats[(size_t)i * IMP_STATS + j] - m;
      ss += d * d;
    }
    const double tot = block_sum(ss);
    if (tid == 0) {
      summ[10 + j] = S > 1 ? sqrt(tot / (double)(S - 1)) : CUDART_NAN;
      if (j == IMP_STATS - 1) summ[15] = m;
    }
  }
  double le = 0.0, ge = 0.0;
  for (int i = tid; i < S; i += 1024) {
    const double ps = stats[(size_t)i * IMP_STATS + 1];
    le += (a.obs_sum <= ps) ? 1.0 : 0.0;
    ge += (a.obs_sum >= ps) ? 1.0 : 0.0;
  }
  const double tle = block_sum(le);
  const double tge = block_sum(ge);
  if (tid == 0) { summ[16] = tle; summ[17] = tge; }
}

}  // namespace ci
