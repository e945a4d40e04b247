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
//   k_impact_rows  32 draws per CTA, 4 per warp (+1 CTA for the predictive mean):
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
// Which column jobs a k_impact_jobs launch runs.  One GPU: everything.  Time-sharded:
// this rank holds prediction columns [t_begin, t_begin + t_count) and cumulative columns
// [c_begin, c_begin + c_count) of ALL draws (trT / cumT are indexed from the block's first column);
// only the rank with do_stats holds the per-draw statistics and writes the summary.
struct ImpactCols { int t_begin, t_count, c_begin, c_count, do_stats; };
// Draws held by each rank of a sharded fit (ci_impact_sharded_d).  One NVSwitch domain.
constexpr int IMP_MAX_RANKS = 16;
struct ShardCounts { int ws; int n[IMP_MAX_RANKS]; };
// Column blocks as the exchange leaves them: the source ranks' blocks one after the other, block
// r = [rows][n_r] (rowsT rows of paths, rowsC rows of cumulative paths preceded by headC rows of
// per-draw statistics).  ws == 0: plain [rows][S] arrays.
struct ColBlocks { int ws, rowsT, rowsC, headC; int n[IMP_MAX_RANKS]; };
// Where k_impact_rows stores its output in a sharded fit: straight into the exchange windows of
// the ranks that own the time blocks (peer memory over NVLink; T[g] / C[g] are rank g's windows
// mapped into this process), in the ColBlocks layout -- this rank's block starts me_off draws in.
// ws == 0: the local [T][S] / [T - t_c0][S] / [5][S] arrays.
struct PeerDest {
  int ws, me_off, n_me;
  int T_base, T_extra, C_base, C_extra;       // balanced split of the T (T - t_c0) columns
  void* T[IMP_MAX_RANKS];
  void* C[IMP_MAX_RANKS];
};
// owner of item x under the balanced split (base = n / ws, extra = n % ws): rank, first item, count
__device__ __forceinline__ void split_owner(int x, int base, int extra, int& g, int& start, int& cnt) {
  const int lim = (base + 1) * extra;
  if (x < lim) { g = x / (base + 1); start = g * (base + 1); cnt = base + 1; }
  else { g = extra + (x - lim) / base; start = lim + (g - extra) * base; cnt = base; }
}

// standardize.py:60-64: (x * stddev) + mean, two roundings like numpy (no FMA contraction)
__device__ __forceinline__ double imp_unscale(double x, double scale, double offset) {
  return __dadd_rn(__dmul_rn(x, scale), offset);
}

constexpr int IMP_TILE = 32;         // draws per CTA == draws per transposed 128-byte run
constexpr int IMP_WARPS = 8;         // warps per CTA
constexpr int IMP_LPW = IMP_TILE / IMP_WARPS;   // load phase: draws per warp (lane <-> step)
constexpr int IMP_CH = 2;            // 32-step sub-chunks per loop iteration
constexpr int IMP_CHUNK = IMP_CH * 32;
constexpr int IMP_SPW = IMP_CHUNK / IMP_WARPS;  // compute phase: steps per warp (lane <-> draw)
constexpr int IMP_SEG = 4 * IMP_CHUNK;          // pre-period steps per transpose-only CTA

// where element (time column, draw) of the transposed outputs goes: the local arrays or, in a
// sharded fit, the owner rank's window (ColBlocks layout) over NVLink
// (the owner of a column is cached in `own`: a warp stores 8 consecutive columns, which cross a
// block boundary at most once -- the two integer divisions of split_owner are paid per boundary,
// not per store)
struct OwnerCache { int g = 0, st0 = 0, cnt = -1; };
template <typename R>
__device__ __forceinline__ R* impact_dst_path(R* trT, const ImpactDev& a, const PeerDest& pd, int tc,
                                              OwnerCache& own) {
  if (pd.ws == 0) return trT + (size_t)tc * a.S;
  if (tc < own.st0 || tc >= own.st0 + own.cnt)
    split_owner(tc, pd.T_base, pd.T_extra, own.g, own.st0, own.cnt);
  return static_cast<R*>(pd.T[own.g]) + (size_t)own.cnt * pd.me_off + (size_t)(tc - own.st0) * pd.n_me;
}
__device__ __forceinline__ double* impact_dst_cum(double* cumT, const ImpactDev& a, const PeerDest& pd,
                                                  int c, OwnerCache& own) {
  if (pd.ws == 0) return cumT + (size_t)c * a.S;
  if (c < own.st0 || c >= own.st0 + own.cnt)
    split_owner(c, pd.C_base, pd.C_extra, own.g, own.st0, own.cnt);
  const int head = own.g == 0 ? IMP_STATS : 0;
  return static_cast<double*>(pd.C[own.g]) + (size_t)(own.cnt + head) * pd.me_off +
         (size_t)(head + c - own.st0) * pd.n_me;
}

// The predictive mean's row (the *_mean series columns and `predicted`): one CTA, warp w takes
// 32 steps of every 256-step round (lane <-> step, warp scan), the warps' totals are chained
// through shared memory.
template <typename R>
__device__ __forceinline__ void impact_mean_row(const R* __restrict__ mean,
                                                const double* __restrict__ obs,
                                                const uint8_t* __restrict__ period,
                                                const ImpactDev& a, double* __restrict__ series,
                                                double* __restrict__ summ) {
  __shared__ double mtot[2][IMP_WARPS];
  __shared__ double msum[IMP_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double carry = 0.0, pred_sum = 0.0;
  int buf = 0;
  for (int base = 0; base < a.T; base += 32 * IMP_WARPS, buf ^= 1) {
    const int t = base + warp * 32 + lane;
    const bool valid = t < a.T;
    const double x = imp_unscale(valid ? (double)mean[t] : 0.0, a.scale, a.offset);
    const double pt = valid ? obs[t] - x : CUDART_NAN;               // lib.py:822-823
    const bool isn = !(pt == pt);
    // lib.py:826-831: effects before the post-period count as 0; NaNs are skipped
    double inc = (valid && t >= a.t_c0 && !isn) ? pt : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += up;
    }
    if (lane == 31) mtot[buf][warp] = inc;
    if (valid && period[t] == 1) pred_sum += x;
    __syncthreads();
    double off = carry, all = 0.0;
#pragma unroll
    for (int w = 0; w < IMP_WARPS; ++w) {
      const double v = mtot[buf][w];
      if (w < warp) off += v;
      all += v;
    }
    carry += all;
    if (valid) {
      double* row = series + (size_t)t * IMP_SERIES_COLS;
      row[0] = x; row[3] = pt;
      row[6] = t < a.t_c0 ? 0.0 : (isn ? CUDART_NAN : off + inc);
    }
  }
  pred_sum = warp_sum(pred_sum);
  if (lane == 0) msum[warp] = pred_sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < IMP_WARPS; ++w) tot += msum[w];
    summ[18] = tot / (double)a.n_post; summ[19] = tot;
  }
}

// CTA (rb, seg) = blockIdx.x % row_ctas, blockIdx.x / row_ctas works on draws 32 rb .. 32 rb + 31:
//   seg = 0   the steps from the 64-aligned start of the post-period to T: cumulative effect
//             paths, post-period per-draw statistics, transposed copies of the raw draws and of
//             the cumulative paths;
//   seg >= 1  pre-period steps [(seg-1) seg_len, seg seg_len): nothing depends on them except
//             their own quantiles, so these CTAs only transpose (and there are many of them).
// Every 64-step chunk is LOADED with lane <-> step (warp w reads draws 4w .. 4w+3: coalesced 128-
// byte runs of the [S,T] array; the next chunk's loads are issued before this one is processed)
// into a shared-memory tile, and PROCESSED with lane <-> draw (warp w takes steps 8w .. 8w+7):
// the running sum over time is then a plain per-thread recurrence -- no shuffles; the 8 warps'
// partial sums are chained through shared memory -- and every store is a coalesced run of 32
// draws of one time step, straight from registers.  With `mean`, CTA rb = row_ctas - 1 of seg 0
// is extra and walks the predictive mean instead (impact_mean_row).
// (History: round 1 -- 1024 threads, lane <-> step, warp-shuffle scans, every CTA walking all T
// steps, ONE CTA per SM: 157 us at S=10000, T=2000 for 210 MB of traffic.  Run 21: 256 threads x
// 4 draws per warp, pre-period segments as separate CTAs: 123 us, but a 1250-draw shard of an
// 8-GPU fit still took 85-100 us -- 3400 dependent instructions per warp at 2 warps per
// scheduler.)
template <typename R>
__global__ void __launch_bounds__(32 * IMP_WARPS, 3)
k_impact_rows(const R* __restrict__ traj, const R* __restrict__ mean,
              const double* __restrict__ obs, const uint8_t* __restrict__ period, ImpactDev a,
              R* __restrict__ trT, double* __restrict__ cumT, double* __restrict__ statsT,
              double* __restrict__ series, double* __restrict__ summ,
              const ImpactSeries* __restrict__ per, int row_ctas, int seg_len, PeerDest pd) {
  if (per) {      // batched: this CTA row works on series blockIdx.y (obs is [N,T], period shared)
    const size_t sidx = blockIdx.y;
    a.scale = per[sidx].scale; a.offset = per[sidx].offset; a.obs_sum = per[sidx].obs_sum;
    traj += sidx * (size_t)a.S * a.T; mean += sidx * (size_t)a.T; obs += sidx * (size_t)a.T;
    trT += sidx * (size_t)a.S * a.T; cumT += sidx * (size_t)a.S * (a.T - a.t_c0);
    statsT += sidx * (size_t)a.S * IMP_STATS;
    series += sidx * (size_t)a.T * IMP_SERIES_COLS; summ += sidx * (size_t)IMP_SUMMARY_LEN;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rb = blockIdx.x % row_ctas, seg = blockIdx.x / row_ctas;
  const bool post = seg == 0;
  if (mean != nullptr && rb == row_ctas - 1) {       // the extra CTA: the predictive mean's row
    if (post) impact_mean_row<R>(mean, obs, period, a, series, summ);
    return;
  }
  __shared__ R tile[2][IMP_TILE][IMP_CHUNK + 1];     // [draw][step] of the chunk, double-buffered
  __shared__ double wtot[2][IMP_WARPS][IMP_TILE];    // per warp: sum of its steps' effects, per draw
  const int post_base = (a.t_c0 / IMP_CHUNK) * IMP_CHUNK;
  int t_begin, t_end;
  if (post) {
    t_begin = post_base; t_end = a.T;
  } else {
    t_begin = (seg - 1) * seg_len;
    t_end = min(t_begin + seg_len, post_base);
    if (t_begin >= t_end) return;
  }
  const int r0 = rb * IMP_TILE;
  // load phase: this warp's draws
  const R* src[IMP_LPW];
  bool lok[IMP_LPW];
#pragma unroll
  for (int k = 0; k < IMP_LPW; ++k) {
    const int r = r0 + warp * IMP_LPW + k;
    lok[k] = r < a.S;
    src[k] = traj + (size_t)(lok[k] ? r : 0) * a.T;
  }
  // compute phase: this lane's draw
  const int rr = r0 + lane;
  const bool dok = rr < a.S;
  double carry = 0.0, pred_sum = 0.0, eff_sum = 0.0;
  int eff_cnt = 0;
  R nxt[IMP_LPW][IMP_CH];
#pragma unroll
  for (int k = 0; k < IMP_LPW; ++k)
#pragma unroll
    for (int h = 0; h < IMP_CH; ++h) {
      const int t = t_begin + h * 32 + lane;
      nxt[k][h] = (lok[k] && t < t_end) ? src[k][t] : (R)0;
    }
  OwnerCache ownT, ownC;
  int buf = 0;
  for (int base = t_begin; base < t_end; base += IMP_CHUNK, buf ^= 1) {
#pragma unroll
    for (int k = 0; k < IMP_LPW; ++k)
#pragma unroll
      for (int h = 0; h < IMP_CH; ++h) {
        tile[buf][warp * IMP_LPW + k][h * 32 + lane] = nxt[k][h];
        const int t = base + IMP_CHUNK + h * 32 + lane;          // next chunk: loads in flight
        nxt[k][h] = (lok[k] && t < t_end) ? src[k][t] : (R)0;
      }
    __syncthreads();          // tile[buf] is whole (and everybody is done with tile[buf] of 2 chunks ago)
    const int j0 = warp * IMP_SPW;
    if (!post) {
#pragma unroll
      for (int jj = 0; jj < IMP_SPW; ++jj) {
        const int tc = base + j0 + jj;
        if (dok && tc < t_end) impact_dst_path<R>(trT, a, pd, tc, ownT)[rr] = tile[buf][lane][j0 + jj];
      }
      continue;
    }
    // post-period chunk: this thread walks draw rr over steps base + j0 .. + IMP_SPW - 1
    double run = 0.0, s[IMP_SPW];
    unsigned nanmask = 0;
#pragma unroll
    for (int jj = 0; jj < IMP_SPW; ++jj) {
      const int t = base + j0 + jj;
      const bool valid = dok && t < a.T;
      const R raw = tile[buf][lane][j0 + jj];
      const double x = imp_unscale((double)raw, a.scale, a.offset);
      const double o = t < a.T ? obs[t] : CUDART_NAN;                // (the same for the whole warp)
      const int pdv = t < a.T ? (int)period[t] : 0;
      const double pt = valid ? o - x : CUDART_NAN;                  // lib.py:822-823
      const bool isn = !(pt == pt);
      // lib.py:826-831: effects before the post-period count as 0; NaNs are skipped
      run += (valid && t >= a.t_c0 && !isn) ? pt : 0.0;
      s[jj] = run;
      nanmask |= (isn ? 1u : 0u) << jj;
      if (valid && pdv == 1) {                       // inside the post-period (lib.py:966-1011)
        pred_sum += x;
        if (!isn) { eff_sum += pt; ++eff_cnt; }
      }
      if (valid) impact_dst_path<R>(trT, a, pd, t, ownT)[rr] = raw;
    }
    wtot[buf][warp][lane] = run;
    __syncthreads();
    double off = carry, all = 0.0;
#pragma unroll
    for (int w = 0; w < IMP_WARPS; ++w) {
      const double v = wtot[buf][w][lane];
      if (w < warp) off += v;
      all += v;
    }
    carry += all;
#pragma unroll
    for (int jj = 0; jj < IMP_SPW; ++jj) {
      const int t = base + j0 + jj;
      if (dok && t < a.T && t >= a.t_c0)
        impact_dst_cum(cumT, a, pd, t - a.t_c0, ownC)[rr] = ((nanmask >> jj) & 1u) ? CUDART_NAN : off + s[jj];
    }
  }
  if (!post) return;
  // per-draw statistics: the warps' partial sums, in warp order
  __shared__ double sred[2][IMP_WARPS][IMP_TILE];
  __shared__ int scnt[IMP_WARPS][IMP_TILE];
  sred[0][warp][lane] = pred_sum; sred[1][warp][lane] = eff_sum; scnt[warp][lane] = eff_cnt;
  __syncthreads();
  if (warp == 0 && dok) {
    double ps = 0.0, es = 0.0;
    int ec = 0;
#pragma unroll
    for (int w = 0; w < IMP_WARPS; ++w) { ps += sred[0][w][lane]; es += sred[1][w][lane]; ec += scnt[w][lane]; }
    double* sd = statsT;
    if (pd.ws) {                     // the statistics ride in front of rank 0's cumulative block
      const int cnt0 = pd.C_base + (pd.C_extra > 0 ? 1 : 0);
      sd = static_cast<double*>(pd.C[0]) + (size_t)(cnt0 + IMP_STATS) * pd.me_off;
    }
    sd[0 * (size_t)a.S + rr] = ps / (double)a.n_post;
    sd[1 * (size_t)a.S + rr] = ps;
    sd[2 * (size_t)a.S + rr] = ec > 0 ? es / (double)ec : CUDART_NAN;
    sd[3 * (size_t)a.S + rr] = es;
    sd[4 * (size_t)a.S + rr] = a.obs_sum / ps - 1.0;                 // lib.py:1010-1011
  }
}

// Grid of k_impact_rows: (row CTAs incl. the mean's, number of segments incl. the post one).
inline int impact_seg_len() {      // CI_B200_IMP_SEG: tuning knob (steps, rounded to whole chunks)
  static const int v = [] {
    const char* e = getenv("CI_B200_IMP_SEG");
    const int n = e ? atoi(e) : 0;
    return n > 0 ? ((n + IMP_CHUNK - 1) / IMP_CHUNK) * IMP_CHUNK : IMP_SEG;
  }();
  return v;
}
inline void impact_rows_grid(int S, int T, int t_c0, bool with_mean, int* row_ctas, int* nseg) {
  *row_ctas = (S + IMP_TILE - 1) / IMP_TILE + (with_mean ? 1 : 0);
  const int post_base = (t_c0 / IMP_CHUNK) * IMP_CHUNK;
  *nseg = 1 + (post_base + impact_seg_len() - 1) / impact_seg_len();
}

// numpy.lib._function_base_impl._lerp in float64
__device__ __forceinline__ double imp_lerp(double va, double vb, double g) {
  const double diff = vb - va;
  double r = va + diff * g;
  if (g >= 0.5) r = vb - diff * (1.0 - g);
  if (g == 0.0) r = va;
  return r;
}

// Fold a thread's smallest / largest valid key into sh.kmin / sh.kmax (what binned_select needs first)
template <typename V>
__device__ __forceinline__ void fold_minmax(typename KeyOf<V>::type mn, typename KeyOf<V>::type mx,
                                            SelectShared<V>& sh) {
  using Key = typename KeyOf<V>::type;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const Key a = __shfl_xor_sync(FULL, mn, o), b = __shfl_xor_sync(FULL, mx, o);
    mn = a < mn ? a : mn; mx = b > mx ? b : mx;
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&sh.kmin, mn); atomicMax(&sh.kmax, mx); }
}

// Contiguous column -> shared-memory keys (+ their min / max in sh); returns the number of non-NaN values.
template <typename V>
__device__ __forceinline__ int load_contig_keys(const V* __restrict__ col, int S,
                                                typename KeyOf<V>::type* keys, int* n_valid,
                                                SelectShared<V>& sh) {
  using Key = typename KeyOf<V>::type;
  const int tid = threadIdx.x, nt = blockDim.x;
  const Key NANK = KeyOf<V>::nan_key();
  if (tid == 0) { *n_valid = 0; sh.kmin = NANK; sh.kmax = 0; }
  __syncthreads();
  int cnt = 0;
  Key mn = NANK, mx = 0;
  for (int i = tid; i < S; i += nt) {
    const V v = col[i];
    const bool ok = (v == v);
    const Key k = ok ? KeyOf<V>::enc(v) : NANK;
    keys[i] = k;
    cnt += ok ? 1 : 0;
    if (ok) { mn = k < mn ? k : mn; mx = k > mx ? k : mx; }
  }
  fold_minmax<V>(mn, mx, sh);
  cnt = __reduce_add_sync(FULL, cnt);
  if ((tid & 31) == 0 && cnt) atomicAdd(n_valid, cnt);
  __syncthreads();
  return *n_valid;
}

// Row `row` of column blocks (ColBlocks layout, `rows` rows per block) -> shared-memory keys.
template <typename V>
__device__ __forceinline__ int load_block_keys(const V* __restrict__ base, int rows, int row,
                                               const ColBlocks& cb,
                                               typename KeyOf<V>::type* keys, int* n_valid,
                                               SelectShared<V>& sh) {
  using Key = typename KeyOf<V>::type;
  const int tid = threadIdx.x, nt = blockDim.x;
  const Key NANK = KeyOf<V>::nan_key();
  if (tid == 0) { *n_valid = 0; sh.kmin = NANK; sh.kmax = 0; }
  __syncthreads();
  int cnt = 0;
  Key mn = NANK, mx = 0;
  // one flat sweep over the S draws, four loads in flight per thread; (seg, off, blk) = the
  // block that holds draw i, advanced as i grows (a thread's draws are nt apart: rarely)
  int S = 0;
  for (int r = 0; r < cb.ws; ++r) S += cb.n[r];
  int seg = 0, off = 0;
  size_t blk = 0;
  for (int i0 = tid; i0 < S; i0 += 4 * nt) {
    V v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nt;
      if (i < S) {
        while (i >= off + cb.n[seg]) { off += cb.n[seg]; blk += (size_t)rows * cb.n[seg]; ++seg; }
        v[u] = base[blk + (size_t)row * cb.n[seg] + (i - off)];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nt;
      if (i < S) {
        const bool ok = (v[u] == v[u]);
        const Key k = ok ? KeyOf<V>::enc(v[u]) : NANK;
        keys[i] = k;
        cnt += ok ? 1 : 0;
        if (ok) { mn = k < mn ? k : mn; mx = k > mx ? k : mx; }
      }
    }
  }
  fold_minmax<V>(mn, mx, sh);
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
                                                double (&g)[2], const ColBlocks* cb = nullptr,
                                                int blk_rows = 0, int blk_row = 0) {
  using Key = typename KeyOf<V>::type;
  Key* keys = reinterpret_cast<Key*>(key_mem);
  int* n_valid = ibuf;                // [0]; [1] = number of ranks; [2..9] = slots
  const GlobalKeys<V> gkeys{col, (size_t)1};
  // column blocks (cb): `col` is the base of the blocks; always staged in shared memory
  const bool staged = cb != nullptr || in_smem;
  const int n = cb ? load_block_keys<V>(col, blk_rows, blk_row, *cb, keys, n_valid, sh)
                   : in_smem ? load_contig_keys<V>(col, S, keys, n_valid, sh)
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
  if (staged) radix_select_multi<V>(SmemKeys<V>{keys}, S, ibuf[1], sh, true);
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
                              int in_smem, const ImpactSeries* __restrict__ per, ImpactCols jc,
                              ColBlocks cb) {
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
  const int S = a.S;
  const int ns = jc.do_stats ? IMP_STATS : 0;
  const int b = blockIdx.x, tid = threadIdx.x;
  double res[2][4], g[2];
  if (b < jc.c_count + ns) {                        // cumulative-effect / per-draw statistic column
    const bool is_cum = b < jc.c_count;
    // sharded: the cumulative paths are read as the exchange left them (column blocks); the
    // per-draw statistics were laid side by side (they also feed impact_summary_block)
    const bool blocks = is_cum && cb.ws > 0;
    const double* col = blocks ? cumT : is_cum ? cumT + (size_t)b * S : statsT + (size_t)(b - jc.c_count) * S;
    SelectShared<double>& sh = *reinterpret_cast<SelectShared<double>*>(sel_raw);
    const int n = column_quantiles<double, false>(col, S, a.q_lo, a.q_hi, key_mem, in_smem, sh, ibuf, res, g,
                                                  blocks ? &cb : nullptr, cb.rowsC, cb.headC + b);
    if (tid == 0) {
      double* out = is_cum ? series + (size_t)(a.t_c0 + jc.c_begin + b) * IMP_SERIES_COLS + 7
                           : summ + 2 * (b - jc.c_count);
      for (int iq = 0; iq < 2; ++iq)
        out[iq] = n ? imp_lerp(res[iq][0], res[iq][1], g[iq]) : CUDART_NAN;
    }
    return;
  }
  if (b < jc.c_count + ns + jc.t_count) {           // prediction + point-effect column
    const int tl = b - jc.c_count - ns, t = jc.t_begin + tl;
    SelectShared<R>& sh = *reinterpret_cast<SelectShared<R>*>(sel_raw);
    const int n = column_quantiles<R, true>(cb.ws > 0 ? trT : trT + (size_t)tl * S, S, a.q_lo, a.q_hi,
                                            key_mem, in_smem, sh, ibuf, res, g,
                                            cb.ws > 0 ? &cb : nullptr, cb.rowsT, tl);
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

// ---- pieces of the time-sharded impact stage (ci_impact_sharded_d) ----------------
// What the exchange leaves: the blocks of the source ranks one after the other, block r =
// [src_rows][n_r].  Lay the first gridDim.y rows side by side: out [rows][S], draws in rank order.
template <typename V>
__global__ void __launch_bounds__(256)
k_merge_blocks(const V* __restrict__ recv, V* __restrict__ out, int src_rows, ShardCounts sc, int S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
  if (i >= S) return;
  int r = 0, off = 0;
  size_t base = 0;
  while (i >= off + sc.n[r]) { off += sc.n[r]; base += (size_t)src_rows * sc.n[r]; ++r; }
  out[(size_t)row * S + i] = recv[base + (size_t)row * sc.n[r] + (i - off)];
}

// Predictive mean over all draws from the ranks' means of their own draws: draw-count-weighted
// float64 sum in rank order (the same on every rank).
template <typename R>
__global__ void __launch_bounds__(256)
k_mean_combine(const R* __restrict__ parts, ShardCounts sc, int S, int T, R* __restrict__ mean) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  double acc = 0.0;
  for (int r = 0; r < sc.ws; ++r)
    if (sc.n[r]) acc += (double)parts[(size_t)r * T + t] * ((double)sc.n[r] / (double)S);
  mean[t] = (R)acc;
}

}  // namespace ci
