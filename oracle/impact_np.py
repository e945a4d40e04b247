"""ORACLE (test infrastructure only): numpy restatement of the O(S*T) arithmetic of the
reference's _compute_impact -- what the device kernels of ci_impact (csrc/ci_impact.cuh)
are checked against.  Pinned to the reference itself: tests/golden/postproc_*.npz hold
outputs of the reference's own functions (oracle/make_golden_postproc.py), and
tests/test_postproc_golden.py runs this file through the product's packaging code
against them.

Follows (relative to /root/reference):
  causalimpact/causalimpact_lib.py:793-837   point / cumulative effect paths
  causalimpact/causalimpact_lib.py:840-931   per-time quantiles of the three families
  causalimpact/causalimpact_lib.py:934-1093  post-period summary statistics
  causalimpact/posterior_processing.py:25-60 quantiles (pandas, linear interpolation)
  causalimpact/standardize.py:60-64          inverse scaling
"""
import numpy as np

from oracle import quantiles_np

SERIES_COLS = 9
SUMMARY_LEN = 20


def _nan_cumsum(a, axis):
  """pandas cumsum(skipna=True): NaNs stay NaN but do not poison later rows."""
  nan = np.isnan(a)
  out = np.cumsum(np.where(nan, 0.0, a), axis=axis)
  out[nan] = np.nan
  return out


def impact_arrays(traj, mean, observed, period, scale, offset, q_lo, q_hi, obs_sum,
                  row_quantiles=quantiles_np.row_quantiles):
  """Same contract as ci_impact (include/ci_b200.h): returns (series [T,9], summary [20])."""
  traj = np.asarray(traj, dtype=np.float64) * scale + offset          # [S, T]
  mean = np.asarray(mean, dtype=np.float64).reshape(-1) * scale + offset
  observed = np.asarray(observed, dtype=np.float64)
  period = np.asarray(period)
  S, T = traj.shape
  q = np.array([q_lo, q_hi])
  before_post, in_post = period == 0, period == 1

  point = observed[None, :] - traj
  cum = _nan_cumsum(np.where(before_post[None, :], 0.0, point), axis=1)
  q_pred = row_quantiles(traj, q)
  q_point = row_quantiles(point, q)
  q_cum = row_quantiles(cum, q)
  point_mean = observed - mean
  cum_mean = _nan_cumsum(np.where(before_post, 0.0, point_mean), axis=0)
  series = np.column_stack([mean, q_pred[:, 0], q_pred[:, 1], point_mean, q_point[:, 0],
                            q_point[:, 1], cum_mean, q_cum[:, 0], q_cum[:, 1]])

  traj_post, point_post = traj[:, in_post], point[:, in_post]
  pred_mean_s, pred_sum_s = traj_post.mean(axis=1), traj_post.sum(axis=1)
  with np.errstate(invalid="ignore"), __import__("warnings").catch_warnings():
    __import__("warnings").simplefilter("ignore")
    eff_mean_s = np.nanmean(point_post, axis=1)
  eff_sum_s = np.nansum(point_post, axis=1)
  rel_s = obs_sum / pred_sum_s - 1.0
  per_draw = np.column_stack([pred_mean_s, pred_sum_s, eff_mean_s, eff_sum_s, rel_s])
  qd = row_quantiles(per_draw, q)                                      # [5, 2]
  sd = per_draw.std(axis=0, ddof=1) if S > 1 else np.full(5, np.nan)
  summ = np.empty(SUMMARY_LEN)
  summ[0:10] = qd.reshape(-1)
  summ[10:15] = sd
  summ[15] = rel_s.mean()
  summ[16] = np.sum(obs_sum <= pred_sum_s)
  summ[17] = np.sum(obs_sum >= pred_sum_s)
  summ[18], summ[19] = mean[in_post].mean(), mean[in_post].sum()
  return series, summ
