"""Panels of independent series without per-series pandas.

``fit_causalimpact_panel`` takes ONE array ``values [N, T, 1 + k]`` (column 0 = outcome) with a
shared index and shared periods -- the "thousands of geographies" shape -- and returns arrays.
Everything the reference does per series in pandas (data.py:77-137: split, nan-aware
standardisation with the pre-period statistics, intercept column, masked outcome; then
causalimpact_lib.py:892-931 / :1021-1091: series and summary frames) is done once for the whole
panel ON THE DEVICE by one kernel (``ci_set_panel``: standardisation, intercept, masks, tiles, Gram
matrices, priors); sampling is one batched launch (``ci_gibbs_run_batch_d``), predictive mean and
impact are one batched launch each (grid.y = series) and everything is read back with a single
copy.  Frames are built only on demand (``PanelResult.frames(i)``).

The arithmetic per series is the one of ``fit_causalimpact``; only the summation order inside the
pre-period mean / sd can differ from pandas' (1 ulp in float64 before the cast to the engine
dtype), so results agree with ``fit_causalimpact_many`` to rounding, not bit for bit
(tests/test_gpu_batch.py).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional

import numpy as np
import pandas as pd

from . import frame as _frame
from . import impact as _impact
from . import shard as _shard
from .model import build_problem, build_seasonal

SUMMARY_COLUMNS = ["actual", "predicted", "predicted_lower", "predicted_upper", "predicted_sd",
                   "abs_effect", "abs_effect_lower", "abs_effect_upper", "abs_effect_sd",
                   "rel_effect", "rel_effect_lower", "rel_effect_upper", "rel_effect_sd", "p_value",
                   "alpha"]


@dataclasses.dataclass
class PanelResult:
  """Arrays for the series this rank owns (``series_ids``); frames on demand."""
  index: pd.Index                 # the caller's index
  series_ids: np.ndarray          # [n] positions in the input panel
  series: np.ndarray              # [n, len(index), 10] float64, impact.SERIES_VALUE_COLUMNS
  summary: np.ndarray             # [n, 2, 15] float64: (average, cumulative) x SUMMARY_COLUMNS
  observation_noise_scale: np.ndarray   # [n, S]
  level_scale: np.ndarray         # [n, S]
  weights: Optional[np.ndarray]   # [n, S, k + 1] or None
  level: Optional[np.ndarray]     # [n, S, T_model] when keep_level
  seasonal_levels: Optional[np.ndarray]         # [n, S, T_model, K] when keep_level and seasons
  seasonal_drift_scales: Optional[np.ndarray]   # [n, S, K] with seasons
  pre_period: tuple
  post_period: tuple
  inclusion: np.ndarray           # [n, chains, k + 1]

  def frames(self, i: int):
    """(series, summary) DataFrames of the i-th owned series, as fit_causalimpact returns."""
    ser = pd.DataFrame(self.series[i], index=self.index, columns=_impact.SERIES_VALUE_COLUMNS)
    ser["pre_period_start"] = self.pre_period[0]
    ser["pre_period_end"] = self.pre_period[1]
    ser["post_period_start"] = self.post_period[0]
    ser["post_period_end"] = self.post_period[1]
    summ = pd.DataFrame(self.summary[i], index=["average", "cumulative"], columns=SUMMARY_COLUMNS)
    return ser, summ


def panel_layout(index, pre_period, post_period):
  """The O(T) row layout every series of a panel shares: validated periods, the modelled span
  (rows of the pre-period and everything after it: contiguous, ``row0`` .. end) and the number of
  pre-period rows."""
  index = pd.Index(index)
  probe = pd.DataFrame({"y": np.zeros(len(index))}, index=index)
  pre, post = _frame.parse_and_validate_date_data(probe, pre_period, post_period)
  in_pre = np.asarray((index >= pre[0]) & (index <= pre[1]))
  after = np.asarray(index > pre[1])
  rows = np.flatnonzero(in_pre | after)
  if rows.size == 0 or not np.array_equal(rows, np.arange(rows[0], len(index))):
    raise ValueError("the panel's index must be sorted: the modelled span has to be one block of rows")
  return dict(index=index, pre=pre, post=post, rows=rows, row0=int(rows[0]), n_pre=int(in_pre.sum()))


def prepare_panel(values, index, pre_period, post_period, standardize_data=True, dtype=np.float32):
  """Vectorised data.py:77-137 for N series at once -- the HOST restatement of what
  ``ci_set_panel`` (csrc/abi_panel.cu) does on the device; the product path does not call it (the
  CPU test double, tests/fake_engine.py, and the parity tests do).

  Returns dict: y_ext [N, Tm] (float64 view of the dtype-rounded standardized outcome, NaN where
  masked), design [N, Tm, k+1] or None, outcome_sd [N], y_scale / y_offset [N], model rows
  (positions of the modelled span in ``index``), periods, observed-scale outcome [N, Tm]."""
  values = np.asarray(values, dtype=np.float64)
  if values.ndim != 3 or values.shape[2] < 1:
    raise ValueError("values must be [n_series, T, 1 + n_covariates]")
  index = pd.Index(index)
  N, T, ncol = values.shape
  if len(index) != T:
    raise ValueError("index length differs from values.shape[1]")
  probe = pd.DataFrame({"y": np.zeros(T)}, index=index)
  pre, post = _frame.parse_and_validate_date_data(probe, pre_period, post_period)
  in_pre = np.asarray((index >= pre[0]) & (index <= pre[1]))
  after = np.asarray(index > pre[1])
  rows = np.flatnonzero(in_pre | after)                 # the modelled span, in index order
  n_pre = int(in_pre.sum())
  y_all = values[:, :, 0]
  # input validation of data.py:140-190, for every series
  with np.errstate(invalid="ignore"):
    if np.any(np.nanstd(y_all, axis=1) == 0):
      raise ValueError("Input response cannot be constant.")
  if np.any(np.sum(~np.isnan(y_all), axis=1) < 3):
    raise ValueError("Input data must have at least 3 observations.")
  if ncol > 1 and np.isnan(values[:, :, 1:]).any():
    raise ValueError("Input data cannot have any missing values.")
  pre_vals = values[:, in_pre, :]                       # [N, n_pre, ncol]
  model_vals = values[:, rows, :]                       # [N, Tm, ncol]
  if standardize_data:
    mean = np.nanmean(pre_vals, axis=1)                 # standardize.py:42-47 (ddof = 1)
    std = np.nanstd(pre_vals, axis=1, ddof=1)
    ok = std > 0
    scaled = np.where(ok[:, None, :], (model_vals - mean[:, None, :]) / np.where(ok, std, 1.0)[:, None, :],
                      model_vals)
    y_scale, y_offset = std[:, 0].copy(), mean[:, 0].copy()
  else:
    scaled = model_vals
    y_scale, y_offset = np.ones(N), np.zeros(N)
  np_dt = np.dtype(dtype)
  y_pre = scaled[:, :n_pre, 0].astype(np_dt)
  y_ext = np.concatenate([y_pre, np.full((N, len(rows) - n_pre), np.nan, dtype=np_dt)], axis=1)
  design = None
  if ncol > 1:
    design = np.concatenate([scaled[:, :, 1:], np.ones((N, len(rows), 1))], axis=2).astype(np_dt)
  outcome_sd = np.nanstd(y_pre, axis=1, ddof=1).astype(np_dt).astype(np.float64)   # lib.py:563-564
  return dict(y_ext=y_ext.astype(np.float64), design=None if design is None else design.astype(np.float64),
              outcome_sd=outcome_sd, y_scale=y_scale, y_offset=y_offset, rows=rows, n_pre=n_pre,
              pre=pre, post=post, y_model=model_vals[:, :, 0], index=index)


def _empty_result(values, index, pre_period, post_period, np_dt, seasons, keep_level) -> PanelResult:
  """The result of a rank that owns no series (more ranks than series)."""
  index = pd.Index(index)
  probe = pd.DataFrame({"y": np.zeros(len(index))}, index=index)
  pre, post = _frame.parse_and_validate_date_data(probe, pre_period, post_period)
  k1 = values.shape[2]                       # covariates + intercept
  z = lambda *shape: np.zeros(shape, np_dt)
  return PanelResult(
      index=index, series_ids=np.zeros(0, np.int64), series=np.zeros((0, len(index), 10)),
      summary=np.zeros((0, 2, 15)), observation_noise_scale=z(0, 0), level_scale=z(0, 0),
      weights=z(0, 0, k1) if k1 > 1 else None, level=z(0, 0, 0) if keep_level else None,
      seasonal_levels=z(0, 0, 0, 0) if (seasons and keep_level) else None,
      seasonal_drift_scales=z(0, 0, 0) if seasons else None, pre_period=pre, post_period=post,
      inclusion=np.zeros((0, 0, k1 if k1 > 1 else 0), np.float32))


def fit_causalimpact_panel(values, index, pre_period, post_period, alpha: float = 0.05, seed=None,
                           data_options=None, model_options=None, inference_options=None,
                           engine_options=None, keep_level: bool = False) -> PanelResult:
  """Fit every series of ``values [N, T, 1 + k]`` (shared ``index`` and periods); see module
  docstring.  Series are sharded over the ranks of an initialised process group (contiguous
  ranges, no collective); the result holds this rank's series (``series_ids``)."""
  from . import api as _api
  data_options = data_options if data_options is not None else _api.DataOptions()
  model_options = model_options if model_options is not None else _api.ModelOptions()
  inference_options = inference_options if inference_options is not None else _api.InferenceOptions()
  opts = engine_options or _api.EngineOptions()
  seasons = list(model_options.seasons or ())
  if not 0 < alpha < 1:
    raise ValueError("`alpha` must be between 0 and 1.")
  np_dt = _api._np_dtype(data_options.dtype)
  seed64 = _api._seed_to_u64(seed)
  values = np.asarray(values)
  rank, ws = _shard.world()
  s0, n_local = _shard.split_range(values.shape[0], ws, rank)
  eng = _api._resolve_engine(opts)
  if seed is None and ws > 1:          # fresh entropy: every rank must use rank 0's
    seed64 = _shard.broadcast_u64(seed64, getattr(eng, "torch_device", lambda: None)())
  if n_local == 0:                     # more ranks than series: nothing to fit on this one
    return _empty_result(values, index, pre_period, post_period, np_dt, bool(seasons), keep_level)
  # ---- O(T) layout shared by every series (host) ----
  lay = panel_layout(index, pre_period, post_period)
  vals_local = np.ascontiguousarray(values[s0:s0 + n_local], dtype=np.float64)
  if vals_local.ndim != 3 or vals_local.shape[2] < 1:
    raise ValueError("values must be [n_series, T, 1 + n_covariates]")
  if vals_local.shape[1] != len(lay["index"]):
    raise ValueError("index length differs from values.shape[1]")
  # ---- data.py:77-137 + priors + tiles for the whole panel: ONE kernel (ci_set_panel) ----
  stats = eng.set_panel(vals_local, row0=lay["row0"], n_pre=lay["n_pre"],
                        standardize=data_options.standardize_data, dtype=np_dt,
                        prior_level_sd=model_options.prior_level_sd,
                        ub_on_scale=opts.upper_bound_on == "scale")
  y_scale, y_offset, outcome_sd = stats[:, 0], stats[:, 1], stats[:, 2]
  N, Tm = vals_local.shape[0], len(lay["rows"])
  p = vals_local.shape[2] if vals_local.shape[2] > 1 else 0
  S = inference_options.num_results
  C = max(int(opts.num_chains), 1)
  n_per = max(1, math.ceil(S / C))
  n_warm = max(int(inference_options.num_warmup_steps), int(opts.gibbs_min_warmup))
  seas = drift = None
  stride = C if opts.decorrelate_series else 0          # see EngineOptions.decorrelate_series
  bkw = dict(n_warmup=n_warm, n_results=n_per, seed=seed64, chain_id0=s0 * stride, sparse=True,
             ssvs_order=opts.ssvs_order, series_stride=stride)
  if seasons:
    # one calendar for the panel; the priors scale with every series' own outcome sd (lib.py:472-489)
    sched = build_seasonal(seasons, Tm, 1.0)
    eng.set_seasonal_batch(sched, init_sd=outcome_sd, drift_scale=5e-7 * outcome_sd ** 2,
                           drift_ub=outcome_sd)
    theta, level, latent, traj, seas, drift, incl = eng.gibbs_seasonal_run_batch_t(C, **bkw)
  else:
    theta, level, traj, incl = eng.gibbs_run_batch_t(C, **bkw)
    latent = level
  S = min(S, C * n_per)

  # ---- O(T) metadata of the impact stage, shared / vectorised (impact.prepare per series) ----
  mi = lay["index"][lay["rows"]]
  pre, post = lay["pre"], lay["post"]
  in_pre = np.asarray(mi <= pre[1])
  in_post = np.asarray((mi >= post[0]) & (mi <= post[1]))
  period = np.where(np.asarray(mi < post[0]), 0, np.where(in_post, 1, 2)).astype(np.uint8)
  y_model = vals_local[:, lay["row0"]:, 0]
  observed = np.where((in_pre | in_post)[None, :], y_model, np.nan)              # [N, Tm]
  hide = np.asarray(((mi > pre[1]) & (mi < post[0])) | (mi > post[1]))[None, :] | np.isnan(observed)
  y_post = observed[:, in_post]
  obs_mean, obs_sum = np.nanmean(y_post, axis=1), np.nansum(y_post, axis=1)
  q_lo, q_hi = _impact._percentile_q(alpha / 2.0), _impact._percentile_q(1.0 - alpha / 2.0)

  # ---- predictive mean + impact of EVERY series: three launches, one read-back ----
  mean_all = eng.predictive_mean_batch_t(theta[:, :S], latent[:, :S])
  ser_d, sum_d = eng.impact_batch_t(traj[:, :S], mean_all, scale=y_scale, offset=y_offset,
                                    obs_sum=obs_sum, observed=observed, period=period,
                                    q_lo=q_lo, q_hi=q_hi)
  series9 = eng.to_host(ser_d).reshape(N, Tm, 9)
  summ = eng.to_host(sum_d)

  # ---- lib.py:892-931 for all series: NaN rules, re-index to the caller's index ----
  cols = np.concatenate([observed[:, :, None], series9], axis=2)     # [N, Tm, 10]
  cols[:, :, 4:][hide] = np.nan
  full = np.full((N, len(lay["index"]), 10), np.nan)
  full[:, lay["rows"], :] = cols
  full[:, :, 0] = values[s0:s0 + n_local, :, 0]                      # `observed` = the input column
  # ---- lib.py:1021-1091 for all series ----
  qd = summ[:, 0:10].reshape(N, 5, 2); sd = summ[:, 10:15]
  avg_pred, cum_pred, rel_mean = summ[:, 18], summ[:, 19], summ[:, 15]
  table = np.empty((N, 2, 15))
  for r, (act, pred, k_pred, k_eff) in enumerate(((obs_mean, avg_pred, 0, 2), (obs_sum, cum_pred, 1, 3))):
    table[:, r, 0] = act; table[:, r, 1] = pred
    table[:, r, 2] = qd[:, k_pred, 0]; table[:, r, 3] = qd[:, k_pred, 1]; table[:, r, 4] = sd[:, k_pred]
    table[:, r, 5] = act - pred
    table[:, r, 6] = qd[:, k_eff, 0]; table[:, r, 7] = qd[:, k_eff, 1]; table[:, r, 8] = sd[:, k_eff]
    table[:, r, 9] = rel_mean
    table[:, r, 10] = qd[:, 4, 0]; table[:, r, 11] = qd[:, 4, 1]; table[:, r, 12] = sd[:, 4]
    table[:, r, 13] = np.minimum((summ[:, 16] + 1.0) / (S + 1.0), (summ[:, 17] + 1.0) / (S + 1.0))
    table[:, r, 14] = alpha
  th = eng.to_host(theta[:, :S]).astype(np.float64)
  return PanelResult(
      index=lay["index"], series_ids=np.arange(s0, s0 + n_local), series=full, summary=table,
      observation_noise_scale=np.exp(0.5 * th[:, :, p]).astype(np_dt),
      level_scale=np.exp(0.5 * th[:, :, p + 1]).astype(np_dt),
      weights=th[:, :, :p].astype(np_dt) if p else None,
      level=eng.to_host(level[:, :S]).astype(np_dt, copy=False) if keep_level else None,
      seasonal_levels=(eng.to_host(seas[:, :S]).astype(np_dt, copy=False)
                       if (seasons and keep_level) else None),
      seasonal_drift_scales=(np.exp(0.5 * eng.to_host(drift[:, :S]).astype(np.float64)).astype(np_dt)
                             if seasons else None),
      pre_period=pre, post_period=post, inclusion=incl)
