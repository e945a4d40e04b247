"""Input frame handling: period parsing, validation, standardisation.

Host-side (pandas/numpy, O(T p) once per fit) -- not part of the CUDA hot path,
but it sits directly in front of it, so it keeps the reference's semantics and
error behaviour.  No TensorFlow objects: ``outcome_ts`` is a plain
``MaskedSeries`` (numpy) instead of ``tfp.sts.MaskedTimeSeries``.

Mirrors (reference, relative to /root/reference):
  causalimpact/indices.py:30-149      period parsing + alignment + checks
  causalimpact/standardize.py:26-64   nan-aware standardiser (ddof = 1)
  causalimpact/data.py:77-190         CausalImpactData
"""
from __future__ import annotations

import dataclasses
import datetime
from typing import Optional, Tuple, Union

import numpy as np
import pandas as pd

InputDateType = Union[str, int, datetime.datetime]


# ---------------------------------------------------------------------------
# periods (indices.py)
# ---------------------------------------------------------------------------
def _to_index_value(value, frame: pd.DataFrame):
  """str -> Timestamp, int -> positional lookup, datetime -> itself (indices.py:137-149)."""
  if isinstance(value, str):
    return pd.to_datetime(value)
  if isinstance(value, (int, np.integer)):
    return frame.index[value]
  if isinstance(value, datetime.datetime):
    return value
  raise ValueError(f"Expected argument to be str, int, or datetime. Got {type(value)}")


def align_period(period, frame: pd.DataFrame):
  """Snap a (start, end) pair onto the index, shrinking rather than growing
  (start -> next index value, end -> previous one; indices.py:100-134)."""
  start, end = period
  if start > end:
    raise ValueError(f"Period end must be after period start. Got {period}")
  pos = frame.index.get_indexer([start], method="bfill")[0]
  if pos == -1:
    raise ValueError("Aligned period start not found in the index.")
  snapped_start = frame.index[pos]
  pos = frame.index.get_indexer([end], method="ffill")[0]
  if pos == -1:
    raise ValueError("Aligned period end not found in the index.")
  return snapped_start, frame.index[pos]


def validate_periods(pre_period, post_period, frame: pd.DataFrame):
  """indices.py:57-97."""
  pre = align_period(pre_period, frame)
  post = align_period(post_period, frame)
  n_pre = int(((frame.index >= pre[0]) & (frame.index <= pre[1])).sum())
  if pre[1] >= post[0]:
    raise ValueError("pre_period and post_period cannot overlap.")
  if n_pre < 3:
    raise ValueError("pre_period must span at least 3 time points. Got %s" % n_pre)
  if pre[1] < pre[0]:
    raise ValueError("pre_period last number must be bigger than its first.")
  if post[1] < post[0]:
    raise ValueError("post_period last number must be bigger than its first.")
  return pre, post


def parse_and_validate_date_data(data: pd.DataFrame, pre_period, post_period):
  """Entry point with the reference's name and signature (indices.py:30-54)."""
  pre = tuple(_to_index_value(v, data) for v in pre_period)
  post = tuple(_to_index_value(v, data) for v in post_period)
  return validate_periods(pre, post, data)


# ---------------------------------------------------------------------------
# standardiser (standardize.py)
# ---------------------------------------------------------------------------
class NotFittedError(ValueError, AttributeError):
  """Scaler used before fit()."""


class Scaler:
  """Column-wise (x - mean) / std with NaNs ignored in fit and kept in transform;
  unbiased std by default; zero-variance columns pass through unchanged."""

  def __init__(self, ddof: int = 1):
    self.ddof = ddof
    self._ready = False

  def fit(self, frame) -> "Scaler":
    self.mean_ = np.nanmean(frame, axis=0)
    self.stddev_ = np.nanstd(frame, axis=0, ddof=self.ddof)
    self._ready = True
    return self

  def _need_fit(self):
    if not self._ready:
      raise NotFittedError("Must call `.fit(df)` before using Scaler to transform!")

  def transform(self, frame: pd.DataFrame) -> pd.DataFrame:
    self._need_fit()
    # plain ndarray arithmetic: the same IEEE operations as the frame expression
    # (frame - mean) / std, without building three intermediate DataFrames
    vals = np.asarray(frame, dtype=np.float64) if not isinstance(frame, np.ndarray) else frame
    with np.errstate(divide="ignore", invalid="ignore"):       # zero-variance columns pass through
      scaled = np.where(self.stddev_ > 0, (vals - self.mean_) / self.stddev_, vals)
    return pd.DataFrame(scaled, index=frame.index, columns=frame.columns)

  def fit_transform(self, frame: pd.DataFrame) -> pd.DataFrame:
    return self.fit(frame).transform(frame)

  def inverse_transform(self, values):
    self._need_fit()
    return values * self.stddev_ + self.mean_


# ---------------------------------------------------------------------------
# the prepared frame (data.py)
# ---------------------------------------------------------------------------
@dataclasses.dataclass
class MaskedSeries:
  """numpy stand-in for tfp.sts.MaskedTimeSeries (data.py:125-128)."""
  time_series: np.ndarray
  is_missing: np.ndarray


def _select_columns(data: pd.DataFrame, outcome_column: Optional[str]):
  """Column defaults and input validation (data.py:140-190)."""
  if outcome_column is None:
    outcome_column = data.columns[0]
  if outcome_column not in data.columns:
    raise KeyError(f"Specified `outcome_column` ({outcome_column}) not found in data")
  if data[outcome_column].std(skipna=True, ddof=0) == 0:
    raise ValueError("Input response cannot be constant.")
  features = [c for c in data.columns if c != outcome_column] if data.shape[1] > 1 else None
  data = data[[outcome_column] + (features or [])]
  if data[outcome_column].count() < 3:
    raise ValueError("Input data must have at least 3 observations.")
  if data[features or []].isna().values.any():
    raise ValueError("Input data cannot have any missing values.")
  if not data.dtypes.map(pd.api.types.is_numeric_dtype).all():
    raise ValueError("Input data must contain only numeric values.")
  return data, outcome_column, features


class CausalImpactData:
  """Validated, split and (optionally) standardised input.

  Same attribute names as the reference class (data.py:26-137) so downstream
  code written against it keeps working: ``data, pre_period, post_period,
  outcome_column, feature_columns, standardize_data, pre_data, after_pre_data,
  num_steps_forecast, model_pre_data, model_after_pre_data, outcome_scaler,
  feature_ts, outcome_ts``.
  """

  def __init__(self, data, pre_period: Tuple[InputDateType, InputDateType],
               post_period: Tuple[InputDateType, InputDateType],
               outcome_column: Optional[str] = None, standardize_data: bool = True,
               dtype=np.float32):
    data = pd.DataFrame(data)
    self.pre_period, self.post_period = parse_and_validate_date_data(data, pre_period,
                                                                     post_period)
    self.data, self.outcome_column, self.feature_columns = _select_columns(data, outcome_column)
    del data
    self.standardize_data = standardize_data
    idx = self.data.index
    self.pre_data = self.data.loc[(idx >= self.pre_period[0]) & (idx <= self.pre_period[1])]
    # everything after the pre-period: the gap and the tail are forecast too
    self.after_pre_data = self.data.loc[idx > self.pre_period[1]]
    self.num_steps_forecast = len(self.after_pre_data.index)
    if standardize_data:
      all_cols = Scaler().fit(self.pre_data)
      self.outcome_scaler = Scaler().fit(self.pre_data[self.outcome_column])
      self.model_pre_data = all_cols.transform(self.pre_data)
      self.model_after_pre_data = all_cols.transform(self.after_pre_data)
    else:
      self.outcome_scaler = None
      self.model_pre_data = self.pre_data
      self.model_after_pre_data = self.after_pre_data
    np_dtype = np.dtype(dtype)
    series = np.asarray(self.model_pre_data[self.outcome_column], dtype=np_dtype)
    self.outcome_ts = MaskedSeries(time_series=series, is_missing=np.isnan(series))
    if self.feature_columns is not None:
      # FULL length: post-period covariates drive the counterfactual
      self.feature_ts = pd.concat([self.model_pre_data[self.feature_columns],
                                   self.model_after_pre_data[self.feature_columns]], axis=0)
      self.feature_ts["intercept_"] = 1.
    else:
      self.feature_ts = None

  # ---- what the engine needs (causalimpact_lib.py:545-564) ----
  def engine_inputs(self, dtype=np.float32):
    """(y_ext [T] with NaN for masked steps, design [T,p] or None, outcome_sd)."""
    np_dtype = np.dtype(dtype)
    y_pre = self.outcome_ts.time_series.astype(np_dtype)
    y_ext = np.concatenate([y_pre, np.full(self.num_steps_forecast, np.nan, dtype=np_dtype)])
    design = None if self.feature_ts is None else np.asarray(self.feature_ts.values,
                                                             dtype=np_dtype)
    outcome_sd = float(np.asarray(np.nanstd(y_pre, ddof=1), dtype=np_dtype))
    return y_ext.astype(np.float64), None if design is None else design.astype(np.float64), \
        outcome_sd
