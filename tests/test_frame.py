"""Host-side data preparation (frame.py) against the behaviours the reference's
own unit tests pin: indices_test.py:47-153 (period parsing, alignment, error
strings), standardize_test.py:24-118 (nan-aware ddof=1 scaler) and
data_test.py:40-154 (column defaults, splits, validation).  CPU only."""
import os

import numpy as np
import pandas as pd
import pytest

from causalimpact_b200 import frame as fr
from test_postproc_golden import GOLDEN, load_case


@pytest.fixture(scope="module")
def csv():
  """The reference's 91-row, 10-second fixture, rebuilt from the golden file."""
  g, data, _, _ = load_case([p for p in GOLDEN if p.endswith("postproc_csv.npz")][0])
  data = data.copy()
  data.iloc[[1, 3, 7], 0] = np.array([125.0, 128.0, 124.0])   # the golden has NaNs there
  return data


@pytest.mark.parametrize("period,msg", [
    ((pd.Timestamp("2016-01-20 22:41:20"), pd.Timestamp("2016-01-21 22:41:20")),
     "Aligned period end not found in the index."),
    ((pd.Timestamp("2022-01-20 22:41:20"), pd.Timestamp("2022-01-21 22:41:20")),
     "Aligned period start not found in the index."),
    ((pd.Timestamp("2016-02-20 22:41:50"), pd.Timestamp("2016-02-20 22:41:20")),
     "Period end must be after period start. "),
])
def test_align_period_errors(csv, period, msg):                    # indices_test.py:47-71
  with pytest.raises(ValueError, match=msg):
    fr.align_period(period, csv)


@pytest.mark.parametrize("pre,post,msg", [
    ((pd.Timestamp("2016-02-20 22:41:20"), pd.Timestamp("2016-02-20 22:41:50")),
     (pd.Timestamp("2016-02-20 22:41:40"), pd.Timestamp("2016-02-20 22:41:50")),
     "pre_period and post_period cannot overlap."),
    ((pd.Timestamp("2016-02-20 22:41:20"), pd.Timestamp("2016-02-20 22:41:30")),
     (pd.Timestamp("2016-02-20 22:41:40"), pd.Timestamp("2016-02-20 22:41:50")),
     "pre_period must span at least 3 time points."),
])
def test_validate_periods_errors(csv, pre, post, msg):             # indices_test.py:73-95
  with pytest.raises(ValueError, match=msg):
    fr.validate_periods(pre, post, csv)


@pytest.mark.parametrize("pre,post", [
    ((0, 10), (11, 90)),
    (("2016-02-20 22:41:20", "2016-02-20 22:43:00"), ("2016-02-20 22:43:10", "2016-02-20 22:56:20")),
    (("2016-02-20 22:41:15", "2016-02-20 22:43:05"), ("2016-02-20 22:43:06", "2016-02-20 22:56:28")),
    ((pd.Timestamp("2016-02-20 22:41:11"), pd.Timestamp("2016-02-20 22:43:07")),
     (pd.Timestamp("2016-02-20 22:43:09"), pd.Timestamp("2016-02-20 22:56:23"))),
])
def test_period_formats_round_inwards(csv, pre, post):             # indices_test.py:97-139
  got_pre, got_post = fr.parse_and_validate_date_data(csv, pre, post)
  assert got_pre == (pd.Timestamp("2016-02-20 22:41:20"), pd.Timestamp("2016-02-20 22:43:00"))
  assert got_post == (pd.Timestamp("2016-02-20 22:43:10"), pd.Timestamp("2016-02-20 22:56:20"))


def test_integer_index(csv):                                       # indices_test.py:141-153
  data = csv.copy()
  data.index = np.arange(len(data))
  assert fr.parse_and_validate_date_data(data, (0, 10), (11, 90)) == ((0, 10), (11, 90))
  with pytest.raises(ValueError, match="Expected argument to be str, int, or datetime"):
    fr.parse_and_validate_date_data(data, (0.5, 10), (11, 90))


def test_scaler_matches_reference_semantics():                     # standardize_test.py:24-118
  idx = pd.date_range("2022-01-01", periods=3, freq="h")
  df = pd.DataFrame({"x1": [4., 5., 6.], "x2": [100., 101., 102.]}, index=idx)
  sc = fr.Scaler()
  z = sc.fit_transform(df)
  pd.testing.assert_frame_equal(z, pd.DataFrame({"x1": [-1., 0., 1.], "x2": [-1., 0., 1.]},
                                                index=idx))
  pd.testing.assert_frame_equal(sc.inverse_transform(z), df)
  ints = pd.DataFrame({"x": np.int32([4, 5, 6, 12])})
  pd.testing.assert_frame_equal(
      fr.Scaler().fit_transform(ints),
      pd.DataFrame({"x": np.float64([-0.7651691780042776, -0.48692584054817667,
                                     -0.20868250309207573, 1.46077752164453])}))
  idx4 = pd.date_range("2022-01-01", periods=4, freq="h")
  nan_df = pd.DataFrame({"x1": [4., 5., np.nan, 6.], "x2": [98., np.nan, 102., 106.]}, index=idx4)
  pd.testing.assert_frame_equal(
      fr.Scaler().fit_transform(nan_df),
      pd.DataFrame({"x1": [-1., 0., np.nan, 1.], "x2": [-1., np.nan, 0., 1.]}, index=idx4))
  with pytest.raises(fr.NotFittedError):
    fr.Scaler().transform(df)
  const = pd.DataFrame({"c": [3., 3., 3.]})
  pd.testing.assert_frame_equal(fr.Scaler().fit_transform(const), const)   # zero variance: untouched


def test_causalimpact_data_defaults_and_splits(csv):               # data_test.py:40-135
  pre = (csv.index[0], csv.index[59]); post = (csv.index[60], csv.index[-1])
  d = fr.CausalImpactData(csv, pre, post)
  assert d.outcome_column == "y" and d.feature_columns == ["x1", "x2"]
  assert len(d.pre_data) == 60 and len(d.after_pre_data) == 31 and d.num_steps_forecast == 31
  assert list(d.feature_ts.columns) == ["x1", "x2", "intercept_"] and len(d.feature_ts) == 91
  assert d.outcome_ts.time_series.dtype == np.float32 and d.outcome_ts.time_series.shape == (60,)
  np.testing.assert_allclose(d.model_pre_data["y"].mean(), 0.0, atol=1e-12)
  np.testing.assert_allclose(d.model_pre_data["y"].std(ddof=1), 1.0, rtol=1e-12)
  y_ext, design, sd = d.engine_inputs(np.float32)
  assert y_ext.shape == (91,) and np.isnan(y_ext[60:]).all() and not np.isnan(y_ext[:60]).any()
  assert design.shape == (91, 3) and np.all(design[:, -1] == 1.0)
  assert abs(sd - 1.0) < 1e-6
  d2 = fr.CausalImpactData(csv, pre, post, outcome_column="x1")
  assert d2.outcome_column == "x1" and d2.feature_columns == ["y", "x2"]
  d3 = fr.CausalImpactData(csv["y"], pre, post)                    # a Series is accepted
  assert d3.feature_ts is None and d3.feature_columns is None
  d4 = fr.CausalImpactData(csv, pre, post, standardize_data=False)
  assert d4.outcome_scaler is None and d4.model_pre_data is d4.pre_data


def test_causalimpact_data_validation(csv):                        # data_test.py:90-154
  pre = (csv.index[0], csv.index[59]); post = (csv.index[60], csv.index[-1])
  with pytest.raises(KeyError):
    fr.CausalImpactData(csv, pre, post, outcome_column="nope")
  bad = csv.copy(); bad.iloc[5, 1] = np.nan
  with pytest.raises(ValueError, match="cannot have any missing values"):
    fr.CausalImpactData(bad, pre, post)
  const = csv.copy(); const["y"] = 1.0
  with pytest.raises(ValueError, match="cannot be constant"):
    fr.CausalImpactData(const, pre, post)
  text = csv.copy(); text["x1"] = "a"
  with pytest.raises(ValueError, match="only numeric"):
    fr.CausalImpactData(text, pre, post)


def test_seasonal_schedule_matches_oracle_and_reference_examples():
  """model.build_seasonal (product, vectorised) vs oracle/seasonal_np.season_schedule (loop):
  int / per-season tuple / per-cycle nested tuple run lengths, as the reference's own test
  passes them (causalimpact_lib_test.py:741-755); priors of causalimpact_lib.py:471-474."""
  import types
  from causalimpact_b200 import model
  from oracle import seasonal_np as S
  cases = [(4, (2, 1, 1, 1)), (7, 1), (6, ((2, 2, 1, 1, 1, 1), (2, 2, 1, 1, 1, 1))),
           (2, ((1, 2), (3, 1))), (7, 24), (3, (5, 1, 2))]
  T = 211
  seasons = [types.SimpleNamespace(num_seasons=n, num_steps_per_season=s) for n, s in cases]
  sch = model.build_seasonal(seasons, T, 1.3)
  assert sch.K == len(cases) and sch.active.shape == (len(cases), T)
  for k, (n, s) in enumerate(cases):
    idx, ends = S.season_schedule(n, s, T)
    np.testing.assert_array_equal(sch.active[k], idx)
    np.testing.assert_array_equal(sch.ends[k].astype(bool), ends)
  assert sch.init_sd == 1.3 and sch.drift_conc == 0.005 and sch.drift_ub == 1.3
  np.testing.assert_allclose(sch.drift_scale, 5e-7 * 1.3 ** 2)
  assert model.build_seasonal([], T, 1.0) is None
  for bad in ((4, (1, 2, 3)), (3, 0), (3, (1.5, 1, 1)), (1, 1)):
    with pytest.raises(ValueError):
      model.build_seasonal([types.SimpleNamespace(num_seasons=bad[0], num_steps_per_season=bad[1])],
                           T, 1.0)


def test_split_rhat():
  from causalimpact_b200.api import split_rhat
  rng = np.random.default_rng(0)
  good = rng.normal(size=(16, 200, 3))
  r = split_rhat(good)
  assert r.shape == (3,) and np.all(np.abs(r - 1.0) < 0.02)
  bad = good.copy(); bad[:8, :, 1] += 3.0                  # half of the chains sit elsewhere
  r = split_rhat(bad)
  assert r[1] > 1.5 and abs(r[0] - 1.0) < 0.02
  assert np.isnan(split_rhat(rng.normal(size=(4, 3, 2)))).all()


def test_prepare_panel_matches_the_pandas_data_prep():
  """panel.prepare_panel (vectorised numpy over N series) vs CausalImpactData.engine_inputs
  (the mirror of data.py:77-137) series by series: the engine inputs agree to one ulp of the
  engine dtype; validation errors are the reference's."""
  import causalimpact_b200 as ci
  rng = np.random.default_rng(0)
  N, T, k = 6, 120, 2
  idx = pd.date_range("2021-01-01", periods=T, freq="D")
  vals = np.empty((N, T, 1 + k))
  for s in range(N):
    xs = 100 + np.cumsum(rng.normal(size=(T, k)), axis=0) * 0.3
    y = xs[:, 0] + rng.normal(size=T); y[90:] += 3
    if s % 2:
      y[4 + s] = np.nan
    vals[s] = np.column_stack([y, xs])
  pre, post = (idx[0], idx[89]), (idx[95], idx[-3])
  prep = ci.prepare_panel(vals, idx, pre, post)
  assert prep["y_ext"].shape == (N, T) and prep["design"].shape == (N, T, k + 1)
  eps = np.finfo(np.float32).eps
  for s in range(N):
    cid = fr.CausalImpactData(pd.DataFrame(vals[s], index=idx, columns=["y", "a", "b"]), pre, post)
    y_ext, design, sd = cid.engine_inputs(np.float32)
    np.testing.assert_allclose(prep["y_ext"][s], y_ext, rtol=2 * eps, atol=2 * eps, equal_nan=True)
    np.testing.assert_allclose(prep["design"][s], design, rtol=2 * eps, atol=2 * eps)
    np.testing.assert_allclose(prep["outcome_sd"][s], sd, rtol=2 * eps)
    np.testing.assert_allclose(prep["y_scale"][s], float(cid.outcome_scaler.stddev_), rtol=1e-14)
    np.testing.assert_allclose(prep["y_offset"][s], float(cid.outcome_scaler.mean_), rtol=1e-14)
  no_cov = ci.prepare_panel(vals[:, :, :1], idx, pre, post)
  assert no_cov["design"] is None
  raw = ci.prepare_panel(vals, idx, pre, post, standardize_data=False)
  np.testing.assert_array_equal(raw["y_scale"], 1.0)
  with pytest.raises(ValueError, match="missing"):
    bad = vals.copy(); bad[1, 5, 1] = np.nan
    ci.prepare_panel(bad, idx, pre, post)
  with pytest.raises(ValueError, match="overlap"):
    ci.prepare_panel(vals, idx, (idx[0], idx[50]), (idx[40], idx[-1]))
