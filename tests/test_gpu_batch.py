"""Batches of independent series: ci_set_data_batch,
ci_batch_select, ci_gibbs_run_batch_d and fit_causalimpact_many.

Parity bar: BIT-IDENTICAL to fitting every series on its own (same kernels, same Philox
keys) -- series, summary and latent draws; plus the select path against a fresh upload."""
import numpy as np
import pandas as pd
import pytest

import causalimpact_b200 as ci
from causalimpact_b200 import EngineError
from causalimpact_b200 import frame as fr

pytestmark = pytest.mark.gpu


def panel(n_series, n, n_cov, seed, treat):
  rng = np.random.default_rng(seed)
  idx = pd.date_range("2021-01-01", periods=n, freq="D")
  out = []
  for s in range(n_series):
    xs = 100 + np.cumsum(rng.normal(size=(n, n_cov)), axis=0) * 0.3
    y = (1.0 + 0.1 * s) * xs[:, 0] + rng.normal(size=n) * (0.5 + 0.2 * s)
    y[treat:] += 2.0 + s
    if s % 2:
      y[3 + s] = np.nan                               # a missing pre-period point
    out.append(pd.DataFrame(np.column_stack([y, xs]), index=idx,
                            columns=["y"] + [f"x{i}" for i in range(n_cov)]))
  return out


@pytest.mark.parametrize("n_cov", [2, 5], ids=["dense_prior", "spike_and_slab"])
def test_many_equals_one_by_one(n_cov):
  dfs = panel(5, 120, n_cov, 3, 90)
  pre, post = (dfs[0].index[0], dfs[0].index[89]), (dfs[0].index[90], dfs[0].index[-1])
  kw = dict(seed=(2, 5), inference_options=ci.InferenceOptions(num_results=150))
  eo = ci.EngineOptions(num_chains=12, sampler="gibbs", decorrelate_series=False)
  many = ci.fit_causalimpact_many(dfs, pre, post, engine_options=eo, **kw)
  assert len(many) == 5
  for df, got in zip(dfs, many):
    one = ci.fit_causalimpact(df, pre, post, engine_options=eo, **kw)
    pd.testing.assert_frame_equal(got.series, one.series)
    pd.testing.assert_frame_equal(got.summary, one.summary)
    np.testing.assert_array_equal(got.posterior_samples.level, one.posterior_samples.level)
    np.testing.assert_array_equal(got.posterior_samples.weights, one.posterior_samples.weights)
    np.testing.assert_array_equal(got.posterior_samples.observation_noise_scale,
                                  one.posterior_samples.observation_noise_scale)
    np.testing.assert_array_equal(got.diagnostics["inclusion"], one.diagnostics["inclusion"])
  # the series really differ (no accidental aliasing of one series' tiles)
  e = [float(r.summary.loc["average", "abs_effect"]) for r in many]
  assert len({round(v, 6) for v in e}) == 5 and all(abs(v - (2.0 + s)) < 1.5 for s, v in enumerate(e))


def test_select_runs_single_series_entry_points(engine):
  dfs = panel(4, 300, 3, 8, 210)
  specs = []
  for d in dfs:
    cid = fr.CausalImpactData(d, (d.index[0], d.index[209]), (d.index[210], d.index[-1]))
    y_ext, design, sd = cid.engine_inputs(np.float32)
    specs.append(ci.build_problem(y_ext, design, outcome_sd=sd))
  rng = np.random.default_rng(0)
  th = np.tile(ci.initial_theta(specs[0]), (9, 1)) + 0.1 * rng.normal(size=(9, specs[0].dim))
  want = []
  for sp in specs:
    engine.set_data(sp)
    want.append(engine.logprob_grad(th, with_prior=True))
  engine.set_data_batch(specs)
  for i in (2, 0, 3, 1):
    engine.batch_select(i)
    v, g = engine.logprob_grad(th, with_prior=True)
    np.testing.assert_array_equal(v, want[i][0])
    np.testing.assert_array_equal(g, want[i][1])
  with pytest.raises(EngineError, match="out of range"):
    engine.batch_select(4)
  bad = ci.build_problem(specs[0].y[:-1], specs[0].X[:-1])
  with pytest.raises(ValueError, match="share"):
    engine.set_data_batch([specs[0], bad])
  engine.set_data(specs[0])
  with pytest.raises(EngineError, match="ci_set_data_batch"):
    engine._check(engine._lib.ci_batch_select(engine._ctx, 0))


def test_no_covariates_batch():
  rng = np.random.default_rng(4)
  idx = pd.date_range("2020-01-01", periods=100, freq="D")
  dfs = [pd.DataFrame({"y": 5 + np.cumsum(rng.normal(size=100)) * 0.1 + rng.normal(size=100)},
                      index=idx) for _ in range(3)]
  for d in dfs:
    d.iloc[70:, 0] += 3
  many = ci.fit_causalimpact_many(dfs, (idx[0], idx[69]), (idx[70], idx[-1]), seed=1,
                                  inference_options=ci.InferenceOptions(num_results=64),
                                  engine_options=ci.EngineOptions(decorrelate_series=False))
  for df, got in zip(dfs, many):
    one = ci.fit_causalimpact(df, (idx[0], idx[69]), (idx[70], idx[-1]), seed=1,
                              inference_options=ci.InferenceOptions(num_results=64),
                              engine_options=ci.EngineOptions(sampler="gibbs"))
    pd.testing.assert_frame_equal(got.series, one.series)
    assert got.posterior_samples.weights is None


def test_panel_arrays_match_the_frame_path():
  """fit_causalimpact_panel (vectorised numpy prep, one read-back, arrays out) vs
  fit_causalimpact_many (pandas per series): same kernels and Philox keys; the inputs can differ
  by one float64 ulp in the pre-period mean / sd, so agreement is to rounding (rtol 2e-4 on the
  original scale), with the dense prior (<= 2 covariates) where the sweep is continuous in its
  inputs."""
  dfs = panel(5, 120, 2, 11, 90)
  idx = dfs[0].index
  pre, post = (idx[0], idx[89]), (idx[92], idx[-2])          # a gap and a tail
  # (decorrelate_series=False keeps fit_causalimpact_many on its per-frame pandas preparation;
  # with the default these stackable frames would take the panel route themselves)
  kw = dict(seed=7, inference_options=ci.InferenceOptions(num_results=120),
            engine_options=ci.EngineOptions(num_chains=10, decorrelate_series=False))
  many = ci.fit_causalimpact_many(dfs, pre, post, **kw)
  values = np.stack([d.values for d in dfs])
  res = ci.fit_causalimpact_panel(values, idx, pre, post, keep_level=True, **kw)
  assert res.series.shape == (5, 120, 10) and res.summary.shape == (5, 2, 15)
  assert res.level.shape == (5, 120, 120) and res.weights.shape == (5, 120, 3)
  vals = ci.impact.SERIES_VALUE_COLUMNS
  for i, one in enumerate(many):
    want = one.series[vals].values.astype(float)
    scale = np.nanmax(np.abs(want))
    np.testing.assert_allclose(res.series[i], want, rtol=2e-4, atol=2e-4 * scale, equal_nan=True)
    np.testing.assert_array_equal(np.isnan(res.series[i]), np.isnan(want))
    np.testing.assert_allclose(res.summary[i], one.summary.values.astype(float), rtol=2e-4,
                               atol=2e-4 * scale)
    ser, summ = res.frames(i)
    assert list(ser.columns) == list(one.series.columns)
    assert list(summ.columns) == list(one.summary.columns)
    np.testing.assert_allclose(res.level[i], one.posterior_samples.level, rtol=2e-3, atol=2e-3)
  with pytest.raises(ValueError, match="constant"):
    bad = values.copy(); bad[2, :, 0] = 1.0
    ci.fit_causalimpact_panel(bad, idx, pre, post, **kw)


def test_seasonal_panel_and_many_equal_single_seasonal_fits():
  """Seasonal components in the batched paths: fit_causalimpact_many is bit-identical to the
  single seasonal fits (per-series prior scales travel with the batch), the panel agrees to
  rounding."""
  rng = np.random.default_rng(21)
  n = 140
  idx = pd.date_range("2022-01-03", periods=n, freq="D")
  pat = np.array([1.0, 4.0, 5.0, 2.0, -1.0, -2.0, -3.0])
  dfs = []
  for s in range(3):
    x = 100 + np.cumsum(rng.normal(size=n)) * 0.3
    y = x + (0.5 + 0.2 * s) * pat[np.arange(n) % 7] + 0.3 * rng.normal(size=n)
    y[100:] += 2.0
    dfs.append(pd.DataFrame({"y": y, "x": x}, index=idx))
  pre, post = (idx[0], idx[99]), (idx[100], idx[-1])
  mo = ci.ModelOptions(seasons=[ci.Seasons(num_seasons=7)])
  kw = dict(seed=5, model_options=mo, inference_options=ci.InferenceOptions(num_results=96),
            engine_options=ci.EngineOptions(num_chains=8, decorrelate_series=False))
  many = ci.fit_causalimpact_many(dfs, pre, post, **kw)
  res = ci.fit_causalimpact_panel(np.stack([d.values for d in dfs]), idx, pre, post,
                                  keep_level=True, **kw)
  assert res.seasonal_levels.shape == (3, 96, n, 1) and res.seasonal_drift_scales.shape == (3, 96, 1)
  vals = ci.impact.SERIES_VALUE_COLUMNS
  for i, df in enumerate(dfs):
    one = ci.fit_causalimpact(df, pre, post, **kw)
    pd.testing.assert_frame_equal(many[i].series, one.series)
    pd.testing.assert_frame_equal(many[i].summary, one.summary)
    np.testing.assert_array_equal(many[i].posterior_samples.seasonal_levels,
                                  one.posterior_samples.seasonal_levels)
    np.testing.assert_array_equal(many[i].posterior_samples.seasonal_drift_scales,
                                  one.posterior_samples.seasonal_drift_scales)
    want = one.series[vals].values.astype(float)
    np.testing.assert_allclose(res.series[i], want, rtol=5e-4, atol=5e-4 * np.nanmax(np.abs(want)),
                               equal_nan=True)
    # the weekly pattern is found in every series
    contrib = one.posterior_samples.seasonal_levels.numpy()[:, :98, 0].mean(0).reshape(14, 7).mean(0)
    assert np.corrcoef(contrib, pat)[0, 1] > 0.95


def test_device_panel_prep_matches_the_host_restatement(engine):
  """ci_set_panel (data.py:77-137 + priors + tiles + Gram matrices for N series in ONE kernel) vs the
  host path (panel.prepare_panel + build_problem + ci_set_data_batch): the per-series statistics
  agree to float64 rounding, and every single-series entry point sees the same problem
  (log-posterior and gradient of series i through ci_batch_select agree to float32 rounding: a
  standardized value can differ by one float32 ulp where the two summation orders of the
  pre-period mean / sd differ in the last float64 bit)."""
  from causalimpact_b200 import panel as pn
  rng = np.random.default_rng(12)
  N, T, k = 6, 150, 3
  xs = 100 + np.cumsum(rng.normal(size=(N, T, k)), axis=1) * 0.3
  y = xs[:, :, 0] * 1.1 - 0.4 * xs[:, :, 1] + rng.normal(size=(N, T))
  y[:, 100:] += 2.0
  y[1, 7] = np.nan; y[4, 0] = np.nan; y[4, 33] = np.nan            # missing pre-period points
  vals = np.concatenate([y[:, :, None], xs], axis=2)
  row0, n_pre = 5, 95                                               # rows before the pre-period are ignored
  for std in (True, False):
    stats = engine.set_panel(vals, row0=row0, n_pre=n_pre, standardize=std, prior_level_sd=0.02)
    prep = pn.prepare_panel(vals, np.arange(T), (row0, row0 + n_pre - 1), (row0 + n_pre, T - 1), std,
                            np.float32)
    np.testing.assert_allclose(stats[:, 0], prep["y_scale"], rtol=1e-12)
    np.testing.assert_allclose(stats[:, 1], prep["y_offset"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(stats[:, 2], prep["outcome_sd"], rtol=2e-6)
    assert np.array_equal(stats[:, 3], np.sum(~np.isnan(prep["y_ext"]), axis=1))
    specs = [ci.build_problem(prep["y_ext"][i], prep["design"][i], prior_level_sd=0.02,
                              outcome_sd=float(prep["outcome_sd"][i])) for i in range(N)]
    rng2 = np.random.default_rng(1)
    th = np.tile(ci.initial_theta(specs[0], 0.02), (5, 1)) + 0.05 * rng2.normal(size=(5, specs[0].dim))
    got = []
    for i in range(N):
      engine.batch_select(i)
      got.append(engine.logprob_grad(th, with_prior=True))
    for i in range(N):
      engine.set_data(specs[i])
      v, g = engine.logprob_grad(th, with_prior=True)
      np.testing.assert_allclose(got[i][0], v, rtol=1e-4, atol=1e-2)
      np.testing.assert_allclose(got[i][1], g, rtol=2e-3, atol=5e-2)
  # the reference's input errors (data.py:140-190), raised from the device-side validation
  bad = vals.copy(); bad[2, :, 0] = 3.0
  with pytest.raises(ValueError, match="constant"):
    engine.set_panel(bad, row0=row0, n_pre=n_pre)
  bad = vals.copy(); bad[3, 40, 2] = np.nan
  with pytest.raises(ValueError, match="missing values"):
    engine.set_panel(bad, row0=row0, n_pre=n_pre)
  # no covariates
  stats = engine.set_panel(vals[:, :, :1], row0=0, n_pre=100)
  assert engine.spec.p == 0 and np.all(stats[:, 3] >= 98)
  d, l, t, inc = engine.gibbs_run_batch_t(3, n_warmup=5, n_results=3, seed=1)
  assert bool(d.isfinite().all()) and tuple(d.shape) == (N, 9, 2)
