"""End-to-end ``fit_causalimpact`` on the GPU.  These are the reference's own
hot-path property tests (causalimpact_lib_test.py, line numbers cited per
test) re-run against the B200 engine, plus parity against the golden vectors
and the restated reference sampler."""
import glob
import os

import numpy as np
import pandas as pd
import pytest

import causalimpact_b200 as ci
from causalimpact_b200 import frame as fr
from causalimpact_b200 import impact
from oracle import gibbs_np as G
from oracle import kalman_np as K
from oracle import smoother_np as SM
from test_postproc_golden import GOLDEN, load_case, oracle_impact

pytestmark = pytest.mark.gpu


def csv_data():
  """The reference's 91-row fixture with y[1,3,7] = NaN (lib_test.py:205-221),
  rebuilt from the golden file (no reference files are read at test time)."""
  _, data, _, _ = load_case([p for p in GOLDEN if p.endswith("postproc_csv.npz")][0])
  data.index.freq = "10s"
  pre = (data.index[0], data.index[59])
  post = (data.index[60], data.index[-1])
  return data, pre, post


def synthetic(n=100, treat=50, amt=5.0, seed=0):
  """y = 1.2 x + N(0,1), +amt after `treat` (lib_test.py:33-46 with an AR(1) x)."""
  rng = np.random.default_rng(seed)
  a = np.zeros(n)
  for t in range(1, n):
    a[t] = 0.9 * a[t - 1] + rng.normal()
  x = 100 + a
  y = 1.2 * x + rng.normal(size=n)
  df = pd.DataFrame({"y": y, "x": x}, index=pd.date_range("2018-01-01", periods=n, freq="D"))
  df.loc[df.index > df.index[treat], "y"] += amt
  return df


def test_unexpected_kwargs_raise():                    # lib_test.py:231-240
  data, pre, post = csv_data()
  with pytest.raises(TypeError):
    ci.fit_causalimpact(some_unknown_arg=3, data=data, pre_period=pre, post_period=post,
                        inference_options=ci.InferenceOptions(num_results=10), seed=(1, 2))


def test_unsupported_model_options_fail_loudly():
  data, pre, post = csv_data()
  with pytest.raises(ValueError, match="state dimension"):           # 1 + 200 > 192
    ci.fit_causalimpact(data, pre, post, seed=1,
                        model_options=ci.ModelOptions(seasons=[ci.Seasons(num_seasons=200)]))
  with pytest.raises(NotImplementedError):
    ci.fit_causalimpact(data, pre, post, seed=1, experimental_model=object())
  # seasons themselves are supported (csrc/ci_seasonal.cuh): data.csv with a weekly component
  res = ci.fit_causalimpact(data, pre, post, seed=1,
                            model_options=ci.ModelOptions(seasons=[ci.Seasons(num_seasons=7)]),
                            inference_options=ci.InferenceOptions(num_results=64))
  assert res.posterior_samples.seasonal_levels.shape == (64, len(data), 1)
  assert res.posterior_samples.weights.shape == (64, 3)


@pytest.mark.parametrize("prior_level_sd", [0.01, 0.1, 0.5])
def test_prior_level_sd_is_used(prior_level_sd):       # lib_test.py:242-271
  data, _, _ = csv_data()
  res = ci.fit_causalimpact(
      data, (data.index[0], data.index[19]), (data.index[20], data.index[-1]),
      inference_options=ci.InferenceOptions(num_results=100, num_warmup_steps=100),
      model_options=ci.ModelOptions(prior_level_sd=prior_level_sd), seed=(0, 0))
  np.testing.assert_allclose(np.mean(res.posterior_samples.level_scale), prior_level_sd,
                             atol=0.2 * prior_level_sd)


def test_intercept_and_shapes_with_covariates():       # lib_test.py:286-295, 319-338, 361-379
  data, pre, post = csv_data()
  res = ci.fit_causalimpact(data, pre, post, seed=(1, 1),
                            inference_options=ci.InferenceOptions(num_results=10,
                                                                  num_warmup_steps=100))
  ps = res.posterior_samples
  assert ps.weights.shape == (10, 3)                     # 2 features + intercept
  assert not np.any(np.isnan(ps.level.numpy())) and not np.any(np.isnan(ps.weights.numpy()))
  assert np.all(ps.observation_noise_scale.numpy() <= 1.2 + 1e-6)
  assert np.all(ps.level_scale.numpy() <= 1.0 + 1e-6)
  assert np.all(ps.weights.numpy() != 0.0)               # inclusion prob 1 for p <= 3
  assert ps.level.shape == (10, 91)
  assert res.series.index.equals(data.index)


def test_prediction_dims_no_covariates():              # lib_test.py:340-359
  data, pre, post = csv_data()
  res = ci.fit_causalimpact(data[["y"]], pre, post, seed=3,
                            inference_options=ci.InferenceOptions(num_results=17))
  assert res.posterior_samples.weights is None
  assert res.series.index.equals(data.index)
  assert res.posterior_samples.level.shape[0] == 17
  assert res.posterior_samples.observation_noise_scale.shape == (17,)


def test_same_seed_is_bit_identical():                 # lib_test.py:462-502
  df = synthetic(seed=13)
  kw = dict(pre_period=(df.index[0], df.index[49]), post_period=(df.index[50], df.index[-1]),
            inference_options=ci.InferenceOptions(num_results=50))
  for seed in ((13, 37), 14):
    a = ci.fit_causalimpact(df, seed=seed, **kw)
    b = ci.fit_causalimpact(df, seed=seed, **kw)
    pd.testing.assert_frame_equal(a.series, b.series)
    pd.testing.assert_frame_equal(a.summary, b.summary)
  c = ci.fit_causalimpact(df, seed=15, **kw)
  assert not a.summary.equals(c.summary)


@pytest.mark.parametrize("n", [100, 150])
def test_summary_recovers_effect(n):                   # lib_test.py:504-535
  df = synthetic(n=n, seed=1)
  res = ci.fit_causalimpact(df, (df.index[0], df.index[49]), (df.index[50], df.index[99]),
                            seed=0, inference_options=ci.InferenceOptions(num_results=100))
  np.testing.assert_allclose(res.summary.loc["cumulative", "abs_effect"], 250, rtol=0.2)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_numeric_impact_values(dtype):                 # lib_test.py:655-702
  rng = np.random.default_rng(5)
  y = rng.normal(0, 1e-2, size=100)
  y[50:] += 5.0
  df = pd.DataFrame({"y": y})
  res = ci.fit_causalimpact(df, (0, 49), (50, 99), seed=1,
                            data_options=ci.DataOptions(dtype=dtype),
                            inference_options=ci.InferenceOptions(num_results=1000))
  s = res.summary
  np.testing.assert_allclose(s.loc["average", "abs_effect"], 5.0, rtol=1e-3)
  np.testing.assert_allclose(s.loc["cumulative", "abs_effect"], 250.0, rtol=1e-3)
  width = (s.loc["average", "abs_effect_upper"] - s.loc["average", "abs_effect_lower"]) / 5.0
  assert width <= 0.01
  assert res.posterior_samples.level.dtype == dtype


def test_gap_between_pre_and_post_period():            # lib_test.py:564-653
  rng = np.random.default_rng(2)
  n = 100
  y = 0.01 * rng.normal(size=n)
  y[60:] += 3.0
  df = pd.DataFrame({"y": y}, index=pd.date_range("2020-01-01", periods=n))
  res = ci.fit_causalimpact(df, (df.index[5], df.index[39]), (df.index[60], df.index[89]),
                            seed=4, inference_options=ci.InferenceOptions(num_results=50))
  s = res.series
  eff = ["point_effects_mean", "point_effects_lower", "cumulative_effects_mean",
         "cumulative_effects_upper"]
  assert s.iloc[:5][eff + ["posterior_mean"]].isna().all().all()        # before pre-period
  assert s.iloc[5:40][eff].notna().all().all()                           # pre-period
  assert (s.iloc[5:40]["cumulative_effects_mean"] == 0).all()
  assert s.iloc[40:60][eff].isna().all().all()                           # the gap
  assert s.iloc[40:60]["posterior_mean"].notna().all()
  assert s.iloc[60:90][eff].notna().all().all()                          # post-period
  assert s.iloc[90:][eff].isna().all().all()                             # after post
  assert s.iloc[90:]["posterior_mean"].notna().all()
  np.testing.assert_allclose(res.summary.loc["average", "abs_effect"], 3.0, rtol=0.05)


def test_missing_pre_period_observations():            # lib_test.py:814-844
  rng = np.random.default_rng(8)
  df = pd.DataFrame(rng.normal(size=(200, 3)), columns=["y", "x1", "x2"])
  df.iloc[2:5, 0] = np.nan
  res = ci.fit_causalimpact(df, (0, 99), (100, 199), seed=2,
                            inference_options=ci.InferenceOptions(num_results=20))
  eff = [c for c in res.series.columns if c.startswith(("point_", "cumulative_"))]
  assert res.series.iloc[2:5][eff].isna().all().all()
  assert res.series.iloc[5:100][eff].notna().all().all()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[9:-4] for p in GOLDEN])
def test_postprocessing_on_gpu_matches_reference_golden(engine, path):
  """_compute_impact through ci_impact (host buffers, float64) vs the reference's own
  outputs (golden files)."""
  g, data, pre, post = load_case(path)
  cid = fr.CausalImpactData(data, pre, post, standardize_data=bool(g["standardize"]))
  series, summary = impact.compute_impact(g["posterior_means"], g["posterior_trajectories"], cid,
                                          float(g["alpha"]), engine.impact)
  cols = [str(c) for c in g["series_columns"]]
  rtol, atol = (1e-11, 1e-11) if bool(g["standardize"]) else \
      (1e-5, 4e-6 * float(np.nanmax(np.abs(g["series_values"]))))
  np.testing.assert_allclose(series[cols].values.astype(float), g["series_values"], rtol=rtol,
                             atol=atol, equal_nan=True)
  np.testing.assert_allclose(summary.values.astype(float), g["summary_values"], rtol=rtol,
                             atol=atol, equal_nan=True)


def test_statistical_parity_with_restated_reference_sampler():
  """BASELINE config 1 (quickstart shape): posterior mean and credible interval
  of the counterfactual vs the restated reference Gibbs sampler + the
  reference's own post-processing semantics.  Tolerance: 5 combined MC standard
  errors, floor 1e-3 * sd_y."""
  df = synthetic(n=100, treat=70, amt=10.0, seed=3)
  pre, post = (df.index[0], df.index[69]), (df.index[70], df.index[-1])
  res = ci.fit_causalimpact(df, pre, post, seed=11,
                            inference_options=ci.InferenceOptions(num_results=2000))
  cid = fr.CausalImpactData(df, pre, post)
  y_ext, design, sd = cid.engine_inputs(np.float32)
  prob = K.default_problem(y_ext, design, outcome_sd=sd)
  gb = G.run(prob, n_results=4000, n_warmup=500, seed=5)
  rng = np.random.default_rng(0)
  loc = gb["level"] + gb["w"] @ design.T
  traj = loc + np.sqrt(gb["s_e"])[:, None] * rng.normal(size=loc.shape)
  ser_o, sum_o = impact.compute_impact(loc.mean(0), traj, cid, 0.05, oracle_impact)
  sd_y = float(np.nanstd(df["y"].values[:70], ddof=1))
  for col in ("abs_effect", "abs_effect_lower", "abs_effect_upper", "predicted"):
    a, b = res.summary.loc["average", col], sum_o.loc["average", col]
    se = res.summary.loc["average", "abs_effect_sd"] * np.sqrt(1 / 100 + 1 / 200) * 3
    assert abs(a - b) < max(5 * se, 1e-3 * sd_y), (col, a, b, se)
  post_rows = res.series.index >= post[0]
  d = (res.series.loc[post_rows, "posterior_mean"] - ser_o.loc[post_rows, "posterior_mean"]).abs()
  assert d.max() < 0.1 * sd_y


def test_large_prior_level_sd_starts_inside_support():
  """prior_level_sd > 1 puts the reference's initial level scale above its own
  upper bound (lib.py:432, 572); chains must still start inside the support."""
  df = synthetic(n=80, treat=50, seed=4)
  res = ci.fit_causalimpact(df, (df.index[0], df.index[49]), (df.index[50], df.index[-1]), seed=2,
                            model_options=ci.ModelOptions(prior_level_sd=1.5),
                            inference_options=ci.InferenceOptions(num_results=40))
  assert np.all(np.isfinite(res.summary.values.astype(float)))
  assert res.diagnostics["accept_rate"].mean() > 0.3
  assert np.all(res.posterior_samples.level_scale.numpy() <= 1.0 + 1e-6)


def test_ten_covariates_match_slab_only_gibbs():
  """BASELINE config-2 family (10 covariates): the engine samples the SLAB-ONLY
  model.  Check the whole fit -- whitening, HMC, mapping the
  weights back, predictive draws -- against the restated Gibbs sampler with every
  feature included: regression weights and the counterfactual agree within MC error."""
  rng = np.random.default_rng(12)
  n, k = 300, 10
  xs = 100 + np.cumsum(rng.normal(size=(n, k)), axis=0) * 0.3
  beta = np.zeros(k); beta[:3] = (1.2, 0.6, -0.4)
  y = xs @ beta + rng.normal(size=n)
  y[210:] += 8.0
  df = pd.DataFrame(np.column_stack([y, xs]), columns=["y"] + [f"x{i}" for i in range(k)])
  res = ci.fit_causalimpact(df, (0, 209), (210, 299), seed=3,
                            inference_options=ci.InferenceOptions(num_results=1500),
                            engine_options=ci.EngineOptions(num_chains=128, min_warmup=500,
                                                            sampler="hmc"))
  assert res.diagnostics["sampler"] == "hmc"
  cid = fr.CausalImpactData(df, (0, 209), (210, 299))
  y_ext, design, sd = cid.engine_inputs(np.float32)
  prob = K.default_problem(y_ext, design, outcome_sd=sd)
  gb = G.run(prob, n_results=6000, n_warmup=1000, seed=8)
  w_hmc = res.posterior_samples.weights.numpy()
  assert w_hmc.shape == (1500, k + 1)
  for j in range(k):                          # slopes (the intercept is confounded with the level)
    a, b = w_hmc[:, j], gb["w"][:, j]
    se = np.sqrt(a.var() / (a.size / 10) + b.var() / (b.size / 30))
    assert abs(a.mean() - b.mean()) < max(5 * se, 5e-3), (j, a.mean(), b.mean(), se)
  loc = gb["level"] + gb["w"] @ design.T
  mean_o = impact.unscale(loc.mean(0), cid)
  post = res.series.index >= 210
  d = np.abs(res.series.loc[post, "posterior_mean"].values - mean_o[210:])
  sd_y = float(np.std(y[:210], ddof=1))
  assert d.max() < 0.05 * sd_y, d.max()
  assert res.diagnostics["n_divergent"].sum() <= 15


def test_auto_sampler_reproduces_reference_spike_and_slab_posterior():
  """With > 2 covariates the reference's prior is spike-and-slab (inclusion prob
  3/p, lib.py:449-450).  sampler="auto" then runs the GPU Gibbs kernel; the
  whole fit is compared with the restated reference sampler (sparse=True):
  sparse weights, inclusion pattern, counterfactual and its uncertainty."""
  rng = np.random.default_rng(12)
  n, k = 300, 10
  xs = 100 + np.cumsum(rng.normal(size=(n, k)), axis=0) * 0.3
  beta = np.zeros(k); beta[:3] = (1.2, 0.6, -0.4)
  y = xs @ beta + rng.normal(size=n)
  y[210:] += 8.0
  df = pd.DataFrame(np.column_stack([y, xs]), columns=["y"] + [f"x{i}" for i in range(k)])
  res = ci.fit_causalimpact(df, (0, 209), (210, 299), seed=3,
                            inference_options=ci.InferenceOptions(num_results=1920))
  assert res.diagnostics["sampler"] == "gibbs"
  cid = fr.CausalImpactData(df, (0, 209), (210, 299))
  y_ext, design, sd = cid.engine_inputs(np.float32)
  prob = K.default_problem(y_ext, design, outcome_sd=sd)
  gb = G.run(prob, n_results=5000, n_warmup=800, seed=8, sparse=True)
  w = res.posterior_samples.weights.numpy()
  inc_gpu, inc_ref = (w != 0).mean(0), (gb["w"] != 0).mean(0)
  np.testing.assert_allclose(inc_gpu, inc_ref, atol=0.12)
  assert (w == 0).any()                                    # exact zeros, like the reference
  loc = gb["level"] + gb["w"] @ design.T
  rng2 = np.random.default_rng(0)
  traj = loc + np.sqrt(gb["s_e"])[:, None] * rng2.normal(size=loc.shape)
  ser_o, sum_o = impact.compute_impact(loc.mean(0), traj, cid, 0.05, oracle_impact)
  sd_y = float(np.std(y[:210], ddof=1))
  for col in ("abs_effect", "abs_effect_lower", "abs_effect_upper", "predicted"):
    a, b = res.summary.loc["average", col], sum_o.loc["average", col]
    assert abs(a - b) < 0.05 * sd_y, (col, a, b)
  ratio = res.summary.loc["average", "abs_effect_sd"] / sum_o.loc["average", "abs_effect_sd"]
  assert 0.75 < ratio < 1.33, ratio


def test_shortest_period_after_pre_period():           # lib_test.py:222-229
  data, _, _ = csv_data()
  res = ci.fit_causalimpact(data, (data.index[0], data.index[-2]), (data.index[-1], data.index[-1]),
                            inference_options=ci.InferenceOptions(num_results=10), seed=(1, 2))
  assert res is not None and res.series.shape[0] == data.shape[0]
  assert np.isfinite(res.summary.loc["average", "abs_effect"])


def test_no_datetime_index_succeeds():                 # lib_test.py:273-284
  data, _, _ = csv_data()
  data = data.copy(); data.index = np.arange(data.shape[0])
  res = ci.fit_causalimpact(data, (data.index[0], data.index[19]), (data.index[20], data.index[-1]),
                            inference_options=ci.InferenceOptions(num_results=10), seed=(0, 0))
  assert res is not None and res.series.index.equals(data.index)


def test_non_aligned_start_time():                     # lib_test.py:537-562
  rng = np.random.default_rng(2)
  y = rng.normal(size=100, scale=0.0001); y[50:] += 5.0
  df = pd.DataFrame({"y": y}, index=pd.date_range("2018-01-07", periods=100, freq="W"))
  res = ci.fit_causalimpact(df, pre_period=("2018-01-10", "2018-01-30"),
                            post_period=("2018-02-02", "2018-02-23"), seed=1,
                            inference_options=ci.InferenceOptions(num_results=10))
  assert res.series.loc[pd.to_datetime("2018-01-28"), "cumulative_effects_mean"] == 0
  assert res.series.loc[pd.to_datetime("2018-02-04"), "cumulative_effects_mean"] != 0


def test_missing_and_healthy_input():                  # lib_test.py:793-811, 778-788
  with pytest.raises(TypeError):
    ci.fit_causalimpact()                              # pylint: disable=no-value-for-parameter
  rng = np.random.default_rng(3)
  data = pd.DataFrame({"y": rng.normal(size=200), "x1": rng.normal(size=200),
                       "x2": rng.normal(size=200)})
  res = ci.fit_causalimpact(data, pre_period=(0, 100), post_period=(101, 199), seed=4,
                            inference_options=ci.InferenceOptions(num_results=10))
  assert data.shape[0] == res.series.shape[0]
  assert "Posterior Inference" in ci.summary(res)
  assert "Posterior Inference" in ci.summary(res, output_format="summary")
  assert "Analysis report" in ci.summary(res, output_format="report")
  assert "Analysis report" in ci.summary(res, "report")
  with pytest.raises(ValueError):
    ci.summary(res, output_format="foo")
  assert res.diagnostics["rhat_log_variances"].shape == (2,)
