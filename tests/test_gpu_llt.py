"""GPU parity for the local-linear-trend model (BASELINE config 3 family;
capability extension -- the reference has no slope, causalimpact_lib.py:496).
CUDA scan kernels (Sarkka/Garcia-Fernandez elements + congruence adjoints) vs
the generic float64 oracle (oracle/kalman_np.py gen_filter_grad).
Tolerances: float32 value 3e-5 rel + 5e-3, gradient 5e-3 rel + 5e-2;
float64 1e-9 / 1e-6."""
import numpy as np
import pytest

import causalimpact_b200 as cib
from causalimpact_b200 import _engine
from conftest import make_series, make_thetas
from oracle import hmc_np as H
from oracle import kalman_np as K

pytestmark = pytest.mark.gpu
LLT = _engine.MODEL_LOCAL_LINEAR_TREND


@pytest.mark.parametrize("T,n_cov,C", [(100, 1, 8), (257, 0, 4), (700, 12, 20), (600, 40, 7)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_llt_logprob_and_grad_match_oracle(engine, T, n_cov, C, dtype):
  y, X, _ = make_series(T, n_cov, 300 + T, nan_frac=0.02)
  spec = cib.build_problem(y, X, prior_level_sd=0.05, model=LLT, dtype=dtype)
  engine.set_data(spec)
  prob = K.default_problem(y, X, prior_level_sd=0.05, model=K.MODEL_LOCAL_LINEAR_TREND)
  assert spec.dim == prob.dim == (0 if X is None else X.shape[1]) + 3
  th = make_thetas(spec.dim, spec.p, C, 5, d=2).astype(dtype).astype(np.float64)
  for with_prior in (False, True):
    val, grad = engine.logprob_grad(th, with_prior=with_prior)
    v_only = engine.logprob(th, with_prior=with_prior)
    ov, og = K.log_post_grad(prob, th) if with_prior else K.log_lik_grad(prob, th)
    rv, av, rg, ag = (3e-5, 5e-3, 5e-3, 5e-2) if dtype == np.float32 else (1e-9, 1e-8, 1e-6, 1e-6)
    np.testing.assert_allclose(val, ov, rtol=rv, atol=av)
    np.testing.assert_allclose(v_only, val, rtol=1e-7, atol=1e-6)
    np.testing.assert_allclose(grad, og, rtol=rg, atol=ag)


def test_llt_config3_shape_streaming(engine):
  """BASELINE config 3 shape (T=5000, 50 covariates): tiles stream through the
  mbarrier ring (1 MB of [X|y] does not fit in shared memory)."""
  y, X, _ = make_series(5000, 50, 2023)
  prob = K.default_problem(y, X, model=K.MODEL_LOCAL_LINEAR_TREND)
  th = make_thetas(prob.dim, prob.p, 6, 8, d=2)
  # float32: the configuration itself (54 KB tiles, 3-stage ring)
  spec32 = cib.build_problem(y, X, model=LLT, dtype=np.float32)
  engine.set_data(spec32)
  th32 = th.astype(np.float32).astype(np.float64)
  v32, g32 = engine.logprob_grad(th32, with_prior=True)
  ov, og = K.log_post_grad(prob, th32)
  np.testing.assert_allclose(v32, ov, rtol=5e-5, atol=5e-2)
  np.testing.assert_allclose(g32, og, rtol=2e-2, atol=0.5)
  # float64: 109 KB tiles -> single-stage ring (correct, no copy/compute overlap)
  spec = cib.build_problem(y, X, model=LLT, dtype=np.float64)
  engine.set_data(spec)
  val, grad = engine.logprob_grad(th, with_prior=True)
  ov, og = K.log_post_grad(prob, th)
  np.testing.assert_allclose(val, ov, rtol=1e-9, atol=1e-7)
  np.testing.assert_allclose(grad, og, rtol=1e-6, atol=1e-5)


def test_llt_hmc_float64_pathwise(engine):
  y, X, _ = make_series(100, 1, 31)
  spec = cib.build_problem(y, X, prior_level_sd=0.05, model=LLT, dtype=np.float64)
  engine.set_data(spec)
  prob = K.default_problem(y, X, prior_level_sd=0.05, model=K.MODEL_LOCAL_LINEAR_TREND)
  th0 = np.tile(cib.initial_theta(spec, 0.05), (4, 1))
  th0[:, :spec.p] += 0.1 * np.random.default_rng(0).normal(size=(4, spec.p))
  kw = dict(n_warmup=25, n_results=8, seed=3, max_leapfrog=4, init_step=0.01)
  draws, stats = engine.hmc_run(th0, **kw)
  od, ost = H.run(lambda t: K.log_post_grad(prob, t), th0, **kw)
  np.testing.assert_allclose(draws, od, rtol=1e-6, atol=1e-6)
  assert np.array_equal(stats["n_leapfrog"], ost["n_leapfrog"])


@pytest.mark.parametrize("T,n_cov", [(100, 1), (700, 3), (260, 0)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_llt_posterior_predict_pathwise(engine, T, n_cov, dtype):
  """d = 2 simulation smoother + predictive draws vs oracle/smoother_np.posterior_predict_llt with
  the SAME Philox normals: level / trajectory paths agree draw by draw (float32: 5e-3 of the
  series' unit scale; float64: 1e-7), the mean is the average of level + X.w."""
  from oracle import smoother_np as SM
  y, X, _ = make_series(T, n_cov, 70 + T, nan_frac=0.03)
  spec = cib.build_problem(y, X, prior_level_sd=0.05, model=LLT, dtype=dtype)
  engine.set_data(spec)
  prob = K.default_problem(y, X, prior_level_sd=0.05, model=K.MODEL_LOCAL_LINEAR_TREND)
  th = make_thetas(spec.dim, spec.p, 6, 3, d=2).astype(dtype).astype(np.float64)
  th[:, spec.p + 1] = np.log(0.05 ** 2) + 0.3 * np.arange(6)        # a range of level / slope scales
  th[:, spec.p + 2] = np.log(0.01 ** 2) + 0.5 * np.arange(6)
  level, traj, mean = engine.posterior_predict(th, seed=21, draw_id0=9)
  ol, _, ot, om = SM.posterior_predict_llt(prob, th, 21, 9)
  tol = 5e-3 if dtype == np.float32 else 1e-7
  np.testing.assert_allclose(level, ol, rtol=tol, atol=tol)
  np.testing.assert_allclose(traj, ot, rtol=tol, atol=tol)
  np.testing.assert_allclose(mean, om, rtol=tol, atol=tol)
  # draws are keyed by their global id: a split batch reproduces the rows bit for bit
  l2, t2, _ = engine.posterior_predict(th[2:5], seed=21, draw_id0=11)
  assert np.array_equal(l2, level[2:5]) and np.array_equal(t2, traj[2:5])


def test_llt_config3_full_size_1024_chains(engine):
  """BASELINE configs[2] at FULL size (T=5000, 50 covariates, 1024 chains): finite, a subset
  equals the float64 oracle, and a chain's result does not depend on its position in the batch."""
  y, X, _ = make_series(5000, 50, 20243)
  spec = cib.build_problem(y, X, model=LLT)
  engine.set_data(spec)
  th = make_thetas(spec.dim, spec.p, 1024, 4, d=2).astype(np.float32).astype(np.float64)
  val, grad = engine.logprob_grad(th, with_prior=True)
  assert np.all(np.isfinite(val)) and np.all(np.isfinite(grad))
  prob = K.default_problem(y, X, model=K.MODEL_LOCAL_LINEAR_TREND)
  sub = np.arange(0, 1024, 171)
  from oracle import c_port
  ov, og, _ = c_port.logpost_grad(prob, th[sub])
  np.testing.assert_allclose(val[sub], ov, rtol=5e-5, atol=5e-2)
  np.testing.assert_allclose(grad[sub], og, rtol=2e-2, atol=0.5)
  perm = np.random.default_rng(0).permutation(1024)
  v2, g2 = engine.logprob_grad(th[perm], with_prior=True)
  assert np.array_equal(v2, val[perm]) and np.array_equal(g2, grad[perm])


def test_fit_causalimpact_with_local_linear_trend():
  """configs[2] end to end through the reference's entry point: a series with a steady drift the
  covariate does not explain.  The local-level model has to absorb the drift in its level noise;
  the trend extension extrapolates it, so its counterfactual keeps rising over the post-period
  and the estimated effect is the planted one."""
  import pandas as pd
  rng = np.random.default_rng(3)
  n = 240
  x = 100 + np.cumsum(rng.normal(size=n)) * 0.2
  y = 0.8 * x + 0.08 * np.arange(n) + rng.normal(size=n) * 0.3
  y[170:] += 4.0
  df = pd.DataFrame({"y": y, "x": x})
  kw = dict(seed=5, inference_options=cib.InferenceOptions(num_results=600))
  res = cib.fit_causalimpact(df, (0, 169), (170, n - 1),
                             engine_options=cib.EngineOptions(local_linear_trend=True), **kw)
  assert res.posterior_samples.slope_scale is not None
  assert res.posterior_samples.slope_scale.shape == (600,)
  assert res.posterior_samples.level.shape == (600, n)
  eff = res.summary.loc["average", "abs_effect"]
  lo, hi = res.summary.loc["average", "abs_effect_lower"], res.summary.loc["average", "abs_effect_upper"]
  assert lo < 4.0 < hi and abs(eff - 4.0) < 1.5, (eff, lo, hi)
  # the counterfactual keeps the pre-period drift
  post = res.series["posterior_mean"].values[170:]
  assert post[-1] - post[0] > 0.5 * 0.08 * (n - 171) * 0.5
  with pytest.raises(NotImplementedError):
    cib.fit_causalimpact(df, (0, 169), (170, n - 1), seed=1,
                         model_options=cib.ModelOptions(seasons=[cib.Seasons(num_seasons=7)]),
                         engine_options=cib.EngineOptions(local_linear_trend=True))
