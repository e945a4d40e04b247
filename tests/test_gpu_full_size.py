"""BASELINE.json configs at FULL size, checked through size-independent
properties (the float64 oracle is only run on a small subset of chains)."""
import numpy as np
import pytest

import causalimpact_b200 as cib
from conftest import make_series, make_thetas
from oracle import kalman_np as K

pytestmark = pytest.mark.gpu


def test_config4_long_series_512_chains(engine):
  """T=20000, 512 chains, tiles streamed through the mbarrier ring."""
  y, X, _ = make_series(20000, 1, 20244)
  spec = cib.build_problem(y, X)
  engine.set_data(spec)
  th = make_thetas(spec.dim, spec.p, 512, 4).astype(np.float32).astype(np.float64)
  val, grad = engine.logprob_grad(th, with_prior=True)
  assert np.all(np.isfinite(val)) and np.all(np.isfinite(grad))
  prob = K.default_problem(y, X)
  sub = np.arange(0, 512, 37)
  ov, og = K.log_post_grad(prob, th[sub])
  np.testing.assert_allclose(val[sub], ov, rtol=2e-5, atol=5e-2)
  np.testing.assert_allclose(grad[sub], og, rtol=5e-3, atol=5e-2)
  # permutation invariance: a chain's result does not depend on its position in the batch
  perm = np.random.default_rng(0).permutation(512)
  v2, g2 = engine.logprob_grad(th[perm], with_prior=True)
  assert np.array_equal(v2, val[perm]) and np.array_equal(g2, grad[perm])


def test_config5_forecast_10000_draws(engine):
  """10000-draw posterior forecast at T=2000 (the per-GPU work of config 5 is
  1250 draws; here all 10000 on one GPU) + per-time quantiles."""
  y, X, _ = make_series(2000, 10, 20245)
  spec = cib.build_problem(y, X)
  engine.set_data(spec)
  rng = np.random.default_rng(1)
  th = np.tile(make_thetas(spec.dim, spec.p, 1, 6), (10000, 1))
  th[:, :spec.p] += 0.01 * rng.normal(size=(10000, spec.p))
  level, traj, mean = engine.posterior_predict(th, seed=11)
  assert level.shape == traj.shape == (10000, 2000) and np.all(np.isfinite(traj))
  q = engine.row_quantiles(traj, [0.025, 0.5, 0.975])
  assert q.shape == (2000, 3)
  assert np.all(q[:, 0] <= q[:, 1]) and np.all(q[:, 1] <= q[:, 2])          # sortedness
  loc = level + th[:, :spec.p].astype(np.float32) @ X.T.astype(np.float32)
  np.testing.assert_allclose(mean, loc.mean(0), atol=2e-3)
  # the median of the predictive draws tracks the mean of the noise-free predictive
  assert np.max(np.abs(q[:, 1] - mean)) < 0.05
  # coverage of the 95 % band by the draws themselves
  inside = ((traj >= q[:, 0]) & (traj <= q[:, 2])).mean()
  assert abs(inside - 0.95) < 0.002
  # first 1250 draws == what rank 0 of an 8-GPU job computes (global draw ids 0..1249)
  l8, t8, _ = engine.posterior_predict(th[:1250], seed=11, draw_id0=0)
  assert np.array_equal(t8, traj[:1250]) and np.array_equal(l8, level[:1250])


def test_config2_hmc_256_chains_full(engine):
  """configs[1]: 256 chains, T=1000, 10 covariates -- a short HMC run must be finite,
  adapt, and agree in distribution across halves of the chain batch."""
  y, X, _ = make_series(1000, 10, 20242)
  spec = cib.build_problem(y, X)
  engine.set_data(spec)
  th0 = np.tile(cib.initial_theta(spec), (256, 1))
  th0[:, :spec.p] = make_thetas(spec.dim, spec.p, 256, 2)[:, :spec.p]
  draws, stats = engine.hmc_run(th0, n_warmup=300, n_results=40, seed=3, init_step=0.01)
  assert np.all(np.isfinite(draws))
  assert stats["n_divergent"].sum() <= 0.02 * 256 * 40
  assert 0.55 < stats["accept_rate"].mean() < 0.98
  a = draws[:, :128].reshape(-1, spec.dim); b = draws[:, 128:].reshape(-1, spec.dim)
  for j in (spec.p, spec.p + 1):
    se = np.sqrt(a[:, j].var() / 300 + b[:, j].var() / 300)
    assert abs(a[:, j].mean() - b[:, j].mean()) < 6 * se + 1e-3
