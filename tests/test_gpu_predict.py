"""GPU parity for K4 (simulation smoother + predictive draw) and K5 (per-time
quantiles) through the C ABI.

K4 is compared PATHWISE: the oracle restates the engine's Philox streams, so
the same normals drive both.  Tolerances: float32 |d| <= 1e-3*|x| + 3e-3
(standardized scale, values O(1)); float64 1e-8.
K5 is compared with pandas DataFrame.quantile(axis=1) -- the very call the
reference makes (posterior_processing.py:56): float64 to 1e-13, float32 2e-6.
"""
import numpy as np
import pandas as pd
import pytest

import causalimpact_b200 as cib
from conftest import make_series, make_thetas
from oracle import kalman_np as K
from oracle import smoother_np as SM

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,n_cov,S", [(100, 1, 8), (257, 0, 5), (1000, 10, 40), (600, 40, 9)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_posterior_predict_matches_oracle_pathwise(engine, T, n_cov, S, dtype):
  y, X, _ = make_series(T, n_cov, 50 + T, nan_frac=0.02)
  spec = cib.build_problem(y, X, prior_level_sd=0.05, dtype=dtype)
  engine.set_data(spec)
  prob = K.default_problem(y, X, prior_level_sd=0.05)
  th = make_thetas(spec.dim, spec.p, S, 3).astype(dtype).astype(np.float64)
  th[:, spec.p + 1] += 2.0            # livelier level so the smoother matters
  level, traj, mean = engine.posterior_predict(th, seed=1234, draw_id0=17)
  ol, ot, om = SM.posterior_predict(prob, th, seed=1234, draw_id0=17)
  rt, at = (1e-3, 3e-3) if dtype == np.float32 else (1e-8, 1e-8)
  np.testing.assert_allclose(level, ol, rtol=rt, atol=at)
  np.testing.assert_allclose(traj, ot, rtol=rt, atol=at)
  np.testing.assert_allclose(mean, om, rtol=rt, atol=at)


def test_predict_independent_of_batch_split(engine):
  """Draw s depends only on (seed, global draw id): splitting the batch (as the
  multi-GPU sharding does) must give bit-identical rows."""
  y, X, _ = make_series(700, 3, 9)
  spec = cib.build_problem(y, X)
  engine.set_data(spec)
  th = make_thetas(spec.dim, spec.p, 24, 4)
  l_all, t_all, _ = engine.posterior_predict(th, seed=99, draw_id0=0)
  l_a, t_a, _ = engine.posterior_predict(th[:10], seed=99, draw_id0=0)
  l_b, t_b, _ = engine.posterior_predict(th[10:], seed=99, draw_id0=10)
  assert np.array_equal(l_all, np.concatenate([l_a, l_b]))
  assert np.array_equal(t_all, np.concatenate([t_a, t_b]))
  l2, t2, _ = engine.posterior_predict(th, seed=99, draw_id0=0)
  assert np.array_equal(l_all, l2) and np.array_equal(t_all, t2)   # run-to-run determinism


def test_predictive_moments_long_series(engine):
  """Property check at T=2000 / S=2000 (config-5 shape per GPU): traj - level - Xw
  is N(0, sigma_obs^2) noise and the mean excludes it (lib.py:717-722)."""
  y, X, _ = make_series(2000, 10, 2025)
  spec = cib.build_problem(y, X)
  engine.set_data(spec)
  th = np.tile(make_thetas(spec.dim, spec.p, 1, 5), (2000, 1))
  level, traj, mean = engine.posterior_predict(th, seed=7)
  th32 = th[0].astype(np.float32).astype(np.float64)
  xw = X @ th32[:spec.p]
  noise = traj - level - xw[None, :]
  s_e = np.exp(th32[spec.p])
  assert abs(noise.mean()) < 5 * np.sqrt(s_e / noise.size) + 1e-5
  assert abs(noise.var() / s_e - 1) < 0.01
  np.testing.assert_allclose(mean, level.mean(0) + xw, atol=2e-4)


@pytest.mark.parametrize("S,T", [(10, 7), (900, 91), (1000, 300), (10000, 64), (3, 2), (1, 4)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_row_quantiles_match_pandas(engine, S, T, dtype):
  rng = np.random.default_rng(S * 1000 + T)
  a = rng.normal(size=(S, T)).astype(dtype)
  if S >= 10:
    a[rng.integers(0, S, 5), rng.integers(0, T, 5)] = np.nan     # pandas skips NaN
    if T > 5:
      a[:, 0] = np.nan                                            # an all-NaN column
  q = np.array([0.025, 0.975, 0.5, 0.0, 1.0])
  out = engine.row_quantiles(a, q)
  want = pd.DataFrame(a.T.astype(np.float64)).quantile(q=list(q), axis=1).transpose().values
  tol = 1e-13 if dtype == np.float64 else 2e-6
  np.testing.assert_allclose(out, want, rtol=tol, atol=tol, equal_nan=True)


def test_row_quantiles_rejects_bad_input(engine):
  a = np.zeros((4, 4))
  with pytest.raises(cib.EngineError):
    engine.row_quantiles(a, [1.5])
  with pytest.raises(cib.EngineError):
    engine.row_quantiles(a, list(np.linspace(0, 1, 9)))          # at most 8 quantiles per call


def test_predict_team_and_single_warp_paths_agree(monkeypatch):
  """T=1000: the opt-in TEAM kernel (CI_B200_PREDICT_TEAM=1: one warp per tile, no
  replay) vs the default one-warp-per-draw kernel: same Philox normals, so the
  paths agree to float32 rounding and both match the oracle.  (The default never
  depends on the batch size, so draws are bit-identical under any split.)"""
  y, X, _ = make_series(1000, 10, 77, nan_frac=0.02)
  spec = cib.build_problem(y, X, prior_level_sd=0.05)
  prob = K.default_problem(y, X, prior_level_sd=0.05)
  th = make_thetas(spec.dim, spec.p, 37, 3).astype(np.float32).astype(np.float64)
  th[:, spec.p + 1] += 2.0
  out = {}
  for mode in ("1", "0"):
    monkeypatch.setenv("CI_B200_PREDICT_TEAM", mode)
    eng = cib.Engine(0)
    eng.set_data(spec)
    out[mode] = eng.posterior_predict(th, seed=5, draw_id0=100)
    eng.close()
  ol, ot, om = SM.posterior_predict(prob, th, seed=5, draw_id0=100)
  for mode in out:
    np.testing.assert_allclose(out[mode][0], ol, rtol=1e-3, atol=3e-3)
    np.testing.assert_allclose(out[mode][1], ot, rtol=1e-3, atol=3e-3)
    np.testing.assert_allclose(out[mode][2], om, rtol=1e-3, atol=3e-3)
  np.testing.assert_allclose(out["1"][1], out["0"][1], rtol=1e-4, atol=5e-4)


def test_row_quantiles_large_draw_counts(engine):
  """Columns that fit shared memory (float64 up to ~25k draws, float32 ~50.9k) are selected
  there; longer ones straight from global memory -- same exact results, no size limit.  Includes
  a column with heavy ties and one with an extreme outlier (both defeat the value binning and
  take the bisection fallback)."""
  rng = np.random.default_rng(0)
  a = rng.normal(size=(20000, 3))
  out = engine.row_quantiles(a, [0.025, 0.975])
  np.testing.assert_allclose(out, np.quantile(a, [0.025, 0.975], axis=0).T, rtol=1e-13)
  b = rng.normal(size=(60000, 4))                           # float64, global-memory path
  b[:, 2] = np.round(b[:, 2])                               # ties
  b[17, 3] = 1e12                                           # outlier: everything else in one bin
  b[5, 0] = np.nan
  out = engine.row_quantiles(b, [0.025, 0.5, 0.975])
  assert out.dtype == np.float64
  np.testing.assert_allclose(out, np.nanquantile(b, [0.025, 0.5, 0.975], axis=0).T, rtol=1e-13)
  c = rng.normal(size=(200000, 2)).astype(np.float32)       # float32, global-memory path
  out = engine.row_quantiles(c, [0.01, 0.5, 0.99])
  np.testing.assert_allclose(out, np.quantile(c.astype(np.float64), [0.01, 0.5, 0.99], axis=0).T,
                             rtol=2e-6, atol=2e-6)


def test_quantiles_of_standard_normal_draws(engine):
  """posterior_processing_test.py:26-45: the reference draws 1e7 N(0,1) per time point for 10
  time points and expects +-1.96 within 0.01.  Same assertion and tolerance with 2e6 draws per
  time point (MC sd of the quantile 0.002; 80 MB instead of 400 MB through PCIe)."""
  rng = np.random.default_rng(1)
  a = rng.standard_normal((2_000_000, 10), dtype=np.float32)
  q = engine.row_quantiles(a, [0.025, 0.975])
  assert q.shape == (10, 2)
  np.testing.assert_allclose(q[:, 0], -1.96, atol=0.01)
  np.testing.assert_allclose(q[:, 1], 1.96, atol=0.01)
