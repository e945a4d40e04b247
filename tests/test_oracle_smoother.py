"""Pins the FFBS oracle and the Philox restatement (CPU only)."""
import numpy as np

from conftest import make_series
from oracle import kalman_np as K
from oracle import philox_np as PH
from oracle import smoother_np as SM


def test_philox_known_answer():
  # Random123 known-answer vectors for philox4x32-10
  out = PH.philox4x32(0, 0, 0, 0, 0)
  assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
  seed = (0xffffffff << 32) | 0xffffffff
  out = PH.philox4x32(seed, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
  assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
  seed = (0x299f31d0 << 32) | 0xa4093822
  out = PH.philox4x32(seed, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)
  assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_normals_are_standard():
  zs, zp = PH.predict_normals(123, 7, 200000)
  for z in (zs, zp):
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
  assert abs(np.corrcoef(zs, zp)[0, 1]) < 0.01
  assert abs(np.corrcoef(zs[:-1], zs[1:])[0, 1]) < 0.01


def test_ffbs_moments_match_dense_smoother():
  y, X, _ = make_series(40, 1, 3, nan_frac=0.05)
  prob = K.default_problem(y, X, prior_level_sd=0.3)
  th = np.array([0.5, 0.1, np.log(0.2), np.log(0.05)])
  S = 20000
  level, traj, mean = SM.posterior_predict(prob, np.tile(th, (S, 1)), seed=5)
  mu, cov = SM.smoother_moments_dense(prob, th)
  se = np.sqrt(np.diag(cov) / S)
  assert np.all(np.abs(level.mean(0) - mu) < 5 * se + 1e-9)
  emp = np.cov(level.T)
  assert np.max(np.abs(emp - cov)) < 0.05 * np.max(np.abs(cov)) + 5e-3
  # predictive: loc + sigma_obs noise
  xw = prob.X @ th[:2]
  np.testing.assert_allclose(mean, level.mean(0) + xw, atol=1e-12)
  resid_var = np.var(traj - level - xw)
  assert abs(resid_var - 0.2) < 0.01


def test_llt_ffbs_mean_path_is_the_dense_smoother_mean():
  """Local linear trend: with zero normals the backward recursion returns the RTS smoothed
  means, which must equal the conditional mean of the dense joint Gaussian (exact); with real
  normals the draws have the dense conditional covariance (statistical)."""
  from conftest import make_series, make_thetas
  from oracle import kalman_np as K
  from oracle import smoother_np as SM
  y, X, _ = make_series(40, 2, 3, nan_frac=0.1)
  prob = K.default_problem(y, X, prior_level_sd=0.2, model=K.MODEL_LOCAL_LINEAR_TREND)
  th = make_thetas(prob.dim, prob.p, 1, 4, d=2)[0]
  th[prob.p + 1] = np.log(0.05 ** 2); th[prob.p + 2] = np.log(0.02 ** 2)
  s_e, q1, q2 = np.exp(th[prob.p]), np.exp(th[prob.p + 1]), np.exp(th[prob.p + 2])
  r = prob.y - prob.X @ th[:prob.p]
  m, C = SM.filtered_moments_llt(r, prob.mask, s_e, q1, q2, [prob.m0, prob.m0_slope],
                                 np.diag([prob.P0, prob.P0_slope]))
  mean, cov = SM.smoother_moments_dense_llt(prob, th)
  x0 = SM.ffbs_path_llt(m, C, q1, q2, np.zeros((prob.T, 2)))
  np.testing.assert_allclose(x0[:, 0], mean, rtol=1e-8, atol=1e-9)
  rng = np.random.default_rng(0)
  draws = np.stack([SM.ffbs_path_llt(m, C, q1, q2, rng.normal(size=(prob.T, 2)))[:, 0]
                    for _ in range(4000)])
  np.testing.assert_allclose(draws.mean(0), mean, atol=5 * np.sqrt(np.diag(cov).max() / 4000))
  emp = np.cov(draws.T)
  for i, j in ((0, 0), (5, 5), (5, 6), (20, 30), (39, 39)):
    assert abs(emp[i, j] - cov[i, j]) < 0.12 * np.sqrt(cov[i, i] * cov[j, j]) + 1e-6, (i, j)
  # the engine's streams: deterministic, draw-id keyed
  l1, s1, t1, mu1 = SM.posterior_predict_llt(prob, np.stack([th, th]), seed=9, draw_id0=3)
  l2, _, _, _ = SM.posterior_predict_llt(prob, th[None], seed=9, draw_id0=4)
  assert np.array_equal(l1[1], l2[0]) and not np.array_equal(l1[0], l1[1])
