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
