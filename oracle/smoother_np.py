"""Simulation smoother (FFBS) + one-step predictive draw, float64 NumPy.
TEST INFRASTRUCTURE ONLY.

Restates, for the local-level model:
* the latent resampling inside the reference's sampler -- TFP
  ``LinearGaussianStateSpaceModel.posterior_sample`` called from
  ``gibbs_sampler._resample_latents`` (call site causalimpact_lib.py:365-388).
  TFP uses the Durbin-Koopman mean-correction smoother; forward-filter /
  backward-sample draws from exactly the same conditional law
  p(level_{0:T-1} | y, theta), which is what is restated here (and what the
  CUDA kernel does, because its backward recursion is an affine scan).
* ``_get_posterior_means_and_trajectories`` (causalimpact_lib.py:609-632):
  loc = level + X.w (zero-step prediction), scale = sigma_obs; the mean is the
  average of loc over draws (:627), one Normal sample per (draw, t) (:629-631).

Parity status: unpinned against the reference (needs TFP); pinned to the
analytic smoother moments by tests/test_oracle_smoother.py.
"""
import numpy as np

from oracle import philox_np as PH


def filtered_moments(r, mask, s_e, s_h, m0, P0):
  T = r.shape[0]
  m = np.empty(T); Cv = np.empty(T)
  a, P = m0, P0
  for t in range(T):
    if not mask[t]:
      F = P + s_e; K = P / F
      a = a + K * (r[t] - a); P = P * (1 - K)
    m[t], Cv[t] = a, P
    P = P + s_h
  return m, Cv


def ffbs_path(m, Cv, s_h, z):
  """Backward sampling given filtered moments and standard normals z[T]."""
  T = m.shape[0]
  x = np.empty(T)
  x[T - 1] = m[T - 1] + np.sqrt(Cv[T - 1]) * z[T - 1]
  for t in range(T - 2, -1, -1):
    J = Cv[t] / (Cv[t] + s_h)
    x[t] = m[t] + J * (x[t + 1] - m[t]) + np.sqrt(Cv[t] * (1 - J)) * z[t]
  return x


def posterior_predict(prob, theta_draws, seed, draw_id0=0):
  """Returns level [S,T], traj [S,T], mean [T] using the engine's RNG streams."""
  th = np.atleast_2d(np.asarray(theta_draws, np.float64))
  S, T, p = th.shape[0], prob.T, prob.p
  level = np.empty((S, T)); traj = np.empty((S, T)); loc = np.empty((S, T))
  mask = prob.mask
  for s in range(S):
    w = th[s, :p]; s_e = np.exp(th[s, p]); s_h = np.exp(th[s, p + 1])
    xw = prob.X @ w if p else np.zeros(T)
    r = prob.y - xw
    m, Cv = filtered_moments(r, mask, s_e, s_h, prob.m0, prob.P0)
    zs, zp = PH.predict_normals(seed, draw_id0 + s, T)
    level[s] = ffbs_path(m, Cv, s_h, zs)
    loc[s] = level[s] + xw
    traj[s] = loc[s] + np.sqrt(s_e) * zp
  return level, traj, loc.mean(axis=0)


def smoother_moments_dense(prob, theta_row):
  """Analytic E[level | y], Cov[level | y] from the dense joint Gaussian (small T)."""
  th = np.asarray(theta_row, np.float64)
  T, p = prob.T, prob.p
  s_e, s_h = np.exp(th[p]), np.exp(th[p + 1])
  xw = prob.X @ th[:p] if p else np.zeros(T)
  r = prob.y - xw
  tt = np.arange(T)
  Sx = prob.P0 + s_h * np.minimum(tt[:, None], tt[None, :])     # Cov(level)
  obs = ~prob.mask
  Syy = Sx[np.ix_(obs, obs)] + s_e * np.eye(obs.sum())
  Sxy = Sx[:, obs]
  G = Sxy @ np.linalg.inv(Syy)
  mean = prob.m0 + G @ (r[obs] - prob.m0)
  cov = Sx - G @ Sxy.T
  return mean, cov
