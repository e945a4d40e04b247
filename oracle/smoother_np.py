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


# --------------------------------------------------------------------------------------------
# Local linear trend (d = 2): the predictive path of BASELINE.json configs[2].  The reference has
# no slope (causalimpact_lib.py:496); same conventions as oracle/kalman_np.gen_filter:
# state (level, slope), A = [[1,1],[0,1]], h = (1,0), Q = diag(q1, q2), prior at t = 0.
# --------------------------------------------------------------------------------------------
def filtered_moments_llt(r, mask, s_e, q1, q2, a0, P0):
  """Filtered means m [T,2] and covariances C [T,2,2] (after the update at each step)."""
  T = r.shape[0]
  A = np.array([[1.0, 1.0], [0.0, 1.0]])
  Q = np.diag([q1, q2])
  m = np.empty((T, 2)); C = np.empty((T, 2, 2))
  a = np.asarray(a0, float).copy(); P = np.asarray(P0, float).copy()
  for t in range(T):
    if not mask[t]:
      F = P[0, 0] + s_e
      K = P[:, 0] / F
      a = a + K * (r[t] - a[0])
      P = P - np.outer(K, P[0, :])
    m[t], C[t] = a, P
    a = A @ a
    P = A @ P @ A.T + Q
  return m, C


def _chol2(V):
  l00 = np.sqrt(max(V[0, 0], 0.0))
  l10 = V[0, 1] / l00 if l00 > 0 else 0.0
  l11 = np.sqrt(max(V[1, 1] - l10 * l10, 0.0))
  return np.array([[l00, 0.0], [l10, l11]])


def ffbs_path_llt(m, C, q1, q2, z):
  """Backward sampling of the (level, slope) path given filtered moments and normals z [T,2]."""
  T = m.shape[0]
  A = np.array([[1.0, 1.0], [0.0, 1.0]])
  Q = np.diag([q1, q2])
  x = np.empty((T, 2))
  x[T - 1] = m[T - 1] + _chol2(C[T - 1]) @ z[T - 1]
  for t in range(T - 2, -1, -1):
    R = A @ C[t] @ A.T + Q
    J = C[t] @ A.T @ np.linalg.inv(R)
    V = C[t] - J @ R @ J.T
    x[t] = m[t] + J @ (x[t + 1] - A @ m[t]) + _chol2(0.5 * (V + V.T)) @ z[t]
  return x


def predict_normals_llt(seed, draw_id, T):
  """(z_state [T,2], z_pred [T]) of draw `draw_id`: one Philox call per step (counter word c3 = 1
  separates the stream from the local-level one)."""
  x0, x1, x2, x3 = PH.philox4x32(seed, np.uint64(draw_id) & PH.MASK, PH._c1(PH.RNG_SMOOTH, draw_id),
                                 np.arange(T), 1)
  za, zb = PH.box_muller(x0, x1)
  zp, _ = PH.box_muller(x2, x3)
  return np.stack([za, zb], axis=1), zp


def posterior_predict_llt(prob, theta_draws, seed, draw_id0=0):
  """Local linear trend: level [S,T], slope [S,T], traj [S,T], mean [T] with the engine's streams."""
  th = np.atleast_2d(np.asarray(theta_draws, np.float64))
  S, T, p = th.shape[0], prob.T, prob.p
  level = np.empty((S, T)); slope = np.empty((S, T)); traj = np.empty((S, T)); loc = np.empty((S, T))
  for s in range(S):
    w = th[s, :p]; s_e, q1, q2 = np.exp(th[s, p]), np.exp(th[s, p + 1]), np.exp(th[s, p + 2])
    xw = prob.X @ w if p else np.zeros(T)
    r = prob.y - xw
    m, C = filtered_moments_llt(r, prob.mask, s_e, q1, q2, [prob.m0, prob.m0_slope],
                                np.diag([prob.P0, prob.P0_slope]))
    zs, zp = predict_normals_llt(seed, draw_id0 + s, T)
    x = ffbs_path_llt(m, C, q1, q2, zs)
    level[s], slope[s] = x[:, 0], x[:, 1]
    loc[s] = level[s] + xw
    traj[s] = loc[s] + np.sqrt(s_e) * zp
  return level, slope, traj, loc.mean(axis=0)


def smoother_moments_dense_llt(prob, theta_row):
  """Analytic E[level | y], Cov[level | y] of the local linear trend from the dense joint Gaussian
  (small T): level_t = l_0 + t s_0 + sum of level / slope innovations."""
  th = np.asarray(theta_row, np.float64)
  T, p = prob.T, prob.p
  s_e, q1, q2 = np.exp(th[p]), np.exp(th[p + 1]), np.exp(th[p + 2])
  xw = prob.X @ th[:p] if p else np.zeros(T)
  r = prob.y - xw
  # level_t = l0 + t * s0 + sum_{u<t} eta_u + sum_{u<t} (t - 1 - u) zeta_u
  tt = np.arange(T, dtype=float)
  Sx = prob.P0 + prob.P0_slope * np.outer(tt, tt)
  for u in range(T - 1):
    ind = (tt > u).astype(float)
    Sx += q1 * np.outer(ind, ind)
    lag = np.maximum(tt - 1 - u, 0.0)
    Sx += q2 * np.outer(lag, lag)
  mu = prob.m0 + prob.m0_slope * tt
  obs = ~prob.mask
  Syy = Sx[np.ix_(obs, obs)] + s_e * np.eye(obs.sum())
  G = Sx[:, obs] @ np.linalg.inv(Syy)
  return mu + G @ (r[obs] - mu[obs]), Sx - G @ Sx[:, obs].T
