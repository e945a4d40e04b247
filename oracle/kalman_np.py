"""Float64 NumPy restatement of the Kalman log-prob path.  TEST INFRASTRUCTURE.

What it restates (reference file:line, relative to /root/reference):

* model + priors ............ causalimpact/causalimpact_lib.py:398-500
  (``_build_default_gibbs_model``): local level, InvGamma variance priors
  with ``upper_bound`` clamps (:424-443), slab precision
  ``0.01 * set_diag(0.5 X'X, diag(X'X)) / T`` over the FULL design matrix
  (:451-459), initial level ``N(y[0], sd)`` (:467-469), no slope (:496).
* mask extension / init ...... causalimpact/causalimpact_lib.py:548-581
* Kalman recursion ........... NOT in the reference tree.  It is TFP's
  ``tfd.LinearGaussianStateSpaceModel`` (tensorflow-probability, un-pinned:
  pyproject.toml:22).  Restated from the published algorithm with TFP's
  convention: the prior ``N(m0, P0)`` is on the state AT t=0; each step first
  *updates* with ``y_t`` (skipped when masked) and then *predicts*.

Parity status: **unpinned against the reference** (no TFP here, no golden
vectors in the reference for this path); pinned to ground truth by
tests/test_oracle_kalman.py (dense Gaussian marginal, finite differences,
steady-state gain).

theta layout (one row per chain), everything unconstrained:
    [ w_0 .. w_{p-1},  u = log sigma_obs^2,  l = log sigma_level^2
      (, s = log sigma_slope^2  when model == 1) ]
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

LOG2PI = float(np.log(2.0 * np.pi))

MODEL_LOCAL_LEVEL = 0
MODEL_LOCAL_LINEAR_TREND = 1


@dataclasses.dataclass
class Problem:
  """Plain parameter struct replacing the TFP model object (A2 in SURVEY §8a)."""
  model: int
  y: np.ndarray                 # [T] float64, NaN == missing (pre NaNs + whole post period)
  X: Optional[np.ndarray]       # [T, p] float64 or None
  Omega: Optional[np.ndarray]   # [p, p] slab precision (scaled by 1/sigma_obs^2)
  m0: float                     # initial level mean           (lib.py:467-469)
  P0: float                     # initial level variance = sd^2
  obs_conc: float               # InvGamma on sigma_obs^2      (lib.py:434-441)
  obs_scale: float
  obs_ub: float                 # the prior's `upper_bound` (lib.py:442-443); see ub_on_scale
  lvl_conc: float               # InvGamma on sigma_level^2    (lib.py:424-431)
  lvl_scale: float
  lvl_ub: float                 # the prior's `upper_bound` (lib.py:432)
  slope_conc: float = 16.0      # extension: no reference counterpart (lib.py:496)
  slope_scale: float = 0.0
  slope_ub: float = np.inf
  m0_slope: float = 0.0
  P0_slope: float = 1.0
  # what the *_ub limit: the reference puts `upper_bound` on InverseGamma priors over VARIANCES
  # and TFP clips the variance draw (min(variance, upper_bound)) -> False; True = the scale
  ub_on_scale: bool = False

  def ub_var(self, ub: float) -> float:
    """The bound expressed on the variance."""
    return ub * ub if self.ub_on_scale else ub

  @property
  def T(self) -> int:
    return int(self.y.shape[0])

  @property
  def p(self) -> int:
    return 0 if self.X is None else int(self.X.shape[1])

  @property
  def d(self) -> int:
    return 2 if self.model == MODEL_LOCAL_LINEAR_TREND else 1

  @property
  def dim(self) -> int:
    return self.p + 1 + self.d

  @property
  def mask(self) -> np.ndarray:
    return np.isnan(self.y)


def slab_precision(X_full: np.ndarray) -> np.ndarray:
  """Omega = 0.01 * (0.5 X'X + 0.5 diag(X'X)) / T      (lib.py:451-453)."""
  xtx = X_full.T @ X_full
  om = 0.5 * xtx
  om[np.diag_indices_from(om)] = np.diag(xtx)
  return 0.01 * om / X_full.shape[0]


def default_problem(y_ext: np.ndarray, X_full: Optional[np.ndarray], *,
                    prior_level_sd: float = 0.01, model: int = MODEL_LOCAL_LEVEL,
                    outcome_sd: Optional[float] = None) -> Problem:
  """Priors exactly as the reference builds them (lib.py:398-500, 563-572).

  ``y_ext`` is the standardized outcome over pre + after-pre with NaN for every
  masked step (lib.py:548-562).  ``outcome_sd`` = nanstd(pre y, ddof=1) (:563).
  """
  y_ext = np.asarray(y_ext, dtype=np.float64)
  obs = y_ext[~np.isnan(y_ext)]
  sd = float(np.std(obs, ddof=1)) if outcome_sd is None else float(outcome_sd)
  has_x = X_full is not None and X_full.shape[1] > 0
  level_scale0 = prior_level_sd * sd                      # lib.py:572
  # lib.py:467-469 uses y[0]; that is NaN when the first point is missing
  # (untested upstream) -- we fall back to the first observed value.
  m0 = float(obs[0]) if np.isnan(y_ext[0]) else float(y_ext[0])
  return Problem(
      model=model, y=y_ext,
      X=None if not has_x else np.asarray(X_full, dtype=np.float64),
      Omega=None if not has_x else slab_precision(np.asarray(X_full, np.float64)),
      m0=m0, P0=sd * sd,
      obs_conc=25.0 if has_x else 0.005,                  # lib.py:434-441
      obs_scale=(5.0 if has_x else 0.005) * sd * sd,
      obs_ub=1.2 * sd,                                    # lib.py:442-443
      lvl_conc=16.0, lvl_scale=16.0 * level_scale0 ** 2,  # lib.py:424-431
      lvl_ub=sd,                                          # lib.py:432
      slope_conc=16.0, slope_scale=16.0 * level_scale0 ** 2, slope_ub=sd,
      m0_slope=0.0, P0_slope=sd * sd)


def initial_theta(prob: Problem, prior_level_sd: float = 0.01) -> np.ndarray:
  """The reference's initial sampler state (lib.py:566-581) in theta coordinates."""
  sd = np.sqrt(prob.P0)
  th = np.zeros(prob.dim)
  sig_obs = np.sqrt(1.0 - 0.8) * sd if prob.p > 0 else sd
  th[prob.p] = np.log(sig_obs ** 2)
  th[prob.p + 1] = np.log((prior_level_sd * sd) ** 2)
  if prob.d == 2:
    th[prob.p + 2] = np.log((prior_level_sd * sd) ** 2)
  return th


def residuals(prob: Problem, W: np.ndarray) -> np.ndarray:
  """r[c, t] = y_t - x_t . w_c   (NaN where masked)."""
  C = W.shape[0]
  if prob.p == 0:
    return np.broadcast_to(prob.y, (C, prob.T)).copy()
  return prob.y[None, :] - W @ prob.X.T


# --------------------------------------------------------------------------
# Local level (d = 1): scalar recursion, batched over chains.
# --------------------------------------------------------------------------
def ll_filter(R, mask, s_e, s_h, m0, P0, return_path=False):
  """Forward filter.  R [C,T]; mask [T] bool; s_e, s_h [C] variances.

  Returns ll [C] (and the per-step path).  Convention: (a_t, P_t) are the
  predicted moments at t; update (if observed) then predict.
  """
  C, T = R.shape
  a = np.full(C, float(m0))
  P = np.full(C, float(P0))
  ll = np.zeros(C)
  if return_path:
    A_ = np.empty((T, C)); P_ = np.empty((T, C))
  for t in range(T):
    if return_path:
      A_[t] = a; P_[t] = P
    if not mask[t]:
      v = R[:, t] - a
      F = P + s_e
      K = P / F
      ll += -0.5 * (LOG2PI + np.log(F) + v * v / F)
      a = a + K * v
      P = P * (1.0 - K)
    P = P + s_h
  if return_path:
    return ll, A_, P_
  return ll


def ll_filter_grad(R, mask, s_e, s_h, m0, P0):
  """Value + reverse-mode gradient of the local-level filter log-likelihood.

  Returns ll [C], rbar [C,T] (d ll / d r_t, 0 where masked), d ll/d s_e [C],
  d ll/d s_h [C].
  """
  C, T = R.shape
  ll, A_, P_ = ll_filter(R, mask, s_e, s_h, m0, P0, return_path=True)
  abar = np.zeros(C); Pbar = np.zeros(C)
  g_e = np.zeros(C); g_h = np.zeros(C)
  rbar = np.zeros((C, T))
  for t in range(T - 1, -1, -1):
    g_h += Pbar                       # P_{t+1} = (...) + s_h
    if mask[t]:
      continue                        # a, P pass through
    a = A_[t]; P = P_[t]
    v = R[:, t] - a
    F = P + s_e
    K = P / F
    dF = -0.5 * (1.0 / F - v * v / (F * F))
    rbar[:, t] = K * abar - v / F
    g_e += K * K * Pbar - abar * v * P / (F * F) + dF
    Pbar_new = (1.0 - K) ** 2 * Pbar + abar * v * s_e / (F * F) + dF
    abar = (1.0 - K) * abar + v / F
    Pbar = Pbar_new
  return ll, rbar, g_e, g_h


# --------------------------------------------------------------------------
# General d-dimensional filter (used for the local linear trend, d = 2).
# Observation vector h = e_0; transition A; Q = diag(q).
# --------------------------------------------------------------------------
def _llt_mats():
  A = np.array([[1.0, 1.0], [0.0, 1.0]])
  h = np.array([1.0, 0.0])
  return A, h


def gen_filter(R, mask, s_e, Qdiag, a0, P0, A, h, return_path=False):
  """Generic filter.  R [C,T]; s_e [C]; Qdiag [C,d]; a0 [d]; P0 [d,d]."""
  C, T = R.shape
  d = A.shape[0]
  a = np.broadcast_to(np.asarray(a0, float), (C, d)).copy()
  P = np.broadcast_to(np.asarray(P0, float), (C, d, d)).copy()
  ll = np.zeros(C)
  if return_path:
    A_ = np.empty((T, C, d)); P_ = np.empty((T, C, d, d))
  for t in range(T):
    if return_path:
      A_[t] = a; P_[t] = P
    if not mask[t]:
      v = R[:, t] - a @ h
      Ph = P @ h
      F = Ph @ h + s_e
      K = Ph / F[:, None]
      ll += -0.5 * (LOG2PI + np.log(F) + v * v / F)
      a = a + K * v[:, None]
      P = P - K[:, :, None] * (np.einsum('i,cij->cj', h, P))[:, None, :]
    a = a @ A.T
    P = np.einsum('ij,cjk,lk->cil', A, P, A)
    idx = np.arange(d)
    P[:, idx, idx] += Qdiag
  if return_path:
    return ll, A_, P_
  return ll


def gen_filter_grad(R, mask, s_e, Qdiag, a0, P0, A, h):
  """Reverse-mode gradient of gen_filter: ll, rbar [C,T], d/ds_e [C], d/dQdiag [C,d]."""
  C, T = R.shape
  d = A.shape[0]
  ll, A_, P_ = gen_filter(R, mask, s_e, Qdiag, a0, P0, A, h, return_path=True)
  abar = np.zeros((C, d)); Pbar = np.zeros((C, d, d))
  g_e = np.zeros(C); g_q = np.zeros((C, d))
  rbar = np.zeros((C, T))
  idx = np.arange(d)
  for t in range(T - 1, -1, -1):
    # predict:  a' = A a+,  P' = A P+ A' + Q
    g_q += Pbar[:, idx, idx]
    abar_f = abar @ A                                  # A' abar'
    Pbar_f = np.einsum('ji,cjk,kl->cil', A, Pbar, A)   # A' Pbar' A
    if mask[t]:
      abar, Pbar = abar_f, Pbar_f
      continue
    a = A_[t]; P = P_[t]
    v = R[:, t] - a @ h
    Ph = P @ h
    hP = np.einsum('i,cij->cj', h, P)
    F = Ph @ h + s_e
    K = Ph / F[:, None]
    # a+ = a + K v
    abar_n = abar_f.copy()
    Kbar = abar_f * v[:, None]
    vbar = np.einsum('ci,ci->c', K, abar_f)
    # P+ = P - K (h'P)
    Pbar_n = Pbar_f - h[None, :, None] * np.einsum('ci,cij->cj', K, Pbar_f)[:, None, :]
    Kbar = Kbar - np.einsum('cij,cj->ci', Pbar_f, hP)
    # K = P h / F
    Pbar_n = Pbar_n + Kbar[:, :, None] * h[None, None, :] / F[:, None, None]
    Fbar = -np.einsum('ci,ci->c', Kbar, K) / F
    # ll_t
    vbar = vbar - v / F
    Fbar = Fbar - 0.5 * (1.0 / F - v * v / (F * F))
    # F = h'Ph + s_e
    Pbar_n = Pbar_n + Fbar[:, None, None] * np.outer(h, h)[None]
    g_e += Fbar
    # v = r - h.a
    rbar[:, t] = vbar
    abar_n = abar_n - vbar[:, None] * h[None, :]
    abar, Pbar = abar_n, Pbar_n
  return ll, rbar, g_e, g_q


# --------------------------------------------------------------------------
# The target: Kalman log-likelihood (+ log prior + Jacobians).
# --------------------------------------------------------------------------
def _unpack(prob: Problem, theta: np.ndarray):
  theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
  p = prob.p
  W = theta[:, :p]
  s_e = np.exp(theta[:, p])
  s_h = np.exp(theta[:, p + 1])
  s_z = np.exp(theta[:, p + 2]) if prob.d == 2 else None
  return theta, W, s_e, s_h, s_z


def in_support(prob: Problem, theta) -> np.ndarray:
  """upper_bound clamps of the reference (lib.py:432, 442-443) as a truncation."""
  theta, W, s_e, s_h, s_z = _unpack(prob, theta)
  ok = (s_e <= prob.ub_var(prob.obs_ub)) & (s_h <= prob.ub_var(prob.lvl_ub))
  if s_z is not None:
    ok &= s_z <= prob.ub_var(prob.slope_ub)
  return ok & np.all(np.isfinite(theta), axis=1)


def log_prior(prob: Problem, theta):
  """log prior + log|Jacobian| in theta coordinates, and its gradient."""
  theta, W, s_e, s_h, s_z = _unpack(prob, theta)
  p = prob.p
  u = theta[:, p]; l = theta[:, p + 1]
  g = np.zeros_like(theta)
  lp = -(prob.obs_conc + 1.0) * u - prob.obs_scale / s_e + u
  g[:, p] = -(prob.obs_conc + 1.0) + prob.obs_scale / s_e + 1.0
  lp += -(prob.lvl_conc + 1.0) * l - prob.lvl_scale / s_h + l
  g[:, p + 1] = -(prob.lvl_conc + 1.0) + prob.lvl_scale / s_h + 1.0
  if s_z is not None:
    s = theta[:, p + 2]
    lp += -(prob.slope_conc + 1.0) * s - prob.slope_scale / s_z + s
    g[:, p + 2] = -(prob.slope_conc + 1.0) + prob.slope_scale / s_z + 1.0
  if p > 0:
    Ow = W @ prob.Omega.T
    q = np.einsum('cj,cj->c', W, Ow)
    lp += -0.5 * p * u - 0.5 * q / s_e
    g[:, :p] = -Ow / s_e[:, None]
    g[:, p] += -0.5 * p + 0.5 * q / s_e
  return lp, g


def log_lik(prob: Problem, theta) -> np.ndarray:
  """The LGSSM log_prob of the observed series given theta (no priors)."""
  theta, W, s_e, s_h, s_z = _unpack(prob, theta)
  R = residuals(prob, W)
  if prob.d == 1:
    return ll_filter(R, prob.mask, s_e, s_h, prob.m0, prob.P0)
  A, h = _llt_mats()
  return gen_filter(R, prob.mask, s_e, np.stack([s_h, s_z], 1),
                    [prob.m0, prob.m0_slope], np.diag([prob.P0, prob.P0_slope]), A, h)


def log_lik_grad(prob: Problem, theta):
  theta, W, s_e, s_h, s_z = _unpack(prob, theta)
  p = prob.p
  R = residuals(prob, W)
  g = np.zeros_like(theta)
  if prob.d == 1:
    ll, rbar, g_e, g_h = ll_filter_grad(R, prob.mask, s_e, s_h, prob.m0, prob.P0)
    g[:, p] = g_e * s_e
    g[:, p + 1] = g_h * s_h
  else:
    A, h = _llt_mats()
    ll, rbar, g_e, g_q = gen_filter_grad(
        R, prob.mask, s_e, np.stack([s_h, s_z], 1), [prob.m0, prob.m0_slope],
        np.diag([prob.P0, prob.P0_slope]), A, h)
    g[:, p] = g_e * s_e
    g[:, p + 1] = g_q[:, 0] * s_h
    g[:, p + 2] = g_q[:, 1] * s_z
  if p > 0:
    g[:, :p] = -rbar @ prob.X          # r = y - Xw
  return ll, g


def log_post(prob: Problem, theta) -> np.ndarray:
  """Un-normalised log posterior in theta coordinates; -inf outside the bounds."""
  ll = log_lik(prob, theta)
  lp, _ = log_prior(prob, theta)
  out = ll + lp
  out[~in_support(prob, theta)] = -np.inf
  return out


def log_post_grad(prob: Problem, theta):
  ll, g = log_lik_grad(prob, theta)
  lp, gp = log_prior(prob, theta)
  val = ll + lp
  val[~in_support(prob, theta)] = -np.inf
  return val, g + gp


# --------------------------------------------------------------------------
# Independent ground truth (used only to pin the oracle itself).
# --------------------------------------------------------------------------
def dense_marginal_loglik(prob: Problem, theta_row) -> float:
  """log N(y_obs; mean, Sigma) with the dense marginal covariance (T <= ~400)."""
  from scipy.stats import multivariate_normal
  theta, W, s_e, s_h, s_z = _unpack(prob, theta_row)
  T = prob.T
  r = residuals(prob, W)[0]
  obs = ~prob.mask
  tt = np.arange(T)
  if prob.d == 1:
    # level_t = level_0 + sum_{k<=t, k>=1} eta_k ;  Var(level_0)=P0
    cov = prob.P0 + s_h[0] * np.minimum(tt[:, None], tt[None, :])
    mean = np.full(T, prob.m0)
  else:
    # mu_t = mu_0 + t*delta_0 + sum eta + sum_{j} (t-j) zeta_j terms
    cov = prob.P0 + prob.P0_slope * np.outer(tt, tt) \
        + s_h[0] * np.minimum(tt[:, None], tt[None, :])
    # slope noise: delta_k = delta_0 + sum_{j=1..k} zeta_j; mu_t = ... + sum_{k=0}^{t-1} delta_k
    # coefficient of zeta_j in mu_t is max(t - j, 0), j >= 1
    j = np.arange(1, T)
    Cz = np.maximum(tt[:, None] - j[None, :], 0).astype(float)
    cov = cov + s_z[0] * (Cz @ Cz.T)
    mean = prob.m0 + prob.m0_slope * tt
  cov = cov + s_e[0] * np.eye(T)
  return float(multivariate_normal.logpdf(r[obs], mean=mean[obs], cov=cov[np.ix_(obs, obs)],
                                           allow_singular=False))
