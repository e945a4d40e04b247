"""NumPy restatement of the reference's Gibbs sampler.  TEST INFRASTRUCTURE.

The reference fits with ``gibbs_sampler.fit_with_gibbs_sampling`` from
tensorflow_probability.experimental.sts_gibbs (call site
causalimpact/causalimpact_lib.py:365-388; model lib.py:398-500; initial state
lib.py:566-581).  TFP is not vendored and not installable here, so the sweep is
restated from the published algorithm (Scott & Varian 2014 spike-and-slab
regression + FFBS + conjugate InverseGamma draws), in the order TFP's
``_build_sampler_loop_body`` runs them:

  1. (covariates) targets = y - level on observed steps;
     (sigma_obs^2, w) ~ conjugate Normal-InverseGamma with slab
     w | sigma^2 ~ N(0, sigma^2 Omega^-1)  and every feature included
     (inclusion probability min(1, 3/p) = 1 for p <= 3, lib.py:449-450; for
     p > 3 the reference additionally samples inclusion indicators; restated
     behind ``sparse=True`` from Scott & Varian's marginal, feature order and
     TFP's `experimental_use_weight_adjustment` are unknown -- see DESIGN.md
     "sampler choice").
  2. level ~ p(level | y - Xw, sigma's)   (FFBS, oracle/smoother_np.py)
  3. sigma_level^2 ~ InvGamma(conc + (T-1)/2, scale + 1/2 sum (d level)^2)
  4. (no covariates) sigma_obs^2 ~ InvGamma(conc + n/2, scale + 1/2 SSE)
  ``upper_bound`` clamps (lib.py:432, 442-443) are applied as min(scale, ub).

Parity status: unpinned against TFP (see oracle/__init__.py).  Its stationary
law, with level integrated out, is exactly oracle/kalman_np.log_post (up to
the rarely-binding clamps), which is what tests/test_oracle_samplers.py checks
against oracle/hmc_np.py.
"""
from __future__ import annotations

import numpy as np

from oracle import smoother_np as SM


def _log_marginal(gamma, XtX, Xty, yty, Omega, n_obs, conc0, scale0):
  """log p(gamma | targets) up to a constant, Scott & Varian (2014) eq. (6) with
  b = 0: |Omega_g|^{1/2} |Lambda_g|^{-1/2} (scale0 + SS_g / 2)^{-(conc0 + n/2)}."""
  idx = np.flatnonzero(gamma)
  if idx.size == 0:
    return -(conc0 + 0.5 * n_obs) * np.log(scale0 + 0.5 * yty)
  Om = Omega[np.ix_(idx, idx)]
  Lam = XtX[np.ix_(idx, idx)] + Om
  Lc = np.linalg.cholesky(Lam)
  z = np.linalg.solve(Lc, Xty[idx])
  ss = yty - z @ z
  return (0.5 * np.linalg.slogdet(Om)[1] - np.sum(np.log(np.diag(Lc)))
          - (conc0 + 0.5 * n_obs) * np.log(scale0 + 0.5 * ss))


def run(prob, *, n_results, n_warmup, seed, prior_level_sd=0.01, sparse=False):
  """Single chain, like the reference.  Returns dict of stacked draws.

  sparse=True adds the spike-and-slab inclusion step the reference uses when
  p > 3 (inclusion probability min(1, 3/p), lib.py:449-450): one Gibbs pass over
  the features, each indicator drawn from its conditional given the others
  (stochastic search variable selection), then sigma^2 and the active weights
  from their conjugate conditionals; inactive weights are exactly 0."""
  from oracle.kalman_np import initial_theta
  rng = np.random.Generator(np.random.PCG64(seed))
  T, p = prob.T, prob.p
  obs = ~prob.mask
  n_obs = int(obs.sum())
  th0 = initial_theta(prob, prior_level_sd)
  s_e, s_h = np.exp(th0[p]), np.exp(th0[p + 1])
  w = np.zeros(p)
  level = np.zeros(T)
  y0 = np.where(obs, prob.y, 0.0)
  if p:
    Xo = prob.X[obs]
    Lam = Xo.T @ Xo + prob.Omega
    Lam_chol = np.linalg.cholesky(Lam)
  out = dict(w=[], s_e=[], s_h=[], level=[])
  if p and sparse:
    XtX = Xo.T @ Xo
    pi = min(1.0, 3.0 / p)
    logit_pi = np.log(pi) - np.log1p(-pi) if pi < 1 else np.inf
    gamma = np.zeros(p, bool)                    # initial weights are all zero (lib.py:575-578)
  for it in range(n_warmup + n_results):
    if p and sparse:
      targ = (y0 - level)[obs]
      Xty = Xo.T @ targ
      yty = float(targ @ targ)
      for j in rng.permutation(p):
        g1 = gamma.copy(); g1[j] = True
        g0 = gamma.copy(); g0[j] = False
        l1 = _log_marginal(g1, XtX, Xty, yty, prob.Omega, n_obs, prob.obs_conc, prob.obs_scale)
        l0 = _log_marginal(g0, XtX, Xty, yty, prob.Omega, n_obs, prob.obs_conc, prob.obs_scale)
        d = l1 - l0 + logit_pi
        gamma[j] = rng.random() < 1.0 / (1.0 + np.exp(-d))
      idx = np.flatnonzero(gamma)
      w = np.zeros(p)
      if idx.size:
        Lg = XtX[np.ix_(idx, idx)] + prob.Omega[np.ix_(idx, idx)]
        Lc = np.linalg.cholesky(Lg)
        wbar = np.linalg.solve(Lg, Xty[idx])
        sse = yty - wbar @ Lg @ wbar
      else:
        sse = yty
      s_e = 1.0 / rng.gamma(prob.obs_conc + 0.5 * n_obs, 1.0 / (prob.obs_scale + 0.5 * sse))
      s_e = min(s_e, prob.ub_var(prob.obs_ub))
      if idx.size:
        w[idx] = wbar + np.sqrt(s_e) * np.linalg.solve(Lc.T, rng.normal(size=idx.size))
      r = prob.y - prob.X @ w
    elif p:
      targ = (y0 - level)[obs]
      b = Xo.T @ targ
      wbar = np.linalg.solve(Lam, b)
      sse = float(targ @ targ - wbar @ Lam @ wbar)
      s_e = 1.0 / rng.gamma(prob.obs_conc + 0.5 * n_obs, 1.0 / (prob.obs_scale + 0.5 * sse))
      s_e = min(s_e, prob.ub_var(prob.obs_ub))
      w = wbar + np.sqrt(s_e) * np.linalg.solve(Lam_chol.T, rng.normal(size=p))
      r = prob.y - prob.X @ w
    else:
      r = prob.y
    m, Cv = SM.filtered_moments(r, prob.mask, s_e, s_h, prob.m0, prob.P0)
    level = SM.ffbs_path(m, Cv, s_h, rng.normal(size=T))
    dl = np.diff(level)
    s_h = 1.0 / rng.gamma(prob.lvl_conc + 0.5 * (T - 1), 1.0 / (prob.lvl_scale + 0.5 * dl @ dl))
    s_h = min(s_h, prob.ub_var(prob.lvl_ub))
    if not p:
      e = (prob.y - level)[obs]
      s_e = 1.0 / rng.gamma(prob.obs_conc + 0.5 * n_obs, 1.0 / (prob.obs_scale + 0.5 * e @ e))
      s_e = min(s_e, prob.ub_var(prob.obs_ub))
    if it >= n_warmup:
      out["w"].append(w.copy()); out["s_e"].append(s_e); out["s_h"].append(s_h)
      out["level"].append(level.copy())
  return {k: np.asarray(v) for k, v in out.items()}
