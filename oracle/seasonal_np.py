"""ORACLE (test infrastructure only): the reference's seasonal model and its Gibbs sweep,
float64 NumPy.  SURVEY section 8 row f3.

The reference adds one ``tfp.sts.Seasonal(num_seasons, num_steps_per_season,
allow_drift=True, constrain_mean_effect_to_zero=True, drift_scale_prior=sqrt of
InverseGamma(0.005, 5e-7 sd^2) [variance bounded by sd], initial_effect_prior=N(0, sd))``
per ``Seasons`` option (causalimpact/causalimpact_lib.py:471-489) to the Gibbs model
(:491-500); its sampler then draws level and seasonal latents JOINTLY with the LGSSM
simulation smoother and the drift scales from their InverseGamma conditionals
(call site :365-388; initial drift scale 0.01 sd, :573-574).

TFP is not vendored / installable here, so this file restates the PUBLISHED
model (tfp.sts.Seasonal docstring + Harvey 1989 dummy-seasonal form) -- "parity unpinned"
against TFP itself, pinned instead by exact linear-Gaussian identities
(tests/test_oracle_seasonal.py):

  * TFP's construction: a latent vector of the num_seasons effects, ROTATED every time a
    season ends so that the current season's effect is element 0 (observation picks element
    0); at a season end the just-finished effect receives N(0, drift^2) noise; with
    constrain_mean_effect_to_zero the effects are replaced by num_seasons-1 "residuals"
    through E2R = (I - 11'/n)[:-1] and R2E = pinv(E2R) (transition E2R.Perm.R2E, noise
    E2R Q E2R', prior E2R S E2R').  ``tfp_constrained_matrices`` builds exactly that.
  * The form the ENGINE uses ("effects space"): keep all n effects in fixed positions (no
    rotation, transition = identity), let the observation pick the current season, and
    project every random input onto the zero-sum subspace: prior sd^2 C, season-end noise
    drift^2 (C e_j)(C e_j)',  C = I - 11'/n.  Because R2E.E2R = C and C commutes with the
    rotation, both forms give the SAME law for y and for each season's contribution; the
    test checks the dense T x T marginal covariances agree to 1e-12.

The simulation smoother is Durbin & Koopman's (2002) mean-correction sampler with the
fast state smoother (Koopman 1993) -- the algorithm the CUDA kernel runs, and the one TFP's
``posterior_sample`` uses; it is checked against the dense Gaussian conditional.
"""
from __future__ import annotations

import dataclasses
from typing import List, Sequence

import numpy as np


# ---------------------------------------------------------------------------
# schedule: which season is active at step t, and whether it ends after step t
# ---------------------------------------------------------------------------
def season_schedule(num_seasons: int, num_steps_per_season, T: int):
  """(idx [T] season index at t, ends [T] True when the season is over after step t).

  ``num_steps_per_season``: int | [num_seasons] | [num_cycles, num_seasons]
  (reference Seasons docstring, causalimpact_lib.py:162-180); step 0 is the first step of
  season 0 of cycle 0; cycles repeat."""
  steps = np.asarray(num_steps_per_season, dtype=np.int64)
  if steps.ndim == 0:
    steps = np.full((1, num_seasons), int(steps))
  elif steps.ndim == 1:
    steps = steps.reshape(1, -1)
  if steps.ndim != 2 or steps.shape[1] != num_seasons or np.any(steps < 1):
    raise ValueError("num_steps_per_season must be an int, [num_seasons] or "
                     "[num_cycles, num_seasons] of positive ints")
  flat = steps.reshape(-1)
  idx = np.empty(T, np.int64); ends = np.zeros(T, bool)
  pos, left = 0, int(flat[0])
  for t in range(T):
    idx[t] = pos % num_seasons
    left -= 1
    if left == 0:
      ends[t] = True
      pos = (pos + 1) % flat.size
      left = int(flat[pos])
  return idx, ends


@dataclasses.dataclass
class SeasonalSpec:
  """Effects-space description of K seasonal components on T steps."""
  n: List[int]                 # num_seasons per component
  idx: np.ndarray              # [K, T] active season
  ends: np.ndarray             # [K, T] season ends after step t
  init_sd: float               # initial_effect_prior scale (lib.py:489: outcome_sd)
  drift_conc: float            # InverseGamma on the drift VARIANCE (lib.py:472-473)
  drift_scale: float
  drift_ub: float              # the prior's upper_bound (lib.py:474), applied like gibbs_np (Problem.ub_on_scale)

  @property
  def K(self):
    return len(self.n)

  @property
  def offsets(self):
    return np.concatenate([[1], 1 + np.cumsum(self.n)[:-1]]).astype(int) if self.n else \
        np.zeros(0, int)

  @property
  def d(self):
    return 1 + int(sum(self.n))


def make_spec(seasons: Sequence, T: int, outcome_sd: float) -> SeasonalSpec:
  """``seasons``: objects with num_seasons / num_steps_per_season (the API's Seasons)."""
  idx, ends = [], []
  for s in seasons:
    i, e = season_schedule(int(s.num_seasons), s.num_steps_per_season, T)
    idx.append(i); ends.append(e)
  return SeasonalSpec(n=[int(s.num_seasons) for s in seasons],
                      idx=np.asarray(idx).reshape(len(seasons), T),
                      ends=np.asarray(ends).reshape(len(seasons), T), init_sd=float(outcome_sd),
                      drift_conc=0.005, drift_scale=5e-7 * outcome_sd ** 2,
                      drift_ub=float(outcome_sd))


# ---------------------------------------------------------------------------
# effects-space state-space pieces
# ---------------------------------------------------------------------------
def obs_cols(sp: SeasonalSpec, t: int):
  """State indices the observation sums at step t: level + each component's active season."""
  return [0] + [int(sp.offsets[k] + sp.idx[k, t]) for k in range(sp.K)]


def noise_dir(sp: SeasonalSpec, k: int, t: int):
  """c = C e_j on component k's block (d-vector), j = the season that ends after step t."""
  c = np.zeros(sp.d)
  o, n = sp.offsets[k], sp.n[k]
  c[o:o + n] = -1.0 / n
  c[o + sp.idx[k, t]] += 1.0
  return c


def prior_cov(sp: SeasonalSpec, P0_level: float):
  P = np.zeros((sp.d, sp.d))
  P[0, 0] = P0_level
  for k in range(sp.K):
    o, n = sp.offsets[k], sp.n[k]
    P[o:o + n, o:o + n] = sp.init_sd ** 2 * (np.eye(n) - np.ones((n, n)) / n)
  return P


def dense_moments(sp: SeasonalSpec, T: int, s_e, s_h, s_d, m0, P0_level):
  """Brute-force Gaussian moments of the effects-space model.
  Returns mu_y [T], Syy [T,T], Sxy [T,d,T] (Cov(x_t, y)), Sxx_diag [T,d,d] (Cov(x_t))."""
  d = sp.d
  Pt = prior_cov(sp, P0_level)
  # Cov(x_s, x_t) = Cov(x_min(s,t)) because increments are independent and transition = I
  covs = np.empty((T, d, d))
  for t in range(T):
    covs[t] = Pt
    Q = np.zeros((d, d)); Q[0, 0] = s_h
    for k in range(sp.K):
      if sp.ends[k, t]:
        c = noise_dir(sp, k, t)
        Q += s_d[k] * np.outer(c, c)
    Pt = Pt + Q
  H = np.zeros((T, d))
  for t in range(T):
    H[t, obs_cols(sp, t)] = 1.0
  Syy = np.empty((T, T)); Sxy = np.empty((T, d, T))
  for s in range(T):
    for t in range(T):
      Syy[s, t] = H[s] @ covs[min(s, t)] @ H[t]
      Sxy[s, :, t] = covs[min(s, t)] @ H[t]
  Syy += s_e * np.eye(T)
  mu_y = np.full(T, m0)
  return mu_y, Syy, Sxy, covs


def tfp_constrained_matrices(n: int):
  """E2R [n-1, n] and R2E [n, n-1] of TFP's constrained seasonal (restated from the
  tfp.sts.Seasonal source: effects_to_residuals = (I - 11'/n) without its last row,
  residuals_to_effects = its pseudo-inverse)."""
  E2R = (np.eye(n) - np.ones((n, n)) / n)[:-1]
  return E2R, np.linalg.pinv(E2R)


def tfp_form_y_cov(n, idx_ends, T, s_e, s_d, init_sd):
  """Dense Cov(y) of ONE constrained seasonal component + observation noise, built the way
  TFP builds it (rotating latent, observation = element 0 of the effects)."""
  E2R, R2E = tfp_constrained_matrices(n)
  rot = np.zeros((n, n))
  for i in range(n):
    rot[i, (i + 1) % n] = 1.0                      # new effect i = old effect i+1
  ends = idx_ends
  A_change = E2R @ rot @ R2E
  Qe = np.zeros((n, n)); Qe[n - 1, n - 1] = s_d     # noise on the bottom (just-finished) effect
  Q_change = E2R @ Qe @ E2R.T
  h = (np.eye(n)[0] @ R2E)                          # observation picks effect 0
  P = E2R @ (init_sd ** 2 * np.eye(n)) @ E2R.T
  # propagate Cov(z_s, z_t) = Phi(t<-s) Cov(z_s)
  covs, Phis = [], []
  for t in range(T):
    covs.append(P)
    A = A_change if ends[t] else np.eye(n - 1)
    Q = Q_change if ends[t] else np.zeros((n - 1, n - 1))
    Phis.append(A)
    P = A @ P @ A.T + Q
  Syy = np.empty((T, T))
  for s in range(T):
    Phi = np.eye(n - 1)
    for t in range(s, T):
      Syy[s, t] = Syy[t, s] = h @ Phi @ covs[s] @ h
      Phi = Phis[t] @ Phi
  return Syy + s_e * np.eye(T)


# ---------------------------------------------------------------------------
# Durbin-Koopman simulation smoother with the fast state smoother (what the kernel runs)
# ---------------------------------------------------------------------------
def dk_mean_correction(sp: SeasonalSpec, ystar, mask, s_e, s_h, s_d, P0_level):
  """E[x_t | ystar] - E[x_t] for the zero-prior-mean model; ystar [T] (ignored where mask).
  Passes A (filter: K_t, e_t), B (backward r recursion), C (forward state recursion).
  Returns xhat [T, d]."""
  T, d = ystar.shape[0], sp.d
  P = prior_cov(sp, P0_level)
  a = np.zeros(d)
  Kt = np.zeros((T, d)); et = np.zeros(T)
  Qs = []
  for t in range(T):                                # ---- pass A
    cols = obs_cols(sp, t)
    if not mask[t]:
      Ph = P[:, cols].sum(axis=1)
      F = Ph[cols].sum() + s_e
      v = ystar[t] - a[cols].sum()
      Kt[t] = Ph / F; et[t] = v / F
      a = a + Kt[t] * v
      P = P - np.outer(Ph, Ph) / F
    Q = np.zeros((d, d)); Q[0, 0] = s_h
    for k in range(sp.K):
      if sp.ends[k, t]:
        c = noise_dir(sp, k, t)
        Q += s_d[k] * np.outer(c, c)
    Qs.append(Q)
    P = P + Q
  r = np.zeros(d)
  rt = np.zeros((T, d))
  for t in range(T - 1, -1, -1):                    # ---- pass B
    rt[t] = r                                       # r_t: used by x_{t+1} = x_t + Q_t r_t
    cols = obs_cols(sp, t)
    r = r.copy()
    r[cols] += et[t] - Kt[t] @ rt[t]
  xhat = np.empty((T, d))
  x = prior_cov(sp, P0_level) @ r                   # x_0 = P_0 r_{-1}
  for t in range(T):                                # ---- pass C
    xhat[t] = x
    x = x + Qs[t] @ rt[t]
  return xhat


def prior_draw(sp: SeasonalSpec, T, s_e, s_h, s_d, m0, P0_level, z_init, z_eta, z_eps, z_drift):
  """x+ [T, d], y+ [T] from standard normals: z_init [d], z_eta [T], z_eps [T],
  z_drift [K, T] (only entries at season ends are used)."""
  d = sp.d
  x = np.zeros(d)
  x[0] = m0 + np.sqrt(P0_level) * z_init[0]
  for k in range(sp.K):
    o, n = sp.offsets[k], sp.n[k]
    zb = z_init[o:o + n]
    x[o:o + n] = sp.init_sd * (zb - zb.mean())
  xs = np.empty((T, d)); ys = np.empty(T)
  for t in range(T):
    xs[t] = x
    ys[t] = x[obs_cols(sp, t)].sum() + np.sqrt(s_e) * z_eps[t]
    x = x.copy()
    x[0] += np.sqrt(s_h) * z_eta[t]
    for k in range(sp.K):
      if sp.ends[k, t]:
        x += np.sqrt(s_d[k]) * z_drift[k, t] * noise_dir(sp, k, t)
  return xs, ys


def posterior_state_draw(sp, r, mask, s_e, s_h, s_d, m0, P0_level, rng):
  """One draw of x_{0:T-1} | r (r = y - X.w, NaN / masked steps ignored)."""
  T, d = r.shape[0], sp.d
  xp, yp = prior_draw(sp, T, s_e, s_h, s_d, m0, P0_level, rng.normal(size=d),
                      rng.normal(size=T), rng.normal(size=T), rng.normal(size=(max(sp.K, 1), T)))
  ystar = np.where(mask, 0.0, r - yp)
  return xp + dk_mean_correction(sp, ystar, mask, s_e, s_h, s_d, P0_level)


# ---------------------------------------------------------------------------
# the Gibbs sweep with seasonal components
# ---------------------------------------------------------------------------
def run(prob, sp: SeasonalSpec, *, n_results, n_warmup, seed, prior_level_sd=0.01, sparse=False):
  """Single chain like the reference.  Same structure as oracle/gibbs_np.run with the level
  draw replaced by the joint (level, seasonal effects) draw and one InverseGamma draw per
  seasonal drift variance.  Returns stacked draws: w, s_e, s_h, s_d [., K], level [., T],
  seasonal [., T, K] (each component's contribution at every step)."""
  from oracle import gibbs_np as G
  from oracle.kalman_np import initial_theta
  rng = np.random.Generator(np.random.PCG64(seed))
  T, p, K = prob.T, prob.p, sp.K
  obs = ~prob.mask
  n_obs = int(obs.sum())
  th0 = initial_theta(prob, prior_level_sd)
  s_e, s_h = np.exp(th0[p]), np.exp(th0[p + 1])
  s_d = np.full(K, (0.01 * sp.init_sd) ** 2)        # lib.py:573-574
  w = np.zeros(p)
  y0 = np.where(obs, prob.y, 0.0)
  latent = np.zeros(T)                              # level + seasonal contributions
  if p:
    Xo = prob.X[obs]
    XtX = Xo.T @ Xo
    pi = min(1.0, 3.0 / p)
    logit_pi = np.log(pi) - np.log1p(-pi) if pi < 1 else np.inf
    gamma = np.zeros(p, bool) if sparse else np.ones(p, bool)
  n_ends = sp.ends[:, :T - 1].sum(axis=1) if K else np.zeros(0)
  out = dict(w=[], s_e=[], s_h=[], s_d=[], level=[], seasonal=[])
  for it in range(n_warmup + n_results):
    targ = (y0 - latent)[obs]
    yty = float(targ @ targ)
    if p:
      Xty = Xo.T @ targ
      if sparse:
        for j in range(p):
          g1 = gamma.copy(); g1[j] = True
          g0 = gamma.copy(); g0[j] = False
          l1 = G._log_marginal(g1, XtX, Xty, yty, prob.Omega, n_obs, prob.obs_conc, prob.obs_scale)
          l0 = G._log_marginal(g0, XtX, Xty, yty, prob.Omega, n_obs, prob.obs_conc, prob.obs_scale)
          gamma[j] = rng.random() < 1.0 / (1.0 + np.exp(-(l1 - l0 + logit_pi)))
      idx = np.flatnonzero(gamma)
      w = np.zeros(p)
      sse = yty
      if idx.size:
        Lg = XtX[np.ix_(idx, idx)] + prob.Omega[np.ix_(idx, idx)]
        Lc = np.linalg.cholesky(Lg)
        wbar = np.linalg.solve(Lg, Xty[idx])
        sse = yty - wbar @ Lg @ wbar
      s_e = min(1.0 / rng.gamma(prob.obs_conc + 0.5 * n_obs, 1.0 / (prob.obs_scale + 0.5 * sse)),
                prob.ub_var(prob.obs_ub))
      if idx.size:
        w[idx] = wbar + np.sqrt(s_e) * np.linalg.solve(Lc.T, rng.normal(size=idx.size))
      r = prob.y - prob.X @ w
    else:
      # no covariates: sigma_obs^2 from y - level - seasonal (sweep 0: latent = 0)
      s_e = min(1.0 / rng.gamma(prob.obs_conc + 0.5 * n_obs, 1.0 / (prob.obs_scale + 0.5 * yty)),
                prob.ub_var(prob.obs_ub))
      r = prob.y
    x = posterior_state_draw(sp, np.where(obs, r, 0.0), prob.mask, s_e, s_h, s_d, prob.m0, prob.P0,
                             rng)
    level = x[:, 0]
    seas = np.stack([x[np.arange(T), sp.offsets[k] + sp.idx[k]] for k in range(K)], axis=1) \
        if K else np.zeros((T, 0))
    latent = level + seas.sum(axis=1)
    dl = np.diff(level)
    s_h = min(1.0 / rng.gamma(prob.lvl_conc + 0.5 * (T - 1), 1.0 / (prob.lvl_scale + 0.5 * dl @ dl)),
              prob.ub_var(prob.lvl_ub))
    for k in range(K):
      # the season-end increment is u * C e_j; its j-th entry is u (1 - 1/n)
      tt = np.flatnonzero(sp.ends[k, :T - 1])
      j = sp.offsets[k] + sp.idx[k, tt]
      u = (x[tt + 1, j] - x[tt, j]) / (1.0 - 1.0 / sp.n[k])
      s_d[k] = min(1.0 / rng.gamma(sp.drift_conc + 0.5 * n_ends[k],
                                   1.0 / (sp.drift_scale + 0.5 * u @ u)), prob.ub_var(sp.drift_ub))
    if it >= n_warmup:
      out["w"].append(w.copy()); out["s_e"].append(s_e); out["s_h"].append(s_h)
      out["s_d"].append(s_d.copy()); out["level"].append(level.copy()); out["seasonal"].append(seas)
  return {k: np.asarray(v) for k, v in out.items()}
