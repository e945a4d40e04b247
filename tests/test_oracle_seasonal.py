"""Pins oracle/seasonal_np.py (the seasonal model + Gibbs sweep the CUDA kernel is checked
against) to exact linear-Gaussian identities -- TFP itself is unobtainable:
  * the season schedule against hand-written cases (int / tuple / nested-tuple
    num_steps_per_season, as the reference's own test uses, causalimpact_lib_test.py:741-755);
  * the engine's "effects space" form == TFP's rotating, constrained construction: equal
    dense covariance of y;
  * Durbin-Koopman mean correction == dense Gaussian conditional mean; draws have the
    dense conditional covariance;
  * the Gibbs sweep recovers a planted seasonal pattern."""
import types

import numpy as np

from oracle import kalman_np as K
from oracle import seasonal_np as S


def seasons(*args):
  return [types.SimpleNamespace(num_seasons=n, num_steps_per_season=s) for n, s in args]


def test_schedule():
  idx, ends = S.season_schedule(3, 1, 7)
  assert idx.tolist() == [0, 1, 2, 0, 1, 2, 0] and ends.all()
  idx, ends = S.season_schedule(4, (2, 1, 1, 1), 11)
  assert idx.tolist() == [0, 0, 1, 2, 3, 0, 0, 1, 2, 3, 0]
  assert ends.tolist() == [False, True, True, True, True, False, True, True, True, True, False]
  idx, ends = S.season_schedule(2, ((1, 2), (3, 1)), 10)
  assert idx.tolist() == [0, 1, 1, 0, 0, 0, 1, 0, 1, 1]
  assert ends.tolist() == [True, False, True, False, False, True, True, True, False, True]
  idx, ends = S.season_schedule(7, 24, 24 * 7 * 2)
  assert idx[23] == 0 and idx[24] == 1 and ends[23] and not ends[22] and idx[24 * 7] == 0


def test_effects_space_equals_tfp_constrained_form():
  T, s_e = 23, 0.3
  for n, steps, s_d, sd in ((4, (2, 1, 1, 1), 0.05, 1.3), (7, 1, 0.2, 0.8), (3, 2, 0.0, 1.0)):
    sp = S.make_spec(seasons((n, steps)), T, sd)
    # the level is switched off: P0 = 0, s_h = 0 -> y = seasonal contribution + noise
    _, Syy, _, _ = S.dense_moments(sp, T, s_e, 0.0, [s_d], 0.0, 0.0)
    want = S.tfp_form_y_cov(n, sp.ends[0], T, s_e, s_d, sd)
    np.testing.assert_allclose(Syy, want, rtol=1e-12, atol=1e-12)


def _setup(T=40, seed=0):
  rng = np.random.default_rng(seed)
  sp = S.make_spec(seasons((4, (2, 1, 1, 1)), (7, 1)), T, 1.1)
  mask = np.zeros(T, bool); mask[[3, 11]] = True; mask[30:] = True
  y = rng.normal(size=T) + 0.5
  return sp, mask, y, rng


def test_dk_mean_correction_is_the_gaussian_conditional_mean():
  sp, mask, y, _ = _setup()
  s_e, s_h, s_d, m0, P0 = 0.2, 0.01, [0.03, 0.002], 0.4, 1.5
  mu, Syy, Sxy, _ = S.dense_moments(sp, 40, s_e, s_h, s_d, m0, P0)
  o = ~mask
  # E[x_t | y_obs] = E[x_t] + Cov(x_t, y_o) Syy_oo^-1 (y_o - mu_o); E[x_t] = (m0, 0, ...)
  want = np.stack([Sxy[t][:, o] @ np.linalg.solve(Syy[np.ix_(o, o)], (y - mu)[o])
                   for t in range(40)])
  got = S.dk_mean_correction(sp, np.where(mask, 0.0, y - mu), mask, s_e, s_h, s_d, P0)
  np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-10)


def test_dk_draws_have_the_conditional_covariance():
  sp, mask, y, rng = _setup(T=24, seed=1)
  mask = mask[:24]; y = y[:24]; mask[20:] = True
  sp = S.make_spec(seasons((4, (2, 1, 1, 1)), (7, 1)), 24, 1.1)
  s_e, s_h, s_d, m0, P0 = 0.2, 0.02, [0.03, 0.01], 0.4, 1.5
  mu, Syy, Sxy, covs = S.dense_moments(sp, 24, s_e, s_h, s_d, m0, P0)
  o = ~mask
  n = 4000
  draws = np.stack([S.posterior_state_draw(sp, np.where(mask, 0.0, y), mask, s_e, s_h, s_d, m0,
                                           P0, rng) for _ in range(n)])
  for t in (0, 7, 19, 23):
    G = Sxy[t][:, o]
    cond_cov = covs[t] - G @ np.linalg.solve(Syy[np.ix_(o, o)], G.T)
    cond_mean = np.r_[m0, np.zeros(sp.d - 1)] + G @ np.linalg.solve(Syy[np.ix_(o, o)], (y - mu)[o])
    emp = draws[:, t, :]
    se = np.sqrt(np.diag(cond_cov) / n) + 1e-12
    assert np.all(np.abs(emp.mean(0) - cond_mean) < 5 * se + 1e-9)
    sd_ref = np.sqrt(np.diag(cond_cov))
    np.testing.assert_allclose(np.cov(emp.T), cond_cov,
                               atol=6 * np.outer(sd_ref, sd_ref).max() / np.sqrt(n))
    # each block stays in the zero-sum subspace (constrain_mean_effect_to_zero)
    for k in range(sp.K):
      blk = emp[:, sp.offsets[k]:sp.offsets[k] + sp.n[k]]
      assert np.abs(blk.sum(1)).max() < 1e-9


def test_gibbs_recovers_a_planted_seasonal_pattern():
  rng = np.random.default_rng(4)
  T = 140
  pat = np.array([1.0, 4.0, 5.0, 2.0, -1.0, -2.0, -3.0]); pat -= pat.mean()
  y = 0.3 * pat[np.arange(T) % 7] + 0.15 * rng.normal(size=T)
  sd = y[:100].std(ddof=1)
  ys = (y - y[:100].mean()) / sd
  y_ext = ys.copy(); y_ext[100:] = np.nan
  prob = K.default_problem(y_ext, None, outcome_sd=1.0)
  sp = S.make_spec(seasons((7, 1)), T, 1.0)
  out = S.run(prob, sp, n_results=150, n_warmup=100, seed=2)
  contrib = out["seasonal"][:, :, 0].mean(0)
  truth = 0.3 * pat[np.arange(T) % 7] / sd
  assert np.abs(contrib - truth).max() < 0.25
  # with the pattern explained, the noise scale is far below the outcome sd (no covariates:
  # initial value is sd itself, lib.py:566-571)
  assert np.sqrt(out["s_e"]).mean() < 0.45
  assert out["s_d"].shape == (150, 1) and out["level"].shape == (150, T)
