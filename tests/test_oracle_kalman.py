"""Pins the float64 oracle itself to ground truth (no GPU, no reference)."""
import numpy as np
import pytest

from conftest import make_series
from oracle import kalman_np as K
from oracle import scan_np as S


def _toy(model, seed=0, T=60, p=3):
  rng = np.random.default_rng(seed)
  X = rng.normal(size=(T, p)); X[:, -1] = 1
  y = X @ np.array([.5, -.3, .1]) + np.cumsum(rng.normal(size=T)) * .1 + rng.normal(size=T) * .4
  y[[3, 7, 20]] = np.nan; y[45:] = np.nan
  prob = K.default_problem(y, X, model=model, prior_level_sd=0.1)
  th = rng.normal(size=(4, prob.dim)) * .3
  th[:, p] += np.log(.2); th[:, p + 1] += np.log(.01)
  if model == 1:
    th[:, p + 2] += np.log(.001)
  return prob, th


@pytest.mark.parametrize("model", [0, 1])
def test_loglik_equals_dense_gaussian_marginal(model):
  prob, th = _toy(model)
  ll = K.log_lik(prob, th)
  dense = np.array([K.dense_marginal_loglik(prob, t) for t in th])
  np.testing.assert_allclose(ll, dense, rtol=0, atol=1e-9)


@pytest.mark.parametrize("model", [0, 1])
def test_gradient_equals_central_differences(model):
  prob, th = _toy(model, seed=1)
  _, g = K.log_post_grad(prob, th)
  eps = 1e-6
  fd = np.zeros_like(g)
  for j in range(prob.dim):
    tp, tm = th.copy(), th.copy()
    tp[:, j] += eps; tm[:, j] -= eps
    fd[:, j] = (K.log_post(prob, tp) - K.log_post(prob, tm)) / (2 * eps)
  np.testing.assert_allclose(g, fd, rtol=1e-6, atol=1e-6)


def test_steady_state_gain_closed_form():
  # P_bar (predicted) solves P = P s_e/(P+s_e) + s_h  =>  P = (s_h + sqrt(s_h^2 + 4 s_h s_e))/2
  s_e, s_h = 0.3, 0.02
  r = np.zeros((1, 400))
  _, _, P = K.ll_filter(r, np.zeros(400, bool), np.array([s_e]), np.array([s_h]), 0.0, 1.0,
                        return_path=True)
  closed = 0.5 * (s_h + np.sqrt(s_h * s_h + 4 * s_h * s_e))
  assert abs(P[-1, 0] - closed) < 1e-12


def test_generic_filter_reduces_to_scalar():
  prob, th = _toy(0, seed=2)
  th, W, se, sh, _ = K._unpack(prob, th)
  R = K.residuals(prob, W)
  a = K.ll_filter_grad(R, prob.mask, se, sh, prob.m0, prob.P0)
  b = K.gen_filter_grad(R, prob.mask, se, sh[:, None], [prob.m0], [[prob.P0]],
                        np.array([[1.0]]), np.array([1.0]))
  np.testing.assert_allclose(a[0], b[0], atol=1e-11)
  np.testing.assert_allclose(a[1], b[1], atol=1e-11)
  np.testing.assert_allclose(a[2], b[2], atol=1e-10)
  np.testing.assert_allclose(a[3], b[3][:, 0], atol=1e-10)


def test_scan_formulation_equals_sequential():
  rng = np.random.default_rng(3)
  T = 700
  r = rng.normal(size=T); mask = rng.random(T) < 0.1; mask[500:] = True; r[mask] = np.nan
  ll, rb, ge, gh = K.ll_filter_grad(r[None], mask, np.array([0.3]), np.array([0.02]), 0.2, 1.1)
  l2, rb2, ge2, gh2 = S.ll_scan_value_grad(r, mask, 0.3, 0.02, 0.2, 1.1)
  assert abs(ll[0] - l2) < 1e-10
  np.testing.assert_allclose(rb[0], rb2, atol=1e-12)
  assert abs(ge[0] - ge2) < 1e-9 and abs(gh[0] - gh2) < 1e-9


def test_out_of_support_is_minus_inf():
  prob, th = _toy(0)
  th[0, prob.p] = np.log(prob.ub_var(prob.obs_ub) * 1.02)
  th[1, prob.p + 1] = np.log(prob.ub_var(prob.lvl_ub) * 1.02)
  v = K.log_post(prob, th)
  assert np.isneginf(v[0]) and np.isneginf(v[1]) and np.isfinite(v[2])


def test_priors_match_reference_constants():
  """Constants of causalimpact_lib.py:424-453 / :566-572."""
  y, X, _ = make_series(200, 2, 5)
  prob = K.default_problem(y, X, prior_level_sd=0.01)
  sd = np.nanstd(y, ddof=1)
  assert prob.obs_conc == 25.0 and np.isclose(prob.obs_scale, 5 * sd ** 2)
  assert prob.lvl_conc == 16.0 and np.isclose(prob.lvl_scale, 16 * (0.01 * sd) ** 2)
  assert np.isclose(prob.obs_ub, 1.2 * sd) and np.isclose(prob.lvl_ub, sd)
  xtx = X.T @ X
  om = 0.5 * xtx; om[np.diag_indices(3)] = np.diag(xtx)
  np.testing.assert_allclose(prob.Omega, 0.01 * om / 200)
  prob0 = K.default_problem(y, None)
  assert prob0.obs_conc == 0.005 and np.isclose(prob0.obs_scale, 0.005 * sd ** 2)


def test_llt_scan_formulation_equals_generic_filter():
  """The d=2 formulation of csrc/ci_llt.cuh (filtering elements + congruence
  adjoints) against the generic reverse-mode oracle."""
  rng = np.random.default_rng(0)
  T = 90
  r = rng.normal(size=T) * 0.5 + np.cumsum(rng.normal(size=T)) * 0.05
  mask = rng.random(T) < 0.1; mask[70:] = True; mask[0] = False
  r[mask] = np.nan
  s_e, q1, q2 = 0.2, 0.01, 0.001
  m0 = np.array([0.3, 0.0]); P0 = np.diag([1.1, 0.9])
  A = np.array([[1., 1.], [0., 1.]]); h = np.array([1., 0.])
  ll, rbar, ge, gq = K.gen_filter_grad(r[None], mask, np.array([s_e]), np.array([[q1, q2]]), m0,
                                       P0, A, h)
  l2, rb2, ge2, g1, g2 = S.llt_scan_value_grad(r, mask, s_e, q1, q2, m0, P0)
  assert abs(ll[0] - l2) < 1e-10
  np.testing.assert_allclose(rbar[0], rb2, atol=1e-12)
  assert abs(ge[0] - ge2) < 1e-10 and abs(gq[0, 0] - g1) < 1e-10 and abs(gq[0, 1] - g2) < 1e-9
