"""The C port (CPU baseline) must equal the NumPy oracle to rounding."""
import numpy as np
import pytest

from conftest import make_series, make_thetas
from oracle import c_port
from oracle import kalman_np as K


@pytest.mark.parametrize("T,n_cov", [(100, 1), (500, 10), (300, 0)])
def test_c_port_equals_numpy_oracle(T, n_cov):
  y, X, _ = make_series(T, n_cov, 3)
  prob = K.default_problem(y, X)
  th = make_thetas(prob.dim, prob.p, 9, 2)
  th[0, prob.p] = np.log(prob.ub_var(prob.obs_ub) * 1.2)        # out of support
  v, g, used = c_port.logpost_grad(prob, th)
  ov, og = K.log_post_grad(prob, th)
  np.testing.assert_allclose(v, ov, rtol=1e-12, atol=1e-10)
  np.testing.assert_allclose(g, og, rtol=1e-9, atol=1e-9)
  assert used >= 1
  v2, _, _ = c_port.logpost_grad(prob, th, with_prior=False, want_grad=False)
  np.testing.assert_allclose(v2, K.log_lik(prob, th), rtol=1e-12)


@pytest.mark.parametrize("T,n_cov", [(120, 2), (400, 6)])
def test_c_port_local_linear_trend_equals_numpy_oracle(T, n_cov):
  """BASELINE configs[2] model (d = 2) in the C port == the generic NumPy filter + adjoint."""
  y, X, _ = make_series(T, n_cov, 8, nan_frac=0.03)
  prob = K.default_problem(y, X, model=1)
  th = make_thetas(prob.dim, prob.p, 7, 4, d=2)
  th[1, prob.p + 2] = np.log(prob.ub_var(prob.slope_ub) * 1.3)   # out of support
  v, g, _ = c_port.logpost_grad(prob, th)
  ov, og = K.log_post_grad(prob, th)
  np.testing.assert_allclose(v, ov, rtol=1e-11, atol=1e-9)
  np.testing.assert_allclose(g, og, rtol=1e-8, atol=1e-8)
  v2, _, _ = c_port.logpost_grad(prob, th, with_prior=False, want_grad=False)
  np.testing.assert_allclose(v2, K.log_lik(prob, th), rtol=1e-11)


def test_c_port_posterior_predict_equals_numpy_oracle():
  """The C smoother + predictive draws (CPU baseline of posterior draws/s) == smoother_np,
  same Philox streams; independent of the thread count."""
  from oracle import smoother_np as SM
  y, X, _ = make_series(150, 3, 5, nan_frac=0.05)
  prob = K.default_problem(y, X)
  th = make_thetas(prob.dim, prob.p, 6, 9)
  l, t, m, used = c_port.posterior_predict(prob, th, seed=77, draw_id0=5)
  ol, ot, om = SM.posterior_predict(prob, th, 77, 5)
  np.testing.assert_allclose(l, ol, rtol=1e-10, atol=1e-10)
  np.testing.assert_allclose(t, ot, rtol=1e-10, atol=1e-10)
  np.testing.assert_allclose(m, om, rtol=1e-10, atol=1e-10)
  l1, t1, m1, _ = c_port.posterior_predict(prob, th, seed=77, draw_id0=5, nthreads=1)
  assert np.array_equal(t1, t) and np.array_equal(l1, l)
  np.testing.assert_allclose(m1, m, rtol=1e-12)
