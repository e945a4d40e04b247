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
