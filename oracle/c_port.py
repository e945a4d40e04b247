"""ctypes loader for oracle/_ref/libci_oracle.so (the C port of kalman_np).
TEST INFRASTRUCTURE / CPU BASELINE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libci_oracle.so")
_lib = None


def load():
  global _lib
  if _lib is None:
    if not os.path.exists(LIB):
      subprocess.run(["make", "-s", "-C", HERE], check=True)
    _lib = C.CDLL(LIB)
    dp = C.POINTER(C.c_double)
    _lib.ci_oracle_logpost_grad.argtypes = [dp, dp, dp, C.c_int, C.c_int, dp, dp, C.c_int, dp,
                                            dp, C.c_int, C.c_int]
    _lib.ci_oracle_logpost_grad.restype = C.c_int
  return _lib


def _dp(a):
  return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def logpost_grad(prob, theta, with_prior=True, want_grad=True, nthreads=0):
  """prob: oracle.kalman_np.Problem (local level).  Returns (val, grad, threads)."""
  lib = load()
  assert prob.d == 1
  theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
  n, p = theta.shape[0], prob.p
  y = np.ascontiguousarray(prob.y, np.float64)
  X = np.ascontiguousarray(prob.X if p else np.zeros((prob.T, 0)), np.float64)
  Om = np.ascontiguousarray(prob.Omega if p else np.zeros((0, 0)), np.float64)
  # (the C port bounds the scale: hand it the square root of the variance bound)
  prior = np.array([prob.m0, prob.P0, prob.obs_conc, prob.obs_scale,
                    np.sqrt(prob.ub_var(prob.obs_ub)), prob.lvl_conc, prob.lvl_scale,
                    np.sqrt(prob.ub_var(prob.lvl_ub))], np.float64)
  val = np.empty(n); grad = np.empty_like(theta) if want_grad else None
  used = lib.ci_oracle_logpost_grad(_dp(y), _dp(X), _dp(Om), prob.T, p, _dp(prior), _dp(theta),
                                    n, _dp(val), _dp(grad), int(with_prior), nthreads)
  return val, grad, used
