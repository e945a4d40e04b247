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
    _lib.ci_oracle_llt_logpost_grad.argtypes = _lib.ci_oracle_logpost_grad.argtypes
    _lib.ci_oracle_llt_logpost_grad.restype = C.c_int
    _lib.ci_oracle_predict.argtypes = [dp, dp, C.c_int, C.c_int, dp, dp, C.c_int, C.c_uint64,
                                       C.c_uint64, dp, dp, dp, C.c_int]
    _lib.ci_oracle_predict.restype = C.c_int
  return _lib


def _dp(a):
  return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def logpost_grad(prob, theta, with_prior=True, want_grad=True, nthreads=0):
  """prob: oracle.kalman_np.Problem (local level or local linear trend).  Returns
  (val, grad, threads)."""
  lib = load()
  if prob.d == 2:
    return _llt_logpost_grad(lib, prob, theta, with_prior, want_grad, nthreads)
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


def _llt_logpost_grad(lib, prob, theta, with_prior, want_grad, nthreads):
  theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
  n, p = theta.shape[0], prob.p
  y = np.ascontiguousarray(prob.y, np.float64)
  X = np.ascontiguousarray(prob.X if p else np.zeros((prob.T, 0)), np.float64)
  Om = np.ascontiguousarray(prob.Omega if p else np.zeros((0, 0)), np.float64)
  sq = lambda ub: np.sqrt(prob.ub_var(ub)) if np.isfinite(ub) else 1e300
  prior = np.array([prob.m0, prob.P0, prob.obs_conc, prob.obs_scale, sq(prob.obs_ub),
                    prob.lvl_conc, prob.lvl_scale, sq(prob.lvl_ub), prob.slope_conc,
                    prob.slope_scale, sq(prob.slope_ub), prob.m0_slope, prob.P0_slope], np.float64)
  val = np.empty(n); grad = np.empty_like(theta) if want_grad else None
  used = lib.ci_oracle_llt_logpost_grad(_dp(y), _dp(X), _dp(Om), prob.T, p, _dp(prior),
                                        _dp(theta), n, _dp(val), _dp(grad), int(with_prior),
                                        nthreads)
  return val, grad, used


def posterior_predict(prob, theta_draws, seed, draw_id0=0, want_level=True, nthreads=0):
  """C port of oracle.smoother_np.posterior_predict (local level): (level [S,T] or None,
  traj [S,T], mean [T], threads)."""
  lib = load()
  assert prob.d == 1
  th = np.ascontiguousarray(np.atleast_2d(theta_draws), dtype=np.float64)
  S, T, p = th.shape[0], prob.T, prob.p
  y = np.ascontiguousarray(prob.y, np.float64)
  X = np.ascontiguousarray(prob.X if p else np.zeros((T, 0)), np.float64)
  prior = np.array([prob.m0, prob.P0], np.float64)
  level = np.empty((S, T)) if want_level else None
  traj = np.empty((S, T)); loc = np.empty(T)
  used = lib.ci_oracle_predict(_dp(y), _dp(X), T, p, _dp(prior), _dp(th), S, int(seed) & (2**64 - 1),
                               int(draw_id0), _dp(level), _dp(traj), _dp(loc), nthreads)
  return level, traj, loc / S, used
