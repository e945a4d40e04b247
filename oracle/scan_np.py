"""Time-parallel (associative-scan) restatement of the local-level filter and
its adjoint, in float64 NumPy.  TEST INFRASTRUCTURE ONLY.

This is the *formulation* the CUDA scan kernel implements (csrc/ci_filter.cuh):
the variance recursion is a Moebius map P' = (aP+b)/(cP+d) composed as 2x2
matrices, the mean recursion and both adjoint recursions are first-order affine
maps x' = m x + c.  Here every scan is evaluated with explicit prefix products
so that tests/test_oracle_scan.py can show it equals the sequential recursion
of oracle/kalman_np.py (which restates TFP's LinearGaussianStateSpaceModel
conventions; the reference call site is causalimpact_lib.py:365-388).
"""
from __future__ import annotations

import numpy as np

from oracle.kalman_np import LOG2PI


def _prefix_affine(m, c, x0):
  """x_{t+1} = m_t x_t + c_t, returns x_0..x_{T-1} (exclusive prefix) and x_T."""
  T = m.shape[0]
  M = np.ones(T + 1); Cc = np.zeros(T + 1)
  for t in range(T):            # inclusive composition (later o earlier)
    M[t + 1] = m[t] * M[t]
    Cc[t + 1] = m[t] * Cc[t] + c[t]
  x = M * x0 + Cc
  return x[:-1], x[-1]


def _prefix_mobius(obs, s_e, s_h, P0):
  """P path via 2x2 matrix prefix products (normalised each step)."""
  T = obs.shape[0]
  Mo = np.array([[s_e + s_h, s_e * s_h], [1.0, s_e]])
  Mm = np.array([[1.0, s_h], [0.0, 1.0]])
  acc = np.eye(2)
  P = np.empty(T)
  for t in range(T):
    P[t] = (acc[0, 0] * P0 + acc[0, 1]) / (acc[1, 0] * P0 + acc[1, 1])
    acc = (Mo if obs[t] else Mm) @ acc
    acc = acc / acc.sum()
  Pout = (acc[0, 0] * P0 + acc[0, 1]) / (acc[1, 0] * P0 + acc[1, 1])
  return P, Pout


def ll_scan_value_grad(r, mask, s_e, s_h, m0, P0):
  """Single chain.  Returns ll, rbar[T], d/ds_e, d/ds_h via scans only."""
  T = r.shape[0]
  obs = ~mask
  P, _ = _prefix_mobius(obs, s_e, s_h, P0)
  F = P + s_e
  K = np.where(obs, P / F, 0.0)
  rr = np.where(obs, r, 0.0)
  a, _ = _prefix_affine(1.0 - K, K * rr, m0)
  v = np.where(obs, rr - a, 0.0)
  ll = -0.5 * np.sum(np.where(obs, LOG2PI + np.log(F) + v * v / F, 0.0))
  # reverse affine scans: flip time
  g = np.where(obs, v / F, 0.0)
  ab_rev, _ = _prefix_affine((1.0 - K)[::-1], g[::-1], 0.0)
  abn = ab_rev[::-1]                       # abar_{t+1} seen from step t
  dF = np.where(obs, -0.5 * (1.0 / F - v * v / (F * F)), 0.0)
  q = np.where(obs, abn * v * s_e / (F * F) + dF, 0.0)
  mult = np.where(obs, (1.0 - K) ** 2, 1.0)
  pb_rev, _ = _prefix_affine(mult[::-1], q[::-1], 0.0)
  pbn = pb_rev[::-1]                       # Pbar_{t+1} seen from step t
  rbar = np.where(obs, K * abn - v / F, 0.0)
  g_h = np.sum(pbn)
  g_e = np.sum(np.where(obs, K * K * pbn - abn * v * P / (F * F) + dF, 0.0))
  return ll, rbar, g_e, g_h
