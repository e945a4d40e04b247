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


# --------------------------------------------------------------------------
# Local linear trend (d = 2): the formulation of csrc/ci_llt.cuh in float64.
# Forward: Sarkka & Garcia-Fernandez filtering elements (A, b, C, eta, J);
# backward: abar / Pbar recursions with the shared multiplier G = A (I - K h').
# --------------------------------------------------------------------------
_A2 = np.array([[1.0, 1.0], [0.0, 1.0]])
_H2 = np.array([1.0, 0.0])


def llt_element(t, r, mask, s_e, q1, q2, m0, P0):
  Q = np.diag([q1, q2])
  Z2, z2 = np.zeros((2, 2)), np.zeros(2)
  if t == 0:
    if mask[0]:
      return (Z2, np.array(m0, float), np.array(P0, float), z2, Z2)
    S = P0[0][0] + s_e
    Kk = np.asarray(P0) @ _H2 / S
    return (Z2, np.asarray(m0) + Kk * (r[0] - m0[0]), np.asarray(P0) - np.outer(Kk, Kk) * S, z2, Z2)
  if mask[t]:
    return (_A2.copy(), z2, Q, z2, Z2)
  S = q1 + s_e
  Kk = Q @ _H2 / S
  M = np.eye(2) - np.outer(Kk, _H2)
  ah = _A2.T @ _H2
  return (M @ _A2, Kk * r[t], M @ Q, ah * r[t] / S, np.outer(ah, ah) / S)


def llt_combine(ei, ej):
  Ai, bi, Ci, hi, Ji = ei
  Aj, bj, Cj, hj, Jj = ej
  I = np.eye(2)
  X = Aj @ np.linalg.inv(I + Ci @ Jj)
  Y = Ai.T @ np.linalg.inv(I + Jj @ Ci)
  return (X @ Ai, X @ (bi + Ci @ hj) + bj, X @ Ci @ Aj.T + Cj, Y @ (hj - Jj @ bi) + hi,
          Y @ Jj @ Ai + Ji)


def llt_scan_value_grad(r, mask, s_e, q1, q2, m0, P0, block=8):
  """Single chain: ll, rbar[T], d/ds_e, d/dq1, d/dq2 via block-combined elements
  (forward) and the hand-derived symmetric adjoint (backward)."""
  T = r.shape[0]
  Q = np.diag([q1, q2])
  # forward: combine per-block aggregates first (exercises associativity)
  blocks = []
  for b0 in range(0, T, block):
    e = llt_element(b0, r, mask, s_e, q1, q2, m0, P0)
    for t in range(b0 + 1, min(b0 + block, T)):
      e = llt_combine(e, llt_element(t, r, mask, s_e, q1, q2, m0, P0))
    blocks.append(e)
  path = []
  ll = 0.0
  pre = None
  for bi, b0 in enumerate(range(0, T, block)):
    cur = pre
    for t in range(b0, min(b0 + block, T)):
      if t == 0:
        a, P = np.array(m0, float), np.array(P0, float)
      else:
        a, P = _A2 @ cur[1], _A2 @ cur[2] @ _A2.T + Q
      path.append((a, P))
      if not mask[t]:
        v = r[t] - a[0]; F = P[0, 0] + s_e
        ll += -0.5 * (np.log(2 * np.pi) + np.log(F) + v * v / F)
      el = llt_element(t, r, mask, s_e, q1, q2, m0, P0)
      cur = el if cur is None else llt_combine(cur, el)
    pre = blocks[bi] if pre is None else llt_combine(pre, blocks[bi])
  # backward
  ab = np.zeros(2); Pb = np.zeros((2, 2)); g_e = 0.0; g_q = np.zeros(2); rb = np.zeros(T)
  for t in range(T - 1, -1, -1):
    a, P = path[t]
    g_q += np.diag(Pb)
    if mask[t]:
      ab = _A2.T @ ab; Pb = _A2.T @ Pb @ _A2
      continue
    v = r[t] - a[0]; F = P[0, 0] + s_e; Kk = P @ _H2 / F
    G = _A2 @ (np.eye(2) - np.outer(Kk, _H2))
    abp = _A2.T @ ab; Pbp = _A2.T @ Pb @ _A2
    dF = -0.5 * (1 / F - v * v / F ** 2)
    rb[t] = Kk @ abp - v / F
    g_e += dF - (abp @ Kk) * v / F + Kk @ Pbp @ Kk
    u = G.T @ ab
    E = np.array([[u[0], u[1] / 2], [u[1] / 2, 0.0]]) * v / F + dF * np.outer(_H2, _H2)
    Pb = G.T @ Pb @ G + E
    ab = G.T @ ab + _H2 * v / F
  return ll, rb, g_e, g_q[0], g_q[1]
