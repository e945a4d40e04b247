"""Float64 NumPy restatement of the engine's HMC driver.  TEST INFRASTRUCTURE.

The reference does NOT run HMC: it calls TFP's single-chain Gibbs sampler
(causalimpact_lib.py:365-388).  BASELINE.json's north_star replaces that call
with "custom HMC over the filter kernel"; this file restates that HMC
(csrc/ci_hmc.cuh) step for step -- same Philox streams, same Stan-style
windowed adaptation -- so the CUDA kernel can be checked pathwise in float64
and statistically in float32.  The target density is oracle/kalman_np.log_post
(model + priors of causalimpact_lib.py:398-500).  oracle/gibbs_np.py restates
the reference's own sampler for the statistical comparison.

Algorithm (per chain; L and the adaptation schedule are shared by all chains):
  rho = z / sqrt(minv),  z ~ N(0, I)           (momentum ~ N(0, M), M = 1/minv)
  L leapfrog steps of size eps, L ~ U{1..max_leapfrog} keyed by iteration only
  accept with prob min(1, exp(H0 - H1)); non-finite H1 => reject ("divergent")
  warm-up: dual averaging (gamma=0.05, t0=10, kappa=0.75, mu=log(10 eps)),
           Welford variances over doubling windows -> minv (Stan regularisation)
"""
from __future__ import annotations

import numpy as np

from oracle import philox_np as PH

GAMMA, T0, KAPPA = 0.05, 10.0, 0.75


def adapt_schedule(n_warmup: int):
  """Stan's windowed adaptation.  Returns (init_buffer, slow_end, window_ends)."""
  W = int(n_warmup)
  if W < 20:
    return W, W, []
  init, term, base = 75, 50, 25
  if init + base + term > W:
    init = int(0.15 * W); term = int(0.1 * W); base = W - init - term
  last = W - term - 1
  ends = []
  size = base
  nxt = init + base - 1
  while True:
    ends.append(nxt)
    if nxt == last:
      break
    size *= 2
    n2 = nxt + size
    if n2 != last and n2 + 2 * size >= W - term:
      n2 = last
    if n2 > last:
      n2 = last
    nxt = n2
  return init, W - term, ends


def run(logpost_grad, theta0, *, n_warmup, n_results, seed, chain_id0=0, max_leapfrog=8,
        init_step=0.05, target_accept=0.8, adapt_mass=True):
  """logpost_grad(theta[C,dim]) -> (value[C], grad[C,dim]).  Returns draws
  [n_results, C, dim] and a dict of per-chain stats."""
  th = np.array(np.atleast_2d(theta0), dtype=np.float64)
  C, dim = th.shape
  lp, g = logpost_grad(th)
  minv = np.ones((C, dim))
  eps = np.full(C, float(init_step))
  mu = np.log(10.0 * eps); hbar = np.zeros(C); leb = np.zeros(C); dac = np.zeros(C)
  wn = np.zeros(C); wmean = np.zeros((C, dim)); wm2 = np.zeros((C, dim))
  init_buf, slow_end, ends = adapt_schedule(n_warmup)
  draws = np.empty((n_results, C, dim))
  acc_sum = np.zeros(C); n_div = np.zeros(C, int); n_leap = np.ones(C, int)
  for it in range(n_warmup + n_results):
    L = PH.leapfrog_count(seed, it, max_leapfrog)
    z = np.stack([PH.momentum_normals(seed, chain_id0 + c, it, dim) for c in range(C)])
    rho = z / np.sqrt(minv)
    H0 = -lp + 0.5 * np.sum(minv * rho * rho, axis=1)
    thn, gn, lpn = th.copy(), g.copy(), lp.copy()
    rho = rho + 0.5 * eps[:, None] * gn
    for i in range(L):
      thn = thn + eps[:, None] * minv * rho
      lpn, gn = logpost_grad(thn)
      rho = rho + (1.0 if i < L - 1 else 0.5) * eps[:, None] * gn
    n_leap += L
    H1 = -lpn + 0.5 * np.sum(minv * rho * rho, axis=1)
    dH = H0 - H1
    fin = np.isfinite(dH)
    alpha = np.where(fin, np.minimum(1.0, np.exp(np.minimum(dH, 0.0))), 0.0)
    div = (~fin) | (dH < -1000.0)
    u = np.array([PH.accept_uniform(seed, chain_id0 + c, it) for c in range(C)])
    acc = u < alpha
    th = np.where(acc[:, None], thn, th)
    g = np.where(acc[:, None], gn, g)
    lp = np.where(acc, lpn, lp)
    if it < n_warmup:
      dac += 1.0
      eta = 1.0 / (dac + T0)
      hbar = (1.0 - eta) * hbar + eta * (target_accept - alpha)
      le = mu - hbar * np.sqrt(dac) / GAMMA
      ex = dac ** (-KAPPA)
      leb = (1.0 - ex) * leb + ex * le
      eps = np.exp(le)
      if adapt_mass and init_buf <= it < slow_end:
        wn += 1.0
        d = th - wmean
        wmean = wmean + d / wn[:, None]
        wm2 = wm2 + d * (th - wmean)
        if it in ends:
          var = wm2 / (wn[:, None] - 1.0)
          minv = (wn / (wn + 5.0))[:, None] * var + 1e-3 * (5.0 / (wn + 5.0))[:, None]
          wn[:] = 0; wmean[:] = 0; wm2[:] = 0
          mu = np.log(10.0 * eps); hbar[:] = 0; leb[:] = 0; dac[:] = 0
      if it == n_warmup - 1:
        eps = np.exp(leb)
    else:
      draws[it - n_warmup] = th
      acc_sum += alpha
      n_div += div
  stats = dict(accept_rate=acc_sum / max(n_results, 1), step_size=eps, n_divergent=n_div,
               n_leapfrog=n_leap, minv=minv)
  return draws, stats
