"""CPU-only: the oracle HMC targets the same posterior as the restated
reference Gibbs sampler (inclusion probability 1 regime), and the adaptation
schedule follows Stan's windowing."""
import numpy as np

from conftest import make_series
from oracle import c_port
from oracle import gibbs_np as G
from oracle import hmc_np as H
from oracle import kalman_np as K
from oracle import philox_np as PH


def test_adapt_schedule():
  assert H.adapt_schedule(1000) == (75, 950, [99, 149, 249, 449, 949])
  assert H.adapt_schedule(100) == (15, 90, [89])
  assert H.adapt_schedule(300) == (75, 250, [99, 149, 249])
  assert H.adapt_schedule(10) == (10, 10, [])


def test_leapfrog_count_range():
  Ls = [PH.leapfrog_count(42, it, 8) for it in range(2000)]
  assert min(Ls) == 1 and max(Ls) == 8
  assert abs(np.mean(Ls) - 4.5) < 0.2


def test_hmc_oracle_agrees_with_restated_reference_gibbs():
  y, X, _ = make_series(100, 1, 7)
  prob = K.default_problem(y, X, prior_level_sd=0.01)
  f = lambda th: c_port.logpost_grad(prob, th)[:2]
  rng = np.random.default_rng(0)
  th0 = np.tile(K.initial_theta(prob), (16, 1))
  th0[:, :prob.p] += 0.1 * rng.normal(size=(16, prob.p))
  draws, st = H.run(f, th0, n_warmup=300, n_results=300, seed=11, max_leapfrog=8, init_step=0.05)
  assert st["n_divergent"].sum() == 0
  d = draws.reshape(-1, prob.dim)
  g = G.run(prob, n_results=4000, n_warmup=500, seed=3)
  p = prob.p
  pairs = [(d[:, 0], g["w"][:, 0]),
           (np.exp(d[:, p] / 2), np.sqrt(g["s_e"])),
           (np.exp(d[:, p + 1] / 2), np.sqrt(g["s_h"]))]
  for a, b in pairs:
    se = np.sqrt(a.var() / (a.size / 20) + b.var() / (b.size / 20))
    assert abs(a.mean() - b.mean()) < max(5 * se, 2e-3)
    assert abs(a.std() / b.std() - 1) < 0.2


def test_spike_and_slab_oracle_selects_true_features():
  """The restated spike-and-slab sweep (Scott & Varian marginal, inclusion prior
  3/p as in causalimpact_lib.py:449-450) keeps the 3 true covariates and drops
  the 7 noise ones; with inclusion probability 1 it reduces to the dense sweep."""
  y, X, _ = make_series(300, 10, 2022)
  prob = K.default_problem(y, X)
  sp = G.run(prob, n_results=600, n_warmup=200, seed=1, sparse=True)
  inc = (sp["w"] != 0).mean(0)
  assert np.all(inc[:3] > 0.9) and np.all(inc[3:10] < 0.2)
  y2, X2, _ = make_series(100, 1, 7)                      # p = 2 -> inclusion prob 1
  prob2 = K.default_problem(y2, X2)
  a = G.run(prob2, n_results=1500, n_warmup=200, seed=2, sparse=True)
  b = G.run(prob2, n_results=1500, n_warmup=200, seed=3, sparse=False)
  assert np.all(a["w"] != 0)
  assert abs(a["w"][:, 0].mean() - b["w"][:, 0].mean()) < 0.02
  assert abs(np.sqrt(a["s_e"]).mean() - np.sqrt(b["s_e"]).mean()) < 0.01
