"""GPU parity for K6 (persistent batched-chain HMC) through the C ABI.

* float64 kernels vs oracle/hmc_np.py PATHWISE (same Philox streams, same
  adaptation schedule): draws agree to 1e-6 over a short run.
* float32 kernels STATISTICALLY: posterior means / sds of a longer run agree
  with the float64 oracle HMC and with the restated reference Gibbs sampler
  within max(5 combined MC standard errors, 2e-3)  (SURVEY section 8c).
* determinism: same seed -> bit-identical draws; chain c depends only on
  (seed, global chain id), not on the batch split (multi-GPU sharding).
"""
import numpy as np
import pytest

import causalimpact_b200 as cib
from conftest import make_series
from oracle import c_port
from oracle import gibbs_np as G
from oracle import hmc_np as H
from oracle import kalman_np as K

pytestmark = pytest.mark.gpu


def _setup(engine, T, n_cov, seed, dtype, prior_level_sd=0.01):
  y, X, _ = make_series(T, n_cov, seed)
  spec = cib.build_problem(y, X, prior_level_sd=prior_level_sd, dtype=dtype)
  engine.set_data(spec)
  prob = K.default_problem(y, X, prior_level_sd=prior_level_sd)
  return spec, prob


def _init(spec, C, seed, prior_level_sd=0.01):
  rng = np.random.default_rng(seed)
  th0 = np.tile(cib.initial_theta(spec, prior_level_sd), (C, 1))
  th0[:, :spec.p] += 0.1 * rng.normal(size=(C, spec.p))
  return th0


@pytest.mark.parametrize("T,n_cov", [(100, 1), (300, 3)])
def test_hmc_float64_matches_oracle_pathwise(engine, T, n_cov):
  spec, prob = _setup(engine, T, n_cov, 21, np.float64)
  th0 = _init(spec, 6, 1)
  kw = dict(n_warmup=40, n_results=15, seed=77, max_leapfrog=5, init_step=0.02)
  draws, stats = engine.hmc_run(th0, chain_id0=3, **kw)
  f = lambda th: c_port.logpost_grad(prob, th)[:2]
  od, ost = H.run(f, th0, chain_id0=3, **kw)
  np.testing.assert_allclose(draws, od, rtol=1e-6, atol=1e-6)
  np.testing.assert_allclose(stats["step_size"], ost["step_size"], rtol=1e-5)
  np.testing.assert_allclose(stats["accept_rate"], ost["accept_rate"], atol=1e-5)
  assert np.array_equal(stats["n_leapfrog"], ost["n_leapfrog"])


def test_hmc_float32_statistical_parity(engine):
  """BASELINE config 1 (quickstart: local level + 1 covariate, T=100)."""
  spec, prob = _setup(engine, 100, 1, 7, np.float32)
  C = 64
  th0 = _init(spec, C, 0)
  draws, stats = engine.hmc_run(th0, n_warmup=400, n_results=300, seed=5, max_leapfrog=8,
                                init_step=0.05)
  assert np.all(np.isfinite(draws))
  assert stats["n_divergent"].sum() <= 0.01 * C * 300
  assert 0.6 < stats["accept_rate"].mean() < 0.99
  d = draws.reshape(-1, spec.dim)
  # float64 oracle HMC (16 chains) and restated reference Gibbs (1 chain)
  f = lambda th: c_port.logpost_grad(prob, th)[:2]
  od, _ = H.run(f, th0[:16], n_warmup=400, n_results=300, seed=6, max_leapfrog=8, init_step=0.05)
  od = od.reshape(-1, spec.dim)
  gb = G.run(prob, n_results=4000, n_warmup=500, seed=3)
  p = spec.p
  quantities = {
      "w0": (d[:, 0], od[:, 0], gb["w"][:, 0]),
      "sigma_obs": (np.exp(d[:, p] / 2), np.exp(od[:, p] / 2), np.sqrt(gb["s_e"])),
      "sigma_level": (np.exp(d[:, p + 1] / 2), np.exp(od[:, p + 1] / 2), np.sqrt(gb["s_h"])),
  }
  for name, (a, b, c) in quantities.items():
    # effective sample sizes are conservatively taken as n/20
    for other in (b, c):
      se = np.sqrt(a.var() / (a.size / 20) + other.var() / (other.size / 20))
      assert abs(a.mean() - other.mean()) < max(5 * se, 2e-3), (name, a.mean(), other.mean(), se)
      assert abs(a.std() / other.std() - 1) < 0.2, (name, a.std(), other.std())


def test_hmc_deterministic_and_split_invariant(engine):
  spec, _ = _setup(engine, 300, 2, 5, np.float32)
  th0 = _init(spec, 12, 2)
  kw = dict(n_warmup=30, n_results=10, seed=123, max_leapfrog=4, init_step=0.03)
  d1, s1 = engine.hmc_run(th0, **kw)
  d2, s2 = engine.hmc_run(th0, **kw)
  assert np.array_equal(d1, d2) and np.array_equal(s1, s2)
  da, _ = engine.hmc_run(th0[:5], chain_id0=0, **kw)
  db, _ = engine.hmc_run(th0[5:], chain_id0=5, **kw)
  assert np.array_equal(d1, np.concatenate([da, db], axis=1))
  d3, _ = engine.hmc_run(th0, **dict(kw, seed=124))
  assert not np.array_equal(d1, d3)


def test_hmc_streaming_tiles(engine):
  """Long series: tiles stream through the mbarrier ring during the whole run."""
  spec, prob = _setup(engine, 20000, 1, 4, np.float64)
  th0 = _init(spec, 3, 3)
  kw = dict(n_warmup=6, n_results=4, seed=9, max_leapfrog=3, init_step=0.01)
  draws, stats = engine.hmc_run(th0, **kw)
  f = lambda th: c_port.logpost_grad(prob, th)[:2]
  od, _ = H.run(f, th0, **kw)
  np.testing.assert_allclose(draws, od, rtol=1e-6, atol=1e-6)


def test_hmc_rejects_bad_options(engine):
  spec, _ = _setup(engine, 100, 1, 7, np.float32)
  th0 = _init(spec, 2, 0)
  with pytest.raises(cib.EngineError):
    engine.hmc_run(th0, n_warmup=10, n_results=0, seed=1)
  with pytest.raises(cib.EngineError):
    engine.hmc_run(th0, n_warmup=10, n_results=5, seed=1, init_step=-1.0)
