"""GPU parity for the Gibbs kernel with spike-and-slab regression -- the
reference's own sampler (causalimpact_lib.py:365-388) on the B200 -- against the
NumPy restatement oracle/gibbs_np.py.

RNG streams differ (the oracle uses PCG64; TFP's stream is unobtainable), so
the comparison is STATISTICAL: posterior means within max(5 combined MC
standard errors, floor), inclusion frequencies within 0.1; plus exact
invariants (inactive weights are exactly 0, clamps, determinism, independence
of the chain split)."""
import numpy as np
import pytest

import causalimpact_b200 as cib
from conftest import make_series
from oracle import gibbs_np as G
from oracle import kalman_np as K

pytestmark = pytest.mark.gpu


def _close(a, b, floor, ess_a=30.0, ess_b=30.0):
  se = np.sqrt(a.var() / max(a.size / ess_a, 2) + b.var() / max(b.size / ess_b, 2))
  return abs(a.mean() - b.mean()) < max(5 * se, floor), (a.mean(), b.mean(), se)


def test_spike_and_slab_matches_restated_reference_sampler(engine):
  """BASELINE config-2 family: 10 covariates + intercept, inclusion prob 3/11."""
  y, X, _ = make_series(300, 10, 2022)
  spec = cib.build_problem(y, X)
  engine.set_data(spec)
  prob = K.default_problem(y, X)
  C, n = 64, 150
  draws, level, traj, incl = engine.gibbs_run(C, n_warmup=300, n_results=n, seed=7, sparse=True)
  assert draws.shape == (n, C, spec.dim) and level.shape == (n, C, spec.T) == traj.shape
  assert np.all(np.isfinite(draws)) and np.all(np.isfinite(level)) and np.all(np.isfinite(traj))
  d = draws.reshape(-1, spec.dim)
  w = d[:, :spec.p]
  ob = G.run(prob, n_results=4000, n_warmup=500, seed=1, sparse=True)
  inc_gpu, inc_ref = incl.mean(0), (ob["w"] != 0).mean(0)
  np.testing.assert_allclose(inc_gpu, inc_ref, atol=0.1)
  np.testing.assert_allclose((w != 0).mean(0), inc_gpu, atol=1e-6)   # zeros are EXACT zeros
  for j in np.flatnonzero(inc_ref > 0.9):
    ok, info = _close(w[:, j], ob["w"][:, j], 5e-3)
    assert ok, (j, info)
  for a, b, name in ((np.exp(d[:, spec.p] / 2), np.sqrt(ob["s_e"]), "sigma_obs"),
                     (np.exp(d[:, spec.p + 1] / 2), np.sqrt(ob["s_h"]), "sigma_level")):
    ok, info = _close(a, b, 2e-3)
    assert ok, (name, info)
  # the reference's upper bounds limit the VARIANCES (ci_problem.ub_on_scale = 0)
  assert np.all(np.exp(d[:, spec.p]) <= spec.obs_ub * (1 + 1e-6))
  assert np.all(np.exp(d[:, spec.p + 1]) <= spec.lvl_ub * (1 + 1e-6))
  # counterfactual (level + X w) over the masked post-period
  post = np.isnan(y); post[:210] = False
  loc_gpu = (level.reshape(-1, spec.T) + w @ X.T)[:, post].mean(1)
  loc_ref = (ob["level"] + ob["w"] @ X.T)[:, post].mean(1)
  ok, info = _close(loc_gpu, loc_ref, 1e-2)
  assert ok, info
  assert abs(loc_gpu.std() / loc_ref.std() - 1) < 0.25
  # predictive draws = loc + sigma_obs noise
  noise = traj.reshape(-1, spec.T) - level.reshape(-1, spec.T) - w @ X.T
  assert abs(noise.var() / np.exp(d[:, spec.p]).mean() - 1) < 0.05


def test_dense_mode_matches_hmc_target(engine):
  """sparse=False on the quickstart shape (p = 2 <= 3: the reference's prior IS the
  slab): Gibbs kernel, oracle Gibbs and the HMC kernel sample one posterior."""
  y, X, _ = make_series(100, 1, 7)
  spec = cib.build_problem(y, X)
  engine.set_data(spec)
  prob = K.default_problem(y, X)
  draws, level, _, incl = engine.gibbs_run(64, n_warmup=300, n_results=100, seed=3, sparse=False,
                                           want_traj=False)
  assert np.all(incl == 1.0)
  d = draws.reshape(-1, spec.dim)
  ob = G.run(prob, n_results=4000, n_warmup=500, seed=2, sparse=False)
  th0 = np.tile(cib.initial_theta(spec), (64, 1))
  hd, _ = engine.hmc_run(th0, n_warmup=400, n_results=100, seed=5)
  h = hd.reshape(-1, spec.dim)
  p = spec.p
  for a, b, c, name in ((d[:, 0], ob["w"][:, 0], h[:, 0], "w0"),
                        (np.exp(d[:, p] / 2), np.sqrt(ob["s_e"]), np.exp(h[:, p] / 2), "sigma_obs"),
                        (np.exp(d[:, p + 1] / 2), np.sqrt(ob["s_h"]), np.exp(h[:, p + 1] / 2),
                         "sigma_level")):
    for other in (b, c):
      ok, info = _close(a, other, 2e-3)
      assert ok, (name, info)


def test_no_covariates(engine):
  y, _, _ = make_series(200, 0, 4)
  spec = cib.build_problem(y, None, prior_level_sd=0.1)
  engine.set_data(spec)
  prob = K.default_problem(y, None, prior_level_sd=0.1)
  draws, level, traj, _ = engine.gibbs_run(32, n_warmup=200, n_results=100, seed=9)
  d = draws.reshape(-1, spec.dim)
  ob = G.run(prob, n_results=3000, n_warmup=300, seed=4, prior_level_sd=0.1)
  for a, b, name in ((np.exp(d[:, 0] / 2), np.sqrt(ob["s_e"]), "sigma_obs"),
                     (np.exp(d[:, 1] / 2), np.sqrt(ob["s_h"]), "sigma_level")):
    ok, info = _close(a, b, 2e-3)
    assert ok, (name, info)
  obs = ~np.isnan(y)
  ok, info = _close(level.reshape(-1, spec.T)[:, obs].mean(1), ob["level"][:, obs].mean(1), 5e-3)
  assert ok, info


def test_gibbs_deterministic_and_split_invariant(engine):
  y, X, _ = make_series(300, 5, 11)
  spec = cib.build_problem(y, X)
  engine.set_data(spec)
  kw = dict(n_warmup=20, n_results=6, seed=42)
  a = engine.gibbs_run(10, **kw)
  b = engine.gibbs_run(10, **kw)
  for x, z in zip(a, b):
    assert np.array_equal(x, z)
  lo = engine.gibbs_run(4, chain_id0=0, **kw)
  hi = engine.gibbs_run(6, chain_id0=4, **kw)
  assert np.array_equal(a[0], np.concatenate([lo[0], hi[0]], axis=1))
  assert np.array_equal(a[1], np.concatenate([lo[1], hi[1]], axis=1))
  c = engine.gibbs_run(10, **dict(kw, seed=43))
  assert not np.array_equal(a[0], c[0])


def test_gibbs_float64_and_streaming(engine):
  """float64 kernels, and a long series whose tiles stream through the ring."""
  y, X, _ = make_series(5000, 1, 5)
  for dt in (np.float64, np.float32):
    spec = cib.build_problem(y, X, dtype=dt)
    engine.set_data(spec)
    draws, level, _, _ = engine.gibbs_run(3, n_warmup=10, n_results=4, seed=1, want_traj=False)
    assert np.all(np.isfinite(draws)) and np.all(np.isfinite(level))
    sig = np.exp(draws[..., spec.p] / 2)
    assert 0.05 < sig.mean() < 1.2


def test_gibbs_rejects_bad_options(engine):
  y, X, _ = make_series(100, 1, 7)
  engine.set_data(cib.build_problem(y, X))
  with pytest.raises(cib.EngineError):
    engine.gibbs_run(2, n_warmup=5, n_results=0, seed=1)
  with pytest.raises(cib.EngineError):
    engine.gibbs_run(2, n_warmup=5, n_results=2, seed=1, nonzero_prob=1.5)


def test_team_and_one_warp_gibbs_paths_agree(monkeypatch):
  """T <= 2048 runs the TEAM sweep (one warp per tile, speculative inclusion draws); with
  CI_B200_GIBBS_TEAM=0 the same problem runs one warp per chain.  Same Philox keys, same
  decisions: the two differ only by the summation order of the sufficient statistics, so the
  first sweeps agree to float32 rounding (later ones drift apart chaotically) and long runs
  agree in distribution.  Checked with the spike-and-slab prior and both visiting orders."""
  y, X, _ = make_series(700, 10, 2030, nan_frac=0.02)
  spec = cib.build_problem(y, X)
  res = {}
  for mode in ("1", "0"):
    monkeypatch.setenv("CI_B200_GIBBS_TEAM", mode)
    eng = cib.Engine(0)
    eng.set_data(spec)
    res[mode] = dict(
        short=eng.gibbs_run(24, n_warmup=0, n_results=2, seed=5, sparse=True),
        short_idx=eng.gibbs_run(24, n_warmup=0, n_results=2, seed=5, sparse=True, ssvs_order="index"),
        long=eng.gibbs_run(48, n_warmup=200, n_results=100, seed=9, sparse=True))
    eng.close()
  for key in ("short", "short_idx"):
    a, b = res["1"][key], res["0"][key]
    np.testing.assert_array_equal(a[0][..., :spec.p] != 0, b[0][..., :spec.p] != 0)   # same inclusions
    np.testing.assert_allclose(a[0], b[0], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(a[1], b[1], rtol=2e-3, atol=5e-3)                      # level paths
    np.testing.assert_array_equal(a[3], b[3])
  a, b = res["1"]["long"], res["0"]["long"]
  np.testing.assert_allclose(a[3].mean(0), b[3].mean(0), atol=0.08)                   # inclusion freq
  da, db = a[0].reshape(-1, spec.dim), b[0].reshape(-1, spec.dim)
  for j in (0, 1, 2, spec.p, spec.p + 1):
    ok, info = _close(da[:, j], db[:, j], 5e-3)
    assert ok, (j, info)


def test_team_gibbs_edge_shapes(engine):
  """Team sweep at its limits: a single tile (W = 1), 8 tiles (T = 2048), a ragged last tile,
  no covariates, float64; finite draws, clamps respected, deterministic."""
  for T, n_cov, dt in ((200, 3, np.float32), (2048, 2, np.float32), (1800, 0, np.float32),
                       (600, 12, np.float64)):
    y, X, _ = make_series(T, n_cov, 40 + T)
    spec = cib.build_problem(y, X, dtype=dt)
    engine.set_data(spec)
    a = engine.gibbs_run(9, n_warmup=15, n_results=5, seed=3)
    b = engine.gibbs_run(9, n_warmup=15, n_results=5, seed=3)
    for x, z in zip(a, b):
      assert np.array_equal(x, z)
    assert all(np.all(np.isfinite(x)) for x in a[:3])
    assert np.all(np.exp(a[0][..., spec.p]) <= spec.obs_ub * (1 + 1e-6))
    lo = engine.gibbs_run(4, n_warmup=15, n_results=5, seed=3)
    assert np.array_equal(a[0][:, :4], lo[0]) and np.array_equal(a[1][:, :4], lo[1])
