"""GPU parity: CUDA Kalman log-prob / gradient vs the float64 oracle, through
the C ABI.  Tolerances (float32 kernels): value  |d| <= 2e-5*|ll| + 2e-3,
gradient  |d| <= 2e-3*|g| + 2e-2  (float64 kernels: 1e-9 / 1e-7 relative)."""
import numpy as np
import pytest

import causalimpact_b200 as cib
from conftest import make_series, make_thetas
from oracle import kalman_np as K

pytestmark = pytest.mark.gpu

CASES = [  # (name, T, n_cov, C)
    ("quickstart", 100, 1, 8),          # BASELINE config 1
    ("one_tile_edge", 256, 2, 5),
    ("ragged", 257, 0, 3),              # no covariates, T = TB + 1
    ("cfg2", 1000, 10, 256),            # BASELINE config 2
    ("wide", 600, 40, 33),              # p > 32: two covariate slots per lane
    ("tiny", 3, 1, 2),
]


@pytest.mark.parametrize("name,T,n_cov,C", CASES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("variant", [1, 0], ids=["scan", "seq"])
def test_logprob_and_grad_match_oracle(engine, name, T, n_cov, C, dtype, variant):
  y, X, _ = make_series(T, n_cov, 100 + T, nan_frac=0.02 if T > 50 else 0.0)
  spec = cib.build_problem(y, X, prior_level_sd=0.01, dtype=dtype)
  engine.set_data(spec)
  prob = K.default_problem(y, X, prior_level_sd=0.01)
  th = make_thetas(spec.dim, spec.p, C, 7).astype(dtype).astype(np.float64)
  for with_prior in (False, True):
    val, grad = engine.logprob_grad(th, with_prior=with_prior, variant=variant)
    v_only = engine.logprob(th, with_prior=with_prior, variant=variant)
    if with_prior:
      ov, og = K.log_post_grad(prob, th)
    else:
      ov, og = K.log_lik_grad(prob, th)
    rt_v, at_v, rt_g, at_g = (2e-5, 2e-3, 2e-3, 2e-2) if dtype == np.float32 else \
                             (1e-10, 1e-8, 1e-7, 1e-7)
    np.testing.assert_allclose(val, ov, rtol=rt_v, atol=at_v)
    np.testing.assert_allclose(v_only, val, rtol=1e-7, atol=1e-6)
    np.testing.assert_allclose(grad, og, rtol=rt_g, atol=at_g)


def test_streaming_pipeline_long_series(engine):
  """T = 20000 (BASELINE config 4): tiles do not fit in shared memory, so the
  mbarrier ring streams them; checked on a subset of chains against the oracle."""
  y, X, _ = make_series(20000, 1, 2024)
  spec = cib.build_problem(y, X)
  engine.set_data(spec)
  prob = K.default_problem(y, X)
  th = make_thetas(spec.dim, spec.p, 24, 9).astype(np.float32).astype(np.float64)
  ov, og = K.log_post_grad(prob, th)
  for variant in (1, 0):
    val, grad = engine.logprob_grad(th, with_prior=True, variant=variant)
    np.testing.assert_allclose(val, ov, rtol=2e-5, atol=2e-2)
    np.testing.assert_allclose(grad, og, rtol=5e-3, atol=5e-2)


def test_out_of_support_and_bad_inputs(engine):
  y, X, _ = make_series(300, 2, 5)
  spec = cib.build_problem(y, X)
  engine.set_data(spec)
  th = make_thetas(spec.dim, spec.p, 4, 1)
  th[0, spec.p] = np.log(spec.ub_variance(spec.obs_ub) * 1.1)
  v = engine.logprob(th, with_prior=True)
  assert np.isneginf(v[0]) and np.all(np.isfinite(v[1:]))
  with pytest.raises(ValueError):
    engine.logprob(th[:, :-1])
  with pytest.raises(cib.EngineError):
    engine.logprob(th, variant=7)


def test_linearity_property_full_size(engine):
  """Size-independent property at BASELINE config-2 size: with sigma's fixed the
  log-lik is quadratic in w, so the gradient is affine along any line."""
  y, X, _ = make_series(1000, 10, 2022)
  spec = cib.build_problem(y, X, dtype=np.float64)
  engine.set_data(spec)
  rng = np.random.default_rng(0)
  t0 = make_thetas(spec.dim, spec.p, 1, 3)[0]
  dirn = np.zeros(spec.dim); dirn[:spec.p] = rng.normal(size=spec.p)
  ths = np.stack([t0 + s * dirn for s in (0.0, 0.5, 1.0)])
  _, g = engine.logprob_grad(ths)
  np.testing.assert_allclose(g[1, :spec.p], 0.5 * (g[0, :spec.p] + g[2, :spec.p]),
                             rtol=1e-8, atol=1e-8)


def test_team_and_single_warp_paths_agree(monkeypatch):
  """T=1000 has 4 tiles: the default is TEAM mode (one warp per tile); with
  CI_B200_TEAM=0 the same problem runs on the one-warp-per-chain path.  Both
  must match the oracle, and each other to float32 rounding."""
  y, X, _ = make_series(1000, 10, 2023, nan_frac=0.02)
  spec = cib.build_problem(y, X)
  prob = K.default_problem(y, X)
  th = make_thetas(spec.dim, spec.p, 70, 11).astype(np.float32).astype(np.float64)
  out = {}
  for mode in ("1", "0"):
    monkeypatch.setenv("CI_B200_TEAM", mode)
    eng = cib.Engine(0)
    eng.set_data(spec)
    out[mode] = eng.logprob_grad(th, with_prior=True)
    eng.close()
  ov, og = K.log_post_grad(prob, th)
  for mode in out:
    np.testing.assert_allclose(out[mode][0], ov, rtol=2e-5, atol=2e-3)
    np.testing.assert_allclose(out[mode][1], og, rtol=2e-3, atol=2e-2)
  np.testing.assert_allclose(out["1"][0], out["0"][0], rtol=1e-5, atol=1e-3)


def test_pinned_zero_copy_path_matches_staged_copies(engine):
  """Host-pointer ABI: pinned caller buffers are read / written by the kernel
  directly (zero-copy), pageable ones go through staged copies -- same bits."""
  import torch
  from causalimpact_b200 import _engine
  y, X, _ = make_series(1000, 10, 2022)
  spec = cib.build_problem(y, X)
  engine.set_data(spec)
  th = make_thetas(spec.dim, spec.p, 64, 3).astype(np.float32)
  v_staged, g_staged = engine.logprob_grad(th.astype(np.float64), with_prior=True)
  th_pin = torch.from_numpy(th).pin_memory()
  v_pin = torch.empty(64, dtype=torch.float32).pin_memory()
  g_pin = torch.empty(64, spec.dim, dtype=torch.float32).pin_memory()
  engine.logprob_grad_ptr(th_pin.data_ptr(), 64, v_pin.data_ptr(), g_pin.data_ptr(),
                          _engine.VARIANT_SCAN, _engine.WITH_PRIOR, host=True)
  assert np.array_equal(v_pin.numpy(), v_staged) and np.array_equal(g_pin.numpy(), g_staged)
