"""GPU parity of the long-series team kernels (csrc/ci_team_stream.cuh): W warps per chain
walking the series in rounds of W tiles, tiles resident or streamed through the ring.
Checked against the float64 oracle and against the one-warp-per-chain path
(CI_B200_TSTREAM=0).  Tolerances as tests/test_gpu_logprob.py (float32: value 2e-5 rel +
2e-2 abs at T >= 5000 -- the log-lik is O(T) --, gradient 5e-3 rel + 5e-2; float64: 1e-9)."""
import numpy as np
import pytest

import causalimpact_b200 as cib
from conftest import make_series, make_thetas
from oracle import kalman_np as K

pytestmark = pytest.mark.gpu

CASES = [  # (name, T, n_cov, C, W)   W = CI_B200_TSW (0 = default)
    ("resident_12_tiles", 3000, 1, 9, 0),        # more tiles than a resident team, all in smem
    ("partial_last_round", 2304 + 5, 2, 7, 4),   # 10 tiles, W = 4: last round has 2 tiles
    ("w_eq_8", 5000, 3, 5, 8),
    ("w_eq_3_no_cov", 4000, 0, 4, 3),
    ("streamed_wide", 6000, 20, 6, 4),           # p > 16: transposed X'rbar through shared memory
]


def _engine_with(monkeypatch, **env):
  for k, v in env.items():
    monkeypatch.setenv(k, str(v))
  return cib.Engine(0)


@pytest.mark.parametrize("name,T,n_cov,C,W", CASES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_tstream_matches_oracle_and_one_warp_path(monkeypatch, name, T, n_cov, C, W, dtype):
  y, X, _ = make_series(T, n_cov, 500 + T, nan_frac=0.02)
  spec = cib.build_problem(y, X, dtype=dtype)
  prob = K.default_problem(y, X)
  th = make_thetas(spec.dim, spec.p, C, 3).astype(dtype).astype(np.float64)
  ov, og = K.log_post_grad(prob, th)
  out = {}
  for mode in ("1", "0"):
    eng = _engine_with(monkeypatch, CI_B200_TSTREAM=mode, CI_B200_TSW=W)
    eng.set_data(spec)
    out[mode] = eng.logprob_grad(th, with_prior=True)
    v_only = eng.logprob(th, with_prior=True)
    np.testing.assert_allclose(v_only, out[mode][0], rtol=1e-7, atol=1e-6)
    eng.close()
  rt_v, at_v, rt_g, at_g = (2e-5, 2e-2, 5e-3, 5e-2) if dtype == np.float32 else \
                           (1e-10, 1e-8, 1e-7, 1e-7)
  for mode in out:
    np.testing.assert_allclose(out[mode][0], ov, rtol=rt_v, atol=at_v)
    np.testing.assert_allclose(out[mode][1], og, rtol=rt_g, atol=at_g)


def test_tstream_result_does_not_depend_on_batch_position_or_size(monkeypatch):
  """A chain's value / gradient are bit-identical wherever it sits in the batch and however
  many chains share its CTA (teams per CTA change with the batch size)."""
  y, X, _ = make_series(5000, 1, 77)
  spec = cib.build_problem(y, X)
  eng = _engine_with(monkeypatch, CI_B200_TSTREAM="1")
  eng.set_data(spec)
  th = make_thetas(spec.dim, spec.p, 700, 5).astype(np.float32).astype(np.float64)
  v, g = eng.logprob_grad(th, with_prior=True)
  perm = np.random.default_rng(1).permutation(700)
  v2, g2 = eng.logprob_grad(th[perm], with_prior=True)
  assert np.array_equal(v2, v[perm]) and np.array_equal(g2, g[perm])
  v3, g3 = eng.logprob_grad(th[:5], with_prior=True)
  assert np.array_equal(v3, v[:5]) and np.array_equal(g3, g[:5])
  eng.close()


def test_tstream_hmc_long_series(monkeypatch):
  """Persistent HMC over the long-series team evaluator: hundreds of evaluations through the
  same ring (the stream position must stay in step with the producer); draws agree in
  distribution with the one-warp HMC kernel."""
  y, X, _ = make_series(3000, 1, 31)
  spec = cib.build_problem(y, X)
  th0 = np.tile(cib.initial_theta(spec), (48, 1))
  res = {}
  for mode in ("1", "0"):
    eng = _engine_with(monkeypatch, CI_B200_TSTREAM=mode)
    eng.set_data(spec)
    res[mode] = eng.hmc_run(th0, n_warmup=150, n_results=60, seed=5, init_step=0.01)
    eng.close()
  for mode, (draws, stats) in res.items():
    assert np.all(np.isfinite(draws)), mode
    assert 0.5 < stats["accept_rate"].mean() < 0.99, mode
  a = res["1"][0].reshape(-1, spec.dim); b = res["0"][0].reshape(-1, spec.dim)
  for j in range(spec.dim):
    se = np.sqrt(a[:, j].var() / 200 + b[:, j].var() / 200)      # ~200 effective draws each
    assert abs(a[:, j].mean() - b[:, j].mean()) < 6 * se + 1e-3, j
