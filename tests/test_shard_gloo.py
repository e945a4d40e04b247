"""N > 1 host path on CPU: world_size-2 gloo run of the sharded fit must equal
the single-process result bit for bit (chains / draws are keyed by global ids;
exactly one all-gather).  The engine is the oracle-backed FakeEngine."""
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np
import pandas as pd
import pytest

from causalimpact_b200 import shard

WORKER = r'''
import os, pickle, sys
import numpy as np, pandas as pd
ROOT = sys.argv[1]
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch.distributed as dist
from fake_engine import FakeEngine
import causalimpact_b200 as cib
from causalimpact_b200 import api
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
  dist.init_process_group("gloo")
fake = FakeEngine()
api._resolve_engine = lambda opts: fake
rng = np.random.default_rng(3)
n = 60
n_cov = int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else ""
seasonal = "seasonal" in mode
exchange = "columns" if "columns" in mode else "draws"
xs = 100 + np.cumsum(rng.normal(size=(n, n_cov)), axis=0); y = 1.2 * xs[:, 0] + rng.normal(size=n); y[40:] += 4
df = pd.DataFrame(np.column_stack([y, xs]), columns=["y"] + [f"x{i}" for i in range(n_cov)],
                  index=pd.date_range("2021-01-01", periods=n))
mo = cib.ModelOptions(seasons=[cib.Seasons(num_seasons=4), cib.Seasons(num_seasons=3, num_steps_per_season=2)]) \
    if seasonal else None
ci = cib.fit_causalimpact(df, (df.index[0], df.index[39]), (df.index[40], df.index[-1]), seed=(1, 2),
    model_options=mo, inference_options=cib.InferenceOptions(num_results=22),
    engine_options=cib.EngineOptions(num_chains=5, min_warmup=25, max_leapfrog=3,
                                     gibbs_min_warmup=10, exchange=exchange))
if int(os.environ.get("RANK", "0")) == 0:
  vals = [c for c in ci.series.columns if not c.endswith(("_start", "_end"))]
  pickle.dump(dict(series=ci.series[vals].values, summary=ci.summary.values,
                   level=np.asarray(ci.posterior_samples.level),
                   weights=np.asarray(ci.posterior_samples.weights),
                   seasonal=np.asarray(ci.posterior_samples.seasonal_levels),
                   drift=None if ci.posterior_samples.seasonal_drift_scales is None
                   else np.asarray(ci.posterior_samples.seasonal_drift_scales)), open(sys.argv[2], "wb"))
if world > 1:
  dist.destroy_process_group()
'''


def _run(world, out, n_cov=1, mode=""):
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
    f.write(WORKER)
    script = f.name
  env = dict(os.environ, OMP_NUM_THREADS="1")
  if world == 1:
    cmd = [sys.executable, script, root, out, str(n_cov), mode]
  else:
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29731",
           script, root, out, str(n_cov), mode]
  res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
  os.unlink(script)
  assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
  return pickle.load(open(out, "rb"))


def test_split_range_is_a_partition():
  for n in (1, 5, 64, 257):
    for w in (1, 2, 3, 8):
      parts = [shard.split_range(n, w, r) for r in range(w)]
      assert parts[0][0] == 0 and sum(c for _, c in parts) == n
      for (s0, c0), (s1, _) in zip(parts, parts[1:]):
        assert s0 + c0 == s1
      assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


@pytest.mark.parametrize("n_cov", [1, 4], ids=["hmc_path", "gibbs_path"])
def test_world2_gloo_equals_single_process(tmp_path, n_cov):
  """n_cov = 1 -> sampler auto = HMC; n_cov = 4 (p = 5 > 3) -> the Gibbs kernel path."""
  one = _run(1, str(tmp_path / "w1.pkl"), n_cov)
  two = _run(2, str(tmp_path / "w2.pkl"), n_cov)
  assert one.pop("drift") is None and two.pop("drift") is None      # no seasonal components
  for k in one:
    assert np.array_equal(one[k], two[k], equal_nan=True), k
  assert one["level"].shape == (22, 60)


def test_world2_gloo_equals_single_process_with_seasonal_components(tmp_path):
  """The seasonal branch of the fit (six gathered parts: theta, level, trajectory, latent,
  per-component contributions, drift variances) under a 2-rank gloo group == single process."""
  one = _run(1, str(tmp_path / "s1.pkl"), 1, "seasonal")
  two = _run(2, str(tmp_path / "s2.pkl"), 1, "seasonal")
  for k in one:
    assert np.array_equal(one[k], two[k], equal_nan=True), k
  assert one["seasonal"].shape == (22, 60, 2) and one["drift"].shape == (22, 2)


@pytest.mark.parametrize("n_cov,mode", [(1, "columns"), (4, "columns"), (1, "seasonal-columns")],
                         ids=["hmc_path", "gibbs_path", "seasonal"])
def test_world2_gloo_time_sharded_impact_equals_single_process(tmp_path, n_cov, mode):
  """EngineOptions.exchange = "columns": the trajectories are never gathered; the ranks swap
  time blocks of the transposed paths (two all-to-alls), select the quantiles of their half of
  the time axis, and sum the disjoint results (shard.impact_sharded).  22 draws from 5 chains of
  5 -> rank 0 holds 15 draws, rank 1 the remaining 7 (25 truncated to 22): ragged shards.
  Latent draws and every quantile column equal the single-process fit exactly; the
  mean-derived columns agree to the float32 rounding of the per-rank partial means."""
  one = _run(1, str(tmp_path / "c1.pkl"), n_cov, mode.replace("columns", "").strip("-"))
  two = _run(2, str(tmp_path / "c2.pkl"), n_cov, mode)
  for k in ("level", "weights", "seasonal", "drift"):
    if one[k] is not None:
      assert np.array_equal(one[k], two[k], equal_nan=True), k
  # series value columns: observed, posterior_{mean,lower,upper}, point_effects_*, cumulative_*
  a, b = one["series"].astype(np.float64), two["series"].astype(np.float64)
  mean_cols = [1, 4, 7]
  quant_cols = [c for c in range(a.shape[1]) if c not in mean_cols]
  assert np.array_equal(a[:, quant_cols], b[:, quant_cols], equal_nan=True)
  np.testing.assert_allclose(a[:, mean_cols], b[:, mean_cols], rtol=1e-5, atol=1e-4)
  np.testing.assert_allclose(one["summary"].astype(np.float64), two["summary"].astype(np.float64),
                             rtol=1e-5, atol=1e-4)


MANY_WORKER = r'''
import os, pickle, sys
import numpy as np, pandas as pd
ROOT = sys.argv[1]
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch.distributed as dist
from fake_engine import FakeEngine
import causalimpact_b200 as cib
from causalimpact_b200 import api
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
  dist.init_process_group("gloo")
fake = FakeEngine()
api._resolve_engine = lambda opts: fake
rng = np.random.default_rng(5)
idx = pd.date_range("2021-01-01", periods=50)
dfs = []
for s in range(5):
  x = 100 + np.cumsum(rng.normal(size=50)); y = (1 + 0.2 * s) * x + rng.normal(size=50); y[35:] += 3 + s
  dfs.append(pd.DataFrame({"y": y, "x": x}, index=idx))
res = cib.fit_causalimpact_many(dfs, (idx[0], idx[34]), (idx[35], idx[-1]), seed=(1, 2),
    inference_options=cib.InferenceOptions(num_results=12),
    engine_options=cib.EngineOptions(num_chains=3, gibbs_min_warmup=8))
mine = {i: dict(series=r.series[[c for c in r.series.columns if not c.endswith(("_start", "_end"))]].values,
                summary=r.summary.values) for i, r in enumerate(res) if r is not None}
pickle.dump(mine, open(sys.argv[2] + "." + os.environ.get("RANK", "0"), "wb"))
if world > 1:
  dist.destroy_process_group()
'''


def test_many_series_are_sharded_by_series_without_a_collective(tmp_path):
  """fit_causalimpact_many under a 2-rank gloo group: rank r owns a contiguous range of the
  series (None elsewhere), and every owned result equals the single-process one bit for bit."""
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  script = str(tmp_path / "many_worker.py")
  open(script, "w").write(MANY_WORKER)
  env = dict(os.environ, OMP_NUM_THREADS="1")
  one, two = str(tmp_path / "one"), str(tmp_path / "two")
  r = subprocess.run([sys.executable, script, root, one], capture_output=True, text=True, env=env,
                     timeout=600)
  assert r.returncode == 0, r.stderr[-3000:]
  r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                      "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29733",
                      script, root, two], capture_output=True, text=True, env=env, timeout=600)
  assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
  ref = pickle.load(open(one + ".0", "rb"))
  parts = [pickle.load(open(f"{two}.{k}", "rb")) for k in (0, 1)]
  assert sorted(ref) == [0, 1, 2, 3, 4]
  assert sorted(parts[0]) == [0, 1, 2] and sorted(parts[1]) == [3, 4]      # split_range(5, 2, r)
  for part in parts:
    for i, got in part.items():
      np.testing.assert_array_equal(got["series"], ref[i]["series"])
      np.testing.assert_array_equal(got["summary"], ref[i]["summary"])


def test_panel_host_logic_matches_the_frame_path(monkeypatch):
  """fit_causalimpact_panel's vectorised packaging (NaN rules, re-indexing, summary table) vs
  the pandas packaging of fit_causalimpact_many, both on the oracle-backed FakeEngine (CPU):
  a gap between the periods, a tail after the post-period, a missing pre-period point."""
  from fake_engine import FakeEngine
  import causalimpact_b200 as cib
  from causalimpact_b200 import api
  fake = FakeEngine()
  monkeypatch.setattr(api, "_resolve_engine", lambda opts: fake)
  rng = np.random.default_rng(8)
  idx = pd.date_range("2021-03-01", periods=60)
  dfs = []
  for s in range(3):
    x = 100 + np.cumsum(rng.normal(size=60)); y = (1 + 0.3 * s) * x + rng.normal(size=60)
    y[40:] += 4
    if s == 1:
      y[5] = np.nan
    dfs.append(pd.DataFrame({"y": y, "x": x}, index=idx))
  pre, post = (idx[0], idx[37]), (idx[41], idx[55])
  # decorrelate_series=False: fit_causalimpact_many then prepares every frame with pandas (the
  # default would route these stackable frames to the panel path itself)
  kw = dict(seed=(3, 1), inference_options=cib.InferenceOptions(num_results=10),
            engine_options=cib.EngineOptions(num_chains=2, gibbs_min_warmup=6,
                                             decorrelate_series=False))
  many = cib.fit_causalimpact_many(dfs, pre, post, **kw)
  assert not isinstance(many[0], cib.PanelAnalysis)
  res = cib.fit_causalimpact_panel(np.stack([d.values for d in dfs]), idx, pre, post,
                                   keep_level=True, **kw)
  vals = cib.impact.SERIES_VALUE_COLUMNS
  for i, one in enumerate(many):
    want = one.series[vals].values.astype(float)
    np.testing.assert_array_equal(np.isnan(res.series[i]), np.isnan(want))
    np.testing.assert_allclose(res.series[i], want, rtol=1e-5, atol=1e-5, equal_nan=True)
    np.testing.assert_allclose(res.summary[i], one.summary.values.astype(float), rtol=1e-5, atol=1e-6)
    ser, summ = res.frames(i)
    assert list(ser.columns) == list(one.series.columns) and ser.index.equals(one.series.index)
    assert list(summ.columns) == list(one.summary.columns) and list(summ.index) == list(one.summary.index)
    np.testing.assert_allclose(res.level[i], one.posterior_samples.level, rtol=1e-4, atol=1e-4)


def test_many_stackable_frames_take_the_panel_route(monkeypatch):
  """fit_causalimpact_many with frames that share index and columns (default options): ONE
  fit_causalimpact_panel call, results behind the CausalImpactAnalysis interface with frames built
  on first access -- equal to the panel call's arrays, same columns / index / sample shapes as the
  per-frame results; a named outcome column that is not first is moved first; frames that do not
  stack (another index) fall back to the per-frame path."""
  from fake_engine import FakeEngine
  import causalimpact_b200 as cib
  from causalimpact_b200 import api
  fake = FakeEngine()
  monkeypatch.setattr(api, "_resolve_engine", lambda opts: fake)
  rng = np.random.default_rng(18)
  idx = pd.date_range("2021-03-01", periods=50)
  dfs = []
  for s in range(3):
    x = 100 + np.cumsum(rng.normal(size=50)); y = (1 + 0.3 * s) * x + rng.normal(size=50)
    y[35:] += 4
    dfs.append(pd.DataFrame({"x": x, "y": y}, index=idx if s else idx.copy()))
  pre, post = (idx[0], idx[33]), (idx[35], idx[48])
  kw = dict(seed=(3, 1), inference_options=cib.InferenceOptions(num_results=8),
            engine_options=cib.EngineOptions(num_chains=2, gibbs_min_warmup=5))
  do = cib.DataOptions(outcome_column="y")
  many = cib.fit_causalimpact_many(dfs, pre, post, data_options=do, **kw)
  assert all(isinstance(m, cib.PanelAnalysis) and isinstance(m, cib.CausalImpactAnalysis) for m in many)
  res = cib.fit_causalimpact_panel(np.stack([d[["y", "x"]].values for d in dfs]), idx, pre, post,
                                   keep_level=True, **kw)
  slow = cib.fit_causalimpact_many(dfs, pre, post, data_options=do, seed=(3, 1),
                                   inference_options=cib.InferenceOptions(num_results=8),
                                   engine_options=cib.EngineOptions(num_chains=2, gibbs_min_warmup=5,
                                                                    decorrelate_series=False))
  vals = cib.impact.SERIES_VALUE_COLUMNS
  for i, m in enumerate(many):
    assert m._frames is None                                        # nothing built yet
    np.testing.assert_array_equal(m.series[vals].values, res.series[i])
    np.testing.assert_array_equal(m.summary.values, res.summary[i])
    assert list(m.series.columns) == list(slow[i].series.columns) and m.series.index.equals(idx)
    assert list(m.summary.columns) == list(slow[i].summary.columns)
    ps, qs = m.posterior_samples, slow[i].posterior_samples
    np.testing.assert_array_equal(ps.level, res.level[i])
    for f in ("observation_noise_scale", "level_scale", "level", "weights", "seasonal_levels"):
      assert getattr(ps, f).shape == getattr(qs, f).shape, f
    assert ps.seasonal_drift_scales is None and qs.seasonal_drift_scales is None
    assert m.diagnostics["inclusion"].shape == slow[i].diagnostics["inclusion"].shape
    assert ps.level.numpy().dtype == np.float32
  # another index in one frame: not stackable -> the per-frame path (and its own error / result)
  odd = [dfs[0], dfs[1].set_index(idx + pd.Timedelta(days=1))]
  with pytest.raises(Exception):
    cib.fit_causalimpact_many(odd, pre, (idx[35], idx[49]), data_options=do, **kw)
  # return_level=False: no level paths in the lazy record either
  nl = cib.fit_causalimpact_many(dfs[:2], pre, post, data_options=do, seed=1,
                                 inference_options=cib.InferenceOptions(num_results=4),
                                 engine_options=cib.EngineOptions(num_chains=2, gibbs_min_warmup=3,
                                                                  return_level=False))
  assert nl[0].posterior_samples.level is None and nl[0].posterior_samples.seasonal_levels is None


PIECES_WORKER = r'''
import os, sys
import numpy as np
ROOT = sys.argv[1]
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch, torch.distributed as dist
from causalimpact_b200 import shard
dist.init_process_group("gloo")
rank, ws = shard.world()
counts = [5, 0, 3][:ws] if ws == 3 else [4, 3]
T = 6
full = np.arange(sum(counts) * T, dtype=np.float32).reshape(sum(counts), T)
s0 = sum(counts[:rank])
local = torch.from_numpy(full[s0:s0 + counts[rank]].copy())
sd = shard.ShardedDraws(local, counts)
assert sd.shape == (sum(counts), T)
assert np.array_equal(np.asarray(sd), full)                      # gathered on demand, rank order
part = local.double().mean(0).float() if counts[rank] else torch.zeros(T)
sm = shard.ShardedMean(None, part, counts)
np.testing.assert_allclose(np.asarray(sm), full.mean(0), rtol=1e-6)
assert sm.shape == (T,)
# time blocks of ragged shards: every rank ends up with all draws of ITS columns, in rank order
splits = [shard.split_range(T, ws, r) for r in range(ws)]
mine, head = shard._exchange_columns(local.t().contiguous(), splits, counts, rank)
t0, tn = splits[rank]
assert head is None and np.array_equal(mine.numpy(), full.T[t0:t0 + tn])
dist.barrier(); dist.destroy_process_group()
if rank == 0:
  open(sys.argv[2], "w").write("ok")
'''


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_containers_and_column_exchange_under_gloo(tmp_path, world):
  """shard.ShardedDraws / ShardedMean (what a fit with exchange="columns" hands to the impact stage)
  and the all-to-all by time block, with ragged and EMPTY shards, 2 and 3 ranks."""
  script = tmp_path / "w.py"
  script.write_text(PIECES_WORKER)
  ok = tmp_path / "ok"
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29741",
                        str(script), root, str(ok)], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))
  assert res.returncode == 0 and ok.exists(), res.stdout[-2000:] + res.stderr[-4000:]
