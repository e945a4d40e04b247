"""``fit_causalimpact`` on the B200 engine -- same entry point, option structs and
result object as the reference (causalimpact/causalimpact_lib.py:44-339).

What changed underneath (BASELINE.json north_star): the TFP Gibbs call
(causalimpact_lib.py:365-388) is replaced by batched-chain HMC over the CUDA
Kalman log-prob kernel (ci_hmc_run), the per-draw predictive sampling
(causalimpact_lib.py:609-632) by the CUDA simulation smoother
(ci_posterior_predict), and the per-time quantiles
(posterior_processing.py:25-60) by ci_row_quantiles.  No TensorFlow.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import List, Optional, Tuple, Union

import numpy as np
import pandas as pd

from . import frame as _frame
from . import impact as _impact
from . import shard as _shard
from ._engine import DeviceArray, Engine, ProblemSpec
from .model import build_problem, build_seasonal, initial_theta


class Samples(np.ndarray):
  """ndarray that also answers ``.numpy()`` (the reference returns tf.Tensors and
  its users call ``.numpy()``, e.g. causalimpact_lib_test.py:269, 335-338)."""

  def __new__(cls, arr):
    return np.asarray(arr).view(cls)

  def numpy(self) -> np.ndarray:
    return np.asarray(self)


@dataclasses.dataclass
class CausalImpactPosteriorSamples:
  """Draws of the model's latents (reference causalimpact_lib.py:44-58)."""
  observation_noise_scale: Samples            # [S]
  level_scale: Optional[Samples]              # [S]
  level: Optional[Samples]                    # [S, T]
  weights: Optional[Samples]                  # [S, covariates + 1 intercept] or None
  seasonal_drift_scales: Optional[Samples]    # None (no seasonal components)
  seasonal_levels: Optional[Samples]          # [S, T, 0]
  slope_scale: Optional[Samples] = None       # [S]; extension: EngineOptions.local_linear_trend


@dataclasses.dataclass
class CausalImpactAnalysis:
  """``series`` / ``summary`` frames + latent draws (reference :61-144).

  ``diagnostics`` is extra (not in the reference): per-chain HMC statistics.
  """
  series: pd.DataFrame
  summary: pd.DataFrame
  posterior_samples: CausalImpactPosteriorSamples
  diagnostics: Optional[dict] = None


@dataclasses.dataclass
class DataOptions:
  """reference :147-159.  ``dtype``: numpy float32 / float64 (or anything
  np.dtype() understands, or an object with ``as_numpy_dtype`` such as a tf dtype)."""
  outcome_column: Optional[str] = None
  standardize_data: bool = True
  dtype: object = np.float32


@dataclasses.dataclass(frozen=True)
class Seasons:
  """reference :162-180.  One seasonal effect (e.g. day of week): ``num_seasons`` effects,
  each lasting ``num_steps_per_season`` steps (int, per-season tuple, or per-cycle tuple of
  tuples).  Sampled by the seasonal Gibbs kernel (csrc/ci_seasonal.cuh); the sum of
  num_seasons over all components is limited to 191 (week-of-year and hour-of-week fit), at most 7
  components."""
  num_seasons: int
  num_steps_per_season: Union[int, Tuple[int], Tuple[Tuple[int]]] = 1


@dataclasses.dataclass
class ModelOptions:
  """reference :183-203."""
  prior_level_sd: float = 0.01
  seasons: List[Seasons] = dataclasses.field(default_factory=list)


@dataclasses.dataclass
class InferenceOptions:
  """reference :206-220 (warm-up defaults to ceil(num_results / 9))."""
  num_results: int = 900
  num_warmup_steps: Optional[int] = None

  def __post_init__(self):
    if self.num_warmup_steps is None:
      self.num_warmup_steps = math.ceil(self.num_results / 9)


@dataclasses.dataclass
class EngineOptions:
  """Knobs of the B200 engine (new; passed as ``engine_options=`` kwarg).

  num_chains: HMC chains (one warp each), sharded over the ranks of an
    initialised torch.distributed process group.
  min_warmup: HMC needs more adaptation than a Gibbs sweep count suggests; the
    warm-up run is max(InferenceOptions.num_warmup_steps, min_warmup).
  sampler: "auto" (default) = the reference's posterior: batched-chain HMC when the
    spike-and-slab inclusion probability min(1, 3/p) is 1 (<= 2 covariates -- the prior
    is then the continuous slab and HMC targets exactly the reference's posterior),
    the GPU Gibbs kernel with spike-and-slab otherwise; "hmc" forces HMC on the
    slab-only model, "gibbs" forces the Gibbs kernel.
  upper_bound_on: what the reference's ``prior.upper_bound`` attributes (lib.py:432,
    442-443, 474) limit: "variance" (default; TFP clips the sampled variance with
    ``min(variance, upper_bound)``) or "scale" (sigma <= bound; the round-1 reading).
  ssvs_order: order in which a Gibbs sweep visits the spike-and-slab inclusion indicators:
    "random" (a fresh Philox permutation per sweep, like TFP's sampler) or "index".
  decorrelate_series: batched fits (``fit_causalimpact_many`` / ``_panel``) give series i the
    global chain ids ``i * num_chains + c``, so the Monte-Carlo errors of different series are
    independent and panel aggregates average them out.  False: every series consumes the same
    streams and result i is bit-identical to ``fit_causalimpact(datas[i], sampler="gibbs")``.
  return_level: False leaves ``posterior_samples.level`` (and ``seasonal_levels``) as None: the
    [S, T] level paths are then neither all-gathered across ranks nor copied to the host (they
    are the largest part of both); series and summary are unaffected.
  exchange: what a fit sharded over several GPUs moves between them.  "draws" (default): one
    all-gather of every rank's result rows, after which each rank holds all draws and computes the
    impact itself -- results are bit-identical for any GPU count.  "columns": the predictive
    trajectories stay where they were drawn; the impact stage swaps TIME BLOCKS of the transposed
    paths (all-to-all) and each rank selects the quantiles of T/world steps (shard.impact_sharded).
    Only theta (and the level paths if ``return_level``) are gathered.  Quantiles are identical to
    "draws"; mean-derived columns agree to the rounding of the per-rank partial means.
  local_linear_trend: EXTENSION (the reference's model has no slope, lib.py:496; BASELINE.json
    configs[2] asks for it): the level follows a local linear trend -- state (level, slope),
    slope variance ~ InverseGamma like the level's -- sampled by batched-chain HMC over the d = 2
    scan filter (csrc/ci_llt.cuh) with the d = 2 simulation smoother for the predictive draws
    (csrc/ci_llt_predict.cuh).  ``posterior_samples.slope_scale`` holds the extra draws.
  profile: record the wall time of every phase of the fit in ``diagnostics["phases_ms"]`` (the
    device is synchronised at each phase boundary, so the phases do not overlap).
  """
  num_chains: int = 64
  sampler: str = "auto"
  gibbs_min_warmup: int = 100
  max_leapfrog: int = 8
  min_warmup: int = 300
  init_step: float = 0.05
  target_accept: float = 0.8
  device: Optional[int] = None
  whiten: bool = True
  upper_bound_on: str = "variance"
  ssvs_order: str = "random"
  decorrelate_series: bool = True
  return_level: bool = True
  exchange: str = "draws"
  profile: bool = False
  local_linear_trend: bool = False


_ENGINES = {}


class _Phases:
  """Wall time per phase of a fit (EngineOptions.profile); a no-op when off."""

  def __init__(self, on: bool, eng=None):
    import time
    self.on, self.t, self._eng, self._clock = on, {}, eng, time.perf_counter
    self._last = self._clock()

  def mark(self, name: str):
    if not self.on:
      return
    sync = getattr(self._eng, "synchronize", None)
    if sync is not None:
      sync()
    now = self._clock()
    self.t[name] = self.t.get(name, 0.0) + (now - self._last) * 1e3
    self._last = now


def _engine_for(device: int) -> Engine:
  if device not in _ENGINES:
    _ENGINES[device] = Engine(device)
  return _ENGINES[device]


def _resolve_engine(opts: Optional["EngineOptions"]) -> Engine:
  device = None if opts is None else opts.device
  if device is None:
    device = int(os.environ.get("LOCAL_RANK", "0")) if _shard.world()[1] > 1 else 0
  return _engine_for(device)


def _np_dtype(dtype) -> np.dtype:
  if hasattr(dtype, "as_numpy_dtype"):
    dtype = dtype.as_numpy_dtype
  dt = np.dtype(dtype)
  if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
    raise ValueError(f"dtype must be float32 or float64, got {dt}")
  return dt


def _seed_to_u64(seed) -> int:
  """int -> (0, seed) like the reference (causalimpact_lib.py:535-539); a pair is
  packed into 64 bits; None -> fresh entropy (non-deterministic, as upstream)."""
  if seed is None:
    return int(np.random.SeedSequence().generate_state(2, dtype=np.uint32).astype(np.uint64)
               @ np.array([1 << 32, 1], dtype=np.uint64))
  if isinstance(seed, (int, np.integer)):
    seed = (0, int(seed))
  a, b = (int(v) for v in np.asarray(seed).reshape(-1)[:2])
  return ((a & 0xFFFFFFFF) << 32) | (b & 0xFFFFFFFF)


@dataclasses.dataclass
class _Whitening:
  """w = z @ Linv with L L' = X_obs' X_obs + Omega: the design matrix handed to
  the engine has an identity Gram matrix, so HMC sees an almost isotropic
  posterior in the regression block.  Purely a change of variables (constant
  Jacobian); draws are mapped back before they are returned."""
  Linv: np.ndarray

  @staticmethod
  def build(design, observed, omega) -> "_Whitening":
    gram = design[observed].T @ design[observed] + omega
    L = np.linalg.cholesky(gram)
    return _Whitening(Linv=np.linalg.inv(L))

  def design(self, design):
    return design @ self.Linv.T

  def omega(self, omega):
    return self.Linv @ omega @ self.Linv.T

  def to_weights(self, z):
    return z @ self.Linv


def _train_causalimpact_sts(*, ci_data, prior_level_sd, seed, num_results: int,
                            num_warmup_steps: int, model=None, dtype=np.float32, seasons=(),
                            experimental_tf_function_cache_key_addition: int = 0,
                            engine_options: Optional[EngineOptions] = None):
  """The engine swap point: same contract as the reference's
  ``_train_causalimpact_sts`` (causalimpact_lib.py:503-606) -- returns
  ``(posterior_samples, posterior_means [T], posterior_trajectories [S, T])`` on
  the standardized scale.  The last two are ``DeviceArray``s: they stay in HBM for
  ``ci_impact`` and are copied to the host only if the caller asks (``np.asarray``)."""
  del experimental_tf_function_cache_key_addition      # no tracing cache here
  if model is not None:
    raise NotImplementedError("experimental_model needs TFP objects; not supported by the "
                              "B200 engine")
  opts = engine_options or EngineOptions()
  seasons = list(seasons or ())
  if seasons and opts.sampler == "hmc":
    raise NotImplementedError("seasonal components are sampled by the Gibbs kernel "
                              "(EngineOptions.sampler 'auto' or 'gibbs'), not by HMC")
  np_dt = _np_dtype(dtype)
  seed64 = _seed_to_u64(seed)

  rank, ws = _shard.world()
  eng = _resolve_engine(opts)
  ph = _Phases(opts.profile, eng)
  if seed is None and ws > 1:          # fresh entropy: every rank must use rank 0's
    seed64 = _shard.broadcast_u64(seed64, getattr(eng, "torch_device", lambda: None)())

  y_ext, design, outcome_sd = ci_data.engine_inputs(np_dt)
  if opts.upper_bound_on not in ("variance", "scale"):
    raise ValueError(f"EngineOptions.upper_bound_on must be variance|scale, got {opts.upper_bound_on!r}")
  if opts.local_linear_trend and (seasons or opts.sampler == "gibbs"):
    raise NotImplementedError("local_linear_trend is sampled by HMC and has no seasonal components")
  from ._engine import MODEL_LOCAL_LEVEL, MODEL_LOCAL_LINEAR_TREND
  spec = build_problem(y_ext, design, prior_level_sd=prior_level_sd, outcome_sd=outcome_sd,
                       dtype=np_dt, ub_on_scale=opts.upper_bound_on == "scale",
                       model=MODEL_LOCAL_LINEAR_TREND if opts.local_linear_trend else MODEL_LOCAL_LEVEL)
  p, T = spec.p, spec.T
  if opts.sampler not in ("auto", "hmc", "gibbs"):
    raise ValueError(f"EngineOptions.sampler must be auto|hmc|gibbs, got {opts.sampler!r}")
  use_gibbs = (opts.sampler == "gibbs" or (opts.sampler == "auto" and p > 3) or bool(seasons)) \
      and not opts.local_linear_trend
  wh = None
  if p and opts.whiten and not use_gibbs:
    wh = _Whitening.build(design, ~np.isnan(y_ext), spec.Omega)
    spec = dataclasses.replace(spec, X=wh.design(design), Omega=wh.omega(spec.Omega))
  eng.set_data(spec)
  sched = build_seasonal(seasons, T, outcome_sd)                 # lib.py:471-489
  if sched is not None:
    eng.set_seasonal(sched)
  K = 0 if sched is None else sched.K
  ph.mark("prepare_upload")

  # ---- chains: global ids 0..C-1, contiguous shard per rank ----
  # Everything below stays in the engine's HBM (tensors on eng's device; torch is the
  # allocator and the collective plumbing, every kernel is the engine's own): sampler ->
  # smoother -> predictive mean, and later ci_impact, with no host round trip of the
  # [S,T] arrays.  Only theta and the level paths (part of the result object) come back.
  C = max(int(opts.num_chains), 1)
  n_per = max(1, math.ceil(num_results / C))
  c0, c_local = _shard.split_range(C, ws, rank)
  stats = None
  extra_l = []            # seasonal only: [latent, contributions (flattened), log drift variances]
  if sched is not None:
    # the reference's sampler with seasonal components (lib.py:365-388, 471-489): the joint
    # (level, seasonal) draw replaces the level draw; one drift variance per component
    n_warm = max(int(num_warmup_steps), int(opts.gibbs_min_warmup))
    theta_l, level_l, latent_l, traj_l, seas_l, drift_l, incl = eng.gibbs_seasonal_run_t(
        max(c_local, 1), n_warmup=n_warm, n_results=n_per, seed=seed64, chain_id0=c0,
        sparse=True, ssvs_order=opts.ssvs_order)
    extra_l = [latent_l, seas_l.reshape(seas_l.shape[0], T * K), drift_l]
    stats = {"sampler": "gibbs", "inclusion": incl[:c_local]}
  elif use_gibbs:
    # the reference's sampler (lib.py:365-388): spike-and-slab Gibbs sweeps, started
    # from its initial state (lib.py:566-581); chain-major rows like the HMC path
    n_warm = max(int(num_warmup_steps), int(opts.gibbs_min_warmup))
    theta_l, level_l, traj_l, incl = eng.gibbs_run_t(max(c_local, 1), n_warmup=n_warm,
                                                     n_results=n_per, seed=seed64,
                                                     chain_id0=c0, sparse=True,
                                                     ssvs_order=opts.ssvs_order)
    stats = {"sampler": "gibbs", "inclusion": incl[:c_local]}
  else:
    rng = np.random.Generator(np.random.Philox(key=seed64))
    theta0 = np.tile(initial_theta(spec, prior_level_sd), (C, 1))
    if p:
      seen = ~np.isnan(y_ext)
      z_star = spec.X[seen].T @ y_ext[seen] if wh is not None else \
          np.linalg.solve(spec.X[seen].T @ spec.X[seen] + spec.Omega, spec.X[seen].T @ y_ext[seen])
      theta0[:, :p] = z_star + 0.05 * rng.normal(size=(C, p))
    theta0[:, p:] += 0.1 * rng.normal(size=(C, spec.dim - p))
    # keep every start strictly inside the truncated support (lib.py:432, 442-443):
    # a chain that starts at log-density -inf could never move
    theta0[:, p] = np.minimum(theta0[:, p], np.log(0.8 * spec.ub_variance(spec.obs_ub)))
    theta0[:, p + 1] = np.minimum(theta0[:, p + 1], np.log(0.8 * spec.ub_variance(spec.lvl_ub)))
    if spec.d == 2:
      theta0[:, p + 2] = np.minimum(theta0[:, p + 2], np.log(0.8 * spec.ub_variance(spec.slope_ub)))
    n_warm = max(int(num_warmup_steps), int(opts.min_warmup))
    # a rank without chains (more ranks than chains) still runs one throw-away chain so
    # that every rank holds tensors of the right width for the all-gather
    lo = min(c0, C - 1)
    draws, hstats = eng.hmc_run_t(theta0[lo:lo + max(c_local, 1)], n_warmup=n_warm,
                                  n_results=n_per, seed=seed64, chain_id0=lo,
                                  max_leapfrog=opts.max_leapfrog, init_step=opts.init_step,
                                  target_accept=opts.target_accept)
    # chain-major draw ids: g = chain * n_per + iteration  (contiguous per rank)
    theta_l = draws.permute(1, 0, 2).reshape(-1, spec.dim).contiguous()
    ph.mark("sampler")
    level_l, traj_l = eng.posterior_predict_t(theta_l, seed=seed64 ^ 0x9E3779B97F4A7C15,
                                              draw_id0=lo * n_per)
    hstats = hstats[:c_local]
    stats = {"sampler": "hmc", "accept_rate": np.asarray(hstats["accept_rate"]),
             "step_size": np.asarray(hstats["step_size"]),
             "n_divergent": np.asarray(hstats["n_divergent"]),
             "n_leapfrog": np.asarray(hstats["n_leapfrog"])}
  ph.mark("predictive" if stats.get("sampler") == "hmc" else "sampler")
  n_local = c_local * n_per
  parts = [theta_l, level_l, traj_l] + extra_l
  if opts.exchange not in ("draws", "columns"):
    raise ValueError(f"EngineOptions.exchange must be draws|columns, got {opts.exchange!r}")
  by_columns = ws > 1 and opts.exchange == "columns"
  if ws == 1:
    parts = [t[:num_results] for t in parts]
  elif not by_columns:
    # the ONE collective of the fit: every chain contributes n_per result rows
    import torch
    widths = [t.shape[1] for t in parts]
    rows = torch.cat([t[:n_local] for t in parts], dim=1)
    rows = _shard.all_gather_rows(rows, C, rows_per_item=n_per)[:num_results]
    parts = [t.contiguous() for t in torch.split(rows, widths, dim=1)]
  if not by_columns:
    theta_t, level_t, traj_t = parts[:3]
    ph.mark("all_gather")
    # mean of the predictive mixture = average of level (+ seasonal) + X.w over the draws
    # (causalimpact_lib.py:627); fixed summation order over the gathered draws => the same
    # for any GPU count
    mean_out = DeviceArray(eng.predictive_mean_t(theta_t, parts[3] if sched is not None else level_t))
    ph.mark("predictive_mean")
    traj_out = DeviceArray(traj_t)
  else:
    # EngineOptions.exchange == "columns": the trajectories stay sharded (the impact stage swaps
    # time blocks instead, shard.impact_sharded); only what the result object holds is gathered
    import torch
    counts = []
    for r in range(ws):
      r0, rc = _shard.split_range(C, ws, r)
      counts.append(max(0, min((r0 + rc) * n_per, num_results) - r0 * n_per))
    keep = [0] + ([1] if opts.return_level else [])
    if sched is not None:
      keep += ([4] if opts.return_level else []) + [5]
    widths = [parts[i].shape[1] for i in keep]
    rows = torch.cat([parts[i][:n_local] for i in keep], dim=1)
    rows = _shard.all_gather_rows(rows, C, rows_per_item=n_per)[:num_results]
    got = dict(zip(keep, (t.contiguous() for t in torch.split(rows, widths, dim=1))))
    theta_t, level_t = got[0], got.get(1)
    parts = [got.get(i) for i in range(len(parts))]
    ph.mark("all_gather")
    mean_out = _shard.ShardedMean(
        eng, _shard.predictive_mean_part(eng, theta_l, extra_l[0] if sched is not None else level_l,
                                         counts), counts)
    ph.mark("predictive_mean")
    traj_out = _shard.ShardedDraws(traj_l[:counts[rank]], counts)

  samples = _package_samples(eng, theta_t, level_t if opts.return_level else None, p, T, np_dt, wh,
                             (parts[4] if opts.return_level else None, parts[5], K)
                             if sched is not None else None)
  ph.mark("package_samples")
  if stats is not None:
    # convergence across the parallel chains: split-R-hat of (log sigma_obs^2, log sigma_level^2)
    # over the complete chains among the kept draws (rows are chain-major)
    full = (theta_t.shape[0] // n_per) * n_per
    if full >= 2 * n_per:
      th = eng.to_host(theta_t[:full, p:p + 2]).reshape(full // n_per, n_per, 2)
      stats["rhat_log_variances"] = split_rhat(th)
  samples.hmc_stats = stats            # pylint: disable=attribute-defined-outside-init
  samples.phases = ph                  # pylint: disable=attribute-defined-outside-init
  return samples, mean_out, traj_out


def split_rhat(x: np.ndarray) -> np.ndarray:
  """Split-R-hat (Gelman et al. 2013) of draws [chains, iterations, k] -> [k].  The reference
  returns no convergence diagnostics; with many short chains this is the
  natural one.  NaN when a chain has fewer than 4 iterations."""
  x = np.asarray(x, dtype=np.float64)
  c, n = x.shape[:2]
  if n < 4 or c < 1:
    return np.full(x.shape[2:], np.nan)
  h = n // 2
  halves = np.concatenate([x[:, :h], x[:, n - h:]], axis=0)            # [2c, h, k]
  w = halves.var(axis=1, ddof=1).mean(axis=0)
  b = h * halves.mean(axis=1).var(axis=0, ddof=1)
  with np.errstate(divide="ignore", invalid="ignore"):
    return np.sqrt(((h - 1) / h * w + b / h) / w)


def _package_samples(eng, theta_t, level_t, p, T, np_dt, wh=None, seasonal=None):
  """Device draws -> the reference's posterior-sample record (host arrays)."""
  theta = eng.to_host(theta_t).astype(np.float64)
  level = None if level_t is None else eng.to_host(level_t)
  z = theta[:, :p]
  weights = wh.to_weights(z) if wh is not None else z
  S = theta.shape[0]
  if seasonal is not None:
    # each component's contribution at every step: what the reference extracts as the
    # 0-th element of the component's latent (lib.py:299-317), [S, T, K]
    seas_t, drift_t, K = seasonal
    seas_levels = None if seas_t is None else \
        eng.to_host(seas_t).reshape(S, T, K).astype(np_dt, copy=False)
    drift_scales = np.exp(0.5 * eng.to_host(drift_t).astype(np.float64)).astype(np_dt)
  else:
    seas_levels, drift_scales = np.zeros((S, T, 0), np_dt), np.zeros((S, 0), np_dt)
  return CausalImpactPosteriorSamples(
      observation_noise_scale=Samples(np.exp(0.5 * theta[:, p]).astype(np_dt)),
      level_scale=Samples(np.exp(0.5 * theta[:, p + 1]).astype(np_dt)),
      level=None if level is None else Samples(level.astype(np_dt, copy=False)),
      weights=Samples(weights.astype(np_dt)) if p else Samples(np.zeros((S, 0), np_dt)),
      seasonal_drift_scales=Samples(drift_scales),
      seasonal_levels=None if seas_levels is None else Samples(seas_levels),
      slope_scale=(Samples(np.exp(0.5 * theta[:, p + 2]).astype(np_dt))
                   if theta.shape[1] == p + 3 else None))


def fit_causalimpact(data: pd.DataFrame,
                     pre_period: Tuple[_frame.InputDateType, _frame.InputDateType],
                     post_period: Tuple[_frame.InputDateType, _frame.InputDateType],
                     alpha: float = 0.05,
                     seed=None,
                     data_options: Optional[DataOptions] = None,
                     model_options: Optional[ModelOptions] = None,
                     inference_options: Optional[InferenceOptions] = None,
                     **kwargs) -> CausalImpactAnalysis:
  """Fit a CausalImpact model (same signature as the reference,
  causalimpact_lib.py:223-231).

  Extra keyword (experimental, like the reference's): ``engine_options``.
  Unknown keywords raise TypeError (causalimpact_lib.py:272-273).
  """
  data_options = data_options if data_options is not None else DataOptions()
  model_options = model_options if model_options is not None else ModelOptions()
  inference_options = inference_options if inference_options is not None else InferenceOptions()
  experimental_model = kwargs.pop("experimental_model", None)
  cache_key = kwargs.pop("experimental_tf_function_cache_key_addition", 0)
  engine_options = kwargs.pop("engine_options", None)
  if kwargs:
    raise TypeError(f"Received unknown {kwargs=}")

  np_dt = _np_dtype(data_options.dtype)
  ci_data = _frame.CausalImpactData(
      data=data, pre_period=pre_period, post_period=post_period,
      outcome_column=data_options.outcome_column,
      standardize_data=data_options.standardize_data, dtype=np_dt)
  samples, means, trajectories = _train_causalimpact_sts(
      ci_data=ci_data, prior_level_sd=model_options.prior_level_sd, seed=seed,
      num_results=inference_options.num_results,
      num_warmup_steps=inference_options.num_warmup_steps, model=experimental_model,
      dtype=np_dt, seasons=model_options.seasons,
      experimental_tf_function_cache_key_addition=cache_key, engine_options=engine_options)
  eng = _resolve_engine(engine_options)
  impact_fn = eng.impact
  if isinstance(trajectories, _shard.ShardedDraws):
    def impact_fn(traj, mean, meta):       # draws sharded over the ranks: swap time blocks
      T = traj.shape[1]
      flat = eng.to_host(_shard.impact_sharded(eng, traj.local, mean, meta, traj.counts))
      return flat[:T * 9].reshape(T, 9), flat[T * 9:]
  series, summary = _impact.compute_impact(means, trajectories, ci_data, alpha, impact_fn)
  ph = getattr(samples, "phases", None)
  if ph is not None and ph.on:
    ph.mark("impact_and_frames")
    samples.hmc_stats["phases_ms"] = dict(ph.t)
  return _analysis(series, summary, samples)


def _analysis(series, summary, samples) -> CausalImpactAnalysis:
  stats = getattr(samples, "hmc_stats", None)
  result_samples = CausalImpactPosteriorSamples(
      observation_noise_scale=samples.observation_noise_scale,
      level_scale=samples.level_scale, level=samples.level,
      weights=samples.weights if samples.weights.shape[1] > 0 else None,    # :330-331
      seasonal_drift_scales=(samples.seasonal_drift_scales
                             if samples.seasonal_drift_scales.shape[-1] > 0 else None),   # :332-334
      seasonal_levels=samples.seasonal_levels, slope_scale=samples.slope_scale)
  return CausalImpactAnalysis(series, summary, result_samples, stats)


class PanelAnalysis(CausalImpactAnalysis):
  """One series of a panel fit behind the ``CausalImpactAnalysis`` interface: the ``series`` /
  ``summary`` frames and the sample record are built from the panel's arrays on first access (a
  thousand-series call does not pay for two thousand DataFrames nobody may look at)."""

  def __init__(self, panel_result, i: int, np_dt):   # pylint: disable=super-init-not-called
    self._res, self._i, self._np_dt = panel_result, i, np_dt
    self._frames = None
    self._samples = None
    self.diagnostics = {"sampler": "gibbs", "inclusion": panel_result.inclusion[i]}

  def _build(self):
    if self._frames is None:
      self._frames = self._res.frames(self._i)
    return self._frames

  series = property(lambda self: self._build()[0])
  summary = property(lambda self: self._build()[1])

  @property
  def posterior_samples(self) -> CausalImpactPosteriorSamples:
    if self._samples is None:
      r, i = self._res, self._i
      S = r.observation_noise_scale.shape[1]
      seasonal = r.seasonal_drift_scales is not None
      self._samples = CausalImpactPosteriorSamples(
          observation_noise_scale=Samples(r.observation_noise_scale[i]),
          level_scale=Samples(r.level_scale[i]),
          level=None if r.level is None else Samples(r.level[i]),
          weights=None if r.weights is None else Samples(r.weights[i]),
          seasonal_drift_scales=Samples(r.seasonal_drift_scales[i]) if seasonal else None,
          seasonal_levels=(None if r.level is None else
                           Samples(r.seasonal_levels[i]) if seasonal else
                           Samples(np.zeros((S, r.level.shape[2], 0), self._np_dt))))
    return self._samples


def _stack_frames(datas, outcome_column):
  """[N, T, 1 + k] float64 values (outcome first) and the common index of DataFrames that share
  index and columns; None when they do not (or are not all-numeric frames): the caller then takes
  the per-frame path, which raises the reference's errors."""
  first = datas[0]
  if not isinstance(first, pd.DataFrame) or first.shape[1] < 1:
    return None
  cols = list(first.columns)
  if outcome_column is not None:
    if outcome_column not in cols:
      return None
    cols = [outcome_column] + [c for c in cols if c != outcome_column]
  order = None if cols == list(first.columns) else cols
  vals = np.empty((len(datas),) + first.shape, dtype=np.float64)
  for i, d in enumerate(datas):
    if not isinstance(d, pd.DataFrame) or d.shape != first.shape or \
        not (d.index is first.index or d.index.equals(first.index)) or \
        not (d.columns is first.columns or d.columns.equals(first.columns)):
      return None
    try:
      vals[i] = (d if order is None else d[order]).to_numpy(dtype=np.float64, copy=False)
    except (TypeError, ValueError):
      return None
  return vals, first.index


def fit_causalimpact_many(datas, pre_period, post_period, alpha: float = 0.05, seed=None,
                          data_options: Optional[DataOptions] = None,
                          model_options: Optional[ModelOptions] = None,
                          inference_options: Optional[InferenceOptions] = None,
                          engine_options: Optional[EngineOptions] = None):
  """Fit MANY independent series in one go (the "thousands of
  geographies" use case the reference serves with a Python loop of single-chain fits).

  ``datas``: sequence of DataFrames with a common shape -- same index, periods and number of
  covariates.  Every series is prepared exactly as ``fit_causalimpact`` does (data.py:77-137),
  all of them are uploaded together (``ci_set_data_batch``) and ONE launch of the reference's
  sampler (``ci_gibbs_run_batch_d``: spike-and-slab Gibbs, grid.y = series) draws every chain
  of every series; predictive mean and impact follow per series on the device.  The batched
  path always runs the Gibbs kernel (``fit_causalimpact(sampler="auto")`` picks HMC for <= 2
  covariates).  With ``EngineOptions(decorrelate_series=False)`` result i is bit-identical to
  ``fit_causalimpact(datas[i], ..., engine_options=EngineOptions(sampler="gibbs"))`` with the
  same seed; by default every series draws from its own Philox streams.

  Frames that share index and columns (the usual panel of geographies) and the default
  ``decorrelate_series=True`` take the PANEL route: the values are stacked and handed to
  ``fit_causalimpact_panel`` -- data prep on the device, no per-series pandas -- and every result is
  a ``PanelAnalysis`` whose frames are built on first access.  (``decorrelate_series=False`` keeps
  the per-frame preparation: that is what makes result i bit-identical to the single fit.)

  Multi-GPU: series are sharded over the ranks of an initialised process group (contiguous
  ranges, no collective: series are independent); a rank returns ``None`` for the series it
  does not own.  ``ModelOptions.seasons`` are supported (one season calendar for the panel).
  """
  data_options = data_options if data_options is not None else DataOptions()
  model_options = model_options if model_options is not None else ModelOptions()
  inference_options = inference_options if inference_options is not None else InferenceOptions()
  opts = engine_options or EngineOptions()
  if opts.sampler == "hmc":
    raise NotImplementedError("the batched path runs the Gibbs kernel")
  seasons = list(model_options.seasons or ())
  np_dt = _np_dtype(data_options.dtype)
  seed64 = _seed_to_u64(seed)
  datas = list(datas)
  if opts.decorrelate_series and datas:
    stacked = _stack_frames(datas, data_options.outcome_column)
    if stacked is not None:
      from . import panel as _panel
      res = _panel.fit_causalimpact_panel(
          stacked[0], stacked[1], pre_period, post_period, alpha=alpha, seed=seed,
          data_options=data_options, model_options=model_options,
          inference_options=inference_options, engine_options=opts, keep_level=opts.return_level)
      out = [None] * len(datas)
      for j, sid in enumerate(res.series_ids):
        out[int(sid)] = PanelAnalysis(res, j, np_dt)
      return out
  rank, ws = _shard.world()
  s0, n_local = _shard.split_range(len(datas), ws, rank)
  out = [None] * len(datas)
  eng = _resolve_engine(opts)
  if seed is None and ws > 1:
    seed64 = _shard.broadcast_u64(seed64, getattr(eng, "torch_device", lambda: None)())
  if n_local == 0:
    return out
  cids, specs, scheds = [], [], []
  for d in datas[s0:s0 + n_local]:
    cid = _frame.CausalImpactData(data=d, pre_period=pre_period, post_period=post_period,
                                  outcome_column=data_options.outcome_column,
                                  standardize_data=data_options.standardize_data, dtype=np_dt)
    y_ext, design, outcome_sd = cid.engine_inputs(np_dt)
    specs.append(build_problem(y_ext, design, prior_level_sd=model_options.prior_level_sd,
                               outcome_sd=outcome_sd, dtype=np_dt,
                               ub_on_scale=opts.upper_bound_on == "scale"))
    scheds.append(build_seasonal(seasons, specs[-1].T, outcome_sd))
    cids.append(cid)
  p, T = specs[0].p, specs[0].T
  eng.set_data_batch(specs)
  K = 0
  if seasons:
    eng.set_seasonal_batch(scheds)
    K = scheds[0].K
  num_results = inference_options.num_results
  C = max(int(opts.num_chains), 1)
  n_per = max(1, math.ceil(num_results / C))
  n_warm = max(int(inference_options.num_warmup_steps), int(opts.gibbs_min_warmup))
  # series i of the WHOLE input (not of this rank's shard) owns the chain ids i * C .. i * C + C - 1
  stride = C if opts.decorrelate_series else 0
  bkw = dict(n_warmup=n_warm, n_results=n_per, seed=seed64, chain_id0=s0 * stride, sparse=True,
             ssvs_order=opts.ssvs_order, series_stride=stride)
  if seasons:
    theta, level, latent, traj, seas, drift, incl = eng.gibbs_seasonal_run_batch_t(C, **bkw)
  else:
    theta, level, traj, incl = eng.gibbs_run_batch_t(C, **bkw)
    latent = level
  # predictive mean and impact of EVERY series: one batched launch each (grid.y = series), one
  # read-back; the frames are O(T) host packaging per series
  S = min(num_results, C * n_per)
  metas = [_impact.prepare(cid, alpha) for cid in cids]
  mean_all = eng.predictive_mean_batch_t(theta[:, :S], latent[:, :S])
  ser_d, sum_d = eng.impact_batch_t(
      traj[:, :S], mean_all, scale=[m.scale for m in metas], offset=[m.offset for m in metas],
      obs_sum=[m.obs_sum for m in metas], observed=np.stack([m.observed for m in metas]),
      period=metas[0].period, q_lo=metas[0].q_lo, q_hi=metas[0].q_hi)
  series9 = eng.to_host(ser_d).reshape(len(cids), T, 9)
  summ = eng.to_host(sum_d)
  for i, cid in enumerate(cids):
    if not np.array_equal(metas[i].period, metas[0].period):
      raise ValueError("the series of a batch must share their index and periods")
    samples = _package_samples(
        eng, theta[i, :S], level[i, :S], p, T, np_dt,
        seasonal=(seas[i, :S], drift[i, :S], K) if seasons else None)
    samples.hmc_stats = {"sampler": "gibbs", "inclusion": incl[i]}
    series, summary = _impact.package(series9[i], summ[i], S, metas[i], cid, alpha)
    out[s0 + i] = _analysis(series, summary, samples)
  return out
