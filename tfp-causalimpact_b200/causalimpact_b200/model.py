"""Model + priors as a plain parameter struct (no TFP objects).

Host-side mirror of ``_build_default_gibbs_model`` and the initial sampler
state of ``_train_causalimpact_sts`` (reference:
causalimpact/causalimpact_lib.py:398-500 and :563-581).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence

import numpy as np

from ._engine import F32, F64, MODEL_LOCAL_LEVEL, MODEL_LOCAL_LINEAR_TREND, ProblemSpec


def slab_precision(design: np.ndarray) -> np.ndarray:
  """Spike-and-slab *slab* precision, built from the FULL-length design matrix.

  ``0.01 * set_diag(0.5 * X'X, diag(X'X)) / T``  (causalimpact_lib.py:451-459).
  """
  gram = design.T @ design
  prec = 0.5 * gram
  np.fill_diagonal(prec, np.diag(gram))
  return 0.01 * prec / design.shape[0]


def build_problem(y_ext, design: Optional[np.ndarray], *, prior_level_sd: float = 0.01,
                  outcome_sd: Optional[float] = None, model: int = MODEL_LOCAL_LEVEL,
                  dtype=np.float32, ub_on_scale: bool = False) -> ProblemSpec:
  """Assemble the ``ci_problem`` the reference would hand to its sampler.

  Args:
    y_ext: standardized outcome over pre + after-pre; NaN for every masked
      step, i.e. pre-period gaps and the whole after-pre range
      (causalimpact_lib.py:548-562).
    design: [T, p] standardized covariates + intercept column (data.py:129-135)
      or None.
    prior_level_sd: ModelOptions.prior_level_sd (causalimpact_lib.py:202).
    outcome_sd: nanstd(pre y, ddof=1) (causalimpact_lib.py:563-564); computed
      from ``y_ext`` when omitted.
    ub_on_scale: what the reference's ``prior.upper_bound`` attributes (:432, 442-443,
      474) limit.  They sit on InverseGamma priors over VARIANCES and TFP's samplers clip
      the variance draw (``min(variance, upper_bound)``), so the default is False: e.g.
      sigma_obs^2 <= 1.2 sd.  True bounds the scale instead (sigma_obs <= 1.2 sd).
  """
  y_ext = np.asarray(y_ext, dtype=np.float64)
  seen = y_ext[~np.isnan(y_ext)]
  if seen.size < 2:
    raise ValueError("need at least 2 observed points")
  sd = float(np.std(seen, ddof=1)) if outcome_sd is None else float(outcome_sd)
  with_x = design is not None and design.shape[1] > 0
  if with_x:
    design = np.asarray(design, dtype=np.float64)
    if design.shape[0] != y_ext.shape[0]:
      raise ValueError("design and y_ext lengths differ")
  level0 = prior_level_sd * sd                                   # :572
  # :467-469 takes y[0]; when that is missing (untested upstream) use the first
  # observed value instead of propagating a NaN prior mean.
  m0 = float(seen[0]) if np.isnan(y_ext[0]) else float(y_ext[0])
  np_dt = np.dtype(dtype)
  return ProblemSpec(
      model=model, dtype=F64 if np_dt == np.float64 else F32,
      y=y_ext, X=design if with_x else None,
      Omega=slab_precision(design) if with_x else None,
      m0=m0, P0=sd * sd,
      obs_conc=25.0 if with_x else 0.005,                        # :434-441
      obs_scale=(5.0 if with_x else 0.005) * sd * sd,
      obs_ub=1.2 * sd,                                           # :442-443
      lvl_conc=16.0, lvl_scale=16.0 * level0 * level0,           # :424-431
      lvl_ub=sd,                                                 # :432
      slope_conc=16.0, slope_scale=16.0 * level0 * level0, slope_ub=sd,
      m0_slope=0.0, P0_slope=sd * sd, ub_on_scale=bool(ub_on_scale))


def initial_theta(spec: ProblemSpec, prior_level_sd: float = 0.01) -> np.ndarray:
  """The reference's initial sampler state (causalimpact_lib.py:566-581)."""
  sd = float(np.sqrt(spec.P0))
  th = np.zeros(spec.dim)
  sig_obs = np.sqrt(1.0 - 0.8) * sd if spec.p > 0 else sd       # r2 = 0.8, :566-571
  th[spec.p] = np.log(sig_obs ** 2)
  th[spec.p + 1] = np.log((prior_level_sd * sd) ** 2)
  if spec.model == MODEL_LOCAL_LINEAR_TREND:
    th[spec.p + 2] = np.log((prior_level_sd * sd) ** 2)
  return th


# ---------------------------------------------------------------------------
# seasonal components (ModelOptions.seasons; causalimpact_lib.py:162-180, 471-489)
# ---------------------------------------------------------------------------
MAX_SEASONAL_COMPONENTS = 7     # csrc/ci_device.cuh: MAX_SEAS
MAX_SEASONAL_STATE = 192        # csrc/ci_device.cuh: SEAS_MAXD (level + every seasonal effect)


@dataclasses.dataclass
class SeasonalSchedule:
  """What ``ci_set_seasonal`` needs: per component the active season of every step and
  whether that season is over after the step, plus the priors of lib.py:471-489."""
  num_seasons: List[int]
  active: np.ndarray          # [K, T] uint8
  ends: np.ndarray            # [K, T] uint8
  init_sd: float              # initial_effect_prior = Normal(0, outcome_sd)        (:489)
  drift_conc: float = 0.005   # drift variance ~ InverseGamma(0.005, 5e-7 sd^2)     (:472-473)
  drift_scale: float = 5e-7
  drift_ub: float = 1.0       # .upper_bound = outcome_sd                           (:474)

  @property
  def K(self) -> int:
    return len(self.num_seasons)


def _season_steps(num_seasons: int, num_steps_per_season) -> np.ndarray:
  """int | [num_seasons] | [num_cycles, num_seasons]  ->  flat run lengths, cycle-major."""
  steps = np.asarray(num_steps_per_season)
  if steps.dtype.kind not in "iu":
    raise ValueError("num_steps_per_season must be integer valued")
  if steps.ndim == 0:
    steps = np.full((1, num_seasons), int(steps))
  elif steps.ndim == 1:
    steps = steps.reshape(1, -1)
  if steps.ndim != 2 or steps.shape[1] != num_seasons or np.any(steps < 1):
    raise ValueError("num_steps_per_season must be a positive int, a [num_seasons] tuple or a "
                     f"[num_cycles, num_seasons] tuple of tuples; got shape {steps.shape} for "
                     f"num_seasons={num_seasons}")
  return steps.reshape(-1).astype(np.int64)


def build_seasonal(seasons: Sequence, T: int, outcome_sd: float) -> Optional[SeasonalSchedule]:
  """Season calendar of every ``Seasons`` option over the T modelled steps: step 0 is the
  first step of season 0 (of cycle 0); run lengths repeat cyclically (tfp.sts.Seasonal
  semantics, which the reference instantiates at lib.py:475-489)."""
  if not seasons:
    return None
  K = len(seasons)
  # the engine's limits (csrc/ci_device.cuh: MAX_SEAS, SEAS_MAXD), checked here with a clear
  # message before anything is packed into uint8 calendars
  if K > MAX_SEASONAL_COMPONENTS:
    raise ValueError(f"at most {MAX_SEASONAL_COMPONENTS} seasonal components are supported, got {K}")
  total = 1 + sum(int(s.num_seasons) for s in seasons)
  if total > MAX_SEASONAL_STATE:
    raise ValueError(
        f"1 + sum(num_seasons) = {total} exceeds the seasonal state dimension the engine "
        f"supports ({MAX_SEASONAL_STATE})")
  active = np.zeros((K, T), np.uint8)
  ends = np.zeros((K, T), np.uint8)
  ns = []
  for k, s in enumerate(seasons):
    n = int(s.num_seasons)
    if n < 2:
      raise ValueError("Seasons.num_seasons must be at least 2")
    runs = _season_steps(n, s.num_steps_per_season)
    # expand the run lengths until they cover T steps
    reps = int(np.ceil(T / runs.sum())) + 1
    lengths = np.tile(runs, reps)
    season_of_run = np.tile(np.arange(runs.size) % n, reps)
    active[k] = np.repeat(season_of_run, lengths)[:T]
    last = np.cumsum(lengths) - 1
    ends[k, last[last < T]] = 1
    ns.append(n)
  return SeasonalSchedule(num_seasons=ns, active=active, ends=ends, init_sd=float(outcome_sd),
                          drift_conc=0.005, drift_scale=5e-7 * outcome_sd ** 2,
                          drift_ub=float(outcome_sd))
