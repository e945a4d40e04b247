"""Model + priors as a plain parameter struct (no TFP objects).

Host-side mirror of ``_build_default_gibbs_model`` and the initial sampler
state of ``_train_causalimpact_sts`` (reference:
causalimpact/causalimpact_lib.py:398-500 and :563-581).
"""
from __future__ import annotations

from typing import Optional

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
                  dtype=np.float32) -> ProblemSpec:
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
      m0_slope=0.0, P0_slope=sd * sd)


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
