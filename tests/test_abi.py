"""CPU-side checks of the drop-in boundary: the library builds, loads and
exports every symbol include/ci_b200.h declares; host logic agrees with the
oracle's restatement.  No compute calls (no GPU here)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, make_series


def test_library_builds_and_exports_every_declared_symbol():
  import __graft_entry__ as g
  g.build()
  from causalimpact_b200 import _engine
  lib = _engine.load_library()
  header = open(os.path.join(ROOT, "include", "ci_b200.h")).read()
  declared = set(re.findall(r"\b(ci_[a-z_0-9]+)\s*\(", header))
  declared -= {"ci_ctx"}
  assert declared == set(_engine.EXPORTS), declared ^ set(_engine.EXPORTS)
  for name in declared:
    assert getattr(lib, name) is not None
  assert lib.ci_version() == 100


def test_struct_layouts_match_header():
  import ctypes
  from causalimpact_b200 import _engine
  assert ctypes.sizeof(_engine.CiProblem) == 4 * 4 + 13 * 8 + 2 * 4
  assert ctypes.sizeof(_engine.CiHmcOpts) == 4 * 4 + 2 * 8
  assert ctypes.sizeof(_engine.CiHmcStats) == 16 == _engine.HMC_STATS_DTYPE.itemsize
  assert ctypes.sizeof(_engine.CiGibbsOpts) == 4 * 4 + 8 + 2 * 4 + 8
  assert ctypes.sizeof(_engine.CiImpactArgs) == 4 * 4 + 5 * 8
  assert ctypes.sizeof(_engine.CiSeasonal) == 8 * 4 + 2 * 8 + 4 * 8
  assert _engine.MAX_SEASONAL == 7 and _engine.IMPACT_SERIES_COLS == 9
  header = open(os.path.join(ROOT, "include", "ci_b200.h")).read()
  assert "#define CI_MAX_SEASONAL 7" in header and "#define CI_IMPACT_SUMMARY_LEN 20" in header


def test_no_cpu_fallback_without_gpu():
  import torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  import causalimpact_b200 as cib
  with pytest.raises(cib.EngineError, match="no CUDA device|no CPU fallback"):
    cib.Engine(0)


def test_host_problem_builder_matches_oracle_restatement():
  import causalimpact_b200 as cib
  from oracle import kalman_np as K
  for n_cov in (0, 3):
    y, X, _ = make_series(150, n_cov, 11)
    spec = cib.build_problem(y, X, prior_level_sd=0.05)
    prob = K.default_problem(y, X, prior_level_sd=0.05)
    for f in ("m0", "P0", "obs_conc", "obs_scale", "obs_ub", "lvl_conc", "lvl_scale", "lvl_ub"):
      assert np.isclose(getattr(spec, f), getattr(prob, f)), f
    if n_cov:
      np.testing.assert_allclose(spec.Omega, prob.Omega)
    np.testing.assert_allclose(cib.initial_theta(spec, 0.05), K.initial_theta(prob, 0.05))
