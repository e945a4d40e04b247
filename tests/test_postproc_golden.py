"""Host logic + the impact oracle vs golden vectors produced by the REFERENCE's own
code (oracle/make_golden_postproc.py ran causalimpact_lib._compute_impact and
data.CausalImpactData unmodified).  CPU: the device call ci_impact is replaced by its
numpy oracle (oracle/impact_np.py) -- this pins the oracle; the GPU variant of this
test (the real kernels against the same goldens) lives in test_gpu_impact.py."""
import glob
import os

import numpy as np
import pandas as pd
import pytest

from causalimpact_b200 import frame as fr
from causalimpact_b200 import impact
from oracle import impact_np

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "postproc_*.npz")))


def oracle_impact(traj, mean, meta):
  return impact_np.impact_arrays(traj, mean, meta.observed, meta.period, meta.scale, meta.offset,
                                 meta.q_lo, meta.q_hi, meta.obs_sum)


def load_case(path):
  g = np.load(path, allow_pickle=False)
  idx = pd.DatetimeIndex(g["frame_index"]) if bool(g["index_is_datetime"]) \
      else pd.Index(g["frame_index"])
  data = pd.DataFrame(g["frame_values"], index=idx, columns=[str(c) for c in g["frame_columns"]])
  pre = (idx[int(g["pre"][0])], idx[int(g["pre"][1])])
  post = (idx[int(g["post"][0])], idx[int(g["post"][1])])
  return g, data, pre, post


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[9:-4] for p in GOLDEN])
def test_data_prep_matches_reference(path):
  g, data, pre, post = load_case(path)
  ci = fr.CausalImpactData(data, pre, post, standardize_data=bool(g["standardize"]))
  np.testing.assert_allclose(ci.model_pre_data.values, g["model_pre"], rtol=1e-13, equal_nan=True)
  np.testing.assert_allclose(ci.model_after_pre_data.values, g["model_after"], rtol=1e-13,
                             equal_nan=True)
  if g["feature_ts"].size:
    np.testing.assert_allclose(ci.feature_ts.values, g["feature_ts"], rtol=1e-13)
    assert list(ci.feature_ts.columns)[-1] == "intercept_"
  else:
    assert ci.feature_ts is None


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[9:-4] for p in GOLDEN])
def test_compute_impact_matches_reference(path):
  g, data, pre, post = load_case(path)
  ci = fr.CausalImpactData(data, pre, post, standardize_data=bool(g["standardize"]))
  series, summary = impact.compute_impact(g["posterior_means"], g["posterior_trajectories"], ci,
                                          float(g["alpha"]), oracle_impact)
  cols = [str(c) for c in g["series_columns"]]
  assert list(series.columns[:len(cols)]) == cols
  # standardize_data=True: the reference un-scales into float64 first, so we agree
  # to rounding.  standardize_data=False: the reference stays in float32 pandas
  # arithmetic (posterior_processing.py:88-91 skips the float64 scaler); we
  # compute in float64, so agreement is to float32 rounding of the values.
  rtol, atol = (1e-11, 1e-11) if bool(g["standardize"]) else \
      (1e-5, 4e-6 * float(np.nanmax(np.abs(g["series_values"]))))
  np.testing.assert_allclose(series[cols].values.astype(float), g["series_values"], rtol=rtol,
                             atol=atol, equal_nan=True)
  assert [str(c) for c in g["summary_columns"]] == list(summary.columns)
  assert list(summary.index) == ["average", "cumulative"]
  np.testing.assert_allclose(summary.values.astype(float), g["summary_values"], rtol=rtol,
                             atol=atol, equal_nan=True)
  assert series.index.equals(data.index)
  for c in ("pre_period_start", "pre_period_end", "post_period_start", "post_period_end"):
    assert c in series.columns


def test_alpha_is_validated():
  g, data, pre, post = load_case(GOLDEN[0])
  ci = fr.CausalImpactData(data, pre, post)
  with pytest.raises(ValueError, match="alpha"):
    impact.compute_impact(g["posterior_means"], g["posterior_trajectories"], ci, 1.5,
                          oracle_impact)


def test_text_summary_layout():
  """causalimpact.summary(impact) (reference summary.py:133-178): same layout as
  the reference's golden text (testdata/test_summary_output.txt), own formatter."""
  import types
  from causalimpact_b200 import report
  cols = ["actual", "predicted", "predicted_lower", "predicted_upper", "predicted_sd",
          "abs_effect", "abs_effect_lower", "abs_effect_upper", "abs_effect_sd", "rel_effect",
          "rel_effect_lower", "rel_effect_upper", "rel_effect_sd", "p_value", "alpha"]
  tab = pd.DataFrame([[5.3, 4.3, 3.3, 6.3, 0.0, 3.3, 2.3, 6.3, 0.0, .123, .143, .343, .001, .4593, .1],
                      [10.3, 9.3, 8.3, 9.3, 0.1, 10.3, 4.3, 9.3, 0.1, .233, .133, .333, .1, .4593, .1]],
                     index=["average", "cumulative"], columns=cols)
  text = report.summary(types.SimpleNamespace(summary=tab))
  lines = text.splitlines()
  assert lines[0] == "Posterior Inference {CausalImpact}"
  assert lines[2] == "Actual                    5.3                10.3"
  assert lines[3] == "Prediction (s.d.)         4.3 (0.0)          9.3 (0.1)"
  assert lines[4] == "90% CI                    [3.3, 6.3]         [8.3, 9.3]"
  assert lines[9] == "Relative effect (s.d.)    12.3% (0.1%)       23.3% (10.0%)"
  assert "Posterior tail-area probability p: 0.459" in text
  assert "Posterior probability of an effect: 54.07%" in text
  assert "not statistically significant" not in report.summary(
      types.SimpleNamespace(summary=tab), output_format="report")
  with pytest.raises(ValueError):
    report.summary(types.SimpleNamespace(summary=tab), output_format="x")
