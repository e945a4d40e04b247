"""Generates tests/golden/postproc_*.npz by running the REFERENCE's own code.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference
exists):   python oracle/make_golden_postproc.py

The reference's post-processing (causalimpact_lib.py:635-1093,
posterior_processing.py:25-98, data.py, indices.py, standardize.py) is pure
pandas/numpy; only its imports need TensorFlow / TFP.  Those two modules are
replaced by inert stubs (just enough for data.py:125-128 to build its masked
series), then ``causalimpact.causalimpact_lib._compute_impact`` and
``causalimpact.data.CausalImpactData`` -- UNMODIFIED reference code -- are
executed on seeded inputs and their outputs stored as golden vectors.
"""
import os
import sys
import types

import numpy as np
import pandas as pd

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _install_stubs():
  class _Any(types.ModuleType):
    def __getattr__(self, name):
      if name.startswith("__"):
        raise AttributeError(name)
      sub = _Any(self.__name__ + "." + name)
      setattr(self, name, sub)
      return sub

    def __call__(self, *a, **k):
      return None

  tf = _Any("tensorflow")
  tf.float32, tf.float64 = np.float32, np.float64
  tf.convert_to_tensor = lambda x, dtype=None: np.asarray(x, dtype=dtype)
  tf.math.is_nan = np.isnan
  tf.function = lambda *a, **k: (lambda f: f)
  tfp = _Any("tensorflow_probability")

  class MaskedTimeSeries:   # data.py:127
    def __init__(self, time_series, is_missing):
      self.time_series, self.is_missing = time_series, is_missing
  tfp.sts.MaskedTimeSeries = MaskedTimeSeries
  names = ["tensorflow", "tensorflow_probability", "tensorflow_probability.python",
           "tensorflow_probability.python.experimental",
           "tensorflow_probability.python.experimental.distributions",
           "tensorflow_probability.python.experimental.sts_gibbs",
           "tensorflow_probability.python.experimental.sts_gibbs.gibbs_sampler",
           "tensorflow_probability.python.internal",
           "tensorflow_probability.python.internal.prefer_static", "altair", "matplotlib",
           "matplotlib.pyplot"]
  for n in names:
    root = tf if n == "tensorflow" else (tfp if n.startswith("tensorflow_probability") else _Any(n))
    mod = root
    if n.startswith("tensorflow_probability."):
      for part in n.split(".")[1:]:
        mod = getattr(mod, part)
    sys.modules[n] = mod
  sys.modules["tensorflow_probability.python.experimental.distributions"] \
      .MultivariateNormalPrecisionFactorLinearOperator = object


def _scenarios():
  """name -> (frame, pre_period, post_period, standardize, alpha, S)."""
  rng = np.random.Generator(np.random.PCG64(20240))
  out = {}
  # (a) the reference's own CSV fixture, NaNs in the pre-period (lib_test.py:215)
  df = pd.read_csv(os.path.join(REF, "causalimpact", "testdata", "data.csv"))
  df = df.set_index(pd.to_datetime(df["t"])).drop(columns=["t"])
  df.loc[df.index[[1, 3, 7]], "y"] = np.nan
  out["csv"] = (df, (df.index[0], df.index[59]), (df.index[60], df.index[-1]), True, 0.05, 60)
  # (b) gap between pre and post, tail after post, NaN inside the post-period,
  #     lead-in rows before the pre-period, datetime index
  n = 120
  idx = pd.date_range("2020-01-01", periods=n, freq="D")
  x = 100 + np.cumsum(rng.normal(size=n))
  y = 1.2 * x + rng.normal(size=n)
  y[80:] += 5
  fr = pd.DataFrame({"y": y, "x": x}, index=idx)
  fr.iloc[90, 0] = np.nan
  fr.iloc[12, 0] = np.nan
  out["gap"] = (fr, (idx[5], idx[69]), (idx[80], idx[109]), True, 0.1, 37)
  # (c) integer index, no standardisation, no covariates
  n = 60
  y = 10 + np.cumsum(0.1 * rng.normal(size=n)) + rng.normal(size=n)
  y[40:] += 2
  out["int"] = (pd.DataFrame({"y": y}), (0, 39), (40, 59), False, 0.05, 11)
  return out, rng


def main():
  _install_stubs()
  sys.path.insert(0, REF)
  import causalimpact.causalimpact_lib as lib   # noqa: E402  (reference, unmodified)
  import causalimpact.data as cid               # noqa: E402
  os.makedirs(OUT, exist_ok=True)
  scen, rng = _scenarios()
  for name, (frame, pre, post, std, alpha, S) in scen.items():
    ci_data = cid.CausalImpactData(frame, pre, post, standardize_data=std, dtype=np.float32)
    T = len(ci_data.model_pre_data) + len(ci_data.model_after_pre_data)
    base = np.concatenate([np.asarray(ci_data.model_pre_data.iloc[:, 0]),
                           np.asarray(ci_data.model_after_pre_data.iloc[:, 0])])
    base = np.where(np.isnan(base), 0.0, base)
    means = (base + 0.05 * rng.normal(size=T)).astype(np.float32)
    traj = (means[None, :] + 0.3 * rng.normal(size=(S, T))).astype(np.float32)
    series, summary = lib._compute_impact(means, traj, ci_data, alpha)   # pylint: disable=protected-access
    val_cols = [c for c in series.columns if not c.endswith(("_start", "_end"))]
    is_dt = isinstance(frame.index, pd.DatetimeIndex)
    np.savez_compressed(
        os.path.join(OUT, f"postproc_{name}.npz"),
        frame_values=frame.values.astype(np.float64), frame_columns=np.array(list(frame.columns)),
        frame_index=(frame.index.as_unit("ns").asi8 if is_dt
                     else np.asarray(frame.index, dtype=np.int64)),
        index_is_datetime=is_dt,
        pre=np.array([frame.index.get_loc(ci_data.pre_period[0]),
                      frame.index.get_loc(ci_data.pre_period[1])]),
        post=np.array([frame.index.get_loc(ci_data.post_period[0]),
                       frame.index.get_loc(ci_data.post_period[1])]),
        standardize=std, alpha=alpha, posterior_means=means, posterior_trajectories=traj,
        series_columns=np.array(val_cols), series_values=series[val_cols].values.astype(np.float64),
        summary_columns=np.array(list(summary.columns)),
        summary_values=summary.values.astype(np.float64),
        # data-prep goldens (data.py:114-135)
        model_pre=ci_data.model_pre_data.values.astype(np.float64),
        model_after=ci_data.model_after_pre_data.values.astype(np.float64),
        feature_ts=(np.zeros((0, 0)) if ci_data.feature_ts is None
                    else ci_data.feature_ts.values.astype(np.float64)))
    print(name, "series", series.shape, "summary", summary.shape)


if __name__ == "__main__":
  main()
