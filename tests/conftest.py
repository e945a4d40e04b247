import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200")):
  if p not in sys.path:
    sys.path.insert(0, p)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def make_series(T, n_cov, seed, pre_frac=0.7, nan_frac=0.01, effect=10.0):
  """Synthetic CausalImpact input following BASELINE.md section 4.

  Covariates x_j = 100 + AR(1)(phi=0.999); beta = (1.2, 0.6, -0.4, 0, ...);
  y = sum beta_j x_j + N(0,1); pre = first 70 %; y[post] += effect; 1 % of the
  pre-period y set to NaN; standardised with pre-period mean / std(ddof=1);
  intercept column appended (reference data.py:114-135).  Returns
  (y_ext [T] with NaN for masked steps, design [T, n_cov+1] or None, y_std_full).
  """
  rng = np.random.Generator(np.random.PCG64(seed))
  t_pre = int(round(pre_frac * T))
  beta = np.zeros(max(n_cov, 1)); beta[:3] = (1.2, 0.6, -0.4)[:min(3, max(n_cov, 1))]
  xs = np.empty((T, n_cov))
  for j in range(n_cov):
    a = np.empty(T); a[0] = rng.normal()
    eps = rng.normal(size=T)
    for t in range(1, T):
      a[t] = 0.999 * a[t - 1] + eps[t]
    xs[:, j] = 100.0 + a
  if n_cov:
    y = xs @ beta[:n_cov] + rng.normal(size=T)
  else:
    y = 100.0 + np.cumsum(0.05 * rng.normal(size=T)) + rng.normal(size=T)
  y[t_pre:] += effect
  n_nan = int(nan_frac * t_pre)
  if n_nan:
    y[rng.choice(np.arange(1, t_pre), size=n_nan, replace=False)] = np.nan
  mu, sd = np.nanmean(y[:t_pre]), np.nanstd(y[:t_pre], ddof=1)
  y_std = (y - mu) / sd
  y_ext = y_std.copy(); y_ext[t_pre:] = np.nan
  design = None
  if n_cov:
    xm, xsd = xs[:t_pre].mean(0), xs[:t_pre].std(0, ddof=1)
    design = np.concatenate([(xs - xm) / xsd, np.ones((T, 1))], axis=1)
  return y_ext, design, y_std


def make_thetas(spec_dim, p, C, seed, d=1):
  """theta batch of BASELINE.md section 4: w ~ N(beta_std, 0.1^2), log sigma's."""
  rng = np.random.Generator(np.random.PCG64(seed))
  th = np.zeros((C, spec_dim))
  if p:
    th[:, :p] = 0.1 * rng.normal(size=(C, p))
    th[:, :min(3, p)] += np.array([0.6, 0.3, -0.2])[:min(3, p)]
  th[:, p] = 2.0 * (np.log(0.45) + 0.2 * rng.normal(size=C))
  th[:, p + 1] = 2.0 * (np.log(0.01) + 0.5 * rng.normal(size=C))
  if d == 2:
    th[:, p + 2] = 2.0 * (np.log(0.001) + 0.5 * rng.normal(size=C))
  return th


@pytest.fixture(scope="session")
def engine():
  import causalimpact_b200 as cib
  from causalimpact_b200 import _build
  _build.build()          # no-op when lib/libci_b200.so is newer than csrc/ (nvcc is in the image)
  eng = cib.Engine(0)
  yield eng
  eng.close()
