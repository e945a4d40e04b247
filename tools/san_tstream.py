"""Developer tool: one small pass of the long-series team kernels for compute-sanitizer."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import causalimpact_b200 as cib
from conftest import make_series, make_thetas
for T, n_cov, W, dt in ((2600, 1, 4, np.float32), (6000, 20, 4, np.float32), (2600, 2, 3, np.float64)):
  os.environ["CI_B200_TSW"] = str(W)
  eng = cib.Engine(0)
  y, X, _ = make_series(T, n_cov, 1)
  spec = cib.build_problem(y, X, dtype=dt)
  eng.set_data(spec)
  th = make_thetas(spec.dim, spec.p, 5, 1)
  v, g = eng.logprob_grad(th, with_prior=True)
  d, st = eng.hmc_run(th, n_warmup=3, n_results=2, seed=1, init_step=0.01)
  print(T, n_cov, W, float(v.sum()), float(np.abs(g).sum()), float(d.sum()))
  eng.close()
