"""Exercise every kernel once at small sizes (for compute-sanitizer runs).
Developer tool, not a test:  compute-sanitizer python tools/debug_team.py T n_cov C"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import causalimpact_b200 as cib
from conftest import make_series, make_thetas
from oracle import kalman_np as K

T, n_cov, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
y, X, _ = make_series(T, n_cov, 100 + T, nan_frac=0.02)
prob = K.default_problem(y, X)
for dt in (np.float32, np.float64):
  spec = cib.build_problem(y, X, dtype=dt)
  eng = cib.Engine(0)
  eng.set_data(spec)
  th = make_thetas(spec.dim, spec.p, C, 7).astype(np.float32).astype(np.float64)
  val, grad = eng.logprob_grad(th, with_prior=True)
  v2 = eng.logprob(th)
  ov, og = K.log_post_grad(prob, th)
  print(dt.__name__, "logprob max |dv|", np.abs(val - ov).max(), "max |dg|", np.abs(grad - og).max())
  d, st = eng.hmc_run(th, n_warmup=6, n_results=3, seed=1, max_leapfrog=3, init_step=0.01)
  print(" hmc finite", np.isfinite(d).all(), st["n_leapfrog"][:3])
  lv, tr, mn = eng.posterior_predict(th, seed=3)
  print(" predict finite", np.isfinite(lv).all(), np.isfinite(tr).all(), np.isfinite(mn).all())
  q = eng.row_quantiles(tr, [0.025, 0.975])
  print(" quantiles", q.shape, np.isfinite(q).all())
  eng.close()
