"""Developer tool: one short HMC run (for `ncu -k regex:k_hmc`)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import causalimpact_b200 as cib
from conftest import make_series, make_thetas
y, X, _ = make_series(1000, 10, 20242)
spec = cib.build_problem(y, X)
eng = cib.Engine(0); eng.set_data(spec)
th = make_thetas(spec.dim, spec.p, 256, 1)
d, st = eng.hmc_run(th, n_warmup=20, n_results=10, seed=1, max_leapfrog=8, init_step=0.02)
print("evals per chain", st["n_leapfrog"][0], "finite", np.isfinite(d).all())
