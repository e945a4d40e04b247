"""Repro harness for compute-sanitizer runs (developer tool, not a test)."""
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
spec = cib.build_problem(y, X)
eng = cib.Engine(0)
eng.set_data(spec)
th = make_thetas(spec.dim, spec.p, C, 7).astype(np.float32).astype(np.float64)
val, grad = eng.logprob_grad(th, with_prior=True)
prob = K.default_problem(y, X)
ov, og = K.log_post_grad(prob, th)
print("max |dv|", np.abs(val - ov).max(), "max |dg|", np.abs(grad - og).max())
