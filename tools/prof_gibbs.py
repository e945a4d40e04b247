"""Developer tool: one short Gibbs run (for `ncu -k k_gibbs`) + timing."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import causalimpact_b200 as cib
from conftest import make_series
y, X, _ = make_series(1000, 10, 20242)
spec = cib.build_problem(y, X)
eng = cib.Engine(0); eng.set_data(spec)
eng.gibbs_run(256, n_warmup=2, n_results=2, seed=1, want_level=False, want_traj=False)
for sparse in (True, False):
  t0 = time.perf_counter()
  d, _, _, incl = eng.gibbs_run(256, n_warmup=20, n_results=20, seed=1, sparse=sparse, want_level=False, want_traj=False)
  dt = time.perf_counter() - t0
  print(f"sparse={sparse}: 40 sweeps x 256 chains in {dt*1e3:.2f} ms = {dt/40*1e6:.1f} us per sweep; incl", incl.mean(0).round(2))
