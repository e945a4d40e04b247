"""Developer tool: one short seasonal Gibbs run for `ncu -k regex:k_gibbs_seasonal`."""
import os, sys, time, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
from causalimpact_b200 import model
from conftest import make_series
T, C, ncov = int(os.environ.get("T", "1000")), int(os.environ.get("C", "256")), int(os.environ.get("NCOV", "10"))
y, X, _ = make_series(T, ncov, 20242)
eng = cib.Engine(0)
eng.set_data(cib.build_problem(y, X))
ss = [types.SimpleNamespace(num_seasons=7, num_steps_per_season=1),
      types.SimpleNamespace(num_seasons=4, num_steps_per_season=(2, 1, 1, 1))]
eng.set_seasonal(model.build_seasonal(ss, T, 1.0))
n = int(os.environ.get("SWEEPS", "4"))
eng.gibbs_seasonal_run_t(C, n_warmup=1, n_results=1, seed=1)
torch.cuda.synchronize(); t0 = time.perf_counter()
eng.gibbs_seasonal_run_t(C, n_warmup=n // 2, n_results=n // 2, seed=1)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"seasonal gibbs T={T} d=12 C={C}: {n} sweeps in {dt * 1e3:.2f} ms = {dt / n * 1e6:.0f} us per sweep")
