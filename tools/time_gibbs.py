"""Developer tool: wall time per Gibbs sweep (ci_gibbs_run_d, device resident) at BASELINE
configs[1] (T=1000, 10 covariates) for the team and the one-warp kernels."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
from conftest import make_series

def run(T, n_cov, C, env, sweeps=100, sparse=True):
  for k in ("CI_B200_GIBBS_TEAM", "CI_B200_G"):
    os.environ.pop(k, None)
  os.environ.update({k: str(v) for k, v in env.items()})
  eng = cib.Engine(0)
  y, X, _ = make_series(T, n_cov, 20240 + T)
  eng.set_data(cib.build_problem(y, X))
  eng.gibbs_run_t(C, n_warmup=5, n_results=2, seed=1, sparse=sparse)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  eng.gibbs_run_t(C, n_warmup=sweeps - 10, n_results=10, seed=1, sparse=sparse)
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  eng.close()
  return ms * 1e3 / sweeps

for T, n_cov, C in ((1000, 10, 256), (1000, 10, 1024), (2000, 10, 256), (300, 2, 1024), (100, 1, 64)):
  for sparse in (True, False):
    base = run(T, n_cov, C, {"CI_B200_GIBBS_TEAM": 0}, sparse=sparse)
    line = f"T={T} cov={n_cov} C={C} sparse={sparse}: one-warp {base:.1f} us/sweep"
    for G in (0, 1, 2, 4):
      env = {"CI_B200_GIBBS_TEAM": 1}
      if G: env["CI_B200_G"] = G
      t = run(T, n_cov, C, env, sparse=sparse)
      line += f" | team G={G or 'auto'} {t:.1f}"
    print(line, flush=True)
