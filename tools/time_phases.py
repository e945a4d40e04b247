"""Developer tool: device time of each phase of the fit at the config-5 scale
(T=2000, 10 covariates, 256 chains, 10000 draws)."""
import os, sys, time
import numpy as np, pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
from causalimpact_b200 import frame as fr, api
T = int(os.environ.get("T", "2000")); k = int(os.environ.get("K", "10"))
rng = np.random.default_rng(0)
xs = 100 + np.cumsum(rng.normal(size=(T, k)), axis=0) * 0.3
y = 1.2 * xs[:, 0] + 0.6 * xs[:, 1] - 0.4 * xs[:, 2] + rng.normal(size=T); y[int(.7*T):] += 8
df = pd.DataFrame(np.column_stack([y, xs]), columns=["y"] + [f"x{i}" for i in range(k)])
cid = fr.CausalImpactData(df, (0, int(.7*T) - 1), (int(.7*T), T - 1))
y_ext, design, sd = cid.engine_inputs(np.float32)
spec = cib.build_problem(y_ext, design, prior_level_sd=0.01, outcome_sd=sd, dtype=np.float32)
eng = cib.Engine(0); eng.set_data(spec)
def timed(name, f, n=3):
  f(); torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(n): r = f()
  torch.cuda.synchronize()
  print(f"{name}: {(time.perf_counter()-t0)/n*1e3:.2f} ms"); return r
C = 256
for nw, nr in ((100, 40), (0, 40), (100, 1), (200, 1)):
  timed(f"gibbs C={C} warm={nw} res={nr}", lambda: eng.gibbs_run_t(C, n_warmup=nw, n_results=nr, seed=1))
th, lv, tr, _ = eng.gibbs_run_t(C, n_warmup=100, n_results=40, seed=1)
th0 = np.tile(cib.initial_theta(spec, 0.01), (C, 1)); th0 += 0.05 * rng.normal(size=th0.shape)
for nw, nr in ((300, 40), (300, 1), (100, 1)):
  timed(f"hmc C={C} warm={nw} res={nr}", lambda: eng.hmc_run_t(th0, n_warmup=nw, n_results=nr, seed=1))
timed("predict 10240", lambda: eng.posterior_predict_t(th, seed=3))
timed("mean", lambda: eng.predictive_mean_t(th, lv))
timed("to_host level", lambda: eng.to_host(lv))
pin = torch.empty(lv.shape, dtype=lv.dtype).pin_memory()
timed("to pinned host level", lambda: pin.copy_(lv))
meta = __import__("causalimpact_b200.impact", fromlist=["x"]).prepare(cid, 0.05)
mean = eng.predictive_mean_t(th, lv)
timed("impact", lambda: eng.impact(tr, mean, meta))
