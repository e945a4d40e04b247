"""Developer tool: one small pass over every entry point added in session 2, for
compute-sanitizer (impact, seasonal Gibbs, batch Gibbs + select, panel fit)."""
import os, sys, types
import numpy as np, pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
rng = np.random.default_rng(0)
idx = pd.date_range("2021-01-01", periods=90, freq="D")
vals = np.empty((3, 90, 3))
for s in range(3):
  xs = 100 + np.cumsum(rng.normal(size=(90, 2)), axis=0) * 0.3
  y = xs[:, 0] + rng.normal(size=90); y[60:] += 3
  vals[s] = np.column_stack([y, xs])
kw = dict(seed=1, inference_options=cib.InferenceOptions(num_results=24),
          engine_options=cib.EngineOptions(num_chains=4, gibbs_min_warmup=6, min_warmup=20))
r = cib.fit_causalimpact_panel(vals, idx, (idx[0], idx[59]), (idx[60], idx[-1]), **kw)
dfs = [pd.DataFrame(v, index=idx, columns=["y", "a", "b"]) for v in vals]
m = cib.fit_causalimpact_many(dfs, (idx[0], idx[59]), (idx[60], idx[-1]), **kw)
one = cib.fit_causalimpact(dfs[0], (idx[0], idx[59]), (idx[60], idx[-1]), **kw)
sea = cib.fit_causalimpact(dfs[1], (idx[0], idx[59]), (idx[60], idx[-1]),
                           model_options=cib.ModelOptions(seasons=[cib.Seasons(num_seasons=7)]), **kw)
print("ok", r.summary[0, 0, 5], float(m[0].summary.loc["average", "abs_effect"]),
      float(one.summary.loc["average", "abs_effect"]), float(sea.summary.loc["average", "abs_effect"]))
