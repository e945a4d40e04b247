"""Developer tool: wall time of the public fit at the config-5 scale
(T=2000, 10 covariates, 10000 draws), with a per-phase breakdown."""
import os, sys, time, cProfile, pstats
import numpy as np, pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200")):
  sys.path.insert(0, p)
import causalimpact_b200 as cib
rng = np.random.default_rng(0)
n, k = 2000, 10
xs = 100 + np.cumsum(rng.normal(size=(n, k)), axis=0) * 0.3
y = 1.2 * xs[:, 0] + 0.6 * xs[:, 1] - 0.4 * xs[:, 2] + rng.normal(size=n); y[1400:] += 8
df = pd.DataFrame(np.column_stack([y, xs]), columns=["y"] + [f"x{i}" for i in range(k)])
for sampler in ("gibbs", "hmc"):
  kw = dict(inference_options=cib.InferenceOptions(num_results=10000),
            engine_options=cib.EngineOptions(num_chains=256, sampler=sampler))
  cib.fit_causalimpact(df, (0, 1399), (1400, 1999), seed=1,
                       inference_options=cib.InferenceOptions(num_results=100),
                       engine_options=cib.EngineOptions(num_chains=16, sampler=sampler))   # warm
  t0 = time.perf_counter()
  pr = cProfile.Profile(); pr.enable()
  res = cib.fit_causalimpact(df, (0, 1399), (1400, 1999), seed=1, **kw)
  pr.disable()
  dt = time.perf_counter() - t0
  print(f"== sampler={sampler}: fit wall {dt*1e3:.0f} ms; abs_effect {float(res.summary.loc['average','abs_effect']):.3f} "
        f"[{float(res.summary.loc['average','abs_effect_lower']):.3f}, {float(res.summary.loc['average','abs_effect_upper']):.3f}]")
  st = pstats.Stats(pr); st.sort_stats("cumulative")
  import io; buf = io.StringIO(); st.stream = buf; st.print_stats(14); print("\n".join(buf.getvalue().splitlines()[6:26]))
