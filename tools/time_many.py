"""Developer tool: throughput of the batched fit (series per second)."""
import os, sys, time
import numpy as np, pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
N, T, k = int(os.environ.get("N", "256")), int(os.environ.get("T", "300")), 2
rng = np.random.default_rng(0)
idx = pd.date_range("2021-01-01", periods=T, freq="D")
dfs = []
for s in range(N):
  xs = 100 + np.cumsum(rng.normal(size=(T, k)), axis=0) * 0.3
  y = xs[:, 0] + rng.normal(size=T); y[int(.7 * T):] += 3
  dfs.append(pd.DataFrame(np.column_stack([y, xs]), index=idx, columns=["y", "a", "b"]))
pre, post = (idx[0], idx[int(.7 * T) - 1]), (idx[int(.7 * T)], idx[-1])
kw = dict(seed=1, inference_options=cib.InferenceOptions(num_results=400),
          engine_options=cib.EngineOptions(num_chains=8, sampler="gibbs"))
cib.fit_causalimpact_many(dfs[:4], pre, post, **kw)
t0 = time.perf_counter(); res = cib.fit_causalimpact_many(dfs, pre, post, **kw); dt = time.perf_counter() - t0
print(f"fit_causalimpact_many: {N} series (T={T}, {k} covariates, 8 chains, 400 draws) in {dt*1e3:.0f} ms = {N/dt:.0f} series/s")
t0 = time.perf_counter()
for d in dfs[:16]:
  cib.fit_causalimpact(d, pre, post, **kw)
dt1 = (time.perf_counter() - t0) / 16
print(f"fit_causalimpact one by one: {dt1*1e3:.1f} ms per series = {1/dt1:.0f} series/s")
# device part alone
from causalimpact_b200 import frame as fr
eng = cib.Engine(0)
specs = []
for d in dfs:
  cid = fr.CausalImpactData(d, pre, post)
  y_ext, design, sd = cid.engine_inputs(np.float32)
  specs.append(cib.build_problem(y_ext, design, outcome_sd=sd))
eng.set_data_batch(specs)
eng.gibbs_run_batch_t(8, n_warmup=2, n_results=2, seed=1); torch.cuda.synchronize()
t0 = time.perf_counter(); eng.gibbs_run_batch_t(8, n_warmup=100, n_results=50, seed=1); torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"ci_gibbs_run_batch_d: {N} series x 8 chains x 150 sweeps in {dt*1e3:.1f} ms = {N*8*150/dt/1e6:.2f} M sweeps/s")
vals = np.stack([d.values for d in dfs])
cib.fit_causalimpact_panel(vals[:4], idx, pre, post, **kw)
t0 = time.perf_counter(); res = cib.fit_causalimpact_panel(vals, idx, pre, post, **kw); dt = time.perf_counter() - t0
print(f"fit_causalimpact_panel: {N} series in {dt*1e3:.0f} ms = {N/dt:.0f} series/s")
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable(); cib.fit_causalimpact_panel(vals, idx, pre, post, **kw); pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative"); buf = io.StringIO(); st.stream = buf; st.print_stats(14); print("\n".join(buf.getvalue().splitlines()[6:24]))
