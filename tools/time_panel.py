"""Developer tool: where the time of fit_causalimpact_panel goes (second call, warm allocator)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
from causalimpact_b200 import panel as pn, _engine, impact as _impact

for Np in (128, 1024):
  rp = np.random.Generator(np.random.PCG64(77))
  Tp = 300
  xs = 100 + np.cumsum(rp.normal(size=(Np, Tp, 2)), axis=1) * 0.3
  y = xs[:, :, 0] + rp.normal(size=(Np, Tp)); y[:, 210:] += 3.0
  vals = np.concatenate([y[:, :, None], xs], axis=2)
  kw = dict(seed=1, inference_options=cib.InferenceOptions(num_results=400),
            engine_options=cib.EngineOptions(num_chains=8))
  for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cib.fit_causalimpact_panel(vals, np.arange(Tp), (0, 209), (210, Tp - 1), **kw)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"N={Np} call {rep}: {1e3 * (t1 - t0):.2f} ms = {Np / (t1 - t0):.0f} series/s", flush=True)
  # stages
  eng = cib.api._resolve_engine(kw["engine_options"])
  sync = torch.cuda.synchronize
  def T(label, fn):
    sync(); t = time.perf_counter(); r = fn(); sync()
    print(f"   {label}: {1e3 * (time.perf_counter() - t):.2f} ms", flush=True)
    return r
  lay = T("panel_layout", lambda: pn.panel_layout(np.arange(Tp), (0, 209), (210, Tp - 1)))
  v64 = T("ascontiguous", lambda: np.ascontiguousarray(vals, dtype=np.float64))
  stats = T("set_panel (H2D + prep kernel + D2H stats)", lambda: eng.set_panel(v64, row0=0, n_pre=210))
  out = T("gibbs_run_batch_t (150 sweeps)", lambda: eng.gibbs_run_batch_t(8, n_warmup=100, n_results=50, seed=1, series_stride=8))
  theta, level, traj, incl = out
  mean = T("predictive_mean_batch_t", lambda: eng.predictive_mean_batch_t(theta, level))
  per = np.zeros(Tp, np.uint8); per[210:] = 1
  obs = vals[:, :, 0].copy()
  r = T("impact_batch_t", lambda: eng.impact_batch_t(traj, mean, scale=stats[:, 0], offset=stats[:, 1],
                                                    obs_sum=obs[:, 210:].sum(1), observed=obs, period=per,
                                                    q_lo=0.025, q_hi=0.975))
  T("to_host series+summary", lambda: (eng.to_host(r[0]), eng.to_host(r[1])))
  T("to_host theta", lambda: eng.to_host(theta))
