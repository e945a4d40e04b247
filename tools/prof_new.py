"""Developer tool: short runs of the round's new kernels for ncu / compute-sanitizer:
ci_impact (config-5 shape unless SMALL=1) and the seasonal Gibbs kernel."""
import os, sys, time, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
from causalimpact_b200 import model
from conftest import make_series
small = os.environ.get("SMALL") == "1"
eng = cib.Engine(0)
# ---- impact ----
S, T = (512, 300) if small else (10000, 2000)
rng = np.random.default_rng(0)
traj = torch.from_numpy(rng.normal(size=(S, T)).astype(np.float32)).cuda()
mean = traj.mean(0)
per = np.zeros(T, np.uint8); per[int(.7 * T):] = 1
obs = rng.normal(size=T) * 2 + 100
meta = types.SimpleNamespace(observed=obs, period=per, scale=2.0, offset=100.0, q_lo=0.025, q_hi=0.975,
                             obs_sum=float(obs[per == 1].sum()))
for _ in range(3):
  eng.impact(traj, mean, meta)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
  eng.impact(traj, mean, meta)
torch.cuda.synchronize()
print(f"impact S={S} T={T}: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms (incl. the D2H of the result)")
# ---- seasonal gibbs ----
Ts, C = (300, 16) if small else (1000, 256)
y, X, _ = make_series(Ts, 2 if small else 10, 20242)
spec = cib.build_problem(y, X)
eng.set_data(spec)
ss = [types.SimpleNamespace(num_seasons=7, num_steps_per_season=1),
      types.SimpleNamespace(num_seasons=4, num_steps_per_season=(2, 1, 1, 1))]
eng.set_seasonal(model.build_seasonal(ss, Ts, 1.0))
eng.gibbs_seasonal_run_t(C, n_warmup=1, n_results=1, seed=1)
torch.cuda.synchronize(); t0 = time.perf_counter()
nsw = 4 if small else 20
eng.gibbs_seasonal_run_t(C, n_warmup=nsw // 2, n_results=nsw // 2, seed=1)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"seasonal gibbs T={Ts} d=12 C={C}: {nsw} sweeps in {dt * 1e3:.2f} ms = {dt / nsw * 1e6:.0f} us per sweep")
