"""Developer tool: time the HMC kernel (steady-state evaluations) and the
single-launch log-prob for a few launch-shape knobs.  Not a test, not the bench.
  python tools/tune.py            (reads CI_B200_TEAM / CI_B200_G from the env)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
from causalimpact_b200 import _engine
from conftest import make_series, make_thetas

T, n_cov, C = 1000, 10, 256
y, X, _ = make_series(T, n_cov, 20242)
spec = cib.build_problem(y, X)
eng = cib.Engine(0)
eng.set_data(spec)
th = make_thetas(spec.dim, spec.p, C, 1)
kw = dict(n_warmup=100, n_results=100, seed=1, max_leapfrog=8, init_step=0.02)
eng.hmc_run(th, **dict(kw, n_warmup=5, n_results=5))
best = 1e9
for _ in range(3):
  t0 = time.perf_counter(); _, st = eng.hmc_run(th, **kw); best = min(best, time.perf_counter() - t0)
ev = st["n_leapfrog"].sum()
print(f"TEAM={os.environ.get('CI_B200_TEAM','1')} G={os.environ.get('CI_B200_G','auto')} "
      f"hmc {best*1e3:.3f} ms, {ev/best/1e6:.2f} M evals/s, {best/ (ev/C) * 1e6:.2f} us per eval per chain")
dev = torch.device("cuda", 0)
theta = torch.from_numpy(th.astype(np.float32)).to(dev)
val = torch.empty(C, dtype=torch.float32, device=dev); grad = torch.empty(C, spec.dim, dtype=torch.float32, device=dev)
s = torch.cuda.current_stream()
for variant in (1, 0):
  for _ in range(20):
    eng.logprob_grad_ptr(theta.data_ptr(), C, val.data_ptr(), grad.data_ptr(), variant, 1, s.cuda_stream)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(s)
  for _ in range(200):
    eng.logprob_grad_ptr(theta.data_ptr(), C, val.data_ptr(), grad.data_ptr(), variant, 1, s.cuda_stream)
  e1.record(s); torch.cuda.synchronize()
  print(f"  logprob variant={variant}: {e0.elapsed_time(e1)/200*1e3:.2f} us per launch (hot L2, back to back)")
