"""Developer tool: time ci_posterior_predict_d for the library named by CI_B200_LIB."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
from conftest import make_series, make_thetas
from oracle import kalman_np as K, smoother_np as SM
dev = torch.device("cuda", 0); s = torch.cuda.current_stream()
eng = cib.Engine(0); lib, ctx = eng._lib, eng._ctx
for T, S in ((1000, 4096), (2000, 10000), (100, 900)):
  y, X, _ = make_series(T, 10, 20245)
  spec = cib.build_problem(y, X); eng.set_data(spec)
  th = np.tile(make_thetas(spec.dim, spec.p, 1, 6), (S, 1)).astype(np.float32)
  thd = torch.from_numpy(th).to(dev)
  lvl = torch.empty(S, T, dtype=torch.float32, device=dev); trj = torch.empty_like(lvl)
  mean = torch.empty(T, dtype=torch.float32, device=dev)
  run = lambda: lib.ci_posterior_predict_d(ctx, thd.data_ptr(), S, 7, 0, lvl.data_ptr(), trj.data_ptr(), mean.data_ptr(), s.cuda_stream)
  for _ in range(3): assert run() == 0
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(s)
  for _ in range(20): run()
  e1.record(s); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 20
  # accuracy vs the oracle on 4 draws
  prob = K.default_problem(y, X)
  ol, ot, _ = SM.posterior_predict(prob, th[:4].astype(np.float64), seed=7)
  err = np.abs(trj[:4].cpu().numpy() - ot).max()
  print(f"{os.environ.get('CI_B200_LIB','default').split('/')[-1]}: T={T} S={S}: {ms*1e3:.1f} us, {S/ms*1e3/1e6:.2f} M draws/s, max |traj - oracle| = {err:.2e}", flush=True)
