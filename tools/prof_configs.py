"""Developer tool: a few launches of the dominant kernel of one BASELINE config, for ncu.
  python tools/prof_configs.py cfg2|cfg3|cfg4|cfg5|gibbs"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import bench
import causalimpact_b200 as cib
which = sys.argv[1]
dev = torch.device("cuda", 0)
s = torch.cuda.current_stream()
eng = cib.Engine(0)
if which in ("cfg2", "cfg3", "cfg4"):
  cfg = bench.CONFIGS[{"cfg2": 0, "cfg3": 1, "cfg4": 2}[which]]
  y, X, th = bench.make_inputs(cfg)
  spec = cib.build_problem(y, X, model=cfg["model"])
  eng.set_data(spec)
  C = cfg["chains"]
  theta = torch.from_numpy(th.astype(np.float32)).to(dev)
  val = torch.empty(C, dtype=torch.float32, device=dev)
  grad = torch.empty(C, spec.dim, dtype=torch.float32, device=dev)
  for _ in range(4):
    eng.logprob_grad_ptr(theta.data_ptr(), C, val.data_ptr(), grad.data_ptr(), 1, 1, s.cuda_stream)
elif which == "cfg5":
  cfg = bench.CONFIGS[3]
  y, X, th = bench.make_inputs(cfg)
  eng.set_data(cib.build_problem(y, X))
  for _ in range(4):
    eng.posterior_predict_t(th.astype(np.float32), seed=11)
elif which == "gibbs":
  cfg = bench.CONFIGS[0]
  y, X, th = bench.make_inputs(cfg)
  eng.set_data(cib.build_problem(y, X))
  for _ in range(3):
    eng.gibbs_run_t(256, n_warmup=10, n_results=5, seed=1)
torch.cuda.synchronize()
print("ok", which)
