"""Developer tool: a few launches of the long-series team kernel at BASELINE configs[3]
(T=20000, 1 covariate, 512 chains) for ncu."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
from conftest import make_series, make_thetas
T, n_cov, C = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (20000, 1, 512)))
dev = torch.device("cuda", 0)
s = torch.cuda.current_stream()
eng = cib.Engine(0)
y, X, _ = make_series(T, n_cov, 20240 + T)
spec = cib.build_problem(y, X)
eng.set_data(spec)
th = make_thetas(spec.dim, spec.p, C, 1)
theta = torch.from_numpy(th.astype(np.float32)).to(dev)
val = torch.empty(C, dtype=torch.float32, device=dev)
grad = torch.empty(C, spec.dim, dtype=torch.float32, device=dev)
for _ in range(4):
  eng.logprob_grad_ptr(theta.data_ptr(), C, val.data_ptr(), grad.data_ptr(), 1, 1, s.cuda_stream)
torch.cuda.synchronize()
print("ok", float(val.sum()))
