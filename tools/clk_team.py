"""Developer tool: phase timeline of k_logpost_team (needs lib/libci_b200_clk.so built with
-DCI_CLK; run with CI_B200_LIB=tfp-causalimpact_b200/lib/libci_b200_clk.so)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
from causalimpact_b200 import _engine
from conftest import make_series, make_thetas
y, X, _ = make_series(1000, 10, 20242)
spec = cib.build_problem(y, X)
eng = cib.Engine(0); eng.set_data(spec)
C = 256
th = torch.from_numpy(make_thetas(spec.dim, spec.p, C, 1).astype(np.float32)).cuda()
val = torch.empty(C, device="cuda"); grad = torch.empty(C, spec.dim, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = ["entry", "eval start (theta, exp, tile wait done)", "residuals", "F1 scan", "sync1", "F2 (P seq, mean scan)",
         "sync2", "F3 + ll", "B1 scan", "sync3", "abar seq + B2 scan", "sync4", "Pbar seq", "X'rbar",
         "final sync+sums", "prior + store"]
for mode in ("hot", "flushed"):
  for _ in range(3):
    if mode == "flushed": flush.zero_()
    eng.logprob_grad_ptr(th.data_ptr(), C, val.data_ptr(), grad.data_ptr(), 1, 1)
  torch.cuda.synchronize()
  out = (ctypes.c_longlong * 32)()
  assert eng._lib.ci_debug_clocks(out) == 0
  clk = np.array(out).reshape(2, 16)
  for w, label in ((0, "warp 0"), (1, "warp W-1")):
    d = np.diff(clk[w]); tot = clk[w][15] - clk[w][0]
    print(f"[{mode}] {label}: total {tot} cycles")
    for i in range(15):
      print(f"   {names[i + 1]:<45} {d[i]:>7}")
