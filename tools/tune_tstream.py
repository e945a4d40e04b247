"""Developer tool: BASELINE configs[3] (T=20000, 1 covariate, 512 chains) value+gradient
launch time of the long-series team kernels for several (warps per chain, teams per CTA),
against the one-warp-per-chain path.  CUDA events, L2 flushed between launches."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
from conftest import make_series, make_thetas

dev = torch.device("cuda", 0)
s = torch.cuda.current_stream()
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
shapes = [("cfg4", 20000, 1, 512), ("T5000_p11", 5000, 10, 1024), ("T20000_p11", 20000, 10, 512)]
if len(sys.argv) > 1:
  shapes = [sh for sh in shapes if sh[0] in sys.argv[1:]]


def run(T, n_cov, C, env, reps=10):
  for k in ("CI_B200_TSTREAM", "CI_B200_TSW", "CI_B200_G"):
    os.environ.pop(k, None)
  os.environ.update({k: str(v) for k, v in env.items()})
  eng = cib.Engine(0)
  y, X, _ = make_series(T, n_cov, 20240 + T)
  spec = cib.build_problem(y, X)
  eng.set_data(spec)
  th = make_thetas(spec.dim, spec.p, C, 1)
  theta = torch.from_numpy(th.astype(np.float32)).to(dev)
  val = torch.empty(C, dtype=torch.float32, device=dev)
  grad = torch.empty(C, spec.dim, dtype=torch.float32, device=dev)
  try:
    for _ in range(3):
      eng.logprob_grad_ptr(theta.data_ptr(), C, val.data_ptr(), grad.data_ptr(), 1, 1, s.cuda_stream)
    torch.cuda.synchronize()
  except cib.EngineError as e:
    eng.close()
    return None, str(e)
  ts = []
  for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    eng.logprob_grad_ptr(theta.data_ptr(), C, val.data_ptr(), grad.data_ptr(), 1, 1, s.cuda_stream)
    e1.record(s); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
  eng.close()
  return float(np.median(ts)), float(val.sum().item())


for name, T, n_cov, C in shapes:
  base, chk = run(T, n_cov, C, {"CI_B200_TSTREAM": 0})
  print(f"{name}: one warp per chain: {base:.1f} us  (sum of values {chk:.3f})", flush=True)
  for W in (2, 3, 4, 5, 8):
    for G in (0, 1, 2, 3, 4, 5, 8):
      if G * W > 16:
        continue
      env = {"CI_B200_TSTREAM": 1, "CI_B200_TSW": W}
      if G:
        env["CI_B200_G"] = G
      t, chk = run(T, n_cov, C, env)
      print(f"  W={W} G={G or 'auto'}: " + (f"{t:.1f} us  x{base / t:.2f}  (sum {chk:.3f})" if t else f"n/a ({chk})"),
            flush=True)
