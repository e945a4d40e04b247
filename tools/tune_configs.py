"""Developer tool: per-launch times of the other BASELINE configs (3, 4, 5).
They are parity-test cases, not bench lines; the numbers go into DESIGN.md."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
from causalimpact_b200 import _engine
from conftest import make_series, make_thetas
from oracle import c_port, kalman_np as K
import bench as _bench

dev = torch.device("cuda", 0)
s = torch.cuda.current_stream()
eng = cib.Engine(0)


def time_logprob(name, T, n_cov, C, model, reps=20):
  y, X, _ = make_series(T, n_cov, 20240 + T)
  spec = cib.build_problem(y, X, model=model)
  eng.set_data(spec)
  th = make_thetas(spec.dim, spec.p, C, 1, d=spec.d)
  theta = torch.from_numpy(th.astype(np.float32)).to(dev)
  val = torch.empty(C, dtype=torch.float32, device=dev)
  grad = torch.empty(C, spec.dim, dtype=torch.float32, device=dev)
  for _ in range(3):
    eng.logprob_grad_ptr(theta.data_ptr(), C, val.data_ptr(), grad.data_ptr(), 1, 1, s.cuda_stream)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(s)
  for _ in range(reps):
    eng.logprob_grad_ptr(theta.data_ptr(), C, val.data_ptr(), grad.data_ptr(), 1, 1, s.cuda_stream)
  e1.record(s); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / reps
  B = 8 * T * (spec.p + 1) + 4 * (2 * spec.dim + 1)
  line = f"{name}: T={T} p={spec.p} C={C}: {ms*1e3:.1f} us/launch, {C/ms*1e3/1e6:.3f} M evals/s, algorithmic {C*B/ms/1e6:.0f} GB/s"
  if model == 0:
    prob = K.default_problem(y, X)
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < 3.0:
      c_port.logpost_grad(prob, th, nthreads=_bench.usable_cpus()); n += C
    line += f"; CPU port {n/(time.perf_counter()-t0)/1e6:.4f} M evals/s on {_bench.usable_cpus()} threads (cgroup-aware)"
  print(line, flush=True)


time_logprob("config 2 (team)", 1000, 10, 256, 0, reps=100)
time_logprob("config 3 (LLT, streamed)", 5000, 50, 1024, 1)
time_logprob("config 4 (long series, streamed)", 20000, 1, 512, 0)
# config 5: 10000-draw forecast, T=2000
y, X, _ = make_series(2000, 10, 20245)
spec = cib.build_problem(y, X); eng.set_data(spec)
S = 10000
th = np.tile(make_thetas(spec.dim, spec.p, 1, 6), (S, 1)).astype(np.float32)
thd = torch.from_numpy(th).to(dev)
lvl = torch.empty(S, 2000, dtype=torch.float32, device=dev); trj = torch.empty_like(lvl)
mean = torch.empty(2000, dtype=torch.float32, device=dev)
q = np.array([0.025, 0.975]); qd = torch.empty(2000, 2, dtype=torch.float32, device=dev)
import ctypes
lib, ctx = eng._lib, eng._ctx
def run():
  assert lib.ci_posterior_predict_d(ctx, thd.data_ptr(), S, 7, 0, lvl.data_ptr(), trj.data_ptr(), mean.data_ptr(), s.cuda_stream) == 0
  assert lib.ci_row_quantiles_d(ctx, trj.data_ptr(), S, 2000, 0, q.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 2, qd.data_ptr(), s.cuda_stream) == 0
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(5): run()
e1.record(s); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
Bd = 4 * 2000 * (spec.p + 1) + 8 * 2000 + 8 * 600
print(f"config 5: 10000 draws T=2000 p=11 predict+quantiles: {ms:.3f} ms, {S/ms*1e3/1e6:.2f} M draws/s, algorithmic {S*Bd/ms/1e6:.0f} GB/s")
