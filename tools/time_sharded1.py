"""ci_impact_sharded_d on ONE rank (self exchange) vs ci_impact_d on the same draws: isolates the
cost of the sharded code path (window stores, column blocks) from the multi-GPU effects.
CI_B200_TRACE=1 prints the per-step device times."""
import os, sys, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tfp-causalimpact_b200"))
import torch
import causalimpact_b200 as cib

S, T, t_pre = 10000, 2000, 1400
eng = cib.Engine(0)
comm = cib.Comm(eng, cib.comm_unique_id(), 0, 1)
rng = np.random.default_rng(0)
per = np.zeros(T, np.uint8); per[t_pre:] = 1
obs = rng.normal(size=T)
meta = types.SimpleNamespace(observed=obs, period=per, scale=2.0, offset=100.0, q_lo=0.025, q_hi=0.975,
                             obs_sum=float(obs[t_pre:].sum()))
traj = torch.randn(S, T, device="cuda"); mean = traj.mean(0)
out = torch.empty(T * 9 + 20, dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, n=12):
  ts = []
  for _ in range(n):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
  return float(np.median(ts[2:])) * 1e3
print("ci_impact_d          %.1f us" % timed(lambda: eng.impact(traj, mean, meta, out=out)))
print("ci_impact_sharded_d  %.1f us (1 rank)" % timed(lambda: comm.impact_sharded_t(traj, mean, meta, [S])))
comm.close(); eng.close()
