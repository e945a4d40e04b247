"""k_impact_rows alone (ci_impact_rows_d): time vs draws, with / without the predictive-mean CTA.
CI_B200_IMP_SEG (pre-period steps per transpose-only CTA) is read once per process."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tfp-causalimpact_b200"))
import torch
import causalimpact_b200 as cib
from causalimpact_b200 import impact as _imp

T, t_pre = 2000, 1400
eng = cib.Engine(0)
rng = np.random.default_rng(0)
per = np.zeros(T, np.uint8); per[t_pre:] = 1
obs = rng.normal(size=T)
meta = _imp.ImpactMeta(index=None, observed=obs, period=per, hide=None, scale=2.0, offset=100.0,
                       q_lo=0.025, q_hi=0.975, obs_mean=0.0, obs_sum=float(obs[t_pre:].sum()))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = torch.zeros(T * 9 + 20, dtype=torch.float64, device="cuda")
print("IMP_SEG", os.environ.get("CI_B200_IMP_SEG", "default"))
for S in ([int(x) for x in sys.argv[1:]] or (1250, 5000, 10000, 40000)):
  traj = torch.randn(S, T, device="cuda")
  mean = traj.mean(0)
  for with_mean in (True, False):
    ts = []
    for it in range(12):
      flush.zero_()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(); eng.impact_rows_t(traj, mean if with_mean else None, meta, out); e1.record()
      torch.cuda.synchronize()
      ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[2:]))
    gb = S * (T * 8 + (T - t_pre) * 8) / 1e9
    print(f"  S={S:6d} mean={int(with_mean)}: {ms * 1e3:7.1f} us  ({gb / (ms * 1e-3):7.0f} GB/s algorithmic)")
