"""Developer tool: launch shape of the select kernels inside ci_impact_d / ci_impact_batch_d.
Times (CUDA events) the impact stage at BASELINE configs[4] size (S=10000, T=2000) and the batched
panel call (N=1024, S=400, T=300) for thread counts / key staging choices."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch, types
import causalimpact_b200 as cib

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)

def ev_time(fn, reps=10):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps

def single(env):
  for k in ("CI_B200_SEL_NT", "CI_B200_SEL_SMEM"): os.environ.pop(k, None)
  os.environ.update({k: str(v) for k, v in env.items()})
  eng = cib.Engine(0)
  S, T = 10000, 2000
  traj = torch.randn(S, T, device=dev); mean = traj.mean(0)
  per = np.zeros(T, np.uint8); per[1400:] = 1
  obs = rng.normal(size=T)
  meta = types.SimpleNamespace(observed=obs, period=per, scale=2.0, offset=100.0, q_lo=0.025, q_hi=0.975,
                               obs_sum=float(obs[1400:].sum()))
  out = torch.empty(T * 9 + 20, dtype=torch.float64, device=dev)
  t = ev_time(lambda: eng.impact(traj, mean, meta, out=out))
  chk = float(out[:T * 9].reshape(T, 9)[:, 1].sum())
  eng.close()
  return t, chk

def batch(env):
  for k in ("CI_B200_SEL_NT", "CI_B200_SEL_SMEM"): os.environ.pop(k, None)
  os.environ.update({k: str(v) for k, v in env.items()})
  eng = cib.Engine(0)
  N, S, T = 1024, 400, 300
  vals = np.concatenate([rng.normal(size=(N, T, 1)) , 100 + rng.normal(size=(N, T, 2))], axis=2)
  eng.set_panel(vals, row0=0, n_pre=210)
  traj = torch.randn(N, S, T, device=dev); mean = traj.mean(1)
  per = np.zeros(T, np.uint8); per[210:] = 1
  obs = rng.normal(size=(N, T))
  kw = dict(scale=np.ones(N), offset=np.zeros(N), obs_sum=obs[:, 210:].sum(1), observed=obs, period=per,
            q_lo=0.025, q_hi=0.975)
  t = ev_time(lambda: eng.impact_batch_t(traj, mean, **kw), reps=5)
  ser, _ = eng.impact_batch_t(traj, mean, **kw)
  chk = float(ser.reshape(N, T, 9)[:, :, 1].sum())
  eng.close()
  return t, chk

for name, fn in (("single S=10000 T=2000", single), ("batch N=1024 S=400 T=300", batch)):
  t, chk = fn({})
  print(f"{name}: default {t:.3f} ms (chk {chk:.4f})", flush=True)
  for sm in (1, 0):
    for nt in (64, 128, 256, 512, 1024):
      t, chk = fn({"CI_B200_SEL_NT": nt, "CI_B200_SEL_SMEM": sm})
      print(f"   smem={sm} nt={nt}: {t:.3f} ms (chk {chk:.4f})", flush=True)
