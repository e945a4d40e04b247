"""Developer tool: one small pass of the impact-stage kernels for compute-sanitizer (memcheck /
racecheck): k_impact_rows (local and exchange-window destinations, pre-period segments, mean row),
k_impact_jobs (contiguous columns and column blocks), k_merge_blocks, k_mean_combine,
k_mean_partial / k_mean_final, the batched launches."""
import os, sys, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import causalimpact_b200 as cib
from conftest import make_series, make_thetas

rng = np.random.default_rng(1)
eng = cib.Engine(0)
comm = cib.Comm(eng, cib.comm_unique_id(), 0, 1)
for dtype, S, T, t0 in ((np.float32, 300, 517, 310), (np.float64, 97, 200, 130), (np.float32, 33, 70, 40)):
  per = np.zeros(T, np.uint8); per[t0:] = 1
  obs = rng.normal(size=T); obs[5] = np.nan; obs[t0 + 2] = np.nan
  meta = types.SimpleNamespace(observed=obs, period=per, scale=2.0, offset=10.0, q_lo=0.025, q_hi=0.975,
                               obs_sum=float(np.nansum(obs[per == 1])))
  traj = torch.from_numpy(rng.normal(size=(S, T)).astype(dtype)).cuda()
  mean = traj.mean(0)
  s9, summ = eng.impact(traj, mean, meta)
  out, full = comm.impact_sharded_t(traj, mean, meta, [S])
  got = out.cpu().numpy()
  assert np.array_equal(got[:T * 9].reshape(T, 9), s9, equal_nan=True) and np.array_equal(got[T * 9:], summ, equal_nan=True)
  outp = torch.zeros(T * 9 + 20, dtype=torch.float64, device="cuda")
  trT, cumT, stats, packed = eng.impact_rows_t(traj, mean, meta, outp)
  eng.impact_cols_t(trT, 0, cumT, 0, stats, meta, outp)
  assert np.array_equal(outp.cpu().numpy()[:T * 9].reshape(T, 9), s9, equal_nan=True)
  print("impact ok", dtype.__name__, S, T)
comm.close()
y, X, _ = make_series(700, 3, 1)
spec = cib.build_problem(y, X)
eng.set_data(spec)
th = torch.from_numpy(make_thetas(spec.dim, spec.p, 150, 1).astype(np.float32)).cuda()
lvl, trj = eng.posterior_predict_t(th, seed=3)
m = eng.predictive_mean_t(th, lvl)
print("mean ok", float(m.sum()))
eng.close()
