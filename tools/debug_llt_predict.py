import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import causalimpact_b200 as cib
from conftest import make_series, make_thetas
from oracle import kalman_np as K, smoother_np as SM
T, n_cov = 700, 3
y, X, _ = make_series(T, n_cov, 70 + T, nan_frac=0.03)
spec = cib.build_problem(y, X, prior_level_sd=0.05, model=1, dtype=np.float32)
eng = cib.Engine(0); eng.set_data(spec)
prob = K.default_problem(y, X, prior_level_sd=0.05, model=1)
th = make_thetas(spec.dim, spec.p, 6, 3, d=2).astype(np.float32).astype(np.float64)
th[:, spec.p + 1] = np.log(0.05 ** 2) + 0.3 * np.arange(6)
th[:, spec.p + 2] = np.log(0.01 ** 2) + 0.5 * np.arange(6)
level, traj, mean = eng.posterior_predict(th, seed=21, draw_id0=9)
ol, _, ot, om = SM.posterior_predict_llt(prob, th, 21, 9)
d = np.abs(traj - ot); dl = np.abs(level - ol)
bad = np.argwhere(d > 5e-3)
print("n bad", len(bad), "max level diff", dl.max())
print("bad draws", np.unique(bad[:, 0], return_counts=True))
print("bad t (first 40)", bad[:40, 1])
print("nan at bad t?", np.isnan(y)[bad[:40, 1]])
s, t = bad[0]
print("example", s, t, traj[s, t], ot[s, t], level[s, t], ol[s, t], (X @ th[s, :spec.p])[t])
zs, zp = SM.predict_normals_llt(21, 9 + s, T)
print("noise gpu", (traj[s, t] - level[s, t] - (X @ th[s, :spec.p])[t]) / np.exp(0.5 * th[s, spec.p]), "oracle zp", zp[t])
