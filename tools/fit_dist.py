"""Multi-GPU check of the product call (run under torchrun with nccl, and once
single-process): the sharded fit must equal the single-GPU fit bit for bit.
  torchrun --nproc-per-node 2 tools/fit_dist.py out2.pkl ; python tools/fit_dist.py out1.pkl
  python tools/fit_dist.py --compare out1.pkl out2.pkl"""
import os, pickle, sys
import numpy as np, pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200")):
  sys.path.insert(0, p)

if sys.argv[1] == "--compare":
  a, b = (pickle.load(open(f, "rb")) for f in sys.argv[2:4])
  for name in a:
    for k in a[name]:
      same = np.array_equal(a[name][k], b[name][k], equal_nan=True)
      print(name, k, "bit-identical" if same else "DIFFERENT")
      assert same
  print("multi-GPU fit == single-GPU fit")
  sys.exit(0)

import torch
import torch.distributed as dist
import causalimpact_b200 as cib
world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
  dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {}
for name, n_cov in (("hmc_path", 1), ("gibbs_path", 6), ("seasonal_path", 1)):
  rng = np.random.default_rng(5)
  n = 400
  xs = 100 + np.cumsum(rng.normal(size=(n, n_cov)), axis=0) * 0.3
  y = 1.2 * xs[:, 0] + rng.normal(size=n); y[280:] += 6
  df = pd.DataFrame(np.column_stack([y, xs]), columns=["y"] + [f"x{i}" for i in range(n_cov)])
  mo = cib.ModelOptions(seasons=[cib.Seasons(num_seasons=7)]) if name == "seasonal_path" else None
  res = cib.fit_causalimpact(df, (0, 279), (280, 399), seed=(3, 4), model_options=mo,
                             inference_options=cib.InferenceOptions(num_results=500),
                             engine_options=cib.EngineOptions(num_chains=50))
  vals = [c for c in res.series.columns if not c.endswith(("_start", "_end"))]
  out[name] = dict(series=res.series[vals].values, summary=res.summary.values,
                   level=np.asarray(res.posterior_samples.level),
                   weights=np.asarray(res.posterior_samples.weights),
                   seasonal=np.asarray(res.posterior_samples.seasonal_levels))
  if int(os.environ.get("RANK", "0")) == 0:
    print(name, res.diagnostics["sampler"], "abs_effect", float(res.summary.loc["average", "abs_effect"]))
if int(os.environ.get("RANK", "0")) == 0:
  pickle.dump(out, open(sys.argv[1], "wb"))
if world > 1:
  dist.destroy_process_group()
