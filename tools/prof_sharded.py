"""Step-by-step device and host times of the time-sharded impact stage (shard.impact_sharded)
under torchrun:  python -m torch.distributed.run --nproc-per-node N tools/prof_sharded.py
Device time between boundaries from CUDA events, host time from perf_counter (a step whose host
time exceeds its device time is launch-bound)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tfp-causalimpact_b200"))
import torch
import torch.distributed as dist

import causalimpact_b200 as cib
from causalimpact_b200 import impact as _imp


def main():
  rank, world, local = (int(os.environ.get(k, d)) for k, d in
                        (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  S, T, p = 10000, 2000, 10
  rng = np.random.default_rng(5)
  X = rng.normal(size=(T, p)); y = X[:, 0] + np.cumsum(rng.normal(size=T)) * 0.05 + rng.normal(size=T) * 0.3
  spec = cib.build_problem(y, X)
  eng = cib.Engine(local)
  eng.set_data(spec)
  counts = cib.shard.even_counts(S, world)
  s0 = sum(counts[:rank])
  th = np.zeros((counts[rank], spec.dim), np.float32)
  th[:, :p] = 0.1 * rng.normal(size=(counts[rank], p)); th[:, p] = np.log(0.09); th[:, p + 1] = np.log(0.0025)
  th_d = torch.from_numpy(th).cuda()
  lvl, trj = eng.posterior_predict_t(th_d, seed=3, draw_id0=s0)
  t_pre = 1400
  per = np.zeros(T, np.uint8); per[t_pre:] = 1
  obs = rng.normal(size=T)
  meta = _imp.ImpactMeta(index=None, observed=obs, period=per, hide=None, scale=2.0, offset=100.0,
                         q_lo=0.025, q_hi=0.975, obs_mean=0.0, obs_sum=float(obs[t_pre:].sum()))
  names, evs, host = [], [], []

  def trace(name):
    e = torch.cuda.Event(enable_timing=True); e.record()
    names.append(name); evs.append(e); host.append(time.perf_counter())

  acc = {}
  reps = 20
  for it in range(reps + 5):
    del names[:], evs[:], host[:]
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    trace("start")
    if os.environ.get("PROF_TORCH_PATH"):      # the torch.distributed composition
      mean = cib.shard.predictive_mean_sharded(eng, th_d, lvl, counts)
    else:                                      # ci_impact_sharded_d
      mean = cib.shard.ShardedMean(eng, cib.shard.predictive_mean_part(eng, th_d, lvl, counts), counts)
    trace("mean")
    cib.shard.impact_sharded(eng, trj, mean, meta, counts, trace=trace)
    trace("end")
    torch.cuda.synchronize()
    if it >= 5:
      for i in range(1, len(names)):
        d = acc.setdefault(names[i], [0.0, 0.0])
        d[0] += evs[i - 1].elapsed_time(evs[i]); d[1] += (host[i] - host[i - 1]) * 1e3
  if rank == 0:
    print(f"world {world}: S={S} T={T}; per step: device ms / host ms")
    for k, (dv, hs) in acc.items():
      print(f"  {k:22s} {dv / reps:8.4f} {hs / reps:8.4f}")
    print(f"  {'total':22s} {sum(v[0] for v in acc.values()) / reps:8.4f} "
          f"{sum(v[1] for v in acc.values()) / reps:8.4f}")
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
