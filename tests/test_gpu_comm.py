"""ci_comm_* / ci_allgather (include/ci_b200.h): the path's one collective through the C ABI alone
(NCCL resolved with dlopen, no torch.distributed).  One rank always; N ranks (one process per
GPU, the id handed over through a file) when the box has more than one GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, time
import numpy as np
root, path, rank, n = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tfp-causalimpact_b200"))
import torch
import causalimpact_b200 as cib
torch.cuda.set_device(rank)
eng = cib.Engine(rank)
idf = path + ".id"
if rank == 0:
  uid = cib.comm_unique_id()
  open(idf + ".tmp", "wb").write(uid); os.replace(idf + ".tmp", idf)
else:
  t0 = time.time()
  while not os.path.exists(idf):
    if time.time() - t0 > 120: raise SystemExit("no id")
    time.sleep(0.05)
  uid = open(idf, "rb").read()
comm = cib.Comm(eng, uid, rank, n)
local = torch.full((5, 7), float(rank), device=f"cuda:{rank}") + torch.arange(7, device=f"cuda:{rank}")
out = comm.allgather_rows(local)
torch.cuda.synchronize()
want = torch.cat([torch.full((5, 7), float(r)) + torch.arange(7) for r in range(n)])
assert torch.equal(out.cpu(), want), out
comm.close(); eng.close()
open(path + f".ok{rank}", "w").write("ok")
'''


def test_single_rank_comm(engine):
  import torch
  import causalimpact_b200 as cib
  comm = cib.Comm(engine, cib.comm_unique_id(), 0, 1)
  x = torch.arange(24, dtype=torch.float32, device="cuda:0").reshape(6, 4)
  y = comm.allgather_rows(x)
  torch.cuda.synchronize()
  assert torch.equal(x, y)
  comm.close()
  with pytest.raises(ValueError):
    cib.Comm(engine, b"short", 0, 1)


def test_multi_rank_comm_one_process_per_gpu(tmp_path):
  import torch
  n = min(torch.cuda.device_count(), 4)
  if n < 2:
    pytest.skip("one GPU on this box: the multi-rank all-gather needs >= 2 (run under gpurun --gpus 2)")
  script = str(tmp_path / "w.py")
  open(script, "w").write(WORKER)
  base = str(tmp_path / "comm")
  procs = [subprocess.Popen([sys.executable, script, ROOT, base, str(r), str(n)],
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
           for r in range(n)]
  outs = [p.communicate(timeout=300)[0] for p in procs]
  for r, (p, o) in enumerate(zip(procs, outs)):
    assert p.returncode == 0, (r, o[-2000:])
    assert os.path.exists(base + f".ok{r}")


def test_fused_sharded_impact_single_rank_equals_ci_impact(engine):
  """ci_impact_sharded_d with a 1-rank communicator (the exchange is a self send / receive): rows
  kernel, grouped ncclSend/ncclRecv, block merge, mean combine, column jobs and the all-reduce,
  all on one stream -- bit-identical to ci_impact_d."""
  import torch
  import causalimpact_b200 as cib
  from test_gpu_impact import make_meta
  rng = np.random.default_rng(4)
  comm = cib.Comm(engine, cib.comm_unique_id(), 0, 1)
  for dtype, S, T, t_pre, t0, t1 in [(np.float32, 1000, 517, 300, 310, 500), (np.float64, 97, 64, 40, 40, 64),
                                     (np.float32, 3, 40, 37, 37, 40)]:
    m = make_meta(T, t_pre, t0, t1, rng, nan_obs=2)
    traj = rng.normal(size=(S, T)).astype(dtype)
    mean = traj.mean(axis=0).astype(dtype)
    traj_d, mean_d = torch.from_numpy(traj).cuda(), torch.from_numpy(mean).cuda()
    want = engine.impact(traj_d, mean_d, m)
    out, full = comm.impact_sharded_t(traj_d, mean_d, m, [S])
    got = out.cpu().numpy()
    np.testing.assert_array_equal(full.cpu().numpy(), mean)
    np.testing.assert_array_equal(got[:T * 9].reshape(T, 9), want[0])
    np.testing.assert_array_equal(got[T * 9:], want[1])
  comm.close()


SHARDED_WORKER = r'''
import os, sys
import numpy as np
root = sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tfp-causalimpact_b200")); sys.path.insert(0, os.path.join(root, "tests"))
import torch, torch.distributed as dist
import causalimpact_b200 as cib
from causalimpact_b200 import shard
from test_gpu_impact import make_meta
rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
eng = cib.Engine(rank)
rng = np.random.default_rng(11)
for dtype, S, T, t_pre, t0, t1 in [(np.float32, 1003, 517, 300, 310, 500), (np.float64, 97, 64, 40, 40, 64),
                                   (np.float32, ws + 1, 40, 37, 37, 40), (np.float32, 1, 40, 30, 30, 40)]:
  m = make_meta(T, t_pre, t0, t1, rng, nan_obs=2)
  traj = rng.normal(size=(S, T)).astype(dtype)
  traj[::3] = np.round(traj[::3], 1)
  counts = shard.even_counts(S, ws)
  s0 = sum(counts[:rank])
  mine = torch.from_numpy(traj[s0:s0 + counts[rank]]).cuda()
  part = mine.double().mean(0).to(mine.dtype) if counts[rank] else torch.zeros(T, dtype=mine.dtype, device="cuda")
  sm = shard.ShardedMean(eng, part, counts)
  fused = shard.impact_sharded(eng, mine, sm, m, counts).cpu().numpy()            # ci_impact_sharded_d
  composed = shard.impact_sharded(eng, mine, sm.tensor.clone(), m, counts).cpu().numpy()   # torch collectives
  full_mean = sm.tensor
  want = eng.impact(torch.from_numpy(traj).cuda(), full_mean, m)
  for got in (fused, composed):
    np.testing.assert_array_equal(got[:T * 9].reshape(T, 9), want[0])
    np.testing.assert_array_equal(got[T * 9:], want[1])
  np.testing.assert_allclose(full_mean.cpu().numpy(), traj.mean(0), rtol=1e-5, atol=1e-6)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
  open(sys.argv[2], "w").write("ok")
'''


def test_fused_sharded_impact_across_gpus(tmp_path):
  """N >= 2 GPUs (skipped on a one-GPU box): ci_impact_sharded_d and the torch.distributed
  composition of ci_impact_rows_d / ci_impact_cols_d both equal ci_impact_d on the gathered draws
  bit for bit (given the same combined mean), with ragged draw counts and time blocks."""
  import torch
  n = min(torch.cuda.device_count(), 4)
  if n < 2:
    pytest.skip("needs >= 2 GPUs")
  script = tmp_path / "w.py"
  script.write_text(SHARDED_WORKER)
  ok = tmp_path / "ok"
  res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29533",
                        str(script), ROOT, str(ok)], capture_output=True, text=True, timeout=600)
  assert res.returncode == 0 and ok.exists(), res.stdout[-2000:] + res.stderr[-4000:]
