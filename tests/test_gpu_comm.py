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
