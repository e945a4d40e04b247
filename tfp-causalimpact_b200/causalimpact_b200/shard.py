"""Multi-GPU sharding of chains / draws (one process per GPU, torch.distributed).

Chains and posterior draws are independent, and every random number is keyed
by a GLOBAL chain / draw id, so the work is split into contiguous id ranges
with no data-path collective; the only exchange is ONE all-gather of the
per-draw result rows at the end.  The reference has no
distributed counterpart (single process, single chain).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def world() -> Tuple[int, int]:
  """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
  try:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
      return dist.get_rank(), dist.get_world_size()
  except ImportError:
    pass
  return 0, 1


def broadcast_u64(value: int, device=None) -> int:
  """Rank 0's 64-bit value on every rank (one tiny broadcast; identity without a group).
  Used for ``seed=None``: every rank of a sharded fit must key its Philox streams with the
  SAME fresh seed."""
  rank, ws = world()
  if ws == 1:
    return int(value)
  import torch
  import torch.distributed as dist
  if dist.get_backend() == "nccl":
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
  else:
    dev = torch.device("cpu")
  v = int(value) & (2**64 - 1)
  t = torch.tensor([v >> 32, v & 0xFFFFFFFF], dtype=torch.int64, device=dev)
  dist.broadcast(t, src=0)
  hi, lo = (int(x) for x in t.tolist())
  return (hi << 32) | lo


def split_range(n: int, world_size: int, rank: int) -> Tuple[int, int]:
  """Contiguous balanced split of range(n): returns (start, count)."""
  base, extra = divmod(n, world_size)
  start = rank * base + min(rank, extra)
  return start, base + (1 if rank < extra else 0)


def all_gather_rows(local, n_items: int, rows_per_item: int = 1):
  """Concatenate every rank's rows (axis 0) in rank order with ONE all-gather.

  ``local`` is this rank's [n_local_items * rows_per_item, width] tensor (torch; on the
  rank's GPU in production, on the CPU under the gloo tests), where the items (chains)
  are split by ``split_range(n_items, world, rank)``.  NCCL over NVLink when the backend
  is nccl -- the rows never leave the devices -- gloo (CPU tensors) otherwise.  Returns a
  tensor on ``local``'s device.
  """
  rank, ws = world()
  if ws == 1:
    return local
  import torch
  import torch.distributed as dist
  width = local.shape[1]
  counts = [split_range(n_items, ws, r)[1] * rows_per_item for r in range(ws)]
  cap = max(counts)
  on_gpu = dist.get_backend() == "nccl"
  home = local.device
  if on_gpu and not local.is_cuda:
    raise ValueError("all_gather_rows over NCCL needs the rows on this rank's GPU")
  # stage on the device the rows live on (the engine's), NOT torch.cuda.current_device(): a
  # caller that never ran torch.cuda.set_device(LOCAL_RANK) would put every rank on cuda:0
  dev = home if on_gpu else torch.device("cpu")
  send = torch.zeros((cap, width), dtype=local.dtype, device=dev)
  send[:local.shape[0]] = local.to(dev)
  recv = torch.empty((ws * cap, width), dtype=send.dtype, device=dev)
  dist.all_gather_into_tensor(recv, send)
  recv = recv.reshape(ws, cap, width)
  return torch.cat([recv[r, :counts[r]] for r in range(ws)], dim=0).to(home)
