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
  dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
  send = torch.zeros((cap, width), dtype=local.dtype, device=dev)
  send[:local.shape[0]] = local.to(dev)
  recv = torch.empty((ws * cap, width), dtype=send.dtype, device=dev)
  dist.all_gather_into_tensor(recv, send)
  recv = recv.reshape(ws, cap, width)
  return torch.cat([recv[r, :counts[r]] for r in range(ws)], dim=0).to(home)
