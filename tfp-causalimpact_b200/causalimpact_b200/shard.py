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


def _staging_device(t):
  import torch
  import torch.distributed as dist
  return t.device if dist.get_backend() == "nccl" else torch.device("cpu")


def even_counts(n_total: int, world_size: int):
  """Draws per rank under split_range."""
  return [split_range(n_total, world_size, r)[1] for r in range(world_size)]


class ShardedDraws:
  """[S_total, T] draws that stay sharded over the ranks: ``local`` holds this rank's rows
  (``counts[rank]`` of them, in rank order of the global row ids).  What a sharded fit hands to
  the impact stage instead of a gathered array; ``np.asarray`` gathers on demand."""

  def __init__(self, local, counts):
    self.local, self.counts = local, [int(c) for c in counts]
    self.shape = (sum(self.counts), int(local.shape[1]))

  def gathered(self):
    import torch
    import torch.distributed as dist
    rank, ws = world()
    cap = max(self.counts)
    dev = _staging_device(self.local)
    send = torch.zeros((cap, self.shape[1]), dtype=self.local.dtype, device=dev)
    send[:self.counts[rank]] = self.local[:self.counts[rank]].to(dev)
    recv = torch.empty((ws, cap, self.shape[1]), dtype=send.dtype, device=dev)
    dist.all_gather_into_tensor(recv.view(ws * cap, -1), send)
    return torch.cat([recv[r, :self.counts[r]] for r in range(ws)], dim=0).to(self.local.device)

  def __array__(self, dtype=None, copy=None):
    a = self.gathered().cpu().numpy()
    return a if dtype is None else a.astype(dtype, copy=False)


class ShardedMean:
  """The predictive mean of a fit whose draws stay sharded: ``part`` is the mean over THIS rank's
  draws; the mean over all draws (``.tensor``, ``np.asarray``) is combined on first use -- the
  fused impact stage (ci_impact_sharded_d) combines the parts inside its own exchange and never
  asks."""

  def __init__(self, eng, part, counts):
    self._eng, self.part, self.counts = eng, part.reshape(-1), [int(c) for c in counts]
    self.shape = (int(self.part.shape[0]),)
    self._full = None

  def set_full(self, full):
    self._full = full

  @property
  def tensor(self):
    if self._full is None:
      self._full = _combine_means(self.part, self.counts)
    return self._full

  def numpy(self):
    return self.tensor.detach().cpu().numpy()

  def __array__(self, dtype=None, copy=None):
    a = self.numpy()
    return a if dtype is None else a.astype(dtype, copy=False)


def engine_comm(eng):
  """The engine's own NCCL communicator over the ranks of the torch.distributed group (ci_comm;
  created once per engine: rank 0's unique id is broadcast through the group).  None when the
  group is not NCCL or the engine is not the CUDA one (the CPU tests' stand-in)."""
  rank, ws = world()
  if ws == 1 or not hasattr(eng, "_lib"):
    return None
  import torch
  import torch.distributed as dist
  if dist.get_backend() != "nccl":
    return None
  comm = getattr(eng, "_shard_comm", None)
  if comm is None:
    from . import _engine
    dev = eng.torch_device()
    uid = torch.zeros(_engine.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
      uid.copy_(torch.frombuffer(bytearray(_engine.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, src=0)
    comm = _engine.Comm(eng, bytes(uid.cpu().numpy().tobytes()), rank, ws)
    eng._shard_comm = comm
  return comm


_WEIGHTS = {}


def _combine_means(part, counts):
  """[T] mean over this rank's draws -> mean over all draws: one [T] all-gather and a
  draw-count-weighted float64 sum in rank order (the same on every rank)."""
  import torch
  import torch.distributed as dist
  rank, ws = world()
  if ws == 1:
    return part
  T = part.shape[0]
  home = part.device
  dev = _staging_device(part)
  part = part.to(dev).contiguous()
  flat = torch.empty(ws * T, dtype=part.dtype, device=dev)
  dist.all_gather_into_tensor(flat, part)
  key = (tuple(counts), str(dev))
  wts = _WEIGHTS.get(key)
  if wts is None:          # built once per shard layout: torch.tensor(..., device=) is a blocking copy
    n_total = float(sum(counts))
    wts = _WEIGHTS[key] = torch.tensor([c / n_total for c in counts], dtype=torch.float64,
                                       device=dev)
  return (wts @ flat.view(ws, T).to(torch.float64)).to(part.dtype).to(home)


def predictive_mean_part(eng, theta_local, level_local, counts):
  """ci_predictive_mean_d over this rank's ``counts[rank]`` draws (zeros without draws)."""
  import torch
  rank, _ = world()
  if counts[rank]:
    return eng.predictive_mean_t(theta_local[:counts[rank]], level_local[:counts[rank]]).reshape(-1)
  return torch.zeros(level_local.shape[1], dtype=level_local.dtype, device=level_local.device)


def predictive_mean_sharded(eng, theta_local, level_local, counts):
  """The predictive mean over ALL draws from each rank's mean of its own ``counts[rank]`` draws
  (equal to the one-GPU mean up to the rounding of the partial means to the draw dtype)."""
  rank, ws = world()
  if ws == 1:
    return eng.predictive_mean_t(theta_local, level_local)
  return _combine_means(predictive_mean_part(eng, theta_local, level_local, counts), counts)


def _exchange_columns(local, splits, counts, rank, head_rows: int = 0):
  """All-to-all by time block.  ``local`` [head_rows + n_cols, S_local] holds this rank's draws
  of every column (rows contiguous); ``splits[g]`` = (start, count) of the columns rank g owns;
  the ``head_rows`` leading rows all go to rank 0.  Returns ([n_mine, S_total], head
  [head_rows, S_total] or None): the received blocks laid side by side in rank order."""
  import torch
  import torch.distributed as dist
  ws = len(counts)
  s_loc = counts[rank]
  in_splits = [(splits[g][1] + (head_rows if g == 0 else 0)) * s_loc for g in range(ws)]
  rows_me = splits[rank][1] + (head_rows if rank == 0 else 0)
  out_splits = [rows_me * counts[r] for r in range(ws)]
  dev = _staging_device(local)
  send = local.to(dev).contiguous().reshape(-1)
  recv = torch.empty(sum(out_splits), dtype=local.dtype, device=dev)
  dist.all_to_all_single(recv, send, out_splits, in_splits)
  blocks = [b.view(rows_me, counts[r]) for r, b in enumerate(recv.split(out_splits))]
  merged = torch.cat(blocks, dim=1).to(local.device)          # [rows_me, S_total]
  if rank == 0 and head_rows:
    return merged[head_rows:], merged[:head_rows]
  return merged, None


def impact_sharded(eng, traj_local, mean, meta, counts, trace=None):
  """ci_impact over draws sharded ``counts[r]`` per rank WITHOUT gathering them: each rank
  transposes its own paths (ci_impact_rows_d), the ranks swap time blocks (two all-to-alls: the
  float paths, and the float64 cumulative paths with the per-draw statistics riding to rank 0),
  each rank selects the quantiles of its T/world time steps over all draws (ci_impact_cols_d),
  and one all-reduce of the [T*9 + 20] result -- every entry written by exactly one rank, zeros
  elsewhere -- leaves the full series + summary on every rank.  ``mean``: a ShardedMean (on
  GPUs with an NCCL group the whole stage then runs as ONE library call, ci_impact_sharded_d,
  which also combines the mean parts inside its exchange) or the predictive mean over all draws
  as a tensor (the torch.distributed composition below: what the gloo CPU tests exercise).  Returns the float64 device tensor [T*9 + 20]; equal to
  ``eng.impact`` on the gathered draws (the quantile selection is exact).  ``trace(name)`` is
  called at every step boundary (tools/prof_sharded.py records events there)."""
  trace = trace or (lambda name: None)
  rank, ws = world()
  import torch
  import torch.distributed as dist
  T = traj_local.shape[1]
  counts = [int(c) for c in counts]
  comm = engine_comm(eng)
  if comm is not None and isinstance(mean, ShardedMean):
    # the product path on GPUs: the whole stage, exchange included, is ONE call into the library
    out, full = comm.impact_sharded_t(traj_local[:counts[rank]], mean.part, meta, counts)
    mean.set_full(full)
    trace("fused")
    return out
  if isinstance(mean, ShardedMean):
    mean = mean.tensor
  out = torch.zeros(T * 9 + 20, dtype=torch.float64, device=traj_local.device)
  if ws == 1:
    eng.impact(traj_local, mean, meta, out=out)
    return out
  if counts[0] < 1:
    raise ValueError("impact_sharded: rank 0 must hold at least one draw")
  per = np.asarray(meta.period)
  t_c0 = int(np.argmax(per != 0)) if np.any(per != 0) else T
  tsp = [split_range(T, ws, r) for r in range(ws)]
  csp = [split_range(T - t_c0, ws, r) for r in range(ws)]
  if counts[rank]:
    trT, _, _, packed = eng.impact_rows_t(traj_local[:counts[rank]], mean if rank == 0 else None,
                                          meta, out)
  else:                                   # more ranks than draws: nothing to contribute
    trT = torch.empty((T, 0), dtype=traj_local.dtype, device=traj_local.device)
    packed = torch.empty((5 + T - t_c0, 0), dtype=torch.float64, device=traj_local.device)
  trace("rows")
  tr_all, _ = _exchange_columns(trT, tsp, counts, rank)
  trace("alltoall_paths")
  cum_all, stats_all = _exchange_columns(packed, csp, counts, rank, head_rows=5)
  trace("alltoall_cumulative")
  eng.impact_cols_t(tr_all, tsp[rank][0], cum_all, csp[rank][0], stats_all, meta, out)
  trace("columns")
  red = out.to(_staging_device(out))
  dist.all_reduce(red)
  trace("allreduce")
  return red.to(out.device)
