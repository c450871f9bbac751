"""Seed-sharded multi-GPU execution: replicas are independent, so the replica axis is split
contiguously across ranks and there is no data-path collective. The only exchange is the evaluation-return
reduction: a 3-float (sum, sum of squares, count) vector per rank, produced on the device by il_return_stats
and summed with one NCCL all-reduce on the same stream."""
from __future__ import annotations

import ctypes as C
import os
from typing import Tuple

import torch
import torch.distributed as dist

from . import _lib


def init(backend: str = 'nccl') -> Tuple[int, int]:
  """One process per GPU; reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* set by torch.distributed.run."""
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  if backend == 'nccl': torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
  if world > 1 and not dist.is_initialized():
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29512')
    kw = dict(device_id=torch.device('cuda', torch.cuda.current_device())) if backend == 'nccl' else {}
    dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
  return rank, world


def shard(total_replicas: int, rank: int, world: int) -> Tuple[int, int]:
  """Contiguous replica range [lo, hi) owned by `rank` (GPU g owns replicas [g R/G, (g+1) R/G))."""
  base, rem = divmod(total_replicas, world)
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


_nccl_comm = None


def nccl_comm():
  """This process's ncclComm_t for the library's own collectives (il_return_allreduce): created once from an ncclUniqueId
  that rank 0 generates and torch.distributed broadcasts. None for a single process."""
  global _nccl_comm
  if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1): return None
  if _nccl_comm is None:
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = (C.c_uint8 * 128)()
    if rank == 0: _lib.check(_lib.lib().il_nccl_unique_id(uid))
    t = torch.tensor(list(uid), dtype=torch.uint8, device='cuda' if dist.get_backend() == 'nccl' else 'cpu')
    dist.broadcast(t, src=0)
    uid = (C.c_uint8 * 128)(*t.cpu().tolist())
    comm = C.c_void_p()
    _lib.check(_lib.lib().il_nccl_comm_create(uid, rank, world, C.byref(comm)))
    _nccl_comm = comm
  return _nccl_comm


def reduce_stats(stats3: torch.Tensor) -> torch.Tensor:
  """Sum of per-rank (sum, sum of squares, count) vectors; in place. No-op for a single process."""
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1: dist.all_reduce(stats3, op=dist.ReduceOp.SUM)
  return stats3


def return_stats_device(returns: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
  """(sum, sum of squares, count) of the evaluation returns over ALL ranks, left on the device: il_return_allreduce = the
  per-rank reduction kernel + ncclAllReduce enqueued on the same stream (no host involvement)."""
  out = torch.empty(3, device=returns.device, dtype=torch.float32) if out is None else out
  flat = returns.reshape(-1).contiguous()
  _lib.check(_lib.lib().il_return_allreduce(_lib.handle(), nccl_comm(), flat.data_ptr(), flat.numel(), out.data_ptr(), _lib.stream()))
  return out


def return_statistics(returns: torch.Tensor) -> Tuple[float, float, int]:
  """Global mean / std / count of evaluation returns over all ranks (device reduction + one all-reduce)."""
  out = return_stats_device(returns)
  s, s2, n = (float(x) for x in out.cpu())
  mean = s / n
  return mean, max(s2 / n - mean * mean, 0.0) ** 0.5, int(n)


def stats_from_sums(stats3: torch.Tensor) -> Tuple[float, float, int]:
  s, s2, n = (float(x) for x in stats3.cpu())
  mean = s / max(n, 1.0)
  return mean, max(s2 / max(n, 1.0) - mean * mean, 0.0) ** 0.5, int(n)


def barrier():
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1: dist.barrier()


def shutdown():
  """Destroys the library's NCCL communicator and the torch process group (end of train.py / bench.py)."""
  global _nccl_comm
  if _nccl_comm is not None:
    _lib.check(_lib.lib().il_nccl_comm_destroy(_nccl_comm))
    _nccl_comm = None
  if dist.is_available() and dist.is_initialized():
    dist.barrier()
    dist.destroy_process_group()
