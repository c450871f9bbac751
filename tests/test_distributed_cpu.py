"""Host-side logic of the multi-GPU path on CPU (gloo, world_size 2): replica sharding and the evaluation-return
reduction (the one collective of the design). The device-side producer of the (sum, sum^2, count) vector
is covered by the GPU tests; here the vector is built with torch on the CPU."""
import os

import pytest
import torch
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
  os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  import il_b200  # noqa: F401
  from il_b200 import distributed
  r, w = distributed.init('gloo')
  assert (r, w) == (rank, world)
  total = 11
  lo, hi = distributed.shard(total, rank, world)
  returns = torch.arange(total * 3, dtype=torch.float32).reshape(total, 3)[lo:hi]  # [local replicas, episodes]
  stats = torch.tensor([returns.sum(), (returns ** 2).sum(), returns.numel()], dtype=torch.float32)
  distributed.reduce_stats(stats)
  mean, std, n = distributed.stats_from_sums(stats)
  distributed.barrier()
  q.put((rank, lo, hi, mean, std, n))


def test_shard_partitions_all_replicas():
  from il_b200 import distributed
  for total in (1, 7, 1024, 8192):
    for world in (1, 2, 3, 8):
      ranges = [distributed.shard(total, r, world) for r in range(world)]
      assert ranges[0][0] == 0 and ranges[-1][1] == total
      assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
      sizes = [hi - lo for lo, hi in ranges]
      assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_return_reduction_two_ranks_gloo():
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = 29600 + (os.getpid() % 200)
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs: p.start()
  res = sorted(q.get(timeout=100) for _ in range(2))
  for p in procs: p.join(timeout=30)
  assert [r[1:3] for r in res] == [(0, 6), (6, 11)]
  full = torch.arange(33, dtype=torch.float32)
  for r in res:
    assert r[5] == 33
    assert abs(r[3] - full.mean().item()) < 1e-4
    assert abs(r[4] - full.std(unbiased=False).item()) < 1e-3
