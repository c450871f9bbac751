"""Multi-GPU checks (need >= 2 visible GPUs; skipped otherwise): the library's own NCCL communicator and the fused
return reduction (il_return_allreduce: per-rank reduction kernel + ncclAllReduce on the same stream)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, torch
sys.path.insert(0, os.environ['IL_ROOT'])
import il_b200
from il_b200 import distributed
rank, world = distributed.init('nccl')
lo, hi = distributed.shard(10, rank, world)
full = (torch.arange(10 * 30, dtype=torch.float32).reshape(10, 30) * 0.25 - 20)
mine = full[lo:hi].cuda()
mean, std, n = distributed.return_statistics(mine)           # il_return_allreduce over the library's own ncclComm_t
ref = torch.tensor([float(full.sum()), float((full * full).sum()), 300.0])
got = distributed.return_stats_device(mine).cpu()
assert n == 300 and abs(mean - float(full.mean())) < 1e-3 and abs(std - float(full.std(unbiased=False))) < 1e-2, (mean, std, n)
assert torch.allclose(got, ref, rtol=1e-5), (got, ref)
g = torch.cuda.CUDAGraph()                                   # the reduction + all-reduce is capturable: one graph node sequence, no host involvement
out = torch.zeros(3, device='cuda')
distributed.return_stats_device(mine, out)
torch.cuda.synchronize()
with torch.cuda.graph(g): distributed.return_stats_device(mine, out)
out.zero_(); g.replay(); torch.cuda.synchronize()
assert torch.allclose(out.cpu(), ref, rtol=1e-5), out
distributed.barrier()
print('RANK_OK', rank, flush=True)
"""


@pytest.mark.timeout(300)
def test_return_allreduce_over_nccl_two_ranks(tmp_path):
  if torch.cuda.device_count() < 2: pytest.skip('needs 2 GPUs')
  script = tmp_path / 'worker.py'
  script.write_text(WORKER)
  env = dict(os.environ, IL_ROOT=ROOT)
  r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1', '--master-port', '29731', str(script)],
                     env=env, capture_output=True, text=True, timeout=280)
  assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
  assert r.stdout.count('RANK_OK') == 2, r.stdout
