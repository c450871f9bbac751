"""Functional + throughput run of the BASELINE.json configurations (configs[1..4]) for a few steps each, including a
batched evaluation and the NCCL return reduction. Launch with torch.distributed.run for more than one GPU.
  python scripts/run_configs.py [replicas_per_gpu_scale]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import il_b200
from il_b200 import distributed
from il_b200.config import load_config
from il_b200.train import Trainer

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
rank, world = distributed.init('nccl')
CONFIGS = [
  ('GAIL hopper 1024/GPU', ['algorithm=GAIL', 'env=hopper'], 1024),
  ('GMMIL halfcheetah 1024/GPU', ['algorithm=GMMIL', 'env=halfcheetah'], 1024),
  ('SAC ant 512/GPU (4096 over 8)', ['algorithm=SAC', 'env=ant'], 512),
  ('GAIL walker2d GP+SN 1024/GPU (8192 over 8)', ['algorithm=GAIL', 'env=walker2d', 'imitation.grad_penalty=1', 'imitation.spectral_norm=true'], 1024),
  ('PWIL hopper 256/GPU', ['algorithm=PWIL', 'env=hopper'], 256),
]
for name, ov, R in CONFIGS:
  R = max(int(R * scale), 2)
  start, K = 40, 12
  cfg = load_config(ov + [f'steps={start + 3 * K + 16}', f'training.start={start}', 'imitation.trajectories=5', f'replicas={R}', 'evaluation.episodes=4', 'seed=0', 'memory.size=4096'])
  lo, hi = distributed.shard(R * world, rank, world)
  tr = Trainer(cfg, replicas=R, seed_offset=lo, fast_init=True)
  for _ in range(start + K): tr.train_step()
  torch.cuda.synchronize(); distributed.barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(K): tr.train_step()
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / K
  # shortened evaluation episodes keep this script quick
  tr.eval_env._max_episode_steps = 100
  t0 = time.time()
  returns = tr.evaluate()
  mean, std, n = distributed.return_statistics(returns)
  finite = bool(torch.isfinite(tr.actor.mlp.flat).all() and torch.isfinite(tr.critic.mlp.flat).all() and torch.isfinite(tr.log_alpha).all())
  if rank == 0:
    print(json.dumps(dict(config=name, replicas_per_gpu=R, gpus=world, ms_per_step=round(ms, 3), env_steps_per_s=round(R * world / ms * 1e3), eval_mean_return=round(mean, 3),
                          eval_std=round(std, 3), eval_episodes=n, eval_seconds=round(time.time() - t0, 2), params_finite=finite, sac_losses=[round(float(x), 4) for x in tr.sac_out['losses'][0]])),
          flush=True)
  assert finite
  del tr
  torch.cuda.empty_cache()
distributed.barrier()
if torch.distributed.is_initialized(): torch.distributed.destroy_process_group()
