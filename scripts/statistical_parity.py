"""Long-horizon, statistical parity (SURVEY.md §7 "chaotic divergence" level iii): the same GAIL-hopper configuration is
run (a) as independent CPU oracle loops (oracle/loop.py, the reference's own RNG calls) and (b) as one replica-batched GPU
trainer with device RNG, and summary statistics of the learning signals are compared across replicas at checkpoints.
Individual trajectories diverge (different random streams); distributions over replicas must agree.
  python scripts/statistical_parity.py [steps] [cpu_replicas] [gpu_replicas]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
n_cpu = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n_gpu = int(sys.argv[3]) if len(sys.argv) > 3 else 256
start, B, H = 300, 128, 256
checkpoints = [c for c in (500, 1000, 1500, 2000, 3000) if c <= steps]


def cpu_worker(seed):
  torch.set_num_threads(1)
  from oracle import loop
  lp = loop.OracleLoop('GAIL', 'hopper', seed=seed, batch_size=B, start=start, memory_size=steps, hidden_size=H, trajectories=5)
  out = {}
  for s in range(1, steps + 1):
    lp.run_step()
    if s in checkpoints:
      out[s] = dict(alpha=float(lp.agent.log_alpha.exp()), q=float(lp.last['sac']['q_values'].mean()), reward=float(lp.last['rewards'].mean()),
                    entropy=float(-lp.last['sac']['log_probs'].mean()), episodes=len(lp.episode_returns), mean_return=float(np.mean(lp.episode_returns)) if lp.episode_returns else float('nan'))
  return out


if __name__ == '__main__':
  import torch.multiprocessing as mp
  t0 = time.time()
  with mp.get_context('spawn').Pool(min(n_cpu, os.cpu_count() or 1)) as pool:
    cpu = pool.map(cpu_worker, range(n_cpu))
  t_cpu = time.time() - t0
  import il_b200
  from il_b200.config import load_config
  from il_b200.train import Trainer
  cfg = load_config(['algorithm=GAIL', 'env=hopper', f'steps={steps}', f'training.start={start}', f'training.batch_size={B}', 'imitation.trajectories=5', f'replicas={n_gpu}', 'seed=100'])
  t0 = time.time()
  tr = Trainer(cfg)
  gpu = {}
  for s in range(1, steps + 1):
    tr.train_step()
    if s in checkpoints:
      eps = tr.episodes.float().clamp_min(1)
      gpu[s] = dict(alpha=tr.log_alpha.exp().cpu().numpy(), q=tr.sac_out['q_values'].mean(1).cpu().numpy(), reward=tr.batch['rewards'].mean(1).cpu().numpy(),
                    entropy=(-tr.sac_out['log_probs']).mean(1).cpu().numpy(), episodes=tr.episodes.cpu().numpy(), mean_return=(tr.return_sum / eps).cpu().numpy())
  t_gpu = time.time() - t0
  rows = []
  for s in checkpoints:
    row = dict(step=s)
    for k in ('alpha', 'q', 'reward', 'entropy', 'episodes', 'mean_return'):
      c = np.array([r[s][k] for r in cpu], dtype=np.float64)
      g = np.asarray(gpu[s][k], dtype=np.float64)
      row[k] = dict(cpu_mean=round(float(np.nanmean(c)), 4), cpu_std=round(float(np.nanstd(c)), 4), gpu_mean=round(float(np.nanmean(g)), 4), gpu_std=round(float(np.nanstd(g)), 4))
    rows.append(row)
    print(json.dumps(row), flush=True)
  print(json.dumps(dict(cpu_replicas=n_cpu, gpu_replicas=n_gpu, steps=steps, cpu_seconds=round(t_cpu, 1), gpu_seconds=round(t_gpu, 1))), flush=True)
