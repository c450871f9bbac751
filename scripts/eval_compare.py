"""evaluate_agent: the round-1 host loop (one il_actor_forward + il_env_step + il_eval_accumulate per step from Python, a host sync every 50 steps)
against the round-2 device program (il_eval_rollout: CUDA graph with a device-side WHILE node). Same actor, same initial states; prints JSON lines.
  python scripts/eval_compare.py [replicas] [episodes]"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import il_b200
from il_b200 import _lib
from il_b200.environments import D4RLEnv
from il_b200.evaluation import evaluate_agent


class Cfg(dict):
  __getattr__ = dict.__getitem__
  def get(self, k, d=None): return dict.get(self, k, d)


def host_loop(actor, env, E, u, check_every=50):
  """The round-1 implementation (imitation-learning_b200/evaluation.py at commit 6b4ce7f), kept here for the comparison."""
  R, dev = actor.replicas, actor.device
  eb = env.eval_batch(E)
  n, S, A = R * E, eb.S, actor.action_size
  state, nxt = torch.empty(n, S, device=dev), torch.empty(n, S, device=dev)
  reward, returns = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
  done, finished, running = torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
  eb.reset(u, state)
  lib, h, steps = _lib.lib(), _lib.handle(), 0
  for t in range(eb.max_episode_steps):
    action = actor._run(state.view(R, E, S), want=('action', ))['action'].view(n, A)
    eb.step(action, nxt, reward, done, frozen=finished)
    _lib.check(lib.il_eval_accumulate(h, n, reward.data_ptr(), done.data_ptr(), returns.data_ptr(), finished.data_ptr(), running.data_ptr(), _lib.stream()))
    state, nxt = nxt, state
    steps += 1
    if (t + 1) % check_every == 0 and int(running.item()) == 0: break
  return returns.view(R, E), steps


def main():
  R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
  E = int(sys.argv[2]) if len(sys.argv) > 2 else 30
  for Rr in sorted({1, 64, R}):
    env = D4RLEnv('hopper', True, replicas=Rr)
    actor = il_b200.SoftActor(12, 3, Cfg(hidden_size=256, depth=2, activation='relu'), replicas=1)
    actor.mlp.flat = actor.mlp.flat.expand(Rr, -1).contiguous(); actor.mlp.replicas = actor.replicas = Rr
    u = env.reset_noise(Rr * E)
    res = {}
    for name in ('host_loop_r1', 'device_program_r2'):
      for rep in range(2):  # second repetition is the timed one (graph build / allocator warm-up in the first)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if name == 'host_loop_r1':
          ret, iters = host_loop(actor, env, E, u)
          steps = None
        else:
          st = {}
          ret = evaluate_agent(actor, env, E, reset_noise=u, out_stats=st)
          ret = ret if torch.is_tensor(ret) else torch.tensor(ret, device='cuda').view(Rr, E)
          iters, steps = st['iterations'], st['env_steps']
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
      res[name] = dict(seconds=dt, iterations=iters, returns_mean=float(ret.mean()))
      if steps is not None: res[name]['env_steps'] = steps
    es = res['device_program_r2']['env_steps']
    print(json.dumps(dict(replicas=Rr, episodes=E, env_steps=es, host_loop_s=res['host_loop_r1']['seconds'], device_program_s=res['device_program_r2']['seconds'],
                          speedup=res['host_loop_r1']['seconds'] / res['device_program_r2']['seconds'], eval_steps_per_s=es / res['device_program_r2']['seconds'],
                          same_returns=abs(res['host_loop_r1']['returns_mean'] - res['device_program_r2']['returns_mean']) < 1e-3 * max(1.0, abs(res['host_loop_r1']['returns_mean'])))), flush=True)


if __name__ == '__main__':
  main()
