"""evaluate_agent (reference evaluation.py:11-35): the reference runs `num_episodes` greedy episodes one after
another with batch size 1; here all replicas x episodes advance in lock-step on the device with finished episodes
frozen, which is the same computation because episodes are independent given their initial states. The whole
rollout is ONE C-ABI call (il_eval_rollout): a CUDA graph whose loop body (greedy actor forward, env step + return
accumulation) sits in a WHILE conditional node with the condition set on the device — the host launches once and
does not synchronise until the caller reads the returns."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple, Union

import torch
from torch import Tensor

from . import _lib
from .environments import D4RLEnv
from .models import SoftActor


class _EvalState:
  """Device buffers of one (actor, env, episodes) evaluation set-up; cached on the env so repeated evaluations (train.py:213-219)
  reuse the same buffers — and therefore the same cached device graph."""

  def __init__(self, actor: SoftActor, env: D4RLEnv, E: int, trajectories: bool):
    R, dev = actor.replicas, actor.device
    self.eb = eb = env.eval_batch(E)
    self.n, self.S, self.A = n, S, A = R * E, eb.S, actor.action_size
    self.state, self.returns = torch.empty(n, S, device=dev), torch.zeros(n, device=dev)
    self.counters = torch.zeros(2, dtype=torch.int64, device=dev)
    T = eb.max_episode_steps
    self.traj = None
    if trajectories:
      need = n * T * (S + A + 1) * 4
      if need > 8 << 30: raise MemoryError(f'return_trajectories for {n} episodes of up to {T} steps needs {need / 2**30:.1f} GiB; evaluate fewer replicas / episodes at a time')
      self.traj = dict(states=torch.zeros(n, T, S, device=dev), actions=torch.zeros(n, T, A, device=dev), rewards=torch.zeros(n, T, device=dev), len=torch.zeros(n, dtype=torch.int32, device=dev))
    a = self.args = _lib.EvalArgs()
    a.actor, a.env, a.R, a.episodes, a.max_steps, a.traj_T = actor.mlp.c_struct(), eb.c_struct(), R, E, T + 1, (T if trajectories else 0)
    a.state, a.returns, a.out_counters = self.state.data_ptr(), self.returns.data_ptr(), self.counters.data_ptr()
    if trajectories:
      a.traj_states, a.traj_actions, a.traj_rewards, a.traj_len = (self.traj[k].data_ptr() for k in ('states', 'actions', 'rewards', 'len'))
    need = _lib.lib().il_eval_workspace_bytes(C.byref(a))
    assert need > 0
    self.ws = torch.zeros(need, dtype=torch.uint8, device=dev)
    a.workspace, a.workspace_bytes = self.ws.data_ptr(), self.ws.numel()


def evaluate_agent(actor: SoftActor, env: D4RLEnv, num_episodes: int, return_trajectories: bool = False, render: bool = False, reset_noise: Optional[Tensor] = None,
                   out_stats: Optional[Dict[str, int]] = None) -> Union[List[float], Tensor, Tuple]:
  """Returns the list of episode returns (R == 1, like the reference) or a [R, num_episodes] tensor; with
  `return_trajectories` also the per-episode dicts of evaluation.py:30-33 (`states`, `actions`, `rewards`, `terminals`; a list
  of `num_episodes` dicts for R == 1, a list of R such lists otherwise).
  `reset_noise` ([R * num_episodes, obs] U[0,1) draws) injects the initial states; default: the env's seeded stream.
  `out_stats`, when given, receives `iterations` and `env_steps` of the rollout (a host read: synchronises)."""
  R, E, dev = actor.replicas, num_episodes, actor.device
  assert env.replicas == R
  key = (id(actor), E, bool(return_trajectories))
  cache = env.__dict__.setdefault('_eval_states', {})
  es = cache.get(key)
  if es is None or es.args.actor.params != actor.mlp.flat.data_ptr():
    es = cache[key] = _EvalState(actor, env, E, return_trajectories)
  eb, n = es.eb, es.n
  u = env.reset_noise(n) if reset_noise is None else torch.as_tensor(reset_noise, dtype=torch.float32).to(dev).reshape(n, eb.obs).contiguous()
  eb.reset(u, es.state)  # evaluation.py:19
  _lib.check(_lib.lib().il_eval_rollout(_lib.handle(), C.byref(es.args), _lib.stream()))  # evaluation.py:20-28 for every episode
  if out_stats is not None:
    c = es.counters.cpu()
    out_stats['iterations'], out_stats['env_steps'] = int(c[0]), int(c[1])
  returns = es.returns.clone()
  out = [float(x) for x in returns.cpu()] if R == 1 else returns.view(R, E)
  if not return_trajectories: return out
  lens = es.traj['len'].cpu().tolist()
  st, ac, rw = es.traj['states'].cpu(), es.traj['actions'].cpu(), es.traj['rewards'].cpu()
  eps = []
  for i in range(n):  # evaluation.py:30-33
    L = lens[i]
    eps.append({'states': st[i, :L].clone(), 'actions': ac[i, :L].clone(), 'rewards': rw[i, :L].clone(), 'terminals': torch.cat([torch.zeros(L - 1), torch.ones(1)])})
  return out, (eps if R == 1 else [eps[r * E:(r + 1) * E] for r in range(R)])
